"""Host-side WordPiece tokenisation for BERT-class sentence encoders (uncased or cased
``vocab.txt`` models such as all-MiniLM-L6-v2 and bge-base-en-v1.5).

In the reference this happens inside ``SentenceTransformer.encode`` → HF ``BertTokenizer``
(``src/lean_explore/util/embedding_client.py:99``; truncation to ``max_seq_length``, ``:60-63``).
String processing stays on the host; the output (``input_ids`` / ``attention_mask`` int32, right
padded to the longest sequence of the batch) is what ``lxg_encode`` consumes.  The algorithm is
BERT's published one: BasicTokenizer (clean, optional lower-casing + accent stripping, CJK
isolation, punctuation splitting) followed by greedy longest-match-first WordPiece with the ``##``
continuation prefix; ``tests/test_tokenizer.py`` cross-checks it against HF's implementation.
"""

from __future__ import annotations

import unicodedata
from pathlib import Path

import numpy as np


def _is_whitespace(ch: str) -> bool:
    if ch in (" ", "\t", "\n", "\r"):
        return True
    return unicodedata.category(ch) == "Zs"


def _is_control(ch: str) -> bool:
    if ch in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(ch).startswith("C")


def _is_punctuation(ch: str) -> bool:
    cp = ord(ch)
    if (33 <= cp <= 47) or (58 <= cp <= 64) or (91 <= cp <= 96) or (123 <= cp <= 126):
        return True
    return unicodedata.category(ch).startswith("P")


def _is_cjk(cp: int) -> bool:
    return ((0x4E00 <= cp <= 0x9FFF) or (0x3400 <= cp <= 0x4DBF) or (0x20000 <= cp <= 0x2A6DF)
            or (0x2A700 <= cp <= 0x2B73F) or (0x2B740 <= cp <= 0x2B81F) or (0x2B820 <= cp <= 0x2CEAF)
            or (0xF900 <= cp <= 0xFAFF) or (0x2F800 <= cp <= 0x2FA1F))


class WordPieceTokenizer:
    def __init__(self, vocab: dict[str, int] | list[str], do_lower_case: bool = True, unk_token: str = "[UNK]",
                 cls_token: str = "[CLS]", sep_token: str = "[SEP]", pad_token: str = "[PAD]",
                 max_input_chars_per_word: int = 100, model_max_length: int = 512):
        if not isinstance(vocab, dict):
            vocab = {tok: i for i, tok in enumerate(vocab)}
        self.vocab = vocab
        self.do_lower_case = do_lower_case
        self.unk_token = unk_token
        self.unk_id = vocab[unk_token]
        self.cls_id = vocab[cls_token]
        self.sep_id = vocab[sep_token]
        self.pad_id = vocab[pad_token]
        self.never_split = {unk_token, cls_token, sep_token, pad_token, "[MASK]"}
        self.max_input_chars_per_word = max_input_chars_per_word
        self.model_max_length = model_max_length

    @classmethod
    def from_vocab_file(cls, path, **kw) -> "WordPieceTokenizer":
        tokens = Path(path).read_text(encoding="utf-8").split("\n")
        if tokens and tokens[-1] == "":
            tokens.pop()
        return cls({tok.rstrip("\n"): i for i, tok in enumerate(tokens)}, **kw)

    # ---------------------------------------------------------------- BasicTokenizer
    def _basic(self, text: str) -> list[str]:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            if _is_cjk(cp):
                out.append(f" {ch} ")
            elif _is_whitespace(ch):
                out.append(" ")
            else:
                out.append(ch)
        text = unicodedata.normalize("NFC", "".join(out))
        words = []
        for tok in text.strip().split():
            if tok not in self.never_split and self.do_lower_case:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            if tok in self.never_split:
                words.append(tok)
                continue
            cur = []
            for ch in tok:
                if _is_punctuation(ch):
                    if cur:
                        words.append("".join(cur))
                        cur = []
                    words.append(ch)
                else:
                    cur.append(ch)
            if cur:
                words.append("".join(cur))
        return " ".join(words).split()

    # ---------------------------------------------------------------- WordPiece
    def _wordpiece(self, word: str) -> list[int]:
        if len(word) > self.max_input_chars_per_word:
            return [self.unk_id]
        ids, start = [], 0
        while start < len(word):
            end = len(word)
            cur = None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = self.vocab[sub]
                    break
                end -= 1
            if cur is None:
                return [self.unk_id]
            ids.append(cur)
            start = end
        return ids

    def tokenize_ids(self, text: str) -> list[int]:
        ids = []
        for w in self._basic(text):
            ids.extend(self._wordpiece(w))
        return ids

    def encode(self, text: str, max_length: int | None = None) -> list[int]:
        """[CLS] tokens [SEP], truncated to max_length (longest-first == cut the tail)."""
        limit = min(max_length or self.model_max_length, self.model_max_length)
        ids = self.tokenize_ids(text)[: max(0, limit - 2)]
        return [self.cls_id] + ids + [self.sep_id]

    def batch(self, texts: list[str], max_length: int | None = None):
        """-> (input_ids int32 [B, S], attention_mask int32 [B, S]), right padded to the longest."""
        enc = [self.encode(t, max_length) for t in texts]
        s = max((len(e) for e in enc), default=1)
        ids = np.full((len(enc), s), self.pad_id, dtype=np.int32)
        mask = np.zeros((len(enc), s), dtype=np.int32)
        for i, e in enumerate(enc):
            ids[i, : len(e)] = e
            mask[i, : len(e)] = 1
        return ids, mask
