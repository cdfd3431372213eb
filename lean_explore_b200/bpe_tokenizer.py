"""Byte-level BPE tokenizer of the Qwen2 / Qwen3 family on the host (pure Python).

The reference gets tokenisation from third-party wrappers: ``SentenceTransformer`` for the
embedding model (``src/lean_explore/util/embedding_client.py:58``) and
``AutoTokenizer.from_pretrained(model_name, padding_side="left")`` for the reranker
(``src/lean_explore/util/reranker_client.py:71-73,119-125``: ``padding=True, truncation=True,
max_length=...``).  This restates the published algorithm of ``transformers``'
``Qwen2Tokenizer`` (``models/qwen2/tokenization_qwen2.py``): NFC normalisation -> split on the
Qwen2 pre-tokenisation regex -> GPT-2 byte-to-unicode mapping -> BPE merges in rank order; added
(special) tokens are matched verbatim first.  ``tests/test_bpe_tokenizer.py`` checks it against
the ``tokenizers`` library configured the same way.
"""

from __future__ import annotations

import json
import unicodedata
from functools import lru_cache
from pathlib import Path

import numpy as np
import regex

PRETOKENIZE_REGEX = (
    r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+"""
)


@lru_cache(maxsize=1)
def bytes_to_unicode() -> dict[int, str]:
    """GPT-2's reversible byte -> printable unicode character table."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


class ByteLevelBPETokenizer:
    def __init__(self, vocab: dict[str, int], merges: list[tuple[str, str]], added_tokens: dict[str, int] | None = None,
                 pad_token: str = "<|endoftext|>", eos_token: str = "<|endoftext|>", append_eos: bool = False,
                 padding_side: str = "left", model_max_length: int = 32768):
        self.vocab = dict(vocab)
        self.ranks = {pair: i for i, pair in enumerate(merges)}
        self.added = dict(added_tokens or {})
        for tok, i in self.added.items():
            self.vocab.setdefault(tok, i)
        self.pad_id = self.vocab.get(pad_token, 0)
        self.eos_id = self.vocab.get(eos_token, self.pad_id)
        self.append_eos = append_eos
        self.padding_side = padding_side
        self.model_max_length = model_max_length
        self._pat = regex.compile(PRETOKENIZE_REGEX)
        self._b2u = bytes_to_unicode()
        self._added_pat = None
        if self.added:
            alts = sorted(self.added, key=len, reverse=True)
            self._added_pat = regex.compile("(" + "|".join(regex.escape(a) for a in alts) + ")")
        self._cache: dict[str, list[int]] = {}

    # ---------------------------------------------------------------- loading
    @classmethod
    def from_dir(cls, d: str | Path, padding_side: str = "left") -> "ByteLevelBPETokenizer":
        d = Path(d)
        added: dict[str, int] = {}
        append_eos = False
        tj = d / "tokenizer.json"
        tok_json = json.loads(tj.read_text()) if tj.exists() else None
        if (d / "vocab.json").exists() and (d / "merges.txt").exists():
            vocab = json.loads((d / "vocab.json").read_text())
            merges = []
            for line in (d / "merges.txt").read_text().splitlines():
                if not line or line.startswith("#version"):
                    continue
                a, b = line.split(" ")
                merges.append((a, b))
        elif tok_json is not None:
            vocab = tok_json["model"]["vocab"]
            merges = [tuple(m.split(" ")) if isinstance(m, str) else tuple(m) for m in tok_json["model"]["merges"]]
        else:
            raise FileNotFoundError(f"no vocab.json + merges.txt or tokenizer.json under {d}")
        if tok_json is not None:
            for t in tok_json.get("added_tokens", []):
                added[t["content"]] = int(t["id"])
            # Qwen3-Embedding's tokenizer.json appends <|endoftext|> through a TemplateProcessing step
            post = tok_json.get("post_processor") or {}
            procs = post.get("processors", [post]) if post else []
            for pr in procs:
                if pr.get("type") == "TemplateProcessing":
                    single = pr.get("single", [])
                    append_eos = any("SpecialToken" in s for s in single[1:])
        cfg = {}
        tc = d / "tokenizer_config.json"
        if tc.exists():
            cfg = json.loads(tc.read_text())
            for i, t in (cfg.get("added_tokens_decoder") or {}).items():
                added.setdefault(t["content"], int(i))

        def name(v, default):
            if isinstance(v, dict):
                return v.get("content", default)
            return v or default

        mml = cfg.get("model_max_length", 32768)
        return cls(vocab, merges, added, pad_token=name(cfg.get("pad_token"), "<|endoftext|>"),
                   eos_token=name(cfg.get("eos_token"), "<|endoftext|>"), append_eos=append_eos,
                   padding_side=padding_side, model_max_length=int(mml) if mml and mml < 10**9 else 32768)

    # ---------------------------------------------------------------- BPE
    def _bpe(self, word: str) -> list[int]:
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts = list(word)
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.ranks.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, i = [], 0
            while i < len(parts):  # merge every occurrence of the best pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        ids = [self.vocab[p] for p in parts if p in self.vocab]
        if len(self._cache) < 1 << 16:
            self._cache[word] = ids
        return ids

    def _encode_plain(self, text: str) -> list[int]:
        ids: list[int] = []
        for piece in self._pat.findall(unicodedata.normalize("NFC", text)):
            ids.extend(self._bpe("".join(self._b2u[b] for b in piece.encode("utf-8"))))
        return ids

    def tokenize_ids(self, text: str) -> list[int]:
        if self._added_pat is None:
            return self._encode_plain(text)
        ids: list[int] = []
        for chunk in self._added_pat.split(text):
            if not chunk:
                continue
            if chunk in self.added:
                ids.append(self.added[chunk])
            else:
                ids.extend(self._encode_plain(chunk))
        return ids

    def convert_tokens_to_ids(self, token: str) -> int:
        """``AutoTokenizer.convert_tokens_to_ids`` (``reranker_client.py:85-86`` looks up "true" / "false")."""
        return self.vocab[token]

    def encode(self, text: str, max_length: int | None = None) -> list[int]:
        limit = min(max_length or self.model_max_length, self.model_max_length)
        ids = self.tokenize_ids(text)
        if self.append_eos:
            return ids[: max(0, limit - 1)] + [self.eos_id]
        return ids[:limit]

    def batch(self, texts: list[str], max_length: int | None = None):
        """-> (input_ids int32 [B, S], attention_mask int32 [B, S]) padded to the longest on
        ``padding_side`` (left, as both reference clients configure Qwen3 tokenizers)."""
        enc = [self.encode(t, max_length) for t in texts]
        s = max(1, max((len(e) for e in enc), default=1))
        ids = np.full((len(enc), s), self.pad_id, dtype=np.int32)
        mask = np.zeros((len(enc), s), dtype=np.int32)
        for i, e in enumerate(enc):
            if not e:
                continue
            if self.padding_side == "left":
                ids[i, s - len(e):] = e
                mask[i, s - len(e):] = 1
            else:
                ids[i, : len(e)] = e
                mask[i, : len(e)] = 1
        return ids, mask
