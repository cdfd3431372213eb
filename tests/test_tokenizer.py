"""CPU suite: the host WordPiece tokenizer against HF's BertTokenizer (the tokenizer
sentence-transformers uses for BERT-class models) on a synthetic vocabulary."""

import numpy as np
import pytest

from lean_explore_b200.tokenizer import WordPieceTokenizer

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "the", "a", "of", "nat", "##ural", "number", "##s", "add",
         "comm", "##ut", "##ative", "theorem", "lemma", "∀", "→", "(", ")", ":", ".", ",", "=", "+", "x", "y",
         "n", "m", "prime", "is", "in", "##finite", "##ly", "there", "are", "many", "real", "cafe", "中", "文",
         "group", "ring", "##oid", "mon", "list", "map", "_", "succ", "zero", "le", "lt", "##_", "0", "1", "##1", "z"]

TEXTS = [
    "the natural numbers",
    "theorem Nat.add_comm (n m : Nat) : n + m = m + n",
    "There are infinitely many primes.",
    "∀ x y, x + y = y + x → commutative",
    "Café monoid 中文 list.map",
    "   ",
    "",
    "a" * 150 + " the",
    "UNKNOWNWORD the ring",
    "x y�\tz\n the ring\x00\x07",
]


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    from transformers import BertTokenizer

    d = tmp_path_factory.mktemp("vocab")
    vocab = list(dict.fromkeys(VOCAB))
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n", encoding="utf-8")
    ours = WordPieceTokenizer.from_vocab_file(d / "vocab.txt", do_lower_case=True)
    hf = BertTokenizer(str(d / "vocab.txt"), do_lower_case=True)
    return ours, hf


@pytest.mark.parametrize("text", TEXTS)
def test_matches_hf_bert_tokenizer(pair, text):
    ours, hf = pair
    assert ours.encode(text) == hf.encode(text, add_special_tokens=True)


def test_truncation_and_batch_padding(pair):
    ours, hf = pair
    long = "the natural numbers " * 50
    for ml in (8, 16, 512):
        assert ours.encode(long, ml) == hf.encode(long, add_special_tokens=True, truncation=True, max_length=ml)
    ids, mask = ours.batch(["the", long, ""], max_length=12)
    ref = hf(["the", long, ""], padding=True, truncation=True, max_length=12, return_tensors="np")
    assert ids.dtype == np.int32 and mask.dtype == np.int32
    assert np.array_equal(ids, ref["input_ids"]) and np.array_equal(mask, ref["attention_mask"])


def test_cased_vocab_keeps_case(tmp_path):
    (tmp_path / "vocab.txt").write_text("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]", "Nat", "nat"]) + "\n")
    tok = WordPieceTokenizer.from_vocab_file(tmp_path / "vocab.txt", do_lower_case=False)
    assert tok.encode("Nat nat") == [2, 5, 6, 3]
