"""CPU suite: the byte-level BPE tokenizer (Qwen2/Qwen3 family) against the `tokenizers` library
configured exactly as transformers' Qwen2Tokenizer configures it (NFC, Qwen2 split regex,
ByteLevel without its own regex, BPE).  No Qwen vocabulary exists offline, so a small BPE
vocabulary is trained here on a fixed text; the algorithm under test does not depend on it."""

import json

import numpy as np
import pytest

from lean_explore_b200.bpe_tokenizer import PRETOKENIZE_REGEX, ByteLevelBPETokenizer, bytes_to_unicode

TRAIN = [
    "theorem Nat.add_comm (a b : ℕ) : a + b = b + a := by induction a <;> simp [*]",
    "lemma continuous_of_lipschitz {f : ℝ → ℝ} (hf : LipschitzWith K f) : Continuous f",
    "A group homomorphism maps the identity to the identity and inverses to inverses.",
    "The sum of two even numbers is even; the product of any number with an even number is even.",
    "<Instruct>: Find relevant Lean 4 math declarations\n<Query>: prime numbers\n<Document>: Nat.Prime p",
    "def List.map {α β : Type*} (f : α → β) : List α → List β", "we're it's don't I'LL 12345 3.14159  \t tabs\r\n\r\nnewlines   ",
] * 4

SAMPLES = [
    "", " ", "a", "theorem foo : 1 + 1 = 2 := rfl", "  leading and trailing  ", "naïve café ℕ → ℝ ∀ ε > 0, ∃ δ",
    "I'm we'RE don't THEY'LL", "x1y22z333", "line1\nline2\r\n\r\n  line3", "tabs\t\tand   spaces ",
    "<Instruct>: Find relevant Lean 4 math declarations\n<Query>: commutativity of addition\n<Document>: Nat.add_comm",
    "日本語のテキスト and emoji 🎉🎉", "é combining vs é", "a<|endoftext|>b <|im_start|>user", "'s't're",
]


@pytest.fixture(scope="module")
def toks(tmp_path_factory):
    from tokenizers import AddedToken, Regex, Tokenizer, normalizers, pre_tokenizers, trainers
    from tokenizers.models import BPE

    def configure(t):
        t.normalizer = normalizers.NFC()
        t.pre_tokenizer = pre_tokenizers.Sequence([
            pre_tokenizers.Split(Regex(PRETOKENIZE_REGEX), behavior="isolated", invert=False),
            pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])

    t = Tokenizer(BPE(unk_token=None, continuing_subword_prefix="", end_of_word_suffix="", fuse_unk=False, byte_fallback=False))
    configure(t)
    trainer = trainers.BpeTrainer(vocab_size=600, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(), special_tokens=[],
                                  show_progress=False)
    t.train_from_iterator(TRAIN, trainer)
    d = tmp_path_factory.mktemp("qwen_tok")
    t.model.save(str(d))  # vocab.json + merges.txt
    vocab = json.loads((d / "vocab.json").read_text())
    specials = {"<|endoftext|>": len(vocab), "<|im_start|>": len(vocab) + 1, "<|im_end|>": len(vocab) + 2}
    (d / "tokenizer_config.json").write_text(json.dumps({
        "added_tokens_decoder": {str(i): {"content": c, "special": True} for c, i in specials.items()},
        "pad_token": "<|endoftext|>", "eos_token": "<|endoftext|>", "model_max_length": 131072}))
    ref = Tokenizer(BPE.from_file(str(d / "vocab.json"), str(d / "merges.txt")))
    configure(ref)
    ref.add_special_tokens([AddedToken(c, special=True) for c in specials])
    assert all(ref.token_to_id(c) == i for c, i in specials.items())
    return ByteLevelBPETokenizer.from_dir(d), ref, specials


def test_byte_table_is_a_bijection():
    m = bytes_to_unicode()
    assert len(m) == 256 and len(set(m.values())) == 256
    assert m[ord("A")] == "A" and m[ord(" ")] == "Ġ" and m[ord("\n")] == "Ċ"


def test_matches_tokenizers_library(toks):
    mine, ref, _ = toks
    for text in SAMPLES + TRAIN[:7]:
        assert mine.tokenize_ids(text) == ref.encode(text, add_special_tokens=False).ids, repr(text)


def test_random_strings_match(toks):
    mine, ref, _ = toks
    rng = np.random.default_rng(0)
    alphabet = list("abcdefghij xyzABC0123456789.,;:'()[]{}+-=*/\n\t→ℕℝ∀∃éü")
    for _ in range(200):
        text = "".join(rng.choice(alphabet, size=int(rng.integers(1, 60))))
        assert mine.tokenize_ids(text) == ref.encode(text, add_special_tokens=False).ids, repr(text)


def test_wide_unicode_fuzz_matches(toks):
    """Letters / digits / punctuation / whitespace of several scripts, combining marks (NFC matters),
    astral code points, contractions in both cases, runs of newlines and spaces."""
    mine, ref, _ = toks
    rng = np.random.default_rng(1)
    pieces = ["a", "Z", "é", "e\u0301", "ß", "ℕ", "→", "∀", "日", "本", "한", "🎉", "👩\u200d🔬", "0", "7", "٣", "½", " ", "  ", "\t", "\n", "\r\n",
              "\u00a0", "\u2003", "'s", "'RE", "'ll", "'", "\"", ".", ",", "(", ")", "_", "-", "+=", "<|", "|>", "theorem", "Nat", "add_comm"]
    for _ in range(400):
        text = "".join(rng.choice(pieces, size=int(rng.integers(1, 40))))
        assert mine.tokenize_ids(text) == ref.encode(text, add_special_tokens=False).ids, repr(text)


def test_left_padding_truncation_and_eos(toks):
    mine, _, specials = toks
    pad = specials["<|endoftext|>"]
    texts = ["a + b = b + a", "theorem Nat.add_comm (a b : ℕ) : a + b = b + a := by induction a <;> simp [*]", ""]
    ids, mask = mine.batch(texts, max_length=16)
    assert ids.dtype == np.int32 and ids.shape == mask.shape and ids.shape[1] <= 16
    for row_ids, row_mask, text in zip(ids, mask, texts):
        want = mine.tokenize_ids(text)[:16]
        n = len(want)
        assert row_mask.sum() == n and (row_mask[ids.shape[1] - n:] == 1).all()  # left padded
        assert list(row_ids[ids.shape[1] - n:]) == want and (row_ids[: ids.shape[1] - n] == pad).all()
    mine.append_eos = True  # Qwen3-Embedding's post-processor appends <|endoftext|>
    try:
        ids, mask = mine.batch(texts[:2], max_length=8)
        assert ids.shape[1] == 8 and (ids[:, -1] == pad).all() and mask[:, -1].all()
        assert list(ids[1, :7]) == mine.tokenize_ids(texts[1])[:7]
    finally:
        mine.append_eos = False
    assert mine.convert_tokens_to_ids("<|im_end|>") == specials["<|im_end|>"]
