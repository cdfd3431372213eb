"""GPU parity tests of the sentence encoder (lxg_encode through the C ABI) against the HF
BertModel oracle (oracle/bert_encoder.py) and the committed golden vectors.  Tolerance: 1e-3 on
the unit-norm output vectors (fp16 tensor-core compute vs fp32 oracle)."""

import asyncio
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import bert_encoder as be

pytestmark = pytest.mark.gpu

TOL = 1e-3
GOLDEN = Path(__file__).resolve().parent / "golden" / "encoder_golden.npz"


def _encoder(model, cfg, pool):
    from lean_explore_b200.encoder import POOL_CLS, POOL_MEAN, BertSentenceEncoder

    return BertSentenceEncoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                               heads=cfg.num_attention_heads, ffn=cfg.intermediate_size, ln_eps=cfg.layer_norm_eps,
                               pool=POOL_CLS if pool == "cls" else POOL_MEAN)


def _compare(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.isfinite(got).all()
    assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - 1).max() < 1e-5
    err = np.abs(got - want).max()
    assert err < TOL, f"max abs err {err}"
    assert (got * want).sum(1).min() > 0.9999


@pytest.mark.parametrize("geom,b,s,pool", [
    ("tiny", 5, 19, "mean"), ("tiny", 3, 8, "cls"), ("tiny", 1, 1, "mean"),
    ("minilm-l6", 4, 24, "mean"), ("minilm-l6", 9, 33, "mean"), ("minilm-l6", 3, 130, "cls"),
    ("bge-base", 4, 24, "cls"), ("bge-base", 2, 17, "mean"), ("bge-base", 6, 40, "cls"),
])
def test_encoder_matches_hf_oracle(geom, b, s, pool):
    model, cfg = be.make_model(geom, seed=0)
    ids, mask = be.make_inputs(b, s, seed=7)
    want = be.encode(model, ids, mask, pool)
    enc = _encoder(model, cfg, pool)
    got = enc.encode_ids(ids, mask)
    _compare(got, want)
    # the query path (<= 64 tokens, <= 32 for hidden >= 768): the single persistent kernel; more: the layered kernels
    fused = b * s <= (64 if cfg.hidden_size <= 512 else 32)
    assert enc.last_launches() == (1 if fused else 2 + 7 * cfg.num_hidden_layers)


@pytest.mark.parametrize("geom,b,s,pool", [
    ("tiny", 1, 1, "cls"), ("tiny", 7, 9, "mean"), ("tiny", 1, 64, "mean"),
    ("minilm-l6", 1, 5, "mean"), ("minilm-l6", 1, 16, "mean"), ("minilm-l6", 1, 33, "mean"), ("minilm-l6", 1, 64, "cls"),
    ("minilm-l6", 4, 7, "mean"), ("minilm-l6", 8, 8, "mean"), ("minilm-l6", 3, 21, "cls"),
    ("bge-base", 1, 9, "cls"), ("bge-base", 1, 16, "mean"), ("bge-base", 1, 32, "cls"), ("bge-base", 2, 13, "mean"),
    ("bge-base", 3, 10, "cls"),
])
def test_query_path_single_kernel_matches_oracle_and_layered_path(geom, b, s, pool):
    """The query path (<= 64 tokens, EmbeddingClient.embed([query]) at search/engine.py:236) is one
    cooperative kernel: same tolerance against the HF oracle as the layered kernels, and the two
    paths agree with each other far inside it (same operand precision, different summation split)."""
    model, cfg = be.make_model(geom, seed=0)
    ids, mask = be.make_inputs(b, s, seed=3 + b + s)
    want = be.encode(model, ids, mask, pool)
    enc = _encoder(model, cfg, pool)
    got = enc.encode_ids(ids, mask)
    assert enc.last_launches() == 1
    _compare(got, want)
    again = enc.encode_ids(ids, mask)
    assert (again == got).all(), "the single kernel must be deterministic (fixed-order split-K sums)"
    enc.set_fused(False)
    layered = enc.encode_ids(ids, mask)
    assert enc.last_launches() == 2 + 7 * cfg.num_hidden_layers
    assert np.abs(layered - got).max() < 5e-4  # fp16 roundings of different summation orders, 12 layers deep
    enc.set_fused(True)
    # garbage under the mask must not matter, nor a different call in between (workspace / counters reuse)
    enc.encode_ids(ids[:1, : min(3, s)], np.ones((1, min(3, s)), np.int32))
    ids2 = np.where(mask == 1, ids, 777).astype(np.int32)
    assert np.abs(enc.encode_ids(ids2, mask) - got).max() < 1e-6


def test_query_path_fully_masked_sequence_and_device_io():
    model, cfg = be.make_model("minilm-l6", seed=0)
    ids, mask = be.make_inputs(3, 12, seed=5)
    enc = _encoder(model, cfg, "mean")
    got = enc.encode_ids_torch(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    assert enc.last_launches() == 1
    _compare(got.cpu().numpy(), be.encode(model, ids, mask, "mean"))
    # a sequence whose mask is all zero: finite output (zero vector), the other rows unchanged
    mask2 = mask.copy()
    mask2[1] = 0
    got2 = enc.encode_ids(ids, mask2)
    assert np.isfinite(got2).all()
    assert np.abs(got2[[0, 2]] - got.cpu().numpy()[[0, 2]]).max() < 1e-6
    enc.set_fused(False)
    assert np.abs(enc.encode_ids(ids, mask2) - got2).max() < 2e-4


def test_encoder_matches_committed_golden_vectors():
    z = np.load(GOLDEN)
    for key in sorted({k.split("/")[0] for k in z.files}):
        geom, shape, pool = key.rsplit("_", 2)
        model, cfg = be.make_model(geom, seed=0)
        enc = _encoder(model, cfg, pool)
        got = enc.encode_ids(z[key + "/ids"], z[key + "/mask"])
        _compare(got, z[key + "/emb"])


def test_large_batch_many_row_tiles_and_device_io():
    """B*S = 2048 tokens (16 GEMM row tiles), device-resident ids in / embeddings out; the padding
    content of masked positions must not matter."""
    model, cfg = be.make_model("minilm-l6", seed=0)
    ids, mask = be.make_inputs(64, 32, seed=11)
    want = be.encode(model, ids, mask, "mean")
    enc = _encoder(model, cfg, "mean")
    got = enc.encode_ids_torch(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda())
    torch.cuda.synchronize()
    _compare(got.cpu().numpy(), want)
    ids2 = np.where(mask == 1, ids, 12345).astype(np.int32)  # garbage under the mask
    got2 = enc.encode_ids(ids2, mask)
    assert np.abs(got2 - got.cpu().numpy()).max() < 1e-6
    # smaller call on the same handle afterwards (workspace reuse)
    _compare(enc.encode_ids(ids[:3, :9], np.ones((3, 9), np.int32)), be.encode(model, ids[:3, :9], np.ones((3, 9), np.int32)))


def test_encoder_argument_errors():
    from lean_explore_b200 import _lib

    model, cfg = be.make_model("tiny", seed=0)
    enc = _encoder(model, cfg, "mean")
    with pytest.raises(_lib.LxgError):
        enc.encode_ids(np.zeros((1, 600), np.int32), np.ones((1, 600), np.int32))  # longer than the position table
    with pytest.raises(ValueError):
        enc.encode_ids(np.zeros((2, 4), np.int32), np.ones((2, 5), np.int32))
    assert enc.encode_ids(np.zeros((0, 4), np.int32), np.zeros((0, 4), np.int32)).shape == (0, cfg.hidden_size)


WORDS = ["the", "nat", "##ural", "number", "##s", "add", "comm", "theorem", "lemma", "prime", "is", "in", "##finite",
         "there", "are", "many", "real", "group", "ring", "list", "map", "(", ")", ":", ".", ",", "=", "+", "x", "y", "n", "m"]


def _write_model_dir(d: Path, model, cfg, pool):
    from safetensors.torch import save_file

    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n")
    (d / "config.json").write_text(json.dumps(dict(
        model_type="bert", hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, intermediate_size=cfg.intermediate_size,
        max_position_embeddings=cfg.max_position_embeddings, layer_norm_eps=cfg.layer_norm_eps, vocab_size=cfg.vocab_size)))
    (d / "tokenizer_config.json").write_text(json.dumps({"do_lower_case": True}))
    (d / "1_Pooling").mkdir()
    (d / "1_Pooling" / "config.json").write_text(json.dumps({
        "pooling_mode_cls_token": pool == "cls", "pooling_mode_mean_tokens": pool == "mean"}))
    save_file({k: v.contiguous() for k, v in model.state_dict().items()}, str(d / "model.safetensors"))


@pytest.mark.parametrize("pool", ["mean", "cls"])
def test_embedding_client_on_a_model_directory(tmp_path, pool):
    """The reference-facing call: GpuEmbeddingClient(model_name).embed(texts, is_query) ->
    EmbeddingResponse (embedding_client.py:73-106), from a sentence-transformers style directory."""
    from lean_explore_b200.embedding_client import EmbeddingResponse, GpuEmbeddingClient
    from lean_explore_b200.tokenizer import WordPieceTokenizer

    model, cfg = be.make_model("tiny", seed=3, vocab_size=5 + len(WORDS), max_pos=64)
    _write_model_dir(tmp_path, model, cfg, pool)
    client = GpuEmbeddingClient(model_name=str(tmp_path), max_length=16, batch_size=2)
    assert client.model_name == str(tmp_path) and client.batch_size == 2 and client.max_length == 16
    texts = ["there are infinitely many primes .", "theorem add comm ( n m : nat ) : n + m = m + n",
             "the natural numbers", "x", "a ring is a group " * 10]
    resp = asyncio.run(client.embed(texts, is_query=True))
    assert isinstance(resp, EmbeddingResponse) and resp.texts == texts and resp.model == str(tmp_path)
    got = np.array(resp.embeddings, dtype=np.float32)
    assert got.shape == (5, cfg.hidden_size)
    tok = WordPieceTokenizer.from_vocab_file(tmp_path / "vocab.txt")
    for i, t in enumerate(texts):  # oracle one text at a time: batching / length sorting must not matter
        ids, mask = tok.batch([t], 16)
        assert ids.shape[1] <= 16
        want = be.encode(model, ids, mask, pool)
        assert np.abs(got[i] - want[0]).max() < TOL
    with pytest.raises(FileNotFoundError):
        GpuEmbeddingClient(model_name="no/such-model")
