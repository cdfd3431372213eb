"""GPU parity tests of the Qwen3-class decoder (lxg_decoder_embed / lxg_decoder_rerank through the
C ABI) against the HF Qwen3 oracle (oracle/qwen3_decoder.py) and the committed golden vectors.

Tolerances: 1e-3 on the unit-norm embedding vectors (north-star tolerance; fp16 tensor-core
operands vs the fp32 oracle).  Reranker scores are probabilities sigma(logit_true - logit_false):
the reference itself runs this model in fp16 on CUDA (reranker_client.py:77-81), so scores are
compared at 5e-3 absolute and their ranking must agree wherever the oracle separates two documents
by more than twice that."""

import asyncio
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import qwen3_decoder as qd

pytestmark = pytest.mark.gpu

TOL_EMB = 1e-3
TOL_SCORE = 5e-3
GOLDEN = Path(__file__).resolve().parent / "golden" / "decoder_golden.npz"
TOKEN_TRUE, TOKEN_FALSE = 1837, 3082

_models = {}


def _pair(geom):
    """(oracle model, product decoder) per geometry, built once per session."""
    if geom not in _models:
        from lean_explore_b200.decoder import Qwen3Decoder

        model, cfg = qd.make_model(geom, seed=0)
        dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                           heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads,
                           ffn=cfg.intermediate_size, head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps,
                           rope_theta=1e6)
        _models[geom] = (model, cfg, dec)
    return _models[geom]


def _compare_emb(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.isfinite(got).all()
    assert np.abs(np.linalg.norm(got.astype(np.float64), axis=1) - 1).max() < 1e-5
    err = np.abs(got - want).max()
    assert err < TOL_EMB, f"max abs err {err}"
    assert (got * want).sum(1).min() > 0.9999


def _compare_scores(got, want):
    assert got.shape == want.shape and got.dtype == np.float32
    assert ((got >= 0) & (got <= 1)).all()
    err = np.abs(got - want).max()
    assert err < TOL_SCORE, f"max abs err {err}"
    for i in range(len(want)):
        for j in range(len(want)):
            if want[i] - want[j] > 2 * TOL_SCORE:
                assert got[i] > got[j]


@pytest.mark.parametrize("geom,b,s,side", [
    ("tiny", 5, 19, "left"), ("tiny", 1, 1, "left"), ("tiny", 3, 70, "right"), ("tiny", 2, 300, "left"),
    ("small", 4, 33, "left"), ("small", 7, 129, "left"), ("qwen3-0.6b", 3, 21, "left"), ("qwen3-0.6b", 2, 150, "left"),
    # several 128-key chunks per query tile (tcgen05 attention: running-maximum rescale of the accumulator)
    ("tiny", 2, 520, "left"), ("tiny", 3, 260, "right"),
    # at most 32 tokens: the query path's narrow-tile GEMMs (32- / 64-column tiles, 32-row A boxes)
    ("qwen3-0.6b", 1, 24, "left"), ("small", 1, 32, "left"), ("tiny", 2, 16, "left"), ("small", 2, 9, "right"),
])
def test_embedding_matches_hf_oracle(geom, b, s, side):
    model, cfg, dec = _pair(geom)
    ids, mask = qd.make_inputs(b, s, seed=7, side=side)
    _compare_emb(dec.embed_ids(ids, mask), qd.embed(model, ids, mask))
    assert dec.last_launches() == 1 + 8 * cfg.num_hidden_layers


@pytest.mark.parametrize("geom,b,s", [("tiny", 6, 40), ("small", 5, 77), ("qwen3-0.6b", 4, 64)])
def test_reranker_scores_match_hf_oracle(geom, b, s):
    model, cfg, dec = _pair(geom)
    ids, mask = qd.make_inputs(b, s, seed=3, side="left")
    want = qd.rerank(model, ids, mask, TOKEN_TRUE, TOKEN_FALSE)
    _compare_scores(dec.rerank_ids(ids, mask, TOKEN_TRUE, TOKEN_FALSE), want)
    # swapping the two tokens gives the complementary probability
    swapped = dec.rerank_ids(ids, mask, TOKEN_FALSE, TOKEN_TRUE)
    assert np.abs(swapped + dec.rerank_ids(ids, mask, TOKEN_TRUE, TOKEN_FALSE) - 1).max() < 1e-5


@pytest.mark.parametrize("scale", [4.0, 8.0])
def test_large_activations_near_the_fp16_range(scale):
    """Real Qwen3 checkpoints carry large-magnitude channels; the random-init models above do not.
    Scale the MLP input projections (gate / up) and a few residual channels so that SwiGLU outputs and
    the residual stream run far above the init_std = 0.05 regime, and require the same tolerances.
    The product keeps SwiGLU outputs and GEMM operands in fp16 - this is the
    test that says how much head-room that leaves."""
    from lean_explore_b200.decoder import Qwen3Decoder

    model, cfg = qd.make_model("small", seed=2)
    with torch.no_grad():
        for layer in model.model.layers:
            layer.mlp.gate_proj.weight.mul_(scale)
            layer.mlp.up_proj.weight.mul_(scale)
        model.model.embed_tokens.weight[:, :4].mul_(40.0)  # a few "massive" residual channels
    dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                       heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                       head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
    ids, mask = qd.make_inputs(4, 48, seed=9, side="left")
    with torch.no_grad():  # how large the oracle's SwiGLU outputs actually get (reported on failure)
        hs = model.model(input_ids=torch.from_numpy(ids).long(), attention_mask=torch.from_numpy(mask).long(),
                         output_hidden_states=True).hidden_states
        peak = max(float(h.abs().max()) for h in hs)
    assert peak > 50.0, f"the stress model is not stressing anything (peak |h| = {peak})"
    _compare_emb(dec.embed_ids(ids, mask), qd.embed(model, ids, mask))
    _compare_scores(dec.rerank_ids(ids, mask, TOKEN_TRUE, TOKEN_FALSE), qd.rerank(model, ids, mask, TOKEN_TRUE, TOKEN_FALSE))


def test_matches_committed_golden_vectors():
    z = np.load(GOLDEN)
    for key in sorted({k.split("/")[0] for k in z.files}):
        geom = key.rsplit("_", 2)[0]
        _, _, dec = _pair(geom)
        _compare_emb(dec.embed_ids(z[key + "/ids"], z[key + "/mask"]), z[key + "/emb"])
        if key + "/score" in z.files:
            _compare_scores(dec.rerank_ids(z[key + "/ids"], z[key + "/mask"], TOKEN_TRUE, TOKEN_FALSE), z[key + "/score"])


def test_padding_content_graph_replay_and_device_io():
    """Garbage under the mask must not matter; the second call with a shape replays a CUDA graph
    and must give the same bits; device-resident ids in / vectors out."""
    model, cfg, dec = _pair("small")
    ids, mask = qd.make_inputs(16, 48, seed=11, side="left")
    a = dec.embed_ids(ids, mask)
    b = dec.embed_ids(ids, mask)  # graph replay
    assert (a == b).all()
    ids2 = np.where(mask == 1, ids, 777).astype(np.int32)
    assert np.abs(dec.embed_ids(ids2, mask) - a).max() < 1e-6
    assert dec.last_tokens() == int(mask.sum())  # host batch: padding tokens were dropped before layer 0
    want = qd.embed(model, ids, mask)
    got = dec.embed_ids_torch(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda())  # device ids: padded rectangle
    torch.cuda.synchronize()
    assert dec.last_tokens() == ids.size
    _compare_emb(got.cpu().numpy(), want)
    got2 = dec.embed_ids_torch(torch.from_numpy(ids).cuda(), torch.from_numpy(mask).cuda())  # graph replay
    torch.cuda.synchronize()
    assert (got2 == got).all()
    assert np.abs(got.cpu().numpy() - a).max() < TOL_EMB  # packed and padded formulations agree
    _compare_emb(a, want)
    # left padding == the same sequences without padding, one by one (positions shift, RoPE is relative)
    for i in (1, 5):
        n = int(mask[i].sum())
        solo = dec.embed_ids(ids[i : i + 1, 48 - n :], np.ones((1, n), np.int32))
        assert np.abs(solo[0] - a[i]).max() < TOL_EMB


def test_packed_and_padded_paths_agree_and_holes_fall_back(monkeypatch):
    """LXG_DECODER_PACK=0 keeps the padded rectangle for host batches too; a mask with a hole is not
    a padding pattern, so it takes the padded path (HF semantics: positions keep counting, the
    hole is only masked as a key)."""
    from lean_explore_b200.decoder import Qwen3Decoder

    model, cfg, dec = _pair("tiny")
    monkeypatch.setenv("LXG_DECODER_PACK", "0")
    padded = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers,
                          heads=cfg.num_attention_heads, kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size,
                          head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)
    for side in ("left", "right"):
        ids, mask = qd.make_inputs(9, 90, seed=21, side=side)
        a, b = dec.embed_ids(ids, mask), padded.embed_ids(ids, mask)
        assert dec.last_tokens() == int(mask.sum()) and padded.last_tokens() == ids.size
        assert np.abs(a - b).max() < TOL_EMB
        _compare_emb(a, qd.embed(model, ids, mask))
    ids, mask = qd.make_inputs(4, 30, seed=22, side="left")
    mask[2, 20] = 0  # a hole
    got = dec.embed_ids(ids, mask)
    assert dec.last_tokens() == ids.size
    _compare_emb(got, qd.embed(model, ids, mask))
    sc = dec.rerank_ids(ids, mask, TOKEN_TRUE, TOKEN_FALSE)
    _compare_scores(sc, qd.rerank(model, ids, mask, TOKEN_TRUE, TOKEN_FALSE))


def test_argument_errors():
    from lean_explore_b200 import _lib

    _, _, dec = _pair("tiny")
    ids, mask = qd.make_inputs(2, 8, seed=0)
    with pytest.raises(ValueError):
        dec.embed_ids(ids, mask[:, :4])
    with pytest.raises(_lib.LxgError):
        dec.rerank_ids(ids, mask, 10**6, 3)
    assert dec.embed_ids(ids[:0], mask[:0]).shape == (0, 256)


class _Tok:
    """Whitespace 'tokenizer' with the interface the clients use (no Qwen vocabulary exists offline)."""

    def __init__(self, vocab_size):
        self.vocab_size = vocab_size

    def convert_tokens_to_ids(self, t):
        return {"true": TOKEN_TRUE, "false": TOKEN_FALSE}[t]

    def _ids(self, text):
        import zlib

        return [10 + zlib.crc32(w.encode()) % (self.vocab_size - 10) for w in text.split()]

    def batch(self, texts, max_length=None):
        enc = [self._ids(t)[: max_length or 512] for t in texts]
        s = max(1, max(len(e) for e in enc))
        ids = np.zeros((len(enc), s), np.int32)
        mask = np.zeros((len(enc), s), np.int32)
        for i, e in enumerate(enc):
            if e:
                ids[i, s - len(e):] = e
                mask[i, s - len(e):] = 1
        return ids, mask


def test_reranker_client_duck_type():
    """GpuRerankerClient mirrors RerankerClient (reranker_client.py:143-205): scores in input order,
    empty input, inline vs executor batching give the same numbers."""
    from lean_explore_b200.reranker_client import GpuRerankerClient, RerankerResponse

    model, cfg, dec = _pair("small")
    dec.tokenizer = _Tok(cfg.vocab_size)
    client = GpuRerankerClient("Qwen/Qwen3-Reranker-0.6B", model=dec, batch_size=4, max_length=64)
    docs = [f"theorem number {i} about " + " ".join(f"w{j}" for j in range(i % 7 + 1)) for i in range(10)]
    r = client.rerank_sync("prime numbers", docs)
    assert isinstance(r, RerankerResponse) and r.query == "prime numbers" and r.model == "Qwen/Qwen3-Reranker-0.6B"
    assert len(r.scores) == len(docs) and all(0.0 <= s <= 1.0 for s in r.scores)
    pairs = [client._format_pair("prime numbers", d) for d in docs]
    assert pairs[0].startswith("<Instruct>: Find relevant Lean 4 math declarations\n<Query>: prime numbers\n<Document>: ")
    ids, mask = dec.tokenizer.batch(pairs, 64)
    _compare_scores(np.asarray(r.scores, np.float32), qd.rerank(model, ids, mask, TOKEN_TRUE, TOKEN_FALSE))
    r2 = asyncio.run(client.rerank("prime numbers", docs))  # 10 docs > batch_size 4: executor path
    assert np.abs(np.asarray(r2.scores) - np.asarray(r.scores)).max() < TOL_SCORE
    assert asyncio.run(client.rerank("q", [])).scores == []
    dec.tokenizer = None


def test_tma_reduce_epilogue_is_bit_identical_to_the_per_thread_one(tmp_path):
    """o_proj / down_proj accumulate onto the fp32 residual stream either through one
    cp.reduce.async.bulk.tensor per 32 x 32 block (default) or by the row's thread (LXG_GEMM_TMA_ACC=0).
    Every output element is the same single fp32 addition either way, so the two builds of the forward must
    agree bit for bit.  The switch is read once per process: two child processes."""
    import os
    import subprocess
    import sys

    script = tmp_path / "run.py"
    script.write_text(
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {str(Path(__file__).resolve().parents[1])!r})\n"
        "from oracle import qwen3_decoder as qd\n"
        "from lean_explore_b200.decoder import Qwen3Decoder\n"
        "model, cfg = qd.make_model('small', seed=0)\n"
        "dec = Qwen3Decoder(model.state_dict(), hidden=cfg.hidden_size, layers=cfg.num_hidden_layers, heads=cfg.num_attention_heads,\n"
        "                   kv_heads=cfg.num_key_value_heads, ffn=cfg.intermediate_size, head_dim=cfg.head_dim, rms_eps=cfg.rms_norm_eps, rope_theta=1e6)\n"
        "ids, mask = qd.make_inputs(7, 129, seed=7, side='left')\n"
        "np.save(sys.argv[1], dec.embed_ids(ids, mask))\n")
    outs = []
    for flag in ("1", "0"):
        out = tmp_path / f"emb_{flag}.npy"
        env = dict(os.environ, LXG_GEMM_TMA_ACC=flag)
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=env, timeout=300)
        outs.append(np.load(out))
    assert outs[0].shape == (7, outs[0].shape[1]) and np.isfinite(outs[0]).all()
    assert np.array_equal(outs[0], outs[1])
