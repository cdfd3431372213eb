"""CPU suite: host-side logic of the Qwen3 path that needs no GPU - the gate/up interleave the
SwiGLU GEMM epilogue expects, tokenizer loading from the HF file variants, the clients' argument
handling (there is no CPU fallback: they must raise)."""

import json

import numpy as np
import pytest
import torch

from lean_explore_b200.bpe_tokenizer import ByteLevelBPETokenizer, bytes_to_unicode
from lean_explore_b200.decoder import interleave_gate_up, model_type_of


def test_gate_up_interleave_layout():
    """Rows 64 i .. 64 i + 31 = gate[32 i ..], rows 64 i + 32 .. 64 i + 63 = up[32 i ..] (include/lxg.h,
    lxg_qwen3_layer.wgu): the epilogue thread that owns output column c finds gate[c] and up[c] in the
    two 32-column chunks of the same 64-column group."""
    f, h = 128, 8
    gate = torch.arange(f * h, dtype=torch.float32).reshape(f, h)
    up = -torch.arange(f * h, dtype=torch.float32).reshape(f, h) - 1
    w = interleave_gate_up(gate, up)
    assert w.shape == (2 * f, h)
    for c in range(f):
        grp, j = divmod(c, 32)
        assert torch.equal(w[64 * grp + j], gate[c]) and torch.equal(w[64 * grp + 32 + j], up[c])
    with pytest.raises(ValueError):
        interleave_gate_up(gate[:40], up[:40])


def _tiny_vocab():
    b2u = bytes_to_unicode()
    vocab = {b2u[b]: i for i, b in enumerate(range(256))}
    merges = [("t", "h"), ("th", "e"), ("Ġ", "the")]
    for a, b in merges:
        vocab[a + b] = len(vocab)
    return vocab, merges


def test_tokenizer_from_tokenizer_json_only_with_eos_template(tmp_path):
    """Qwen3-Embedding ships a tokenizer.json whose TemplateProcessing appends <|endoftext|>; a
    directory with only that file must load, with the merges in either serialisation."""
    vocab, merges = _tiny_vocab()
    eos = len(vocab)
    for style in ("strings", "pairs"):
        d = tmp_path / style
        d.mkdir()
        (d / "tokenizer.json").write_text(json.dumps({
            "model": {"type": "BPE", "vocab": vocab,
                      "merges": [f"{a} {b}" for a, b in merges] if style == "strings" else [list(m) for m in merges]},
            "added_tokens": [{"id": eos, "content": "<|endoftext|>", "special": True}],
            "post_processor": {"type": "TemplateProcessing",
                               "single": [{"Sequence": {"id": "A", "type_id": 0}},
                                          {"SpecialToken": {"id": "<|endoftext|>", "type_id": 0}}]}}))
        tok = ByteLevelBPETokenizer.from_dir(d)
        assert tok.append_eos and tok.eos_id == eos and tok.pad_id == eos
        ids = tok.encode("the the")
        assert ids[-1] == eos and ids[:-1] == [vocab["the"], vocab["Ġthe"]]
        batch_ids, mask = tok.batch(["the", "the the the"], max_length=3)
        assert batch_ids.shape == (2, 3) and mask.tolist() == [[0, 1, 1], [1, 1, 1]]  # left padded, truncated, EOS kept
        assert batch_ids[1, -1] == eos and batch_ids[0, 0] == eos                     # pad id == <|endoftext|>
    with pytest.raises(FileNotFoundError):
        ByteLevelBPETokenizer.from_dir(tmp_path / "missing")


def test_model_type_dispatch_and_missing_model(tmp_path, monkeypatch):
    monkeypatch.setenv("LEAN_EXPLORE_MODEL_DIR", str(tmp_path))
    (tmp_path / "some-qwen").mkdir()
    (tmp_path / "some-qwen" / "config.json").write_text(json.dumps({"model_type": "qwen3"}))
    (tmp_path / "some-bert").mkdir()
    (tmp_path / "some-bert" / "config.json").write_text(json.dumps({"hidden_size": 384}))
    assert model_type_of("some-qwen") == "qwen3" and model_type_of("org/some-bert") == "bert"
    with pytest.raises(FileNotFoundError, match="not found locally"):
        model_type_of("org/definitely-not-here")


def test_clients_refuse_to_run_without_cuda():
    from lean_explore_b200.embedding_client import GpuEmbeddingClient
    from lean_explore_b200.reranker_client import DEFAULT_INSTRUCTION, GpuRerankerClient

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GpuEmbeddingClient("whatever", device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        GpuRerankerClient("whatever", device="cpu")
    assert DEFAULT_INSTRUCTION == "Find relevant Lean 4 math declarations"  # reranker_client.py:13


def test_reranker_client_batching_rule_with_a_stub_model(monkeypatch):
    """reranker_client.py:143-205: inline for <= batch_size documents, executor batches above it,
    env default for the batch size, scores in input order."""
    import asyncio

    from lean_explore_b200.reranker_client import GpuRerankerClient

    class _Tok:
        def convert_tokens_to_ids(self, t):
            return {"true": 7, "false": 9}[t]

    calls = []

    class _Model:
        tokenizer = _Tok()
        max_length = None

        def score_pairs(self, pairs, tt, tf):
            calls.append(len(pairs))
            assert (tt, tf) == (7, 9)
            return [float(len(p)) for p in pairs]

    monkeypatch.setenv("LEAN_EXPLORE_RERANKER_BATCH_SIZE", "3")
    client = GpuRerankerClient("Qwen/Qwen3-Reranker-0.6B", model=_Model())
    assert client.batch_size == 3 and client.model.max_length == 512
    docs = [f"doc {'x' * i}" for i in range(8)]
    r = asyncio.run(client.rerank("q", docs))
    assert calls == [3, 3, 2] and r.scores == [float(len(client._format_pair("q", d))) for d in docs]
    calls.clear()
    assert asyncio.run(client.rerank("q", docs[:3])).scores == r.scores[:3] and calls == [3]
    assert asyncio.run(client.rerank("q", docs, batch_size=100)).scores == r.scores
