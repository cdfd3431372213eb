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


# ---------------------------------------------------------------------------------------
# The reference's own client tests, replayed against the drop-in classes with a stub model
# (reference tests/util/reranker_client_test.py:19-185, tests/util/embedding_client_test.py:15-118
# mock the third-party model the same way).

class _StubTok:
    def convert_tokens_to_ids(self, t):
        return 1 if t == "true" else 0


class _StubDecoder:
    tokenizer = _StubTok()
    max_length = None

    def score_pairs(self, pairs, tt, tf):
        return [0.5 for _ in pairs]


def _reranker(**kw):
    from lean_explore_b200.reranker_client import GpuRerankerClient

    return GpuRerankerClient(model_name="test-model", model=_StubDecoder(), **kw)


def test_reference_reranker_response_and_instruction_cases():
    import asyncio

    from lean_explore_b200.reranker_client import DEFAULT_INSTRUCTION, RerankerResponse

    response = RerankerResponse(query="test query", scores=[0.9, 0.7, 0.3], model="test-model")
    assert response.query == "test query" and response.scores == [0.9, 0.7, 0.3] and response.model == "test-model"
    assert _reranker().instruction == DEFAULT_INSTRUCTION                      # test_default_instruction
    assert _reranker(instruction="Custom instruction").instruction == "Custom instruction"  # test_custom_instruction
    result = _reranker()._format_pair("search query", "document text")          # test_format_pair
    assert "<Instruct>:" in result and "<Query>: search query" in result and "<Document>: document text" in result
    assert DEFAULT_INSTRUCTION in result                                        # test_format_pair_includes_instruction
    client = _reranker()
    empty = asyncio.run(client.rerank("query", []))                            # test_rerank_empty_documents
    assert isinstance(empty, RerankerResponse) and empty.scores == [] and empty.query == "query"
    assert client.rerank_sync("query", []).scores == []                        # test_rerank_sync_empty_documents
    got = asyncio.run(client.rerank("query", ["doc1", "doc2"]))                # test_rerank_returns_response
    assert isinstance(got, RerankerResponse) and got.model == "test-model" and len(got.scores) == 2
    assert client._token_true_id == 1 and client._token_false_id == 0          # reranker_client.py:85-86


def test_reference_embedding_client_cases(monkeypatch):
    """EmbeddingResponse fields; max_length / batch size plumbing; is_query reaches the encoder as the
    query-prompt switch (the reference passes prompt_name="query", embedding_client.py:97-98)."""
    import asyncio

    import lean_explore_b200.decoder as dec_mod
    import lean_explore_b200.encoder as enc_mod
    from lean_explore_b200.embedding_client import EmbeddingResponse, GpuEmbeddingClient

    response = EmbeddingResponse(texts=["hello", "world"], embeddings=[[0.1, 0.2], [0.3, 0.4]], model="test-model")
    assert response.texts == ["hello", "world"] and len(response.embeddings) == 2 and response.model == "test-model"

    calls = []

    class _Model:
        def __init__(self, max_length):
            self.max_length = max_length

        def encode(self, texts, batch_size=8, is_query=False):
            calls.append((list(texts), batch_size, is_query))
            return np.arange(len(texts) * 4, dtype=np.float32).reshape(len(texts), 4)

    monkeypatch.setattr(dec_mod, "model_type_of", lambda name: "bert")
    monkeypatch.setattr(enc_mod, "load_sentence_encoder", lambda name, device=None, max_length=None: _Model(max_length))
    monkeypatch.setenv("LEAN_EXPLORE_EMBEDDING_BATCH_SIZE", "5")
    client = GpuEmbeddingClient(model_name="test-model", max_length=256)       # test_max_length_setting
    assert client.model.max_length == 256 and client.batch_size == 5 and client.device == "cuda"
    got = asyncio.run(client.embed(["hello", "world"]))                        # test_embed_returns_response
    assert isinstance(got, EmbeddingResponse) and len(got.embeddings) == 2 and got.model == "test-model"
    assert calls[-1] == (["hello", "world"], 5, False)                         # test_embed_without_query_flag
    asyncio.run(client.embed(["search query"], is_query=True))                 # test_embed_with_query_flag
    assert calls[-1] == (["search query"], 5, True)
    assert client.embed_array(["a", "b", "c"], is_query=True).shape == (3, 4)
    assert GpuEmbeddingClient(model_name="test-model", batch_size=2).batch_size == 2
