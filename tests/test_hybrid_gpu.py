"""GPU integration test of BASELINE.json config 5 (end-to-end hybrid: encoder forward + dense
top-k + cross-encoder rerank + BM25 fusion) on a synthetic data directory laid out like the
reference's cache (engine.py:93-104): lean_explore.db, informalization_faiss.index (IVFFlat file),
*_ids_map.json, bm25_name_*/.  Models are HF-layout directories with seeded random Qwen3 weights
and a byte-level BPE vocabulary trained here (no checkpoints or Qwen vocabulary exist offline).

The oracle runs the same request with HF fp32 models, the `tokenizers` library, a numpy flat
search and the (reference-pinned) glue; the GPU pipeline must return the same declarations."""

import asyncio
import json
import sqlite3
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NAMES = ["Nat.add_comm", "Nat.add_assoc", "Nat.mul_comm", "List.map_append", "List.length_append", "Finset.sum_comm",
         "Real.sqrt_nonneg", "Point.mk", "MeasureTheory.integral_add", "Nat.Prime.two_le", "continuous_of_lipschitz",
         "Set.union_comm", "Group.mul_left_cancel", "Matrix.det_mul", "isCompact_iff_finite_subcover",
         "Polynomial.degree_add_le", "ENNReal.tsum_eq_iSup_sum", "Int.emod_emod_of_dvd", "Fin.val_add", "Prod.mk"]
WORDS = ("addition is commutative associative multiplication of natural numbers list append length map sum over finite set "
         "square root nonnegative integral additive prime at least two continuous lipschitz union group cancel determinant "
         "product compact finite subcover degree polynomial series supremum remainder divides value true false").split()


def _make_tokenizer_dir(d):
    from tokenizers import Regex, Tokenizer, normalizers, pre_tokenizers, trainers
    from tokenizers.models import BPE

    from lean_explore_b200.bpe_tokenizer import PRETOKENIZE_REGEX

    t = Tokenizer(BPE(unk_token=None, fuse_unk=False, byte_fallback=False))
    t.normalizer = normalizers.NFC()
    t.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Split(Regex(PRETOKENIZE_REGEX), behavior="isolated"),
                                               pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    text = [" ".join(WORDS)] * 6 + NAMES * 4 + ["<Instruct>: Find relevant Lean 4 math declarations\n<Query>: q\n<Document>: d"] * 4
    text += ["true", "false"] * 30  # standalone words: the vocabulary must hold the bare tokens the reranker reads
    t.train_from_iterator(text, trainers.BpeTrainer(vocab_size=700, initial_alphabet=pre_tokenizers.ByteLevel.alphabet(),
                                                    special_tokens=[], show_progress=False))
    t.model.save(str(d))
    vocab = json.loads((d / "vocab.json").read_text())
    assert "true" in vocab and "false" in vocab
    pad = len(vocab)
    (d / "tokenizer_config.json").write_text(json.dumps({
        "added_tokens_decoder": {str(pad): {"content": "<|endoftext|>", "special": True}},
        "pad_token": "<|endoftext|>", "eos_token": "<|endoftext|>", "model_max_length": 131072}))
    return t, pad + 1


def _make_model_dir(d, seed, vocab_size, sentence_transformer):
    from safetensors.torch import save_file

    from oracle import qwen3_decoder as qd

    d.mkdir(parents=True, exist_ok=True)
    ref_tok, n_vocab = _make_tokenizer_dir(d)
    model, cfg = qd.make_model("small", seed=seed, vocab_size=n_vocab)
    state = {k: v.contiguous() for k, v in model.state_dict().items() if k != "lm_head.weight"}  # tied
    save_file(state, str(d / "model.safetensors"))
    (d / "config.json").write_text(json.dumps({
        "model_type": "qwen3", "hidden_size": cfg.hidden_size, "num_hidden_layers": cfg.num_hidden_layers,
        "num_attention_heads": cfg.num_attention_heads, "num_key_value_heads": cfg.num_key_value_heads,
        "intermediate_size": cfg.intermediate_size, "head_dim": cfg.head_dim, "rms_norm_eps": cfg.rms_norm_eps,
        "rope_theta": 1e6, "vocab_size": n_vocab, "tie_word_embeddings": True}))
    if sentence_transformer:
        (d / "config_sentence_transformers.json").write_text(json.dumps({"prompts": {"query": "Instruct: find\nQuery:"}}))
    return model, ref_tok


def _left_pad(rows, pad_id):
    s = max(len(r) for r in rows)
    ids = np.full((len(rows), s), pad_id, np.int32)
    mask = np.zeros((len(rows), s), np.int32)
    for i, r in enumerate(rows):
        ids[i, s - len(r):] = r
        mask[i, s - len(r):] = 1
    return ids, mask


def test_config5_end_to_end_matches_oracle_pipeline(tmp_path, monkeypatch):
    from lean_explore_b200 import corpus as lc
    from lean_explore_b200 import hybrid as hy
    from lean_explore_b200.embedding_client import GpuEmbeddingClient
    from lean_explore_b200.reranker_client import GpuRerankerClient
    from oracle import faiss_flat as ff
    from oracle import qwen3_decoder as qd

    rng = np.random.default_rng(0)
    emb_model, emb_tok = _make_model_dir(tmp_path / "models" / "emb", 0, None, True)
    rr_model, rr_tok = _make_model_dir(tmp_path / "models" / "rerank", 5, None, False)
    monkeypatch.setenv("LEAN_EXPLORE_MODEL_DIR", str(tmp_path / "models"))
    pad = emb_tok.get_vocab_size()

    # ---- the data directory: 400 declarations, embeddings from the ORACLE embedding model
    names = [n if i < len(NAMES) else f"{n}_{i}" for i, n in ((i, NAMES[i % len(NAMES)]) for i in range(400))]
    infos = [" ".join(rng.choice(WORDS, size=int(rng.integers(4, 14)))) for _ in names]
    enc = [emb_tok.encode(t, add_special_tokens=False).ids for t in infos]
    vecs = []
    for a in range(0, len(enc), 50):
        ids, mask = _left_pad(enc[a : a + 50], pad)
        vecs.append(qd.embed(emb_model, ids, mask))
    matrix = np.concatenate(vecs).astype(np.float32)   # [400, 512]
    data = tmp_path / "cache"
    data.mkdir()
    con = sqlite3.connect(data / "lean_explore.db")
    con.execute("CREATE TABLE declarations (id INTEGER PRIMARY KEY, name TEXT, module TEXT, docstring TEXT, source_text TEXT,"
                " source_link TEXT, dependencies TEXT, informalization TEXT, informalization_embedding BLOB)")
    decl_ids = [100 + 3 * i for i in range(len(names))]
    for i, (cid, n, info) in enumerate(zip(decl_ids, names, infos)):
        deps = json.dumps([names[j] for j in rng.integers(0, len(names), size=3)])
        con.execute("INSERT INTO declarations VALUES (?,?,?,?,?,?,?,?,?)",
                    (cid, n, "Mathlib.Test", None, "theorem " + n, "https://x/" + n, deps, info, lc.pack_embedding(matrix[i].tolist())))
    con.commit()
    con.close()
    nlist = 8
    cent = matrix[rng.choice(len(matrix), nlist, replace=False)]
    lc.write_ivfflat_index(data / "informalization_faiss.index", matrix, np.argmax(matrix @ cent.T, axis=1), cent)
    (data / "informalization_faiss_ids_map.json").write_text(json.dumps(decl_ids))
    hy.Bm25Plus().index([list(set(hy.name_tokens_spaced(n))) for n in names]).save(data / "bm25_name_spaced")
    hy.Bm25Plus().index([list(set(hy.name_token_raw(n))) for n in names]).save(data / "bm25_name_raw")
    (data / "bm25_ids_map.json").write_text(json.dumps(decl_ids))

    # ---- GPU pipeline, everything loaded from disk the way the local backend would
    eng = hy.HybridSearchEngine.from_data_dir(data, embedding_client=GpuEmbeddingClient("emb", max_length=512),
                                              reranker_client=GpuRerankerClient("rerank", max_length=256))
    assert eng.semantic.faiss_informal_index.d == 512 and eng.semantic.faiss_informal_index.ntotal == 400
    queries = ["addition of natural numbers is commutative", "Nat.add_comm", "determinant of a product", "finite subcover compact"]
    got = [asyncio.run(eng.search(q, limit=10, faiss_k=200, rerank_top=20)) for q in queries]
    got_batch = asyncio.run(eng.search_batch(queries, limit=10, faiss_k=200, rerank_top=20))
    assert [[r.id for r in g] for g in got] == [[r.id for r in g] for g in got_batch]

    # ---- oracle pipeline: HF fp32 models + tokenizers library + numpy flat search + the same glue
    tt, tf = rr_tok.token_to_id("true"), rr_tok.token_to_id("false")

    class _Sem:
        async def _retrieve_semantic_candidates(self, query, k):
            ids, mask = _left_pad([emb_tok.encode("Instruct: find\nQuery:" + query, add_special_tokens=False).ids], pad)
            x = qd.embed(emb_model, ids, mask)
            ff.normalize_L2(x)
            D, I = ff.flat_ip_search(matrix, x, k)
            return eng.semantic._to_map(I[0], D[0], decl_ids)

    class _RR:
        async def rerank(self, query, documents):
            pairs = [f"<Instruct>: Find relevant Lean 4 math declarations\n<Query>: {query}\n<Document>: {d}" for d in documents]
            ids, mask = _left_pad([rr_tok.encode(p, add_special_tokens=False).ids for p in pairs], pad)
            return types.SimpleNamespace(scores=qd.rerank(rr_model, ids, mask, tt, tf).tolist())

    oracle = hy.HybridSearchEngine(_Sem(), eng.declarations, eng._bm25_name_spaced, eng._bm25_name_raw, decl_ids, _RR())
    for q, g in zip(queries, got):
        want = asyncio.run(oracle.search(q, limit=10, faiss_k=200, rerank_top=20))
        assert len(g) > 0 and all(isinstance(r, hy.SearchResult) for r in g)
        gi, wi = [r.id for r in g], [r.id for r in want]
        assert set(gi) == set(wi), (q, gi, wi)
        # order: identical unless two blended scores are within the reranker's fp16 noise
        assert sum(a != b for a, b in zip(gi, wi)) <= 2, (q, gi, wi)
        assert g[0].name and g[0].source_link.startswith("https://x/")
