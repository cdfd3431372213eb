#!/usr/bin/env python
"""Generates tests/golden/hybrid_golden.json by EXECUTING the reference's own hybrid-ranking code
(/root/reference/src/lean_explore/search/{scoring,tokenization,engine}.py, unmodified) on seeded
inputs.  The reference's third-party imports that are not installable here (bm25s, sqlalchemy) and
its ORM models are replaced by minimal stand-ins below; the arithmetic that is recorded - rank
fusion, dependency boost, score normalisation, fuzzy matching, the rerank blend, result filtering -
is the reference's.  /root/reference exists only in the build container, so the output is
committed.  Run from the repo root: python tests/golden/make_hybrid_golden.py"""
import asyncio
import json
import random
import sys
import types
from pathlib import Path

REF = Path("/root/reference/src")
OUT = Path(__file__).with_name("hybrid_golden.json")


# ---- stand-ins for what cannot be imported here
class _Col:
    def in_(self, ids):
        return ("in", list(ids))

    def __eq__(self, other):
        return ("eq", other)


class Declaration:
    id = _Col()
    name = _Col()

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _install_stubs(fake_bm25_scores):
    from pydantic import BaseModel

    class SearchResult(BaseModel):
        id: int
        name: str
        module: str
        docstring: str | None
        source_text: str
        source_link: str
        dependencies: str | None
        informalization: str | None

    models = types.ModuleType("lean_explore.models")
    models.Declaration, models.SearchResult, models.SearchResponse = Declaration, SearchResult, object
    sys.modules["lean_explore.models"] = models

    bm25s = types.ModuleType("bm25s")

    class BM25:  # records nothing: returns the scores the generator chose for this candidate list
        def __init__(self, method=None):
            assert method == "bm25+"

        def index(self, tokens):
            self.n = len(tokens)

        def retrieve(self, queries, k):
            sc = fake_bm25_scores[: self.n]
            order = sorted(range(self.n), key=lambda i: -sc[i])[:k]
            return [order], [[sc[i] for i in order]]

    bm25s.BM25 = BM25
    sys.modules["bm25s"] = bm25s

    sa = types.ModuleType("sqlalchemy")

    class _Stmt:
        def where(self, cond):
            self.cond = cond
            return self

    sa.select = lambda *a: _Stmt()
    sys.modules["sqlalchemy"] = sa
    ext = types.ModuleType("sqlalchemy.ext")
    aio = types.ModuleType("sqlalchemy.ext.asyncio")

    class AsyncSession:
        def __init__(self, engine):
            self.rows = engine  # the "engine" of the stand-in is just the list of rows

        async def __aenter__(self):
            return self

        async def __aexit__(self, *a):
            return False

        async def execute(self, stmt):
            kind, ids = stmt.cond
            wanted = set(ids)
            rows = [r for r in self.rows if r.id in wanted]
            return types.SimpleNamespace(scalars=lambda: types.SimpleNamespace(all=lambda: rows))

    aio.AsyncSession, aio.AsyncEngine, aio.create_async_engine = AsyncSession, object, lambda *a, **k: None
    sys.modules["sqlalchemy.ext"], sys.modules["sqlalchemy.ext.asyncio"] = ext, aio
    cfg = types.ModuleType("lean_explore.config")
    cfg.Config = types.SimpleNamespace()
    sys.modules["lean_explore.config"] = cfg


NAMES = ["Nat.add_comm", "Nat.add_assoc", "List.map_append", "Finset.sum_comm", "Real.sqrt_nonneg", "Point.mk",
         "MeasureTheory.integral_add", "Nat.Prime.two_le", "continuous_of_lipschitz", "Set.union_comm", "Prod.mk",
         "Group.mul_left_cancel", "Matrix.det_mul", "isCompact_iff", "Polynomial.degree_add_le", "ENNReal.tsum_eq"]


def make_decls(rng, n):
    decls = []
    for i in range(n):
        base = NAMES[i % len(NAMES)] + ("" if i < len(NAMES) else f"_{i}")
        decls.append(dict(id=1000 + 7 * i, name=base, module=rng.choice(["Mathlib.Data.Nat", "Mathlib.Topology", "Std.Data", "Batteries.List", ""]),
                          docstring=rng.choice([None, "doc " + base]), source_text="theorem " + base, source_link="https://x/" + base,
                          dependencies=None, informalization=rng.choice([None, "", "The statement that " + base.replace(".", " ") + " holds"])))
    names = [d["name"] for d in decls]
    for d in decls:
        r = rng.random()
        if r < 0.15:
            d["dependencies"] = "{not json"
        elif r < 0.8:
            d["dependencies"] = json.dumps(rng.sample(names, k=rng.randint(0, min(5, n))) + ["Outside.decl"])
    return decls


def main():
    rng = random.Random(20240607)
    sys.path.insert(0, str(REF))
    fake_scores = [round(rng.random() * 9, 6) for _ in range(64)]
    _install_stubs(fake_scores)
    import lean_explore.search.scoring as scoring
    import lean_explore.search.tokenization as tok
    from lean_explore.search.engine import SearchEngine

    g = {"fake_bm25_scores": fake_scores}
    texts = ["Nat.add_comm", "List.mapAppend_nil", "isCompact_iff_finite_subcover", "", "HTTPServer.parseURL2", "a.b_c.DéjàVu", "Point.mk",
             "sum of two even numbers is even!", "Finset.sum_comm'", "x", "ℕ → ℝ continuous", "MK.mk.mk", "UPPER lower Mixed_Case.dots"]
    g["tokenize"] = [dict(text=t, spaced=tok.tokenize_spaced(t), raw=tok.tokenize_raw(t), words=tok.tokenize_words(t),
                          autogen=tok.is_autogenerated(t)) for t in texts]
    score_lists = [[], [0.0], [5.0, 5.0], [0.0, 0.0, 0.0], [1e-10, 2e-10], [3.0, -1.0, 2.5, 2.5], [rng.uniform(-3, 9) for _ in range(17)],
                   [1.0, 1.0 + 5e-10], [-2.0, -2.0]]
    g["normalize_scores"] = [dict(scores=s, out=scoring.normalize_scores(s)) for s in score_lists]
    count_lists = [[], [0, 0], [1], [0, 3, 7, 1], [rng.randint(0, 40) for _ in range(23)]]
    g["normalize_dependency_counts"] = [dict(counts=c, out=scoring.normalize_dependency_counts(c)) for c in count_lists]
    g["compute_ranks"] = [dict(scores=s, out=scoring.compute_ranks(s)) for s in score_lists + [[0.5, 1.0, 0.3, 0.8], [1.0, 0.0, 0.5]]]
    rank_sets = [[[1, 2, 3], [3, 1, 2]], [[rng.randint(1, 30) for _ in range(12)] for _ in range(3)]]
    g["rrf_lists"] = [dict(ranks=r, k=k, out=scoring.reciprocal_rank_fusion(r, k=k)) for r in rank_sets for k in (0, 60)]
    fusion_sets = [([[0.0, 0.5, 1.0], [1.0, 0.5, 0.0]], [0.5, 0.5]), ([[3.0, 3.0], [1.0, 2.0]], [0.7, 0.3]), ([], []), ([[], []], [1.0, 1.0]),
                   ([[rng.uniform(0, 5) for _ in range(9)] for _ in range(3)], [1.0, 0.4, 0.2])]
    g["weighted_fusion"] = [dict(scores=sl, weights=w, out=scoring.weighted_score_fusion(sl, w)) for sl, w in fusion_sets]
    pairs = [("add comm", "Nat.add_comm"), ("Nat.add_comm", "Nat.add_comm"), ("prime two", "Nat.Prime.two_le"), ("", "x"), ("", ""),
             ("continuous lipschitz", "continuous_of_lipschitz"), ("SUM_COMM", "Finset.sum_comm"), ("det mul", "Matrix.det_mul"),
             ("nat add comm", "Nat.add_comm"), ("list map", "Set.union_comm")]
    g["fuzzy"] = [dict(query=q, name=n, out=scoring.fuzzy_name_score(q, n)) for q, n in pairs]

    # ---- rank fusion (engine.py:263-300)
    g["rrf"] = []
    for case in range(8):
        nb, ns = rng.randint(0, 40), rng.randint(0, 40)
        ids = rng.sample(range(1, 500), 70)
        bm = {i: round(rng.uniform(0.1, 12), 3) for i in rng.sample(ids, nb)}
        sm = {i: round(rng.uniform(-0.2, 1), 4) for i in rng.sample(ids, ns)}
        if case == 3:  # ties inside both signals
            bm = {i: 2.0 for i in bm}
            sm = {i: 0.5 for i in sm}
        out = SearchEngine._compute_rrf_scores(None, bm, sm)
        g["rrf"].append(dict(bm25=[[k, v] for k, v in bm.items()], semantic=[[k, v] for k, v in sm.items()], out=[[c, s] for c, s in out]))

    # ---- dependency boost (engine.py:302-358), candidate counts (:451-476), rerank blend (:360-416),
    # result filtering (:468-487) and the whole search() control flow (:534-583)
    g["boost"], g["rerank"], g["search"] = [], [], []
    for case in range(6):
        n = rng.randint(3, 40)
        decls = make_decls(rng, n)
        rows = [Declaration(**d) for d in decls]
        ids = [d["id"] for d in decls]
        rng.shuffle(ids)
        rrf = sorted(((cid, round(rng.uniform(0.01, 2), 5)) for cid in ids + [99991, 99992][: case % 3]), key=lambda x: -x[1])
        eng = SearchEngine.__new__(SearchEngine)
        eng.engine = rows
        top_n = [500, 500, 10, 500, 7, 500][case]
        boosted, dmap = asyncio.run(eng._apply_dependency_boost(rrf, top_n=top_n))
        g["boost"].append(dict(decls=decls, rrf=[[c, s] for c, s in rrf], top_n=top_n, out=[[c, s] for c, s in boosted],
                               fetched=sorted(dmap)))
        # rerank
        cand = [(r, 0.0) for r in rows[: rng.randint(1, min(n, 25))]]
        rr = [round(rng.random(), 6) for _ in cand]
        if case == 2:
            rr = [0.5] * len(cand)
        query = ["add comm", "Nat.add_comm", "prime", "sum_comm", "compact", "Matrix.det_mul"][case]

        class _RR:
            async def rerank(self, q, documents):
                self.documents = documents
                return types.SimpleNamespace(scores=rr)

        eng._reranker_client = _RR()
        limit = [50, 3, 50, 5, 50, 2][case]
        res = asyncio.run(eng._rerank_candidates(query, cand, limit))
        g["rerank"].append(dict(decls_of_boost_case=case, candidate_ids=[r.id for r, _ in cand], reranker_scores=rr, query=query, limit=limit,
                                documents=eng._reranker_client.documents,
                                dep_counts=eng._compute_candidate_dependency_counts(cand),
                                bm25_scores=eng._compute_bm25_on_informalizations(query, cand),
                                out_ids=[r.id for r in res]))
        # search(): candidate maps are injected where the reference retrieves them
        bm = {i: round(rng.uniform(0.1, 12), 3) for i in rng.sample(ids, rng.randint(0, n))}
        sm = {i: round(rng.uniform(0, 1), 4) for i in rng.sample(ids, rng.randint(1, n))}
        eng._retrieve_bm25_candidates = lambda q, k, bm=bm: bm

        async def _sem(q, k, sm=sm):
            return sm

        eng._retrieve_semantic_candidates = _sem
        class _RRdoc:  # search(): score by a fixed rule of the document text (tests use the same rule)
            async def rerank(self, q, documents):
                import zlib

                return types.SimpleNamespace(scores=[(zlib.crc32(d.encode()) % 1000) / 1000.0 for d in documents])

        eng._reranker_client = _RRdoc()
        for rerank_top, packages, lim in ((None, None, 50), (0, ["Mathlib"], 4), (5, None, 3), (25, ["Std", "Batteries"], 50)):
            res = asyncio.run(eng.search(query, limit=lim, rerank_top=rerank_top, packages=packages))
            g["search"].append(dict(decls_of_boost_case=case, bm25=[[k, v] for k, v in bm.items()], semantic=[[k, v] for k, v in sm.items()],
                                    query=query, limit=lim, rerank_top=rerank_top, packages=packages,
                                    out_ids=[r.id for r in res], first=res[0].model_dump() if res else None))
    g["search"].append(dict(empty_query=asyncio.run(SearchEngine.__new__(SearchEngine).search("   "))))
    OUT.write_text(json.dumps(g, ensure_ascii=False, indent=0))
    print("wrote", OUT, {k: len(v) for k, v in g.items()})


if __name__ == "__main__":
    main()
