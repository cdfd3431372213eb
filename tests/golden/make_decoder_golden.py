#!/usr/bin/env python
"""Generates tests/golden/decoder_golden.npz: last-token unit vectors and reranker scores of the
HF Qwen3 oracle (oracle/qwen3_decoder.py, fp32, seeded random weights - no checkpoints exist
offline and the reference holds no golden embeddings / scores) for fixed synthetic token ids.
Run from the repo root: python tests/golden/make_decoder_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import qwen3_decoder as qd  # noqa: E402

CASES = [("tiny", 5, 19, "left"), ("tiny", 3, 70, "right"), ("small", 4, 33, "left"), ("qwen3-0.6b", 3, 21, "left")]
TOKEN_TRUE, TOKEN_FALSE = 1837, 3082


def main():
    out = {}
    for geom, b, s, side in CASES:
        model, _ = qd.make_model(geom, seed=0)
        ids, mask = qd.make_inputs(b, s, seed=7, side=side)
        key = f"{geom}_{b}x{s}_{side}"
        out[key + "/emb"] = qd.embed(model, ids, mask)
        if side == "left":  # the reranker reads position -1: only meaningful under left padding
            out[key + "/score"] = qd.rerank(model, ids, mask, TOKEN_TRUE, TOKEN_FALSE)
        out[key + "/ids"] = ids
        out[key + "/mask"] = mask
    np.savez_compressed(Path(__file__).with_name("decoder_golden.npz"), **out)
    print("wrote", len(CASES))


if __name__ == "__main__":
    main()
