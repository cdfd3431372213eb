#!/usr/bin/env python
"""Generates tests/golden/encoder_golden.npz: pooled unit vectors of the HF-BertModel oracle
(oracle/bert_encoder.py, fp32, seeded random weights - no checkpoints exist offline and the
reference holds no golden embeddings) for fixed synthetic token ids.  Run from the repo root:
python tests/golden/make_encoder_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import bert_encoder as be  # noqa: E402

CASES = [("tiny", 5, 19, "mean"), ("tiny", 3, 8, "cls"), ("minilm-l6", 4, 24, "mean")]


def main():
    out = {}
    for geom, b, s, pool in CASES:
        model, _ = be.make_model(geom, seed=0)
        ids, mask = be.make_inputs(b, s, seed=7)
        out[f"{geom}_{b}x{s}_{pool}/emb"] = be.encode(model, ids, mask, pool)
        out[f"{geom}_{b}x{s}_{pool}/ids"] = ids
        out[f"{geom}_{b}x{s}_{pool}/mask"] = mask
    np.savez_compressed(Path(__file__).with_name("encoder_golden.npz"), **out)
    print("wrote", len(CASES))


if __name__ == "__main__":
    main()
