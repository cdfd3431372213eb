"""BERT-class sentence encoder on the B200 kernels (``lxg_encode``), host side.

Replaces what ``SentenceTransformer(model_name)`` / ``.encode`` do for the reference's
``EmbeddingClient`` (``src/lean_explore/util/embedding_client.py:58,99``): load a
sentence-transformers model directory (HF ``config.json`` + weights + ``vocab.txt`` + the
``1_Pooling/config.json`` that selects mean / CLS pooling), tokenise on the host, run
transformer → pooling → L2-normalise on the GPU, return ``float32 [B, H]``.

PyTorch only owns the weight tensors in HBM; all arithmetic is ``liblxg.so``.
"""

from __future__ import annotations

import ctypes
import json
import os
from ctypes import c_void_p
from pathlib import Path

import numpy as np
import torch

from . import _lib
from .tokenizer import WordPieceTokenizer

POOL_MEAN, POOL_CLS = _lib.LXG_POOL_MEAN, _lib.LXG_POOL_CLS


class BertSentenceEncoder:
    """Weights of one BERT-class encoder resident in HBM + the ``lxg_encoder`` handle."""

    def __init__(self, state: dict[str, torch.Tensor], *, hidden: int, layers: int, heads: int, ffn: int,
                 ln_eps: float = 1e-12, pool: int = POOL_MEAN, device: int = 0,
                 tokenizer: WordPieceTokenizer | None = None, max_length: int | None = None,
                 query_prompt: str = ""):
        self.device = int(device)
        self.hidden, self.layers, self.heads, self.ffn = hidden, layers, heads, ffn
        self.pool = pool
        self.tokenizer = tokenizer
        self.max_length = max_length
        self.query_prompt = query_prompt
        self._lib = _lib.init(self.device)
        dev = torch.device("cuda", self.device)
        state = {k[5:] if k.startswith("bert.") else k: v for k, v in state.items()}
        self._keep: list[torch.Tensor] = []

        def mat(name):  # matrices: fp16
            t = state[name].to(device=dev, dtype=torch.float16).contiguous()
            self._keep.append(t)
            return t

        def vec(name):  # vectors: fp32
            t = state[name].to(device=dev, dtype=torch.float32).contiguous()
            self._keep.append(t)
            return t

        word = mat("embeddings.word_embeddings.weight")
        pos = mat("embeddings.position_embeddings.weight")
        typ = mat("embeddings.token_type_embeddings.weight")
        self.vocab, self.max_pos = word.shape[0], pos.shape[0]
        arr = (_lib.BertLayer * layers)()
        for i in range(layers):
            p = f"encoder.layer.{i}."
            wqkv = torch.cat([state[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value")], dim=0)
            bqkv = torch.cat([state[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value")], dim=0)
            wqkv = wqkv.to(device=dev, dtype=torch.float16).contiguous()
            bqkv = bqkv.to(device=dev, dtype=torch.float32).contiguous()
            self._keep += [wqkv, bqkv]
            L = arr[i]
            L.wqkv, L.bqkv = wqkv.data_ptr(), bqkv.data_ptr()
            L.wo, L.bo = mat(p + "attention.output.dense.weight").data_ptr(), vec(p + "attention.output.dense.bias").data_ptr()
            L.ln1_g, L.ln1_b = (vec(p + "attention.output.LayerNorm.weight").data_ptr(),
                                vec(p + "attention.output.LayerNorm.bias").data_ptr())
            L.w1, L.b1 = mat(p + "intermediate.dense.weight").data_ptr(), vec(p + "intermediate.dense.bias").data_ptr()
            L.w2, L.b2 = mat(p + "output.dense.weight").data_ptr(), vec(p + "output.dense.bias").data_ptr()
            L.ln2_g, L.ln2_b = vec(p + "output.LayerNorm.weight").data_ptr(), vec(p + "output.LayerNorm.bias").data_ptr()
        w = _lib.BertWeights()
        w.hidden, w.layers, w.heads, w.ffn, w.vocab, w.max_pos = hidden, layers, heads, ffn, self.vocab, self.max_pos
        w.ln_eps = ln_eps
        w.word_emb, w.pos_emb, w.type_emb = word.data_ptr(), pos.data_ptr(), typ.data_ptr()
        w.emb_ln_g = vec("embeddings.LayerNorm.weight").data_ptr()
        w.emb_ln_b = vec("embeddings.LayerNorm.bias").data_ptr()
        w.layer = arr
        self._arr = arr
        torch.cuda.synchronize(self.device)
        h = c_void_p()
        _lib.check(self._lib.lxg_encoder_create(ctypes.byref(h), ctypes.byref(w)))
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.lxg_encoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ token ids in, vectors out
    def encode_ids(self, input_ids: np.ndarray, attention_mask: np.ndarray, pool: int | None = None) -> np.ndarray:
        """int32 [B, S] ids / mask (host) -> float32 [B, H] unit vectors (host)."""
        ids = np.ascontiguousarray(input_ids, dtype=np.int32)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int32)
        if ids.ndim != 2 or ids.shape != mask.shape:
            raise ValueError("input_ids and attention_mask must be [B, S] and agree")
        out = np.empty((ids.shape[0], self.hidden), dtype=np.float32)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.lxg_encode(self._handle, ids.ctypes.data, mask.ctypes.data, ids.shape[0], ids.shape[1],
                                        self.pool if pool is None else pool, out.ctypes.data, stream))
        return out

    def encode_ids_torch(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pool: int | None = None):
        """Device-resident variant (int32 CUDA tensors in, float32 CUDA tensor out, asynchronous)."""
        ids = input_ids.to(torch.int32).contiguous()
        mask = attention_mask.to(torch.int32).contiguous()
        out = torch.empty((ids.shape[0], self.hidden), dtype=torch.float32, device=ids.device)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.lxg_encode(self._handle, ids.data_ptr(), mask.data_ptr(), ids.shape[0], ids.shape[1],
                                        self.pool if pool is None else pool, out.data_ptr(), stream))
        return out

    def last_launches(self) -> int:
        return int(self._lib.lxg_encoder_last_launches(self._handle))

    def set_fused(self, enabled: bool) -> None:
        """Calls of <= 64 tokens (search queries) run as one persistent kernel; False keeps this
        encoder on the layered kernels."""
        _lib.check(self._lib.lxg_encoder_set_fused(self._handle, int(enabled)))

    def read_trace(self) -> np.ndarray:
        """After ``set_fused(2)`` and a query-path call: uint64 [grid, phases, 6] %globaltimer stamps (ns)."""
        cap = 256 * 260 * 6
        buf = np.zeros(cap, dtype=np.uint64)
        grid, phases = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self._lib.lxg_encoder_read_trace(self._handle, buf.ctypes.data, cap, ctypes.byref(grid), ctypes.byref(phases)))
        return buf[: grid.value * phases.value * 6].reshape(grid.value, phases.value, 6)

    # ------------------------------------------------------------------ text in, vectors out
    def encode(self, texts: list[str], batch_size: int = 8, is_query: bool = False) -> np.ndarray:
        """``SentenceTransformer.encode(texts, batch_size=..., prompt_name="query" if is_query)``:
        sorted by length like sentence-transformers, batches of ``batch_size``, original order
        restored.  Returns float32 [len(texts), H]."""
        if self.tokenizer is None:
            raise RuntimeError("this encoder was built without a tokenizer; use encode_ids")
        if is_query and self.query_prompt:
            texts = [self.query_prompt + t for t in texts]
        out = np.empty((len(texts), self.hidden), dtype=np.float32)
        order = sorted(range(len(texts)), key=lambda i: -len(texts[i]))
        for b0 in range(0, len(order), max(1, batch_size)):
            idx = order[b0 : b0 + batch_size]
            ids, mask = self.tokenizer.batch([texts[i] for i in idx], self.max_length)
            out[idx] = self.encode_ids(ids, mask)
        return out


def _resolve_model_dir(model_name: str) -> Path:
    cands = [Path(model_name)]
    root = os.getenv("LEAN_EXPLORE_MODEL_DIR")
    if root:
        cands += [Path(root) / model_name, Path(root) / model_name.split("/")[-1]]
    for c in cands:
        if (c / "config.json").exists():
            return c
    try:  # an already populated HF cache (never downloads: there may be no network)
        from huggingface_hub import snapshot_download

        return Path(snapshot_download(model_name, local_files_only=True))
    except Exception as e:  # noqa: BLE001
        raise FileNotFoundError(
            f"model {model_name!r} not found locally (looked in {[str(c) for c in cands]} and the HF cache); "
            "set LEAN_EXPLORE_MODEL_DIR to a directory holding the sentence-transformers model"
        ) from e


def load_sentence_encoder(model_name: str, device: str | int = "cuda", max_length: int | None = None) -> BertSentenceEncoder:
    """Load a sentence-transformers BERT-class model directory (all-MiniLM-L6-v2, bge-base-en-v1.5, ...)."""
    d = _resolve_model_dir(model_name)
    cfg = json.loads((d / "config.json").read_text())
    if cfg.get("model_type", "bert") != "bert":
        raise NotImplementedError(f"model_type {cfg.get('model_type')!r}: only BERT-class encoders are built (DESIGN.md section 8)")
    if (d / "model.safetensors").exists():
        from safetensors.torch import load_file

        state = load_file(str(d / "model.safetensors"))
    else:
        state = torch.load(d / "pytorch_model.bin", map_location="cpu", weights_only=True)
    pool = POOL_MEAN
    pc = d / "1_Pooling" / "config.json"
    if pc.exists() and json.loads(pc.read_text()).get("pooling_mode_cls_token"):
        pool = POOL_CLS
    lower = True
    tc = d / "tokenizer_config.json"
    if tc.exists():
        lower = bool(json.loads(tc.read_text()).get("do_lower_case", True))
    st_len = None
    sb = d / "sentence_bert_config.json"
    if sb.exists():
        st_len = json.loads(sb.read_text()).get("max_seq_length")
    prompt = ""
    cs = d / "config_sentence_transformers.json"
    if cs.exists():
        prompt = (json.loads(cs.read_text()).get("prompts") or {}).get("query", "") or ""
    tok = WordPieceTokenizer.from_vocab_file(d / "vocab.txt", do_lower_case=lower,
                                             model_max_length=int(cfg.get("max_position_embeddings", 512)))
    dev = 0 if isinstance(device, str) and ":" not in device else int(str(device).split(":")[-1])
    return BertSentenceEncoder(state, hidden=cfg["hidden_size"], layers=cfg["num_hidden_layers"],
                               heads=cfg["num_attention_heads"], ffn=cfg["intermediate_size"],
                               ln_eps=float(cfg.get("layer_norm_eps", 1e-12)), pool=pool, device=dev, tokenizer=tok,
                               max_length=max_length or st_len, query_prompt=prompt)
