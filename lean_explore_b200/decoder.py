"""Qwen3-class decoder backbone on the B200 kernels (``lxg_decoder_embed`` / ``lxg_decoder_rerank``),
host side.

Replaces, for the two models the reference ships with,

* ``SentenceTransformer("Qwen/Qwen3-Embedding-0.6B").encode`` inside ``EmbeddingClient.embed``
  (``src/lean_explore/util/embedding_client.py:58,88-101``): transformer -> last-token pooling ->
  L2 normalise, query prompt from ``config_sentence_transformers.json`` when ``is_query``;
* ``AutoModelForCausalLM.from_pretrained("Qwen/Qwen3-Reranker-0.6B")`` and the last-token
  ``true`` / ``false`` softmax of ``RerankerClient._compute_scores_sync``
  (``src/lean_explore/util/reranker_client.py:71-87,110-141``).

PyTorch only owns the weight tensors in HBM; all arithmetic is ``liblxg.so``.
"""

from __future__ import annotations

import ctypes
import json
from ctypes import c_void_p
from pathlib import Path

import numpy as np
import torch

from . import _lib

HEAD_DIM = 128


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[F, H] gate_proj and up_proj -> [2F, H] with rows in groups of 32: gate[0:32], up[0:32],
    gate[32:64], up[32:64], ... (the layout ``lxg_qwen3_layer.wgu`` documents: one epilogue thread
    of the GEMM then holds gate and up of the same output column)."""
    f, h = gate.shape
    if f % 32:
        raise ValueError("ffn size must be a multiple of 32")
    return torch.stack([gate.reshape(f // 32, 32, h), up.reshape(f // 32, 32, h)], dim=1).reshape(2 * f, h)


class Qwen3Decoder:
    """Weights of one Qwen3-class model resident in HBM + the ``lxg_decoder`` handle."""

    def __init__(self, state: dict[str, torch.Tensor], *, hidden: int, layers: int, heads: int, kv_heads: int,
                 ffn: int, head_dim: int = HEAD_DIM, rms_eps: float = 1e-6, rope_theta: float = 1e6,
                 device: int = 0, tokenizer=None, max_length: int | None = None, query_prompt: str = "",
                 with_lm_head: bool = True):
        self.device = int(device)
        self.hidden, self.layers, self.heads, self.kv_heads, self.ffn = hidden, layers, heads, kv_heads, ffn
        self.tokenizer, self.max_length, self.query_prompt = tokenizer, max_length, query_prompt
        self._lib = _lib.init(self.device)
        dev = torch.device("cuda", self.device)
        state = {k[6:] if k.startswith("model.") else k: v for k, v in state.items()}
        self._keep: list[torch.Tensor] = []

        def put(t: torch.Tensor, dtype) -> torch.Tensor:
            t = t.to(device=dev, dtype=dtype).contiguous()
            self._keep.append(t)
            return t

        tok = put(state["embed_tokens.weight"], torch.float16)
        self.vocab = tok.shape[0]
        lm = None
        if with_lm_head:
            lm = put(state["lm_head.weight"], torch.float16) if "lm_head.weight" in state else tok  # tied
        # Qwen3RotaryEmbedding (default rope): inv_freq = 1 / theta^(arange(0, dim, 2) / dim), fp32
        inv_freq = 1.0 / (rope_theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).to(torch.float32) / head_dim))
        inv_freq = put(inv_freq, torch.float32)
        arr = (_lib.Qwen3Layer * layers)()
        for i in range(layers):
            p = f"layers.{i}."
            L = arr[i]
            L.ln1 = put(state[p + "input_layernorm.weight"], torch.float32).data_ptr()
            wqkv = torch.cat([state[p + f"self_attn.{n}_proj.weight"] for n in ("q", "k", "v")], dim=0)
            L.wqkv = put(wqkv, torch.float16).data_ptr()
            L.q_norm = put(state[p + "self_attn.q_norm.weight"], torch.float32).data_ptr()
            L.k_norm = put(state[p + "self_attn.k_norm.weight"], torch.float32).data_ptr()
            L.wo = put(state[p + "self_attn.o_proj.weight"], torch.float16).data_ptr()
            L.ln2 = put(state[p + "post_attention_layernorm.weight"], torch.float32).data_ptr()
            L.wgu = put(interleave_gate_up(state[p + "mlp.gate_proj.weight"], state[p + "mlp.up_proj.weight"]),
                        torch.float16).data_ptr()
            L.wdown = put(state[p + "mlp.down_proj.weight"], torch.float16).data_ptr()
        w = _lib.Qwen3Weights()
        w.hidden, w.layers, w.heads, w.kv_heads, w.head_dim, w.ffn, w.vocab = (
            hidden, layers, heads, kv_heads, head_dim, ffn, self.vocab)
        w.rms_eps = rms_eps
        w.tok_emb = tok.data_ptr()
        w.lm_head = lm.data_ptr() if lm is not None else None
        w.final_norm = put(state["norm.weight"], torch.float32).data_ptr()
        w.inv_freq = inv_freq.data_ptr()
        w.layer = arr
        self._arr = arr
        torch.cuda.synchronize(self.device)
        h = c_void_p()
        _lib.check(self._lib.lxg_decoder_create(ctypes.byref(h), ctypes.byref(w)))
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.lxg_decoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def last_launches(self) -> int:
        return int(self._lib.lxg_decoder_last_launches(self._handle))

    def last_tokens(self) -> int:
        """Tokens the last forward computed (padding tokens are dropped for host batches)."""
        return int(self._lib.lxg_decoder_last_tokens(self._handle))

    @staticmethod
    def _prep(input_ids, attention_mask):
        ids = np.ascontiguousarray(input_ids, dtype=np.int32)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int32)
        if ids.ndim != 2 or ids.shape != mask.shape:
            raise ValueError("input_ids and attention_mask must be [B, S] and agree")
        return ids, mask

    # ------------------------------------------------------------------ token ids in
    def embed_ids(self, input_ids: np.ndarray, attention_mask: np.ndarray) -> np.ndarray:
        """int32 [B, S] ids / mask (host) -> float32 [B, H] unit vectors (last-token pooling)."""
        ids, mask = self._prep(input_ids, attention_mask)
        out = np.empty((ids.shape[0], self.hidden), dtype=np.float32)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.lxg_decoder_embed(self._handle, ids.ctypes.data, mask.ctypes.data, ids.shape[0],
                                               ids.shape[1], out.ctypes.data, stream))
        return out

    def embed_ids_torch(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        ids = input_ids.to(torch.int32).contiguous()
        mask = attention_mask.to(torch.int32).contiguous()
        out = torch.empty((ids.shape[0], self.hidden), dtype=torch.float32, device=ids.device)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.lxg_decoder_embed(self._handle, ids.data_ptr(), mask.data_ptr(), ids.shape[0],
                                               ids.shape[1], out.data_ptr(), stream))
        return out

    def rerank_ids(self, input_ids: np.ndarray, attention_mask: np.ndarray, token_true: int,
                   token_false: int) -> np.ndarray:
        """int32 [B, S] ids / mask -> float32 [B] P("true") over {false, true} at the last position."""
        ids, mask = self._prep(input_ids, attention_mask)
        out = np.empty((ids.shape[0],), dtype=np.float32)
        stream = int(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._lib.lxg_decoder_rerank(self._handle, ids.ctypes.data, mask.ctypes.data, ids.shape[0],
                                                ids.shape[1], int(token_true), int(token_false), out.ctypes.data,
                                                stream))
        return out

    # ------------------------------------------------------------------ text in
    def _batch(self, texts: list[str]):
        if self.tokenizer is None:
            raise RuntimeError("this decoder was built without a tokenizer; use the *_ids entry points")
        return self.tokenizer.batch(texts, self.max_length)

    def encode(self, texts: list[str], batch_size: int = 8, is_query: bool = False) -> np.ndarray:
        """``SentenceTransformer.encode(texts, batch_size=..., prompt_name="query" if is_query)``:
        sorted by length, batches of ``batch_size``, original order restored; float32 [n, H]."""
        if is_query and self.query_prompt:
            texts = [self.query_prompt + t for t in texts]
        out = np.empty((len(texts), self.hidden), dtype=np.float32)
        order = sorted(range(len(texts)), key=lambda i: -len(texts[i]))
        for b0 in range(0, len(order), max(1, batch_size)):
            idx = order[b0 : b0 + batch_size]
            ids, mask = self._batch([texts[i] for i in idx])
            out[idx] = self.embed_ids(ids, mask)
        return out

    def score_pairs(self, pairs: list[str], token_true: int, token_false: int) -> list[float]:
        """``RerankerClient._compute_scores_sync`` (``reranker_client.py:110-141``) on one batch."""
        ids, mask = self._batch(pairs)
        return self.rerank_ids(ids, mask, token_true, token_false).tolist()


def load_qwen3(model_name: str, device: str | int = "cuda", max_length: int | None = None,
               with_lm_head: bool = True) -> Qwen3Decoder:
    """Load a Qwen3 model directory (HF layout: config.json, model.safetensors, vocab.json +
    merges.txt [+ tokenizer_config.json], optionally the sentence-transformers side files)."""
    from .bpe_tokenizer import ByteLevelBPETokenizer
    from .encoder import _resolve_model_dir

    d = _resolve_model_dir(model_name)
    cfg = json.loads((d / "config.json").read_text())
    if cfg.get("model_type") != "qwen3":
        raise NotImplementedError(f"model_type {cfg.get('model_type')!r} is not a Qwen3 decoder")
    state: dict[str, torch.Tensor] = {}
    files = sorted(d.glob("model*.safetensors"))
    if files:
        from safetensors.torch import load_file

        for f in files:
            state.update(load_file(str(f)))
    else:
        state = torch.load(d / "pytorch_model.bin", map_location="cpu", weights_only=True)
    prompt, st_len = "", None
    cs = d / "config_sentence_transformers.json"
    if cs.exists():
        prompt = (json.loads(cs.read_text()).get("prompts") or {}).get("query", "") or ""
    sb = d / "sentence_bert_config.json"
    if sb.exists():
        st_len = json.loads(sb.read_text()).get("max_seq_length")
    tok = ByteLevelBPETokenizer.from_dir(d)
    dev = 0 if isinstance(device, str) and ":" not in device else int(str(device).split(":")[-1])
    rope = cfg.get("rope_theta") or (cfg.get("rope_parameters") or {}).get("rope_theta", 1e6)
    return Qwen3Decoder(state, hidden=cfg["hidden_size"], layers=cfg["num_hidden_layers"],
                        heads=cfg["num_attention_heads"], kv_heads=cfg["num_key_value_heads"],
                        ffn=cfg["intermediate_size"], head_dim=cfg.get("head_dim", HEAD_DIM),
                        rms_eps=float(cfg.get("rms_norm_eps", 1e-6)), rope_theta=float(rope), device=dev,
                        tokenizer=tok, max_length=max_length or st_len, query_prompt=prompt,
                        with_lm_head=with_lm_head)


def model_type_of(model_name: str) -> str:
    from .encoder import _resolve_model_dir

    return json.loads((Path(_resolve_model_dir(model_name)) / "config.json").read_text()).get("model_type", "bert")
