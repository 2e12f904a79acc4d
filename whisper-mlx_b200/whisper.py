"""Model container: host mirror of `mlx_whisper/whisper.py` (UPSTREAM; reached from
/root/reference/run:3-6; restated in SURVEY.md A.2).

`ModelDimensions` and the public surface of `Whisper` (dims, is_multilingual, num_languages,
embed_audio / encoder, logits, decode, detect_language) keep the reference names.  The layer loops
themselves run in the native engine (csrc/api.cu) over hand-written sm_100a kernels; this module
owns the weights as torch CUDA tensors and hands raw pointers across the C ABI.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .audio import N_FRAMES


@dataclass
class ModelDimensions:
    n_mels: int
    n_audio_ctx: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_vocab: int
    n_text_ctx: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """Returns sinusoids for positional embedding: [sin | cos] halves."""
    assert channels % 2 == 0
    log_timescale_increment = math.log(max_timescale) / (channels // 2 - 1)
    inv_timescales = torch.exp(-log_timescale_increment * torch.arange(channels // 2, dtype=torch.float64))
    scaled_time = torch.arange(length, dtype=torch.float64)[:, None] * inv_timescales[None, :]
    return torch.cat([torch.sin(scaled_time), torch.cos(scaled_time)], dim=1).to(torch.float32)


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


class Whisper:
    """Whisper weights resident in HBM + the native engine handle."""

    PAGE_SIZE = 16

    def __init__(self, dims: ModelDimensions, weights: Dict[str, torch.Tensor], device=None,
                 dtype: torch.dtype = torch.bfloat16):
        if not torch.cuda.is_available():
            raise RuntimeError("Whisper needs a CUDA device (B200); there is no CPU fallback")
        if dtype not in (torch.bfloat16, torch.float16, torch.float32):
            raise ValueError(f"unsupported dtype {dtype}")
        # the engine computes in bf16 storage / fp32 accumulation whatever `dtype` the caller asks for
        self.dims = dims
        self.dtype = torch.bfloat16
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._keep: List[torch.Tensor] = []
        self._handle = _lib.vp(0)
        import os

        # decode batches of >= 16 windows are split over this many concurrent streams (1 disables)
        # (measured on B200, r01: no gain yet -- the 60 K-register GEMM CTAs cannot co-reside with the cross-attention
        # CTAs, so the two streams serialise; default off)
        self.decode_streams = int(os.environ.get("B200W_DECODE_STREAMS", "1"))
        self._build(weights)
        # alignment heads default: all heads of the last half of the decoder layers (unused by ./run)
        all_heads = np.zeros((dims.n_text_layer, dims.n_text_head), dtype=bool)
        all_heads[dims.n_text_layer // 2:] = True
        self.alignment_heads = np.asarray(all_heads.nonzero()).T

    # ------------------------------------------------------------------ weights
    def _dev(self, t: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
        out = t.detach().to(device=self.device, dtype=dt).contiguous()
        self._keep.append(out)
        return out

    def _build(self, w: Dict[str, torch.Tensor]) -> None:
        dm = self.dims
        bf, f32 = torch.bfloat16, torch.float32
        P = _lib.ptr

        def attn_fused(prefix, d):
            wq, wk, wv = w[prefix + ".query.weight"], w[prefix + ".key.weight"], w[prefix + ".value.weight"]
            bq, bv = w[prefix + ".query.bias"], w[prefix + ".value.bias"]
            zero = torch.zeros(d, dtype=bq.dtype, device=bq.device)
            return (self._dev(torch.cat([wq, wk, wv], 0), bf), self._dev(torch.cat([bq.float(), zero.float(), bv.float()], 0), f32))

        d = dm.n_audio_state
        enc_layers = (_lib.EncLayer * dm.n_audio_layer)()
        for i in range(dm.n_audio_layer):
            p = f"encoder.blocks.{i}"
            wqkv, bqkv = attn_fused(p + ".attn", d)
            L = enc_layers[i]
            L.attn_ln_g, L.attn_ln_b = P(self._dev(w[p + ".attn_ln.weight"], f32)), P(self._dev(w[p + ".attn_ln.bias"], f32))
            L.w_qkv, L.b_qkv = P(wqkv), P(bqkv)
            L.w_out, L.b_out = P(self._dev(w[p + ".attn.out.weight"], bf)), P(self._dev(w[p + ".attn.out.bias"], f32))
            L.mlp_ln_g, L.mlp_ln_b = P(self._dev(w[p + ".mlp_ln.weight"], f32)), P(self._dev(w[p + ".mlp_ln.bias"], f32))
            L.w_mlp1, L.b_mlp1 = P(self._dev(w[p + ".mlp1.weight"], bf)), P(self._dev(w[p + ".mlp1.bias"], f32))
            L.w_mlp2, L.b_mlp2 = P(self._dev(w[p + ".mlp2.weight"], bf)), P(self._dev(w[p + ".mlp2.bias"], f32))
        dt = dm.n_text_state
        dec_layers = (_lib.DecLayer * dm.n_text_layer)()
        for i in range(dm.n_text_layer):
            p = f"decoder.blocks.{i}"
            wqkv, bqkv = attn_fused(p + ".attn", dt)
            L = dec_layers[i]
            L.attn_ln_g, L.attn_ln_b = P(self._dev(w[p + ".attn_ln.weight"], f32)), P(self._dev(w[p + ".attn_ln.bias"], f32))
            L.w_qkv, L.b_qkv = P(wqkv), P(bqkv)
            L.w_out, L.b_out = P(self._dev(w[p + ".attn.out.weight"], bf)), P(self._dev(w[p + ".attn.out.bias"], f32))
            L.cross_ln_g = P(self._dev(w[p + ".cross_attn_ln.weight"], f32))
            L.cross_ln_b = P(self._dev(w[p + ".cross_attn_ln.bias"], f32))
            L.w_cq = P(self._dev(w[p + ".cross_attn.query.weight"], bf))
            L.b_cq = P(self._dev(w[p + ".cross_attn.query.bias"], f32))
            wk, wv = w[p + ".cross_attn.key.weight"], w[p + ".cross_attn.value.weight"]
            bv = w[p + ".cross_attn.value.bias"].float()
            L.w_ckv = P(self._dev(torch.cat([wk, wv], 0), bf))
            L.b_ckv = P(self._dev(torch.cat([torch.zeros_like(bv), bv], 0), f32))
            L.w_cout = P(self._dev(w[p + ".cross_attn.out.weight"], bf))
            L.b_cout = P(self._dev(w[p + ".cross_attn.out.bias"], f32))
            L.mlp_ln_g, L.mlp_ln_b = P(self._dev(w[p + ".mlp_ln.weight"], f32)), P(self._dev(w[p + ".mlp_ln.bias"], f32))
            L.w_mlp1, L.b_mlp1 = P(self._dev(w[p + ".mlp1.weight"], bf)), P(self._dev(w[p + ".mlp1.bias"], f32))
            L.w_mlp2, L.b_mlp2 = P(self._dev(w[p + ".mlp2.weight"], bf)), P(self._dev(w[p + ".mlp2.bias"], f32))

        W = _lib.Weights()
        W.dims = _lib.Dims(**asdict(dm))
        # MLX conv weight (out, k, in) flattens to the (out, 3*in) im2col operand as is
        W.conv1_w = P(self._dev(w["encoder.conv1.weight"].reshape(d, 3 * dm.n_mels), bf))
        W.conv1_b = P(self._dev(w["encoder.conv1.bias"], f32))
        W.conv2_w = P(self._dev(w["encoder.conv2.weight"].reshape(d, 3 * d), bf))
        W.conv2_b = P(self._dev(w["encoder.conv2.bias"], f32))
        W.enc_pos = P(self._dev(sinusoids(dm.n_audio_ctx, d), f32))
        W.ln_post_g, W.ln_post_b = P(self._dev(w["encoder.ln_post.weight"], f32)), P(self._dev(w["encoder.ln_post.bias"], f32))
        W.h_enc_layers = C.cast(enc_layers, C.POINTER(_lib.EncLayer))
        self.token_embedding = self._dev(w["decoder.token_embedding.weight"], bf)
        W.tok_emb = P(self.token_embedding)
        W.dec_pos = P(self._dev(w["decoder.positional_embedding"], bf))
        W.dec_ln_g, W.dec_ln_b = P(self._dev(w["decoder.ln.weight"], f32)), P(self._dev(w["decoder.ln.bias"], f32))
        W.h_dec_layers = C.cast(dec_layers, C.POINTER(_lib.DecLayer))
        handle = _lib.vp(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200w_model_create(C.byref(W), C.byref(handle)))
        self._handle = handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.b200w_model_destroy(self._handle)
                self._handle = _lib.vp(0)
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ------------------------------------------------------------------ reference-facing properties
    @property
    def is_multilingual(self) -> bool:
        return self.dims.n_vocab >= 51865

    @property
    def num_languages(self) -> int:
        return self.dims.n_vocab - 51765 - int(self.is_multilingual)

    @property
    def logits_ld(self) -> int:
        return _round_up(self.dims.n_vocab, 128)

    # ------------------------------------------------------------------ encoder side
    def mel_windows(self, mel: torch.Tensor, gmax: Optional[torch.Tensor], row0, size, gidx) -> torch.Tensor:
        """K1b + gather: (rows, n_mels) f32 log-mel -> (W, 3002, n_mels) bf16 slabs for the conv stem.

        `gmax` None: `mel` is already normalised.  row0/size/gidx: per-window first row, valid frames and
        index of the clamp maximum (python sequences).
        """
        n = len(row0)
        dev = self.device
        row0_t = torch.tensor(list(row0), dtype=torch.int64, device=dev)
        size_t = torch.tensor(list(size), dtype=torch.int32, device=dev)
        gidx_t = torch.tensor(list(gidx), dtype=torch.int32, device=dev)
        dst = torch.empty((n, N_FRAMES + 2, self.dims.n_mels), dtype=torch.bfloat16, device=dev)
        mel = mel.contiguous()
        with torch.cuda.device(dev):
            _lib.check(self._lib.b200w_mel_windows(_lib.ptr(mel), _lib.ptr(gmax), _lib.ptr(row0_t), _lib.ptr(size_t),
                                                   _lib.ptr(gidx_t), n, self.dims.n_mels, _lib.ptr(dst), _lib.stream()))
        return dst

    def _mel_to_slabs(self, mel: torch.Tensor) -> torch.Tensor:
        """(B, 3000, n_mels) / (3000, n_mels) normalised f32 (or 16-bit) mel -> bf16 padded slabs."""
        if mel.ndim == 2:
            mel = mel[None]
        assert mel.shape[1:] == (N_FRAMES, self.dims.n_mels), "incorrect audio shape"
        mel = mel.to(device=self.device, dtype=torch.float32).contiguous()
        B = mel.shape[0]
        return self.mel_windows(mel.view(B * N_FRAMES, -1), None, [b * N_FRAMES for b in range(B)], [N_FRAMES] * B,
                                [0] * B)

    def encode_slabs(self, slabs: torch.Tensor, want_f32: bool = False, stop_after_layers: int = -1):
        """Run the AudioEncoder on (W, 3002, n_mels) bf16 slabs -> (W, 1500, d) bf16 [, f32]."""
        dm = self.dims
        Wn = slabs.shape[0]
        ws_bytes = self._lib.b200w_encoder_workspace_bytes(self._handle, Wn)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        xa = torch.empty((Wn, dm.n_audio_ctx, dm.n_audio_state), dtype=torch.bfloat16, device=self.device)
        xa32 = torch.empty_like(xa, dtype=torch.float32) if (want_f32 or stop_after_layers >= 0) else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200w_encoder_forward(self._handle, _lib.ptr(slabs), Wn, _lib.ptr(ws), ws_bytes,
                                                       _lib.ptr(xa), _lib.ptr(xa32), stop_after_layers, _lib.stream()))
        return (xa, xa32) if (want_f32 or stop_after_layers >= 0) else xa

    def embed_audio(self, mel: torch.Tensor) -> torch.Tensor:
        """`model.encoder(mel)`: (B, 3000, n_mels) normalised log-mel -> (B, 1500, d) audio features (bf16)."""
        return self.encode_slabs(self._mel_to_slabs(mel))

    encoder = embed_audio

    def cross_kv(self, xa: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Cross-attention K/V of every decoder layer: (L, W, 1500, 2d) bf16, rows [K | V]."""
        dm = self.dims
        Wn = xa.shape[0]
        shape = (dm.n_text_layer, Wn, dm.n_audio_ctx, 2 * dm.n_text_state)
        if out is None:
            out = torch.empty(shape, dtype=torch.bfloat16, device=self.device)
        assert tuple(out.shape) == shape and out.is_contiguous()
        xa = xa.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.b200w_cross_kv(self._handle, _lib.ptr(xa), Wn, _lib.ptr(out), out.stride(0), 0,
                                                _lib.stream()))
        return out

    # ------------------------------------------------------------------ decoder side
    def side_streams(self, n: int):
        streams = self.__dict__.setdefault("_side_streams", [])
        while len(streams) < n:
            streams.append(torch.cuda.Stream(device=self.device))
        return streams[:n]

    def decode_session(self, n_audio: int, n_group: int, max_tokens: int, slot: int = 0):
        """A (cached) DecodeSession for batches of this shape; its buffers and CUDA graphs persist between calls.
        `slot` distinguishes sessions of equal shape that are alive at the same time (two-stream decoding)."""
        from .decoding import DecodeSession

        key = (n_audio, n_group, max_tokens, slot)
        cache = self.__dict__.setdefault("_sessions", {})
        if key not in cache:
            if len(cache) >= 6:  # bound the HBM held by idle sessions
                cache.pop(next(iter(cache)))
            cache[key] = DecodeSession(self, n_audio, n_group, max_tokens=max_tokens)
        return cache[key]

    def release_sessions(self) -> None:
        self.__dict__.pop("_sessions", None)

    def logits(self, tokens: torch.Tensor, audio_features: torch.Tensor) -> torch.Tensor:
        """Teacher-forced logits (B, n, V) f32 for `tokens` (B, n) given encoder states.

        The first two tokens go through one multi-token decoder pass, the rest through single-token
        steps over the paged KV cache, so both shapes of the decode path are exercised.
        """
        from .decoding import DecodeSession

        tokens = torch.as_tensor(tokens)
        if tokens.ndim == 1:
            tokens = tokens[None]
        B, n = tokens.shape
        sess = DecodeSession(self, audio_features, n_group=1, max_tokens=n + 1)
        sess.set_tokens(tokens.to(torch.int32))
        out = torch.empty((B, n, self.dims.n_vocab), dtype=torch.float32, device=self.device)
        i = 0
        first = min(2, n) if n > 1 else 1
        while i < n:
            nq = first if i == 0 else 1
            lg = sess.forward(nq)  # logits of the last of the nq tokens
            if nq > 1:
                # logits of the earlier prompt positions: re-run them one position at a time is not possible on
                # a shared cache, so evaluate position 0 via the aux (sot_index) path
                out[:, 0] = sess.aux_logits()[:, : self.dims.n_vocab]
            out[:, i + nq - 1] = lg[:, : self.dims.n_vocab]
            i += nq
        return out

    def decode(self, mel, options=None, **kwargs):
        from .decoding import decode as _decode, DecodingOptions

        return _decode(self, mel, options or DecodingOptions(), **kwargs)

    def detect_language(self, mel, tokenizer=None):
        from .decoding import detect_language as _detect

        return _detect(self, mel, tokenizer)
