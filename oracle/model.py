"""Oracle: Whisper AudioEncoder / TextDecoder (test infrastructure only; see oracle/__init__.py).

Restates `mlx_whisper/whisper.py` (UPSTREAM, not under /root/reference; call site
/root/reference/run:3-6) per SURVEY.md Appendix A.2, on PyTorch-CPU tensors, with the MLX
weight names and layouts of SURVEY.md Appendix B.3.  Cross-pinned against
transformers/models/whisper/modeling_whisper.py in tests/test_oracle_vs_hf.py.

`policy="fp32"` is the mathematical oracle.  `policy="bf16"` rounds activations to bfloat16
at the points where the GPU pipeline stores bf16 tensors (GEMM operands, attention
operands), keeping LayerNorm, softmax, the residual stream and every accumulation in fp32;
it is the checker for bit-level-ish agreement of the 16-bit pipeline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict

import torch
import torch.nn.functional as F


@dataclass(frozen=True)
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

    def to_dict(self):
        return asdict(self)


# SURVEY.md Appendix B.1 (public Whisper architecture)
DIMS = {
    "tiny": ModelDimensions(80, 1500, 384, 6, 4, 51865, 448, 384, 6, 4),
    "small": ModelDimensions(80, 1500, 768, 12, 12, 51865, 448, 768, 12, 12),
    "large-v3": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 32),
    "large-v3-turbo": ModelDimensions(128, 1500, 1280, 20, 32, 51866, 448, 1280, 20, 4),
}


def _round(x: torch.Tensor, policy: str) -> torch.Tensor:
    if policy == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    return x


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> torch.Tensor:
    """[sin | cos] halves with increment ln(10000)/(channels/2 - 1)."""
    assert channels % 2 == 0
    inc = math.log(max_timescale) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2, dtype=torch.float64))
    t = torch.arange(length, dtype=torch.float64)[:, None] * inv[None, :]
    return torch.cat([torch.sin(t), torch.cos(t)], dim=1).to(torch.float32)


def _linear(x, w, b=None):
    return F.linear(x, w, b)


def _layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps=1e-5)


def _gelu(x):
    return F.gelu(x)  # exact erf form


def qkv_attention(q, k, v, n_head: int, mask=None, policy="fp32"):
    """softmax((q*s)(k*s)^T) v with s = hd^-0.25, softmax in fp32."""
    B, n, d = q.shape
    hd = d // n_head
    scale = hd ** -0.25
    q = q.view(B, n, n_head, hd).permute(0, 2, 1, 3) * scale
    k = k.view(B, -1, n_head, hd).permute(0, 2, 3, 1) * scale
    v = v.view(B, -1, n_head, hd).permute(0, 2, 1, 3)
    qk = q @ k
    if mask is not None:
        qk = qk + mask  # (n, n_kv) additive causal mask, -inf above the diagonal
    w = torch.softmax(qk.float(), dim=-1)
    out = (_round(w, policy) @ v).permute(0, 2, 1, 3).reshape(B, n, d)
    return out, qk


def _mha(x, w, prefix, n_head, xa=None, mask=None, kv_cache=None, policy="fp32"):
    """MultiHeadAttention: query/value/out with bias, key without."""
    q = _linear(x, w[prefix + ".query.weight"], w[prefix + ".query.bias"])
    if xa is None:
        k = _linear(x, w[prefix + ".key.weight"])
        v = _linear(x, w[prefix + ".value.weight"], w[prefix + ".value.bias"])
        k, v = _round(k, policy), _round(v, policy)
        if kv_cache is not None:
            k = torch.cat([kv_cache[0], k], dim=1)
            v = torch.cat([kv_cache[1], v], dim=1)
    elif kv_cache is None:
        k = _linear(xa, w[prefix + ".key.weight"])
        v = _linear(xa, w[prefix + ".value.weight"], w[prefix + ".value.bias"])
        k, v = _round(k, policy), _round(v, policy)
    else:
        k, v = kv_cache
    q = _round(q, policy)
    out, qk = qkv_attention(q, k, v, n_head, mask, policy)
    out = _round(out, policy)
    return _linear(out, w[prefix + ".out.weight"], w[prefix + ".out.bias"]), (k, v), qk


def _block(x, w, prefix, n_head, xa=None, mask=None, kv_cache=None, cross=False, policy="fp32"):
    kv, cross_kv = kv_cache if kv_cache else (None, None)
    y, kv, _ = _mha(_round(_layer_norm(x, w[prefix + ".attn_ln.weight"], w[prefix + ".attn_ln.bias"]), policy),
                    w, prefix + ".attn", n_head, mask=mask, kv_cache=kv, policy=policy)
    x = x + y
    cross_qk = None
    if cross:
        y, cross_kv, cross_qk = _mha(
            _round(_layer_norm(x, w[prefix + ".cross_attn_ln.weight"], w[prefix + ".cross_attn_ln.bias"]), policy),
            w, prefix + ".cross_attn", n_head, xa=xa, kv_cache=cross_kv, policy=policy)
        x = x + y
    h = _round(_layer_norm(x, w[prefix + ".mlp_ln.weight"], w[prefix + ".mlp_ln.bias"]), policy)
    h = _round(_gelu(_linear(h, w[prefix + ".mlp1.weight"], w[prefix + ".mlp1.bias"])), policy)
    x = x + _linear(h, w[prefix + ".mlp2.weight"], w[prefix + ".mlp2.bias"])
    return x, (kv, cross_kv), cross_qk


def _conv1d_nlc(x, weight, bias, stride):
    """MLX Conv1d: input (B, L, Cin), weight (Cout, k, Cin), padding 1."""
    y = F.conv1d(x.transpose(1, 2), weight.permute(0, 2, 1), bias, stride=stride, padding=1)
    return y.transpose(1, 2)


@torch.no_grad()
def encoder_forward(w: dict, dims: ModelDimensions, mel: torch.Tensor, policy="fp32", return_stem=False):
    """AudioEncoder. mel: (B, 3000, n_mels) f32 -> (B, 1500, d) f32."""
    x = _round(mel.float(), policy)
    x = _round(_gelu(_conv1d_nlc(x, w["encoder.conv1.weight"], w["encoder.conv1.bias"], 1)), policy)
    x = _gelu(_conv1d_nlc(x, w["encoder.conv2.weight"], w["encoder.conv2.bias"], 2))
    assert x.shape[1:] == (dims.n_audio_ctx, dims.n_audio_state), "incorrect audio shape"
    x = x + sinusoids(dims.n_audio_ctx, dims.n_audio_state)
    if return_stem:
        return x
    for i in range(dims.n_audio_layer):
        x, _, _ = _block(x, w, f"encoder.blocks.{i}", dims.n_audio_head, policy=policy)
    return _layer_norm(x, w["encoder.ln_post.weight"], w["encoder.ln_post.bias"])


@torch.no_grad()
def decoder_forward(w: dict, dims: ModelDimensions, tokens: torch.Tensor, xa: torch.Tensor,
                    kv_cache=None, policy="fp32", return_cross_qk=False):
    """TextDecoder. tokens (B, n) int64, xa (B, 1500, d) -> logits (B, n, V), cache.
    return_cross_qk: also the per-layer cross-attention scores (B, H, n, 1500) before the softmax, as
    Whisper.forward_with_cross_qk hands them to timing.py."""
    offset = kv_cache[0][0][0].shape[1] if kv_cache else 0
    n = tokens.shape[-1]
    x = w["decoder.token_embedding.weight"][tokens] + w["decoder.positional_embedding"][offset : offset + n]
    if kv_cache is None:
        kv_cache = [None] * dims.n_text_layer
    mask = torch.full((n, offset + n), float("-inf")).triu_(offset + 1) if n > 1 else None
    xa = _round(xa, policy)
    new_cache, cross_qk = [], []
    for i in range(dims.n_text_layer):
        x, c, qk = _block(x, w, f"decoder.blocks.{i}", dims.n_text_head, xa=xa, mask=mask,
                          kv_cache=kv_cache[i], cross=True, policy=policy)
        new_cache.append(c)
        cross_qk.append(qk)
    x = _round(_layer_norm(x, w["decoder.ln.weight"], w["decoder.ln.bias"]), policy)
    logits = x @ w["decoder.token_embedding.weight"].T
    if return_cross_qk:
        return logits, new_cache, cross_qk
    return logits, new_cache
