"""Seeded synthetic inputs shared by tests, bench.py and the golden-vector generator.

There is no network: no real checkpoints, no datasets.  This module writes random-init Whisper
weights in the on-disk format the reference path consumes (`config.json` + `weights.safetensors`
with MLX parameter names and layouts, SURVEY.md Appendix B.3 -- the format of
`mlx-community/whisper-large-v3-mlx`, /root/reference/run:4) and makes seeded 16 kHz audio.
It is neither the oracle nor the product; both read the files it writes.
"""
from __future__ import annotations

import json
import os
import zlib

import numpy as np

# SURVEY.md Appendix B.1
DIMS = {
    "tiny": dict(n_mels=80, n_audio_ctx=1500, n_audio_state=384, n_audio_head=6, n_audio_layer=4,
                 n_vocab=51865, n_text_ctx=448, n_text_state=384, n_text_head=6, n_text_layer=4),
    "small": dict(n_mels=80, n_audio_ctx=1500, n_audio_state=768, n_audio_head=12, n_audio_layer=12,
                  n_vocab=51865, n_text_ctx=448, n_text_state=768, n_text_head=12, n_text_layer=12),
    "large-v3": dict(n_mels=128, n_audio_ctx=1500, n_audio_state=1280, n_audio_head=20, n_audio_layer=32,
                     n_vocab=51866, n_text_ctx=448, n_text_state=1280, n_text_head=20, n_text_layer=32),
    "large-v3-turbo": dict(n_mels=128, n_audio_ctx=1500, n_audio_state=1280, n_audio_head=20,
                           n_audio_layer=32, n_vocab=51866, n_text_ctx=448, n_text_state=1280,
                           n_text_head=20, n_text_layer=4),
    # a 2+2-layer toy with the tiny widths: fast enough for pure-CPU unit tests
    "micro": dict(n_mels=80, n_audio_ctx=1500, n_audio_state=128, n_audio_head=2, n_audio_layer=2,
                  n_vocab=51865, n_text_ctx=448, n_text_state=128, n_text_head=2, n_text_layer=2),
}


def weight_shapes(dims: dict) -> dict:
    """MLX parameter names -> shapes."""
    d, dt = dims["n_audio_state"], dims["n_text_state"]
    s = {
        "encoder.conv1.weight": (d, 3, dims["n_mels"]),
        "encoder.conv1.bias": (d,),
        "encoder.conv2.weight": (d, 3, d),
        "encoder.conv2.bias": (d,),
        "encoder.ln_post.weight": (d,),
        "encoder.ln_post.bias": (d,),
        "decoder.token_embedding.weight": (dims["n_vocab"], dt),
        "decoder.positional_embedding": (dims["n_text_ctx"], dt),
        "decoder.ln.weight": (dt,),
        "decoder.ln.bias": (dt,),
    }

    def attn(p, n):
        s[p + ".query.weight"] = (n, n)
        s[p + ".query.bias"] = (n,)
        s[p + ".key.weight"] = (n, n)
        s[p + ".value.weight"] = (n, n)
        s[p + ".value.bias"] = (n,)
        s[p + ".out.weight"] = (n, n)
        s[p + ".out.bias"] = (n,)

    def block(p, n, cross):
        attn(p + ".attn", n)
        s[p + ".attn_ln.weight"] = (n,)
        s[p + ".attn_ln.bias"] = (n,)
        if cross:
            attn(p + ".cross_attn", n)
            s[p + ".cross_attn_ln.weight"] = (n,)
            s[p + ".cross_attn_ln.bias"] = (n,)
        s[p + ".mlp1.weight"] = (4 * n, n)
        s[p + ".mlp1.bias"] = (4 * n,)
        s[p + ".mlp2.weight"] = (n, 4 * n)
        s[p + ".mlp2.bias"] = (n,)
        s[p + ".mlp_ln.weight"] = (n,)
        s[p + ".mlp_ln.bias"] = (n,)

    for i in range(dims["n_audio_layer"]):
        block(f"encoder.blocks.{i}", d, False)
    for i in range(dims["n_text_layer"]):
        block(f"decoder.blocks.{i}", dt, True)
    return s


def _init_one(name: str, shape, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    x = rng.standard_normal(shape, dtype=np.float32)
    if name.endswith("_ln.weight") or name.endswith("ln_post.weight") or name.endswith("decoder.ln.weight"):
        return 1.0 + 0.05 * x
    if name.endswith(".bias"):
        return 0.02 * x
    if "embedding" in name:
        return 0.05 * x
    fan_in = int(np.prod(shape[1:]))
    return x / np.float32(np.sqrt(fan_in))


def random_weights(dims: dict, seed: int = 0, device: str = "cpu"):
    """Yield (name, bf16 tensor).

    device="cpu": NumPy generator, identical on every machine (tests, golden vectors).
    device="cuda": torch CUDA generator with the same per-tensor scaling -- seconds instead of half a
    minute for large-v3; used by bench.py, which hands the very same tensors to the CPU baseline.
    """
    import torch

    if device == "cpu":
        for name, shape in weight_shapes(dims).items():
            yield name, torch.from_numpy(_init_one(name, shape, seed)).to(torch.bfloat16)
        return
    gen = torch.Generator(device=device)
    for name, shape in weight_shapes(dims).items():
        gen.manual_seed((seed << 32) ^ zlib.crc32(name.encode()))
        x = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        if name.endswith("_ln.weight") or name.endswith("ln_post.weight") or name.endswith("decoder.ln.weight"):
            x = 1.0 + 0.05 * x
        elif name.endswith(".bias"):
            x = 0.02 * x
        elif "embedding" in name:
            x = 0.05 * x
        else:
            x = x / float(np.sqrt(np.prod(shape[1:])))
        yield name, x.to(torch.bfloat16)


def write_model(path: str, model: str = "tiny", seed: int = 0, dtype: str = "bfloat16") -> str:
    """Write `config.json` + `weights.safetensors` (MLX names/layout) and return `path`."""
    import torch
    from safetensors.torch import save_file

    dims = DIMS[model] if isinstance(model, str) else dict(model)
    os.makedirs(path, exist_ok=True)
    tdt = getattr(torch, dtype)
    tensors = {k: v.to(tdt).contiguous() for k, v in random_weights(dims, seed)}
    save_file(tensors, os.path.join(path, "weights.safetensors"))
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump({**dims, "model_type": "whisper"}, f)
    return path


def load_weights_f32(path: str):
    """(dims dict, {name: f32 torch tensor}) for the oracle."""
    import torch
    from safetensors.torch import load_file

    with open(os.path.join(path, "config.json")) as f:
        cfg = json.load(f)
    cfg.pop("model_type", None)
    cfg.pop("quantization", None)
    w = {k: v.to(torch.float32) for k, v in load_file(os.path.join(path, "weights.safetensors")).items()}
    return cfg, w


# ----------------------------------------------------------------------------- audio

def white_noise(n_samples: int, seed: int = 0, sigma: float = 0.1) -> np.ndarray:
    return (np.random.default_rng(seed).standard_normal(n_samples) * sigma).astype(np.float32)


def tones(n_samples: int, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """Multi-tone + chirp mix with ~70 dB dynamic range between components."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sr
    x = np.zeros(n_samples)
    for k in range(6):
        f0 = rng.uniform(80, 7000)
        amp = 10.0 ** (-rng.uniform(0, 3.5))
        x += amp * np.sin(2 * np.pi * f0 * t + rng.uniform(0, 6.28))
    f_a, f_b = 200.0, 6000.0
    dur = n_samples / sr
    x += 0.3 * np.sin(2 * np.pi * (f_a * t + (f_b - f_a) * t * t / (2 * dur)))
    x += 1e-4 * rng.standard_normal(n_samples)
    return (0.5 * x / np.abs(x).max()).astype(np.float32)


def silence_click(n_samples: int, seed: int = 0) -> np.ndarray:
    x = np.zeros(n_samples, dtype=np.float32)
    rng = np.random.default_rng(seed)
    for p in rng.integers(0, n_samples, 5):
        x[p] = rng.choice([-0.9, 0.9])
    return x


def clipped(n_samples: int, seed: int = 0) -> np.ndarray:
    return np.clip(white_noise(n_samples, seed, 0.8), -1.0, 1.0)


def speechlike(n_samples: int, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """Amplitude-modulated harmonic bursts separated by near-silence (speech-shaped envelope)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n_samples, dtype=np.float64) / sr
    env = np.clip(np.sin(2 * np.pi * 0.7 * t + rng.uniform(0, 6.28)), 0, None) ** 2
    f0 = 110.0 + 30.0 * np.sin(2 * np.pi * 0.31 * t)
    ph = 2 * np.pi * np.cumsum(f0) / sr
    x = sum(np.sin((h + 1) * ph) / (h + 1) for h in range(12))
    x = env * x + 0.003 * rng.standard_normal(n_samples)
    return (0.4 * x / np.abs(x).max()).astype(np.float32)


AUDIO_KINDS = {"noise": white_noise, "tones": tones, "click": silence_click, "clip": clipped,
               "speech": speechlike}


def make_audio(kind: str, n_samples: int, seed: int = 0) -> np.ndarray:
    return AUDIO_KINDS[kind](n_samples, seed)


def long_audio(seconds: float, seed: int = 0, sr: int = 16000) -> np.ndarray:
    """Long-form mix: alternating 30 s blocks of noise / tones / speech-like (BASELINE config 4)."""
    n = int(round(seconds * sr))
    out = np.empty(n, dtype=np.float32)
    kinds = ["noise", "tones", "speech"]
    blk = 30 * sr
    for i, s in enumerate(range(0, n, blk)):
        m = min(blk, n - s)
        out[s : s + m] = make_audio(kinds[i % 3], m, seed * 1000 + i)
    return out
