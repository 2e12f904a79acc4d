"""Weight loading: host mirror of `mlx_whisper/load_models.py` (UPSTREAM; reached from
/root/reference/run:4 `--model mlx-community/whisper-large-v3-mlx`; restated in SURVEY.md A.3).

Reads the on-disk format the reference consumes -- `config.json` (the ten ModelDimensions fields,
optionally `model_type` / `quantization`) and `weights.safetensors` or `weights.npz` with MLX parameter
names and layouts -- into bf16 / f32 torch CUDA tensors.  MLX group-quantised
checkpoints (2/4/8 bit) are expanded to dense weights at load; the kernels compute in bf16.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .whisper import ModelDimensions, Whisper


def _read_weights(model_path: Path) -> Dict[str, torch.Tensor]:
    st = model_path / "weights.safetensors"
    if st.exists():
        from safetensors.torch import load_file

        return load_file(str(st))
    npz = model_path / "weights.npz"
    if npz.exists():
        with np.load(str(npz)) as z:
            return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    raise FileNotFoundError(f"no weights.safetensors or weights.npz under {model_path}")


def dequantize(wq: torch.Tensor, scales: torch.Tensor, biases: torch.Tensor, group_size: int, bits: int) -> torch.Tensor:
    """MLX affine group quantisation -> f32.  `wq` is (out, in * bits / 32) uint32 with element j of a word at bits
    [j * bits, (j + 1) * bits); every `group_size` consecutive input elements share `scales[o, g]` / `biases[o, g]`:
    w = scale * q + bias (mx.quantize / nn.QuantizedLinear, UPSTREAM)."""
    if bits not in (2, 4, 8):
        raise NotImplementedError(f"MLX quantisation with bits={bits} is not supported (2, 4 or 8)")
    words = wq.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(0, 32, bits, dtype=torch.int64)
    q = ((words[..., None] >> shifts) & ((1 << bits) - 1)).reshape(words.shape[0], -1).to(torch.float32)
    out_dim, in_dim = q.shape
    if in_dim != scales.shape[-1] * group_size or scales.shape != biases.shape:
        raise ValueError(f"quantised weight {tuple(wq.shape)} does not match scales {tuple(scales.shape)} at group_size={group_size}")
    q = q.view(out_dim, -1, group_size)
    return (q * scales.to(torch.float32)[..., None] + biases.to(torch.float32)[..., None]).view(out_dim, in_dim)


def dequantize_weights(weights: Dict[str, torch.Tensor], group_size: int = 64, bits: int = 4) -> Dict[str, torch.Tensor]:
    """Replace every (`X.weight` uint32, `X.scales`, `X.biases`) triple by a dense `X.weight`; the engine stores bf16."""
    out = dict(weights)
    for key in [k for k in weights if k.endswith(".scales")]:
        base = key[: -len(".scales")]
        out[base + ".weight"] = dequantize(weights[base + ".weight"], weights[key], weights[base + ".biases"], group_size, bits)
        del out[key], out[base + ".biases"]
    return out


def resolve_model_path(path_or_hf_repo: str) -> Path:
    """A local directory, or the snapshot of a Hugging Face repo id (downloaded or already cached)."""
    model_path = Path(path_or_hf_repo)
    if not model_path.exists():
        try:
            from huggingface_hub import snapshot_download

            model_path = Path(snapshot_download(repo_id=path_or_hf_repo))
        except Exception as e:  # noqa: BLE001 - offline / unknown repo
            raise FileNotFoundError(
                f"{path_or_hf_repo!r} is neither a local directory nor a downloadable Hugging Face repo ({e})") from e
    return model_path


def read_model_files(model_path) -> Tuple[ModelDimensions, Dict[str, torch.Tensor], Optional[np.ndarray]]:
    """The host half of `load_model`: `config.json` -> dims, `weights.safetensors` / `weights.npz` -> dense tensors under
    the MLX parameter names (quantised triples expanded), and the optional `alignment_heads` table."""
    model_path = Path(model_path)
    with open(str(model_path / "config.json"), "r") as f:
        config = json.loads(f.read())
        config.pop("model_type", None)
        quantization = config.pop("quantization", None)
    model_args = ModelDimensions(**config)
    weights = _read_weights(model_path)
    if quantization is not None:  # MLX-quantised checkpoint (e.g. the 4-bit mlx-community variants): expand at load
        weights = dequantize_weights(weights, int(quantization.get("group_size", 64)), int(quantization.get("bits", 4)))
    alignment_heads = weights.pop("alignment_heads", None)
    weights = {k: v for k, v in weights.items() if not k.endswith("encoder.positional_embedding")}
    if alignment_heads is not None:  # (n, 2) [layer, head] pairs used by word-level timestamps
        alignment_heads = np.asarray(alignment_heads.cpu().numpy(), dtype=np.int64).reshape(-1, 2)
    return model_args, weights, alignment_heads


def load_model(path_or_hf_repo: str, dtype: torch.dtype = torch.bfloat16, device=None) -> Whisper:
    """Load a Whisper model from a local directory or a Hugging Face repo id (MLX-format weights)."""
    model_path = resolve_model_path(path_or_hf_repo)
    model_args, weights, alignment_heads = read_model_files(model_path)
    model = Whisper(model_args, weights, device=device, dtype=dtype)
    if alignment_heads is not None:
        model.alignment_heads = alignment_heads
    model.model_path = str(model_path)
    return model
