"""Weight loading: host mirror of `mlx_whisper/load_models.py` (UPSTREAM; reached from
/root/reference/run:4 `--model mlx-community/whisper-large-v3-mlx`; restated in SURVEY.md A.3).

Reads the on-disk format the reference consumes -- `config.json` (the ten ModelDimensions fields,
optionally `model_type` / `quantization`) and `weights.safetensors` or `weights.npz` with MLX parameter
names and layouts -- into bf16 / f32 torch CUDA tensors.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict

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


def load_model(path_or_hf_repo: str, dtype: torch.dtype = torch.bfloat16, device=None) -> Whisper:
    """Load a Whisper model from a local directory or a Hugging Face repo id (MLX-format weights)."""
    model_path = Path(path_or_hf_repo)
    if not model_path.exists():
        try:
            from huggingface_hub import snapshot_download

            model_path = Path(snapshot_download(repo_id=path_or_hf_repo))
        except Exception as e:  # noqa: BLE001 - offline / unknown repo
            raise FileNotFoundError(
                f"{path_or_hf_repo!r} is neither a local directory nor a downloadable Hugging Face repo ({e})") from e

    with open(str(model_path / "config.json"), "r") as f:
        config = json.loads(f.read())
        config.pop("model_type", None)
        quantization = config.pop("quantization", None)
    if quantization is not None:
        raise NotImplementedError("MLX-quantised checkpoints are not supported yet; use the 16-bit weights")
    model_args = ModelDimensions(**config)
    weights = _read_weights(model_path)
    weights.pop("alignment_heads", None)
    weights = {k: v for k, v in weights.items() if not k.endswith("encoder.positional_embedding")}
    model = Whisper(model_args, weights, device=device, dtype=dtype)
    model.model_path = str(model_path)
    return model
