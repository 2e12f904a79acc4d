"""GPU parity tests proper: every case calls the CUDA path through the C ABI and checks it against the
CPU oracle (oracle/) or a torch fp32 reference of the same op.  Run with `pytest -m gpu` on a B200."""
import pytest

from tests.gpu_cases import CASES


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_case(name, built_lib):
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    metrics = CASES[name]()
    print(name, metrics)


@pytest.mark.gpu
def test_native_library_is_the_path(built_lib):
    """The product path is the CUDA library: kernels were launched, and nothing falls back to the CPU."""
    import numpy as np
    import torch
    from tools import synth
    from whisper_mlx_b200 import log_mel_spectrogram
    from whisper_mlx_b200.decoding import total_kernel_launches

    before = total_kernel_launches()
    out = log_mel_spectrogram(synth.white_noise(16000, 0), n_mels=80)
    assert out.is_cuda and out.shape == (100, 80)
    assert total_kernel_launches() > before
    with pytest.raises(RuntimeError):
        from whisper_mlx_b200.audio import log_mel_unclamped

        log_mel_unclamped(torch.zeros(16000), 80)  # CPU tensors are rejected, never silently computed
