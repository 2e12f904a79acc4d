"""B200-native drop-in for the Whisper transcription path of geosurge-ai/whisper-mlx.

`./run input output` in the reference calls the third-party `mlx_whisper` console script
(/root/reference/run:3-6).  This package keeps that Python surface -- `transcribe`, `load_models`,
`audio`, `decoding`, `tokenizer`, `writers`, `cli` with the reference's names and arguments -- and runs
every stage on hand-written sm_100a kernels behind a C ABI (include/b200_whisper.h).

It is importable as `whisper_mlx_b200` (the directory name `whisper-mlx_b200` is not a Python
identifier; see whisper_mlx_b200/__init__.py at the repository root).
"""
from . import audio, decoding, load_models, tokenizer, writers  # noqa: F401
from ._version import __version__  # noqa: F401
from .transcribe import transcribe, transcribe_many  # noqa: F401
from .load_models import load_model  # noqa: F401
from .audio import log_mel_spectrogram, load_audio, pad_or_trim  # noqa: F401
from .decoding import DecodingOptions, DecodingResult, decode, detect_language  # noqa: F401
