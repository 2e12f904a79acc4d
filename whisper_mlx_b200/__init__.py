"""Import shim: makes the package that lives in `whisper-mlx_b200/` importable as `whisper_mlx_b200`.

The product directory is named after the reference repository plus the target (`whisper-mlx_b200`),
which is not a valid Python identifier.  This one-file package extends its own search path to that
directory, so `import whisper_mlx_b200`, `python -m whisper_mlx_b200` and submodule imports
(`whisper_mlx_b200.audio`, ...) all resolve to the real sources; nothing is duplicated.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "whisper-mlx_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
