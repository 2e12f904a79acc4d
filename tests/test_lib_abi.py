"""The C-ABI library loads without a GPU and exports every symbol include/b200_whisper.h declares."""
import ctypes
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "b200_whisper.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200w_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    names = _declared_symbols()
    assert len(names) >= 20, names
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in the header but not exported"


def test_binding_covers_header(built_lib):
    from whisper_mlx_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_version_and_error_string(built_lib):
    assert built_lib.b200w_version().startswith(b"b200-whisper")
    assert isinstance(built_lib.b200w_last_error(), bytes)
    assert built_lib.b200w_launch_count() == 0 or built_lib.b200w_launch_count() > 0


def test_argument_validation_without_gpu(built_lib):
    """Entry points reject bad arguments before touching the device (no compute calls here)."""
    from whisper_mlx_b200 import _lib

    rc = built_lib.b200w_layernorm(None, None, None, 4, 384, None, None, None)
    assert rc == -1 and b"null" in built_lib.b200w_last_error()
    rc = built_lib.b200w_logmel_finalize(None, None, 0, 0, None)
    assert rc == -1
    assert built_lib.b200w_encoder_workspace_bytes(None, 4) == 0


def test_no_fallback_when_library_missing(monkeypatch, tmp_path):
    """A missing extension is a loud failure, not a CPU fallback."""
    from whisper_mlx_b200 import _lib
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_try_build", lambda: (_ for _ in ()).throw(RuntimeError("nvcc missing")))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
