import os
import sys

import pytest

# random-init weights and no vocabulary file in this image: the tests compare token ids and render text with the
# surrogate vocabulary (the product refuses to do that unless told so; tests/test_tokenizer_assets.py covers the refusal)
os.environ.setdefault("B200W_ALLOW_SURROGATE", "1")

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built_lib():
    """In-tree build of the CUDA library + CPU harness (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge

    ge.build()
    from whisper_mlx_b200 import _lib

    return _lib.load()
