import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def nnp():
    """The product library, initialised on cuda:0. Fails loudly when it cannot run."""
    import nnue_data_compress_b200 as pkg

    pkg.init(int(os.environ.get("LOCAL_RANK", "0")))
    pkg.use_torch_stream()  # tests hand torch device tensors to the *_dev entry points
    yield pkg
    pkg.shutdown()
