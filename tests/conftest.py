import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def sodium():
    """libsodium's ristretto255 API (bundled with pyzmq) -- the independent encoding-level oracle."""
    import ctypes
    import glob
    import zmq
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(zmq.__file__)), "pyzmq.libs", "libsodium*.so*"))
    if not libs:
        pytest.skip("libsodium not bundled in this image")
    lib = ctypes.CDLL(libs[0])
    lib.sodium_init()
    return lib
