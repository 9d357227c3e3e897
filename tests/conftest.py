import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """Build the checkers (and, if missing, the product library) once per session. Building is not using."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "librl_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/RandLAPACK") and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "librl_ref.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "randlapack_b200", "librlb200.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "randlapack_b200", "csrc"), "-j8"], stdout=subprocess.DEVNULL)


_ensure_built()


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import randlapack_b200 as rl
    c = rl.Context(0)
    yield c
    c.close()
