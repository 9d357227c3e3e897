"""The C++ algorithm objects of include/RandLAPACK_B200.hh on a real device:
- standalone build: the reference's constructor arguments / call signatures / malloc ownership, QB invariants
- with-reference build (oracle/_ref/dropin_with_ref, compiled against the unmodified reference headers where
  /root/reference exists): the reference's own RSVD driver running on top of rlb200::QB, and the reference's QB on top
  of rlb200::RF, against the all-reference CPU stack on the same input and RNG state."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(path):
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    return r


def test_standalone_objects():
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_standalone")
    deps = [os.path.join(ROOT, "tests", "cpp", "dropin_test.cc"), os.path.join(ROOT, "include", "RandLAPACK_B200.hh"), os.path.join(ROOT, "include", "rlb200.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), deps[0],
                               "-o", exe, "-L" + os.path.join(ROOT, "randlapack_b200"), "-lrlb200",
                               "-Wl,-rpath,$ORIGIN/../../randlapack_b200"])
    r = _run(exe)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout


def test_reference_drivers_on_top_of_b200_objects():
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_with_ref")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_with_ref not built (reference tree absent at build time)")
    r = _run(exe)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout
