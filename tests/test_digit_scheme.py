"""CPU: the arithmetic of the int8 digit-slice engine (oracle/oz_digits.py restates randlapack_b200/csrc/ozaki.cu) against fp64 A @ B.
Tolerances (stated): componentwise against |A| |B|: 2e-12 for S = 6 (46 bits), 1e-13 for S = 7 (54 bits), 2e-7 for S = 4 (30 bits,
the fp32-storage setting) — the same bounds the GPU tests use (tests/test_gpu_ozaki.py)."""
import numpy as np
import pytest

from oracle import oz_digits as Z

TOL = {4: 2e-7, 5: 1e-9, 6: 2e-12, 7: 1e-13}


@pytest.mark.parametrize("S", [4, 5, 6, 7])
@pytest.mark.parametrize("bad_scaling", [False, True])
def test_digit_gemm_matches_fp64(S, bad_scaling):
    rng = np.random.default_rng(S)
    m, K, N = 37, 300, 11
    A, B = rng.standard_normal((m, K)), rng.standard_normal((K, N))
    if bad_scaling:
        A *= 10.0 ** rng.integers(-40, 40, (m, 1))
        B *= 10.0 ** rng.integers(-30, 30, (1, N))
    C = Z.gemm_nn(A, B, S)
    err = (np.abs(C - A @ B) / (np.abs(A) @ np.abs(B))).max()
    assert err <= TOL[S], err


def test_digits_are_balanced_bytes_and_exact():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1000) * 10.0 ** rng.integers(-5, 5, 1000)
    for S in (3, 4, 6, 7):
        E = Z.exponents(np.abs(x).max(), S)
        d = Z.digits(x, np.full(x.shape, E), S)            # reconstruction is asserted inside
        assert all(t.min() >= -128 and t.max() <= 127 for t in d)
        assert np.abs(d[0]).max() <= 65                    # the top digit only carries 6 bits + a carry


def test_exponent_rule_edge_cases():
    # zero group -> clamped exponent, every value slices to zero digits; a denormal-range group keeps fewer digits but a consistent scale
    for S in (6, 7):
        E0 = Z.exponents(0.0, S)
        assert E0 == Z.P_of(S) - 1023
        assert all(int(t[0]) == 0 for t in Z.digits(np.array([0.0]), np.array([E0]), S))
    x = np.array([1.0, 0.75, -0.5])
    E = Z.exponents(1.0, 6)
    assert E == 1 and np.all(np.abs(x) < 2.0 ** E)          # |x| < 2^E strictly
    A = np.array([[1e-320, 2e-320]]); B = np.array([[1.0], [1.0]])
    C = Z.gemm_nn(A, B, 6)
    assert np.isfinite(C).all() and abs(C[0, 0] - 3e-320) <= 1e-322 + 3e-320   # denormals lose digits, never produce garbage


def test_int32_headroom_at_the_largest_group():
    # worst case digits (all +-128 except the top one) over the largest accumulation group the engine uses (K = 16384)
    S, K = 7, 16384
    worst = S * 128 * 128 * K
    assert worst < 2 ** 31
