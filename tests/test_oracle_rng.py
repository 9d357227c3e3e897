"""Pins for the oracle's random-number layer (CPU only):
- Philox4x32-10 known-answer vectors the reference ships (RandBLAS/test/basic_rng/r123_kat_vectors.txt:19-21)
- counter carry semantics (RandBLAS/test/basic_rng/test_r123.cc:735-796)
- operator entries / next-state: bit-exact against the compiled reference (oracle/_ref) when present, and
  against the committed golden fixtures (generated from it) always
- sub-matrix == slice of the full matrix (RandBLAS/test/datastructures/test_denseskop.cc:162-300)
"""
import ctypes
import os

import numpy as np
import pytest

import _ref
from oracle import rl_oracle as O

u32 = ctypes.c_uint32
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))

KAT = [  # r123_kat_vectors.txt:19-21 — philox4x32 10 rounds: ctr(4) key(2) -> expected(4)
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,exp", KAT)
def test_philox_kat(ctr, key, exp):
    assert list(O.philox4x32_10(ctr, key)) == exp
    R = _ref.ref_lib()
    if R is not None:
        out = (u32 * 4)()
        R.rlref_philox4x32_10((u32 * 4)(*ctr), (u32 * 2)(*key), out)
        assert list(out) == exp


def test_counter_carry():
    # test_r123.cc:735-796: increments propagate carries across all four 32-bit words and wrap at 2^128
    assert O.ctr_incr((0xFFFFFFFF, 0, 0, 0), 1) == (0, 1, 0, 0)
    assert O.ctr_incr((0xFFFFFFFF, 0xFFFFFFFF, 0, 0), 1) == (0, 0, 1, 0)
    assert O.ctr_incr((0xFFFFFFFF,) * 4, 1) == (0, 0, 0, 0)
    assert O.ctr_incr((5, 0, 0, 0), (1 << 64) - 1) == (4, 0, 1, 0)
    assert O.ctr_incr((0, 0xFFFFFFFF, 0xFFFFFFFF, 7), 1 << 32) == (0, 0, 0, 8)
    big = 0x1234567890ABCDEF
    c = (0xFFFFFFF0, 0xFFFFFFFF, 0x10, 0)
    val = (c[0] | c[1] << 32 | c[2] << 64 | c[3] << 96) + big
    assert O.ctr_incr(c, big) == tuple((val >> (32 * i)) & 0xFFFFFFFF for i in range(4))
    R = _ref.ref_lib()
    if R is not None:
        cc = (u32 * 4)(*c)
        R.rlref_ctr_incr(cc, ctypes.c_uint64(big))
        assert tuple(cc) == O.ctr_incr(c, big)


def test_fill_dense_golden():
    L = _ref.oracle_lib()
    for i in range(int(GOLD["fill_count"])):
        nr, nc, fam, ax, lay, sr, sc, ro, co = [int(x) for x in GOLD[f"fill{i}_args"]]
        buf = GOLD[f"fill{i}_buf"]
        rc, got, nxt = _ref.fill_dense(L, "rlo", nr, nc, [int(x) for x in GOLD[f"fill{i}_seed"]], buf.dtype.type, fam, ax, lay,
                                       (sr, sc, ro, co))
        assert rc == 0
        assert np.array_equal(got.view(np.uint8), buf.view(np.uint8)), f"case {i}: entries differ from the reference's"
        assert nxt == [int(x) for x in GOLD[f"fill{i}_next"]], f"case {i}: next state differs"


def test_fill_dense_vs_compiled_reference():
    R = _ref.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    L = _ref.oracle_lib()
    rng = np.random.default_rng(1)
    n_cases = 0
    for dt in (np.float64, np.float32):
        for (nr, nc) in [(7, 5), (5, 7), (1, 1), (13, 4), (4, 13), (100, 3), (3, 100), (9, 9), (1, 17), (17, 1)]:
            for fam in (0, 1):
                for axis in (0, 1):
                    for lay in (0, 1, 2):
                        sr, sc = int(rng.integers(1, nr + 1)), int(rng.integers(1, nc + 1))
                        ro, co = int(rng.integers(0, nr - sr + 1)), int(rng.integers(0, nc - sc + 1))
                        seed = [int(x) for x in rng.integers(0, 2 ** 32, 6)]
                        if lay == 1:
                            seed[0], seed[1] = 0xFFFFFFF0, 0xFFFFFFFF
                        a = _ref.fill_dense(R, "rlref", nr, nc, seed, dt, fam, axis, lay, (sr, sc, ro, co))
                        b = _ref.fill_dense(L, "rlo", nr, nc, seed, dt, fam, axis, lay, (sr, sc, ro, co))
                        assert a[0] == b[0] == 0
                        assert np.array_equal(a[1].view(np.uint8), b[1].view(np.uint8))
                        assert a[2] == b[2]
                        n_cases += 1
    assert n_cases == 240


def test_submatrix_is_slice_of_full():
    # test_denseskop.cc:162-300: any sub-block generated through the offset rule equals the slice of the full sample
    st = O.RNGState(3, (0xFFFFFFFE, 0, 0, 0))
    for (nr, nc) in [(11, 6), (6, 11)]:
        for axis in (O.AXIS_LONG, O.AXIS_SHORT):
            full, _ = O.fill_dense(nr, nc, st, np.float64, major_axis=axis, layout=O.LAYOUT_COLMAJOR)
            for sub in [(3, 2, 2, 1), (nr, 1, 0, nc - 1), (1, nc, nr - 1, 0), (5, 5, 1, 1)]:
                blk, _ = O.fill_dense(nr, nc, st, np.float64, major_axis=axis, layout=O.LAYOUT_COLMAJOR, sub=sub)
                sr, sc, ro, co = sub
                assert np.array_equal(blk, full[ro:ro + sr, co:co + sc])


def test_next_state_arithmetic():
    # test_denseskop.cc:408-490 / dense_skops.hh:169-182: full fill advances by ceil(dim_major/4) * dim_minor
    for (nr, nc) in [(10, 3), (3, 10), (8, 8), (1, 5)]:
        st = O.RNGState(0, (0xFFFFFFFF, 0, 0, 0))
        _, nxt = O.fill_dense(nr, nc, st)
        major, minor = max(nr, nc), min(nr, nc)
        assert nxt.counter == O.ctr_incr(st.counter, ((major + 3) // 4) * minor)
        assert nxt.key == st.key


def test_gaussian_moments():
    # test_denseskop.cc:97-160 style sanity: mean ~ 0, variance ~ 1
    g, _ = O.fill_dense(2000, 50, O.RNGState(7), np.float64)
    assert abs(g.mean()) < 0.02 and abs(g.var() - 1.0) < 0.02
