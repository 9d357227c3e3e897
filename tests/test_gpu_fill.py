"""GPU parity, RandBLAS operator layer: Philox integer stream bit-exact; DenseDist entries vs the oracle.

Tolerance (stated): the Philox words and the counter->entry layout are integer work and must be bit-exact
(checked through the Uniform family, whose uneg11 conversion is exact arithmetic, and through next-state
equality).  Gaussian entries go through sin/cos/log: the device evaluates them in fp64 and rounds once, the
reference's host path calls libm's float routines (glibc documents up to 0.56 / 0.82 ulp for sinf / logf), and an
entry is a product of two such values; we allow <= 4 float ulps per entry and require that at least 90% of the
entries are bit-identical."""
import numpy as np
import pytest
import torch

import _ref
import randlapack_b200 as rl
from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu


def test_philox_stream_bit_exact(ctx):
    L = _ref.oracle_lib()
    for seed in ([0, 0, 0, 0, 0, 0], [0xFFFFFFF0, 0xFFFFFFFF, 0xFFFFFFFF, 3, 0xa4093822, 0x299f31d0]):
        n = 5000
        st = rl.RNGState(key=(seed[4], seed[5]), counter=seed[:4])
        got = rl.philox_stream(ctx, st, n).cpu().numpy().view(np.uint32).reshape(-1)
        exp = np.zeros(4 * n, dtype=np.uint32)
        L.rlo_philox_stream((_ref.u32 * 4)(*seed[:4]), (_ref.u32 * 2)(*seed[4:]), _ref.i64(n), exp.ctypes.data_as(_ref.ctypes.c_void_p))
        assert np.array_equal(got, exp)


CASES = [(7, 5), (5, 7), (1, 1), (13, 4), (4, 13), (100, 3), (3, 100), (9, 9), (1000, 33), (33, 1000), (256, 32), (1024, 256)]


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("family", [rl.FAMILY_UNIFORM, rl.FAMILY_GAUSSIAN])
def test_fill_dense_vs_oracle(ctx, dtype, family):
    rng = np.random.default_rng(11)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    n_entries, n_diff, max_ulp = 0, 0, 0
    for (nr, nc) in CASES:
        for axis in (rl.AXIS_LONG, rl.AXIS_SHORT):
            for lay in (rl.LAYOUT_NATURAL, rl.LAYOUT_COLMAJOR, rl.LAYOUT_ROWMAJOR):
                for trial in range(2):
                    if trial == 0:
                        sub = (nr, nc, 0, 0)
                    else:
                        sr, sc = int(rng.integers(1, nr + 1)), int(rng.integers(1, nc + 1))
                        sub = (sr, sc, int(rng.integers(0, nr - sr + 1)), int(rng.integers(0, nc - sc + 1)))
                    seed = [int(x) for x in rng.integers(0, 2 ** 32, 6)]
                    if trial == 1:
                        seed[0], seed[1] = 0xFFFFFF00, 0xFFFFFFFF   # force carries across words
                    st = rl.RNGState(key=(seed[4], seed[5]), counter=seed[:4])
                    D = rl.DenseDist(nr, nc, family, axis)
                    buf, nxt = rl.fill_dense(ctx, D, st, dtype, lay, sub)
                    rc, exp, exp_next = _ref.fill_dense(_ref.oracle_lib(), "rlo", nr, nc, seed, npdt, family, axis, lay, sub)
                    assert rc == 0
                    got = buf.cpu().numpy()
                    assert list(nxt.counter) + list(nxt.key) == exp_next
                    if family == rl.FAMILY_UNIFORM:
                        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), (nr, nc, axis, lay, sub)
                    else:
                        d = _ref.ulp_diff_f32(got, exp)
                        n_entries += d.size
                        n_diff += int((d > 0).sum())
                        max_ulp = max(max_ulp, int(d.max()))
    if family == rl.FAMILY_GAUSSIAN:
        print(f"gaussian entries: {n_entries}, differing from the host libm path: {n_diff}, max ulp {max_ulp}")
        assert max_ulp <= 4, f"max float-ulp distance {max_ulp}"
        assert n_diff <= 0.10 * n_entries, f"{n_diff}/{n_entries} entries differ from the host libm path"


def test_fill_dense_golden(ctx):
    import os
    GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.npz"))
    for i in range(int(GOLD["fill_count"])):
        nr, nc, fam, ax, lay, sr, sc, ro, co = [int(x) for x in GOLD[f"fill{i}_args"]]
        exp = GOLD[f"fill{i}_buf"]
        seed = [int(x) for x in GOLD[f"fill{i}_seed"]]
        st = rl.RNGState(key=(seed[4], seed[5]), counter=seed[:4])
        buf, nxt = rl.fill_dense(ctx, rl.DenseDist(nr, nc, fam, ax), st, torch.float64 if exp.dtype == np.float64 else torch.float32,
                                 lay, (sr, sc, ro, co))
        assert list(nxt.counter) + list(nxt.key) == [int(x) for x in GOLD[f"fill{i}_next"]]
        assert _ref.ulp_diff_f32(buf.cpu().numpy(), exp).max() <= 4


def test_fill_dense_rejects_bad_args(ctx):
    # randblas_require(D.n_rows >= n_rows + ro) (dense_skops.hh:562-563) -> error, not UB
    with pytest.raises(rl.Error):
        rl.fill_dense(ctx, rl.DenseDist(5, 5), rl.RNGState(), sub=(4, 4, 2, 0))


def test_large_fill_statistics(ctx):
    # size-independent property at a large size: moments of 2^24 entries
    D = rl.DenseDist(1 << 20, 16)
    buf, _ = rl.fill_dense(ctx, D, rl.RNGState(5))
    assert abs(buf.mean().item()) < 2e-3 and abs(buf.var().item() - 1.0) < 2e-3
