"""Host-side logic of the row-sharded (multi-GPU) path, on CPU with the gloo backend (world_size 2):
- shard_rows covers [0, m) with aligned, contiguous blocks
- the rlb200_allreduce_fn hook built over torch.distributed sums a buffer handed to it as a raw C pointer
- the operator rule sharding relies on: a row block of the m x k Gaussian equals the slice of the full sample,
  and stacking per-shard Gram / A^T Y partials reproduces the unsharded CholQR + B^T (oracle level)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import randlapack_b200 as rl
from randlapack_b200 import _capi
from oracle import rl_oracle as O


def test_shard_rows_cover():
    import pytest
    for m in (1, 127, 128, 129, 1000, 1 << 20, (1 << 24) + 5):
        for w in (1, 2, 3, 4, 8):
            per = -(-(-(-m // w)) // 128) * 128
            if (w - 1) * per >= m:
                # a trailing rank would be empty: every rank must fail alike, before any collective
                with pytest.raises(ValueError):
                    rl.shard_rows(m, w, w - 1)
                with pytest.raises(ValueError):
                    rl.shard_rows(m, w, 0)
                continue
            blocks = [rl.shard_rows(m, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 < a1
            assert all(b0 % 128 == 0 for b0, b1 in blocks)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hook = _capi.ALLREDUCE_FN(rl.make_allreduce_hook(None))
        # call it exactly as the C side does: raw pointer, element count, element size, stream (NULL on CPU)
        for dt, es in ((np.float64, 8), (np.float32, 4)):
            buf = np.arange(10, dtype=dt) * (rank + 1)
            rc = hook(None, buf.ctypes.data, 10, es, None)
            assert rc == 0
            assert np.array_equal(buf, np.arange(10, dtype=dt) * sum(range(1, world + 1)))
        # sharded CholQR + B^T on the oracle: per-rank partials + allreduce == unsharded result
        m, n, k = 1024, 32, 8
        A, st = O.gen_poly_mat(m, n, n, 50.0, 2.0, O.RNGState(0))
        Om, _ = O.fill_dense(n, k, st)
        r0, r1 = rl.shard_rows(m, world, rank)
        Y_loc = A[r0:r1] @ Om
        G = np.ascontiguousarray(Y_loc.T @ Y_loc)
        assert hook(None, G.ctypes.data, G.size, 8, None) == 0
        R = np.linalg.cholesky(G).T
        Q_loc = Y_loc @ np.linalg.inv(R)
        BT = np.ascontiguousarray(A[r0:r1].T @ Q_loc)
        assert hook(None, BT.ctypes.data, BT.size, 8, None) == 0
        _, Q_full = O.CholQRQ().call(np.asfortranarray(A @ Om))
        assert np.abs(Q_loc - Q_full[r0:r1]).max() < 1e-11
        assert np.abs(BT - A.T @ Q_full).max() < 1e-11
        # odd-p operator: each shard generates only its rows of the m x k Gaussian (dense_skops.hh:109-123)
        full, nxt_full = O.fill_dense(m, k, st)
        blk, _ = O.fill_dense(m, k, st, sub=(r1 - r0, k, r0, 0))
        assert np.array_equal(blk, full[r0:r1])
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_allreduce_hook_and_sharded_math_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
