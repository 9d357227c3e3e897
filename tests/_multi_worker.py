"""Worker for tests/test_gpu_multi.py (launched under torchrun, one rank per GPU): row-sharded RSVD over NCCL must
reproduce the single-GPU result on the same matrix and the same RNG state."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import randlapack_b200 as rl  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rl.Context(local)
    out = {}
    for (m, n, k, p, engine) in [(40000, 256, 32, 2, "dmma"), (70001, 128, 16, 3, "dmma"), (12800, 64, 64, 0, "dmma"),
                                 (40000, 256, 32, 2, "i8"), (70001, 128, 16, 3, "i8"),
                                 # the headline path: k >= 64 on the engine (fused Gram, R folded in RS, U = Y M on the engine)
                                 (65536, 512, 128, 2, "i8"), (65536, 512, 256, 0, "i8"), (40960, 256, 64, 3, "i8")]:
        ctx.set_fp64_engine(engine)      # "i8": the tall products over A on the tcgen05 int8 digit-slice engine
        # global matrix: planted decaying spectrum so that the factors are well defined; identical on every rank
        g = torch.Generator(device="cuda").manual_seed(1234)
        G1 = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
        sig = torch.logspace(0, -3, n, dtype=torch.float64, device="cuda")
        G2 = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g))[0]
        A_full = rl.to_f((torch.linalg.qr(G1)[0] * sig) @ G2.t())
        dist.broadcast(A_full.t(), src=0)
        r0, r1 = rl.shard_rows(m, world, rank)
        A_loc = rl.to_f(A_full[r0:r1].clone())
        stack = rl.RSVD(rl.QB(rl.RF(rl.RS(rl.CholQRQ(), p, 1), rl.CholQRQ()), rl.CholQRQ(), orth_check=True), k)
        ctx.set_shard(r0, m)
        st = rl.RNGState(3)
        rc, kk, U, S, V = stack.call(ctx, A_loc, k, 0.0, st)
        ctx.clear_shard()
        st1 = rl.RNGState(3)
        rc1, kk1, U1, S1, V1 = stack.call(ctx, A_full, k, 0.0, st1)
        res = {"rc": [rc, rc1], "k": [kk, kk1], "state_equal": st == st1,
               "S_rel": ((S[:kk] - S1[:kk]).abs().max() / S1[0]).item() if kk else 0.0,
               "V_abs": (V[:, :kk].abs() - V1[:, :kk].abs()).abs().max().item() if kk else 0.0,
               "U_abs": (U[:, :kk].abs() - U1[r0:r1, :kk].abs()).abs().max().item() if kk else 0.0}
        out[f"{m}x{n}_k{k}_p{p}_{engine}"] = res
    ctx.set_fp64_engine("i8")
    # row-sharded CQRRPT (sketch + Gram allreduce, R / J / rank replicated) == single-GPU CQRRPT on the same matrix and state
    for (m, n, engine) in [(40000, 200, "dmma"), (65536, 384, "i8")]:
        ctx.set_fp64_engine(engine)
        g = torch.Generator(device="cuda").manual_seed(77)
        A_full = rl.to_f(torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g))
        A_full *= (1.0 + torch.arange(n, device="cuda", dtype=torch.float64))[None, :] ** -1.0
        dist.broadcast(A_full.t(), src=0)
        r0, r1 = rl.shard_rows(m, world, rank)
        A_loc = rl.to_f(A_full[r0:r1].clone())
        A_one = A_full.clone()
        alg = rl.CQRRPT(False, None)
        ctx.set_shard(r0, m)
        st = rl.RNGState(5)
        rc, R, J = alg.call(ctx, A_loc, 1.5, st)
        rk = alg.rank
        ctx.clear_shard()
        st1 = rl.RNGState(5)
        rc1, R1, J1 = alg.call(ctx, A_one, 1.5, st1)
        out[f"cqrrpt_{m}x{n}_{engine}"] = {
            "rc": [rc, rc1], "k": [rk, alg.rank], "state_equal": st == st1, "J_equal": bool(torch.equal(J, J1)),
            "R_rel": ((R - R1).abs().max() / R1.abs().max()).item(),
            "Q_abs": (A_loc[:, :rk] - A_one[r0:r1, :rk]).abs().max().item()}
    ctx.set_fp64_engine("i8")
    # TSQR orthogonaliser: HQRQ on the row-sharded iterate (local Householder QR, stacked k x k factors, redundant QR, local product) as the
    # stabiliser of RS, the orthogonaliser of RF and of QB - once through the context's own NCCL communicator (ncclAllReduce issued from
    # C++) and once through the torch.distributed callback; both against the single-GPU run (Q is unique up to column signs)
    for native in (True, False):
        m, n, k, p = 40960, 256, 64, 2
        g = torch.Generator(device="cuda").manual_seed(4321)
        G1 = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
        sig = torch.logspace(0, -6, n, dtype=torch.float64, device="cuda")
        G2 = torch.linalg.qr(torch.randn((n, n), dtype=torch.float64, device="cuda", generator=g))[0]
        A_full = rl.to_f((torch.linalg.qr(G1)[0] * sig) @ G2.t())
        dist.broadcast(A_full.t(), src=0)
        r0, r1 = rl.shard_rows(m, world, rank)
        A_loc = rl.to_f(A_full[r0:r1].clone())
        stack = rl.RSVD(rl.QB(rl.RF(rl.RS(rl.HQRQ(), p, 1), rl.HQRQ()), rl.HQRQ(), orth_check=True), k)
        ctx.set_shard(r0, m, native=native)
        st = rl.RNGState(9)
        rc, kk, U, S, V = stack.call(ctx, A_loc, k, 0.0, st)
        ctx.clear_shard()
        st1 = rl.RNGState(9)
        rc1, kk1, U1, S1, V1 = stack.call(ctx, A_full, k, 0.0, st1)
        # orthogonality of the sharded U: U^T U summed over the shards
        Gm = U[:, :kk].t() @ U[:, :kk]
        dist.all_reduce(Gm)
        out[f"tsqr_hqrq_native{int(native)}"] = {
            "rc": [rc, rc1], "k": [kk, kk1], "state_equal": st == st1,
            "S_rel": ((S[:kk] - S1[:kk]).abs().max() / S1[0]).item(),
            "V_abs": (V[:, :kk].abs() - V1[:, :kk].abs()).abs().max().item(),
            "U_abs": (U[:, :kk].abs() - U1[r0:r1, :kk].abs()).abs().max().item(),
            "orth": (Gm - torch.eye(kk, dtype=torch.float64, device="cuda")).norm().item()}
    # row-sharded CQRRT == single-GPU CQRRT
    for (m, n) in [(65536, 320)]:
        g = torch.Generator(device="cuda").manual_seed(78)
        A_full = rl.to_f(torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g))
        A_full *= (1.0 + torch.arange(n, device="cuda", dtype=torch.float64))[None, :] ** -1.0
        dist.broadcast(A_full.t(), src=0)
        r0, r1 = rl.shard_rows(m, world, rank)
        A_loc = rl.to_f(A_full[r0:r1].clone())
        A_one = A_full.clone()
        alg = rl.CQRRT(False, None)
        ctx.set_shard(r0, m)
        st = rl.RNGState(5)
        rc, R = alg.call(ctx, A_loc, 1.5, st)
        ctx.clear_shard()
        st1 = rl.RNGState(5)
        rc1, R1 = alg.call(ctx, A_one, 1.5, st1)
        out[f"cqrrt_{m}x{n}"] = {"rc": [rc, rc1], "k": [n, n], "state_equal": st == st1,
                                 "R_rel": ((R.triu() - R1.triu()).abs().max() / R1.abs().max()).item(),
                                 "R_abs_rel": ((R.triu().abs() - R1.triu().abs()).abs().max() / R1.abs().max()).item(),
                                 "Q_abs": (A_loc - A_one[r0:r1]).abs().max().item(),
                                 "Q_absabs": (A_loc.abs() - A_one[r0:r1].abs()).abs().max().item()}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print("MULTI_RESULT " + json.dumps(gathered))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
