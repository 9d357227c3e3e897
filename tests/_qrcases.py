"""Shared by the CPU and GPU QR tests: regenerate the golden cases' inputs with the oracle's restatement of the reference's
test-matrix generators (RandLAPACK/testing/rl_gen.hh) and check the stored digests of the reference's own matrices."""
import os

import numpy as np

from oracle import rl_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "qr_vectors.npz"))


def npdt(tag):
    return np.float64 if str(G[tag]) == "f64" else np.float32


def cq_input(i):
    m, n, k, mt, nnz = [int(x) for x in G[f"cq{i}_args"]]
    cond, expo, scal, df, eps = [float(x) for x in G[f"cq{i}_fargs"]]
    dt = npdt(f"cq{i}_dtype")
    if mt == 0:
        A, st = O.gen_poly_mat(m, n, k if k else min(m, n), cond, expo, O.RNGState(0), dtype=dt)
    else:
        A, st = O.gen_adversarial_mat(m, n, scal, O.RNGState(0), dtype=dt)
    chk = np.array([A.sum(dtype=np.float64), np.abs(A).sum(dtype=np.float64), A[0, 0], A[-1, -1]])
    tol = (1e-9 if dt == np.float64 else 1e-3) * max(1.0, abs(G[f"cq{i}_Achk"][1]))
    assert np.all(np.abs(chk - G[f"cq{i}_Achk"]) <= tol), "oracle-regenerated input differs from the reference's mat_gen"
    assert list(st.words()) == list(G[f"cq{i}_state_in"])
    return A, st, dict(m=m, n=n, k=k, nnz=nnz, d_factor=df, eps=eps, dtype=dt)


def bq_input(i):
    m, n, k, b, qw, qt = [int(x) for x in G[f"bq{i}_args"]]
    cond, expo, df = [float(x) for x in G[f"bq{i}_fargs"]]
    dt = npdt(f"bq{i}_dtype")
    A, st = O.gen_poly_mat(m, n, k, cond, expo, O.RNGState(0), dtype=dt)
    chk = np.array([A.sum(dtype=np.float64), np.abs(A).sum(dtype=np.float64), A[0, 0], A[-1, -1]])
    tol = (1e-9 if dt == np.float64 else 1e-3) * max(1.0, abs(G[f"bq{i}_Achk"][1]))
    assert np.all(np.abs(chk - G[f"bq{i}_Achk"]) <= tol)
    assert list(st.words()) == list(G[f"bq{i}_state_in"])
    return A, st, dict(m=m, n=n, k=k, b=b, qrcp_wide=qw, qr_tall=qt, d_factor=df, dtype=dt)


def qr_invariants(A, Q, R, J, rank):
    """The three quantities test/drivers/test_cqrrpt.cc:60-104 asserts (each must be <= eps^0.75):
    ||A[:,J] - QR||_F / ||A||_F, max column residual / that column's norm, ||Q'Q - I||_F / sqrt(n)."""
    A64, Q64, R64 = A.astype(np.float64), Q[:, :rank].astype(np.float64), R[:rank].astype(np.float64)
    AP = A64[:, np.asarray(J, dtype=np.int64) - 1]
    E = AP - Q64 @ R64
    cn = np.linalg.norm(E, axis=0)
    j = int(np.argmax(cn))
    return (np.linalg.norm(E) / np.linalg.norm(A64), cn[j] / max(np.linalg.norm(AP[:, j]), 1e-300),
            np.linalg.norm(Q64.T @ Q64 - np.eye(rank)) / np.sqrt(A.shape[1]))


def geqp3_format_invariants(A, F, tau, J, rank):
    """test/drivers/test_bqrrp.cc:62-107: Q = ungqr(F[:, :rank], tau), R = triu(F)[:rank]; the three error measures
    (each must be <= eps^0.75): ||A[:,J] - QR||_F/||A||_F, worst column residual, ||Q'Q - I||_F/sqrt(n)."""
    from scipy.linalg import lapack
    m, n = A.shape
    k = int(rank)
    F64 = np.asfortranarray(F.astype(np.float64))
    q, _, info = lapack.dorgqr(np.asfortranarray(F64[:, :k]), np.asarray(tau[:k], dtype=np.float64))
    R = np.triu(F64)[:k, :]
    return qr_invariants(A, q, R, J, k)


def numerical_rank(F, rel=1e-10):
    dg = np.abs(np.diag(F))
    return int(np.sum(dg > rel * dg[0]))
