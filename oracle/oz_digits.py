"""TEST INFRASTRUCTURE ONLY (never imported by the product): numpy restatement of the arithmetic of the int8 digit-slice engine
(randlapack_b200/csrc/ozaki.cu) — exponent rule, balanced base-256 digits, digit-pair GEMMs with exact int32 accumulation, the
int64 three-diagonal recombination and the power-of-two scaling of the epilogue.  It has no counterpart in the reference (the
reference calls blas::gemm, e.g. rl_rs.hh:153,165, rl_rf.hh:123, rl_qb.hh:218); its oracle is plain fp64 A @ B, compared
componentwise against the bound |A| |B| with the tolerance that S digits give."""
import numpy as np


def P_of(S):
    return 8 * S - 2


def exponents(absmax, S):
    """E = max(biased exponent - 1022, P - 1023): |x| < 2^E for the whole group and 2^(P-E) is a normal double (oz_exp_from_field)."""
    f = (np.asarray(absmax, dtype=np.float64).view(np.int64) >> 52) & 0x7FF
    return np.maximum(f - 1022, P_of(S) - 1023)


def digits(x, E, S):
    """Balanced base-256 digits of rn(x * 2^(P-E)) exactly as oz_fixed/oz_pack4 produce them: add 0x80 to each of the S-1 low bytes, xor
    it off again; byte b is digit S-1-b.  Returns S int64 arrays d_t in [-128, 127] with sum_t d_t 256^(S-1-t) = rn(x 2^(P-E))."""
    P = P_of(S)
    F = np.rint(np.asarray(x, dtype=np.float64) * np.exp2((P - E).astype(np.float64))).astype(np.int64)
    low = np.uint64(0x8080808080808080 >> (8 * (9 - S)))
    u = (F.astype(np.uint64) + low) ^ low
    d = [((u >> np.uint64(8 * (S - 1 - t))) & np.uint64(0xFF)).astype(np.uint8).view(np.int8).astype(np.int64) for t in range(S)]
    assert np.array_equal(sum(d[t] * 256 ** (S - 1 - t) for t in range(S)), F)
    return d


def gemm_nn(A, B, S):
    """C = A @ B through the digit pairs s + t <= S - 1 (row scales for A, column scales for B), recombined as the epilogue does."""
    P = P_of(S)
    Ea = exponents(np.abs(A).max(axis=1), S)
    Eb = exponents(np.abs(B).max(axis=0), S)
    da, db = digits(A, Ea[:, None], S), digits(B, Eb[None, :], S)
    acc = [sum(da[s] @ db[d - s] for s in range(d + 1)) for d in range(S)]
    assert max(int(np.abs(a).max()) for a in acc) < 2 ** 31, "int32 accumulators would overflow"
    v = np.zeros(acc[0].shape)
    for g in range((S - 1) // 3, -1, -1):                      # three diagonals at a time, exactly, in int64
        t = sum(acc[3 * g + q] * 256 ** (2 - q) for q in range(3) if 3 * g + q < S)
        assert int(np.abs(t).max()) < 2 ** 51                  # the 1.5 * 2^52 conversion trick is exact
        v = v * 2.0 ** -24 + t.astype(np.float64)
    return v * np.exp2((Ea[:, None] + Eb[None, :] - (2 * P - 16 * (S - 1)) - 16).astype(np.float64))
