"""Generate tests/golden/revd2_vectors.npz from the REAL reference (oracle/_ref/librl_ref.so): SYPS, SYRF and REVD2
(rl_syps.hh:21-143, rl_syrf.hh:21-118, rl_revd2.hh:75-246) on the inputs of test/drivers/test_revd2.cc (A = B^T B with B a polynomial-decay
mat_gen matrix; the Uplo test poisons the unused triangle with NaN).  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref  # noqa: E402
from _evdcases import CASES, evd_matrix  # noqa: E402

R = _ref.ref_lib()
assert R is not None
R.rlref_set_num_threads(1)
out = {}
for i, c in enumerate(CASES):
    A, Asym, st = evd_matrix(c, ref=R)
    rc, k, V, ev, st2 = _ref.ref_revd2(R, c["uplo"], A, c["k_start"], c["tol"], c["p"], c["q"], c["orth"], c["est_p"], st)
    out[f"ev{i}_rc"], out[f"ev{i}_k"] = np.array([rc], dtype=np.int64), np.array([k], dtype=np.int64)
    out[f"ev{i}_state_in"], out[f"ev{i}_state_out"] = np.array(st, dtype=np.uint32), np.array(st2, dtype=np.uint32)
    out[f"ev{i}_eig"] = ev
    out[f"ev{i}_Vhead"] = V[:24, :].copy()
    out[f"ev{i}_Achk"] = np.array([Asym.sum(), np.abs(Asym).sum(), Asym[0, 0], Asym[-1, -1]])
    recon = np.linalg.norm(Asym - (V * ev) @ V.T) / np.linalg.norm(Asym)
    out[f"ev{i}_recon"] = np.array([recon])
    # SYPS / SYRF on the same matrix and state
    ks = min(c["m"], 24)
    rcs, sk, sts = _ref.ref_syps(R, c["uplo"], A, ks, c["p"], c["q"], st)
    rcf, Q, stf = _ref.ref_syrf(R, c["uplo"], A, ks, c["p"], c["q"], c["orth"], st)
    out[f"ev{i}_syps_head"], out[f"ev{i}_syps_state"] = sk[:24, :].copy(), np.array(sts, dtype=np.uint32)
    out[f"ev{i}_syrf_head"], out[f"ev{i}_syrf_state"] = Q[:24, :].copy(), np.array(stf, dtype=np.uint32)
    print("revd2", i, c["m"], "k", c["k_start"], "->", k, "rc", rc, "recon", recon, "eig0", ev[0])
out["ev_count"] = np.array(len(CASES))
np.savez_compressed(os.path.join(HERE, "revd2_vectors.npz"), **out)
