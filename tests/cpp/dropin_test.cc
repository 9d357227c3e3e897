// Drop-in test of include/RandLAPACK_B200.hh (needs a B200 at run time).
//   default build  : self-contained objects, checks the factorisation invariants the reference's tests assert
//                    (test/comps/test_qb.cc:162-174 exponents).
//   -DWITH_REF     : compiled against the UNMODIFIED reference headers; the reference's own RandLAPACK::RSVD driver runs on top
//                    of rlb200::QB (which derives from RandLAPACK::QBalg), and a reference QB on top of rlb200::RF, and both
//                    are compared with the all-reference CPU stack on the same input and RNG state.
#ifdef WITH_REF
#include <RandBLAS.hh>
#include "RandLAPACK/rl_blaspp.hh"
#include "RandLAPACK/rl_lapackpp.hh"
#include "RandLAPACK/rl_exceptions.hh"
#include "RandLAPACK/misc/rl_util.hh"
#include "RandLAPACK/comps/rl_orth.hh"
#include "RandLAPACK/comps/rl_rs.hh"
#include "RandLAPACK/comps/rl_rf.hh"
#include "RandLAPACK/comps/rl_qb.hh"
#include "RandLAPACK/drivers/rl_rsvd.hh"
#include "RandLAPACK/drivers/rl_cqrrpt.hh"
#include "RandLAPACK/drivers/rl_cqrrt.hh"
#include "RandLAPACK/drivers/rl_bqrrp.hh"
#include "RandLAPACK/drivers/rl_revd2.hh"
#include "RandLAPACK/drivers/rl_cholqr_linops.hh"
#include "RandLAPACK/drivers/rl_scholqr3_linops.hh"
#include "RandLAPACK/drivers/rl_cqrrt_linops.hh"
#include "RandLAPACK/testing/rl_gen.hh"
#define RLB200_WITH_RANDLAPACK
#endif
#include "RandLAPACK_B200.hh"

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

static double fro(const std::vector<double>& a) { double s = 0; for (double v : a) s += v * v; return std::sqrt(s); }

// ||Q^T Q - I||_F for an m x k column-major Q
static double orth_err(int64_t m, int64_t k, const double* Q) {
    double s = 0;
    for (int64_t i = 0; i < k; ++i)
        for (int64_t j = 0; j < k; ++j) {
            double d = 0;
            for (int64_t r = 0; r < m; ++r) d += Q[r + i * m] * Q[r + j * m];
            d -= (i == j);
            s += d * d;
        }
    return std::sqrt(s);
}
// ||A - U diag(S) V^T||_F
static double resid(int64_t m, int64_t n, int64_t k, const double* A, const double* U, const double* S, const double* V) {
    double s = 0;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < m; ++i) {
            double d = A[i + j * m];
            for (int64_t l = 0; l < k; ++l) d -= U[i + l * m] * S[l] * V[j + l * n];
            s += d * d;
        }
    return std::sqrt(s);
}

int main() {
    const int64_t m = 600, n = 80, k = 20, p = 2, q = 1;
    int fails = 0;
#ifndef WITH_REF
    // planted rank-k matrix + small noise
    std::mt19937_64 gen(7);
    std::normal_distribution<double> nd;
    std::vector<double> L(m * k), R(n * k), A(m * n, 0.0);
    for (auto& v : L) v = nd(gen);
    for (auto& v : R) v = nd(gen);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < m; ++i) {
            double s = 1e-9 * nd(gen);
            for (int64_t l = 0; l < k; ++l) s += L[i + l * m] * R[j + l * n] / (1.0 + l);
            A[i + j * m] = s;
        }
    rlb200::CholQRQ<double> stab(false, false), orth_rf(false, false), orth_qb(false, false);   // the reference's ctor arguments
    rlb200::RS<double> rs(stab, p, q, false, false);
    rlb200::RF<double> rf(rs, orth_rf, false, false);
    rlb200::QB<double> qb(rf, orth_qb, false, true);
    rlb200::RSVD<double> rsvd(qb, k);
    rlb200::RNGState state(0);
    double *U = nullptr, *S = nullptr, *V = nullptr;
    int64_t kk = k;
    int rc = rsvd.call(m, n, A.data(), kk, 0.0, U, S, V, state);
    const double eps625 = std::pow(std::numeric_limits<double>::epsilon(), 0.625);
    double eu = orth_err(m, kk, U), ev = orth_err(n, kk, V), r = resid(m, n, kk, A.data(), U, S, V) / fro(A);
    std::printf("standalone: rc=%d k=%lld qb_code=%d  ||U'U-I||=%.2e ||V'V-I||=%.2e  resid=%.2e  state.ctr0=%u\n", rc, (long long)kk,
                rsvd.qb_code, eu, ev, r, state.counter[0]);
    fails += !(rc == 0 && kk == k && eu <= eps625 && ev <= eps625 && r <= 1e-7 && state.counter[0] == (uint32_t)(k * ((n + 3) / 4)));
    free(U); free(S); free(V);
    // a CPU stabiliser cannot be mixed into a device stack: rejected, never a silent fallback
    struct HostStab : rlb200::Stabilization<double> { int call(int64_t, int64_t, double*) override { return 0; } } hs;
    rlb200::RS<double> bad(hs, 0, 1, false, false);
    try { rlb200_stack_opts o{}; bad.fill(o); fails += 1; } catch (const std::invalid_argument&) {}
    {
        // CQRRPT and BQRRP on the same planted matrix: the invariants test/drivers/test_cqrrpt.cc:98-104 asserts
        std::vector<double> Aq = A, Rq(n * n, 0.0), tauq(n, 0.0);
        std::vector<int64_t> Jq(n);
        rlb200::CQRRPT<double> cq(false, std::pow(std::numeric_limits<double>::epsilon(), 0.85));
        rlb200::RNGState st(0);
        int rcq = cq.call(m, n, Aq.data(), m, Rq.data(), n, Jq.data(), 2.0, st);
        double eq = orth_err(m, cq.rank, Aq.data());
        std::printf("standalone CQRRPT: rc=%d rank=%lld ||Q'Q-I||=%.2e\n", rcq, (long long)cq.rank, eq);
        fails += !(rcq == 0 && cq.rank >= k && cq.rank <= n && eq <= 1e-9);
        std::vector<double> Ab = A;
        rlb200::BQRRP<double> bq(false, 16);
        rlb200::RNGState st2(0);
        int rcb = bq.call(m, n, Ab.data(), m, 1.0, tauq.data(), Jq.data(), st2);
        std::printf("standalone BQRRP: rc=%d rank=%lld\n", rcb, (long long)bq.rank);
        fails += !(rcb == 0 && bq.rank > 0 && bq.rank <= n);
        // hqrrp with the reference's argument list and CQRRPT with qrcp = hqrrp (rl_cqrrpt.hh:230-231)
        std::vector<double> Ah = A;
        rlb200::RNGState st3(0);
        int64_t rch = rlb200::hqrrp<double>(m, n, Ah.data(), m, Jq.data(), tauq.data(), 16, 4, 1, 0, st3, (double**)nullptr);
        std::vector<int64_t> seen(n, 0);
        for (int64_t i = 0; i < n; ++i) if (Jq[i] >= 1 && Jq[i] <= n) seen[Jq[i] - 1] += 1;
        bool perm_ok = true, diag_ok = true;
        for (int64_t i = 0; i < n; ++i) perm_ok = perm_ok && seen[i] == 1;
        diag_ok = std::abs(Ah[k + k * m]) <= 1e-6 * std::abs(Ah[0]) && std::abs(Ah[(k - 1) + (k - 1) * m]) >= 1e-4 * std::abs(Ah[0]);
        std::printf("standalone hqrrp: rc=%lld permutation %d  rank %lld revealed on R's diagonal %d\n", (long long)rch, (int)perm_ok, (long long)k, (int)diag_ok);
        fails += !(rch == 0 && perm_ok && diag_ok);
        std::vector<double> Aq2 = A;
        std::fill(Rq.begin(), Rq.end(), 0.0);
        rlb200::CQRRPT<double> cqh(false, std::pow(std::numeric_limits<double>::epsilon(), 0.85));
        cqh.qrcp = rlb200::CQRRPTSubroutines::hqrrp;
        cqh.nb_alg = 16; cqh.oversampling = 4;
        rlb200::RNGState st4(0);
        int rcq2 = cqh.call(m, n, Aq2.data(), m, Rq.data(), n, Jq.data(), 2.0, st4);
        double eq2 = orth_err(m, cqh.rank, Aq2.data());
        std::printf("standalone CQRRPT(qrcp = hqrrp): rc=%d rank=%lld ||Q'Q-I||=%.2e\n", rcq2, (long long)cqh.rank, eq2);
        fails += !(rcq2 == 0 && cqh.rank >= k && cqh.rank <= n && eq2 <= 1e-9);
        // CQRRPT_GPU's interface (rl_cqrrpt_gpu.hh:43-149): (verb, time_subroutines, eps), no_hqrrp switches geqp3 / hqrrp; same results as CQRRPT
        std::vector<double> Aq3 = A, Rq3(n * n, 0.0);
        std::vector<int64_t> Jq3(n);
        rlb200::CQRRPT_GPU<double> cqg(false, true, std::pow(std::numeric_limits<double>::epsilon(), 0.85));
        cqg.no_hqrrp = 0; cqg.nb_alg = 16; cqg.oversampling = 4;
        rlb200::RNGState st5(0);
        int rcq3 = cqg.call(m, n, Aq3.data(), m, Rq3.data(), n, Jq3.data(), 2.0, st5);
        double dq = 0;
        for (int64_t i = 0; i < m * n; ++i) dq = std::max(dq, std::abs(Aq3[i] - Aq2[i]));
        std::printf("standalone CQRRPT_GPU(no_hqrrp = 0): rc=%d rank=%lld  max|Q - Q(CQRRPT, qrcp = hqrrp)| %.2e  J equal %d  times %zu\n", rcq3,
                    (long long)cqg.rank, dq, (int)(Jq3 == Jq), cqg.times.size());
        fails += !(rcq3 == 0 && cqg.rank == cqh.rank && dq <= 1e-12 && Jq3 == Jq && cqg.times.size() == 8 && st5.counter[0] == st4.counter[0]);
    }
    {
        // REVD2 on a planted PSD matrix of rank 12, only the lower triangle valid (test/drivers/test_revd2.cc: Uplo)
        const int64_t me = 200, re = 12;
        std::vector<double> G(me * re), Se(me * me, 0.0);
        for (auto& v : G) v = nd(gen);
        for (int64_t j = 0; j < me; ++j)
            for (int64_t i = 0; i < me; ++i) {
                double acc = 0;
                for (int64_t l = 0; l < re; ++l) acc += G[i + l * me] * G[j + l * me] / ((1.0 + l) * (1.0 + l));
                Se[i + j * me] = acc;
            }
        std::vector<double> Al = Se;
        for (int64_t j = 0; j < me; ++j)
            for (int64_t i = 0; i < j; ++i) Al[i + j * me] = std::nan("");
        rlb200::SYPS<double> syps(3, 1, false, false);
        rlb200::HQRQ<double> orth(false, false);
        rlb200::SYRF<double> syrf(syps, orth, false, false);
        rlb200::REVD2<double> revd2(syrf, 10, false);
        rlb200::RNGState st(0);
        std::vector<double> Ve, ee;
        int64_t ke = 2;
        int rce = revd2.call(rlb200::Uplo::Lower, me, Al.data(), ke, 1e-12, Ve, ee, st);
        double num = 0;
        for (int64_t j = 0; j < me; ++j)
            for (int64_t i = 0; i < me; ++i) {
                double d = Se[i + j * me];
                for (int64_t l = 0; l < ke; ++l) d -= Ve[i + l * me] * ee[l] * Ve[j + l * me];
                num += d * d;
            }
        std::printf("standalone REVD2: rc=%d k=%lld  ||A - V E V'||/||A|| = %.2e  err_est=%.2e\n", rce, (long long)ke, std::sqrt(num) / fro(Se),
                    revd2.last_error_estimate);
        fails += !(rce == 0 && ke == 16 && std::sqrt(num) / fro(Se) <= 1e-12 && (int64_t)Ve.size() == me * ke && (int64_t)ee.size() == ke);
        // device-resident linear operators: C = 2 A^T B - C on the planted m x n matrix; symmetric operator from the NaN-poisoned triangle
        rlb200::DenseLinOp<double> Aop(m, n, A.data(), m);
        const int64_t nb = 7;
        std::vector<double> Bm(m * nb), Cm(n * nb, 1.0), Cref(n * nb);
        for (auto& v : Bm) v = nd(gen);
        for (int64_t j = 0; j < nb; ++j)
            for (int64_t i = 0; i < n; ++i) {
                double acc = 0;
                for (int64_t l = 0; l < m; ++l) acc += A[l + i * m] * Bm[l + j * m];
                Cref[i + j * n] = 2.0 * acc - 1.0;
            }
        Aop(rlb200::Layout::ColMajor, rlb200::Op::Trans, rlb200::Op::NoTrans, n, nb, m, 2.0, Bm.data(), m, -1.0, Cm.data(), n);
        double dl = 0, nl = 0;
        for (int64_t i = 0; i < n * nb; ++i) { dl = std::max(dl, std::abs(Cm[i] - Cref[i])); nl = std::max(nl, std::abs(Cref[i])); }
        rlb200::ExplicitSymLinOp<double> Sop(me, rlb200::Uplo::Lower, Al.data(), me);
        std::vector<double> Xs(me * 3), Ys(me * 3, 0.0);
        for (auto& v : Xs) v = nd(gen);
        Sop(rlb200::Layout::ColMajor, 3, 1.0, Xs.data(), me, 0.0, Ys.data(), me);
        double ds = 0, ns = 0;
        for (int64_t j = 0; j < 3; ++j)
            for (int64_t i = 0; i < me; ++i) {
                double acc = 0;
                for (int64_t l = 0; l < me; ++l) acc += Se[i + l * me] * Xs[l + j * me];
                ds = std::max(ds, std::abs(acc - Ys[i + j * me])); ns = std::max(ns, std::abs(acc));
            }
        std::printf("standalone linops: DenseLinOp %.2e  ExplicitSymLinOp %.2e  fro %.6e / %.6e\n", dl / nl, ds / ns, (double)Aop.fro_nrm(), fro(A));
        fails += !(dl <= 1e-12 * nl && ds <= 1e-12 * ns && std::abs(Aop.fro_nrm() - fro(A)) <= 1e-12 * fro(A));
    }
#else
    using RNG = r123::Philox4x32;
    auto run = [&](int which, std::vector<double>& Sout, double& res, double& eu) {
        auto state = RandBLAS::RNGState<RNG>();
        std::vector<double> A(m * n);
        RandLAPACK::gen::mat_gen_info<double> info((int64_t&)m, (int64_t&)n, RandLAPACK::gen::polynomial);
        info.cond_num = 2025; info.rank = n; info.exponent = 2.0;
        RandLAPACK::gen::mat_gen(info, A.data(), state);
        std::vector<double> A0 = A;
        double *U = nullptr, *S = nullptr, *V = nullptr;
        int64_t kk = k;
        int rc = 0;
        if (which == 0) {          // all reference (CPU)
            RandLAPACK::CholQRQ<double> stab(false, false), o1(false, false), o2(false, false);
            RandLAPACK::RS<double, RNG> rs(stab, p, q, false, false);
            RandLAPACK::RF<double, RNG> rf(rs, o1, false, false);
            RandLAPACK::QB<double, RNG> qb(rf, o2, false, false);
            RandLAPACK::RSVD<double, RNG> rsvd(qb, k);
            rc = rsvd.call(m, n, A.data(), kk, 0.0, U, S, V, state);
        } else if (which == 1) {   // reference RSVD driver on top of the B200 QB (derives from RandLAPACK::QBalg)
            rlb200::CholQRQ<double> stab(false, false), o1(false, false), o2(false, false);
            rlb200::RS<double> rs(stab, p, q, false, false);
            rlb200::RF<double> rf(rs, o1, false, false);
            rlb200::QB<double> qb(rf, o2, false, false);
            RandLAPACK::RSVD<double, RNG> rsvd(qb, k);
            rc = rsvd.call(m, n, A.data(), kk, 0.0, U, S, V, state);
        } else {                   // reference QB + RSVD on top of the B200 RF (derives from RandLAPACK::RangeFinder)
            rlb200::CholQRQ<double> stab(false, false), o1(false, false);
            RandLAPACK::CholQRQ<double> o2(false, false);
            rlb200::RS<double> rs(stab, p, q, false, false);
            rlb200::RF<double> rf(rs, o1, false, false);
            RandLAPACK::QB<double, RNG> qb(rf, o2, false, false);
            RandLAPACK::RSVD<double, RNG> rsvd(qb, k);
            rc = rsvd.call(m, n, A.data(), kk, 0.0, U, S, V, state);
        }
        Sout.assign(S, S + kk);
        res = resid(m, n, kk, A0.data(), U, S, V) / fro(A0);
        eu = orth_err(m, kk, U);
        std::printf("with-ref[%d]: rc=%d k=%lld resid=%.12e ||U'U-I||=%.2e state.ctr0=%u\n", which, rc, (long long)kk, res, eu, state.counter.v[0]);
        free(U); free(S); free(V);
        return (int)state.counter.v[0];
    };
    std::vector<double> S0, S1, S2;
    double r0, r1, r2, e0, e1, e2;
    int c0 = run(0, S0, r0, e0), c1 = run(1, S1, r1, e1), c2 = run(2, S2, r2, e2);
    double d1 = 0, d2 = 0;
    for (int64_t i = 0; i < k; ++i) { d1 = std::max(d1, std::abs(S1[i] - S0[i]) / S0[0]); d2 = std::max(d2, std::abs(S2[i] - S0[i]) / S0[0]); }
    std::printf("with-ref: max rel sigma diff  B200-QB under ref RSVD: %.2e   B200-RF under ref QB: %.2e\n", d1, d2);
    // device Gaussian entries differ from the host libm path by a few float ulps => 2e-6 (see tests/test_gpu_fill.py)
    fails += !(c0 == c1 && c1 == c2 && d1 <= 2e-6 && d2 <= 2e-6 && std::abs(r1 - r0) <= 2e-6 && std::abs(r2 - r0) <= 2e-6 && e1 <= 1e-9 && e2 <= 1e-9);
    {
        // CQRRPT and BQRRP called through the reference's abstract bases (CQRRPTalg / BQRRPalg): same pivots, rank and RNG state
        const int64_t mq = 900, nq = 96;
        auto gen = [&](std::vector<double>& A, RandBLAS::RNGState<RNG>& st) {
            A.assign(mq * nq, 0.0);
            RandLAPACK::gen::mat_gen_info<double> info((int64_t&)mq, (int64_t&)nq, RandLAPACK::gen::polynomial);
            info.cond_num = 100; info.rank = nq; info.exponent = 2.0;
            RandLAPACK::gen::mat_gen(info, A.data(), st);
        };
        const double tol = std::pow(std::numeric_limits<double>::epsilon(), 0.85);
        RandLAPACK::CQRRPT<double, RNG> cq_ref(false, tol);
        rlb200::CQRRPT<double> cq_dev(false, tol);
        std::vector<int64_t> J0(nq), J1(nq);
        std::vector<double> R0(nq * nq, 0.0), R1(nq * nq, 0.0), A0, A1;
        auto call_cq = [&](RandLAPACK::CQRRPTalg<double, RNG>& alg, std::vector<double>& A, std::vector<double>& R, std::vector<int64_t>& J) {
            auto st = RandBLAS::RNGState<RNG>();
            gen(A, st);
            alg.call(mq, nq, A.data(), mq, R.data(), nq, J.data(), 2.0, st);
            return (int)st.counter.v[0];
        };
        int s0 = call_cq(cq_ref, A0, R0, J0), s1 = call_cq(cq_dev, A1, R1, J1);
        double dR = 0, nR = 0;
        for (int64_t i = 0; i < nq * nq; ++i) { dR = std::max(dR, std::abs(R0[i] - R1[i])); nR = std::max(nR, std::abs(R0[i])); }
        std::printf("with-ref CQRRPT: rank %lld / %lld  J equal %d  max|dR|/max|R| %.2e  state %d / %d\n", (long long)cq_ref.rank, (long long)cq_dev.rank,
                    (int)(J0 == J1), dR / nR, s0, s1);
        fails += !(cq_ref.rank == cq_dev.rank && J0 == J1 && dR <= 1e-9 * nR && s0 == s1);

        RandLAPACK::BQRRP<double, RNG> bq_ref(false, 32);
        rlb200::BQRRP<double> bq_dev(false, 32);
        std::vector<double> t0(nq, 0.0), t1(nq, 0.0);
        auto call_bq = [&](RandLAPACK::BQRRPalg<double, RNG>& alg, std::vector<double>& A, std::vector<double>& tau, std::vector<int64_t>& J) {
            auto st = RandBLAS::RNGState<RNG>();
            gen(A, st);
            alg.call(mq, nq, A.data(), mq, 1.0, tau.data(), J.data(), st);
            return (int)st.counter.v[0];
        };
        s0 = call_bq(bq_ref, A0, t0, J0); s1 = call_bq(bq_dev, A1, t1, J1);
        double dA = 0, dt = 0;
        for (int64_t i = 0; i < mq * nq; ++i) dA = std::max(dA, std::abs(A0[i] - A1[i]));
        for (int64_t i = 0; i < nq; ++i) dt = std::max(dt, std::abs(t0[i] - t1[i]));
        std::printf("with-ref BQRRP: rank %lld / %lld  J equal %d  max|dA| %.2e  max|dtau| %.2e  state %d / %d\n", (long long)bq_ref.rank,
                    (long long)bq_dev.rank, (int)(J0 == J1), dA, dt, s0, s1);
        fails += !(bq_ref.rank == bq_dev.rank && J0 == J1 && dA <= 1e-9 && dt <= 1e-9 && s0 == s1);

        // hqrrp, the reference's free function (rl_hqrrp.hh:811) and rlb200::hqrrp with the same argument list
        auto call_hq = [&](bool dev, std::vector<double>& A, std::vector<double>& tau, std::vector<int64_t>& J) {
            auto st = RandBLAS::RNGState<RNG>();
            gen(A, st);
            if (dev) rlb200::hqrrp<double>(mq, nq, A.data(), mq, J.data(), tau.data(), 32, 8, 1, 0, st, (double**)nullptr);
            else RandLAPACK::hqrrp(mq, nq, A.data(), mq, J.data(), tau.data(), (int64_t)32, (int64_t)8, (int64_t)1, (int64_t)0, st, (double**)nullptr);
            return (int)st.counter.v[0];
        };
        s0 = call_hq(false, A0, t0, J0); s1 = call_hq(true, A1, t1, J1);
        dA = 0; dt = 0;
        for (int64_t i = 0; i < mq * nq; ++i) dA = std::max(dA, std::abs(A0[i] - A1[i]));
        for (int64_t i = 0; i < nq; ++i) dt = std::max(dt, std::abs(t0[i] - t1[i]));
        std::printf("with-ref hqrrp: J equal %d  max|dA| %.2e  max|dtau| %.2e  state %d / %d\n", (int)(J0 == J1), dA, dt, s0, s1);
        fails += !(J0 == J1 && dA <= 1e-9 && dt <= 1e-9 && s0 == s1);
    }
    {
        // the reference's REVD2 on its own SYRF / SYPS / HQRQ objects vs rlb200::REVD2, same matrix (test_revd2.cc recipe), same state
        int64_t me = 300, re = 60;
        std::vector<double> B(me * me), Sm(me * me, 0.0);
        auto st0 = RandBLAS::RNGState<RNG>();
        RandLAPACK::gen::mat_gen_info<double> info(me, me, RandLAPACK::gen::polynomial);
        info.cond_num = 1e4; info.rank = re; info.exponent = 2.0;
        RandLAPACK::gen::mat_gen(info, B.data(), st0);
        blas::syrk(blas::Layout::ColMajor, blas::Uplo::Lower, blas::Op::Trans, me, me, 1.0, B.data(), me, 0.0, Sm.data(), me);
        for (int64_t j = 0; j < me; ++j)
            for (int64_t i = 0; i < j; ++i) Sm[i + j * me] = Sm[j + i * me];
        using SYPS_t = RandLAPACK::SYPS<double, RNG>;
        using SYRF_t = RandLAPACK::SYRF<SYPS_t, RandLAPACK::HQRQ<double>>;
        SYPS_t syps_r(3, 1, false, false);
        RandLAPACK::HQRQ<double> orth_r(false, false);
        SYRF_t syrf_r(syps_r, orth_r, false, false);
        RandLAPACK::REVD2<SYRF_t> revd2_r(syrf_r, 10, false);
        rlb200::SYPS<double> syps_d(3, 1, false, false);
        rlb200::HQRQ<double> orth_d(false, false);
        rlb200::SYRF<double> syrf_d(syps_d, orth_d, false, false);
        rlb200::REVD2<double> revd2_d(syrf_d, 10, false);
        std::vector<double> V0, e0, V1, e1;
        int64_t k0 = 4, k1 = 4;
        auto sa = RandBLAS::RNGState<RNG>(5), sb = RandBLAS::RNGState<RNG>(5);
        revd2_r.call(blas::Uplo::Upper, me, Sm.data(), k0, 1e-13, V0, e0, sa);
        revd2_d.call(blas::Uplo::Upper, me, Sm.data(), k1, 1e-13, V1, e1, sb);
        double de = 0;
        for (int64_t i = 0; i < std::min(k0, k1); ++i) de = std::max(de, std::abs(e0[i] - e1[i]));
        std::printf("with-ref REVD2: k %lld / %lld  max|d eig| %.2e  state %u / %u\n", (long long)k0, (long long)k1, de, sa.counter.v[0], sb.counter.v[0]);
        fails += !(k0 == k1 && de <= 1e-10 * e0[0] && sa.counter.v[0] == sb.counter.v[0]);
        // the REFERENCE's REVD2 (operator overload, rl_revd2.hh:142-150) on a device-resident symmetric operator
        static_assert(RandLAPACK::linops::SymmetricLinearOperator<rlb200::ExplicitSymLinOp<double>>);
        static_assert(RandLAPACK::linops::LinearOperator<rlb200::DenseLinOp<double>>);
        rlb200::ExplicitSymLinOp<double> Sop(me, blas::Uplo::Upper, Sm.data(), me);
        std::vector<double> V2, e2;
        int64_t k2 = 4;
        auto sc = RandBLAS::RNGState<RNG>(5);
        revd2_r.call(Sop, k2, 1e-13, V2, e2, sc);
        double de2 = 0;
        for (int64_t i = 0; i < std::min(k0, k2); ++i) de2 = std::max(de2, std::abs(e0[i] - e2[i]));
        std::printf("with-ref REVD2 on rlb200::ExplicitSymLinOp: k %lld / %lld  max|d eig| %.2e  device products %lld  e[3] %.17g / %.17g\n", (long long)k0,
                    (long long)k2, de2, (long long)Sop.n_products(), e0[3], e2[3]);
        fails += !(k0 == k2 && de2 <= 1e-10 * e0[0] && sa.counter.v[0] == sc.counter.v[0] && Sop.n_products() > 0);
    }
    {
        // The REFERENCE's operator-templated QR drivers - CholQR_linops (rl_cholqr_linops.hh:35), sCholQR3_linops (rl_scholqr3_linops.hh:50) and
        // CQRRT_linops (rl_cqrrt_linops.hh:33; SASO and dense sketch) - on a device-resident rlb200::DenseLinOp: every operator product, and the
        // sketch S A with S regenerated on the device from S.dist / S.seed_state, runs on the B200.  Same R as on the reference's own DenseLinOp.
        const int64_t ml = 3000, nl = 64;
        std::vector<double> Al(ml * nl, 0.0);
        auto stl = RandBLAS::RNGState<RNG>();
        RandLAPACK::gen::mat_gen_info<double> info((int64_t&)ml, (int64_t&)nl, RandLAPACK::gen::polynomial);
        info.cond_num = 100; info.rank = nl; info.exponent = 2.0;
        RandLAPACK::gen::mat_gen(info, Al.data(), stl);
        RandLAPACK::linops::DenseLinOp<double> Ah(ml, nl, Al.data(), ml, blas::Layout::ColMajor);
        rlb200::DenseLinOp<double> Ad(ml, nl, Al.data(), ml);
        const double tol = std::pow(std::numeric_limits<double>::epsilon(), 0.85);
        auto rdiff = [&](const std::vector<double>& R0, const std::vector<double>& R1) {
            double d = 0, s = 0;
            for (int64_t j = 0; j < nl; ++j)
                for (int64_t i = 0; i <= j; ++i) { d = std::max(d, std::abs(R0[i + j * nl] - R1[i + j * nl])); s = std::max(s, std::abs(R0[i + j * nl])); }
            return d / s;
        };
        std::vector<double> R0(nl * nl, 0.0), R1(nl * nl, 0.0);
        RandLAPACK::CholQR_linops<double> c0(false, tol), c1(false, tol);
        int rc0 = c0.call(Ah, R0.data(), nl), rc1 = c1.call(Ad, R1.data(), nl);
        const double d_chol = rdiff(R0, R1);
        std::fill(R0.begin(), R0.end(), 0.0); std::fill(R1.begin(), R1.end(), 0.0);
        RandLAPACK::sCholQR3_linops<double> s0(false, tol), s1(false, tol);
        int rs0 = s0.call(Ah, R0.data(), nl), rs1 = s1.call(Ad, R1.data(), nl);
        const double d_s3 = rdiff(R0, R1);
        double d_cq[2] = {0, 0};
        int rq[2][2] = {{0, 0}, {0, 0}};
        uint32_t sq[2][2] = {{0, 0}, {0, 0}};
        for (int dense = 0; dense < 2; ++dense) {
            std::fill(R0.begin(), R0.end(), 0.0); std::fill(R1.begin(), R1.end(), 0.0);
            RandLAPACK::CQRRT_linops<double, RNG> q0(false, tol), q1(false, tol);
            q0.use_dense_sketch = q1.use_dense_sketch = dense != 0;
            q0.nnz = q1.nnz = 4;
            auto sa = RandBLAS::RNGState<RNG>(11), sb = RandBLAS::RNGState<RNG>(11);
            rq[dense][0] = q0.call(Ah, R0.data(), nl, 2.0, sa); rq[dense][1] = q1.call(Ad, R1.data(), nl, 2.0, sb);
            sq[dense][0] = sa.counter.v[0]; sq[dense][1] = sb.counter.v[0];
            d_cq[dense] = rdiff(R0, R1);
        }
        std::printf("with-ref linop drivers on rlb200::DenseLinOp: max|dR|/max|R|  CholQR_linops %.2e  sCholQR3_linops %.2e  CQRRT_linops saso %.2e dense %.2e"
                    "  device products %lld\n", d_chol, d_s3, d_cq[0], d_cq[1], (long long)Ad.n_products);
        // the dense sketch's Gaussians differ from the host libm's by a few float ulps (tests/test_gpu_fill.py): R of the preconditioned matrix moves by ~1e-7 relative
        fails += !(rc0 == 0 && rc1 == 0 && rs0 == rs1 && rq[0][0] == rq[0][1] && rq[1][0] == rq[1][1] && sq[0][0] == sq[0][1] && sq[1][0] == sq[1][1] &&
                   d_chol <= 1e-11 && d_s3 <= 1e-11 && d_cq[0] <= 1e-11 && d_cq[1] <= 1e-9 && Ad.n_products >= 8);
    }
    {
        // RandBLAS::sketch_general and rlb200::sketch_general, same argument lists (host pointers): dense and sparse operators, left and right,
        // a RowMajor / transposed combination each, and a long-axis (LASO) operator
        const int64_t ms = 700, ns = 40, ds = 24;
        std::vector<double> As(ms * ns), B0, B1;
        std::mt19937_64 g2(11);
        std::normal_distribution<double> nd2;
        for (auto& v : As) v = nd2(g2);
        auto maxdiff = [&](const std::vector<double>& X, const std::vector<double>& Y) {
            double dmax = 0, smax = 0;
            for (size_t i = 0; i < X.size(); ++i) { dmax = std::max(dmax, std::abs(X[i] - Y[i])); smax = std::max(smax, std::abs(X[i])); }
            return dmax / smax;
        };
        double e[6];
        {   // dense, left, ColMajor, NoTrans / NoTrans with offsets
            RandBLAS::DenseDist D(ds + 2, ms + 3);
            RandBLAS::DenseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(3));
            B0.assign(ds * ns, 0.5); B1 = B0;
            RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.5, S, 1, 2, As.data(), ms, -0.5, B0.data(), ds);
            rlb200::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.5, S, 1, 2, As.data(), ms, -0.5, B1.data(), ds);
            e[0] = maxdiff(B0, B1);
        }
        {   // dense, left, RowMajor, opS = Trans (S is (m x d)), A given as the RowMajor m x n array = the same buffer read as ColMajor n x m
            RandBLAS::DenseDist D(ns + 1, ds + 1);      // here the data matrix is As read as RowMajor (ns x ms)... contraction over ns
            RandBLAS::DenseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(4));
            B0.assign(ds * ms, 0.25); B1 = B0;
            // B (ds x ms, RowMajor) = S[0:ns, 0:ds]^T (ds x ns) * A (ns x ms, RowMajor)
            RandBLAS::sketch_general(blas::Layout::RowMajor, blas::Op::Trans, blas::Op::NoTrans, ds, ms, ns, 1.0, S, 0, 0, As.data(), ms, 2.0, B0.data(), ms);
            rlb200::sketch_general(blas::Layout::RowMajor, blas::Op::Trans, blas::Op::NoTrans, ds, ms, ns, 1.0, S, 0, 0, As.data(), ms, 2.0, B1.data(), ms);
            e[1] = maxdiff(B0, B1);
        }
        {   // dense, right: B (ms x ds) = A (ms x ns) * S (ns x ds)
            RandBLAS::DenseDist D(ns, ds);
            RandBLAS::DenseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(5));
            B0.assign(ms * ds, 0.0); B1 = B0;
            RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ms, ds, ns, 1.0, As.data(), ms, S, 0, 0, 0.0, B0.data(), ms);
            rlb200::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ms, ds, ns, 1.0, As.data(), ms, S, 0, 0, 0.0, B1.data(), ms);
            e[2] = maxdiff(B0, B1);
        }
        {   // sparse (SASO), left, ColMajor
            RandBLAS::SparseDist D(ds, ms, 3);
            RandBLAS::SparseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(6));
            B0.assign(ds * ns, 1.0); B1 = B0;
            RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.0, S, 0, 0, As.data(), ms, 1.0, B0.data(), ds);
            rlb200::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.0, S, 0, 0, As.data(), ms, 1.0, B1.data(), ds);
            e[3] = maxdiff(B0, B1);
        }
        {   // sparse (SASO), right with a tall operator, RowMajor: B (ms x ds) = A (ms x ns) S (ns x ds), everything RowMajor
            RandBLAS::SparseDist D(ns, ds, 2);
            RandBLAS::SparseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(7));
            B0.assign(ms * ds, 0.0); B1 = B0;
            RandBLAS::sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, ms, ds, ns, 1.0, As.data(), ns, S, 0, 0, 0.0, B0.data(), ds);
            rlb200::sketch_general(blas::Layout::RowMajor, blas::Op::NoTrans, blas::Op::NoTrans, ms, ds, ns, 1.0, As.data(), ns, S, 0, 0, 0.0, B1.data(), ds);
            e[4] = maxdiff(B0, B1);
        }
        {   // sparse, long-axis (LASO), left
            RandBLAS::SparseDist D(ds, ms, 8, RandBLAS::Axis::Long);
            RandBLAS::SparseSkOp<double, RNG> S(D, RandBLAS::RNGState<RNG>(8));
            B0.assign(ds * ns, 0.0); B1 = B0;
            RandBLAS::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.0, S, 0, 0, As.data(), ms, 0.0, B0.data(), ds);
            rlb200::sketch_general(blas::Layout::ColMajor, blas::Op::NoTrans, blas::Op::NoTrans, ds, ns, ms, 1.0, S, 0, 0, As.data(), ms, 0.0, B1.data(), ds);
            e[5] = maxdiff(B0, B1);
        }
        std::printf("with-ref sketch_general (RandBLAS vs rlb200, same arguments): dense left %.2e  dense RowMajor/opS=T %.2e  dense right %.2e  saso left %.2e"
                    "  saso right RowMajor %.2e  laso left %.2e\n", e[0], e[1], e[2], e[3], e[4], e[5]);
        // dense: the device's Gaussians are within a few float ulps of the host libm's (2e-6); sparse: +-1 / sqrt(count) entries, exact up to summation order
        fails += !(e[0] <= 2e-6 && e[1] <= 2e-6 && e[2] <= 2e-6 && e[3] <= 1e-13 && e[4] <= 1e-13 && e[5] <= 1e-13);
    }
#endif
    std::printf(fails ? "DROPIN_FAIL\n" : "DROPIN_OK\n");
    return fails;
}
