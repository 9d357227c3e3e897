// Host-side control flow: Stabilization / RS / RF / QB / RSVD on device-resident, optionally row-sharded data.
// Reference `call`s reproduced (paths relative to the reference root):
//   CholQRQ::call  RandLAPACK/comps/rl_orth.hh:68-98      RS::call   RandLAPACK/comps/rl_rs.hh:116-178
//   RF::call       RandLAPACK/comps/rl_rf.hh:106-137      QB::call   RandLAPACK/comps/rl_qb.hh:133-268
//   RSVD::call     RandLAPACK/drivers/rl_rsvd.hh:113-154
#include "drivers.cuh"
#include "philox.cuh"
#include <algorithm>
#include <cmath>
#include <limits>

namespace rlb {

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
int ws_reserve(Ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return 0;
    if (ctx->ws) {
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        RLB_CUDA_OK(ctx, cudaFree(ctx->ws));
        ctx->ws = nullptr; ctx->ws_bytes = 0;
    }
    size_t want = std::max(bytes, (size_t)64 << 20);
    if (cudaMalloc(&ctx->ws, want) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "workspace allocation of " + std::to_string(want) + " bytes failed";
        return RLB200_ERR_ALLOC;
    }
    ctx->ws_bytes = want;
    return 0;
}

static std::vector<ArenaChunk>& chunks_of(Ctx* ctx) { return ctx->arena; }
size_t arena_mark(Ctx* ctx) {
    size_t t = 0;
    for (auto& c : chunks_of(ctx)) t += c.used;
    return t;
}
void* arena_push(Ctx* ctx, size_t bytes) {
    bytes = ws_round(std::max<size_t>(bytes, 1));
    auto& ch = chunks_of(ctx);
    // first chunk (in order) with room whose successors are all empty keeps stack discipline
    for (size_t i = 0; i < ch.size(); ++i) {
        bool later_empty = true;
        for (size_t j = i + 1; j < ch.size(); ++j) later_empty &= (ch[j].used == 0);
        if (later_empty && ch[i].cap - ch[i].used >= bytes) {
            void* p = ch[i].p + ch[i].used;
            ch[i].used += bytes;
            return p;
        }
    }
    size_t cap = std::max(bytes, (size_t)32 << 20);
    void* p = nullptr;
    if (cudaMalloc(&p, cap) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "device allocation of " + std::to_string(cap) + " bytes failed";
        return nullptr;
    }
    ch.push_back({static_cast<char*>(p), cap, bytes});
    return p;
}
void arena_release(Ctx* ctx, size_t mark_total) {
    auto& ch = chunks_of(ctx);
    size_t t = 0;
    for (auto& c : ch) {
        if (t >= mark_total) c.used = 0;
        else if (t + c.used > mark_total) c.used = mark_total - t;
        t += c.used;
    }
}
void arena_destroy(Ctx* ctx) {
    for (auto& c : ctx->arena) cudaFree(c.p);
    ctx->arena.clear();
}

#define RLB_ALLOC(ctx, ptr)                                   \
    do { if (!(ptr)) return RLB200_ERR_ALLOC; } while (0)

template <typename T>
static int allreduce_sum(Ctx* ctx, T* buf, int64_t count) {
    if (!ctx->allreduce) return 0;
    int rc = ctx->allreduce(ctx->allreduce_user, buf, count, (int32_t)sizeof(T), ctx->stream);
    if (rc != 0) { ctx->err = "allreduce hook failed with code " + std::to_string(rc); return RLB200_ERR_COLLECTIVE; }
    return 0;
}

template <typename T>
static int read_scalar(Ctx* ctx, const T* dev, T* host) {
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(ctx->hbox, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(host, ctx->hbox, sizeof(T));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Stabilization
// ------------------------------------------------------------------------------------------------
// cond(R) via the singular values of the k x k factor (util::cond_num_check, rl_util.hh:402-424)
template <typename T>
static int cond_of_square(Ctx* ctx, int64_t k, const T* R, double* cond) {
    ArenaScope as(ctx);
    T* Rc = as.take<T>(k * k); RLB_ALLOC(ctx, Rc);
    T* S = as.take<T>(k); RLB_ALLOC(ctx, S);
    T* W = as.take<T>(k * k); RLB_ALLOC(ctx, W);
    void* ws = arena_push(ctx, svd_ws_bytes(k, k, sizeof(T))); RLB_ALLOC(ctx, ws);
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(Rc, R, sizeof(T) * k * k, cudaMemcpyDeviceToDevice, ctx->stream));
    RLB_CHECK(svd_tall<T>(ctx, k, k, Rc, k, S, W, ws, nullptr));
    std::vector<T> s(k);
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(s.data(), S, sizeof(T) * k, cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *cond = (s[k - 1] == 0) ? std::numeric_limits<double>::infinity() : (double)s[0] / (double)s[k - 1];
    return 0;
}

template <typename T>
static int cholqrq(Ctx* ctx, int64_t m, int64_t k, T* A, bool cond_check, bool rows_sharded, int* chol_fail) {
    RLB_REQUIRE(ctx, k <= 16384);
    if (k == 0) return 0;
    ArenaScope as(ctx);
    T* G = as.take<T>(k * k); RLB_ALLOC(ctx, G);
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * k * k, ctx->stream));
    // G = A^T A, upper triangle (rl_orth.hh:78)
    RLB_CHECK(gemm_tn<T>(ctx, m, k, k, 1.0, A, m, A, m, 0.0, G, k, /*upper_only=*/1));
    if (rows_sharded) RLB_CHECK(allreduce_sum<T>(ctx, G, k * k));
    // R = chol(G) (rl_orth.hh:81); failure => chol_fail, return 1
    int info = 0;
    RLB_CHECK(potrf_blocked<T>(ctx, k, G, k, &info));
    if (info != 0) {
        if (chol_fail) *chol_fail = 1;
        return 1;
    }
    if (cond_check) {   // rl_orth.hh:88-93 — the whole k x k buffer (upper R, zero strictly-lower part) is examined
        if (k > 256) RLB_CHECK(tri_op<T>(ctx, 1, k, k, G, k, G, k));   // the blocked factorisation scribbles below the diagonal
        double cond = 0;
        RLB_CHECK(cond_of_square<T>(ctx, k, G, &cond));
        if (cond > 1.0 / std::sqrt((double)std::numeric_limits<T>::epsilon())) return 1;
    }
    // A <- A R^{-1} (rl_orth.hh:95): blocked right-solve whose blocks are tall tensor-pipe GEMMs
    RLB_CHECK(trsm_right_upper<T>(ctx, m, k, G, k, A, m));
    return 0;
}

// HQRQ::call (rl_orth.hh:144-164): geqrf + ungqr.  Householder QR of the tall panel, then Q = (I - V T V^T) [I; 0] = E - V (T V1^T)
// formed in place by one tall tensor-pipe GEMM.
template <typename T>
static int hqrq(Ctx* ctx, int64_t m, int64_t k, T* A, T* R_out = nullptr) {
    if (m == 0 || k == 0) return 0;
    const int64_t kk = std::min(m, k);
    RLB_REQUIRE(ctx, k <= 256 && m >= k);   // (the in-place product needs one CTA tile across all k columns)
    ArenaScope as(ctx);
    T* tau = as.take<T>(kk); RLB_ALLOC(ctx, tau);
    T* G = as.take<T>(kk * kk); RLB_ALLOC(ctx, G);
    T* Tm = as.take<T>(kk * kk); RLB_ALLOC(ctx, Tm);
    T* V1 = as.take<T>(kk * kk); RLB_ALLOC(ctx, V1);
    T* M = as.take<T>(kk * kk); RLB_ALLOC(ctx, M);
    void* ws = arena_push(ctx, qrcp_ws_bytes(k)); RLB_ALLOC(ctx, ws);
    RLB_CHECK(qr_small<T>(ctx, false, m, k, A, m, nullptr, tau, ws));
    if (R_out) RLB_CHECK(tri_op<T>(ctx, 2, kk, kk, A, m, R_out, kk));               // the triangular factor (clean upper copy, ld kk)
    RLB_CHECK(set_upper_diag<T>(ctx, kk, A, m, (T)1, false));                       // A <- V (unit lower trapezoidal, clean)
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * kk * kk, ctx->stream));
    RLB_CHECK(gemm_tn<T>(ctx, m, kk, kk, 1.0, A, m, A, m, 0.0, G, kk, 0));          // G = V^T V
    RLB_CUDA_OK(ctx, cudaMemsetAsync(Tm, 0, sizeof(T) * kk * kk, ctx->stream));
    RLB_CHECK(larft_from_gram<T>(ctx, kk, G, kk, tau, Tm, kk));
    RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(V1, kk * sizeof(T), A, m * sizeof(T), kk * sizeof(T), kk, cudaMemcpyDeviceToDevice, ctx->stream));
    RLB_CHECK(gemm_nt<T>(ctx, kk, kk, kk, 1.0, Tm, kk, V1, kk, 0.0, M, kk));        // M = T V1^T
    RLB_CHECK(gemm_nn_inplace<T>(ctx, m, kk, kk, -1.0, A, m, M, kk));               // A <- -V M
    RLB_CHECK(set_upper_diag<T>(ctx, kk, A, m, (T)1, true));                        // + [I; 0]
    return 0;
}

// HQRQ on a row-sharded iterate: TSQR (net-new; the reference has no distributed layer).  One level, as wide as the shard count:
//   local Householder QR  A_g = Q_g R_g   (rl_orth.hh:144-164 semantics on the shard's rows),
//   the k x k factors stacked in rank order [R_0; ...; R_{W-1}] on every rank (a sum-allreduce of a zero-padded buffer: adding zeros is
//   exact, so all ranks hold bit-identical stacks), its Householder QR  = Q2 R  computed redundantly, and  Q_g <- Q_g Q2[g k : (g+1) k, :].
// Unlike CholQR this does not square the condition number of the iterate.
template <typename T>
static int hqrq_tsqr(Ctx* ctx, int64_t m, int64_t k, T* A) {
    const int W = ctx->shard_world, g = ctx->shard_rank;
    if (W <= 0 || g < 0) { ctx->err = "HQRQ on a row-sharded iterate (TSQR) needs the shard's rank: rlb200_set_shard_rank / rlb200_comm_init"; return RLB200_ERR_ARG; }
    RLB_REQUIRE(ctx, k <= 256 && m >= k);      // every shard holds at least k rows
    ArenaScope as(ctx);
    T* Rl = as.take<T>(k * k); RLB_ALLOC(ctx, Rl);
    T* St = as.take<T>((int64_t)W * k * k); RLB_ALLOC(ctx, St);
    T* M = as.take<T>(k * k); RLB_ALLOC(ctx, M);
    RLB_CHECK(hqrq<T>(ctx, m, k, A, Rl));
    RLB_CUDA_OK(ctx, cudaMemsetAsync(St, 0, sizeof(T) * (size_t)W * k * k, ctx->stream));
    RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(St + (int64_t)g * k, (size_t)W * k * sizeof(T), Rl, k * sizeof(T), k * sizeof(T), k, cudaMemcpyDeviceToDevice, ctx->stream));
    RLB_CHECK(allreduce_sum<T>(ctx, St, (int64_t)W * k * k));
    RLB_CHECK(hqrq<T>(ctx, (int64_t)W * k, k, St));                                 // St <- Q2 (W k x k)
    RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(M, k * sizeof(T), St + (int64_t)g * k, (size_t)W * k * sizeof(T), k * sizeof(T), k, cudaMemcpyDeviceToDevice, ctx->stream));
    RLB_CHECK(gemm_nn_inplace<T>(ctx, m, k, k, 1.0, A, m, M, k));
    return 0;
}

// R = chol(A^T A) (upper, k x k in G, ld k) without applying it: the first half of CholQRQ::call (rl_orth.hh:78-93).
// Returns 0, 1 (potrf failure / cond check) like the stabiliser.  Used where the triangular solve can be folded into the NEXT
// product algebraically:  A^T (Y R^-1) = (A^T Y) R^-1  and  (Y R^-1) W = Y (R^-1 W)  — the m x k solve (m k^2 flops, a full
// read+write of the tall iterate) becomes an n x k or k x k one.
template <typename T>
static int cholqr_finish(Ctx* ctx, int64_t k, bool cond_check, bool rows_sharded, T* G, int* chol_fail);
template <typename T>
static int cholqr_factor(Ctx* ctx, int64_t m, int64_t k, const T* A, bool cond_check, bool rows_sharded, T* G, int* chol_fail) {
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * k * k, ctx->stream));
    RLB_CHECK(gemm_tn<T>(ctx, m, k, k, 1.0, A, m, A, m, 0.0, G, k, /*upper_only=*/1));
    return cholqr_finish<T>(ctx, k, cond_check, rows_sharded, G, chol_fail);
}
// G (upper triangle of the local Gram matrix) -> R = chol(G) in place, with the reference's failure / condition rules (rl_orth.hh:81-93)
template <typename T>
static int cholqr_finish(Ctx* ctx, int64_t k, bool cond_check, bool rows_sharded, T* G, int* chol_fail) {
    if (rows_sharded) RLB_CHECK(allreduce_sum<T>(ctx, G, k * k));
    int info = 0;
    RLB_CHECK(potrf_blocked<T>(ctx, k, G, k, &info));
    if (info != 0) { if (chol_fail) *chol_fail = 1; return 1; }
    if (cond_check) {
        if (k > 256) RLB_CHECK(tri_op<T>(ctx, 1, k, k, G, k, G, k));
        double cond = 0;
        RLB_CHECK(cond_of_square<T>(ctx, k, G, &cond));
        if (cond > 1.0 / std::sqrt((double)std::numeric_limits<T>::epsilon())) return 1;
    }
    return 0;
}

static inline bool fuse_ok(const rlb200_stack_opts& o) { return (o.reserved & 1) == 0; }

template <typename T>
int stab_call(Ctx* ctx, int kind, int64_t m, int64_t k, T* A, bool cond_check, bool rows_sharded, int* chol_fail) {
    RLB_REQUIRE(ctx, m >= 0 && k >= 0);
    switch (kind) {
        case RLB200_STAB_CHOLQRQ: return cholqrq<T>(ctx, m, k, A, cond_check, rows_sharded, chol_fail);
        case RLB200_STAB_PLUL: {
            if (rows_sharded && ctx->allreduce) {
                ctx->err = "PLUL on a row-sharded iterate needs a cross-rank pivot search (max-allreduce); not offered yet";
                return RLB200_ERR_UNSUPPORTED;
            }
            ArenaScope as(ctx);
            void* ws = arena_push(ctx, plul_ws_bytes(ctx, k)); RLB_ALLOC(ctx, ws);
            return plul<T>(ctx, m, k, A, m, ws);   // rl_orth.hh:211-230: always returns 0
        }
        case RLB200_STAB_HQRQ: {
            if (rows_sharded && ctx->allreduce) return hqrq_tsqr<T>(ctx, m, k, A);
            return hqrq<T>(ctx, m, k, A);
        }
    }
    RLB_REQUIRE(ctx, !"unknown stabiliser kind");
    return RLB200_ERR_ARG;
}

// tall products over the m x n data matrix: the fp64 engine is selectable (DMMA fp64 pipe / tcgen05 int8 digit slices, ozaki.cu);
// Gram matrices of CholQR always stay on the fp64 pipe (their conditioning is squared).
// the digit-slice engine pays off on tall operands only (slicing + launch overheads); shorter products stay on the fp64 pipe
constexpr int64_t kI8MinRows = 16384;
template <typename T>
static int tall_nn(Ctx* ctx, int64_t m, int64_t N, int64_t K, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C,
                   int64_t ldc) {
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= kI8MinRows) {
        if (ozaki2_nn_ok(ctx, m, N, K, A, lda * (int64_t)sizeof(T), C)) return ozaki2_gemm_nn<T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
        return ozaki_gemm_nn<T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
    }
    return gemm_nn<T>(ctx, m, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}
template <typename T>
static int tall_tn(Ctx* ctx, int64_t m, int64_t N1, int64_t N2, double alpha, const T* A, int64_t lda, const T* B, int64_t ldb, double beta, T* C,
                   int64_t ldc, double* a_sumsq_out = nullptr) {
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= kI8MinRows) {
        if (ozaki2_tn_ok(ctx, m, N1, N2, A, lda * (int64_t)sizeof(T))) return ozaki2_gemm_tn<T>(ctx, m, N1, N2, alpha, A, lda, B, ldb, beta, C, ldc, a_sumsq_out);
        return ozaki_gemm_tn<T>(ctx, m, N1, N2, alpha, A, lda, B, ldb, beta, C, ldc, a_sumsq_out);
    }
    return gemm_tn<T>(ctx, m, N1, N2, alpha, A, lda, B, ldb, beta, C, ldc, 0, a_sumsq_out);
}

// Z = A^T Y and G = Y^T Y (upper triangle) in ONE sweep: on the int8-slice engine the Gram tiles ride in the launches of the A^T Y
// product, fed from the same digits of Y (the A^T Y pass is bound by slicing, so they are nearly free).  The fused pass runs with 7
// digits for fp64 (54 bits: the Gram matrix squares the conditioning and potrf must fail exactly where the reference's does).
template <typename T>
static int tall_tn_gram(Ctx* ctx, int64_t m, int64_t n, int64_t k, const T* A, int64_t lda, const T* Y, int64_t ldy, T* Z, int64_t ldz, T* G,
                        double* a_sumsq_out = nullptr, bool full_pairs = false) {
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= kI8MinRows && k >= 64 && getenv("RLB200_NO_GRAM_FUSION") == nullptr) {
        // fused engine: 7 digits are produced, the Gram tiles use all 28 digit pairs, the A^T Y tiles the 21 of the first 6 anti-diagonals
        // unless full_pairs asks for all 28
        if (ozaki2_tn_ok(ctx, m, n, k, A, lda * (int64_t)sizeof(T)))
            return ozaki2_gemm_tn<T>(ctx, m, n, k, 1.0, A, lda, Y, ldy, 0.0, Z, ldz, a_sumsq_out, G, k, full_pairs);
        const int old = ctx->i8_digits;
        if (!old && sizeof(T) == 8) ctx->i8_digits = 7;
        const int rc = ozaki_gemm_tn<T>(ctx, m, n, k, 1.0, A, lda, Y, ldy, 0.0, Z, ldz, a_sumsq_out, false, G, k);
        ctx->i8_digits = old;
        return rc;
    }
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * k * k, ctx->stream));
    RLB_CHECK(gemm_tn<T>(ctx, m, k, k, 1.0, Y, ldy, Y, ldy, 0.0, G, k, /*upper_only=*/1));
    return tall_tn<T>(ctx, m, n, k, 1.0, A, lda, Y, ldy, 0.0, Z, ldz, a_sumsq_out);
}

// ------------------------------------------------------------------------------------------------
// RS  (rl_rs.hh:116-178)
// ------------------------------------------------------------------------------------------------
// full-matrix next state of a DenseDist sample (dense_skops.hh:169-182)
static void dense_next_state(int64_t n_rows, int64_t n_cols, uint32_t state[6]) {
    int64_t major = std::max(n_rows, n_cols), minor = std::min(n_rows, n_cols);
    Ctr128 c;
    for (int i = 0; i < 4; ++i) c.v[i] = state[i];
    c = ctr_add(c, (uint64_t)(((major + 3) / 4) * minor));
    for (int i = 0; i < 4; ++i) state[i] = c.v[i];
}

template <typename T>
int rs_call(Ctx* ctx, int64_t m, int64_t n, const T* A, int64_t k, T* Omega, T* work, uint32_t state[6], const rlb200_stack_opts& o) {
    RLB_REQUIRE(ctx, m > 0 && n > 0 && k > 0);
    const int64_t p = o.passes_over_data, q = o.passes_per_stab;
    RLB_REQUIRE(ctx, p >= 0 && (p == 0 || q >= 1));
    RLB_REQUIRE(ctx, p == 0 || work != nullptr);
    OzConstScope a_const(ctx, A);
    const bool sharded = ctx->m_global >= 0;
    const int64_t mg = sharded ? ctx->m_global : m;
    int64_t p_done = 0;
    T* Omega_1 = work;
    if (p % 2 == 0) {
        // Omega = fill_dense(DenseDist(n, k))  (:132-135); replicated on every shard
        RLB_CHECK(fill_dense_unpacked<T>(ctx, n, k, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, n, k, 0, 0, Omega, state));
    } else {
        // Omega_1 = fill_dense(DenseDist(m, k)) (:136-139); each shard generates only its row block of the
        // m_global x k operator through the sub-matrix counter rule, then the state advances as for the whole matrix.
        uint32_t st_full[6];
        std::memcpy(st_full, state, sizeof st_full);
        if (mg > k) {   // tall: natural layout is column-major, rows can be sliced
            uint32_t st_sub[6];
            std::memcpy(st_sub, state, sizeof st_sub);
            RLB_CHECK(fill_dense_unpacked<T>(ctx, mg, k, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, m, k,
                                             sharded ? ctx->row_offset : 0, 0, Omega_1, st_sub));
        } else {
            RLB_REQUIRE(ctx, !sharded);   // wide/square operator buffers are reinterpreted, cannot be row-sliced
            uint32_t st_sub[6];
            std::memcpy(st_sub, state, sizeof st_sub);
            RLB_CHECK(fill_dense_unpacked<T>(ctx, mg, k, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, mg, k, 0, 0, Omega_1, st_sub));
        }
        dense_next_state(mg, k, st_full);
        std::memcpy(state, st_full, sizeof st_full);
        // Omega = A^T Omega_1 (:142)
        RLB_CHECK(tall_tn<T>(ctx, m, n, k, 1.0, A, m, Omega_1, m, 0.0, Omega, n));
        if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, Omega, n * k));
        ++p_done;
        if (p_done % q == 0) {
            int rc = stab_call<T>(ctx, o.stab, n, k, Omega, o.cond_check, false, nullptr);
            if (rc) return rc < 0 ? rc : 1;
        }
    }
    while (p - p_done > 0) {
        // Omega_1 = A Omega (:153)
        RLB_CHECK(tall_nn<T>(ctx, m, k, n, 1.0, A, m, Omega, n, 0.0, Omega_1, m));
        ++p_done;
        T* Rfold = nullptr;
        ArenaScope as_fold(ctx);
        if (p_done % q == 0) {
            if (o.stab == RLB200_STAB_CHOLQRQ && fuse_ok(o)) {
                // stabilise Omega_1 = Q R implicitly: Omega = A^T Q = (A^T Omega_1) R^-1; Omega_1 itself is scratch (:130) and is never read again
                // (the Gram matrix and A^T Omega_1 come out of one sweep; the Cholesky factor, and its failure code, follow)
                Rfold = as_fold.take<T>(k * k); RLB_ALLOC(ctx, Rfold);
                // all digit pairs: Omega = (A^T Omega_1) R^-1 amplifies the error of A^T Omega_1 by cond(R) = cond(A Omega)
                RLB_CHECK(tall_tn_gram<T>(ctx, m, n, k, A, m, Omega_1, m, Omega, n, Rfold, nullptr, /*full_pairs=*/true));
                int rc = cholqr_finish<T>(ctx, k, o.cond_check, sharded, Rfold, nullptr);
                if (rc) return rc < 0 ? rc : 1;
            } else {
                int rc = stab_call<T>(ctx, o.stab, m, k, Omega_1, o.cond_check, sharded, nullptr);
                if (rc) return rc < 0 ? rc : 1;
            }
        }
        // Omega = A^T Omega_1 (:165)
        if (!Rfold) RLB_CHECK(tall_tn<T>(ctx, m, n, k, 1.0, A, m, Omega_1, m, 0.0, Omega, n));
        if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, Omega, n * k));
        if (Rfold) RLB_CHECK(trsm_right_upper<T>(ctx, n, k, Rfold, k, Omega, n));
        ++p_done;
        if (p_done % q == 0) {
            int rc = stab_call<T>(ctx, o.stab, n, k, Omega, o.cond_check, false, nullptr);
            if (rc) return rc < 0 ? rc : 1;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// RF  (rl_rf.hh:106-137)
// ------------------------------------------------------------------------------------------------
template <typename T>
int rf_call(Ctx* ctx, int64_t m, int64_t n, const T* A, int64_t k, T* Q, uint32_t state[6], const rlb200_stack_opts& o) {
    RLB_REQUIRE(ctx, m > 0 && n > 0 && k > 0);
    OzConstScope a_const(ctx, A);
    ArenaScope as(ctx);
    T* Omega = as.take<T>(n * k); RLB_ALLOC(ctx, Omega);
    // RS's m x k scratch (Omega_1) lives in Q, which is overwritten afterwards anyway
    int rc = rs_call<T>(ctx, m, n, A, k, Omega, Q, state, o);
    if (rc) return rc < 0 ? rc : 1;                                           // :118-120
    RLB_CHECK(tall_nn<T>(ctx, m, k, n, 1.0, A, m, Omega, n, 0.0, Q, m));      // :123
    rc = stab_call<T>(ctx, o.orth_rf, m, k, Q, o.cond_check, ctx->m_global >= 0, nullptr);   // :129
    if (rc) return rc < 0 ? rc : 2;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// QB  (rl_qb.hh:133-268)
// ------------------------------------------------------------------------------------------------
template <typename T>
static int fro_norm(Ctx* ctx, const T* A, int64_t m, int64_t n, int64_t lda, bool rows_sharded, double* out) {
    ArenaScope as(ctx);
    double* part = as.take<double>(sumsq_ws_doubles(ctx) + 1); RLB_ALLOC(ctx, part);
    double* res = part + sumsq_ws_doubles(ctx);
    RLB_CHECK(sumsq<T>(ctx, A, m, n, lda, part, res));
    if (rows_sharded) RLB_CHECK(allreduce_sum<double>(ctx, res, 1));
    double ss = 0;
    RLB_CHECK(read_scalar<double>(ctx, res, &ss));
    *out = std::sqrt(ss);
    return 0;
}

// util::orthogonality_check (rl_util.hh:467-496)
template <typename T>
static int orth_check(Ctx* ctx, int64_t m, int64_t k, const T* Q, bool rows_sharded, bool* lost) {
    ArenaScope as(ctx);
    T* G = as.take<T>(k * k); RLB_ALLOC(ctx, G);
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * k * k, ctx->stream));
    RLB_CHECK(gemm_tn<T>(ctx, m, k, k, 1.0, Q, m, Q, m, 0.0, G, k, 1));
    if (rows_sharded) RLB_CHECK(allreduce_sum<T>(ctx, G, k * k));
    std::vector<T> g(k * k);
    RLB_CUDA_OK(ctx, cudaMemcpyAsync(g.data(), G, sizeof(T) * k * k, cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    double ss = 0;
    for (int64_t j = 0; j < k; ++j)
        for (int64_t i = 0; i <= j; ++i) { double v = (double)g[i + j * k] - (i == j ? 1.0 : 0.0); ss += v * v; }
    const double tol = sizeof(T) == 8 ? 1.0e-10 : 1.0e-2;
    *lost = std::sqrt(ss) / std::sqrt((double)k) > tol;
    return 0;
}

template <typename T>
int qb_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t* k_io, int64_t b_sz, T tol_in, T* Q, T* BT, T* Acpy, uint32_t state[6],
            const rlb200_stack_opts& o) {
    RLB_REQUIRE(ctx, m > 0 && n > 0 && k_io && *k_io > 0 && b_sz > 0);
    const bool sharded = ctx->m_global >= 0;
    const int64_t k = *k_io;
    int64_t curr_sz = 0, next_sz = 0;
    const T tol = std::max(tol_in, (T)100 * std::numeric_limits<T>::epsilon());   // :149
    T norm_B = 0, prev_err = 0, approx_err = 0;
    ArenaScope as(ctx);
    // ||A||_F (:168) is accumulated inside the first block's A^T*Q_i pass (the A tiles are on chip there anyway), which
    // saves one full sweep over A; it is only consumed after that product (:225).
    T norm_A = 0;
    double* nA_dev = as.take<double>(1); RLB_ALLOC(ctx, nA_dev);
    const bool multi_block = b_sz < k;
    T* A_work = A;
    T* QtQi = nullptr;
    if (multi_block) {
        // the reference deflates a copy of A (:162,171,260)
        if (!Acpy) { Acpy = as.take<T>((size_t)m * n); RLB_ALLOC(ctx, Acpy); }
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(Acpy, A, sizeof(T) * m * n, cudaMemcpyDeviceToDevice, ctx->stream));
        A_work = Acpy;
        QtQi = as.take<T>((size_t)k * b_sz); RLB_ALLOC(ctx, QtQi);
    }
    while (curr_sz < k) {
        b_sz = std::min(b_sz, k - curr_sz);
        next_sz = curr_sz + b_sz;
        T* Q_i = Q + m * curr_sz;
        T* BT_i = BT + n * curr_sz;
        OzConstScope a_const(ctx, A_work);     // A_work is constant until the deflation at the end of the block
        int rc = rf_call<T>(ctx, m, n, A_work, b_sz, Q_i, state, o);             // :191
        if (rc < 0) return rc;
        if (rc) { *k_io = curr_sz; return 6; }
        if (o.orth_check) {                                                       // :199-207
            bool lost = false;
            RLB_CHECK(orth_check<T>(ctx, m, b_sz, Q_i, sharded, &lost));
            if (lost) { *k_io = curr_sz; return 4; }
        }
        if (curr_sz != 0) {                                                       // :210-215
            // (the reference stores QtQi with ld = next_sz; a compact ld = curr_sz keeps the allreduce contiguous)
            RLB_CHECK(gemm_tn<T>(ctx, m, curr_sz, b_sz, 1.0, Q, m, Q_i, m, 0.0, QtQi, curr_sz, 0));
            if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, QtQi, curr_sz * b_sz));
            RLB_CHECK(gemm_nn<T>(ctx, m, b_sz, curr_sz, -1.0, Q, m, QtQi, curr_sz, 1.0, Q_i, m));
            int rc2 = stab_call<T>(ctx, o.orth_qb, m, b_sz, Q_i, o.cond_check, sharded, nullptr);
            if (rc2 < 0) return rc2;   // (the reference ignores a numeric failure here, :214)
        }
        RLB_CHECK(tall_tn<T>(ctx, m, n, b_sz, 1.0, A_work, m, Q_i, m, 0.0, BT_i, n, curr_sz == 0 ? nA_dev : nullptr));   // :218
        if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, BT_i, n * b_sz));
        if (curr_sz == 0) {
            if (sharded) RLB_CHECK(allreduce_sum<double>(ctx, nA_dev, 1));
            double ss = 0;
            RLB_CHECK(read_scalar<double>(ctx, nA_dev, &ss));
            norm_A = (T)std::sqrt(ss);
        }
        double nBi = 0;
        RLB_CHECK(fro_norm<T>(ctx, BT_i, n, b_sz, n, false, &nBi));               // :221
        norm_B = (T)std::hypot((T)norm_B, (T)nBi);                                // :222
        prev_err = approx_err;
        approx_err = std::sqrt(std::abs(norm_A - norm_B)) * (std::sqrt(norm_A + norm_B) / norm_A);   // :225
        if (curr_sz > 0 && approx_err > prev_err) { *k_io = curr_sz; return 2; }  // :228-234
        if (o.orth_check) {                                                       // :236-244
            bool lost = false;
            RLB_CHECK(orth_check<T>(ctx, m, next_sz, Q, sharded, &lost));
            if (lost) { *k_io = curr_sz; return 5; }
        }
        curr_sz += b_sz;
        if (approx_err < tol) { *k_io = curr_sz; return 0; }                      // :250-256
        // A_work -= Q_i BT_i^T (:260) — only needed when another block follows (the reference also runs it after
        // the last block, where its result is never read)
        if (curr_sz < k) {
            ctx->oz_row.valid = false; ctx->oz_col.valid = false;
            RLB_CHECK(gemm_nt<T>(ctx, m, n, b_sz, -1.0, Q_i, m, BT_i, n, 1.0, A_work, m));
        }
    }
    return 3;
}

// ------------------------------------------------------------------------------------------------
// RSVD  (rl_rsvd.hh:113-154)
// ------------------------------------------------------------------------------------------------
// RSVD with a single QB block and CholQRQ as RF's orthogonaliser (the configuration BASELINE.json names), with Q = Y R^-1 never
// formed: B^T = A^T Q = (A^T Y) R^-1 and U = Q W = Y (R^-1 W).  Same quantities, return codes, k and RNG advancement as
// RF::call -> QB::call -> RSVD::call (rl_rf.hh:106-137, rl_qb.hh:133-268, rl_rsvd.hh:137-148); two m x k triangular solves fewer.
template <typename T>
static int rsvd_single_block_fused(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t* k_io, T tol_in, T* U, T* S, T* V, uint32_t state[6],
                                   const rlb200_stack_opts& o, int* qb_code) {
    const bool sharded = ctx->m_global >= 0;
    const int64_t k = *k_io;
    const T tol = std::max(tol_in, (T)100 * std::numeric_limits<T>::epsilon());       // rl_qb.hh:149
    OzConstScope a_const(ctx, A);
    ArenaScope as(ctx);
    T* Omega = as.take<T>(n * k); RLB_ALLOC(ctx, Omega);
    T* R = as.take<T>(k * k); RLB_ALLOC(ctx, R);
    T* W = as.take<T>(k * k); RLB_ALLOC(ctx, W);
    T* Rinv = as.take<T>(k * k); RLB_ALLOC(ctx, Rinv);
    T* M = as.take<T>(k * k); RLB_ALLOC(ctx, M);
    double* nA_dev = as.take<double>(1); RLB_ALLOC(ctx, nA_dev);
    if (qb_code) *qb_code = 0;
    // RF (rl_rf.hh:118-129)
    int rc = rs_call<T>(ctx, m, n, A, k, Omega, U, state, o);
    if (rc < 0) return rc;
    if (!rc) {
        RLB_CHECK(tall_nn<T>(ctx, m, k, n, 1.0, A, m, Omega, n, 0.0, U, m));           // Y = A Omega
        // Y = Q R (Q implicit) and B^T = A^T Q = (A^T Y) R^-1 (rl_qb.hh:218): the Gram matrix of Y, A^T Y and ||A||_F (:168) come out of
        // one sweep of A and Y; the Cholesky factor, and the failure code of RF's orthogonaliser (rl_rf.hh:129), follow
        RLB_CHECK(tall_tn_gram<T>(ctx, m, n, k, A, m, U, m, V, n, R, nA_dev));
        rc = cholqr_finish<T>(ctx, k, o.cond_check, sharded, R, nullptr);
        if (rc < 0) return rc;
    }
    if (rc) { *k_io = 0; if (qb_code) *qb_code = 6; return 0; }                        // rl_qb.hh:191-197 -> rl_rsvd.hh:137
    if (sharded) { RLB_CHECK(allreduce_sum<T>(ctx, V, n * k)); RLB_CHECK(allreduce_sum<double>(ctx, nA_dev, 1)); }
    RLB_CHECK(trsm_right_upper<T>(ctx, n, k, R, k, V, n));
    double ss = 0, nB = 0;
    RLB_CHECK(read_scalar<double>(ctx, nA_dev, &ss));
    const T norm_A = (T)std::sqrt(ss);
    RLB_CHECK(fro_norm<T>(ctx, V, n, k, n, false, &nB));                               // :221
    const T norm_B = (T)std::hypot((T)0, (T)nB);                                       // :222
    const T approx_err = std::sqrt(std::abs(norm_A - norm_B)) * (std::sqrt(norm_A + norm_B) / norm_A);   // :225
    if (qb_code) *qb_code = approx_err < tol ? 0 : 3;                                  // :250-256 / :267
    // SVD of B^T and U = Q W = Y (R^-1 W) (rl_rsvd.hh:146-148)
    void* ws = arena_push(ctx, svd_ws_bytes(n, k, sizeof(T))); RLB_ALLOC(ctx, ws);
    RLB_CHECK(svd_tall<T>(ctx, n, k, V, n, S, W, ws, nullptr));
    RLB_CHECK(trtri_upper<T>(ctx, (int)k, R, (int)k, Rinv));
    RLB_CHECK(gemm_nn<T>(ctx, k, k, k, 1.0, Rinv, k, W, k, 0.0, M, k));
    // U = Y (R^-1 W) in place (rl_rsvd.hh:148 with Q = Y R^-1 never formed); on the int8-slice engine when it is selected
    // (fp64: 7 digits - the columns of R^-1 W grow like cond(Y), and with 6 digits the orthogonality of U would stop at ~2e-10)
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= kI8MinRows && k >= 64) {
        const int old = ctx->i8_digits;
        if (!old && sizeof(T) == 8) ctx->i8_digits = 7;
        // Chunk-wise through a scratch copy of 2^19 rows of Y when it can be had: the out-of-place product then runs on the persistent
        // kernel (the tail of a K = 256 tile - 8 K steps - overlaps the next tile; the in-place product pays set-up, cluster barrier and
        // epilogue per tile: 89 ms at 2^24 x 256 against ~55 ms incl. the copies).  Same digits, same arithmetic, same result.
        int rc_u = 1;
        {
            const int64_t cr = std::min<int64_t>(m, (int64_t)1 << 19);
            ArenaScope as_u(ctx);
            T* Ys = (ctx->i8_fused && m > cr && ozaki2_nn_ok(ctx, cr, k, k, U /*alignment only*/, cr * (int64_t)sizeof(T), nullptr))
                        ? as_u.take<T>((size_t)cr * k) : nullptr;
            if (Ys) {
                rc_u = 0;
                for (int64_t r0 = 0; r0 < m && rc_u >= 0; r0 += cr) {
                    const int64_t rows = std::min(cr, m - r0);
                    cudaError_t ce = cudaMemcpy2DAsync(Ys, cr * sizeof(T), U + r0, m * sizeof(T), rows * sizeof(T), k, cudaMemcpyDeviceToDevice, ctx->stream);
                    if (ce != cudaSuccess) { ctx->err = std::string("cudaMemcpy2DAsync: ") + cudaGetErrorString(ce); rc_u = RLB200_ERR_CUDA; break; }
                    rc_u = ozaki2_gemm_nn<T>(ctx, rows, k, k, 1.0, Ys, cr, M, k, 0.0, U + r0, m);
                }
                if (rc_u >= 0) { cudaError_t ce = cudaStreamSynchronize(ctx->stream); if (ce != cudaSuccess) rc_u = RLB200_ERR_CUDA; }   // scratch leaves scope
            } else {
                cudaGetLastError();
            }
        }
        if (rc_u == 1)
            rc_u = ozaki2_nn_ok(ctx, m, k, k, U, m * (int64_t)sizeof(T), U) ? ozaki2_gemm_nn<T>(ctx, m, k, k, 1.0, U, m, M, k, 0.0, U, m)
                                                                           : ozaki_gemm_nn<T>(ctx, m, k, k, 1.0, U, m, M, k, 0.0, U, m);
        ctx->i8_digits = old;
        RLB_CHECK(rc_u);
    } else RLB_CHECK(gemm_nn_inplace<T>(ctx, m, k, k, 1.0, U, m, M, k));
    return 0;
}

template <typename T>
int rsvd_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t* k_io, T tol, T* U, T* S, T* V, T* Acpy, uint32_t state[6],
              const rlb200_stack_opts& o, int* qb_code) {
    // :128-132
    RLB_REQUIRE(ctx, m >= 0);
    RLB_REQUIRE(ctx, n >= 0);
    RLB_REQUIRE(ctx, k_io && *k_io > 0);
    RLB_REQUIRE(ctx, tol >= (T)0);
    RLB_REQUIRE(ctx, !(A == nullptr && m > 0 && n > 0));
    RLB_REQUIRE(ctx, m > 0 && n > 0 && *k_io <= n && o.block_sz > 0);
    // Q lives in U, BT in V; the SVD of BT (n x k) overwrites V with its left singular vectors, which is exactly
    // what the reference returns in V (gesdd's U argument, :146)
    if (o.block_sz >= *k_io && *k_io <= 256 && o.orth_rf == RLB200_STAB_CHOLQRQ && !o.orth_check && fuse_ok(o))
        return rsvd_single_block_fused<T>(ctx, m, n, A, k_io, tol, U, S, V, state, o, qb_code);
    int rc = qb_call<T>(ctx, m, n, A, k_io, o.block_sz, tol, U, V, Acpy, state, o);   // :137 (code ignored by the reference)
    if (rc < 0) return rc;
    if (qb_code) *qb_code = rc;
    const int64_t k = *k_io;
    if (k == 0) return 0;
    ArenaScope as(ctx);
    T* W = as.take<T>(k * k); RLB_ALLOC(ctx, W);
    void* ws = arena_push(ctx, svd_ws_bytes(n, k, sizeof(T))); RLB_ALLOC(ctx, ws);
    RLB_CHECK(svd_tall<T>(ctx, n, k, V, n, S, W, ws, nullptr));                   // :146
    // U = Q * UT_buf^T = Q * W (:148), in place
    if (k <= 256) {
        RLB_CHECK(gemm_nn_inplace<T>(ctx, m, k, k, 1.0, U, m, W, k));
    } else {
        T* tmp = as.take<T>((size_t)m * k); RLB_ALLOC(ctx, tmp);
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(tmp, U, sizeof(T) * m * k, cudaMemcpyDeviceToDevice, ctx->stream));
        RLB_CHECK(gemm_nn<T>(ctx, m, k, k, 1.0, tmp, m, W, k, 0.0, U, m));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CQRRPT  (rl_cqrrpt.hh:146-391)
// ------------------------------------------------------------------------------------------------
template <typename T>
static int read_diag(Ctx* ctx, const T* M, int64_t ld, int64_t k, std::vector<T>& out) {
    out.resize((size_t)k);
    if (k == 0) return 0;
    RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(out.data(), sizeof(T), M, (ld + 1) * sizeof(T), sizeof(T), k, cudaMemcpyDeviceToHost, ctx->stream));
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// A <- A * R^-1 (blas::trsm Right/Upper, rl_cqrrpt.hh:306,342) and the Gram matrix of CholQR (syrk, :314) for tall A.
// On the int8-slice engine the solve is one in-place tall product with the explicit inverse (R^-1 = I * R^-1 through the blocked
// solver on the k x k identity), which keeps all O(m k^2) work on tcgen05; the triangular zeros of R^-1 are skipped.
// fp64 QR factors must pass the reference's eps^0.75 acceptance tests (test_cqrrpt.cc:98-104): 7 digits (54 bits) unless the caller
// fixed the digit count; fp32 storage keeps its default (4 digits, 30 bits).
struct OzQrDigits {
    Ctx* ctx; int old;
    OzQrDigits(Ctx* c, size_t elem) : ctx(c), old(c->i8_digits) { if (!old && elem == 8) c->i8_digits = 7; }
    ~OzQrDigits() { ctx->i8_digits = old; }
};
template <typename T>
static int tall_right_solve(Ctx* ctx, int64_t m, int64_t k, const T* R, int64_t ldr, T* A, int64_t lda) {
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= 8192 && k >= 64 && k <= 16384) {
        OzQrDigits dg(ctx, sizeof(T));
        ArenaScope as(ctx);
        T* Rinv = as.take<T>((size_t)k * k); RLB_ALLOC(ctx, Rinv);
        RLB_CUDA_OK(ctx, cudaMemsetAsync(Rinv, 0, sizeof(T) * k * k, ctx->stream));
        RLB_CHECK(set_upper_diag<T>(ctx, k, Rinv, k, (T)1, false));
        RLB_CHECK(trsm_right_upper<T>(ctx, k, k, R, ldr, Rinv, k));
        return ozaki_gemm_nn<T>(ctx, m, k, k, 1.0, A, lda, Rinv, k, 0.0, A, lda, /*b_upper_tri=*/true);
    }
    return trsm_right_upper<T>(ctx, m, k, R, ldr, A, lda);
}
template <typename T>
static int tall_gram_upper(Ctx* ctx, int64_t m, int64_t k, const T* A, int64_t lda, T* G, int64_t ldg) {
    if (ctx->fp64_engine == RLB200_FP64_I8SLICES && m >= 8192 && k >= 64) {
        OzQrDigits dg(ctx, sizeof(T));
        return ozaki_gemm_tn<T>(ctx, m, k, k, 1.0, A, lda, A, lda, 0.0, G, ldg, nullptr, /*upper_only=*/true);
    }
    return gemm_tn<T>(ctx, m, k, k, 1.0, A, lda, A, lda, 0.0, G, ldg, 1);
}

template <typename T>
int cqrrpt_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J_dev, T d_factor, T eps_user, int64_t nnz,
                int64_t* rank_out, uint32_t state[6]) {
    // :161-168
    RLB_REQUIRE(ctx, m >= 0);
    RLB_REQUIRE(ctx, n >= 0);
    RLB_REQUIRE(ctx, lda >= m);
    RLB_REQUIRE(ctx, ldr >= n);
    RLB_REQUIRE(ctx, d_factor >= (T)1.0);
    RLB_REQUIRE(ctx, !(A == nullptr && m > 0 && n > 0));
    RLB_REQUIRE(ctx, !(R == nullptr && n > 0));
    RLB_REQUIRE(ctx, !(J_dev == nullptr && n > 0));
    RLB_REQUIRE(ctx, rank_out != nullptr);
    const bool sharded = ctx->m_global >= 0;
    const int64_t mg = sharded ? ctx->m_global : m;
    *rank_out = 0;
    if (n == 0) return 0;
    RLB_REQUIRE(ctx, mg > 0);
    int64_t k = n;
    const int64_t d = (int64_t)(d_factor * (T)n);                                             // :197
    RLB_REQUIRE(ctx, d <= mg);                                                                // SparseDist(d, m): the operator must be wide
    const T eps_initial = (T)2 * (T)std::pow((double)std::numeric_limits<T>::epsilon(), 0.95);   // :199
    ArenaScope as(ctx);
    T* A_hat = as.take<T>((size_t)d * n); RLB_ALLOC(ctx, A_hat);
    T* tau = as.take<T>((size_t)n); RLB_ALLOC(ctx, tau);
    // phase times in the reference's order (:371-384): saso, qrcp, rank_reveal, cholqr, a_mod_piv, a_mod_trsm, rest, total
    PhaseTimer pt(ctx);
    long long t_saso = 0, t_qrcp = 0, t_rank = 0, t_chol = 0, t_piv = 0, t_trsm = 0;
    // SASO (:214-221); state <- S.next_state
    RLB_CHECK(sketch_sparse_left<T>(ctx, d, mg, nnz, d, n, m, (T)1, 0, 0, A, lda, (T)0, A_hat, d, state));
    t_saso = pt.lap();
    // QRCP of the sketch: geqp3 (:247), BQRRP with the reference's block ratio (:232-244) or hqrrp (:230-231)
    if (ctx->cqrrpt_qrcp == RLB200_CQRRPT_QRCP_BQRRP) {
        if (sharded) { ctx->err = "CQRRPT with qrcp = bqrrp is not offered on a row-sharded context"; return RLB200_ERR_UNSUPPORTED; }
        const T ratio = n <= 2000 ? (T)1 : (n <= 8000 ? (T)0.5 : (T)1 / (T)32);
        const int64_t bsz = (int64_t)((T)n * ratio);
        int64_t rank_b = 0;
        const double tol_saved = ctx->bqrrp_tol;
        ctx->bqrrp_tol = 0.0;                                                                  // a fresh BQRRP object: tol = eps
        const int rcb = bqrrp_call<T>(ctx, d, n, A_hat, d, (T)1, bsz, RLB200_QRCP_LUQR, RLB200_QRTALL_GEQRF, tau, J_dev, &rank_b, state);
        ctx->bqrrp_tol = tol_saved;
        if (rcb < 0) return rcb;
    } else if (ctx->cqrrpt_qrcp == RLB200_CQRRPT_QRCP_HQRRP) {
        // hqrrp(d, n, A_hat, d, J, tau, nb_alg, oversampling, panel_pivoting, use_cholqr, state, nullptr) (:230-231)
        if (sharded) { ctx->err = "CQRRPT with qrcp = hqrrp is not offered on a row-sharded context"; return RLB200_ERR_UNSUPPORTED; }
        const int rch = hqrrp_call<T>(ctx, d, n, A_hat, d, J_dev, tau, ctx->cqrrpt_nb_alg, ctx->cqrrpt_oversampling, ctx->cqrrpt_panel_pivoting,
                                      ctx->cqrrpt_use_cholqr, state);
        if (rch < 0) return rch;
    } else {
        ArenaScope as2(ctx);
        void* ws = arena_push(ctx, qrcp_ws_bytes(n)); RLB_ALLOC(ctx, ws);
        RLB_CHECK(qr_small<T>(ctx, true, d, n, A_hat, d, J_dev, tau, ws));
    }
    t_qrcp = pt.lap();
    std::vector<T> dg;
    RLB_CHECK(read_diag<T>(ctx, A_hat, d, n, dg));
    if (!dg[0]) return 0;                                                                     // :256-261 all-zero input
    for (int64_t i = 0; i < n; ++i) {                                                         // :267-272
        if (std::abs(dg[i]) / std::abs(dg[0]) < eps_initial) { k = i; break; }
    }
    *rank_out = k;
    int64_t new_rank = k;
    RLB_CHECK(tri_op<T>(ctx, 0, k, k, A_hat, d, R, ldr));                                     // lacpy(Upper) :284
    t_rank = pt.lap();
    // col_swap (:291-292)
    {
        std::vector<int64_t> J((size_t)n);
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(J.data(), J_dev, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        for (auto& v : J) v -= 1;
        RLB_CHECK(col_permute<T>(ctx, m, n, A, lda, J.data()));
    }
    t_piv = pt.lap();
    for (int64_t i = 0; i < k; ++i) if (dg[i] == (T)0) return 1;                              // diag_is_nonzero :300-305
    RLB_CHECK(tall_right_solve<T>(ctx, m, k, R, ldr, A, lda));                                // :306
    t_trsm = pt.lap();
    // CholQR (:314-339): the Gram / Cholesky factor is built in scratch so that only the upper triangle of R is written
    T* G = as.take<T>((size_t)k * k); RLB_ALLOC(ctx, G);
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * k * k, ctx->stream));
    RLB_CHECK(tall_gram_upper<T>(ctx, m, k, A, lda, G, k));
    if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, G, k * k));
    int info = 0;
    RLB_CHECK(potrf_blocked<T>(ctx, k, G, k, &info));
    RLB_CHECK(tri_op<T>(ctx, 0, k, k, G, k, R, ldr));
    if (info != 0) {                                                                          // :311-336 a-posteriori rank estimate
        std::vector<T> rd;
        RLB_CHECK(read_diag<T>(ctx, G, k, k, rd));
        T running_max = rd[0], running_min = rd[0];
        const T cond_threshold = std::sqrt(eps_user / std::numeric_limits<T>::epsilon());
        for (int64_t i = 0; i < k; ++i) {
            const T curr = std::abs(rd[i]);
            running_max = std::max(running_max, curr);
            running_min = std::min(running_min, curr);
            if ((running_min * cond_threshold < running_max) && i > 1) { new_rank = i - 1; break; }
        }
    }
    *rank_out = new_rank;                                                                     // :339
    RLB_CHECK(tall_right_solve<T>(ctx, m, new_rank, R, ldr, A, lda));                         // :342
    if (ctx->cqrrpt_orth) {
        // orthogonalization mode (:343-368): R stays the Cholesky factor; the trailing columns complete the orthonormal set
        const int64_t cols = n - new_rank;
        if (cols > 0) {
            if (sharded) { ctx->err = "CQRRPT orthogonalization mode with rank < n is not offered on a row-sharded context"; return RLB200_ERR_UNSUPPORTED; }
            RLB_REQUIRE(ctx, cols <= 256 && m > cols);
            T* Gc = as.take<T>((size_t)m * cols); RLB_ALLOC(ctx, Gc);
            T* tmp = as.take<T>((size_t)std::max<int64_t>(new_rank, 1) * cols); RLB_ALLOC(ctx, tmp);
            uint32_t st_tmp[6];
            std::memcpy(st_tmp, state, sizeof st_tmp);                                          // :351: fill_dense's next state is discarded
            RLB_CHECK(fill_dense_unpacked<T>(ctx, m, cols, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, m, cols, 0, 0, Gc, st_tmp));
            if (new_rank > 0) {
                RLB_CUDA_OK(ctx, cudaMemsetAsync(tmp, 0, sizeof(T) * new_rank * cols, ctx->stream));
                RLB_CHECK(gemm_tn<T>(ctx, m, new_rank, cols, 1.0, A, lda, Gc, m, 0.0, tmp, new_rank, 0));             // Q^T G  (:357)
                RLB_CHECK(gemm_nn<T>(ctx, m, cols, new_rank, -1.0, A, lda, tmp, new_rank, 1.0, Gc, m));               // G - Q Q^T G  (:359)
            }
            RLB_CHECK(hqrq<T>(ctx, m, cols, Gc));                                                                     // geqrf + orgqr  (:364-365)
            RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(A + new_rank * lda, lda * sizeof(T), Gc, m * sizeof(T), m * sizeof(T), cols, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    } else
    // R <- R[0:new_rank, 0:n] * triu(A_hat[0:n, 0:n])  (trmm :349)
    if (new_rank > 0) {
        T* U = as.take<T>((size_t)n * n); RLB_ALLOC(ctx, U);
        T* Rin = as.take<T>((size_t)new_rank * n); RLB_ALLOC(ctx, Rin);
        RLB_CHECK(tri_op<T>(ctx, 2, n, n, A_hat, d, U, n));
        RLB_CUDA_OK(ctx, cudaMemcpy2DAsync(Rin, new_rank * sizeof(T), R, ldr * sizeof(T), new_rank * sizeof(T), n, cudaMemcpyDeviceToDevice, ctx->stream));
        RLB_CHECK(gemm_nn<T>(ctx, new_rank, n, n, 1.0, Rin, new_rank, U, n, 0.0, R, ldr));
    }
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // scratch is released on return
    t_chol = pt.lap();
    if (pt.on) {
        const long long tot = pt.total();
        ctx->phase_us = {t_saso, t_qrcp, t_rank, t_chol, t_piv, t_trsm, tot - (t_saso + t_qrcp + t_rank + t_chol + t_piv + t_trsm), tot};
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CQRRT  (RandLAPACK/drivers/rl_cqrrt.hh:91-297): unpivoted sketched Cholesky QR - CQRRPT without the column pivoting
// ------------------------------------------------------------------------------------------------
template <typename T>
int cqrrt_call(Ctx* ctx, int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, T d_factor, int64_t nnz, int orthogonalization, int compute_Q,
               uint32_t state[6]) {
    // :103-109
    RLB_REQUIRE(ctx, m >= 0);
    RLB_REQUIRE(ctx, n >= 0);
    RLB_REQUIRE(ctx, lda >= m);
    RLB_REQUIRE(ctx, ldr >= n);
    RLB_REQUIRE(ctx, d_factor >= (T)1.0);
    RLB_REQUIRE(ctx, !(A == nullptr && m > 0 && n > 0));
    RLB_REQUIRE(ctx, !(R == nullptr && n > 0));
    const bool sharded = ctx->m_global >= 0;
    const int64_t mg = sharded ? ctx->m_global : m;
    if (n == 0) return 0;
    RLB_REQUIRE(ctx, mg > 0);
    const int64_t d = (int64_t)(d_factor * (T)n);                                             // :135
    RLB_REQUIRE(ctx, d <= mg);                                                                // SparseDist(d, m): the operator must be wide
    ArenaScope as(ctx);
    T* A_hat = as.take<T>((size_t)d * n); RLB_ALLOC(ctx, A_hat);
    T* tau = as.take<T>((size_t)n); RLB_ALLOC(ctx, tau);
    // phase times in the reference's order (:279-282): saso, qr, trtri (= 0), precond, gram, trmm_gram (= 0), potrf, finalize, rest, total
    // (the Q-factor solve is excluded from the total, as in the reference)
    PhaseTimer pt(ctx);
    long long t_saso = 0, t_qr = 0, t_pre = 0, t_gram = 0, t_potrf = 0, t_q = 0, t_fin = 0;
    // SASO (:144-152); state <- S.next_state
    RLB_CHECK(sketch_sparse_left<T>(ctx, d, mg, nnz, d, n, m, (T)1, 0, 0, A, lda, (T)0, A_hat, d, state));
    t_saso = pt.lap();
    // geqrf of the sketch (:160)
    {
        ArenaScope as2(ctx);
        void* ws = arena_push(ctx, qrcp_ws_bytes(n)); RLB_ALLOC(ctx, ws);
        RLB_CHECK(qr_small<T>(ctx, false, d, n, A_hat, d, nullptr, tau, ws));
    }
    t_qr = pt.lap();
    RLB_CHECK(tri_op<T>(ctx, 0, n, n, A_hat, d, R, ldr));                                     // lacpy(Upper) :167
    std::vector<T> dg;
    RLB_CHECK(read_diag<T>(ctx, A_hat, d, n, dg));
    for (int64_t i = 0; i < n; ++i) if (dg[i] == (T)0) return 1;                              // diag_is_nonzero :173-177
    RLB_CHECK(tall_right_solve<T>(ctx, m, n, R, ldr, A, lda));                                // :178 A <- A R_sk^-1
    t_pre = pt.lap();
    // Gram matrix and its Cholesky factor (:186, :194): built in scratch so that only the upper triangle of R is written
    T* G = as.take<T>((size_t)n * n); RLB_ALLOC(ctx, G);
    RLB_CUDA_OK(ctx, cudaMemsetAsync(G, 0, sizeof(T) * n * n, ctx->stream));
    RLB_CHECK(tall_gram_upper<T>(ctx, m, n, A, lda, G, n));
    if (sharded) RLB_CHECK(allreduce_sum<T>(ctx, G, n * n));
    t_gram = pt.lap();
    int info = 0;
    RLB_CHECK(potrf_blocked<T>(ctx, n, G, n, &info));
    RLB_CHECK(tri_op<T>(ctx, 0, n, n, G, n, R, ldr));
    t_potrf = pt.lap();
    if (info != 0) { RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream)); return 1; }        // :194-198
    if (compute_Q) RLB_CHECK(tall_right_solve<T>(ctx, m, n, R, ldr, A, lda));                 // :238 Q = A R_chol^-1
    t_q = pt.lap();
    if (!orthogonalization) {
        // R <- R_chol * triu(A_hat[0:n, 0:n])  (trmm :249)
        T* U = as.take<T>((size_t)n * n); RLB_ALLOC(ctx, U);
        T* Rin = as.take<T>((size_t)n * n); RLB_ALLOC(ctx, Rin);
        RLB_CHECK(tri_op<T>(ctx, 2, n, n, A_hat, d, U, n));
        RLB_CHECK(tri_op<T>(ctx, 2, n, n, R, ldr, Rin, n));
        RLB_CHECK(gemm_nn<T>(ctx, n, n, n, 1.0, Rin, n, U, n, 0.0, G, n));
        RLB_CHECK(tri_op<T>(ctx, 0, n, n, G, n, R, ldr));
    }
    RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));   // scratch is released on return
    t_fin = pt.lap();
    if (pt.on) {
        const long long tot = pt.total() - t_q;
        ctx->phase_us = {t_saso, t_qr, 0, t_pre, t_gram, 0, t_potrf, t_fin, tot - (t_saso + t_qr + t_pre + t_gram + t_potrf + t_fin), tot};
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// SYPS / SYRF / REVD2  (rl_syps.hh:21-143, rl_syrf.hh:21-118, rl_revd2.hh:20-246)
// ------------------------------------------------------------------------------------------------
// ExplicitSymLinOp reads one triangle of A only (the reference's tests poison the other one with NaN, test_revd2.cc:120-137): the
// triangle is mirrored once into a full m x m scratch matrix and every  A * X  below is an ordinary tall product over it.
template <typename T>
__global__ void __launch_bounds__(256) sym_fill_kernel(int upper, int64_t m, const T* __restrict__ A, int64_t lda, T* __restrict__ F) {
    __shared__ T tile[32][33];
    const int64_t bi = blockIdx.x, bj = blockIdx.y;
    if (bi > bj) return;                       // block pairs (bi <= bj): writes F(bi, bj) and its mirror F(bj, bi)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // source block holding the valid triangle: (bi, bj) for upper, (bj, bi) for lower
    const int64_t sr = (upper ? bi : bj) * 32, sc = (upper ? bj : bi) * 32;
    for (int c = ty; c < 32; c += 8) {
        const int64_t i = sr + tx, j = sc + c;
        T v = (T)0;
        if (i < m && j < m) {
            const bool valid = upper ? (i <= j) : (i >= j);
            v = valid ? A[i + j * lda] : A[j + i * lda];        // diagonal blocks: take the mirrored entry from the valid side
        }
        tile[c][tx] = v;                       // tile[c][r] = S(sr + r, sc + c)
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        const int64_t i = sr + tx, j = sc + c;
        if (i < m && j < m) F[i + j * m] = tile[c][tx];
        const int64_t i2 = sc + tx, j2 = sr + c;               // mirrored block: F(sc + r, sr + c) = S(sr + c, sc + r)
        if (bi != bj && i2 < m && j2 < m) F[i2 + j2 * m] = tile[tx][c];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) col_scale_kernel(int64_t m, int64_t k, const T* __restrict__ V, const T* __restrict__ e, T* __restrict__ out) {
    const int64_t total = m * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) out[i] = V[i] * e[i / m];
}
// y <- a * x + b * y  (b == 0: y is not read)
template <typename T>
__global__ void __launch_bounds__(256) vec_axpby_kernel(int64_t n, T a, const T* __restrict__ x, T b, T* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = (b == (T)0) ? a * x[i] : a * x[i] + b * y[i];
}
template <typename T>
static int vec_axpby(Ctx* ctx, int64_t n, T a, const T* x, T b, T* y) {
    if (n == 0) return 0;
    LaunchScope ls(ctx, RLB200_TIMER_SMALL);
    vec_axpby_kernel<T><<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(n, a, x, b, y);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

template <typename T>
static int sym_full(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, T* F) {
    const int64_t nb = (m + 31) / 32;
    RLB_REQUIRE(ctx, nb < 65536);
    LaunchScope ls(ctx, RLB200_TIMER_SMALL);
    sym_fill_kernel<T><<<dim3((unsigned)nb, (unsigned)nb), 256, 0, ctx->stream>>>(uplo == RLB200_UPLO_UPPER ? 1 : 0, m, A, lda, F);
    RLB_CUDA_OK(ctx, cudaGetLastError());
    return 0;
}

// SYPS::call on the mirrored matrix F (rl_syps.hh:58-137).  skop (m x k): in/out as the reference's skop_buff, work (m x k): scratch.
template <typename T>
static int syps_full(Ctx* ctx, int64_t m, const T* F, int64_t k, int64_t p, int64_t q, T* skop, T* work, uint32_t state[6]) {
    RLB_REQUIRE(ctx, p >= 0 && (p == 0 || q >= 1));
    // skop <- fill_dense(DenseDist(m, k)) (:74-75): the natural-layout buffer, read as m x k column-major whatever its shape
    RLB_CHECK(fill_dense_unpacked<T>(ctx, m, k, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, m, k, 0, 0, skop, state));
    RLB_CUDA_OK(ctx, cudaMemsetAsync(work, 0, sizeof(T) * m * k, ctx->stream));                       // :80
    T* out = work; T* in = skop;
    for (int64_t p_done = 0; p_done < p;) {
        RLB_CHECK(tall_nn<T>(ctx, m, k, m, 1.0, F, m, in, m, 0.0, out, m));                           // :86
        ++p_done;
        if (p_done % q == 0) RLB_CHECK(hqrq<T>(ctx, m, k, out));                                      // geqrf + ungqr (:88-93)
        out = (p_done % 2 == 1) ? skop : work;                                                       // :95-96
        in = (p_done % 2 == 1) ? work : skop;
    }
    if (p % 2 == 1) RLB_CUDA_OK(ctx, cudaMemcpyAsync(skop, work, sizeof(T) * m * k, cudaMemcpyDeviceToDevice, ctx->stream));   // :99-100
    return 0;
}

// SYRF::call (rl_syrf.hh:67-112): Q = orth(A * syps(A)).  Returns 0, or 2 when the orthogonaliser fails (the reference throws).
template <typename T>
static int syrf_full(Ctx* ctx, int64_t m, const T* F, int64_t k, T* Q, T* work, uint32_t state[6], const rlb200_revd2_opts& o) {
    RLB_CHECK(syps_full<T>(ctx, m, F, k, o.syps_passes, o.syps_passes_per_stab, work, Q, state));   // the sketch lands in `work`, Q is its scratch (:83)
    RLB_CHECK(tall_nn<T>(ctx, m, k, m, 1.0, F, m, work, m, 0.0, Q, m));                               // :86
    int rc = stab_call<T>(ctx, o.orth, m, k, Q, false, false, nullptr);                               // :94
    if (rc < 0) return rc;
    if (rc) { ctx->err = "SYRF: orthogonalization failed"; return 2; }
    return 0;
}

template <typename T>
int syps_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t k, int64_t passes, int64_t passes_per_stab, T* skop, T* work,
              uint32_t state[6]) {
    RLB_REQUIRE(ctx, m > 0 && k > 0 && lda >= m && A && skop && work && ctx->m_global < 0);
    RLB_REQUIRE(ctx, uplo == RLB200_UPLO_UPPER || uplo == RLB200_UPLO_LOWER);
    ArenaScope as(ctx);
    T* F = as.take<T>((size_t)m * m); RLB_ALLOC(ctx, F);
    RLB_CHECK(sym_full<T>(ctx, uplo, m, A, lda, F));
    OzConstScope a_const(ctx, F);
    return syps_full<T>(ctx, m, F, k, passes, passes_per_stab, skop, work, state);
}

template <typename T>
int syrf_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t k, T* Q, T* work, uint32_t state[6], const rlb200_revd2_opts& o) {
    RLB_REQUIRE(ctx, m > 0 && k > 0 && lda >= m && A && Q && work && ctx->m_global < 0);
    RLB_REQUIRE(ctx, uplo == RLB200_UPLO_UPPER || uplo == RLB200_UPLO_LOWER);
    ArenaScope as(ctx);
    T* F = as.take<T>((size_t)m * m); RLB_ALLOC(ctx, F);
    RLB_CHECK(sym_full<T>(ctx, uplo, m, A, lda, F));
    OzConstScope a_const(ctx, F);
    return syrf_full<T>(ctx, m, F, k, Q, work, state, o);
}

// power_error_est (rl_revd2.hh:20-71): p steps of the power method on  A - V diag(e) V^T  from the vector in vb[0:m]; vb: 4 m scratch
// entries, Mat: m x k scratch.  The last Rayleigh quotient is returned.
template <typename T>
static int revd2_error_est(Ctx* ctx, int64_t m, const T* F, int64_t k, int p, T* vb, const T* V, T* Mat, const T* e_dev, T* err_out) {
    T err = 0;
    T* g = vb; T* t1 = vb + m; T* t2 = vb + 2 * m; T* t3 = vb + 3 * m;
    ArenaScope as(ctx);
    T* dotv = as.take<T>(1); RLB_ALLOC(ctx, dotv);
    for (int i = 0; i < p; ++i) {
        double gn = 0;
        RLB_CHECK(fro_norm<T>(ctx, g, m, 1, m, false, &gn));
        RLB_CHECK(vec_axpby<T>(ctx, m, (T)(1.0 / gn), g, (T)0, g));                                   // g / ||g||  (:36)
        RLB_CHECK(gemm_tn<T>(ctx, m, k, 1, 1.0, V, m, g, m, 0.0, t1, m, 0));                           // V^T g  (:40)
        {
            LaunchScope ls(ctx, RLB200_TIMER_SMALL);
            col_scale_kernel<T><<<(unsigned)std::min<int64_t>((m * k + 255) / 256, (int64_t)ctx->num_sms * 8), 256, 0, ctx->stream>>>(m, k, V, e_dev, Mat);   // :44-48
            RLB_CUDA_OK(ctx, cudaGetLastError());
        }
        RLB_CHECK(gemm_nn<T>(ctx, m, 1, k, 1.0, Mat, m, t1, m, 0.0, t2, m));                           // V diag(e) V^T g  (:52)
        RLB_CHECK(gemm_nn<T>(ctx, m, 1, m, 1.0, F, m, g, m, 0.0, t3, m));                              // A g  (:55)
        RLB_CHECK(vec_axpby<T>(ctx, m, (T)-1, t2, (T)1, t3));                                          // w  (:60)
        RLB_CHECK(gemm_tn<T>(ctx, m, 1, 1, 1.0, g, m, t3, m, 0.0, dotv, 1, 0));                        // g . w  (:62)
        RLB_CHECK(read_scalar<T>(ctx, dotv, &err));
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(g, t3, sizeof(T) * m, cudaMemcpyDeviceToDevice, ctx->stream));   // :64
    }
    *err_out = err;
    return 0;
}

// REVD2::call (rl_revd2.hh:120-246).  V (m x k_cap) and eigvals (k_cap) are the caller's device buffers; *k_io grows as in the reference
// (k <- 2k, capped at m) until the error estimate passes.  Returns 0; 1 = the Cholesky factorization failed and 2 = the orthogonaliser
// failed (the reference throws std::runtime_error for both); 3 = the next k would exceed k_cap (the outputs hold the last, unconverged
// iterate and *k_io its size; the reference would have grown its vectors).  *err_out: the last error estimate.
template <typename T>
int revd2_call(Ctx* ctx, int uplo, int64_t m, const T* A, int64_t lda, int64_t* k_io, int64_t k_cap, T tol, T* V, T* eigvals, uint32_t state[6],
               const rlb200_revd2_opts& o, T* err_out) {
    RLB_REQUIRE(ctx, m >= 0 && k_io && *k_io > 0 && tol >= (T)0 && !(A == nullptr && m > 0));          // :131-134
    RLB_REQUIRE(ctx, uplo == RLB200_UPLO_UPPER || uplo == RLB200_UPLO_LOWER);
    RLB_REQUIRE(ctx, lda >= m && V && eigvals && k_cap >= *k_io && ctx->m_global < 0 && o.error_est_p >= 0);
    RLB_REQUIRE(ctx, m > 0 && *k_io <= m);
    int64_t k = *k_io;
    uint32_t est_state[6];
    std::memcpy(est_state, state, sizeof est_state);                                                   // :152-153: same counter, key + 1
    if (++est_state[4] == 0) ++est_state[5];
    ArenaScope as(ctx);
    T* F = as.take<T>((size_t)m * m); RLB_ALLOC(ctx, F);
    RLB_CHECK(sym_full<T>(ctx, uplo, m, A, lda, F));
    OzConstScope a_const(ctx, F);
    const T eps = std::numeric_limits<T>::epsilon();
    T err = 0;
    int code = 0;
    while (true) {
        ArenaScope it(ctx);
        const int64_t ko = std::max<int64_t>(k, 4);
        T* Omega = it.take<T>((size_t)m * ko); RLB_ALLOC(ctx, Omega);
        T* Y = it.take<T>((size_t)m * k); RLB_ALLOC(ctx, Y);
        T* work = it.take<T>((size_t)m * k); RLB_ALLOC(ctx, work);
        T* R = it.take<T>((size_t)k * k); RLB_ALLOC(ctx, R);
        T* S = it.take<T>((size_t)k); RLB_ALLOC(ctx, S);
        T* W = it.take<T>((size_t)k * k); RLB_ALLOC(ctx, W);
        int rc = syrf_full<T>(ctx, m, F, k, Omega, work, state, o);                                    // :166
        if (rc) { code = rc; break; }
        RLB_CHECK(gemm_nn<T>(ctx, m, k, m, 1.0, F, m, Omega, m, 0.0, Y, m));                           // Y = A Omega on the fp64 pipe (:169): nu below is eps-sized
        double ynorm = 0;
        RLB_CHECK(fro_norm<T>(ctx, Y, m, k, m, false, &ynorm));
        const T nu = eps * (T)ynorm;                                                                   // :171
        // R = chol(Omega^T Y + nu Omega^T Omega)  (:177-186)
        RLB_CUDA_OK(ctx, cudaMemsetAsync(R, 0, sizeof(T) * k * k, ctx->stream));
        RLB_CHECK(gemm_tn<T>(ctx, m, k, k, (double)nu, Omega, m, Omega, m, 0.0, R, k, 0));
        RLB_CHECK(gemm_tn<T>(ctx, m, k, k, 1.0, Omega, m, Y, m, 1.0, R, k, 0));
        int info = 0;
        RLB_CHECK(potrf_blocked<T>(ctx, k, R, k, &info));
        if (info != 0) { ctx->err = "REVD2: Cholesky decomposition failed"; code = 1; break; }
        RLB_CHECK(trsm_right_upper<T>(ctx, m, k, R, k, Y, m));                                         // B = Y R^-1  (:190)
        {   // [V, S, ~] = svd(B)  (:195)
            void* ws = arena_push(ctx, svd_ws_bytes(m, k, sizeof(T))); RLB_ALLOC(ctx, ws);
            RLB_CHECK(svd_tall<T>(ctx, m, k, Y, m, S, W, ws, nullptr));
        }
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(V, Y, sizeof(T) * m * k, cudaMemcpyDeviceToDevice, ctx->stream));
        std::vector<T> s(k), ev(k);
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(s.data(), S, sizeof(T) * k, cudaMemcpyDeviceToHost, ctx->stream));
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
        int64_t r = 0;
        for (int64_t i = 0; i < k; ++i) { ev[i] = s[i] * s[i]; if (ev[i] > nu) ++r; }                  // :198-207
        for (int64_t i = 0; i < r; ++i) if (!(ev[i] - nu < 0)) ev[i] -= nu;                            // :211-212
        RLB_CUDA_OK(ctx, cudaMemcpyAsync(eigvals, ev.data(), sizeof(T) * k, cudaMemcpyHostToDevice, ctx->stream));
        if (r < k) RLB_CUDA_OK(ctx, cudaMemsetAsync(V + m * r, 0, sizeof(T) * m * (k - r), ctx->stream));   // :214
        // error estimate from a fresh Gaussian vector (:219-223)
        RLB_CHECK(fill_dense_unpacked<T>(ctx, m, 1, RLB200_FAMILY_GAUSSIAN, RLB200_AXIS_LONG, RLB200_LAYOUT_NATURAL, m, 1, 0, 0, Omega, est_state));
        RLB_CHECK(revd2_error_est<T>(ctx, m, F, k, o.error_est_p, Omega, V, Y, eigvals, &err));
        RLB_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));                                          // ev is host memory of this iteration
        if (err <= 5 * std::max(tol, nu) || k == m) break;                                             // :225-231
        const int64_t k_next = (2 * k > m) ? m : 2 * k;
        if (k_next > k_cap) { ctx->err = "REVD2: the next rank estimate exceeds the capacity of V / eigvals"; code = 3; break; }
        k = k_next;
    }
    *k_io = k;
    if (err_out) *err_out = err;
    return code;
}

#define INST(T)                                                                                                             \
    template int cqrrt_call<T>(Ctx*, int64_t, int64_t, T*, int64_t, T*, int64_t, T, int64_t, int, int, uint32_t*);          \
    template int stab_call<T>(Ctx*, int, int64_t, int64_t, T*, bool, bool, int*);                                           \
    template int rs_call<T>(Ctx*, int64_t, int64_t, const T*, int64_t, T*, T*, uint32_t*, const rlb200_stack_opts&);         \
    template int rf_call<T>(Ctx*, int64_t, int64_t, const T*, int64_t, T*, uint32_t*, const rlb200_stack_opts&);             \
    template int qb_call<T>(Ctx*, int64_t, int64_t, T*, int64_t*, int64_t, T, T*, T*, T*, uint32_t*, const rlb200_stack_opts&); \
    template int rsvd_call<T>(Ctx*, int64_t, int64_t, T*, int64_t*, T, T*, T*, T*, T*, uint32_t*, const rlb200_stack_opts&, int*); \
    template int cqrrpt_call<T>(Ctx*, int64_t, int64_t, T*, int64_t, T*, int64_t, int64_t*, T, T, int64_t, int64_t*, uint32_t*); \
    template int syps_call<T>(Ctx*, int, int64_t, const T*, int64_t, int64_t, int64_t, int64_t, T*, T*, uint32_t*);          \
    template int syrf_call<T>(Ctx*, int, int64_t, const T*, int64_t, int64_t, T*, T*, uint32_t*, const rlb200_revd2_opts&);  \
    template int revd2_call<T>(Ctx*, int, int64_t, const T*, int64_t, int64_t*, int64_t, T, T*, T*, uint32_t*, const rlb200_revd2_opts&, T*);
INST(double)
INST(float)

}  // namespace rlb
