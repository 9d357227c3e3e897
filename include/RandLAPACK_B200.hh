// RandLAPACK_B200.hh — algorithm objects for RandLAPACK's sketch-and-factor path, backed by librlb200.so (sm_100a).
//
// Same shape as the reference's objects (constructor arguments, public fields, `call` signatures with HOST pointers,
// malloc-family ownership of QB/RSVD outputs, int return codes), so existing RandLAPACK compositions keep compiling:
//
//   reference (paths relative to the reference root)                       this header
//   RandLAPACK::Stabilization<T>      comps/rl_orth.hh:13-23               rlb200::Stabilization<T>
//   RandLAPACK::CholQRQ/PLUL/HQRQ<T>  comps/rl_orth.hh:25-230              rlb200::CholQRQ / PLUL / HQRQ<T>
//   RandLAPACK::RowSketcher / RS      comps/rl_rs.hh:15-178                rlb200::RS<T>
//   RandLAPACK::RangeFinder / RF      comps/rl_rf.hh:16-137                rlb200::RF<T>
//   RandLAPACK::QBalg / QB            comps/rl_qb.hh:17-268                rlb200::QB<T>
//   RandLAPACK::RSVDalg / RSVD        drivers/rl_rsvd.hh:15-154            rlb200::RSVD<T>
//   RandLAPACK::CQRRPTalg / CQRRPT    drivers/rl_cqrrpt.hh:20-391          rlb200::CQRRPT<T>
//   RandLAPACK::CQRRTalg / CQRRT      drivers/rl_cqrrt.hh:20-297           rlb200::CQRRT<T>
//   RandLAPACK::CQRRPT_GPU_alg / CQRRPT_GPU  drivers/rl_cqrrpt_gpu.hh:15-387  rlb200::CQRRPT_GPU<T>  (host pointers, as in the reference)
//   RandLAPACK::BQRRPalg / BQRRP      drivers/rl_bqrrp.hh:19-665           rlb200::BQRRP<T>
//   RandLAPACK::hqrrp (free function) drivers/rl_hqrrp.hh:811-1196         rlb200::hqrrp<T>
//   RandLAPACK::BQRRP_GPU_alg / BQRRP_GPU  drivers/rl_bqrrp_gpu.hh:27-942   rlb200::BQRRP_GPU<T>   (device pointers, sketch as input)
//   RandLAPACK::linops::DenseLinOp / ExplicitSymLinOp  linops/rl_dense_linop.hh:36, linops/rl_sym_linops.hh   rlb200::DenseLinOp / ExplicitSymLinOp<T>  (matrix resident on the device)
//   RandLAPACK::SYPS / SYRF / REVD2   comps/rl_syps.hh, comps/rl_syrf.hh, drivers/rl_revd2.hh   rlb200::SYPS / SYRF / REVD2<T>  (explicit symmetric A)
//
// Two modes:
//  * default: self-contained (no reference headers needed); `rlb200::RNGState` stands in for RandBLAS::RNGState.
//  * #define RLB200_WITH_RANDLAPACK (after including <RandLAPACK.hh>): every class derives from the reference's own
//    abstract base and takes RandBLAS::RNGState<r123::Philox4x32>&, so e.g. a reference RandLAPACK::RSVD can be
//    constructed on top of an rlb200::QB, or a reference QB on top of an rlb200::RF.
//
// The B200 objects compose with each other on the device in one flattened call (no host round trips between RS, RF,
// QB and RSVD).  A B200 RS/RF/QB needs B200 stabilisers (it asks them for their kind); mixing in a CPU stabiliser is
// rejected with std::invalid_argument rather than silently falling back to the CPU.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rlb200.h"

namespace rlb200 {

// thrown for negative RLB200_ERR_* codes (argument errors are RandLAPACK::Error / RandBLAS::Error in the reference,
// CUDA failures abort there: RandLAPACK/rl_exceptions.hh:37-52, RandLAPACK/gpu_functions/rl_cuda_macros.hh:34-42)
class Error : public std::runtime_error {
public:
    Error(int code, const std::string& what) : std::runtime_error("rlb200 error " + std::to_string(code) + ": " + what), code(code) {}
    int code;
};

// RAII owner of an rlb200_ctx (device + stream + workspaces)
class Context {
public:
    explicit Context(int device = 0, void* stream = nullptr) {
        int rc = rlb200_create(&h_, device, stream);
        if (rc) throw Error(rc, "rlb200_create failed (an sm_100 device is required; there is no CPU fallback)");
    }
    ~Context() { rlb200_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    rlb200_ctx* get() const { return h_; }
    // the reference's `times` vectors (microseconds; rl_cqrrpt.hh:371-384, rl_cqrrt.hh:279-282, rl_bqrrp.hh:582-584)
    void phase_timing(bool on) { check(rlb200_set_phase_timing(h_, on ? 1 : 0)); }
    std::vector<long> phase_times() {
        long long buf[32];
        int n = rlb200_get_phase_times(h_, buf, 32);
        std::vector<long> v;
        for (int i = 0; i < n && i < 32; ++i) v.push_back((long)buf[i]);
        return v;
    }
    // engine of the tall products: RLB200_FP64_I8SLICES (tcgen05 int8 digit slices, default) or RLB200_FP64_DMMA (fp64 pipe);
    // digits: 0 = default (6 for fp64 storage, 7 inside the QR drivers, 4 for fp32), else 3..7
    void set_engine(int engine, int digits = 0) { check(rlb200_set_fp64_engine(h_, engine)); check(rlb200_set_i8_digits(h_, digits)); }
    int check(int rc) const {
        if (rc < 0) throw Error(rc, rlb200_last_error(h_));
        return rc;
    }
private:
    rlb200_ctx* h_ = nullptr;
};

// process-wide context on device 0 for objects built with the reference's own constructor arguments
inline Context& default_context() {
    static Context ctx(0, nullptr);
    return ctx;
}

#ifndef RLB200_WITH_RANDLAPACK
// stand-in for RandBLAS::RNGState<r123::Philox4x32> (RandBLAS/RandBLAS/base.hh:64-164)
struct RNGState {
    uint32_t counter[4] = {0, 0, 0, 0};
    uint32_t key[2] = {0, 0};
    RNGState() = default;
    explicit RNGState(uint64_t k) { key[0] = (uint32_t)k; key[1] = (uint32_t)(k >> 32); }
};
inline void state_to_words(const RNGState& s, uint32_t w[6]) { std::memcpy(w, s.counter, 16); std::memcpy(w + 4, s.key, 8); }
inline void words_to_state(const uint32_t w[6], RNGState& s) { std::memcpy(s.counter, w, 16); std::memcpy(s.key, w + 4, 8); }
using state_t = RNGState;
#define RLB200_OVERRIDE
#else
using state_t = RandBLAS::RNGState<r123::Philox4x32>;
inline void state_to_words(const state_t& s, uint32_t w[6]) { for (int i = 0; i < 4; ++i) w[i] = s.counter.v[i]; w[4] = s.key.v[0]; w[5] = s.key.v[1]; }
inline void words_to_state(const uint32_t w[6], state_t& s) { for (int i = 0; i < 4; ++i) s.counter.v[i] = w[i]; s.key.v[0] = w[4]; s.key.v[1] = w[5]; }
#define RLB200_OVERRIDE override
#endif

namespace detail {
template <typename T> struct abi;
template <> struct abi<double> {
    static constexpr auto stab = rlb200_stab_f64_dev; static constexpr auto rs = rlb200_rs_f64_dev; static constexpr auto rf = rlb200_rf_f64_dev;
    static constexpr auto qb = rlb200_qb_f64_dev; static constexpr auto rsvd_host = rlb200_rsvd_f64_host;
    static constexpr auto cqrrpt_host = rlb200_cqrrpt_f64_host; static constexpr auto bqrrp_host = rlb200_bqrrp_f64_host;
    static constexpr auto hqrrp_host = rlb200_hqrrp_f64_host;
    static constexpr auto bqrrp_dev_sk = rlb200_bqrrp_f64_dev_sk; static constexpr auto cqrrt_host = rlb200_cqrrt_f64_host;
    static constexpr auto syps = rlb200_syps_f64_dev; static constexpr auto syrf = rlb200_syrf_f64_dev; static constexpr auto revd2_host = rlb200_revd2_f64_host;
    static constexpr auto gemm = rlb200_gemm_f64_dev;
    static constexpr auto sketch_sparse_left = rlb200_sketch_sparse_left_f64_dev; static constexpr auto sketch_dense_left = rlb200_sketch_dense_left_f64_dev;
};
template <> struct abi<float> {
    static constexpr auto stab = rlb200_stab_f32_dev; static constexpr auto rs = rlb200_rs_f32_dev; static constexpr auto rf = rlb200_rf_f32_dev;
    static constexpr auto qb = rlb200_qb_f32_dev; static constexpr auto rsvd_host = rlb200_rsvd_f32_host;
    static constexpr auto cqrrpt_host = rlb200_cqrrpt_f32_host; static constexpr auto bqrrp_host = rlb200_bqrrp_f32_host;
    static constexpr auto hqrrp_host = rlb200_hqrrp_f32_host;
    static constexpr auto bqrrp_dev_sk = rlb200_bqrrp_f32_dev_sk; static constexpr auto cqrrt_host = rlb200_cqrrt_f32_host;
    static constexpr auto syps = rlb200_syps_f32_dev; static constexpr auto syrf = rlb200_syrf_f32_dev; static constexpr auto revd2_host = rlb200_revd2_f32_host;
    static constexpr auto gemm = rlb200_gemm_f32_dev;
    static constexpr auto sketch_sparse_left = rlb200_sketch_sparse_left_f32_dev; static constexpr auto sketch_dense_left = rlb200_sketch_dense_left_f32_dev;
};

// device buffer staged from / to a host pointer
template <typename T>
class DevBuf {
public:
    DevBuf(Context& c, int64_t count, const T* init = nullptr) : c_(c), n_(count) {
        c_.check(rlb200_dev_alloc(c_.get(), sizeof(T) * (size_t)count, &p_));
        if (init) c_.check(rlb200_copy_h2d(c_.get(), p_, init, sizeof(T) * (size_t)count));
    }
    ~DevBuf() { rlb200_dev_free(c_.get(), p_); }
    T* ptr() { return static_cast<T*>(p_); }
    void to_host(T* dst, int64_t count) { c_.check(rlb200_copy_d2h(c_.get(), dst, p_, sizeof(T) * (size_t)count)); }
private:
    Context& c_; void* p_ = nullptr; int64_t n_;
};
}  // namespace detail

// ---------------------------------------------------------------------------------------------------------------------
// Stabilization (rl_orth.hh)
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RLB200_WITH_RANDLAPACK
template <typename T>
class Stabilization {
public:
    virtual ~Stabilization() {}
    virtual int call(int64_t m, int64_t k, T* A) = 0;
};
template <typename T> using StabBase = Stabilization<T>;
#else
template <typename T> using StabBase = RandLAPACK::Stabilization<T>;
#endif

// common part of the three B200 stabilisers: in place on a HOST m x k column-major matrix, like the reference
template <typename T>
class DeviceStab : public StabBase<T> {
public:
    DeviceStab(Context& ctx, int kind, bool c_check, bool verb) : cond_check(c_check), verbose(verb), chol_fail(false), ctx_(ctx), kind_(kind) {}
    int call(int64_t m, int64_t k, T* A) override {
        detail::DevBuf<T> d(ctx_, m * k, A);
        int cf = 0;
        int rc = ctx_.check(detail::abi<T>::stab(ctx_.get(), kind_, m, k, d.ptr(), cond_check, &cf));
        chol_fail = cf != 0;
        d.to_host(A, m * k);
        return rc;
    }
    int kind() const { return kind_; }
    Context& context() const { return ctx_; }
    bool cond_check, verbose, chol_fail;
private:
    Context& ctx_;
    int kind_;
};
// constructors: the reference's (c_check, verb) — using default_context() — or with an explicit Context first
template <typename T> struct CholQRQ : DeviceStab<T> {
    CholQRQ(bool c_check, bool verb) : DeviceStab<T>(default_context(), RLB200_STAB_CHOLQRQ, c_check, verb) {}
    CholQRQ(Context& c, bool c_check, bool verb) : DeviceStab<T>(c, RLB200_STAB_CHOLQRQ, c_check, verb) {}
};
template <typename T> struct PLUL : DeviceStab<T> {
    PLUL(bool c_check, bool verb) : DeviceStab<T>(default_context(), RLB200_STAB_PLUL, c_check, verb) {}
    PLUL(Context& c, bool c_check, bool verb) : DeviceStab<T>(c, RLB200_STAB_PLUL, c_check, verb) {}
};
template <typename T> struct HQRQ : DeviceStab<T> {
    HQRQ(bool c_check, bool verb) : DeviceStab<T>(default_context(), RLB200_STAB_HQRQ, c_check, verb) {}
    HQRQ(Context& c, bool c_check, bool verb) : DeviceStab<T>(c, RLB200_STAB_HQRQ, c_check, verb) {}
};

template <typename T>
inline DeviceStab<T>& require_device_stab(StabBase<T>& s, const char* who) {
    auto* d = dynamic_cast<DeviceStab<T>*>(&s);
    if (!d) throw std::invalid_argument(std::string(who) + ": composes only with rlb200 stabilisers (CholQRQ/PLUL/HQRQ); no CPU fallback");
    return *d;
}

// ---------------------------------------------------------------------------------------------------------------------
// RS (rl_rs.hh:31-178)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
class RS
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::RowSketcher<T, r123::Philox4x32>
#endif
{
public:
    RS(StabBase<T>& stab_obj, int64_t p, int64_t q, bool verb, bool cond)
        : Stab_Obj(stab_obj), passes_over_data(p), passes_per_stab(q), verbose(verb), cond_check(cond) {}
    virtual ~RS() {}
    void fill(rlb200_stack_opts& o) const {
        auto& s = require_device_stab<T>(Stab_Obj, "rlb200::RS");
        o.passes_over_data = passes_over_data; o.passes_per_stab = passes_per_stab; o.stab = s.kind();
        o.cond_check = s.cond_check;
    }
    Context& context() const { return require_device_stab<T>(Stab_Obj, "rlb200::RS").context(); }
    // A: m x n (host, column-major); Omega: n x k (host, caller-allocated, as in RF::call rl_rf.hh:116)
    int call(int64_t m, int64_t n, const T*& A, int64_t k, T*& Omega, state_t& state) RLB200_OVERRIDE {
        Context& c = context();
        rlb200_stack_opts o{}; fill(o);
        detail::DevBuf<T> dA(c, m * n, A), dOm(c, n * k), dW(c, passes_over_data > 0 ? m * k : 1);
        uint32_t w[6]; state_to_words(state, w);
        int rc = c.check(detail::abi<T>::rs(c.get(), m, n, dA.ptr(), k, dOm.ptr(), dW.ptr(), w, &o));
        words_to_state(w, state);
        dOm.to_host(Omega, n * k);
        return rc;
    }
    StabBase<T>& Stab_Obj;
    int64_t passes_over_data, passes_per_stab;
    bool verbose, cond_check;
    std::vector<T> cond_nums;   // kept for source compatibility; condition-number logging is not offered on the device
};

// ---------------------------------------------------------------------------------------------------------------------
// RF (rl_rf.hh:31-137)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
class RF
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::RangeFinder<T, r123::Philox4x32>
#endif
{
public:
    RF(RS<T>& rs_obj, StabBase<T>& orth_obj, bool verb, bool cond) : rs(rs_obj), orth(orth_obj), verbose(verb), cond_check(cond) {}
    virtual ~RF() {}
    void fill(rlb200_stack_opts& o) const { rs.fill(o); o.orth_rf = require_device_stab<T>(orth, "rlb200::RF").kind(); }
    Context& context() const { return rs.context(); }
    int call(int64_t m, int64_t n, const T* A, int64_t k, T* Q, state_t& state) RLB200_OVERRIDE {
        Context& c = context();
        rlb200_stack_opts o{}; fill(o);
        detail::DevBuf<T> dA(c, m * n, A), dQ(c, m * k);
        uint32_t w[6]; state_to_words(state, w);
        int rc = c.check(detail::abi<T>::rf(c.get(), m, n, dA.ptr(), k, dQ.ptr(), w, &o));
        words_to_state(w, state);
        dQ.to_host(Q, m * k);
        return rc;
    }
    RS<T>& rs;
    StabBase<T>& orth;
    bool verbose, cond_check;
    std::vector<T> cond_nums;
};

// ---------------------------------------------------------------------------------------------------------------------
// QB (rl_qb.hh:36-268)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
class QB
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::QBalg<T, r123::Philox4x32>
#endif
{
public:
    QB(RF<T>& rf_obj, StabBase<T>& orth_obj, bool verb, bool orth) : rf(rf_obj), orth(orth_obj), verbose(verb), orth_check(orth) {}
    virtual ~QB() {}
    void fill(rlb200_stack_opts& o) const {
        rf.fill(o); o.orth_qb = require_device_stab<T>(orth, "rlb200::QB").kind(); o.orth_check = orth_check;
    }
    Context& context() const { return rf.context(); }
    // Q, BT: nullptr or malloc-family on entry; (re)allocated here with calloc and owned by the caller (rl_qb.hh:154-159)
    int call(int64_t m, int64_t n, T* A, int64_t& k, int64_t block_sz, T tol, T*& Q, T*& BT, state_t& state) RLB200_OVERRIDE {
        Context& c = context();
        rlb200_stack_opts o{}; fill(o); o.block_sz = block_sz;
        if (Q) free(Q);
        if (BT) free(BT);
        const int64_t k_in = k;
        Q = (T*)calloc((size_t)(m * k_in), sizeof(T));
        BT = (T*)calloc((size_t)(n * k_in), sizeof(T));
        detail::DevBuf<T> dA(c, m * n, A), dQ(c, m * k_in), dBT(c, n * k_in);
        uint32_t w[6]; state_to_words(state, w);
        int rc = c.check(detail::abi<T>::qb(c.get(), m, n, dA.ptr(), &k, block_sz, tol, dQ.ptr(), dBT.ptr(), nullptr, w, &o));
        words_to_state(w, state);
        if (k > 0) { dQ.to_host(Q, m * k); dBT.to_host(BT, n * k); }
        return rc;
    }
    RF<T>& rf;
    StabBase<T>& orth;
    bool verbose, orth_check;
};

// ---------------------------------------------------------------------------------------------------------------------
// RSVD (rl_rsvd.hh:34-154)
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
class RSVD
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::RSVDalg<T, r123::Philox4x32>
#endif
{
public:
    RSVD(QB<T>& qb_obj, int64_t b_sz) : QB_Obj(qb_obj), block_sz(b_sz) {}
    virtual ~RSVD() {}
    // U (m x k), S (k), V (n x k) are calloc'd here and free()d by the caller (rl_rsvd.hh:141-143, test_rsvd.cc:162-164)
    int call(int64_t m, int64_t n, T* A, int64_t& k, T tol, T*& U, T*& S, T*& V, state_t& state) RLB200_OVERRIDE {
        Context& c = QB_Obj.context();
        rlb200_stack_opts o{}; QB_Obj.fill(o); o.block_sz = block_sz;
        const int64_t k_in = k;
        if (k_in <= 0) throw Error(RLB200_ERR_ARG, "target rank k must be > 0");          // rl_rsvd.hh:130
        U = (T*)calloc((size_t)(m * k_in), sizeof(T));
        S = (T*)calloc((size_t)k_in, sizeof(T));
        V = (T*)calloc((size_t)(n * k_in), sizeof(T));
        uint32_t w[6]; state_to_words(state, w);
        int rc = c.check(detail::abi<T>::rsvd_host(c.get(), m, n, A, &k, tol, U, S, V, w, &o, &qb_code));
        words_to_state(w, state);
        return rc;
    }
    QB<T>& QB_Obj;
    int64_t block_sz;
    int qb_code = 0;   // the code QB returned (the reference discards it, rl_rsvd.hh:137)
};

// ---------------------------------------------------------------------------------------------------------------------
// hqrrp (rl_hqrrp.hh:811-1196): the reference's free function, same argument list (HOST pointers).  `timing`: when non-null, *timing is
// (re)allocated with realloc to 26 entries as the reference does (:1137) and the nine leading entries (:1140-1148) are filled, in
// microseconds; the 2 x 9 per-routine entries of the unblocked QR loops, which have no counterpart here, are zero.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
int64_t hqrrp(Context& c, int64_t m_A, int64_t n_A, T* buff_A, int64_t ldim_A, int64_t* buff_jpvt, T* buff_tau, int64_t nb_alg, int64_t pp,
              int64_t panel_pivoting, int64_t qr_type, state_t& state, T** timing) {
    uint32_t w[6]; state_to_words(state, w);
    if (timing) c.phase_timing(true);
    int rc = c.check(detail::abi<T>::hqrrp_host(c.get(), m_A, n_A, buff_A, ldim_A, buff_jpvt, buff_tau, nb_alg, pp, (int)panel_pivoting,
                                                (int)qr_type, w));
    if (timing) {
        std::vector<long> t = c.phase_times();
        c.phase_timing(false);
        T* out = static_cast<T*>(std::realloc(*timing, 26 * sizeof(T)));
        if (out) { for (int i = 0; i < 26; ++i) out[i] = i < (int)t.size() ? (T)t[i] : (T)0; *timing = out; }
    }
    words_to_state(w, state);
    return rc;
}
template <typename T>
int64_t hqrrp(int64_t m_A, int64_t n_A, T* buff_A, int64_t ldim_A, int64_t* buff_jpvt, T* buff_tau, int64_t nb_alg, int64_t pp,
              int64_t panel_pivoting, int64_t qr_type, state_t& state, T** timing) {
    return hqrrp<T>(default_context(), m_A, n_A, buff_A, ldim_A, buff_jpvt, buff_tau, nb_alg, pp, panel_pivoting, qr_type, state, timing);
}

// ---------------------------------------------------------------------------------------------------------------------
// CQRRPT (rl_cqrrpt.hh:20-391): same constructor (time_subroutines, eps), public fields and call signature (HOST pointers).
// ---------------------------------------------------------------------------------------------------------------------
struct CQRRPTSubroutines {
    enum QRCP { geqp3 = RLB200_CQRRPT_QRCP_GEQP3, bqrrp = RLB200_CQRRPT_QRCP_BQRRP, hqrrp = RLB200_CQRRPT_QRCP_HQRRP };     // rl_cqrrpt.hh:39-43
};
template <typename T>
class CQRRPT
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::CQRRPTalg<T, r123::Philox4x32>
#endif
{
public:
    using Subroutines = CQRRPTSubroutines;
    CQRRPT(bool time_subroutines, T ep) : CQRRPT(default_context(), time_subroutines, ep) {}
    CQRRPT(Context& c, bool time_subroutines, T ep)
        : timing(time_subroutines), eps(ep), rank(0), nnz(2), qrcp(Subroutines::geqp3), orthogonalization(false), nb_alg(64), oversampling(10),
          panel_pivoting(1), use_cholqr(0), ctx_(&c) {}                                      // HQRRP defaults: rl_cqrrpt.hh:60-63
    virtual ~CQRRPT() {}
    // A (m x n, lda) <- Q; R (ldr >= n): rank x n; J: n 1-based pivots (rl_cqrrpt.hh:146-156)
    int call(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J, T d_factor, state_t& state) RLB200_OVERRIDE {
        uint32_t w[6]; state_to_words(state, w);
        int64_t r = 0;
        if (timing) ctx_->phase_timing(true);
        ctx_->check(rlb200_set_cqrrpt_qrcp(ctx_->get(), (int)qrcp));
        ctx_->check(rlb200_set_cqrrpt_orthogonalization(ctx_->get(), orthogonalization ? 1 : 0));
        ctx_->check(rlb200_set_cqrrpt_hqrrp_opts(ctx_->get(), nb_alg, oversampling, (int)panel_pivoting, (int)use_cholqr));
        int rc = ctx_->check(detail::abi<T>::cqrrpt_host(ctx_->get(), m, n, A, lda, R, ldr, J, d_factor, eps, nnz, &r, w));
        if (timing) { times = ctx_->phase_times(); ctx_->phase_timing(false); }
        words_to_state(w, state);
        rank = r;
        return rc;
    }
    bool timing;
    T eps;
    int64_t rank;
    std::vector<long> times;   // 8 entries when `timing` (rl_cqrrpt.hh:371-384)
    int64_t nnz;
    Subroutines::QRCP qrcp;    // QRCP of the sketch: geqp3 (default), bqrrp or hqrrp (rl_cqrrpt.hh:230-247)
    bool orthogonalization;    // rl_cqrrpt.hh:139-142
    int64_t nb_alg, oversampling, panel_pivoting, use_cholqr;     // HQRRP-related (rl_cqrrpt.hh:134-137)
private:
    Context* ctx_;
};

// ---------------------------------------------------------------------------------------------------------------------
// CQRRT (rl_cqrrt.hh:20-297): unpivoted sketched Cholesky QR; same constructor (time_subroutines, eps), public fields and call signature
// (HOST pointers).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
class CQRRT
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::CQRRTalg<T, r123::Philox4x32>
#endif
{
public:
    CQRRT(bool time_subroutines, T ep) : CQRRT(default_context(), time_subroutines, ep) {}
    CQRRT(Context& c, bool time_subroutines, T ep)
        : timing(time_subroutines), eps(ep), nnz(2), orthogonalization(false), compute_Q(true), ctx_(&c) {}
    virtual ~CQRRT() {}
    // A (m x n, lda) <- Q; R (ldr >= n): n x n upper triangular (rl_cqrrt.hh:91-101)
    int call(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, T d_factor, state_t& state) RLB200_OVERRIDE {
        uint32_t w[6]; state_to_words(state, w);
        if (timing) ctx_->phase_timing(true);
        int rc = ctx_->check(detail::abi<T>::cqrrt_host(ctx_->get(), m, n, A, lda, R, ldr, d_factor, nnz, orthogonalization ? 1 : 0,
                                                        compute_Q ? 1 : 0, w));
        if (timing) { times = ctx_->phase_times(); ctx_->phase_timing(false); }
        words_to_state(w, state);
        return rc;
    }
    bool timing;
    T eps;
    std::vector<long> times;
    int64_t nnz;
    bool orthogonalization;
    bool compute_Q;
private:
    Context* ctx_;
};

// ---------------------------------------------------------------------------------------------------------------------
// BQRRP (rl_bqrrp.hh:19-665): same constructor (time_subroutines, b_sz), public fields and call signature (HOST pointers).
// ---------------------------------------------------------------------------------------------------------------------
struct BQRRPSubroutines {
    enum QRCPWide { luqr = RLB200_QRCP_LUQR, geqp3 = RLB200_QRCP_GEQP3 };
    enum QRTall { geqrf = RLB200_QRTALL_GEQRF, cholqr = RLB200_QRTALL_CHOLQR, geqrt = RLB200_QRTALL_GEQRT };
};
template <typename T>
class BQRRP
#ifdef RLB200_WITH_RANDLAPACK
    : public RandLAPACK::BQRRPalg<T, r123::Philox4x32>
#endif
{
public:
    using Subroutines = BQRRPSubroutines;
    BQRRP(bool time_subroutines, int64_t b_sz) : BQRRP(default_context(), time_subroutines, b_sz) {}
    BQRRP(Context& c, bool time_subroutines, int64_t b_sz)
        : timing(time_subroutines), rank(0), block_size(b_sz), internal_nb(b_sz), tol(std::numeric_limits<T>::epsilon()),
          qrcp_wide(Subroutines::luqr), qr_tall(Subroutines::geqrf), ctx_(&c) {
        if (b_sz <= 0) throw Error(RLB200_ERR_ARG, "BQRRP block size b_sz must be > 0");      // rl_bqrrp.hh:66
    }
    virtual ~BQRRP() {}
    int call(int64_t m, int64_t n, T* A, int64_t lda, T d_factor, T* tau, int64_t* J, state_t& state) RLB200_OVERRIDE {
        uint32_t w[6]; state_to_words(state, w);
        int64_t r = 0;
        if (timing) ctx_->phase_timing(true);
        ctx_->check(rlb200_set_bqrrp_tol(ctx_->get(), (double)tol));
        int rc = ctx_->check(detail::abi<T>::bqrrp_host(ctx_->get(), m, n, A, lda, d_factor, block_size, (int)qrcp_wide, (int)qr_tall, tau, J, &r, w));
        if (timing) { times = ctx_->phase_times(); ctx_->phase_timing(false); }
        words_to_state(w, state);
        rank = r;
        return rc;
    }
    bool timing;
    int64_t rank, block_size, internal_nb;
    T tol;
    std::vector<long> times;
    Subroutines::QRCPWide qrcp_wide;
    Subroutines::QRTall qr_tall;
private:
    Context* ctx_;
};

// CQRRPT_GPU (rl_cqrrpt_gpu.hh:15-387): the reference's hybrid driver - HOST pointers in and out, the O(m n^2) part on the GPU.  Same
// constructor (verb, time_subroutines, eps), public fields and call signature; here every phase runs on the device (rlb200_cqrrpt_*_host).
// QRCP of the sketch: geqp3 when `no_hqrrp` (the constructor default, :62), else hqrrp with nb_alg / oversampling / panel_pivoting / use_cholqr
// (:222-226).  Define RLB200_WITH_RANDLAPACK_GPU after including the reference's header to derive from RandLAPACK::CQRRPT_GPU_alg.
template <typename T>
class CQRRPT_GPU
#if defined(RLB200_WITH_RANDLAPACK) && defined(RLB200_WITH_RANDLAPACK_GPU)
    : public RandLAPACK::CQRRPT_GPU_alg<T, r123::Philox4x32>
#endif
{
public:
    CQRRPT_GPU(bool verb, bool time_subroutines, T ep) : CQRRPT_GPU(default_context(), verb, time_subroutines, ep) {}
    CQRRPT_GPU(Context& c, bool verb, bool time_subroutines, T ep)
        : verbosity(verb), timing(time_subroutines), eps(ep), rank(0), num_threads(0), nnz(2), no_hqrrp(1), nb_alg(64), oversampling(10),
          panel_pivoting(1), use_cholqr(0), ctx_(&c) {}
    virtual ~CQRRPT_GPU() {}
    int call(int64_t m, int64_t n, T* A, int64_t lda, T* R, int64_t ldr, int64_t* J, T d_factor, state_t& state)
#if defined(RLB200_WITH_RANDLAPACK) && defined(RLB200_WITH_RANDLAPACK_GPU)
        override
#endif
    {
        uint32_t w[6]; state_to_words(state, w);
        int64_t r = 0;
        if (timing) ctx_->phase_timing(true);
        ctx_->check(rlb200_set_cqrrpt_qrcp(ctx_->get(), no_hqrrp ? RLB200_CQRRPT_QRCP_GEQP3 : RLB200_CQRRPT_QRCP_HQRRP));
        ctx_->check(rlb200_set_cqrrpt_orthogonalization(ctx_->get(), 0));
        ctx_->check(rlb200_set_cqrrpt_hqrrp_opts(ctx_->get(), nb_alg, oversampling, (int)panel_pivoting, (int)use_cholqr));
        int rc = ctx_->check(detail::abi<T>::cqrrpt_host(ctx_->get(), m, n, A, lda, R, ldr, J, d_factor, eps, nnz, &r, w));
        if (timing) { times = ctx_->phase_times(); ctx_->phase_timing(false); }
        words_to_state(w, state);
        rank = r;
        return rc;
    }
    bool verbosity;
    bool timing;
    T eps;
    int64_t rank;
    std::vector<long> times;   // 8 entries when `timing` (rl_cqrrpt_gpu.hh:133-134)
    int num_threads;           // SASO tuning knob of the CPU sketch (:137); unused here
    int64_t nnz;
    int no_hqrrp;              // :141
    int64_t nb_alg, oversampling, panel_pivoting, use_cholqr;
private:
    Context* ctx_;
};

// BQRRP_GPU (rl_bqrrp_gpu.hh:27-942): same constructor (time_subroutines, b_sz), public fields and call signature.  As in the
// reference ALL pointers are DEVICE pointers and the d x n sketch A_sk is an input (overwritten).  qr_tall: geqrf (ctor default,
// rl_bqrrp_gpu.hh:86) or cholqr.  The reference's base class lives in a header that needs its CUDA build (USE_CUDA + blaspp's device
// API); define RLB200_WITH_RANDLAPACK_GPU after including it to derive from RandLAPACK::BQRRP_GPU_alg.
struct BQRRPGPUSubroutines {
    enum QRTall { cholqr = RLB200_QRTALL_CHOLQR, geqrf = RLB200_QRTALL_GEQRF };
};
template <typename T>
class BQRRP_GPU
#if defined(RLB200_WITH_RANDLAPACK) && defined(RLB200_WITH_RANDLAPACK_GPU)
    : public RandLAPACK::BQRRP_GPU_alg<T, r123::Philox4x32>
#endif
{
public:
    using GPUSubroutine = BQRRPGPUSubroutines;
    BQRRP_GPU(bool time_subroutines, int64_t b_sz) : BQRRP_GPU(default_context(), time_subroutines, b_sz) {}
    BQRRP_GPU(Context& c, bool time_subroutines, int64_t b_sz)
        : timing(time_subroutines), rank(0), block_size(b_sz), tol(std::numeric_limits<T>::epsilon()), qr_tall(GPUSubroutine::geqrf), ctx_(&c) {
        if (b_sz <= 0) throw Error(RLB200_ERR_ARG, "BQRRP_GPU block size b_sz must be > 0");
    }
    virtual ~BQRRP_GPU() {}
    int call(int64_t m, int64_t n, T* A, int64_t lda, T* A_sk, int64_t d, T* tau, int64_t* J)
#if defined(RLB200_WITH_RANDLAPACK) && defined(RLB200_WITH_RANDLAPACK_GPU)
        override
#endif
    {
        int64_t r = 0;
        if (timing) ctx_->phase_timing(true);
        ctx_->check(rlb200_set_bqrrp_tol(ctx_->get(), (double)tol));
        int rc = ctx_->check(detail::abi<T>::bqrrp_dev_sk(ctx_->get(), m, n, A, lda, A_sk, d, block_size, (int)qr_tall, tau, J, &r));
        if (timing) { times = ctx_->phase_times(); ctx_->phase_timing(false); }
        rank = r;
        return rc;
    }
    bool timing;
    state_t state;
    int64_t rank, block_size;
    std::vector<long> times;
    T tol;
    GPUSubroutine::QRTall qr_tall;
private:
    Context* ctx_;
};

// ---------------------------------------------------------------------------------------------------------------------
// SYPS / SYRF / REVD2 (rl_syps.hh:21-143, rl_syrf.hh:21-118, rl_revd2.hh:75-246) for an explicit symmetric matrix: the reference's
// constructors, public fields and the `call(uplo, m, A, ...)` overloads with HOST pointers (the SymmetricLinearOperator overloads are
// not offered: the device path needs the matrix itself).  Failures the reference reports by throwing std::runtime_error throw here too.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RLB200_WITH_RANDLAPACK
enum class Uplo : char { Upper = 'U', Lower = 'L' };
using uplo_t = Uplo;
inline int uplo_code(Uplo u) { return u == Uplo::Upper ? RLB200_UPLO_UPPER : RLB200_UPLO_LOWER; }
#else
using uplo_t = blas::Uplo;
inline int uplo_code(blas::Uplo u) { return u == blas::Uplo::Upper ? RLB200_UPLO_UPPER : RLB200_UPLO_LOWER; }
#endif

template <typename T>
class SYPS {
public:
    using scalar_t = T;
    SYPS(int64_t p, int64_t q, bool verb, bool cond) : SYPS(default_context(), p, q, verb, cond) {}
    SYPS(Context& c, int64_t p, int64_t q, bool verb, bool cond) : passes_over_data(p), passes_per_stab(q), verbose(verb), cond_check(cond), ctx_(&c) {}
    // skop_buff (m x k; allocated with new[] when null, as the reference does) <- the power sketch; work_buff is not needed (rl_syps.hh:47-57)
    int call(uplo_t uplo, int64_t m, const T* A, int64_t lda, int64_t k, state_t& state, T*& skop_buff, T* /*work_buff*/) {
        if (!skop_buff) skop_buff = new T[m * k];
        std::vector<T> Ac((size_t)m * m);
        for (int64_t j = 0; j < m; ++j) std::memcpy(Ac.data() + j * m, A + j * lda, sizeof(T) * m);
        detail::DevBuf<T> dA(*ctx_, m * m, Ac.data()), dS(*ctx_, m * k), dW(*ctx_, m * k);
        uint32_t w[6]; state_to_words(state, w);
        int rc = ctx_->check(detail::abi<T>::syps(ctx_->get(), uplo_code(uplo), m, dA.ptr(), m, k, passes_over_data, passes_per_stab, dS.ptr(), dW.ptr(), w));
        words_to_state(w, state);
        dS.to_host(skop_buff, m * k);
        return rc;
    }
    Context& context() const { return *ctx_; }
    int64_t passes_over_data, passes_per_stab;
    bool verbose, cond_check;
    std::vector<T> cond_nums;
private:
    Context* ctx_;
};

template <typename T>
class SYRF {
public:
    SYRF(SYPS<T>& syps_obj, StabBase<T>& orth_obj, bool verb = false, bool cond = false) : syps(syps_obj), orth(orth_obj), verbose(verb), cond_check(cond) {}
    // Q (resized to m x k) <- orth(A * syps(A))  (rl_syrf.hh:43-56)
    int call(uplo_t uplo, int64_t m, const T* A, int64_t k, std::vector<T>& Q, state_t& state, T* /*work_buff*/) {
        Context& c = syps.context();
        rlb200_revd2_opts o = opts(0);
        if ((int64_t)Q.size() < m * k) Q.resize(m * k);
        detail::DevBuf<T> dA(c, m * m, A), dQ(c, m * k), dW(c, m * k);
        uint32_t w[6]; state_to_words(state, w);
        int rc = c.check(detail::abi<T>::syrf(c.get(), uplo_code(uplo), m, dA.ptr(), m, k, dQ.ptr(), dW.ptr(), w, &o));
        words_to_state(w, state);
        if (rc == 2) throw std::runtime_error("Orthogonalization failed.");
        dQ.to_host(Q.data(), m * k);
        return rc;
    }
    rlb200_revd2_opts opts(int error_est_p) const {
        rlb200_revd2_opts o;
        o.syps_passes = syps.passes_over_data; o.syps_passes_per_stab = syps.passes_per_stab;
        o.orth = require_device_stab<T>(orth, "rlb200::SYRF").kind(); o.error_est_p = error_est_p;
        return o;
    }
    SYPS<T>& syps;
    StabBase<T>& orth;
    bool verbose, cond_check;
    std::vector<T> cond_nums;
};

template <typename T>
class REVD2 {
public:
    REVD2(SYRF<T>& syrf_obj, int error_est_power_iters, bool verb = false) : syrf(syrf_obj), error_est_p(error_est_power_iters), verbose(verb) {}
    // V (resized to m x k) and eigvals (k) for the final k (rl_revd2.hh:120-139)
    int call(uplo_t uplo, int64_t m, const T* A, int64_t& k, T tol, std::vector<T>& V, std::vector<T>& eigvals, state_t& state) {
        Context& c = syrf.syps.context();
        rlb200_revd2_opts o = syrf.opts(error_est_p);
        // k doubles until the error estimate passes (up to m): the outputs are sized for 8 k and, if the rank estimate outgrows that
        // (code 3), the call is repeated from the caller's state with four times the room - the algorithm is deterministic, so the
        // repeated prefix reproduces itself and V never needs m x m entries up front
        const int64_t k_in = k;
        int64_t k_cap = std::min<int64_t>(m, std::max<int64_t>(8 * k_in, 64));
        uint32_t w[6];
        T err = 0;
        int rc = 0;
        while (true) {
            state_to_words(state, w);
            k = k_in;
            V.resize((size_t)m * k_cap); eigvals.resize((size_t)k_cap);
            rc = c.check(detail::abi<T>::revd2_host(c.get(), uplo_code(uplo), m, A, m, &k, k_cap, tol, V.data(), eigvals.data(), w, &o, &err));
            if (rc != 3 || k_cap >= m) break;
            k_cap = std::min<int64_t>(m, 4 * k_cap);
        }
        words_to_state(w, state);
        if (rc == 1) throw std::runtime_error("Cholesky decomposition failed.");
        if (rc == 2) throw std::runtime_error("Orthogonalization failed.");
        V.resize((size_t)m * k); eigvals.resize((size_t)k);
        last_error_estimate = err;
        return rc;
    }
    SYRF<T>& syrf;
    int error_est_p;
    bool verbose;
    T last_error_estimate = 0;
};

// ---------------------------------------------------------------------------------------------------------------------
// Linear operators with the matrix RESIDENT ON THE DEVICE.  They satisfy the reference's LinearOperator /
// SymmetricLinearOperator concepts (linops/rl_concepts.hh:30-57: n_rows / n_cols / dim members and the GEMM- / SYMM-like call with HOST
// B and C), so the reference's operator-templated algorithms (SYPS / SYRF / REVD2 `call(SLO&, ...)`, rl_revd2.hh:142-150) run their
// products on the B200 while the matrix crosses PCIe once, at construction.  ColMajor only; op(A), op(B) not both transposed.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RLB200_WITH_RANDLAPACK
enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T' };
enum class Side : char { Left = 'L', Right = 'R' };
using layout_t = Layout; using op_t = Op; using side_t = Side;
inline bool is_colmajor(Layout l) { return l == Layout::ColMajor; }
inline int op_code(Op o) { return o == Op::NoTrans ? 0 : 1; }
inline bool is_left(Side s) { return s == Side::Left; }
#else
using layout_t = blas::Layout; using op_t = blas::Op; using side_t = blas::Side;
inline bool is_colmajor(blas::Layout l) { return l == blas::Layout::ColMajor; }
inline int op_code(blas::Op o) { return o == blas::Op::NoTrans ? 0 : 1; }
inline bool is_left(blas::Side s) { return s == blas::Side::Left; }
#endif

template <typename T>
struct DenseLinOp {
    using scalar_t = T;
    const int64_t n_rows;
    const int64_t n_cols;
    // A_host: n_rows x n_cols column-major (lda >= n_rows, rl_dense_linop.hh:52-58); copied to the device here
    DenseLinOp(int64_t rows, int64_t cols, const T* A_host, int64_t lda) : DenseLinOp(default_context(), rows, cols, A_host, lda) {}
    DenseLinOp(Context& c, int64_t rows, int64_t cols, const T* A_host, int64_t lda) : n_rows(rows), n_cols(cols), ctx_(&c) {
        if (lda < rows) throw Error(RLB200_ERR_ARG, "DenseLinOp: lda must be >= n_rows under ColMajor");
        std::vector<T> packed;
        const T* src = A_host;
        double ss = 0;
        if (lda != rows) {
            packed.resize((size_t)rows * cols);
            for (int64_t j = 0; j < cols; ++j) std::memcpy(packed.data() + j * rows, A_host + j * lda, sizeof(T) * rows);
            src = packed.data();
        }
        for (int64_t i = 0; i < rows * cols; ++i) ss += (double)src[i] * (double)src[i];
        fro_ = (T)std::sqrt(ss);
        dA_ = std::make_shared<detail::DevBuf<T>>(c, rows * cols, src);
    }
    T fro_nrm() { return fro_; }
    // C := alpha * op(A) * op(B) + beta * C with HOST B and C (rl_dense_linop.hh:70-84)
    void operator()(layout_t layout, op_t trans_A, op_t trans_B, int64_t m, int64_t n, int64_t k, T alpha, const T* B, int64_t ldb, T beta, T* C,
                    int64_t ldc) {
        if (!is_colmajor(layout)) throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::DenseLinOp: ColMajor only");
        const int ta = op_code(trans_A), tb = op_code(trans_B);
        const int64_t rows_A = ta ? k : m, cols_A = ta ? m : k, rows_B = tb ? n : k, cols_B = tb ? k : n;
        if (rows_A != n_rows || cols_A != n_cols) throw Error(RLB200_ERR_ARG, "DenseLinOp: (m, k, trans_A) do not match the operator");   // :103-106
        if (ldb < rows_B || ldc < m) throw Error(RLB200_ERR_ARG, "DenseLinOp: ldb / ldc too small");                                        // :109-111
        std::vector<T> Bp((size_t)rows_B * cols_B), Cp((size_t)m * n);
        for (int64_t j = 0; j < cols_B; ++j) std::memcpy(Bp.data() + j * rows_B, B + j * ldb, sizeof(T) * rows_B);
        if (beta != (T)0) for (int64_t j = 0; j < n; ++j) std::memcpy(Cp.data() + j * m, C + j * ldc, sizeof(T) * m);
        detail::DevBuf<T> dB(*ctx_, rows_B * cols_B, Bp.data()), dC(*ctx_, m * n, beta != (T)0 ? Cp.data() : nullptr);
        ctx_->check(detail::abi<T>::gemm(ctx_->get(), ta, tb, m, n, k, alpha, dA_->ptr(), n_rows, dB.ptr(), rows_B, beta, dC.ptr(), m));
        dC.to_host(Cp.data(), m * n);
        for (int64_t j = 0; j < n; ++j) std::memcpy(C + j * ldc, Cp.data() + j * m, sizeof(T) * m);
        ++n_products;
    }
    // The overload with an explicit side (rl_dense_linop.hh:91-142): Side::Left as above; Side::Right: C := alpha * op(B) * op(A) + beta * C.
    // This is the form CholQR_linops / sCholQR3_linops / CQRRT_linops call (rl_cholqr_linops.hh:169-171, rl_cqrrt_linops.hh:275-283).
    void operator()(side_t side, layout_t layout, op_t trans_A, op_t trans_B, int64_t m, int64_t n, int64_t k, T alpha, const T* B, int64_t ldb,
                    T beta, T* C, int64_t ldc) {
        if (is_left(side)) { (*this)(layout, trans_A, trans_B, m, n, k, alpha, B, ldb, beta, C, ldc); return; }
        if (!is_colmajor(layout)) throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::DenseLinOp: ColMajor only");
        const int ta = op_code(trans_A), tb = op_code(trans_B);
        const int64_t rows_B = tb ? k : m, cols_B = tb ? m : k, rows_A = ta ? n : k, cols_A = ta ? k : n;
        if (rows_A != n_rows || cols_A != n_cols) throw Error(RLB200_ERR_ARG, "DenseLinOp: (k, n, trans_A) do not match the operator");   // :131-132
        if (ldb < rows_B || ldc < m) throw Error(RLB200_ERR_ARG, "DenseLinOp: ldb / ldc too small");                                        // :135-137
        std::vector<T> Bp((size_t)rows_B * cols_B), Cp((size_t)m * n);
        for (int64_t j = 0; j < cols_B; ++j) std::memcpy(Bp.data() + j * rows_B, B + j * ldb, sizeof(T) * rows_B);
        if (beta != (T)0) for (int64_t j = 0; j < n; ++j) std::memcpy(Cp.data() + j * m, C + j * ldc, sizeof(T) * m);
        detail::DevBuf<T> dB(*ctx_, rows_B * cols_B, Bp.data()), dC(*ctx_, m * n, beta != (T)0 ? Cp.data() : nullptr);
        ctx_->check(detail::abi<T>::gemm(ctx_->get(), tb, ta, m, n, k, alpha, dB.ptr(), rows_B, dA_->ptr(), n_rows, beta, dC.ptr(), m));
        dC.to_host(Cp.data(), m * n);
        for (int64_t j = 0; j < n; ++j) std::memcpy(C + j * ldc, Cp.data() + j * m, sizeof(T) * m);
        ++n_products;
    }
#ifdef RLB200_WITH_RANDLAPACK
    // Sketching-operator overloads (rl_dense_linop.hh:236-300), Side::Right, ColMajor, NoTrans / NoTrans - the call CQRRT_linops makes
    // (rl_cqrrt_linops.hh:204, 211): C (d x n, HOST) := alpha * S * A + beta * C.  The operator never crosses PCIe: the device regenerates it
    // from S.dist and S.seed_state (bit-exact triplets / counter rule) and applies it to the resident matrix.
    void operator()(side_t side, layout_t layout, op_t trans_A, op_t trans_S, int64_t d, int64_t n, int64_t m, T alpha,
                    RandBLAS::SparseSkOp<T, r123::Philox4x32>& S, T beta, T* C, int64_t ldc) {
        sk_check(side, layout, trans_A, trans_S, d, n, m, S.dist.n_rows, S.dist.n_cols, ldc);
        if (S.dist.major_axis != RandBLAS::Axis::Short) throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::DenseLinOp: short-axis (SASO) sparse operators only");
        uint32_t w[6]; state_to_words(S.seed_state, w);
        sk_apply(d, n, beta, C, ldc, [&](T* dC) {
            return detail::abi<T>::sketch_sparse_left(ctx_->get(), S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz, d, n, m, alpha, 0, 0, dA_->ptr(), n_rows, beta, dC, d, w);
        });
    }
    void operator()(side_t side, layout_t layout, op_t trans_A, op_t trans_S, int64_t d, int64_t n, int64_t m, T alpha,
                    RandBLAS::DenseSkOp<T, r123::Philox4x32>& S, T beta, T* C, int64_t ldc) {
        sk_check(side, layout, trans_A, trans_S, d, n, m, S.dist.n_rows, S.dist.n_cols, ldc);
        uint32_t w[6]; state_to_words(S.seed_state, w);
        const int fam = S.dist.family == RandBLAS::ScalarDist::Uniform ? RLB200_FAMILY_UNIFORM : RLB200_FAMILY_GAUSSIAN;
        const int ax = S.dist.major_axis == RandBLAS::Axis::Short ? RLB200_AXIS_SHORT : RLB200_AXIS_LONG;
        sk_apply(d, n, beta, C, ldc, [&](T* dC) {
            return detail::abi<T>::sketch_dense_left(ctx_->get(), S.dist.n_rows, S.dist.n_cols, fam, ax, d, n, m, alpha, 0, 0, dA_->ptr(), n_rows, beta, dC, d, w);
        });
    }
#endif
    T* device_ptr() { return dA_->ptr(); }
    int64_t n_products = 0;      // products executed on the device so far
private:
#ifdef RLB200_WITH_RANDLAPACK
    void sk_check(side_t side, layout_t layout, op_t trans_A, op_t trans_S, int64_t d, int64_t n, int64_t m, int64_t S_rows, int64_t S_cols, int64_t ldc) {
        if (is_left(side) || !is_colmajor(layout) || op_code(trans_A) || op_code(trans_S))
            throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::DenseLinOp with a sketching operator: Side::Right, ColMajor, NoTrans / NoTrans only");
        if (m != n_rows || n != n_cols || S_rows != d || S_cols != m) throw Error(RLB200_ERR_ARG, "DenseLinOp: (d, n, m) do not match the operators");
        if (ldc < d) throw Error(RLB200_ERR_ARG, "DenseLinOp: ldc too small");
    }
    template <typename F>
    void sk_apply(int64_t d, int64_t n, T beta, T* C, int64_t ldc, F&& f) {
        std::vector<T> Cp((size_t)d * n);
        if (beta != (T)0) for (int64_t j = 0; j < n; ++j) std::memcpy(Cp.data() + j * d, C + j * ldc, sizeof(T) * d);
        detail::DevBuf<T> dC(*ctx_, d * n, beta != (T)0 ? Cp.data() : nullptr);
        ctx_->check(f(dC.ptr()));
        dC.to_host(Cp.data(), d * n);
        for (int64_t j = 0; j < n; ++j) std::memcpy(C + j * ldc, Cp.data() + j * d, sizeof(T) * d);
        ++n_products;
    }
#endif
    Context* ctx_;
    std::shared_ptr<detail::DevBuf<T>> dA_;
    T fro_ = 0;
};

// ExplicitSymLinOp (linops/rl_sym_linops.hh:40-99): only the `uplo` triangle of A_host is read; the mirrored matrix lives on the device
template <typename T>
struct ExplicitSymLinOp {
    using scalar_t = T;
    const int64_t dim;
    const int64_t n_rows;
    const int64_t n_cols;
    ExplicitSymLinOp(int64_t d, uplo_t uplo, const T* A_host, int64_t lda) : ExplicitSymLinOp(default_context(), d, uplo, A_host, lda) {}
    ExplicitSymLinOp(Context& c, int64_t d, uplo_t uplo, const T* A_host, int64_t lda) : dim(d), n_rows(d), n_cols(d), op_(c, d, d, mirror(d, uplo, A_host, lda).data(), d) {}
    // C := alpha * A * B + beta * C, B and C with n columns (rl_sym_linops.hh:76-99)
    void operator()(layout_t layout, int64_t n, T alpha, const T* B, int64_t ldb, T beta, T* C, int64_t ldc) {
#ifndef RLB200_WITH_RANDLAPACK
        op_(layout, Op::NoTrans, Op::NoTrans, dim, n, dim, alpha, B, ldb, beta, C, ldc);
#else
        op_(layout, blas::Op::NoTrans, blas::Op::NoTrans, dim, n, dim, alpha, B, ldb, beta, C, ldc);
#endif
    }
    int64_t n_products() const { return op_.n_products; }
private:
    static std::vector<T> mirror(int64_t d, uplo_t uplo, const T* A, int64_t lda) {
        if (lda < d) throw Error(RLB200_ERR_ARG, "ExplicitSymLinOp: lda must be >= dim");
        const bool upper = uplo_code(uplo) == RLB200_UPLO_UPPER;
        std::vector<T> F((size_t)d * d);
        for (int64_t j = 0; j < d; ++j)
            for (int64_t i = 0; i < d; ++i) {
                const bool valid = upper ? (i <= j) : (i >= j);
                F[i + j * d] = valid ? A[i + j * lda] : A[j + i * lda];
            }
        return F;
    }
    DenseLinOp<T> op_;
};

#ifdef RLB200_WITH_RANDLAPACK
// ---------------------------------------------------------------------------------------------------------------------
// RandBLAS::sketch_general (RandBLAS/RandBLAS/skge.hh:859-960 left, :1031-1131 right) with the reference's argument lists: the sketching
// operator object is only read for its distribution and seed state - the device regenerates it - so nothing but A and B crosses PCIe
// (sketch_general: HOST pointers, copies inside, the drop-in form) or nothing at all (sketch_general_dev: DEVICE pointers).
// Dense operators: every layout / transposition flag.  Sparse operators: short-axis with every flag (op(S) wide), long-axis wide operators
// from the left with ColMajor / NoTrans / NoTrans.  S.next_state is what the reference computed at construction; it is not touched.
// ---------------------------------------------------------------------------------------------------------------------
namespace detail {
inline int layout_code(blas::Layout l) { return l == blas::Layout::ColMajor ? RLB200_LAYOUT_COLMAJOR : RLB200_LAYOUT_ROWMAJOR; }
inline int64_t buf_len(blas::Layout l, int64_t rows, int64_t cols, int64_t ld) { return l == blas::Layout::ColMajor ? ld * cols : rows * ld; }
template <typename T> struct skabi;
template <> struct skabi<double> {
    static constexpr auto dl = rlb200_sketch_general_dense_left_f64_dev; static constexpr auto dr = rlb200_sketch_general_dense_right_f64_dev;
    static constexpr auto sl = rlb200_sketch_general_sparse_left_f64_dev; static constexpr auto sr = rlb200_sketch_general_sparse_right_f64_dev;
    static constexpr auto laso = rlb200_sketch_sparse_left_laso_f64_dev;
};
template <> struct skabi<float> {
    static constexpr auto dl = rlb200_sketch_general_dense_left_f32_dev; static constexpr auto dr = rlb200_sketch_general_dense_right_f32_dev;
    static constexpr auto sl = rlb200_sketch_general_sparse_left_f32_dev; static constexpr auto sr = rlb200_sketch_general_sparse_right_f32_dev;
    static constexpr auto laso = rlb200_sketch_sparse_left_laso_f32_dev;
};
}  // namespace detail

// B(d x n) = alpha * op(submat(S)) * op(A) + beta * B, DEVICE A and B
template <typename T>
void sketch_general_dev(Context& c, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                        const RandBLAS::DenseSkOp<T, r123::Philox4x32>& S, int64_t ro_s, int64_t co_s, const T* A_dev, int64_t lda, T beta, T* B_dev, int64_t ldb) {
    uint32_t w[6]; state_to_words(S.seed_state, w);
    const int fam = S.dist.family == RandBLAS::ScalarDist::Uniform ? RLB200_FAMILY_UNIFORM : RLB200_FAMILY_GAUSSIAN;
    const int ax = S.dist.major_axis == RandBLAS::Axis::Short ? RLB200_AXIS_SHORT : RLB200_AXIS_LONG;
    c.check(detail::skabi<T>::dl(c.get(), detail::layout_code(layout), op_code(opS), op_code(opA), d, n, m, alpha, S.dist.n_rows, S.dist.n_cols, fam, ax,
                                 ro_s, co_s, A_dev, lda, beta, B_dev, ldb, w));
}
template <typename T>
void sketch_general_dev(Context& c, blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha,
                        const RandBLAS::SparseSkOp<T, r123::Philox4x32>& S, int64_t ro_s, int64_t co_s, const T* A_dev, int64_t lda, T beta, T* B_dev, int64_t ldb) {
    uint32_t w[6]; state_to_words(S.seed_state, w);
    if (S.dist.major_axis == RandBLAS::Axis::Long) {
        if (layout != blas::Layout::ColMajor || op_code(opS) || op_code(opA))
            throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::sketch_general: long-axis sparse operators with ColMajor / NoTrans / NoTrans only");
        c.check(detail::skabi<T>::laso(c.get(), S.dist.n_rows, S.dist.n_cols, S.dist.vec_nnz, d, n, m, alpha, ro_s, co_s, A_dev, lda, beta, B_dev, ldb, w));
        return;
    }
    c.check(detail::skabi<T>::sl(c.get(), detail::layout_code(layout), op_code(opS), op_code(opA), d, n, m, alpha, S.dist.n_rows, S.dist.n_cols,
                                 S.dist.vec_nnz, ro_s, co_s, A_dev, lda, beta, B_dev, ldb, w));
}
// B(m x d) = alpha * op(A) * op(submat(S)) + beta * B, DEVICE A and B
template <typename T>
void sketch_general_dev(Context& c, blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A_dev, int64_t lda,
                        const RandBLAS::DenseSkOp<T, r123::Philox4x32>& S, int64_t ro_s, int64_t co_s, T beta, T* B_dev, int64_t ldb) {
    uint32_t w[6]; state_to_words(S.seed_state, w);
    const int fam = S.dist.family == RandBLAS::ScalarDist::Uniform ? RLB200_FAMILY_UNIFORM : RLB200_FAMILY_GAUSSIAN;
    const int ax = S.dist.major_axis == RandBLAS::Axis::Short ? RLB200_AXIS_SHORT : RLB200_AXIS_LONG;
    c.check(detail::skabi<T>::dr(c.get(), detail::layout_code(layout), op_code(opA), op_code(opS), m, d, n, alpha, A_dev, lda, S.dist.n_rows, S.dist.n_cols,
                                 fam, ax, ro_s, co_s, beta, B_dev, ldb, w));
}
template <typename T>
void sketch_general_dev(Context& c, blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A_dev, int64_t lda,
                        const RandBLAS::SparseSkOp<T, r123::Philox4x32>& S, int64_t ro_s, int64_t co_s, T beta, T* B_dev, int64_t ldb) {
    if (S.dist.major_axis == RandBLAS::Axis::Long) throw Error(RLB200_ERR_UNSUPPORTED, "rlb200::sketch_general: long-axis sparse operators from the left only");
    uint32_t w[6]; state_to_words(S.seed_state, w);
    c.check(detail::skabi<T>::sr(c.get(), detail::layout_code(layout), op_code(opA), op_code(opS), m, d, n, alpha, A_dev, lda, S.dist.n_rows, S.dist.n_cols,
                                 S.dist.vec_nnz, ro_s, co_s, beta, B_dev, ldb, w));
}
// HOST pointers: RandBLAS::sketch_general's own signatures (left: skge.hh:859-905 / 907-960; right: :1031-1076 / 1078-1131)
template <typename T, typename SKOP>
void sketch_general(blas::Layout layout, blas::Op opS, blas::Op opA, int64_t d, int64_t n, int64_t m, T alpha, const SKOP& S, int64_t ro_s, int64_t co_s,
                    const T* A, int64_t lda, T beta, T* B, int64_t ldb) {
    Context& c = default_context();
    const int64_t ra = op_code(opA) ? n : m, ca = op_code(opA) ? m : n;
    const int64_t la = detail::buf_len(layout, ra, ca, lda), lb = detail::buf_len(layout, d, n, ldb);
    detail::DevBuf<T> dA(c, std::max<int64_t>(la, 1), la > 0 ? A : nullptr), dB(c, std::max<int64_t>(lb, 1), lb > 0 ? B : nullptr);
    sketch_general_dev<T>(c, layout, opS, opA, d, n, m, alpha, S, ro_s, co_s, dA.ptr(), lda, beta, dB.ptr(), ldb);
    if (lb > 0) dB.to_host(B, lb);
}
template <typename T, typename SKOP>
void sketch_general(blas::Layout layout, blas::Op opA, blas::Op opS, int64_t m, int64_t d, int64_t n, T alpha, const T* A, int64_t lda, const SKOP& S,
                    int64_t ro_s, int64_t co_s, T beta, T* B, int64_t ldb) {
    Context& c = default_context();
    const int64_t ra = op_code(opA) ? n : m, ca = op_code(opA) ? m : n;
    const int64_t la = detail::buf_len(layout, ra, ca, lda), lb = detail::buf_len(layout, m, d, ldb);
    detail::DevBuf<T> dA(c, std::max<int64_t>(la, 1), la > 0 ? A : nullptr), dB(c, std::max<int64_t>(lb, 1), lb > 0 ? B : nullptr);
    sketch_general_dev<T>(c, layout, opA, opS, m, d, n, alpha, dA.ptr(), lda, S, ro_s, co_s, beta, dB.ptr(), ldb);
    if (lb > 0) dB.to_host(B, lb);
}
#endif

}  // namespace rlb200
