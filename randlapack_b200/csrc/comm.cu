// Native collectives for row-sharded operation: an NCCL communicator owned by the context, so that a C/C++ caller of the reference's
// API gets the NCCL/NVLink data plane without writing one (the function-pointer hook of rlb200_set_shard stays for tests and for hosts
// that already own a communicator).  The reference has no distributed layer: this is net-new.
//
// libnccl is opened at run time (dlopen): librlb200.so keeps no link-time dependency on it, single-GPU users never load it, and inside a
// process that already has NCCL loaded (e.g. PyTorch) the same library instance is reused.
#include "drivers.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace rlb {

namespace {
// the few declarations of nccl.h that are needed (NCCL 2.x ABI: ncclUniqueId is 128 bytes; data types and reduction ops are stable enums)
typedef void* nccl_comm_t;
struct nccl_uid { char internal[128]; };
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT32 = 2, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(nccl_comm_t*, int, nccl_uid, int);
typedef int (*fn_destroy)(nccl_comm_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_errstr)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_errstr errstr = nullptr;
};
NcclApi g_api;

int load_nccl(std::string* err) {
    if (g_api.lib) return 0;
    const char* names[] = {getenv("RLB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { if (err) *err = std::string("libnccl could not be opened (set RLB200_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return RLB200_ERR_COLLECTIVE; }
    NcclApi a;
    a.lib = lib;
    a.get_uid = (fn_get_uid)dlsym(lib, "ncclGetUniqueId");
    a.init_rank = (fn_init_rank)dlsym(lib, "ncclCommInitRank");
    a.destroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
    a.allreduce = (fn_allreduce)dlsym(lib, "ncclAllReduce");
    a.errstr = (fn_errstr)dlsym(lib, "ncclGetErrorString");
    if (!a.get_uid || !a.init_rank || !a.destroy || !a.allreduce) { if (err) *err = "libnccl lacks the expected entry points"; return RLB200_ERR_COLLECTIVE; }
    g_api = a;
    return 0;
}

// the rlb200_allreduce_fn of a context that owns a communicator: sum-allreduce in place on the context's stream
int native_allreduce(void* user, void* buf, int64_t count, int32_t elem_size, void* stream) {
    Ctx* ctx = static_cast<Ctx*>(user);
    if (!ctx->nccl_comm) return 1;
    const int dt = elem_size == 8 ? NCCL_FLOAT64 : NCCL_FLOAT32;
    const int rc = g_api.allreduce(buf, buf, (size_t)count, dt, NCCL_SUM, ctx->nccl_comm, static_cast<cudaStream_t>(stream));
    return rc == NCCL_SUCCESS ? 0 : 100 + rc;
}
}  // namespace

int comm_unique_id(unsigned char out[128], std::string* err) {
    int rc = load_nccl(err);
    if (rc) return rc;
    nccl_uid id;
    const int r = g_api.get_uid(&id);
    if (r != NCCL_SUCCESS) { if (err) *err = std::string("ncclGetUniqueId: ") + (g_api.errstr ? g_api.errstr(r) : "failed"); return RLB200_ERR_COLLECTIVE; }
    std::memcpy(out, id.internal, 128);
    return 0;
}

int comm_init(Ctx* ctx, int nranks, int rank, const unsigned char id_bytes[128]) {
    RLB_REQUIRE(ctx, nranks >= 1 && rank >= 0 && rank < nranks && id_bytes != nullptr);
    RLB_CHECK(load_nccl(&ctx->err));
    if (ctx->nccl_comm) { g_api.destroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
    nccl_uid id;
    std::memcpy(id.internal, id_bytes, 128);
    nccl_comm_t comm = nullptr;
    const int r = g_api.init_rank(&comm, nranks, id, rank);
    if (r != NCCL_SUCCESS) { ctx->err = std::string("ncclCommInitRank: ") + (g_api.errstr ? g_api.errstr(r) : "failed"); return RLB200_ERR_COLLECTIVE; }
    ctx->nccl_comm = comm;
    ctx->nccl_lib = g_api.lib;
    ctx->shard_rank = rank;
    ctx->shard_world = nranks;
    ctx->allreduce = native_allreduce;
    ctx->allreduce_user = ctx;
    return 0;
}

void comm_destroy(Ctx* ctx) {
    if (ctx->nccl_comm && g_api.destroy) g_api.destroy(ctx->nccl_comm);
    if (ctx->allreduce == native_allreduce) { ctx->allreduce = nullptr; ctx->allreduce_user = nullptr; }
    ctx->nccl_comm = nullptr;
}

bool comm_is_native(const Ctx* ctx) { return ctx->nccl_comm != nullptr && ctx->allreduce == native_allreduce; }
// (re)install the context's own communicator as the sum-allreduce; false when the context has none
bool comm_use_native(Ctx* ctx) {
    if (!ctx->nccl_comm) return false;
    ctx->allreduce = native_allreduce;
    ctx->allreduce_user = ctx;
    return true;
}

}  // namespace rlb
