// Internal declarations shared by the translation units of libtrx2dyn.so.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/trx2dyn.h"

namespace trx {

constexpr int LANES = 32;        // decoys per group: one warp lane per decoy
constexpr int TILE = 16;         // residues per block row / block column
constexpr int REC_ELEMS = TILE * 9 * LANES;  // one partial-gradient record (N,CA,CB x xyz)
constexpr int MAXK = 64;         // max spline knots per restraint: distance tables (af2 variant: 60 listed + 2 end knots)
constexpr int MAXK_ANG = 32;     // ... and of the three angular types (28 / 28 / 16 listed + 2)
// The restraint kernel runs 4-warp CTAs: a step of the per-tile schedule is a matching of <= 4
// residue pairs, which fills ~91 % of the warp slots on protein-like contact maps (8-warp steps:
// 72 %, bounded by the busiest row/column of a tile).
constexpr int K1_WARPS = 4;
constexpr int K1_THREADS = K1_WARPS * 32;

void set_error(const char *fmt, ...);
void ctx_release(trx_ctx *ctx);   // drops one reference (context.cu)

#define TRX_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (call);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::trx::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return TRX_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define TRX_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::trx::set_error(__VA_ARGS__);     \
            return TRX_ERR_INVALID;            \
        }                                      \
    } while (0)

inline int padded_length(int L) { return (L + TILE - 1) / TILE * TILE; }
inline int num_groups(int N) { return (N + LANES - 1) / LANES; }

struct KernelTimer {
    double total_ms = 0;
    long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

}  // namespace trx

struct trx_ctx {
    // The creator's reference plus one per live tables / fold batch / dynamics state: those read the context when they
    // are destroyed, and a garbage collector (Python at interpreter exit) may destroy the context first.
    // trx_ctx_destroy drops the creator's reference; the context goes with the last one.
    std::atomic<int> refs{1};
    void retain() { refs.fetch_add(1); }
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int timing = 0;              // 0 off, 1 every kernel, 2 only the restraint kernel and the whole fold
    long long launches = 0;
    std::map<std::string, trx::KernelTimer> timers;
    // scratch device buffers reused across calls, keyed by a tag
    std::map<std::string, std::pair<void *, size_t>> scratch;

    // Device blocks released by destroyed tables / fold batches / dynamics states, kept for the next create on this
    // context: a dynamics loop rebuilds all three every iteration, and ~30 cudaMalloc + ~30 cudaFree (each a device-wide
    // synchronisation) cost more than the fold they bracket.  Every user of a block works on ctx->stream and every
    // destroy synchronises that stream before releasing, so a reused block has no reader left.
    std::mutex pool_mu;
    std::multimap<size_t, void *> pool_free;       // capacity -> released block
    std::unordered_map<void *, size_t> pool_cap;   // every block dev_alloc handed out -> its capacity
    size_t pool_bytes = 0;                         // sum over pool_free
    std::vector<void *> pinned_free;               // released PINNED_BLOCK-byte host blocks
    static constexpr size_t PINNED_BLOCK = 256;
    cudaError_t dev_alloc_bytes(void **p, size_t bytes);
    template <typename T>
    cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc_bytes((void **)p, bytes); }
    void dev_free(void *p);
    cudaError_t pinned_alloc(void **p);            // one PINNED_BLOCK-byte pinned host block
    void pinned_release(void *p);
    void pool_trim();                              // cudaFree everything the pool holds

    int get_scratch(const char *tag, size_t bytes, void **out);
    void time_begin(const char *name);
    void time_end(const char *name);
    void collect_timers();
};

namespace trx {

// One spline interval: S(x) = c0 + c1 u + c2 u^2 + c3 u^3 with u = x - x_k (16 B in fp32).
template <typename T>
struct alignas(4 * sizeof(T)) Coef {
    T c0, c1, c2, c3;
};

// Spline knot geometry of one restraint type, as the kernel keeps it in shared memory.
template <typename T>
struct KnotGeom {
    T x[MAXK];      // knot abscissae
    T gx0, ginv;    // interval guess: k = floor((x-gx0)*ginv)+goff
    int goff;
    int K;
    int urun0, urun1;   // intervals [urun0, urun1) are (nearly) uniformly spaced
};

// Work decomposition of the restraint kernel for a given number of decoy groups.
struct Plan {
    int groups = 0;
    int nwork = 0;     // CTAs per decoy group
    int nrec = 0;      // partial-gradient records per decoy group
    int *d_work = nullptr;      // [nwork][4]: block row I, first tile, tile count, row record id
    int *d_blk_ptr = nullptr;   // [nb+1] CSR over block -> records
    int *d_blk_rec = nullptr;
};

}  // namespace trx

struct trx_tables {
    trx_ctx *ctx = nullptr;
    int L = 0, Lpad = 0, nb = 0;
    int n[4] = {0, 0, 0, 0};
    int K[4] = {0, 0, 0, 0};
    double *d_tab64[4] = {nullptr, nullptr, nullptr, nullptr};  // [n][K] Coef<double>
    float *d_tab32[4] = {nullptr, nullptr, nullptr, nullptr};   // [n][K] Coef<float>
    double *d_y2[4] = {nullptr, nullptr, nullptr, nullptr};     // [n][K] fitted second derivatives (parity tests)
    trx::KnotGeom<double> *d_geom64 = nullptr;                  // [4]
    trx::KnotGeom<float> *d_geom32 = nullptr;                   // [4]
    int dist_ca = 0;                 // AtomPair restraints on CA instead of CB (af2 variant; no angular restraints then)
    int ntiles = 0;
    std::vector<int> tileI, tileJ;   // host copies, tiles sorted by (I, J)
    int *d_tileJ = nullptr;          // [ntiles]
    int *d_pairrec = nullptr;        // [ntiles][16][16][8]: mask, 6 restraint indices, pad
    unsigned short *d_sched = nullptr; // [steps][8 warps]: row | column<<4 of the pair, 0xffff = idle
    int *d_nsteps = nullptr;         // [ntiles+1] first step of each tile (CSR)
    int total_steps = 0;
    long long active_pairs = 0;
    std::map<int, trx::Plan> plans;  // keyed by number of decoy groups
    int get_plan(int groups, trx::Plan **out);
};

#ifdef __CUDACC__
namespace trx {
// mbarrier + bulk asynchronous copy (global -> shared, completion counted on an mbarrier; SASS UBLKCP): the TMA path
// for contiguous data, used by the L-BFGS sweeps (history ring) and the restraint kernel (pair records of a tile)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE_%=;\n"
        "bra LAB_WAIT_%=;\n"
        "LAB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace trx
#endif

namespace trx {
template <typename T>
int k1_launch(trx_ctx *ctx, trx_tables *tb, int Gtot, int g0, int ng, const T *d_xyz, int xstride, const float *wl,
              const double *w, const int *gactive, double *d_E, T *d_grad);
}
