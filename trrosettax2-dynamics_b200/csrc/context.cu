// Context, error reporting, per-kernel timing and the natural <-> grouped layout kernels.
#include <algorithm>
#include <cstdarg>
#include <cstring>

#include "internal.cuh"

namespace trx {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// natural [n][res][A][3] -> grouped [n/32][res_pad][A*3][32]; one thread per (decoy, residue)
// reads its A*3 contiguous values; the transposed store is the coalesced side.
template <typename T>
__global__ void to_grouped_kernel(int N, int L, int Lpad, int A3, const T *__restrict__ nat, T *__restrict__ grp)
{
    int lane = threadIdx.x;
    int res = blockIdx.x * blockDim.y + threadIdx.y;
    int g = blockIdx.y;
    int n = g * LANES + lane;
    if (res >= Lpad) return;
    T *dst = grp + ((size_t)g * Lpad + res) * A3 * LANES + lane;
    if (n < N && res < L) {
        const T *src = nat + ((size_t)n * L + res) * A3;
        for (int c = 0; c < A3; ++c) dst[c * LANES] = src[c];
    } else {
        // padding decoys replicate nothing: spread padded points so no geometry degenerates
        for (int c = 0; c < A3; ++c) dst[c * LANES] = (T)(0.37 * (c + 1) + 1.3 * (res % 7) + 0.11 * lane);
    }
}

template <typename T>
__global__ void from_grouped_kernel(int N, int L, int Lpad, int A3, const T *__restrict__ grp, T *__restrict__ nat)
{
    int lane = threadIdx.x;
    int res = blockIdx.x * blockDim.y + threadIdx.y;
    int g = blockIdx.y;
    int n = g * LANES + lane;
    if (res >= L || n >= N) return;
    const T *src = grp + ((size_t)g * Lpad + res) * A3 * LANES + lane;
    T *dst = nat + ((size_t)n * L + res) * A3;
    for (int c = 0; c < A3; ++c) dst[c] = src[c * LANES];
}

}  // namespace trx

namespace trx {
void ctx_release(trx_ctx *ctx)
{
    if (ctx->refs.fetch_sub(1) != 1) return;   // tables / fold batches / dynamics states of this context are still alive
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->collect_timers();
    for (auto &kv : ctx->scratch)
        if (kv.second.first) cudaFree(kv.second.first);
    ctx->pool_trim();
    for (void *h : ctx->pinned_free) cudaFreeHost(h);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
}  // namespace trx

// ---- the per-context block pool (see internal.cuh)
static constexpr size_t POOL_BLOCK_MAX = (size_t)512 << 20;   // larger blocks go straight back to the driver
static constexpr size_t POOL_TOTAL_MAX = (size_t)4 << 30;

cudaError_t trx_ctx::dev_alloc_bytes(void **p, size_t bytes)
{
    size_t want = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        auto it = pool_free.lower_bound(want);
        if (it != pool_free.end() && it->first <= 2 * want + (64 << 10)) {
            *p = it->second;
            pool_bytes -= it->first;
            pool_cap[*p] = it->first;
            pool_free.erase(it);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {   // give the pooled blocks back and try once more
        (void)cudaGetLastError();
        pool_trim();
        e = cudaMalloc(p, want);
    }
    if (e != cudaSuccess) {
        *p = nullptr;
        return e;
    }
    std::lock_guard<std::mutex> lk(pool_mu);
    pool_cap[*p] = want;
    return cudaSuccess;
}

void trx_ctx::dev_free(void *p)
{
    if (!p) return;
    size_t cap = 0;
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        auto it = pool_cap.find(p);
        if (it != pool_cap.end()) {
            cap = it->second;
            pool_cap.erase(it);
            if (cap <= POOL_BLOCK_MAX && pool_bytes + cap <= POOL_TOTAL_MAX) {
                pool_free.emplace(cap, p);
                pool_bytes += cap;
                return;
            }
        }
    }
    cudaFree(p);
}

void trx_ctx::pool_trim()
{
    std::lock_guard<std::mutex> lk(pool_mu);
    for (auto &kv : pool_free) cudaFree(kv.second);
    pool_free.clear();
    pool_bytes = 0;
}

cudaError_t trx_ctx::pinned_alloc(void **p)
{
    {
        std::lock_guard<std::mutex> lk(pool_mu);
        if (!pinned_free.empty()) {
            *p = pinned_free.back();
            pinned_free.pop_back();
            return cudaSuccess;
        }
    }
    return cudaMallocHost(p, PINNED_BLOCK);
}

void trx_ctx::pinned_release(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(pool_mu);
    pinned_free.push_back(p);
}

int trx_ctx::get_scratch(const char *tag, size_t bytes, void **out)
{
    auto &s = scratch[tag];
    if (s.second < bytes) {
        if (s.first) {
            TRX_CUDA(cudaStreamSynchronize(stream));
            TRX_CUDA(cudaFree(s.first));
            s.first = nullptr;
            s.second = 0;
        }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&s.first, want);
        if (e != cudaSuccess) {
            trx::set_error("cudaMalloc(%zu bytes) for '%s' failed: %s", want, tag, cudaGetErrorString(e));
            s.first = nullptr;
            return TRX_ERR_NOMEM;
        }
        s.second = want;
    }
    *out = s.first;
    return TRX_OK;
}

// timing == 2: only the dominant kernel and the whole fold are bracketed with events (two event records per launch are
// not free when a round is a dozen launches of a few microseconds each)
static inline bool timed(int mode, const char *name)
{
    return mode == 1 || (mode == 2 && (!strcmp(name, "restraints") || !strcmp(name, "fold_device")));
}

void trx_ctx::time_begin(const char *name)
{
    ++launches;
    if (!timed(timing, name)) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, stream);
    timers[name].pending.emplace_back(a, b);
}

void trx_ctx::time_end(const char *name)
{
    if (!timed(timing, name)) return;
    auto &t = timers[name];
    cudaEventRecord(t.pending.back().second, stream);
}

void trx_ctx::collect_timers()
{
    for (auto &kv : timers) {
        for (auto &p : kv.second.pending) {
            cudaEventSynchronize(p.second);
            float ms = 0;
            cudaEventElapsedTime(&ms, p.first, p.second);
            kv.second.total_ms += ms;
            kv.second.launches += 1;
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
        kv.second.pending.clear();
    }
}

extern "C" {

int trx_abi_version(void) { return TRX_ABI_VERSION; }
const char *trx_last_error(void) { return trx::g_err; }

int trx_ctx_create(int device, void *stream, trx_ctx **out)
{
    TRX_REQUIRE(out != nullptr, "trx_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        trx::set_error("trx_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return TRX_ERR_CUDA;
    }
    TRX_REQUIRE(device >= 0 && device < count, "trx_ctx_create: device %d out of range [0,%d)", device, count);
    TRX_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TRX_CUDA(cudaGetDeviceProperties(&prop, device));
    TRX_REQUIRE(prop.major == 10, "trx_ctx_create: built for sm_100a only, device is sm_%d%d", prop.major, prop.minor);
    trx_ctx *c = new trx_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        TRX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    *out = c;
    return TRX_OK;
}

int trx_ctx_destroy(trx_ctx *ctx)
{
    if (!ctx) return TRX_OK;
    trx::ctx_release(ctx);
    return TRX_OK;
}

int trx_ctx_sync(trx_ctx *ctx)
{
    TRX_REQUIRE(ctx, "trx_ctx_sync: ctx is NULL");
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRX_OK;
}

int trx_ctx_set_timing(trx_ctx *ctx, int enabled)
{
    TRX_REQUIRE(ctx, "trx_ctx_set_timing: ctx is NULL");
    ctx->timing = enabled < 0 ? 0 : (enabled > 2 ? 1 : enabled);
    return TRX_OK;
}

int trx_ctx_get_timing(trx_ctx *ctx, const char *name, double *total_ms, long long *launches)
{
    TRX_REQUIRE(ctx && name, "trx_ctx_get_timing: NULL argument");
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->collect_timers();
    auto it = ctx->timers.find(name);
    if (total_ms) *total_ms = it == ctx->timers.end() ? 0.0 : it->second.total_ms;
    if (launches) *launches = it == ctx->timers.end() ? 0 : it->second.launches;
    return TRX_OK;
}

int trx_ctx_reset_timing(trx_ctx *ctx)
{
    TRX_REQUIRE(ctx, "trx_ctx_reset_timing: ctx is NULL");
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->collect_timers();
    ctx->timers.clear();
    return TRX_OK;
}

long long trx_ctx_launch_count(trx_ctx *ctx) { return ctx ? ctx->launches : 0; }

int trx_padded_length(int L) { return trx::padded_length(L); }

int trx_to_grouped(trx_ctx *ctx, int N, int L, int n_atoms, int precision, const void *d_nat, void *d_grp)
{
    TRX_REQUIRE(ctx && d_nat && d_grp && N > 0 && L > 0 && n_atoms > 0, "trx_to_grouped: bad argument");
    TRX_REQUIRE(precision == TRX_F64 || precision == TRX_F32, "trx_to_grouped: precision must be 64 or 32");
    int Lpad = trx::padded_length(L), G = trx::num_groups(N);
    dim3 block(32, 8), grid((Lpad + 7) / 8, G);
    ctx->time_begin("layout");
    if (precision == TRX_F64)
        trx::to_grouped_kernel<double><<<grid, block, 0, ctx->stream>>>(N, L, Lpad, n_atoms * 3, (const double *)d_nat, (double *)d_grp);
    else
        trx::to_grouped_kernel<float><<<grid, block, 0, ctx->stream>>>(N, L, Lpad, n_atoms * 3, (const float *)d_nat, (float *)d_grp);
    ctx->time_end("layout");
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

int trx_from_grouped(trx_ctx *ctx, int N, int L, int n_atoms, int precision, const void *d_grp, void *d_nat)
{
    TRX_REQUIRE(ctx && d_nat && d_grp && N > 0 && L > 0 && n_atoms > 0, "trx_from_grouped: bad argument");
    TRX_REQUIRE(precision == TRX_F64 || precision == TRX_F32, "trx_from_grouped: precision must be 64 or 32");
    int Lpad = trx::padded_length(L), G = trx::num_groups(N);
    dim3 block(32, 8), grid((Lpad + 7) / 8, G);
    ctx->time_begin("layout");
    if (precision == TRX_F64)
        trx::from_grouped_kernel<double><<<grid, block, 0, ctx->stream>>>(N, L, Lpad, n_atoms * 3, (const double *)d_grp, (double *)d_nat);
    else
        trx::from_grouped_kernel<float><<<grid, block, 0, ctx->stream>>>(N, L, Lpad, n_atoms * 3, (const float *)d_grp, (float *)d_nat);
    ctx->time_end("layout");
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

}  // extern "C"
