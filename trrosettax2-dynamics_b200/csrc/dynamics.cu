// The outer "dynamics" loop's distogram update on device.
//
// Replaces, per iteration of run_inference.generate_npz_and_pdb (run_inference.py:97-139), the chain
//   PDB text -> Bio.PDB -> get_neighbors (utils_trX2dy/utils.py:125-182) -> pros (:185-249) ->
//   process_distribution_with_pred_distribution (:379-403) x 5 -> npz file
// by two kernels on arrays that stay on the device between iterations: the decoy's 6D geometry is binned
// (one thread per residue pair) and the four distograms plus the un-normalised `tmp` map are decayed,
// renormalised and Gaussian-smoothed (one thread per residue pair and map).
//
// The arithmetic restates the reference's numpy / scipy operations in their own precision and ORDER, so the maps
// are bit-identical to the host path (trx2dyn.dynamics, itself pinned to reference-run golden vectors):
//   * float32 row sums in numpy's pairwise order (8 running sums, combined ((0+1)+(2+3))+((4+5)+(6+7)), tail added
//     in sequence), float32 division;
//   * scipy.ndimage.gaussian_filter1d: float64, symmetric 9-tap correlation accumulated centre first then the
//     taps from the outside in, 'reflect' boundary (d c b a | a b c d | d c b a), result rounded to float32;
//   * the quirks: phi is binned from the THETA values (utils.py:226); the decay skips the last bin (:392).
// Every rounding-sensitive operation uses the _rn intrinsics (never contracted into FMAs).  The geometry itself
// (float64 atan2 / acos) only decides bins; a last-ulp difference to glibc matters only for a value exactly on a
// bin edge.
#include <cmath>
#include <cstring>
#include <vector>

#include "internal.cuh"

namespace trx {

constexpr int NB_D = 37, NB_A = 25, NB_P = 13, NB_MAX = 37;
constexpr float P_CONF = 0.5f, P_CUT = 0.05f, DECAY = 0.5f;   // params("0HD") of utils.py:331-332

struct d3 { double x, y, z; };
__device__ __forceinline__ d3 sub(d3 a, d3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ double dot3(d3 a, d3 b) { return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z)); }
__device__ __forceinline__ d3 cross3(d3 a, d3 b)
{
    return {__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
            __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}
__device__ __forceinline__ d3 scale(double s, d3 a) { return {__dmul_rn(s, a.x), __dmul_rn(s, a.y), __dmul_rn(s, a.z)}; }
__device__ __forceinline__ d3 divide(d3 a, double s) { return {__ddiv_rn(a.x, s), __ddiv_rn(a.y, s), __ddiv_rn(a.z, s)}; }
__device__ __forceinline__ d3 add(d3 a, d3 b) { return {__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)}; }
__device__ __forceinline__ double norm3(d3 a) { return __dsqrt_rn(dot3(a, a)); }

// get_dihedrals (utils.py:97-110)
__device__ __forceinline__ double dihedral_ref(d3 a, d3 b, d3 c, d3 d)
{
    const d3 b0 = sub(a, b), b2 = sub(d, c);
    d3 b1 = sub(c, b);
    b1 = divide(b1, norm3(b1));
    const d3 v = sub(b0, scale(dot3(b0, b1), b1)), w = sub(b2, scale(dot3(b2, b1), b1));
    return atan2(dot3(cross3(b1, v), w), dot3(v, w));
}
// get_angles (utils.py:113-122)
__device__ __forceinline__ double angle_ref(d3 a, d3 b, d3 c)
{
    d3 v = sub(a, b), w = sub(c, b);
    v = divide(v, norm3(v));
    w = divide(w, norm3(w));
    return acos(dot3(v, w));
}

// Virtual CB of every residue (utils.py:132-135), the file's CB where the residue is not Gly (:145-150).
__global__ void dyn_cb_kernel(int L, const double *__restrict__ n, const double *__restrict__ ca, const double *__restrict__ c,
                              const double *__restrict__ cb, const unsigned char *__restrict__ use_cb, double *__restrict__ vcb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const d3 N = {n[i * 3], n[i * 3 + 1], n[i * 3 + 2]}, CA = {ca[i * 3], ca[i * 3 + 1], ca[i * 3 + 2]}, C = {c[i * 3], c[i * 3 + 1], c[i * 3 + 2]};
    const d3 b = sub(CA, N), cc = sub(C, CA);
    d3 v = add(sub(add(scale(-0.58273431, cross3(b, cc)), scale(0.56802827, b)), scale(0.54067466, cc)), CA);
    if (use_cb[i]) v = {cb[i * 3], cb[i * 3 + 1], cb[i * 3 + 2]};
    vcb[i * 3] = v.x; vcb[i * 3 + 1] = v.y; vcb[i * 3 + 2] = v.z;
}

// get_neighbors + pros: the bin every residue pair realises in the four maps.
__global__ void dyn_bins_kernel(int L, const double *__restrict__ n, const double *__restrict__ ca, const double *__restrict__ vcb, double dmax,
                                int *__restrict__ jd, int *__restrict__ jo, int *__restrict__ jt, int *__restrict__ jp)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= L) return;
    auto P = [&](const double *p, int r) -> d3 { return {p[r * 3], p[r * 3 + 1], p[r * 3 + 2]}; };
    const d3 Bi = P(vcb, i), Bj = P(vcb, j);
    double dist = 0.0, omega = 0.0, theta = 0.0;
    const double d0 = norm3(sub(Bi, Bj));
    if (i != j && d0 <= dmax) {
        dist = norm3(sub(Bj, Bi));
        omega = dihedral_ref(P(ca, i), Bi, Bj, P(ca, j));
        theta = dihedral_ref(P(n, i), P(ca, i), Bi, Bj);
    }
    const double pi = 3.141592653589793, step = pi / 12;
    int kd = 0, ko = 0, kt = 0, kp = 0;
    for (int k = 0; k < 37; ++k) kd += __dadd_rn(2.0, __dmul_rn((double)k, 0.5)) < dist;       // np.arange(2, 20.5, 0.5)
    if (kd >= 37) kd = 0;
    for (int k = 0; k < 24; ++k) {
        const double e = __dadd_rn(-pi, __dmul_rn((double)k, step));                            // np.arange(-pi, pi, pi/12)
        ko += e < omega;
        kt += e < theta;
    }
    for (int k = 0; k < 12; ++k) kp += __dmul_rn((double)k, step) < theta;                     // np.arange(0, pi, pi/12), sic: theta (utils.py:226)
    const bool gone = kd == 0;
    const size_t o = (size_t)i * L + j;
    jd[o] = kd; jo[o] = gone ? 0 : ko; jt[o] = gone ? 0 : kt; jp[o] = gone ? 0 : kp;
}

// numpy's float32 pairwise sum of a contiguous row of n < 128 values
__device__ __forceinline__ float np_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float r = 0.f;
        for (int k = 0; k < n; ++k) r = __fadd_rn(r, a[k]);
        return r;
    }
    float r[8];
    for (int k = 0; k < 8; ++k) r[k] = a[k];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i]);
    return res;
}

// process_distribution_with_pred_distribution (utils.py:379-403) of one map, one thread per residue pair.
// in: the map before the update; bin: realised bins; out: the processed map (renormalised + smoothed) or, with
// norm == 0, the decayed un-normalised one (`tmp`).  w: the 9 gaussian taps (float64, as scipy builds them).
// maxchg (may be NULL): max |out - in| over the map (the loop's convergence signal on `tmp`).
template <int NB>
__global__ void dyn_process_kernel(int npair, const float *__restrict__ in, const int *__restrict__ bin, int norm, const double *__restrict__ w,
                                   float *__restrict__ out, unsigned int *__restrict__ maxchg)
{
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    float chg = 0.f;
    if (pidx < npair) {
        float row[NB];
        float mx = -INFINITY;
        for (int k = 0; k < NB; ++k) { row[k] = in[(size_t)pidx * NB + k]; mx = fmaxf(mx, row[k]); }
        if (mx < P_CONF) {
            const int kb = bin[pidx];
            if (kb <= NB - 2) {                       // the reference's slice is empty for the last bin
                const float v = row[kb];
                row[kb] = v < P_CUT ? v : __fmul_rn(v, DECAY);
            }
            if (norm) {
                const float sum = np_sum_f32(row, NB);
                double a[NB];
                for (int k = 0; k < NB; ++k) a[k] = (double)__fdiv_rn(row[k], sum);
                for (int l = 0; l < NB; ++l) {
                    double t = __dmul_rn(a[l], w[4]);
                    for (int jj = -4; jj < 0; ++jj) {
                        int lo = l + jj, hi = l - jj;
                        if (lo < 0) lo = -lo - 1;
                        if (hi >= NB) hi = 2 * NB - 1 - hi;
                        t = __dadd_rn(t, __dmul_rn(__dadd_rn(a[lo], a[hi]), w[jj + 4]));
                    }
                    row[l] = (float)t;
                }
            }
        }
        for (int k = 0; k < NB; ++k) {
            chg = fmaxf(chg, fabsf(row[k] - in[(size_t)pidx * NB + k]));
            out[(size_t)pidx * NB + k] = row[k];
        }
    }
    if (maxchg) {
        for (int o = 16; o > 0; o >>= 1) chg = fmaxf(chg, __shfl_xor_sync(0xffffffffu, chg, o));
        if ((threadIdx.x & 31) == 0) atomicMax(maxchg, __float_as_uint(chg));   // non-negative floats order like their bit patterns
    }
}

}  // namespace trx

using namespace trx;

struct trx_dyn {
    trx_ctx *ctx = nullptr;
    int L = 0, angle = 0;
    float *map[4] = {nullptr, nullptr, nullptr, nullptr};   // dist, omega, theta, phi (current maps)
    float *tmp = nullptr;                                   // un-normalised decayed dist map
    float *scratch = nullptr;                               // one map (largest)
    int *bins = nullptr;                                    // [4][L][L]
    double *xyz = nullptr;                                  // n, ca, c, cb, vcb: [5][L][3]
    unsigned char *use_cb = nullptr;
    double *w = nullptr;                                    // 9 taps
    unsigned int *maxchg = nullptr;
};

extern "C" {

int trx_dyn_destroy(trx_dyn *d)
{
    if (!d) return TRX_OK;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    for (int t = 0; t < 4; ++t) d->ctx->dev_free(d->map[t]);
    d->ctx->dev_free(d->tmp);
    d->ctx->dev_free(d->scratch);
    d->ctx->dev_free(d->bins);
    d->ctx->dev_free(d->xyz);
    d->ctx->dev_free(d->use_cb);
    d->ctx->dev_free(d->w);
    d->ctx->dev_free(d->maxchg);
    trx_ctx *ctx = d->ctx;
    delete d;
    ctx_release(ctx);
    return TRX_OK;
}

int trx_dyn_create(trx_ctx *ctx, int L, const float *dist, const float *omega, const float *theta, const float *phi, trx_dyn **out)
{
    TRX_REQUIRE(ctx && dist && out && L > 0, "trx_dyn_create: bad argument");
    TRX_REQUIRE((omega && theta && phi) || (!omega && !theta && !phi), "trx_dyn_create: give all three orientation maps or none");
    TRX_CUDA(cudaSetDevice(ctx->device));
    trx_dyn *d = new trx_dyn();
    d->ctx = ctx; d->L = L; d->angle = omega ? 1 : 0;
    ctx->retain();   // given back by trx_dyn_destroy (the failure path below goes through it too)
    const size_t np = (size_t)L * L;
    const int nb[4] = {NB_D, NB_A, NB_A, NB_P};
    const float *src[4] = {dist, omega, theta, phi};
    cudaError_t e = cudaSuccess;
    for (int t = 0; t < 4 && e == cudaSuccess; ++t)
        if (src[t]) {
            e = ctx->dev_alloc(&d->map[t], np * nb[t] * sizeof(float));
            if (e == cudaSuccess) e = cudaMemcpyAsync(d->map[t], src[t], np * nb[t] * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);
        }
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->tmp, np * NB_D * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d->tmp, dist, np * NB_D * sizeof(float), cudaMemcpyHostToDevice, ctx->stream);   // 'tmp' starts as dist (run_inference.py:116-133)
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->scratch, np * NB_MAX * sizeof(float));
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->bins, 4 * np * sizeof(int));
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->xyz, (size_t)5 * L * 3 * sizeof(double));
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->use_cb, L);
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->w, 9 * sizeof(double));
    if (e == cudaSuccess) e = ctx->dev_alloc(&d->maxchg, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        set_error("trx_dyn_create: %s", cudaGetErrorString(e));
        trx_dyn_destroy(d);
        return TRX_ERR_CUDA;
    }
    *out = d;
    return TRX_OK;
}

/* One iteration of the outer loop (get_npz_from_pred_pdb twice, run_inference.py:116-133): the decoy's backbone
 * n, ca, c, cb ([L][3] double; cb used where use_cb[i] != 0, the virtual CB elsewhere) updates the four maps and
 * `tmp` in place.  w9: the gaussian taps as scipy.ndimage builds them for the chosen sigma (truncate 4, radius 4).
 * *max_tmp_change: max |tmp_new - tmp_old| (the reference stops below 0.01, run_inference.py:135). */
int trx_dyn_step(trx_dyn *d, const double *n, const double *ca, const double *c, const double *cb, const unsigned char *use_cb,
                 const double *w9, double *max_tmp_change)
{
    TRX_REQUIRE(d && n && ca && c && cb && use_cb && w9, "trx_dyn_step: NULL argument");
    trx_ctx *ctx = d->ctx;
    TRX_CUDA(cudaSetDevice(ctx->device));
    const int L = d->L;
    const size_t v = (size_t)L * 3, np = (size_t)L * L;
    TRX_CUDA(cudaMemcpyAsync(d->xyz, n, v * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d->xyz + v, ca, v * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d->xyz + 2 * v, c, v * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d->xyz + 3 * v, cb, v * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d->use_cb, use_cb, L, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d->w, w9, 9 * 8, cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemsetAsync(d->maxchg, 0, sizeof(unsigned int), ctx->stream));
    ctx->time_begin("dynamics");
    dyn_cb_kernel<<<(L + 127) / 128, 128, 0, ctx->stream>>>(L, d->xyz, d->xyz + v, d->xyz + 2 * v, d->xyz + 3 * v, d->use_cb, d->xyz + 4 * v);
    ctx->time_end("dynamics");
    int *jd = d->bins, *jo = jd + np, *jt = jo + np, *jp = jt + np;
    ctx->time_begin("dynamics");
    dyn_bins_kernel<<<dim3((L + 127) / 128, L), 128, 0, ctx->stream>>>(L, d->xyz, d->xyz + v, d->xyz + 4 * v, 20.0, jd, jo, jt, jp);
    ctx->time_end("dynamics");
    const int grid = (int)((np + 127) / 128);
    auto swap_in = [&](float *&m, size_t bytes) -> int {   // the processed map (in scratch) becomes the current one
        TRX_CUDA(cudaMemcpyAsync(m, d->scratch, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return TRX_OK;
    };
    int rc;
    ctx->time_begin("dynamics");
    dyn_process_kernel<NB_D><<<grid, 128, 0, ctx->stream>>>((int)np, d->tmp, jd, 0, d->w, d->scratch, d->maxchg);
    ctx->time_end("dynamics");
    if ((rc = swap_in(d->tmp, np * NB_D * 4))) return rc;
    ctx->time_begin("dynamics");
    dyn_process_kernel<NB_D><<<grid, 128, 0, ctx->stream>>>((int)np, d->map[0], jd, 1, d->w, d->scratch, nullptr);
    ctx->time_end("dynamics");
    if ((rc = swap_in(d->map[0], np * NB_D * 4))) return rc;
    if (d->angle) {
        ctx->time_begin("dynamics");
        dyn_process_kernel<NB_A><<<grid, 128, 0, ctx->stream>>>((int)np, d->map[1], jo, 1, d->w, d->scratch, nullptr);
        ctx->time_end("dynamics");
        if ((rc = swap_in(d->map[1], np * NB_A * 4))) return rc;
        ctx->time_begin("dynamics");
        dyn_process_kernel<NB_A><<<grid, 128, 0, ctx->stream>>>((int)np, d->map[2], jt, 1, d->w, d->scratch, nullptr);
        ctx->time_end("dynamics");
        if ((rc = swap_in(d->map[2], np * NB_A * 4))) return rc;
        ctx->time_begin("dynamics");
        dyn_process_kernel<NB_P><<<grid, 128, 0, ctx->stream>>>((int)np, d->map[3], jp, 1, d->w, d->scratch, nullptr);
        ctx->time_end("dynamics");
        if ((rc = swap_in(d->map[3], np * NB_P * 4))) return rc;
    }
    unsigned int bits = 0;
    TRX_CUDA(cudaMemcpyAsync(&bits, d->maxchg, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    TRX_CUDA(cudaGetLastError());
    if (max_tmp_change) {
        float f;
        memcpy(&f, &bits, sizeof(f));
        *max_tmp_change = (double)f;
    }
    return TRX_OK;
}

/* Current maps to the host (any pointer may be NULL): dist [L][L][37], omega / theta [L][L][25], phi [L][L][13],
 * tmp [L][L][37]; bins (may be NULL): [4][L][L] realised bins of the last step (dist, omega, theta, phi). */
int trx_dyn_get(trx_dyn *d, float *dist, float *omega, float *theta, float *phi, float *tmp, int *bins)
{
    TRX_REQUIRE(d, "trx_dyn_get: NULL handle");
    trx_ctx *ctx = d->ctx;
    TRX_CUDA(cudaSetDevice(ctx->device));
    const size_t np = (size_t)d->L * d->L;
    float *dst[4] = {dist, omega, theta, phi};
    const int nb[4] = {NB_D, NB_A, NB_A, NB_P};
    for (int t = 0; t < 4; ++t)
        if (dst[t]) {
            TRX_REQUIRE(d->map[t], "trx_dyn_get: the state holds no orientation maps");
            TRX_CUDA(cudaMemcpyAsync(dst[t], d->map[t], np * nb[t] * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
    if (tmp) TRX_CUDA(cudaMemcpyAsync(tmp, d->tmp, np * NB_D * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (bins) TRX_CUDA(cudaMemcpyAsync(bins, d->bins, 4 * np * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRX_OK;
}

}  // extern "C"
