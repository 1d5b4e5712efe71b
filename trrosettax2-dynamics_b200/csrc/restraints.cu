// K1: batched restraint energy + analytic gradient for all four restraint types.
//
// Replaces, for N decoys at once, what PyRosetta does per decoy inside
// ScoreFunction::score / the derivative pass of MinMover for the score terms
// atom_pair_constraint, dihedral_constraint, angle_constraint
// (reference call sites folding/folding.py:74-104,164-171).
//
// Mapping (B200): lane = decoy (32 decoys per group, coordinates stored
// [group][residue][9][32] so every load/store is one full 128 B / 256 B line and
// all 32 lanes walk the SAME restraint => the spline table of a restraint is read
// once per warp from L1/L2, never per decoy from HBM).  A CTA of 8 warps owns a
// block row of 16 residues for one decoy group and walks its active 16x16 tiles;
// warp w owns rows w and w+8 (row gradients live in registers), column gradients
// are accumulated in shared memory with a staggered column schedule
// (warp w touches column (2w+s)&15 at step s) so that no two warps ever touch the
// same column in a step: no atomics, bit-reproducible sums.  Per-tile column
// gradients and per-CTA row gradients are written as partial records that the
// reduce kernel sums in a fixed order.
#include "internal.cuh"

namespace trx {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

__device__ __forceinline__ float t_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double t_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float t_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double t_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ float t_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double t_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float t_atan2(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double t_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float t_acos(float x) { return acosf(x); }
__device__ __forceinline__ double t_acos(double x) { return acos(x); }
__device__ __forceinline__ float t_floor(float x) { return floorf(x); }
__device__ __forceinline__ double t_floor(double x) { return floor(x); }
template <typename T> __device__ __forceinline__ T t_tiny();
template <> __device__ __forceinline__ float t_tiny<float>() { return 1e-12f; }
template <> __device__ __forceinline__ double t_tiny<double>() { return 1e-24; }

template <typename T>
struct K1Params {
    const T *X;                                  // [G][Lpad][9][32]
    const typename Vec2<T>::type *tab[4];        // [n][K] (y, y'')
    const KnotGeom<T> *geom;                     // [4]
    const int *pairrec;                          // [ntiles][16][16][8]
    const int *tileJ;                            // [ntiles]
    const int *work;                             // [nwork][4]
    T *recs;                                     // [G][nrec][REC_ELEMS]
    double *Epart;                               // [G][nwork][3][32]
    int Lpad, nrec, nwork;
    int xstride;                                 // values per residue in X (9: N,CA,CB; 15: fold layout)
    int g0;                                      // first decoy group of this launch
    const int *gactive;                          // per-group flag or NULL (all active)
    const float *wl;                             // per-decoy weights [3][Npad] or NULL (use w0..w2)
    int Npad;
    T w0, w1, w2;
};

// Rosetta SplineFunc (weight 1): cubic inside [x_0, x_{K-1}], flat outside.
// NR splint on the interval found from a uniform-grid guess corrected against the
// true knots (the reference's %.3f rounding makes the angular grids slightly uneven).
template <typename T>
__device__ __forceinline__ void spline_eval(const KnotGeom<T> &kn, const typename Vec2<T>::type *__restrict__ tab,
                                            T x, T &f, T &df)
{
    const int K = kn.K;
    if (x < kn.x[0]) { f = tab[0].x; df = (T)0; return; }
    if (x > kn.x[K - 1]) { f = tab[K - 1].x; df = (T)0; return; }
    int k = (int)t_floor((x - kn.gx0) * kn.ginv) + kn.goff;
    k = min(max(k, 0), K - 2);
    while (k > 0 && x < kn.x[k]) --k;
    while (k < K - 2 && x >= kn.x[k + 1]) ++k;
    const typename Vec2<T>::type lo = tab[k], hi = tab[k + 1];
    const T rh = kn.rh[k];
    const T a = (kn.x[k + 1] - x) * rh;
    const T b = (x - kn.x[k]) * rh;
    f = a * lo.x + b * hi.x + ((a * a * a - a) * lo.y + (b * b * b - b) * hi.y) * kn.h2_6[k];
    df = (hi.x - lo.x) * rh - (((T)3 * a * a - (T)1) * lo.y - ((T)3 * b * b - (T)1) * hi.y) * kn.h_6[k];
}

#define CROSS(o, a, b)                       \
    o##x = a##y * b##z - a##z * b##y;        \
    o##y = a##z * b##x - a##x * b##z;        \
    o##z = a##x * b##y - a##y * b##x;
#define DOT(a, b) (a##x * b##x + a##y * b##y + a##z * b##z)

// Dihedral p1-p2-p3-p4 from F = p1-p2, G = p2-p3, H = p4-p3 (IUPAC sign, equal to the
// reference's numpy get_dihedrals, utils_trX2dy/utils.py:97-110), spline energy, and
// gradient (Blondel & Karplus 1996) accumulated into g1..g4 (arrays of 3).
template <typename T>
__device__ __forceinline__ void dihedral_term(const KnotGeom<T> &kn, const typename Vec2<T>::type *__restrict__ tab,
                                              T Fx, T Fy, T Fz, T Gx, T Gy, T Gz, T Hx, T Hy, T Hz, T w,
                                              double &e, T *g1, T *g2, T *g3, T *g4)
{
    T Ax, Ay, Az, Bx, By, Bz;
    CROSS(A, F, G);
    CROSS(B, H, G);
    const T G2 = max(DOT(G, G), t_tiny<T>());
    const T rG = t_rsqrt(G2);
    const T Gn = G2 * rG;
    const T phi = t_atan2(-Gn * DOT(F, B), DOT(A, B));
    T f, df;
    spline_eval(kn, tab, phi, f, df);
    e += (double)f;
    if (df != (T)0) {
        const T s = w * df;
        const T iA2 = t_rcp(max(DOT(A, A), t_tiny<T>()));
        const T iB2 = t_rcp(max(DOT(B, B), t_tiny<T>()));
        const T c1 = -s * Gn * iA2, c4 = s * Gn * iB2;
        const T tA = s * DOT(F, G) * iA2 * rG, tB = s * DOT(H, G) * iB2 * rG;
        const T u1x = c1 * Ax, u1y = c1 * Ay, u1z = c1 * Az;
        const T u4x = c4 * Bx, u4y = c4 * By, u4z = c4 * Bz;
        const T tx = tA * Ax - tB * Bx, ty = tA * Ay - tB * By, tz = tA * Az - tB * Bz;
        g1[0] += u1x; g1[1] += u1y; g1[2] += u1z;
        g4[0] += u4x; g4[1] += u4y; g4[2] += u4z;
        g2[0] += tx - u1x; g2[1] += ty - u1y; g2[2] += tz - u1z;
        g3[0] -= tx + u4x; g3[1] -= ty + u4y; g3[2] -= tz + u4z;
    }
}

// Angle p1-p2-p3 at vertex p2 from U = p1-p2, V = p3-p2 (reference get_angles,
// utils_trX2dy/utils.py:113-122), spline energy and gradient into g1,g2,g3.
template <typename T>
__device__ __forceinline__ void angle_term(const KnotGeom<T> &kn, const typename Vec2<T>::type *__restrict__ tab,
                                           T Ux, T Uy, T Uz, T Vx, T Vy, T Vz, T w, double &e, T *g1, T *g2, T *g3)
{
    const T rU = t_rsqrt(max(DOT(U, U), t_tiny<T>()));
    const T rV = t_rsqrt(max(DOT(V, V), t_tiny<T>()));
    T c = DOT(U, V) * rU * rV;
    c = min(max(c, (T)-1), (T)1);
    const T ang = t_acos(c);
    T f, df;
    spline_eval(kn, tab, ang, f, df);
    e += (double)f;
    if (df != (T)0) {
        const T sn = t_sqrt(max((T)1 - c * c, t_tiny<T>()));
        const T s = -w * df * t_rcp(sn);
        const T ux = Ux * rU, uy = Uy * rU, uz = Uz * rU;
        const T vx = Vx * rV, vy = Vy * rV, vz = Vz * rV;
        const T a1 = s * rU, a3 = s * rV;
        const T p1x = a1 * (vx - c * ux), p1y = a1 * (vy - c * uy), p1z = a1 * (vz - c * uz);
        const T p3x = a3 * (ux - c * vx), p3y = a3 * (uy - c * vy), p3z = a3 * (uz - c * vz);
        g1[0] += p1x; g1[1] += p1y; g1[2] += p1z;
        g3[0] += p3x; g3[1] += p3y; g3[2] += p3z;
        g2[0] -= p1x + p3x; g2[1] -= p1y + p3y; g2[2] -= p1z + p3z;
    }
}

// All restraints of the unordered residue pair (i = row, j = column, i < j).
// ri/cj: coordinates N(0..2) CA(3..5) CB(6..8); rg/cg: gradient accumulators.
template <typename T>
__device__ __forceinline__ void pair_eval(const K1Params<T> &p, const KnotGeom<T> *geom, const int4 ia, const int4 ib,
                                          const T *ri, const T *cj, T *rg, T *cg, const T w0, const T w1, const T w2,
                                          double &e0, double &e1, double &e2)
{
    using T2 = typename Vec2<T>::type;
    const int mask = ia.x;
    const T Dx = cj[6] - ri[6], Dy = cj[7] - ri[7], Dz = cj[8] - ri[8];        // CB_j - CB_i
    const T Px = ri[3] - ri[6], Py = ri[4] - ri[7], Pz = ri[5] - ri[8];        // CA_i - CB_i
    const T Qx = cj[3] - cj[6], Qy = cj[4] - cj[7], Qz = cj[5] - cj[8];        // CA_j - CB_j
    if (mask & 1) {  // dist: AtomPair CB_i CB_j
        const T d2 = max(DOT(D, D), t_tiny<T>());
        const T rd = t_rsqrt(d2);
        T f, df;
        spline_eval(geom[0], p.tab[0] + (size_t)ia.y * geom[0].K, d2 * rd, f, df);
        e0 += (double)f;
        if (df != (T)0) {
            const T s = w0 * df * rd;
            cg[6] += s * Dx; cg[7] += s * Dy; cg[8] += s * Dz;
            rg[6] -= s * Dx; rg[7] -= s * Dy; rg[8] -= s * Dz;
        }
    }
    if (mask & 2)    // omega: Dihedral CA_i CB_i CB_j CA_j
        dihedral_term<T>(geom[1], p.tab[1] + (size_t)ia.z * geom[1].K, Px, Py, Pz, -Dx, -Dy, -Dz, Qx, Qy, Qz, w1, e1,
                         rg + 3, rg + 6, cg + 6, cg + 3);
    if (mask & 4)    // theta(i,j): Dihedral N_i CA_i CB_i CB_j
        dihedral_term<T>(geom[2], p.tab[2] + (size_t)ia.w * geom[2].K, ri[0] - ri[3], ri[1] - ri[4], ri[2] - ri[5], Px, Py, Pz,
                         Dx, Dy, Dz, w1, e1, rg + 0, rg + 3, rg + 6, cg + 6);
    if (mask & 8)    // theta(j,i): Dihedral N_j CA_j CB_j CB_i
        dihedral_term<T>(geom[2], p.tab[2] + (size_t)ib.x * geom[2].K, cj[0] - cj[3], cj[1] - cj[4], cj[2] - cj[5], Qx, Qy, Qz,
                         -Dx, -Dy, -Dz, w1, e1, cg + 0, cg + 3, cg + 6, rg + 6);
    if (mask & 16)   // phi(i,j): Angle CA_i CB_i CB_j
        angle_term<T>(geom[3], p.tab[3] + (size_t)ib.y * geom[3].K, Px, Py, Pz, Dx, Dy, Dz, w2, e2, rg + 3, rg + 6, cg + 6);
    if (mask & 32)   // phi(j,i): Angle CA_j CB_j CB_i
        angle_term<T>(geom[3], p.tab[3] + (size_t)ib.z * geom[3].K, Qx, Qy, Qz, -Dx, -Dy, -Dz, w2, e2, cg + 3, cg + 6, rg + 6);
    (void)sizeof(T2);
}

template <typename T>
__global__ void __launch_bounds__(K1_THREADS, (sizeof(T) == 4 ? 2 : 1)) restraints_kernel(const K1Params<T> p)
{
    __shared__ KnotGeom<T> geom[4];
    __shared__ __align__(16) T colg[REC_ELEMS];
    double(*ered)[3][LANES] = reinterpret_cast<double(*)[3][LANES]>(colg);  // reused after the last flush
    static_assert(sizeof(double) * K1_WARPS * 3 * LANES <= sizeof(T) * REC_ELEMS, "energy scratch must fit");
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.y, g = p.g0 + blockIdx.x;   // group is the fast grid index: co-resident CTAs share tiles
    if (p.gactive && !p.gactive[g]) return;
    {
        const int *src = reinterpret_cast<const int *>(p.geom);
        int *dst = reinterpret_cast<int *>(geom);
        for (int e = threadIdx.x; e < (int)(sizeof(geom) / sizeof(int)); e += K1_THREADS) dst[e] = src[e];
        for (int e = threadIdx.x; e < REC_ELEMS; e += K1_THREADS) colg[e] = (T)0;
    }
    const int I = p.work[q * 4 + 0], t0 = p.work[q * 4 + 1], nt = p.work[q * 4 + 2], rowrec = p.work[q * 4 + 3];
    const int xs = p.xstride;
    const T *__restrict__ Xg = p.X + (size_t)g * p.Lpad * xs * LANES + lane;
    T w0 = p.w0, w1 = p.w1, w2 = p.w2;
    if (p.wl) {
        w0 = (T)p.wl[0 * (size_t)p.Npad + g * LANES + lane];
        w1 = (T)p.wl[1 * (size_t)p.Npad + g * LANES + lane];
        w2 = (T)p.wl[2 * (size_t)p.Npad + g * LANES + lane];
    }
    T ri[2][9], rg[2][9];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int res = I * TILE + w + K1_WARPS * r;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            ri[r][c] = Xg[((size_t)res * xs + c) * LANES];
            rg[r][c] = (T)0;
        }
    }
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    __syncthreads();

    for (int t = t0; t < t0 + nt; ++t) {
        const int J = p.tileJ[t];
        const int *__restrict__ rec_t = p.pairrec + (size_t)t * TILE * TILE * 8;
        for (int s = 0; s < TILE; ++s) {
            const int c = (2 * w + s) & (TILE - 1);
            const int4 *r0 = reinterpret_cast<const int4 *>(rec_t + (w * TILE + c) * 8);
            const int4 *r1 = reinterpret_cast<const int4 *>(rec_t + ((w + K1_WARPS) * TILE + c) * 8);
            const int4 a0 = __ldg(r0), a1 = __ldg(r1);
            if (a0.x | a1.x) {
                T cj[9], cg[9];
                const int res = J * TILE + c;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    cj[k] = Xg[((size_t)res * xs + k) * LANES];
                    cg[k] = (T)0;
                }
                if (a0.x) pair_eval<T>(p, geom, a0, __ldg(r0 + 1), ri[0], cj, rg[0], cg, w0, w1, w2, e0, e1, e2);
                if (a1.x) pair_eval<T>(p, geom, a1, __ldg(r1 + 1), ri[1], cj, rg[1], cg, w0, w1, w2, e0, e1, e2);
#pragma unroll
                for (int k = 0; k < 9; ++k) colg[(c * 9 + k) * LANES + lane] += cg[k];
            }
            __syncthreads();
        }
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + t) * REC_ELEMS;
        for (int e = threadIdx.x; e < REC_ELEMS; e += K1_THREADS) {
            dst[e] = colg[e];
            colg[e] = (T)0;
        }
        __syncthreads();
    }
    {
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + rowrec) * REC_ELEMS;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int k = 0; k < 9; ++k) dst[((w + K1_WARPS * r) * 9 + k) * LANES + lane] = rg[r][k];
    }
    ered[w][0][lane] = e0;
    ered[w][1][lane] = e1;
    ered[w][2][lane] = e2;
    __syncthreads();
    if (w < 3) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < K1_WARPS; ++k) s += ered[k][w][lane];
        p.Epart[(((size_t)g * p.nwork + q) * 3 + w) * LANES + lane] = s;
    }
}

// Sums the partial records of every residue block in a fixed order (deterministic), and
// the per-CTA energies.  grid = (nb + 1, G): x < nb -> gradient block, x == nb -> energies.
template <typename T>
__global__ void __launch_bounds__(256) reduce_kernel(const T *__restrict__ recs, int nrec, const int *__restrict__ blk_ptr,
                                                     const int *__restrict__ blk_rec, T *__restrict__ grad, int Lpad, int nb,
                                                     int accumulate, const double *__restrict__ Epart, int nwork,
                                                     double *__restrict__ E, int Npad, int g0, const int *__restrict__ gactive)
{
    const int g = g0 + blockIdx.y;
    if (gactive && !gactive[g]) return;
    if ((int)blockIdx.x == nb) {
        if (threadIdx.x < 3 * LANES) {
            const int term = threadIdx.x / LANES, lane = threadIdx.x % LANES;
            double s = 0.0;
            for (int q = 0; q < nwork; ++q) s += Epart[(((size_t)g * nwork + q) * 3 + term) * LANES + lane];
            E[(size_t)term * Npad + g * LANES + lane] = s;
        }
        return;
    }
    const int B = blockIdx.x;
    const int r0 = blk_ptr[B], r1 = blk_ptr[B + 1];
    T *__restrict__ dst = grad + ((size_t)g * Lpad + (size_t)B * TILE) * 9 * LANES;
    for (int e = threadIdx.x; e < REC_ELEMS; e += 256) {
        T s = accumulate ? dst[e] : (T)0;
        for (int r = r0; r < r1; ++r) s += recs[((size_t)g * nrec + blk_rec[r]) * REC_ELEMS + e];
        dst[e] = s;
    }
}

// One restraint evaluation of decoy groups [g0, g0+ng) out of Gtot against tables tb.
// X/grad/E are indexed by the GLOBAL group; the plan (work decomposition) by ng.
template <typename T>
int k1_launch(trx_ctx *ctx, trx_tables *tb, int Gtot, int g0, int ng, const T *d_xyz, int xstride, const float *wl,
              const double *w, const int *gactive, double *d_E, T *d_grad)
{
    Plan *plan = nullptr;
    int rc = tb->get_plan(ng, &plan);
    if (rc) return rc;
    void *recs = nullptr, *epart = nullptr;
    rc = ctx->get_scratch(sizeof(T) == 8 ? "k1_recs64" : "k1_recs32", (size_t)Gtot * std::max(1, tb->ntiles + tb->ntiles + tb->nb) * REC_ELEMS * sizeof(T), &recs);
    if (rc) return rc;
    rc = ctx->get_scratch("k1_epart", (size_t)Gtot * std::max(1, tb->ntiles + tb->nb) * 3 * LANES * sizeof(double), &epart);
    if (rc) return rc;
    K1Params<T> p;
    p.X = d_xyz;
    for (int t = 0; t < 4; ++t)
        p.tab[t] = reinterpret_cast<const typename Vec2<T>::type *>(sizeof(T) == 8 ? (const void *)tb->d_tab64[t] : (const void *)tb->d_tab32[t]);
    p.geom = reinterpret_cast<const KnotGeom<T> *>(sizeof(T) == 8 ? (const void *)tb->d_geom64 : (const void *)tb->d_geom32);
    p.pairrec = tb->d_pairrec;
    p.tileJ = tb->d_tileJ;
    p.work = plan->d_work;
    p.recs = (T *)recs;
    p.Epart = (double *)epart;
    p.Lpad = tb->Lpad;
    p.nrec = plan->nrec;
    p.nwork = plan->nwork;
    p.xstride = xstride;
    p.g0 = g0;
    p.gactive = gactive;
    p.wl = wl;
    p.Npad = Gtot * LANES;
    p.w0 = (T)(w ? w[0] : 0.0);
    p.w1 = (T)(w ? w[1] : 0.0);
    p.w2 = (T)(w ? w[2] : 0.0);
    if (plan->nwork > 0) {
        ctx->time_begin("restraints");
        restraints_kernel<T><<<dim3(ng, plan->nwork), K1_THREADS, 0, ctx->stream>>>(p);
        ctx->time_end("restraints");
        TRX_CUDA(cudaGetLastError());
    }
    ctx->time_begin("reduce");
    const int nbr = d_grad ? tb->nb : 0;
    reduce_kernel<T><<<dim3(nbr + 1, ng), 256, 0, ctx->stream>>>((const T *)recs, plan->nrec, plan->d_blk_ptr, plan->d_blk_rec,
                                                                   d_grad, tb->Lpad, nbr, 0, (const double *)epart, plan->nwork,
                                                                   d_E, Gtot * LANES, g0, gactive);
    ctx->time_end("reduce");
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}
template int k1_launch<float>(trx_ctx *, trx_tables *, int, int, int, const float *, int, const float *, const double *,
                              const int *, double *, float *);
template int k1_launch<double>(trx_ctx *, trx_tables *, int, int, int, const double *, int, const float *, const double *,
                               const int *, double *, double *);

}  // namespace trx

using namespace trx;

extern "C" {

int trx_energy_grad_device(trx_ctx *ctx, trx_tables *tb, int N, int precision, const void *d_xyz, const double w[3],
                           double *d_E, void *d_grad)
{
    TRX_REQUIRE(ctx && tb && d_xyz && w && d_E, "trx_energy_grad_device: NULL argument");
    TRX_REQUIRE(tb->ctx == ctx, "trx_energy_grad_device: tables belong to another context");
    TRX_REQUIRE(N > 0, "trx_energy_grad_device: N must be positive");
    TRX_REQUIRE(precision == TRX_F64 || precision == TRX_F32, "trx_energy_grad_device: precision must be 64 or 32");
    TRX_CUDA(cudaSetDevice(ctx->device));
    const int G = num_groups(N);
    if (precision == TRX_F64)
        return k1_launch<double>(ctx, tb, G, 0, G, (const double *)d_xyz, 9, nullptr, w, nullptr, d_E, (double *)d_grad);
    return k1_launch<float>(ctx, tb, G, 0, G, (const float *)d_xyz, 9, nullptr, w, nullptr, d_E, (float *)d_grad);
}

int trx_energy_grad(trx_ctx *ctx, trx_tables *tb, int N, int precision, const void *xyz, const double w[3], double *E,
                    void *grad)
{
    TRX_REQUIRE(ctx && tb && xyz && w && E, "trx_energy_grad: NULL argument");
    TRX_REQUIRE(tb->ctx == ctx, "trx_energy_grad: tables belong to another context");
    TRX_REQUIRE(N > 0, "trx_energy_grad: N must be positive");
    TRX_REQUIRE(precision == TRX_F64 || precision == TRX_F32, "trx_energy_grad: precision must be 64 or 32");
    TRX_CUDA(cudaSetDevice(ctx->device));
    const size_t es = precision == TRX_F64 ? 8 : 4;
    const int L = tb->L, Lpad = tb->Lpad, G = num_groups(N), Npad = G * LANES;
    const size_t nat_bytes = (size_t)N * L * 9 * es, grp_bytes = (size_t)G * Lpad * 9 * LANES * es;
    void *d_nat, *d_grp, *d_grad, *d_E;
    int rc;
    if ((rc = ctx->get_scratch("eg_nat", nat_bytes, &d_nat))) return rc;
    if ((rc = ctx->get_scratch("eg_grp", grp_bytes, &d_grp))) return rc;
    if ((rc = ctx->get_scratch("eg_grad", grp_bytes, &d_grad))) return rc;
    if ((rc = ctx->get_scratch("eg_E", (size_t)3 * Npad * sizeof(double), &d_E))) return rc;
    TRX_CUDA(cudaMemcpyAsync(d_nat, xyz, nat_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = trx_to_grouped(ctx, N, L, 3, precision, d_nat, d_grp))) return rc;
    if ((rc = trx_energy_grad_device(ctx, tb, N, precision, d_grp, w, (double *)d_E, grad ? d_grad : nullptr))) return rc;
    std::vector<double> Eh((size_t)3 * Npad);
    if (grad) {
        if ((rc = trx_from_grouped(ctx, N, L, 3, precision, d_grad, d_nat))) return rc;
        TRX_CUDA(cudaMemcpyAsync(grad, d_nat, nat_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRX_CUDA(cudaMemcpyAsync(Eh.data(), d_E, Eh.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int n = 0; n < N; ++n)
        for (int t = 0; t < 3; ++t) E[(size_t)n * 3 + t] = Eh[(size_t)t * Npad + n];
    return TRX_OK;
}

}  // extern "C"
