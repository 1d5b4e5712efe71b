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
// once per warp from L1/L2, never per decoy from HBM).  A CTA of K1_WARPS warps owns
// a block row of 16 residues for one decoy group and walks its active 16x16 tiles
// through a host-built step schedule: in a step every warp evaluates one residue pair
// and the pairs of a step have distinct rows and distinct columns, so row and column
// gradients are accumulated in shared memory with plain read-modify-writes: no
// atomics, bit-reproducible sums.  Per-tile column gradients and per-CTA row
// gradients are written as partial records that the reduce kernel sums in a fixed order.
#include <algorithm>
#include <cstdlib>

#include "internal.cuh"

// Register budget of the fp32 kernel: 5 CTAs/SM -> 96 registers, no spills.  Measured on B200
// (tools/k1_bench.py, L=300): 6 CTAs (80 regs, spills) 1.42 / 5.11 ms sparse / dense, 5 CTAs 1.25 / 4.52 ms,
// 4 CTAs (124 regs) 1.33 / 4.80 ms.
#ifndef TRX_K1_MINBLOCKS
#define TRX_K1_MINBLOCKS 5
#endif
// Shared-memory carve-out (percent of the 256 KB L1/shared array): 72 % = 164 KB keeps FOUR CTAs
// (16 warps) resident and leaves ~90 KB of L1 for the coordinate lines the warps of a tile re-read;
// the maximum carve-out (5 CTAs resident, ~28 KB of L1) is 2 % slower.
#ifndef TRX_K1_CARVEOUT
#define TRX_K1_CARVEOUT 72
#endif

namespace trx {

// every argument is clamped to >= 1e-12 before it gets here, so the denormal rescaling of rsqrtf
// (4 extra instructions per call, ~7 calls per pair) is dead weight: MUFU.RSQ directly
__device__ __forceinline__ float t_rsqrt(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double t_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float t_rcp(float x)   // same reasoning: MUFU.RCP without the denormal path
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double t_rcp(double x) { return 1.0 / x; }
#define ATAN_C0 9.999993354e-01f
#define ATAN_C1 -3.332986002e-01f
#define ATAN_C2 1.994655755e-01f
#define ATAN_C3 -1.390859233e-01f
#define ATAN_C4 9.642109384e-02f
#define ATAN_C5 -5.591120728e-02f
#define ATAN_C6 2.186222795e-02f
#define ATAN_C7 -4.054375906e-03f
// Branch-free atan2 for the throughput (fp32) mode: octant reduction to a = min/max in [0,1],
// degree-7 polynomial in a^2 for atan(a)/a (max error 4e-8 rad before rounding, ~1e-7 in fp32),
// selects instead of the special-case branches of atan2f (which break the instruction stream
// into reconvergence regions and stop the six restraints of a pair from overlapping).
__device__ __forceinline__ float t_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    const float a = mn * t_rcp(mx), s = a * a;
    float r = ATAN_C7;
    r = r * s + ATAN_C6; r = r * s + ATAN_C5; r = r * s + ATAN_C4; r = r * s + ATAN_C3;
    r = r * s + ATAN_C2; r = r * s + ATAN_C1; r = r * s + ATAN_C0;
    r *= a;
    r = ay > ax ? 1.57079632679f - r : r;
    r = x < 0.f ? 3.14159265359f - r : r;
    return copysignf(r, y);
}
__device__ __forceinline__ double t_atan2(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float t_floor(float x) { return floorf(x); }
__device__ __forceinline__ double t_floor(double x) { return floor(x); }
template <typename T> __device__ __forceinline__ T t_tiny();
template <> __device__ __forceinline__ float t_tiny<float>() { return 1e-12f; }
template <> __device__ __forceinline__ double t_tiny<double>() { return 1e-24; }

// What the kernel keeps of a KnotGeom in shared memory: the scalars of the interval guess and
// (in one array for the four types) the knot abscissae.  Sized so that six CTAs fit an SM.
template <typename T>
struct KnotHead {
    T gx0, ginv;
    int goff, K, urun0, pad;
};
constexpr int KX_TOTAL = MAXK + 3 * MAXK_ANG;
constexpr int K1_MAXSTEPS = TILE * TILE;   // a tile has at most 256 pairs, hence at most 256 steps
// dynamic shared memory of the restraint kernel: two gradient records, the tile's pair records, its schedule, one mbarrier
// (with the coordinate stage only what stage 2 asks for: shared memory decides how many CTAs are resident)
__host__ __device__ constexpr size_t k1_dyn_bytes(size_t elem, int stage) { return 2 * REC_ELEMS * elem + (stage > 0 ? TILE * TILE * 8 * sizeof(int) + K1_MAXSTEPS * K1_WARPS * sizeof(unsigned short) + 16 : 0) + (stage > 1 ? 2 * REC_ELEMS * elem : 0); }
__host__ __device__ constexpr int kx_off(int type) { return type == 0 ? 0 : MAXK + (type - 1) * MAXK_ANG; }

template <typename T>
struct K1Params {
    const T *X;                                  // [G][Lpad][xstride][32]
    const Coef<T> *tab[4];                       // [n][K] cubic per interval (+ flat tail entry)
    const KnotGeom<T> *geom;                     // [4]
    const int *pairrec;                          // [ntiles][16][16][8]
    const int *tileJ;                            // [ntiles]
    const unsigned short *sched;                 // [steps][8] pair of each warp: row | col<<4, 0xffff idle
    const int *nsteps;                           // [ntiles+1] CSR of steps per tile
    const int *work;                             // [nwork][4]
    T *recs;                                     // [G][nrec][REC_ELEMS]
    double *Epart;                               // [G][nwork][3][32]
    int Lpad, nrec, nwork;
    int xstride;                                 // values per residue in X (9: N,CA,CB; 15: fold layout)
    int dist_ca;                                 // distance restraints on CA-CA (af2 variant, distance-only tables)
    int g0;                                      // first decoy group of this launch
    int stage;                                   // staged in shared memory by bulk async copies: 1 pair records + schedule of the tile, 2 + its coordinates
    const int *gactive;                          // per-group flag or NULL (all active)
    const float *wl;                             // per-decoy weights [3][Npad] or NULL (use w0..w2)
    int Npad;
    T w0, w1, w2;
};

// Rosetta SplineFunc (weight 1): the clamped cubic spline inside [x_0, x_{K-1}], flat outside.
// Each interval is stored as a cubic in (x - x_k) -- algebraically the NR splint expression --
// and the flat tail beyond the last knot as one more "interval", so the common case is
// branch-free: interval index from the uniform part of the grid, one 16 B load, Horner.
// fp64 (parity mode) always corrects the index against the true knots (the reference's %.3f
// rounding makes the angular grids uneven at the 5e-4 level); fp32 only below the uniform run.
// Split in two so that a pair can issue the table loads of all its restraints together.
template <typename T>
__device__ __forceinline__ int spline_locate_exact(const KnotHead<T> &kn, const T *__restrict__ kx, T x, T &u)
{
    const int K = kn.K;
    int k = (int)t_floor((x - kn.gx0) * kn.ginv) + kn.goff;
    k = min(max(k, 0), K - 1);
    while (k > 0 && x < kx[k]) --k;
    while (k < K - 1 && x >= kx[k + 1]) ++k;
    u = x - kx[k];
    if (k == 0) u = max(u, (T)0);   // below the first knot: clamped spline => value y_0, slope 0
    return k;
}

// fp32: no loops, no divergent branches.  Inside the uniform run the guess is the interval
// (a guess that is one off next to a knot evaluates the neighbouring cubic <= 5e-4 outside
// its interval: error ~1e-8, the spline being C2); the few uneven intervals below the run
// (the 0 / 2 / 3.5 A knots of the distance grid; at most 6, checked at table creation) are
// counted with compares.
// HEAD: the grid may have uneven leading intervals (the distance grid); the angular grids are
// uniform from their first knot (checked at table creation), so their lookups skip the test.
template <bool HEAD>
__device__ __forceinline__ int spline_locate(const KnotHead<float> &kn, const float *__restrict__ kx, float x, float &u)
{
    const int K = kn.K, r0 = HEAD ? kn.urun0 : 0;
    int k = (int)floorf((x - kn.gx0) * kn.ginv) + kn.goff;
    if (HEAD && r0 > 0 && __any_sync(0xffffffffu, k < r0)) {   // rare: some decoy below the uniform run (d < 4.25 A)
        int kh = 0;
#pragma unroll
        for (int m = 1; m <= 6; ++m) kh += (m <= r0 && x >= kx[m]) ? 1 : 0;
        k = k >= r0 ? k : kh;
    }
    k = min(max(k, 0), K - 1);
    u = x - kx[k];
    u = k == 0 ? fmaxf(u, 0.f) : u;
    return k;
}
template <bool HEAD>
__device__ __forceinline__ int spline_locate(const KnotHead<double> &kn, const double *__restrict__ kx, double x, double &u)
{
    return spline_locate_exact(kn, kx, x, u);
}

// off = element offset of the restraint's first interval (< 2^31 by construction): one
// 32-bit add and one widening multiply-add instead of 64-bit pointer arithmetic per load
template <typename T>
__device__ __forceinline__ Coef<T> spline_load(const Coef<T> *__restrict__ tab, int off, int k) { return tab[(unsigned)(off + k)]; }
#ifdef TRX_K1_TAB_NOALLOC
// experiment: table gathers bypass L1 allocation, leaving L1 to the coordinate lines a tile re-reads
template <>
__device__ __forceinline__ Coef<float> spline_load<float>(const Coef<float> *__restrict__ tab, int off, int k)
{
    Coef<float> c;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c.c0), "=f"(c.c1), "=f"(c.c2), "=f"(c.c3) : "l"(tab + (unsigned)(off + k)));
    return c;
}
#endif

#define SPLINE_F(c, u) ((c).c0 + (u) * ((c).c1 + (u) * ((c).c2 + (u) * (c).c3)))
#define SPLINE_DF(c, u) ((c).c1 + (u) * ((T)2 * (c).c2 + (T)3 * (c).c3 * (u)))

#define CROSS(o, a, b)                       \
    o##x = a##y * b##z - a##z * b##y;        \
    o##y = a##z * b##x - a##x * b##z;        \
    o##z = a##x * b##y - a##y * b##x;
#define DOT(a, b) (a##x * b##x + a##y * b##y + a##z * b##z)

// Per-column quantities shared by the two rows a warp pairs with the column.
template <typename T>
struct ColGeom {
    T Bx, By, Bz;     // CB_j
    T Qx, Qy, Qz;     // CA_j - CB_j
    T Ux, Uy, Uz;     // N_j - CA_j
    T Wx, Wy, Wz;     // U x Q  (plane normal of N_j, CA_j, CB_j)
    T qq, ww, uq;
};

// All restraints of the unordered residue pair (i = row, j = column, i < j).
//   dist  AtomPair CB_i CB_j               omega  Dihedral CA_i CB_i CB_j CA_j
//   theta Dihedral N_a CA_a CB_a CB_b      phi    Angle CA_a CB_a CB_b     (both directions)
// With D = CB_j-CB_i, P = CA_i-CB_i, Q = CA_j-CB_j the two cross products X = D x P and
// Y = D x Q are the plane normals of ALL five angular restraints (Blondel-Karplus A/B
// vectors up to sign), and |X|, |Y| give the sines of the two phi angles, so the geometry
// is computed once per pair.  Phase 1 computes the six geometric values and issues the six
// table loads (independent gathers in flight together); phase 2 evaluates the cubics and
// accumulates the gradients.  Dihedral sign and range: IUPAC, equal to the reference's numpy
// get_dihedrals (utils_trX2dy/utils.py:97-110); angle: get_angles (:113-122).
// row: CB_i(0..2) P(3..5) U=N_i-CA_i(6..8); rg/cg: gradient accumulators N(0..2) CA(3..5) CB(6..8).
template <typename T, bool ALL>
__device__ __forceinline__ void pair_eval(const K1Params<T> &p, const KnotHead<T> *geom, const T *__restrict__ kx, const int4 ia, const int4 ib,
                                          const T *row, const ColGeom<T> &c, T *rg, T *cg, const T w0, const T w1, const T w2,
                                          T &e0, T &e1, T &e2)
{
    const int mask = ia.x;
    const T Dx = c.Bx - row[0], Dy = c.By - row[1], Dz = c.Bz - row[2];
    const T Px = row[3], Py = row[4], Pz = row[5];
    const T Qx = c.Qx, Qy = c.Qy, Qz = c.Qz;
    const T dd = max(DOT(D, D), t_tiny<T>());
    const T rd = t_rsqrt(dd);
    const T d = dd * rd;
    T Xx, Xy, Xz, Yx, Yy, Yz;
    CROSS(X, D, P);
    CROSS(Y, D, Q);
    const T xx = max(DOT(X, X), t_tiny<T>()), yy = max(DOT(Y, Y), t_tiny<T>());
    const T pd = DOT(P, D), qd = DOT(Q, D);
    const T pp = max(DOT(P, P), t_tiny<T>());
    const T rX = t_rsqrt(xx), rY = t_rsqrt(yy);
    // theta(i,j) plane normal of N_i, CA_i, CB_i
    const T Ux = row[6], Uy = row[7], Uz = row[8];
    T Wx, Wy, Wz;
    CROSS(W, U, P);
    const T ww = max(DOT(W, W), t_tiny<T>());
    const T rp = t_rsqrt(pp), np_ = pp * rp;
    const T rq = t_rsqrt(c.qq), nq = c.qq * rq;

    // ---- phase 1: values, intervals, loads
    T u0 = (T)0, u1 = (T)0, u2 = (T)0, u3 = (T)0, u4 = (T)0, u5 = (T)0;
    Coef<T> c0 = {(T)0, (T)0, (T)0, (T)0}, c1 = c0, c2 = c0, c3 = c0, c4 = c0, c5 = c0;
    if (ALL || (mask & 1)) c0 = spline_load(p.tab[0], ia.y, spline_locate<true>(geom[0], kx + kx_off(0), d, u0));
    if (ALL || (mask & 2))    // omega: F = P, G = -D, H = Q  =>  A = X, B = Y
        c1 = spline_load(p.tab[1], ia.z, spline_locate<false>(geom[1], kx + kx_off(1), t_atan2(-d * DOT(P, Y), DOT(X, Y)), u1));
    if (ALL || (mask & 4))    // theta(i,j): F = U_i, G = P, H = D  =>  A = U_i x P, B = X
        c2 = spline_load(p.tab[2], ia.w, spline_locate<false>(geom[2], kx + kx_off(2), t_atan2(-np_ * DOT(U, X), DOT(W, X)), u2));
    if (ALL || (mask & 8))    // theta(j,i): F = U_j, G = Q, H = -D  =>  A = W_j, B = -Y
        c3 = spline_load(p.tab[2], ib.x,
                         spline_locate<false>(geom[2], kx + kx_off(2), t_atan2(nq * (c.Ux * Yx + c.Uy * Yy + c.Uz * Yz), -(c.Wx * Yx + c.Wy * Yy + c.Wz * Yz)), u3));
    if (ALL || (mask & 16))   // phi(i,j): angle between P and D at CB_i; sin = |X| / (|P||D|)
        c4 = spline_load(p.tab[3], ib.y, spline_locate<false>(geom[3], kx + kx_off(3), t_atan2(xx * rX, pd), u4));
    if (ALL || (mask & 32))   // phi(j,i): angle between Q and -D at CB_j; sin = |Y| / (|Q||D|)
        c5 = spline_load(p.tab[3], ib.z, spline_locate<false>(geom[3], kx + kx_off(3), t_atan2(yy * rY, -qd), u5));

    // ---- phase 2: energies and gradients
    const T ixx = rX * rX, iyy = rY * rY;
    e0 += SPLINE_F(c0, u0);
    e1 += SPLINE_F(c1, u1) + SPLINE_F(c2, u2) + SPLINE_F(c3, u3);
    e2 += SPLINE_F(c4, u4) + SPLINE_F(c5, u5);
    {
        const T s = w0 * SPLINE_DF(c0, u0) * rd;
        cg[6] += s * Dx; cg[7] += s * Dy; cg[8] += s * Dz;
        rg[6] -= s * Dx; rg[7] -= s * Dy; rg[8] -= s * Dz;
    }
    if (ALL || (mask & 2)) {
        const T s = w1 * SPLINE_DF(c1, u1);
        const T a1 = -s * d * ixx, a4 = s * d * iyy, tA = -s * pd * ixx * rd, tB = -s * qd * iyy * rd;
        const T tx = tA * Xx - tB * Yx, ty = tA * Xy - tB * Yy, tz = tA * Xz - tB * Yz;
        rg[3] += a1 * Xx; rg[4] += a1 * Xy; rg[5] += a1 * Xz;                  // CA_i
        cg[3] += a4 * Yx; cg[4] += a4 * Yy; cg[5] += a4 * Yz;                  // CA_j
        rg[6] += tx - a1 * Xx; rg[7] += ty - a1 * Xy; rg[8] += tz - a1 * Xz;   // CB_i
        cg[6] -= tx + a4 * Yx; cg[7] -= ty + a4 * Yy; cg[8] -= tz + a4 * Yz;   // CB_j
    }
    if (ALL || (mask & 4)) {
        const T s = w1 * SPLINE_DF(c2, u2), iww = t_rcp(ww), up = DOT(U, P);
        const T a1 = -s * np_ * iww, a4 = s * np_ * ixx, tA = s * up * iww * rp, tB = s * pd * ixx * rp;
        const T tx = tA * Wx - tB * Xx, ty = tA * Wy - tB * Xy, tz = tA * Wz - tB * Xz;
        rg[0] += a1 * Wx; rg[1] += a1 * Wy; rg[2] += a1 * Wz;                  // N_i
        cg[6] += a4 * Xx; cg[7] += a4 * Xy; cg[8] += a4 * Xz;                  // CB_j
        rg[3] += tx - a1 * Wx; rg[4] += ty - a1 * Wy; rg[5] += tz - a1 * Wz;   // CA_i
        rg[6] -= tx + a4 * Xx; rg[7] -= ty + a4 * Xy; rg[8] -= tz + a4 * Xz;   // CB_i
    }
    if (ALL || (mask & 8)) {
        const T s = w1 * SPLINE_DF(c3, u3), iww = t_rcp(c.ww);
        const T a1 = -s * nq * iww, a4 = s * nq * iyy, tA = s * c.uq * iww * rq, tB = s * qd * iyy * rq;
        const T tx = tA * c.Wx - tB * Yx, ty = tA * c.Wy - tB * Yy, tz = tA * c.Wz - tB * Yz;
        cg[0] += a1 * c.Wx; cg[1] += a1 * c.Wy; cg[2] += a1 * c.Wz;                    // N_j
        rg[6] -= a4 * Yx; rg[7] -= a4 * Yy; rg[8] -= a4 * Yz;                          // CB_i  (a4 * B, B = -Y)
        cg[3] += tx - a1 * c.Wx; cg[4] += ty - a1 * c.Wy; cg[5] += tz - a1 * c.Wz;     // CA_j
        cg[6] -= tx - a4 * Yx; cg[7] -= ty - a4 * Yy; cg[8] -= tz - a4 * Yz;           // CB_j
    }
    if (ALL || (mask & 16)) {
        const T s = w2 * SPLINE_DF(c4, u4) * rX, a = pd * rp * rp, b = pd * rd * rd;
        const T ux = -s * (Dx - a * Px), uy = -s * (Dy - a * Py), uz = -s * (Dz - a * Pz);   // d/dCA_i
        const T vx = -s * (Px - b * Dx), vy = -s * (Py - b * Dy), vz = -s * (Pz - b * Dz);   // d/dCB_j
        rg[3] += ux; rg[4] += uy; rg[5] += uz;
        cg[6] += vx; cg[7] += vy; cg[8] += vz;
        rg[6] -= ux + vx; rg[7] -= uy + vy; rg[8] -= uz + vz;
    }
    if (ALL || (mask & 32)) {
        const T s = w2 * SPLINE_DF(c5, u5) * rY, a = qd * rq * rq, b = qd * rd * rd;
        const T ux = s * (Dx - a * Qx), uy = s * (Dy - a * Qy), uz = s * (Dz - a * Qz);      // d/dCA_j
        const T vx = -s * (Qx - b * Dx), vy = -s * (Qy - b * Dy), vz = -s * (Qz - b * Dz);   // d/dCB_i
        cg[3] += ux; cg[4] += uy; cg[5] += uz;
        rg[6] += vx; rg[7] += vy; rg[8] += vz;
        cg[6] -= ux + vx; cg[7] -= uy + vy; cg[8] -= uz + vz;
    }
}

// ---- packed fp32 path (sm_100 f32x2 arithmetic) -------------------------------------------------------
// Blackwell issues a two-wide fp32 FMA (SASS FFMA2 / FMUL2 / FADD2; PTX fma/mul/add.f32x2 on a 64-bit register
// pair) in one slot, and the restraint kernel is bound by instruction issue, not by bytes.  The restraints of a
// residue pair are symmetric under exchange of the two residues: theta(j,i) and phi(j,i) are theta(i,j) and
// phi(i,j) with the roles "self" and "other" swapped and D = CB_other - CB_self negated, and the omega
// gradient has the same structure.  A lane therefore evaluates BOTH SIDES of its pair at once, side 0
// (self = i) in the low half and side 1 (self = j) in the high half of every packed value: the geometry
// (P|Q, U_i|U_j, the plane normals X|Y' and W_i|W_j, their norms and dot products), the atan2 polynomials of
// the two thetas and the two phis, and the whole gradient assembly.  Registers and occupancy stay those of
// the one-decoy-per-lane kernel (the packed value P|Q replaces the two scalars P and Q).
// Measured alternative (profiles/r2_k1_variants.md): packing two DECOYS per lane needs 248 registers and twice
// the shared memory, halves the resident warps and is 1.35x slower than the scalar kernel.
struct F2 {
    unsigned long long v;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ F2(float a) { asm("mov.b64 %0, {%1, %1};" : "=l"(v) : "f"(a)); }
    __device__ __forceinline__ F2(float a, float b) { asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); }
};
__device__ __forceinline__ float lo(F2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
// no rounding modifier: ptxas may contract mul + add into FFMA2, exactly as it does for scalar float code
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; asm("add.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; asm("sub.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; asm("mul.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a) { return F2(0.f) - a; }   // one FADD2 with a negated operand
__device__ __forceinline__ F2 &operator+=(F2 &a, F2 b) { a = a + b; return a; }
__device__ __forceinline__ F2 &operator-=(F2 &a, F2 b) { a = a - b; return a; }
__device__ __forceinline__ F2 rsqrt2(F2 x) { return F2(t_rsqrt(lo(x)), t_rsqrt(hi(x))); }
__device__ __forceinline__ F2 rcp2(F2 x) { return F2(t_rcp(lo(x)), t_rcp(hi(x))); }
__device__ __forceinline__ F2 fmax2(F2 a, float m) { return F2(fmaxf(lo(a), m), fmaxf(hi(a), m)); }
// atan2 of both halves: the reduction to [0,1] and the final selects per half, the polynomial packed
__device__ __forceinline__ void atan2_2(F2 y, F2 x, float &rl, float &rh)
{
    const float xl = lo(x), xh = hi(x), yl = lo(y), yh = hi(y);
    const float axl = fabsf(xl), ayl = fabsf(yl), axh = fabsf(xh), ayh = fabsf(yh);
    const float mxl = fmaxf(fmaxf(axl, ayl), 1e-30f), mnl = fminf(axl, ayl);
    const float mxh = fmaxf(fmaxf(axh, ayh), 1e-30f), mnh = fminf(axh, ayh);
    const F2 a = F2(mnl, mnh) * F2(t_rcp(mxl), t_rcp(mxh)), s = a * a;
    F2 r = F2(ATAN_C7);
    r = r * s + F2(ATAN_C6); r = r * s + F2(ATAN_C5); r = r * s + F2(ATAN_C4); r = r * s + F2(ATAN_C3);
    r = r * s + F2(ATAN_C2); r = r * s + F2(ATAN_C1); r = r * s + F2(ATAN_C0);
    r = r * a;
    rl = lo(r); rh = hi(r);
    rl = ayl > axl ? 1.57079632679f - rl : rl;
    rh = ayh > axh ? 1.57079632679f - rh : rh;
    rl = xl < 0.f ? 3.14159265359f - rl : rl;
    rh = xh < 0.f ? 3.14159265359f - rh : rh;
    rl = copysignf(rl, yl);
    rh = copysignf(rh, yh);
}

// All restraints of the residue pair (i = row, j = column), both sides at once.  self[k] = (row | column) value k of
// N(0..2) CA(3..5) CB(6..8).  Packed conventions, half 0 / half 1:
//   Dp = CB_other - CB_self = ( D | -D ),  Pp = CA_self - CB_self = ( P | Q ),  Up = N_self - CA_self,
//   Xp = Dp x Pp = ( X | Y' = -Y ),  Wp = Up x Pp,  pdp = Pp.Dp = ( pd | -qd )
// theta(self,other) = atan2(-|Pp| Up.Xp, Wp.Xp),  phi(self,other) = atan2(|Xp|, pdp),  omega = atan2(d P.Y', -X.Y').
// Absent restraints of a pair get zero coefficients (value 0, slope 0): one straight-line path for every pair.
// G[9]: gradient of (row | column) atoms; adds to `other` CB cross over at the end.
__device__ __forceinline__ void pair_eval_sym(const K1Params<float> &p, const KnotHead<float> *geom, const float *__restrict__ kx, const int4 ia,
                                              const int4 ib, const F2 *self, F2 *G, const float w0, const float w1, const float w2,
                                              float &e0, float &e1, float &e2)
{
    typedef float T;
    const int mask = ia.x;
    const Coef<float> czero = {0.f, 0.f, 0.f, 0.f};
    const float Dx = hi(self[6]) - lo(self[6]), Dy = hi(self[7]) - lo(self[7]), Dz = hi(self[8]) - lo(self[8]);
    const F2 Dpx = F2(Dx, -Dx), Dpy = F2(Dy, -Dy), Dpz = F2(Dz, -Dz);
    const F2 Ppx = self[3] - self[6], Ppy = self[4] - self[7], Ppz = self[5] - self[8];
    const F2 Upx = self[0] - self[3], Upy = self[1] - self[4], Upz = self[2] - self[5];
    const float dd = fmaxf(DOT(D, D), 1e-12f), rd = t_rsqrt(dd), d = dd * rd;
    F2 Xpx, Xpy, Xpz, Wpx, Wpy, Wpz;
    CROSS(Xp, Dp, Pp);
    CROSS(Wp, Up, Pp);
    const F2 xxp = fmax2(DOT(Xp, Xp), 1e-12f), ppp = fmax2(DOT(Pp, Pp), 1e-12f), wwp = fmax2(DOT(Wp, Wp), 1e-12f);
    const F2 pdp = DOT(Pp, Dp), upp = DOT(Up, Pp);
    const F2 rXp = rsqrt2(xxp), rpp = rsqrt2(ppp), npp = ppp * rpp;

    // ---- phase 1: values, intervals, loads (six independent gathers in flight)
    float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f, u4 = 0.f, u5 = 0.f;
    Coef<float> c0 = czero, c1 = czero, c2 = czero, c3 = czero, c4 = czero, c5 = czero;
    {
        const int k = spline_locate<true>(geom[0], kx + kx_off(0), d, u0);
        if (mask & 1) c0 = spline_load(p.tab[0], ia.y, k);
    }
    {   // omega: F = P, G = -D, H = Q  =>  A = X, B = Y = -Y'
        const float Px = lo(Ppx), Py = lo(Ppy), Pz = lo(Ppz), Xx = lo(Xpx), Xy = lo(Xpy), Xz = lo(Xpz);
        const float Yx = hi(Xpx), Yy = hi(Xpy), Yz = hi(Xpz);   // Y'
        const int k = spline_locate<false>(geom[1], kx + kx_off(1), t_atan2(d * DOT(P, Y), -DOT(X, Y)), u1);
        if (mask & 2) c1 = spline_load(p.tab[1], ia.z, k);
    }
    {   // theta(i,j) | theta(j,i)
        float al, ah;
        atan2_2(-(npp * DOT(Up, Xp)), DOT(Wp, Xp), al, ah);
        const int kl = spline_locate<false>(geom[2], kx + kx_off(2), al, u2), kh = spline_locate<false>(geom[2], kx + kx_off(2), ah, u3);
        if (mask & 4) c2 = spline_load(p.tab[2], ia.w, kl);
        if (mask & 8) c3 = spline_load(p.tab[2], ib.x, kh);
    }
    {   // phi(i,j) | phi(j,i): sin = |Xp| / (|Pp||D|)
        float al, ah;
        atan2_2(xxp * rXp, pdp, al, ah);
        const int kl = spline_locate<false>(geom[3], kx + kx_off(3), al, u4), kh = spline_locate<false>(geom[3], kx + kx_off(3), ah, u5);
        if (mask & 16) c4 = spline_load(p.tab[3], ib.y, kl);
        if (mask & 32) c5 = spline_load(p.tab[3], ib.z, kh);
    }

    // ---- phase 2: energies and gradients
    const F2 ixxp = rXp * rXp;
    e0 += SPLINE_F(c0, u0);
    e1 += SPLINE_F(c1, u1) + SPLINE_F(c2, u2) + SPLINE_F(c3, u3);
    e2 += SPLINE_F(c4, u4) + SPLINE_F(c5, u5);
    F2 Ox = F2(0.f), Oy = F2(0.f), Oz = F2(0.f);   // gradient on the OTHER residue's CB, per side
    {   // dist: CB_self -= s Dp
        const F2 s = F2(w0 * SPLINE_DF(c0, u0) * rd);
        G[6] -= s * Dpx; G[7] -= s * Dpy; G[8] -= s * Dpz;
    }
    {   // omega: CA_self += a1 Xp;  CB_self += (tA - a1) Xp - tA_other X_other
        const float s = w1 * SPLINE_DF(c1, u1);
        const F2 sx = F2(s) * ixxp, a1 = -(F2(d) * sx), tA = -(pdp * sx * F2(rd)), za = tA - a1;
        G[3] += a1 * Xpx; G[4] += a1 * Xpy; G[5] += a1 * Xpz;
        G[6] += za * Xpx; G[7] += za * Xpy; G[8] += za * Xpz;
        Ox -= tA * Xpx; Oy -= tA * Xpy; Oz -= tA * Xpz;
    }
    {   // theta: N_self += a1 Wp;  CB_other += a4 Xp;  CA_self += t - a1 Wp;  CB_self -= t + a4 Xp,  t = tA Wp - tB Xp
        const F2 s = F2(w1) * F2(SPLINE_DF(c2, u2), SPLINE_DF(c3, u3)), iww = rcp2(wwp);
        const F2 sn = s * npp, a1 = -(sn * iww), a4 = sn * ixxp, tA = s * upp * iww * rpp, tB = s * pdp * ixxp * rpp;
        const F2 tx = tA * Wpx - tB * Xpx, ty = tA * Wpy - tB * Xpy, tz = tA * Wpz - tB * Xpz;
        G[0] += a1 * Wpx; G[1] += a1 * Wpy; G[2] += a1 * Wpz;
        Ox += a4 * Xpx; Oy += a4 * Xpy; Oz += a4 * Xpz;
        G[3] += tx - a1 * Wpx; G[4] += ty - a1 * Wpy; G[5] += tz - a1 * Wpz;
        G[6] -= tx + a4 * Xpx; G[7] -= ty + a4 * Xpy; G[8] -= tz + a4 * Xpz;
    }
    {   // phi: CA_self += u;  CB_other += v;  CB_self -= u + v
        const F2 s = F2(w2) * F2(SPLINE_DF(c4, u4), SPLINE_DF(c5, u5)) * rXp, a = pdp * rpp * rpp, b = pdp * F2(rd * rd);
        const F2 ux = s * (a * Ppx - Dpx), uy = s * (a * Ppy - Dpy), uz = s * (a * Ppz - Dpz);
        const F2 vx = s * (b * Dpx - Ppx), vy = s * (b * Dpy - Ppy), vz = s * (b * Dpz - Ppz);
        G[3] += ux; G[4] += uy; G[5] += uz;
        Ox += vx; Oy += vy; Oz += vz;
        G[6] -= ux + vx; G[7] -= uy + vy; G[8] -= uz + vz;
    }
    // the `other` of side 0 is the column residue, of side 1 the row residue
    G[6] += F2(hi(Ox), lo(Ox)); G[7] += F2(hi(Oy), lo(Oy)); G[8] += F2(hi(Oz), lo(Oz));
}

// One residue pair of one decoy per lane: coordinates in, gradients of the row residue (rg) and of the column
// residue (cg) out -- N(0..2) CA(3..5) CB(6..8) -- and the pair's energies added to e0..e2.
template <typename T, bool SYM>
__device__ __forceinline__ void eval_pair(const K1Params<T> &p, const KnotHead<T> *geom, const T *__restrict__ kx, const int4 ia, const int4 ib,
                                          const T *__restrict__ xr, const T *__restrict__ xc, T *rg, T *cg, const T w0, const T w1, const T w2,
                                          T &f0, T &f1, T &f2)
{
    T ri[9], cj[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        ri[k] = xr[k * LANES];
        cj[k] = xc[k * LANES];
        rg[k] = (T)0;
        cg[k] = (T)0;
    }
    if constexpr (SYM) {
        if (!p.dist_ca) {
            F2 self[9], G[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) { self[k] = F2(ri[k], cj[k]); G[k] = F2(0.f); }
            pair_eval_sym(p, geom, kx, ia, ib, self, G, w0, w1, w2, f0, f1, f2);
#pragma unroll
            for (int k = 0; k < 9; ++k) { rg[k] = lo(G[k]); cg[k] = hi(G[k]); }
            return;
        }
    }
    T row[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        row[k] = ri[6 + k];
        row[3 + k] = ri[3 + k] - ri[6 + k];
        row[6 + k] = ri[k] - ri[3 + k];
    }
    ColGeom<T> cgm;
    cgm.Bx = cj[6]; cgm.By = cj[7]; cgm.Bz = cj[8];
    cgm.Qx = cj[3] - cj[6]; cgm.Qy = cj[4] - cj[7]; cgm.Qz = cj[5] - cj[8];
    cgm.Ux = cj[0] - cj[3]; cgm.Uy = cj[1] - cj[4]; cgm.Uz = cj[2] - cj[5];
    cgm.Wx = cgm.Uy * cgm.Qz - cgm.Uz * cgm.Qy;
    cgm.Wy = cgm.Uz * cgm.Qx - cgm.Ux * cgm.Qz;
    cgm.Wz = cgm.Ux * cgm.Qy - cgm.Uy * cgm.Qx;
    cgm.qq = max(cgm.Qx * cgm.Qx + cgm.Qy * cgm.Qy + cgm.Qz * cgm.Qz, t_tiny<T>());
    cgm.ww = max(cgm.Wx * cgm.Wx + cgm.Wy * cgm.Wy + cgm.Wz * cgm.Wz, t_tiny<T>());
    cgm.uq = cgm.Ux * cgm.Qx + cgm.Uy * cgm.Qy + cgm.Uz * cgm.Qz;
    if (p.dist_ca) {   // 'AtomPair CA a CA b' (utils_ros.py:191): the only restraint of such a pair
        const T Dx = cj[3] - ri[3], Dy = cj[4] - ri[4], Dz = cj[5] - ri[5];
        const T dd = max(DOT(D, D), t_tiny<T>()), rd = t_rsqrt(dd);
        T u0;
        const Coef<T> c0 = spline_load(p.tab[0], ia.y, spline_locate<true>(geom[0], kx + kx_off(0), dd * rd, u0));
        f0 += SPLINE_F(c0, u0);
        const T sg = w0 * SPLINE_DF(c0, u0) * rd;
        cg[3] += sg * Dx; cg[4] += sg * Dy; cg[5] += sg * Dz;
        rg[3] -= sg * Dx; rg[4] -= sg * Dy; rg[5] -= sg * Dz;
    } else
    // pairs carrying all six restraints take a straight-line path (no per-restraint branches)
    if (ia.x == 63) pair_eval<T, true>(p, geom, kx, ia, ib, row, cgm, rg, cg, w0, w1, w2, f0, f1, f2);
    else pair_eval<T, false>(p, geom, kx, ia, ib, row, cgm, rg, cg, w0, w1, w2, f0, f1, f2);
}

// ---- the barrier-free fp32 kernel ------------------------------------------------------------------------
// The step schedule above exists so that plain shared-memory read-modify-writes never collide -- at the price of a
// block-wide barrier per step, which is the kernel's top stall (ncu, round 1) and couples the latencies of the
// four warps of a tile.  Here the row and column gradients of a tile are accumulated with native 32-bit
// SHARED-MEMORY INTEGER ATOMICS ON TWO-LEVEL FIXED POINT: a value v is split exactly into h = rint(v 2^10) and
// l = rint((v - h 2^-10) 2^34), added to an int32 `hi` and an int32 `lo` accumulator (range +-2^21, resolution
// 5.8e-11; at most 16 contributions per accumulator and tile, so `lo` cannot overflow).  Integer sums do not depend
// on the order of the additions: the result is bit-reproducible with no ordering at all between the warps.  The
// pairs of a tile are dealt round-robin to K1F_WARPS warps that never wait for each other: no per-step barrier,
// no idle slots from imperfect matchings, twice the warps per tile for the same shared memory per warp.
// (A 64-bit fixed-point atomicAdd on shared memory is a compare-and-swap loop in SASS -- ATOMS.CAST.SPIN.64.)
constexpr int K1F_WARPS = 8;
constexpr int K1F_THREADS = K1F_WARPS * 32;
constexpr float K1F_HI = 1024.0f, K1F_LO = 17179869184.0f;   // 2^10, 2^34
__host__ __device__ constexpr size_t k1f_dyn_bytes() { return 4 * REC_ELEMS * sizeof(int) + TILE * TILE * 8 * sizeof(int) + K1_MAXSTEPS * K1_WARPS * sizeof(unsigned short) + 16 + 2 * REC_ELEMS * sizeof(float); }
#ifndef TRX_K1F_MINBLOCKS
#define TRX_K1F_MINBLOCKS 2
#endif

__device__ __forceinline__ float fx_value(int h, int l) { return (float)((double)h * (1.0 / 1024.0) + (double)l * (1.0 / 17179869184.0)); }

template <bool SYM>
__global__ void __launch_bounds__(K1F_THREADS, TRX_K1F_MINBLOCKS) restraints_free_kernel(const K1Params<float> p)
{
    typedef float T;
    __shared__ KnotHead<T> geom[4];
    __shared__ T kx[KX_TOTAL];
    extern __shared__ __align__(16) unsigned char k1_dyn[];
    int *colhi = reinterpret_cast<int *>(k1_dyn), *collo = colhi + REC_ELEMS;   // column-block gradient of the current tile
    int *rowhi = collo + REC_ELEMS, *rowlo = rowhi + REC_ELEMS;                 // row-block gradient of the whole work item
    double(*ered)[3][LANES] = reinterpret_cast<double(*)[3][LANES]>(colhi);     // reused after the last flush
    int *recs_s = reinterpret_cast<int *>(k1_dyn + 4 * REC_ELEMS * sizeof(int));
    unsigned short *sched_s = reinterpret_cast<unsigned short *>(recs_s + TILE * TILE * 8);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sched_s + K1_MAXSTEPS * K1_WARPS);
    T *xrow_s = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(bar) + 16), *xcol_s = xrow_s + REC_ELEMS;   // staged coordinates
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.y, g = p.g0 + blockIdx.x;
    if (p.gactive && !p.gactive[g]) return;
    {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 4) {
            const KnotGeom<T> &gs = p.geom[threadIdx.x];
            KnotHead<T> h;
            h.gx0 = gs.gx0; h.ginv = gs.ginv; h.goff = gs.goff; h.K = gs.K; h.urun0 = gs.urun0; h.pad = 0;
            geom[threadIdx.x] = h;
        }
        for (int e = threadIdx.x; e < KX_TOTAL; e += K1F_THREADS) {
            const int t = e < MAXK ? 0 : 1 + (e - MAXK) / MAXK_ANG, k = e - kx_off(t);
            kx[e] = p.geom[t].x[k];
        }
        for (int e = threadIdx.x; e < 4 * REC_ELEMS; e += K1F_THREADS) colhi[e] = 0;
    }
    const int I = p.work[q * 4 + 0], t0 = p.work[q * 4 + 1], nt = p.work[q * 4 + 2], rowrec = p.work[q * 4 + 3];
    const int xs = p.xstride;
    size_t gbase = (size_t)g * p.Lpad * xs * LANES + lane;
    asm volatile("" : "+l"(gbase));
    const T *__restrict__ Xg = p.X + gbase;
    const unsigned xrow = (unsigned)xs * LANES;
    T w0 = p.w0, w1 = p.w1, w2 = p.w2;
    if (p.wl) {
        w0 = p.wl[0 * (size_t)p.Npad + g * LANES + lane];
        w1 = p.wl[1 * (size_t)p.Npad + g * LANES + lane];
        w2 = p.wl[2 * (size_t)p.Npad + g * LANES + lane];
    }
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    __syncthreads();

    for (int t = t0; t < t0 + nt; ++t) {
        const int J = p.tileJ[t];
        const int *__restrict__ rec_t = p.pairrec + (size_t)t * TILE * TILE * 8;
        T f0 = 0.f, f1 = 0.f, f2 = 0.f;   // this warp's partial energies of the tile
        const int s0 = p.nsteps[t], s1 = p.nsteps[t + 1];
        {   // the tile's pair records by one bulk asynchronous copy, its pair list by one coalesced load per thread
            const unsigned cb = 9 * LANES * sizeof(T);
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, TILE * TILE * 8 * sizeof(int) + (t == t0 ? 2 : 1) * TILE * cb);
                bulk_g2s(recs_s, rec_t, TILE * TILE * 8 * sizeof(int), bar);
            }
            if (w == 0) {
                __syncwarp();
                const T *__restrict__ Xgrp = p.X + (size_t)g * p.Lpad * xs * LANES;
                if (lane < TILE) {
                    if (t == t0) bulk_g2s(xrow_s + lane * 9 * LANES, Xgrp + (size_t)(I * TILE + lane) * xrow, cb, bar);
                } else {
                    bulk_g2s(xcol_s + (lane - TILE) * 9 * LANES, Xgrp + (size_t)(J * TILE + lane - TILE) * xrow, cb, bar);
                }
            }
            const uint2 *__restrict__ src = reinterpret_cast<const uint2 *>(p.sched + (size_t)s0 * K1_WARPS);
            for (int e = threadIdx.x; e < s1 - s0; e += K1F_THREADS) reinterpret_cast<uint2 *>(sched_s)[e] = src[e];
            mbar_wait(bar, (t - t0) & 1);
            __syncthreads();
        }
        // pair list of the tile = the entries of its step schedule (idle entries skipped), dealt round-robin to the warps
        const int nent = (s1 - s0) * K1_WARPS;
        for (int idx = w; idx < nent; idx += K1F_WARPS) {
            const int e = sched_s[idx];
            if (e == 0xffff) continue;
            const int r = e & 15, c = (e >> 4) & 15;
            const int4 *rp = reinterpret_cast<const int4 *>(recs_s + (r * TILE + c) * 8);
            const int4 ia = rp[0], ib = rp[1];
            T rg[9], cg[9];
            eval_pair<T, SYM>(p, geom, kx, ia, ib, xrow_s + r * 9 * LANES + lane, xcol_s + c * 9 * LANES + lane, rg, cg, w0, w1, w2, f0, f1, f2);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float hr = rintf(rg[k] * K1F_HI), hc = rintf(cg[k] * K1F_HI);
                const float lr = fmaf(hr, -1.0f / K1F_HI, rg[k]), lc = fmaf(hc, -1.0f / K1F_HI, cg[k]);   // exact remainders
                atomicAdd(&rowhi[(r * 9 + k) * LANES + lane], __float2int_rn(hr));
                atomicAdd(&rowlo[(r * 9 + k) * LANES + lane], __float2int_rn(lr * K1F_LO));
                atomicAdd(&colhi[(c * 9 + k) * LANES + lane], __float2int_rn(hc));
                atomicAdd(&collo[(c * 9 + k) * LANES + lane], __float2int_rn(lc * K1F_LO));
            }
        }
        e0 += (double)f0; e1 += (double)f1; e2 += (double)f2;
        __syncthreads();   // every pair of the tile is in
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + t) * REC_ELEMS;
        for (int e = threadIdx.x * 4; e < REC_ELEMS; e += K1F_THREADS * 4) {   // 16-byte stores
            const int4 h = *reinterpret_cast<const int4 *>(colhi + e), l = *reinterpret_cast<const int4 *>(collo + e);
            *reinterpret_cast<float4 *>(dst + e) = make_float4(fx_value(h.x, l.x), fx_value(h.y, l.y), fx_value(h.z, l.z), fx_value(h.w, l.w));
            *reinterpret_cast<int4 *>(colhi + e) = make_int4(0, 0, 0, 0);
            *reinterpret_cast<int4 *>(collo + e) = make_int4(0, 0, 0, 0);
        }
        __syncthreads();
    }
    {
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + rowrec) * REC_ELEMS;
        for (int e = threadIdx.x * 4; e < REC_ELEMS; e += K1F_THREADS * 4) {
            const int4 h = *reinterpret_cast<const int4 *>(rowhi + e), l = *reinterpret_cast<const int4 *>(rowlo + e);
            *reinterpret_cast<float4 *>(dst + e) = make_float4(fx_value(h.x, l.x), fx_value(h.y, l.y), fx_value(h.z, l.z), fx_value(h.w, l.w));
        }
    }
    __syncthreads();
    ered[w][0][lane] = e0;
    ered[w][1][lane] = e1;
    ered[w][2][lane] = e2;
    __syncthreads();
    if (w < 3) {
        double sum = 0.0;
#pragma unroll
        for (int k = 0; k < K1F_WARPS; ++k) sum += ered[k][w][lane];
        p.Epart[(((size_t)g * p.nwork + q) * 3 + w) * LANES + lane] = sum;
    }
}

// SYM (fp32 only): both sides of a pair packed in f32x2 (pair_eval_sym); otherwise the scalar formulas (fp64 parity
// mode; fp32 with TRX_K1_SCALAR=1 for A/B measurements)
template <typename T, bool SYM>
__global__ void __launch_bounds__(K1_THREADS, (sizeof(T) == 4 ? TRX_K1_MINBLOCKS : 1)) restraints_kernel(const K1Params<T> p)
{
    __shared__ KnotHead<T> geom[4];
    __shared__ T kx[KX_TOTAL];
    extern __shared__ __align__(16) unsigned char k1_dyn[];
    T *colg = reinterpret_cast<T *>(k1_dyn);      // column-block gradient of the current tile
    T *rowg = colg + REC_ELEMS;                    // row-block gradient of the whole work item
    double(*ered)[3][LANES] = reinterpret_cast<double(*)[3][LANES]>(colg);  // reused after the last flush
    static_assert(sizeof(double) * K1_WARPS * 3 * LANES <= sizeof(T) * REC_ELEMS, "energy scratch must fit");
    constexpr int KV = 16 / sizeof(T);
    struct alignas(16) KVec { T v[KV]; };
    KVec kzero;
#pragma unroll
    for (int k = 0; k < KV; ++k) kzero.v[k] = (T)0;
    // The tile's pair records (8 KB, one contiguous block) arrive by ONE bulk asynchronous copy (cp.async.bulk ->
    // UBLKCP, completion on an mbarrier) and its step schedule by one coalesced load per thread, while the CTA
    // zeroes its accumulators: the per-step chain schedule entry -> pair record -> table gather, three dependent
    // L2 round trips, becomes two shared-memory reads and one gather.
    int *recs_s = reinterpret_cast<int *>(k1_dyn + 2 * REC_ELEMS * sizeof(T));
    unsigned short *sched_s = reinterpret_cast<unsigned short *>(recs_s + TILE * TILE * 8);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sched_s + K1_MAXSTEPS * K1_WARPS);
    // The coordinates of the tile's 16 row and 16 column residues (N, CA, CB of 32 decoys: 1152 B per residue in
    // fp32) are staged the same way, 32 bulk copies issued by one warp: a pair then reads its 18 coordinate lines
    // from shared memory instead of L1/L2 -- each line is fetched once per tile instead of once per pair of its row
    // or column (8x on protein-like contact maps; L1 held only 12 % of these re-reads, ncu round 1).
    T *xrow_s = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(bar) + 16), *xcol_s = xrow_s + REC_ELEMS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.y, g = p.g0 + blockIdx.x;   // group is the fast grid index: co-resident CTAs share tiles
    if (p.gactive && !p.gactive[g]) return;
    {
        if (p.stage && threadIdx.x == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 4) {
            const KnotGeom<T> &gs = p.geom[threadIdx.x];
            KnotHead<T> h;
            h.gx0 = gs.gx0; h.ginv = gs.ginv; h.goff = gs.goff; h.K = gs.K; h.urun0 = gs.urun0; h.pad = 0;
            geom[threadIdx.x] = h;
        }
        for (int e = threadIdx.x; e < KX_TOTAL; e += K1_THREADS) {
            const int t = e < MAXK ? 0 : 1 + (e - MAXK) / MAXK_ANG, k = e - kx_off(t);
            kx[e] = p.geom[t].x[k];
        }
        for (int e = threadIdx.x; e < REC_ELEMS; e += K1_THREADS) { colg[e] = (T)0; rowg[e] = (T)0; }
    }
    const int I = p.work[q * 4 + 0], t0 = p.work[q * 4 + 1], nt = p.work[q * 4 + 2], rowrec = p.work[q * 4 + 3];
    const int xs = p.xstride;
    size_t gbase = (size_t)g * p.Lpad * xs * LANES + lane;
    asm volatile("" : "+l"(gbase));   // keep the group offset in registers (the compiler otherwise rebuilds it from %ctaid every step)
    const T *__restrict__ Xg = p.X + gbase;
    const unsigned xrow = (unsigned)xs * LANES;   // elements per residue; offsets within a group fit 32 bits
    T w0 = p.w0, w1 = p.w1, w2 = p.w2;
    if (p.wl) {
        w0 = (T)p.wl[0 * (size_t)p.Npad + g * LANES + lane];
        w1 = (T)p.wl[1 * (size_t)p.Npad + g * LANES + lane];
        w2 = (T)p.wl[2 * (size_t)p.Npad + g * LANES + lane];
    }
    double e0 = 0.0, e1 = 0.0, e2 = 0.0;
    __syncthreads();

    for (int t = t0; t < t0 + nt; ++t) {
        const int J = p.tileJ[t];
        const int *__restrict__ rec_t = p.pairrec + (size_t)t * TILE * TILE * 8;
        T f0 = (T)0, f1 = (T)0, f2 = (T)0;   // per-tile partial energies
        // host-built schedule: the (<= K1_WARPS) pairs of a step have distinct rows and distinct columns,
        // so one warp per pair can add its row and column gradients to shared memory without atomics
        const int s0 = p.nsteps[t], s1 = p.nsteps[t + 1];
        if (p.stage) {
            if (threadIdx.x == 0) {
                const unsigned cb = 9 * LANES * sizeof(T);
                mbar_expect_tx(bar, TILE * TILE * 8 * sizeof(int) + (p.stage > 1 ? (t == t0 ? 2 : 1) * TILE * cb : 0u));
                bulk_g2s(recs_s, rec_t, TILE * TILE * 8 * sizeof(int), bar);
            }
            if (p.stage > 1 && w == 0) {
                __syncwarp();   // the expect_tx above precedes every copy of this warp
                const unsigned cb = 9 * LANES * sizeof(T);
                const T *__restrict__ Xgrp = p.X + (size_t)g * p.Lpad * xs * LANES;
                if (lane < TILE) {
                    if (t == t0) bulk_g2s(xrow_s + lane * 9 * LANES, Xgrp + (size_t)(I * TILE + lane) * xrow, cb, bar);
                } else {
                    bulk_g2s(xcol_s + (lane - TILE) * 9 * LANES, Xgrp + (size_t)(J * TILE + lane - TILE) * xrow, cb, bar);
                }
            }
            const uint2 *__restrict__ src = reinterpret_cast<const uint2 *>(p.sched + (size_t)s0 * K1_WARPS);   // 8 B per step
            for (int e = threadIdx.x; e < s1 - s0; e += K1_THREADS) reinterpret_cast<uint2 *>(sched_s)[e] = src[e];
            mbar_wait(bar, (t - t0) & 1);
            __syncthreads();
        }
        for (int s = s0; s < s1; ++s) {
            const int e = p.stage ? sched_s[(s - s0) * K1_WARPS + w] : p.sched[(size_t)s * K1_WARPS + w];
            if (e != 0xffff) {
                const int r = e & 15, c = (e >> 4) & 15;
                int4 ia, ib;
                if (p.stage) {
                    const int4 *rp = reinterpret_cast<const int4 *>(recs_s + (r * TILE + c) * 8);
                    ia = rp[0]; ib = rp[1];
                } else {
                    const int4 *rp = reinterpret_cast<const int4 *>(rec_t + (r * TILE + c) * 8);
                    ia = __ldg(rp); ib = __ldg(rp + 1);
                }
                T rg[9], cg[9];
                if (p.stage > 1) eval_pair<T, SYM>(p, geom, kx, ia, ib, xrow_s + r * 9 * LANES + lane, xcol_s + c * 9 * LANES + lane, rg, cg, w0, w1, w2, f0, f1, f2);
                else eval_pair<T, SYM>(p, geom, kx, ia, ib, Xg + (unsigned)(I * TILE + r) * xrow, Xg + (unsigned)(J * TILE + c) * xrow, rg, cg, w0, w1, w2, f0, f1, f2);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    rowg[(r * 9 + k) * LANES + lane] += rg[k];
                    colg[(c * 9 + k) * LANES + lane] += cg[k];
                }
            }
            __syncthreads();
        }
        e0 += (double)f0; e1 += (double)f1; e2 += (double)f2;
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + t) * REC_ELEMS;
        for (int e = threadIdx.x * KV; e < REC_ELEMS; e += K1_THREADS * KV) {   // 16-byte copies
            *reinterpret_cast<KVec *>(dst + e) = *reinterpret_cast<const KVec *>(colg + e);
            *reinterpret_cast<KVec *>(colg + e) = kzero;
        }
        __syncthreads();
    }
    {
        T *__restrict__ dst = p.recs + ((size_t)g * p.nrec + rowrec) * REC_ELEMS;
        for (int e = threadIdx.x * KV; e < REC_ELEMS; e += K1_THREADS * KV)
            *reinterpret_cast<KVec *>(dst + e) = *reinterpret_cast<const KVec *>(rowg + e);
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
    // 16-byte accesses: V elements of a record per thread and load (records are REC_ELEMS contiguous values)
    constexpr int V = 16 / sizeof(T);
    struct alignas(16) Vec { T v[V]; };
    for (int e = threadIdx.x * V; e < REC_ELEMS; e += 256 * V) {
        Vec s;
        if (accumulate) s = *reinterpret_cast<const Vec *>(dst + e);
        else
#pragma unroll
            for (int k = 0; k < V; ++k) s.v[k] = (T)0;
        for (int r = r0; r < r1; ++r) {
            const Vec x = *reinterpret_cast<const Vec *>(recs + ((size_t)g * nrec + blk_rec[r]) * REC_ELEMS + e);
#pragma unroll
            for (int k = 0; k < V; ++k) s.v[k] += x.v[k];
        }
        *reinterpret_cast<Vec *>(dst + e) = s;
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
        p.tab[t] = reinterpret_cast<const Coef<T> *>(sizeof(T) == 8 ? (const void *)tb->d_tab64[t] : (const void *)tb->d_tab32[t]);
    p.geom = reinterpret_cast<const KnotGeom<T> *>(sizeof(T) == 8 ? (const void *)tb->d_geom64 : (const void *)tb->d_geom32);
    p.pairrec = tb->d_pairrec;
    p.tileJ = tb->d_tileJ;
    p.sched = tb->d_sched;
    p.nsteps = tb->d_nsteps;
    p.work = plan->d_work;
    p.recs = (T *)recs;
    p.Epart = (double *)epart;
    p.Lpad = tb->Lpad;
    p.nrec = plan->nrec;
    p.nwork = plan->nwork;
    p.xstride = xstride;
    p.dist_ca = tb->dist_ca;
    p.g0 = g0;
    static const int stage = [] { const char *ev = getenv("TRX_K1_STAGE"); return ev && ev[0] ? atoi(ev) : 0; }();   // 0 nothing staged (default: the smallest footprint wins, profiles/r2_k1_variants.md), 1 pair records, 2 + coordinates
    p.stage = sizeof(T) == 8 ? std::min(stage, 1) : stage;
    p.gactive = gactive;
    p.wl = wl;
    p.Npad = Gtot * LANES;
    p.w0 = (T)(w ? w[0] : 0.0);
    p.w1 = (T)(w ? w[1] : 0.0);
    p.w2 = (T)(w ? w[2] : 0.0);
    if (plan->nwork > 0) {
        ctx->time_begin("restraints");
        const size_t dyn = k1_dyn_bytes(sizeof(T), p.stage);
        // Default fp32 kernel: step schedule, scalar formulas, nothing staged (the smallest footprint).  The measured
        // alternatives stay selectable (profiles/r2_k1_variants.md): TRX_K1_SYM=1 both sides of a pair packed in f32x2,
        // TRX_K1_FREE=1 barrier-free fixed-point accumulation, TRX_K1_STAGE=0|1|2.
        static const bool scalar_f32 = [] { const char *ev = getenv("TRX_K1_SYM"); return !(ev && ev[0] && ev[0] != '0'); }();
        auto launch = [&](auto kern) -> int {
            static bool attr_set_dev[64] = {};   // function attributes are per device (one flag per instantiation of this lambda)
            if (!attr_set_dev[ctx->device & 63]) {
                TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_dyn_bytes(sizeof(T), sizeof(T) == 4 ? 2 : 1)));
                int carve = p.stage > 1 ? 86 : TRX_K1_CARVEOUT;   // with staged coordinates: 196 KB of shared memory, two 84 KB CTAs resident
                if (const char *ev = getenv("TRX_K1_CARVEOUT")) carve = atoi(ev);   // development knob: percent of the L1/shared array
                TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                attr_set_dev[ctx->device & 63] = true;
            }
            kern<<<dim3(ng, plan->nwork), K1_THREADS, dyn, ctx->stream>>>(p);
            return TRX_OK;
        };
        static const bool stepped_f32 = [] { const char *ev = getenv("TRX_K1_FREE"); return !(ev && ev[0] && ev[0] != '0'); }();
        auto launch_free = [&](auto kern) -> int {
            static bool attr_set_dev[64] = {};
            if (!attr_set_dev[ctx->device & 63]) {
                TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1f_dyn_bytes()));
                int carve = 86;   // 196 KB of shared memory: two 8-warp CTAs resident
                if (const char *ev = getenv("TRX_K1_CARVEOUT")) carve = atoi(ev);
                TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                attr_set_dev[ctx->device & 63] = true;
            }
            kern<<<dim3(ng, plan->nwork), K1F_THREADS, k1f_dyn_bytes(), ctx->stream>>>(p);
            return TRX_OK;
        };
        int rcl;
        if constexpr (sizeof(T) == 4) {
            if (stepped_f32) rcl = scalar_f32 ? launch(restraints_kernel<float, false>) : launch(restraints_kernel<float, true>);
            else rcl = scalar_f32 ? launch_free(restraints_free_kernel<false>) : launch_free(restraints_free_kernel<true>);
        } else rcl = launch(restraints_kernel<T, false>);
        if (rcl) return rcl;
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

/* Development aid (not part of include/trx2dyn.h): resident CTAs per SM of the fp32 restraint kernel. */
int trx_debug_k1_occupancy(int *out)
{
    const size_t dyn = k1_dyn_bytes(sizeof(float), 1);
    TRX_CUDA(cudaFuncSetAttribute(restraints_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    TRX_CUDA(cudaFuncSetAttribute(restraints_kernel<float, true>, cudaFuncAttributePreferredSharedMemoryCarveout, TRX_K1_CARVEOUT));
    TRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, restraints_kernel<float, true>, K1_THREADS, dyn));
    return TRX_OK;
}

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
