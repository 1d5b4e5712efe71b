// The centroid fold on device: NeRF (K2), soft-sphere vdw (K4), reverse-mode torsion
// gradient + rama/omega (K3), batched L-BFGS with non-monotone Armijo (K5) and the host
// loop that walks the reference's staged schedule.
//
// Replaces, for N decoys at once, folding/folding.py:109-171 (pose_from_sequence,
// set_random_dihedral, remove_clash, RepeatMover(min_mover,3), remove_clash) where every
// step is a PyRosetta MinMover.apply on one decoy in one process.
//
// Layout: every per-decoy vector is [group][element][32 lanes] so a warp is 32 decoys
// and all loads/stores are full lines.  All decoys advance in lock-step evaluation
// rounds; each decoy carries its own position in the schedule (run, weights, line-search
// state), so finished or back-tracking decoys never stall the others.  The non-restraint
// terms (vdw, rama, omega, cart_bonded) are approximations of Rosetta's
// (include/trx_centroid_model.h).
//
// The schedule is cut into segments of torsion-space runs and Cartesian runs
// (min_mover_cart, folding.py:100-102,170); a segment changes the degrees of freedom
// (torsions -> coordinates by NeRF, coordinates -> torsions by reading the dihedrals back),
// and the same L-BFGS kernels serve both: they only see a vector of ndof floats per decoy.
// A call folds any number of decoys through the batch's resident positions (continuous
// batching): within a segment a position whose decoy has left it is parked into a queue
// store and refilled in the same round; the queue is walked segment by segment.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "internal.cuh"
#include "../../include/trx_centroid_model.h"

namespace trx {

constexpr int NAT3 = TRX_NAT * 3;  // 15 values per residue
constexpr int NATP = 16;           // ... padded to 64 B in the natural-layout copies (xnat, gnat): a decoy's residue is four float4

// 15 values of one residue of one decoy <-> the padded natural layout, as four 16 B accesses
__device__ __forceinline__ void nat_store(float *__restrict__ p, const float *v)
{
    float4 *q = reinterpret_cast<float4 *>(p);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
    q[2] = make_float4(v[8], v[9], v[10], v[11]);
    q[3] = make_float4(v[12], v[13], v[14], 0.f);
}
__device__ __forceinline__ void nat_load(const float *__restrict__ p, float *v)
{
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = q[0], b = q[1], c = q[2], d = q[3];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y; v[14] = d.z;
}

struct Run {
    float w[TRX_NTERM];
    int max_iter;
    float tol;
    int clash_check;
    float clash_thr;
    int skip_to;
    int cartesian;
};

struct Model {   // centroid model constants in device-friendly (float) form
    float cen_s[20], r_cen[20], r_bb[5];
    float rama[2][TRX_RAMA_NB][5];
    float rama_off[2];
};
__constant__ Model c_model;

// status of a position between evaluation rounds.  ST_FINAL: the decoy has left the segment in progress; its
// accepted point is evaluated once more (all terms, trial point = accepted point) so that what is parked --
// torsions, coordinates, terms -- is one consistent evaluation; ST_DONE: nothing to do (finished or empty).
enum { ST_INIT = 0, ST_LS = 1, ST_DONE = 2, ST_FINAL = 3 };

// Monte-Carlo extension (trx_fold_mc): options of the call; cycles == 0 switches it off
struct McOpts {
    unsigned long long seed, id_offset;
    int cycles, block_min, block_max, mc_run;
    float sigma, kT;
};

struct FoldState {
    int N, Npad, G, L, Lpad, ndof, m, nruns;
    int ndof_t, ndof_c;      // torsion space: 3 L; Cartesian: 15 Lpad (the layout of X)
    int cart;                // the segment in progress is Cartesian
    int has_cart;            // the schedule has a Cartesian run (the queue keeps held coordinates)
    int k1skip;              // skip the restraint kernel for slot groups whose runs do not score restraints
    int seg_lo, seg_hi;      // runs of the segment in progress: [seg_lo, seg_hi)
    int seg_last;            // ... it is the last one: a decoy that leaves it is finished (results are written)
    int park_X;              // ... the next one is Cartesian: a decoy that leaves it parks its coordinates too
    // vectors [G][ndof][32]
    float *x, *g, *d, *xt, *gt;
    float *S, *Y;            // [G][ndof][m][32]
    float *gram;             // [G][2][M*M][32]: s_i.y_j and y_i.y_j of the stored pairs (M = padded history)
    float *lbpart;           // [G][LB_MAXCH][5M+8][32] partial sums of the chunks of sweep A
    float *lbcoef;           // [G][2M+3][32] coefficients of sweep B
    // per decoy [Npad]
    double *f, *fmem;        // accepted energy, last 3 accepted energies [3][Npad]
    float *alpha, *slope;
    int lb_M;                // compile-time history bound of the L-BFGS kernel in use (8, 16, 20 or 24)
    int *nmem, *hist, *head, *iter, *run, *bt, *status, *restart;
    int *evals, *iters;
    int *flags;              // [Npad] TRX_DECOY_* bits: what went wrong for the decoy, if anything (failure reporting)
    double *terms;           // [TRX_NTERM][Npad] unweighted terms of the last evaluation
    double *ft;              // [Npad] weighted total of the last evaluation
    float *wl;               // [TRX_NTERM][Npad] weights in force per decoy
    // geometry
    float *X;                // [G][Lpad][15][32]
    float *xnat, *gnat;      // [Npad][L][16] (15 values + pad)
    float *gk1;              // [G][Lpad][9][32]
    double *E3;              // [3][Npad]
    double *Evdw;            // [Npad]
    double *Ehb;             // [Npad] backbone hydrogen-bond term (slot order, like Evdw)
    // Verlet list of the vdw / hydrogen-bond pair search, per position: the residue pairs whose bounding spheres were
    // within reach + NBL_SKIN when the list was built, and the spheres at that time.  The list stands while no residue
    // has used up half the skin; the kernel rebuilds it in the same pass as a full scan otherwise.
    int *nbl_ok;             // [Npad] the position's list is usable
    float4 *nbl_ref;         // [Npad][L] bounding spheres (centre, radius) at build time
    unsigned short *nbl_j;   // [Npad][L][NBL_W] partners j > i of residue i, ascending
    unsigned short *nbl_cnt; // [Npad][L]
    int *gactive;            // [G] decoy group has an unfinished decoy (L-BFGS kernel)
    // slot space: the unfinished decoys of each table block, compacted to the front of the
    // block every round, so the evaluation kernels only touch live lanes
    int *perm;               // [Npad] slot -> position, -1 for an empty slot
    int *slot_of;            // [Npad] position -> slot of the evaluation round just made
    long long *k1count;      // [16] decoy evaluations the restraint kernel made per table block (roofline accounting)
    int *gslot;              // [G] slot group holds a live slot
    int *gneedk1;            // [G] ... and one whose run scores the restraints (others skip the restraint kernel)
    int *nslot;              // [16] live slots per table block
    float *wslot;            // [TRX_NTERM][Npad] weights in slot order
    int ntab;
    int tab_d0[16], tab_n[16];   // first decoy and decoy count of each table block
    // Monte-Carlo extension: state saved before a perturbation
    float *xsave;            // [G][ndof][32]
    double *fsave;           // [Npad]
    int *naccept;            // [Npad]
    int *mccyc;              // [Npad] perturbations made so far
    McOpts mc;
    // a decoy that went through a Cartesian run HOLDS those coordinates (in the queue store) and their
    // terms until a torsion-space run starts and rebuilds it with ideal bond geometry
    int *held;               // [Npad]
    double *theld;           // [TRX_NTERM][Npad]
    // Continuous batching.  A call folds nq_tab[t] decoys per table block through the tab_n[t] POSITIONS of
    // the block: every round the positions whose decoy has left the segment in progress are parked into
    // the queue store and refilled with the next waiting decoy, so the batch stays full until the queue
    // drains.  The schedule is walked segment by segment over the whole queue (a segment changes the
    // degrees of freedom); between segments a decoy is its queue record: torsions, held coordinates,
    // terms, run index, counters.  Queue ids are block-major, each block starting at a multiple of 32.
    int nq_tab[16], qd0[16], qc0[16];   // decoys of the call per block; first queue id; first caller index
    int Nqpad;
    int *qcursor;            // [16] queue entries of each block handed out so far
    int *qocc;               // [16] occupied positions of each block after the last turnover
    int *newid;              // [Npad] turnover plan: queue id to load, -1 = leave empty, -2 = no change
    float *q_tors;           // [Nqpad/32][3L][32]
    float *q_X;              // [Nqpad/32][Lpad*15][32] (only with a Cartesian run)
    double *q_terms;         // [TRX_NTERM][Nqpad]
    int *q_run, *q_held, *q_evals, *q_iters, *q_flags;   // [Nqpad]
    // results in the caller's order
    float *o_tors, *o_xyz;   // [Nq][3L], [Nq][L][15] (o_xyz may be NULL)
    double *o_terms;         // [Nq][TRX_NTERM]
    long long *o_stats;      // [Nq][3] evaluations, accepted iterations, accepted MC moves
    int *o_flags;            // [Nq] TRX_DECOY_* bits
    // Migration: the position of a decoy in the arrays above is not its identity.  When fewer than
    // half of a table block's positions hold unfinished decoys, the unfinished ones are moved to the
    // front of the block (swapped with finished ones), so the L-BFGS kernels -- which stream a whole
    // 32-decoy group if any of its decoys is unfinished -- stream only ~live/32 groups.
    int *orig;               // [Npad] queue id of the decoy held at each position, -1 = empty
    int *mig_a, *mig_b;      // [Npad] position pairs to swap (per table block, from its first position)
    int *mig_n;              // [16] pairs per table block
    const int *aa;           // [L]
    const Run *runs;
};

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ f3 unit(f3 a) { float r = rsqrtf(dot(a, a)); return r * a; }

// NeRF placement: |cd| = bond, angle(b,c,d) = ang (ca, sa its cos/sin), dihedral(a,b,c,d) = tor (ct, st).
__device__ __forceinline__ f3 place(f3 a, f3 b, f3 c, float bond, float ca, float sa, float ct, float st)
{
    f3 bc = unit(c - b);
    f3 n = unit(cross(b - a, bc));
    f3 m = cross(n, bc);
    return c + (-bond * ca) * bc + (bond * sa * ct) * m + (bond * sa * st) * n;
}

// K2: NeRF.  One CTA per slot group (lane = decoy), SEG_WARPS warps: warp w builds the chain
// segment [w*Lseg, (w+1)*Lseg) in its own local frame (first residue in canonical position),
// plus the first residue of the next segment, whose N/CA/C define the rigid transform from the
// next segment's frame to this one.  After one barrier every warp composes the transforms of
// the segments before it and maps its residues to the global frame: a two-level prefix over
// rigid motions instead of a 3L-long dependent chain.
// xt[g][L*3][32] -> X[g][Lpad][15][32] and the natural-layout copy xnat[slot][L][15].
// 32 segments (1024 threads, <= 64 registers): the dependent chain of a segment is L/32 residues;
// measured on the L=300 bench: torsion gradient 234 -> 171 ms per fold against 16 segments
#ifndef TRX_SEG_WARPS
#define TRX_SEG_WARPS 32
#endif
constexpr int SEG_WARPS = TRX_SEG_WARPS;
constexpr int SEG_THREADS = SEG_WARPS * 32;

__global__ void __launch_bounds__(SEG_THREADS) nerf_kernel(FoldState s)
{
    __shared__ float frame[SEG_WARPS][12][LANES];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;   // n = slot
    if (!s.gslot[g]) return;
    const int dec = s.perm[n];
    const bool live = dec >= 0;
    const int dn = live ? dec : 0;      // empty slots replay decoy 0's torsions: finite, ignored
    const float *__restrict__ t = s.xt + (size_t)(dn / LANES) * s.ndof * LANES + dn % LANES;
    float *__restrict__ X = s.X + (size_t)g * s.Lpad * NAT3 * LANES + lane;
    float *__restrict__ xn = s.xnat + (size_t)n * s.L * NATP;
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < TRX_NTERM; ++k) s.wslot[(size_t)k * s.Npad + n] = live ? s.wl[(size_t)k * s.Npad + dec] : 0.f;
    }
    const float caNCAC = cosf((float)TRX_A_N_CA_C), saNCAC = sinf((float)TRX_A_N_CA_C);
    const float caCACN = cosf((float)TRX_A_CA_C_N), saCACN = sinf((float)TRX_A_CA_C_N);
    const float caCNCA = cosf((float)TRX_A_C_N_CA), saCNCA = sinf((float)TRX_A_C_N_CA);
    const float caCACO = cosf((float)TRX_A_CA_C_O), saCACO = sinf((float)TRX_A_CA_C_O);
    const int L = s.L, Lseg = (L + SEG_WARPS - 1) / SEG_WARPS;
    const int r0 = warp * Lseg, r1 = min(L, r0 + Lseg);
    if (r0 < L) {
        f3 N = {0.f, 0.f, 0.f}, CA = {(float)TRX_B_N_CA, 0.f, 0.f};
        f3 C = {(float)TRX_B_N_CA - (float)TRX_B_CA_C * caNCAC, (float)TRX_B_CA_C * saNCAC, 0.f};
        for (int i = r0; i <= r1 && i < L; ++i) {
            if (i > r0) {
                const float psi_p = t[((i - 1) * 3 + 1) * LANES], omg_p = t[((i - 1) * 3 + 2) * LANES], phi = t[(i * 3 + 0) * LANES];
                float sn, cs;
                sincosf(psi_p, &sn, &cs);
                f3 Nn = place(N, CA, C, (float)TRX_B_C_N, caCACN, saCACN, cs, sn);
                sincosf(omg_p, &sn, &cs);
                f3 CAn = place(CA, C, Nn, (float)TRX_B_N_CA, caCNCA, saCNCA, cs, sn);
                sincosf(phi, &sn, &cs);
                f3 Cn = place(C, Nn, CAn, (float)TRX_B_CA_C, caNCAC, saNCAC, cs, sn);
                N = Nn; CA = CAn; C = Cn;
            }
            if (i == r1) {   // first residue of the next segment: its frame in this segment's coordinates
                const f3 e1 = unit(CA - N), ez = unit(cross(e1, C - N)), e2 = cross(ez, e1);
                float *fr = &frame[warp + 1][0][lane];
                fr[0 * LANES] = e1.x; fr[1 * LANES] = e2.x; fr[2 * LANES] = ez.x;
                fr[3 * LANES] = e1.y; fr[4 * LANES] = e2.y; fr[5 * LANES] = ez.y;
                fr[6 * LANES] = e1.z; fr[7 * LANES] = e2.z; fr[8 * LANES] = ez.z;
                fr[9 * LANES] = N.x; fr[10 * LANES] = N.y; fr[11 * LANES] = N.z;
                break;
            }
            const float psi = t[(i * 3 + 1) * LANES];
            float sn, cs;
            sincosf(psi, &sn, &cs);
            f3 O = place(N, CA, C, (float)TRX_B_C_O, caCACO, saCACO, -cs, -sn);   // torsion psi + pi
            f3 b = CA - N, c = C - CA, a = cross(b, c);
            f3 CB = (float)TRX_CB_A * a + (float)TRX_CB_B * b + (float)TRX_CB_C * c + CA;
            const float v[NAT3] = {N.x, N.y, N.z, CA.x, CA.y, CA.z, CB.x, CB.y, CB.z, C.x, C.y, C.z, O.x, O.y, O.z};
#pragma unroll
            for (int k = 0; k < NAT3; ++k) X[((size_t)i * NAT3 + k) * LANES] = v[k];
        }
    }
    __syncthreads();
    if (r0 >= L) return;
    // global frame of this segment: A_1 o A_2 o ... o A_warp
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, T[3] = {0.f, 0.f, 0.f};
    for (int j = 1; j <= warp; ++j) {
        float A[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) A[k] = frame[j][k][lane];
        float Rn[9], Tn[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = R[r * 3 + 0] * A[0 * 3 + c] + R[r * 3 + 1] * A[1 * 3 + c] + R[r * 3 + 2] * A[2 * 3 + c];
            Tn[r] = R[r * 3 + 0] * A[9] + R[r * 3 + 1] * A[10] + R[r * 3 + 2] * A[11] + T[r];
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) T[k] = Tn[k];
    }
    for (int i = r0; i < r1; ++i) {
        float v[NAT3];
#pragma unroll
        for (int k = 0; k < NAT3; ++k) v[k] = X[((size_t)i * NAT3 + k) * LANES];
#pragma unroll
        for (int a = 0; a < TRX_NAT; ++a) {
            const float x = v[a * 3], y = v[a * 3 + 1], z = v[a * 3 + 2];
            v[a * 3 + 0] = R[0] * x + R[1] * y + R[2] * z + T[0];
            v[a * 3 + 1] = R[3] * x + R[4] * y + R[5] * z + T[1];
            v[a * 3 + 2] = R[6] * x + R[7] * y + R[8] * z + T[2];
        }
        if (warp > 0) {
#pragma unroll
            for (int k = 0; k < NAT3; ++k) X[((size_t)i * NAT3 + k) * LANES] = v[k];
        }
        if (live) nat_store(xn + (size_t)i * NATP, v);
    }
}

// K4: soft-sphere repulsion of one decoy per CTA.  Atoms N,CA,CB,C,O + CEN (on the CA->CB
// ray) staged in shared memory as float4 (x,y,z,radius); warps scan residue rows for close
// CA pairs (ballot), then spread the 36 atom pairs of each close residue pair over lanes.
// Gradients of clashing atom pairs go through 32-bit fixed-point shared atomics (native
// ATOMS.ADD, no CAS loop): integer sums are order-independent => bit-reproducible.
constexpr int VDW_THREADS = 256;
constexpr float VDW_FIX = 65536.0f;   // 2^16: gradients as 32-bit fixed point (range +-32768, step 1.5e-5)
// two residues whose bounding spheres are farther apart than this beyond touching cannot hold a hydrogen bond:
// the spheres contain the N and O spheres, so |N - O| >= gap + r_N + r_O
constexpr float HB_MARGIN = (float)(TRX_HB_D0 + TRX_HB_W - 1.40 - 1.35) + 1e-3f;
constexpr int NBL_W = 64;            // list slots per residue (a residue with more partners in reach marks the list unusable)
constexpr float NBL_SKIN = 2.0f;     // A; a list stands while every residue's sphere has moved / grown by less than half of it
constexpr double VDW_EFIX = 4294967296.0;   // energies are summed as 64-bit fixed point (2^-32): exact, order-independent sums, so
                                             // an evaluation through the list and one through the full scan are the same bits

__global__ void __launch_bounds__(VDW_THREADS) vdw_kernel(FoldState s)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;   // slot
    if (s.perm[n] < 0) return;
    const int L = s.L;
    float4 *at = reinterpret_cast<float4 *>(smem_raw);                                   // [L][6]
    int *acc = reinterpret_cast<int *>(smem_raw + sizeof(float4) * 6 * L);               // [L][6][3]
    float4 *bsph = reinterpret_cast<float4 *>(smem_raw + sizeof(float4) * 6 * L + ((sizeof(int) * 18 * L + 15) / 16) * 16);   // [L] bounding spheres
    float2 *hreach = reinterpret_cast<float2 *>(bsph + L);   // [L] |CA - N|, |CA - O|: how far a residue's donor / acceptor atom is from its CA
    __shared__ long long ered[2 * (VDW_THREADS / 32)];
    const float *__restrict__ xn = s.xnat + (size_t)n * L * NATP;
    for (int i = threadIdx.x; i < L; i += VDW_THREADS) {
        const int aa = s.aa[i];
        float v[NAT3];
        nat_load(xn + (size_t)i * NATP, v);
#pragma unroll
        for (int a = 0; a < 5; ++a) at[i * 6 + a] = make_float4(v[a * 3], v[a * 3 + 1], v[a * 3 + 2], c_model.r_bb[a]);
        const float cs = c_model.cen_s[aa];
        at[i * 6 + 5] = make_float4(v[3] + cs * (v[6] - v[3]), v[4] + cs * (v[7] - v[4]), v[5] + cs * (v[8] - v[5]), c_model.r_cen[aa]);
        // bounding sphere of the residue's six soft spheres: centre = mean position, radius =
        // farthest sphere surface.  Two residues can only touch when their bounding spheres do.
        float cx = 0.f, cy = 0.f, cz = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) { cx += at[i * 6 + a].x; cy += at[i * 6 + a].y; cz += at[i * 6 + a].z; }
        cx *= (1.0f / 6.0f); cy *= (1.0f / 6.0f); cz *= (1.0f / 6.0f);
        float rb = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const float4 q = at[i * 6 + a];
            rb = fmaxf(rb, sqrtf((q.x - cx) * (q.x - cx) + (q.y - cy) * (q.y - cy) + (q.z - cz) * (q.z - cz)) + q.w);
        }
        bsph[i] = make_float4(cx, cy, cz, rb + 1e-3f);
        hreach[i] = make_float2(sqrtf((v[0] - v[3]) * (v[0] - v[3]) + (v[1] - v[4]) * (v[1] - v[4]) + (v[2] - v[5]) * (v[2] - v[5])) + 1e-3f,
                                sqrtf((v[12] - v[3]) * (v[12] - v[3]) + (v[13] - v[4]) * (v[13] - v[4]) + (v[14] - v[5]) * (v[14] - v[5])) + 1e-3f);
    }
    for (int e = threadIdx.x; e < L * 18; e += VDW_THREADS) acc[e] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // gradients are accumulated already weighted (vdw and the hydrogen-bond term have their own weights)
    const float w_vdw = s.wslot[(size_t)TRX_T_VDW * s.Npad + n], w_hb = s.wslot[(size_t)TRX_T_HB * s.Npad + n];
    long long e_thread = 0, e_hb = 0;
    // Backbone hydrogen bond donor residue i (N-H) -> acceptor residue j (C=O), include/trx_centroid_model.h.
    // Atoms: N_i, CA_i, C_{i-1}, O_j, C_j.  Called by one lane per queued residue pair and direction.
    auto hbond = [&](int i, int j) {
        if (i < 1 || abs(i - j) < TRX_HB_MINSEP || s.aa[i] == TRX_AA_PRO) return;
        const float4 N = at[i * 6 + TRX_AT_N], O = at[j * 6 + TRX_AT_O];
        const float rx = O.x - N.x, ry = O.y - N.y, rz = O.z - N.z, d2 = rx * rx + ry * ry + rz * rz;
        const float dmax = (float)(TRX_HB_D0 + TRX_HB_W), dmin = (float)(TRX_HB_D0 - TRX_HB_W);
        if (d2 >= dmax * dmax || d2 <= dmin * dmin) return;
        const float4 CA = at[i * 6 + TRX_AT_CA], Cp = at[(i - 1) * 6 + TRX_AT_C], C = at[j * 6 + TRX_AT_C];
        const float vx = 2.f * N.x - Cp.x - CA.x, vy = 2.f * N.y - Cp.y - CA.y, vz = 2.f * N.z - Cp.z - CA.z;
        const float qx = O.x - C.x, qy = O.y - C.y, qz = O.z - C.z;
        const float id = rsqrtf(d2), iv = rsqrtf(vx * vx + vy * vy + vz * vz), iq = rsqrtf(qx * qx + qy * qy + qz * qz);
        const float rhx = rx * id, rhy = ry * id, rhz = rz * id, vhx = vx * iv, vhy = vy * iv, vhz = vz * iv;
        const float qhx = qx * iq, qhy = qy * iq, qhz = qz * iq;
        const float c1 = vhx * rhx + vhy * rhy + vhz * rhz, c2 = -(qhx * rhx + qhy * rhy + qhz * rhz);
        if (c1 <= 0.f || c2 <= 0.f) return;
        const float d = d2 * id, t = (d - (float)TRX_HB_D0) * (float)(1.0 / TRX_HB_W), u = 1.f - t * t;
        const float F = u * u, dF = -4.f * u * t * (float)(1.0 / TRX_HB_W), G1 = c1 * c1, G2 = c2 * c2;
        e_hb += __float2ll_rn(-(float)TRX_HB_EPS * F * G1 * G2 * (float)VDW_EFIX);
        const float a = -(float)TRX_HB_EPS * w_hb;
        const float kd = a * dF * G1 * G2, k1 = a * F * 2.f * c1 * G2, k2 = a * F * G1 * 2.f * c2;
        const float grx = kd * rhx + (k1 * (vhx - c1 * rhx) - k2 * (qhx + c2 * rhx)) * id;
        const float gry = kd * rhy + (k1 * (vhy - c1 * rhy) - k2 * (qhy + c2 * rhy)) * id;
        const float grz = kd * rhz + (k1 * (vhz - c1 * rhz) - k2 * (qhz + c2 * rhz)) * id;
        const float gvx = k1 * (rhx - c1 * vhx) * iv, gvy = k1 * (rhy - c1 * vhy) * iv, gvz = k1 * (rhz - c1 * vhz) * iv;
        const float gqx = -k2 * (rhx + c2 * qhx) * iq, gqy = -k2 * (rhy + c2 * qhy) * iq, gqz = -k2 * (rhz + c2 * qhz) * iq;
        auto add3 = [&](int res, int atom, float x, float y, float z) {
            int *pa = acc + (res * 6 + atom) * 3;
            atomicAdd(pa + 0, __float2int_rn(x * VDW_FIX)); atomicAdd(pa + 1, __float2int_rn(y * VDW_FIX)); atomicAdd(pa + 2, __float2int_rn(z * VDW_FIX));
        };
        add3(j, TRX_AT_O, grx + gqx, gry + gqy, grz + gqz);
        add3(j, TRX_AT_C, -gqx, -gqy, -gqz);
        add3(i, TRX_AT_N, 2.f * gvx - grx, 2.f * gvy - gry, 2.f * gvz - grz);
        add3(i, TRX_AT_CA, -gvx, -gvy, -gvz);
        add3(i - 1, TRX_AT_C, -gvx, -gvy, -gvz);
    };
    // per-warp queue of close residue pairs
    __shared__ int queue[VDW_THREADS / 32][64];
    int qn = 0;
    // lane -> (queued pair q = lane / 6, atom a = lane % 6 of residue i), fixed for the whole kernel; a pass
    // takes 5 queued residue pairs (30 lanes) and every lane tests its atom against the 6 atoms of residue j
    // (broadcast reads): 180 atom pairs per pass with no index arithmetic in the loop
    const int fq = lane / 6, fa = lane - 6 * fq;
    auto flush = [&](int count) {
        for (int base = 0; base < count; base += 5) {
            const int q = base + fq;
            if (fq < 5 && q < count) {
                const int pr = queue[warp][q];
                const int i = pr >> 16, j = pr & 0xffff;
                const float4 pa = at[i * 6 + fa];
                int *pi = acc + (i * 6 + fa) * 3;
                // second-level filter: this atom against the other residue's bounding sphere (most atoms of two residues
                // whose spheres touch are still out of each other's reach)
                const float4 sj = bsph[j];
                const float ex = pa.x - sj.x, ey = pa.y - sj.y, ez = pa.z - sj.z, er = pa.w + sj.w;
                const bool near = ex * ex + ey * ey + ez * ez < er * er;
#pragma unroll
                for (int b = 0; near && b < 6; ++b) {
                    const float4 pb = at[j * 6 + b];
                    const float dx = pa.x - pb.x, dy = pa.y - pb.y, dz = pa.z - pb.z;
                    const float r = pa.w + pb.w, r2 = r * r, d2 = dx * dx + dy * dy + dz * dz;
                    if (d2 < r2) {
                        const float c = r2 - d2, ir2 = 1.0f / r2;
                        e_thread += __float2ll_rn((float)TRX_VDW_SCALE * c * c * ir2 * (float)VDW_EFIX);
                        const float f = -4.0f * (float)TRX_VDW_SCALE * c * ir2 * w_vdw;
                        const int gx = __float2int_rn(f * dx * VDW_FIX), gy = __float2int_rn(f * dy * VDW_FIX), gz = __float2int_rn(f * dz * VDW_FIX);
                        int *pj = acc + (j * 6 + b) * 3;
                        atomicAdd(pi + 0, gx); atomicAdd(pi + 1, gy); atomicAdd(pi + 2, gz);
                        atomicAdd(pj + 0, -gx); atomicAdd(pj + 1, -gy); atomicAdd(pj + 2, -gz);
                    }
                }
            }
        }
    };
    // residue pairs in reach of a hydrogen bond have their own per-warp queue: one lane per pair, both directions
    __shared__ int hqueue[VDW_THREADS / 32][64];
    int hn = 0;
    auto flush_hb = [&](int count) {
        if (lane < count) {
            const int pr = hqueue[warp][lane];
            const int i = pr >> 16, j = pr & 0xffff;
            hbond(i, j);
            hbond(j, i);
        }
    };
    // a candidate pair (i, j) of this warp's row goes to the queue when its spheres are within reach now
    auto push = [&](bool close, bool touching, int i, int j) {
        // a hydrogen bond between the two residues needs N...O < D0 + W, hence CA...CA < D0 + W + |CA-N| + |CA-O| of the
        // donor / acceptor (the residues' own distances, so the filter is exact for any geometry)
        bool hb = close && j - i >= TRX_HB_MINSEP;
        if (hb) {
            const float4 a = at[i * 6 + TRX_AT_CA], b = at[j * 6 + TRX_AT_CA];
            const float2 ri = hreach[i], rj = hreach[j];
            const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
            const float reach = (float)(TRX_HB_D0 + TRX_HB_W) + fmaxf(ri.x + rj.y, rj.x + ri.y);
            hb = dx * dx + dy * dy + dz * dz < reach * reach;
        }
        const unsigned mh = __ballot_sync(0xffffffffu, hb);
        if (mh) {
            const int pos = hn + __popc(mh & ((1u << lane) - 1));
            if (hb) hqueue[warp][pos] = (i << 16) | j;
            hn += __popc(mh);
            __syncwarp();
            if (hn >= 32) {
                flush_hb(32);
                __syncwarp();
                const int rest = hn - 32;
                const int v = lane < rest ? hqueue[warp][32 + lane] : 0;
                __syncwarp();
                if (lane < rest) hqueue[warp][lane] = v;
                hn = rest;
                __syncwarp();
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, close && touching);
        if (m) {
            const int pos = qn + __popc(m & ((1u << lane) - 1));
            if (close && touching) queue[warp][pos] = (i << 16) | j;
            qn += __popc(m);
            __syncwarp();
            if (qn >= 30) {   // 30 pairs = 6 full passes of 5 pairs
                flush(30);
                __syncwarp();
                const int rest = qn - 30;   // < 32: at most 31 queued before this scan step
                const int v = lane < rest ? queue[warp][30 + lane] : 0;
                __syncwarp();
                if (lane < rest) queue[warp][lane] = v;
                qn = rest;
                __syncwarp();
            }
        }
    };
    // + the reach of a hydrogen bond beyond touching N and O spheres (N...O up to D0 + W)
    auto gap2 = [&](const float4 ci, const float4 cj, float extra, float &cut) {
        const float dx = ci.x - cj.x, dy = ci.y - cj.y, dz = ci.z - cj.z;
        cut = ci.w + cj.w + HB_MARGIN + extra;
        return dx * dx + dy * dy + dz * dz;
    };
    auto touch = [&](const float4 ci, const float4 cj, float d2) { const float r = ci.w + cj.w; return d2 < r * r; };
    const int dec = s.perm[n];
    bool use_list = false;
    if (s.nbl_j) {   // does the position's list stand?  every residue must have moved / grown by less than half the skin
        const int ok = s.nbl_ok[dec];
        int stale = ok ? 0 : 1;
        if (ok) {
            const float4 *__restrict__ ref = s.nbl_ref + (size_t)dec * L;
            for (int i = threadIdx.x; i < L; i += VDW_THREADS) {
                const float4 b = bsph[i], r0 = ref[i];
                const float dx = b.x - r0.x, dy = b.y - r0.y, dz = b.z - r0.z;
                if (sqrtf(dx * dx + dy * dy + dz * dz) + fmaxf(b.w - r0.w, 0.f) > 0.5f * NBL_SKIN) stale = 1;
            }
        }
        use_list = !__syncthreads_or(stale);
    }
    if (use_list) {
        const unsigned short *__restrict__ lj = s.nbl_j + (size_t)dec * L * NBL_W, *__restrict__ lc = s.nbl_cnt + (size_t)dec * L;
        for (int i = warp; i < L - TRX_VDW_MINSEP; i += VDW_THREADS / 32) {
            const float4 ci = bsph[i];
            const int cnt = lc[i];
            for (int e0 = 0; e0 < cnt; e0 += 32) {
                bool close = false, tch = false;
                int j = 0;
                if (e0 + lane < cnt) {
                    j = lj[(size_t)i * NBL_W + e0 + lane];
                    float cut;
                    const float4 cj = bsph[j];
                    const float d2 = gap2(ci, cj, 0.f, cut);
                    close = d2 < cut * cut;
                    tch = touch(ci, cj, d2);
                }
                push(close, tch, i, j);
            }
        }
    } else {
        // full scan; with lists enabled it also rebuilds the position's list (partners within reach + skin, in order)
        unsigned short *__restrict__ lj = s.nbl_j ? s.nbl_j + (size_t)dec * L * NBL_W : nullptr;
        unsigned short *__restrict__ lc = s.nbl_j ? s.nbl_cnt + (size_t)dec * L : nullptr;
        int overflow = 0;
        if (lj) {
            float4 *__restrict__ ref = s.nbl_ref + (size_t)dec * L;
            for (int i = threadIdx.x; i < L; i += VDW_THREADS) ref[i] = bsph[i];
            for (int i = L - TRX_VDW_MINSEP + (int)threadIdx.x; i < L; i += VDW_THREADS) if (i >= 0) lc[i] = 0;
        }
        for (int i = warp; i < L - TRX_VDW_MINSEP; i += VDW_THREADS / 32) {
            const float4 ci = bsph[i];
            int cnt = 0;
            // One candidate j per lane and half-step, two half-steps per iteration; most iterations find nothing in
            // reach, so the common path is two sphere tests and ONE vote.
            auto handle = [&](int j, const float4 cj, float d2, float cut) {
                const bool close = d2 < cut * cut, tch = touch(ci, cj, d2);
                if (lj) {
                    const bool inl = d2 < (cut + NBL_SKIN) * (cut + NBL_SKIN);
                    const unsigned ml = __ballot_sync(0xffffffffu, inl);
                    const int at_ = cnt + __popc(ml & ((1u << lane) - 1));
                    if (inl && at_ < NBL_W) lj[(size_t)i * NBL_W + at_] = (unsigned short)j;
                    cnt += __popc(ml);
                }
                push(close, tch, i, j);
            };
            const float extra = lj ? NBL_SKIN : 0.f;
            for (int j0 = i + TRX_VDW_MINSEP; j0 < L; j0 += 64) {
                const int ja = j0 + lane, jb = ja + 32;
                float4 cja = ci, cjb = ci;
                float d2a = 1e30f, d2b = 1e30f, cuta = 0.f, cutb = 0.f;
                if (ja < L) { cja = bsph[ja]; d2a = gap2(ci, cja, 0.f, cuta); }
                if (jb < L) { cjb = bsph[jb]; d2b = gap2(ci, cjb, 0.f, cutb); }
                const bool hit = d2a < (cuta + extra) * (cuta + extra) || d2b < (cutb + extra) * (cutb + extra);
                if (!__any_sync(0xffffffffu, hit)) continue;
                handle(ja, cja, d2a, cuta);
                handle(jb, cjb, d2b, cutb);
            }
            if (lj) {
                if (cnt > NBL_W) overflow = 1;
                if (lane == 0) lc[i] = (unsigned short)min(cnt, NBL_W);
            }
        }
        if (lj) {
            const int any = __syncthreads_or(overflow);
            if (threadIdx.x == 0) s.nbl_ok[dec] = any ? 0 : 1;
        }
    }
    flush(qn);
    flush_hb(hn);
    // energy: fixed-shape reduction
    for (int o = 16; o > 0; o >>= 1) { e_thread += __shfl_down_sync(0xffffffffu, e_thread, o); e_hb += __shfl_down_sync(0xffffffffu, e_hb, o); }
    if (lane == 0) { ered[warp] = e_thread; ered[VDW_THREADS / 32 + warp] = e_hb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long e = 0, eh = 0;
        for (int k = 0; k < VDW_THREADS / 32; ++k) { e += ered[k]; eh += ered[VDW_THREADS / 32 + k]; }
        s.Evdw[n] = (double)e * (1.0 / VDW_EFIX);
        s.Ehb[n] = (double)eh * (1.0 / VDW_EFIX);
    }
    // gradient out (already weighted), CEN folded into CA and CB
    const float w = 1.0f;
    float *__restrict__ gn = s.gnat + (size_t)n * L * NATP;
    for (int i = threadIdx.x; i < L; i += VDW_THREADS) {
        const float cs = c_model.cen_s[s.aa[i]];
        float gv[18];
#pragma unroll
        for (int k = 0; k < 18; ++k) gv[k] = w * ((float)acc[i * 18 + k] * (1.0f / VDW_FIX));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            gv[TRX_AT_CA * 3 + k] += (1.0f - cs) * gv[15 + k];
            gv[TRX_AT_CB * 3 + k] += cs * gv[15 + k];
        }
        nat_store(gn + (size_t)i * NATP, gv);
    }
}

// K3: reverse mode.  dE/dtorsion = u.(F1 - p x F2) with F1 = sum x_a x g_a, F2 = sum g_a over
// the atoms the torsion moves (a suffix of the atom sequence N,CA,CB,C,O).  One CTA per slot
// group (lane = decoy), SEG_WARPS warps: each warp walks its chain segment backwards with
// segment-local sums, the segment totals are exchanged through shared memory, and a second,
// dependency-free sweep adds the contribution of everything downstream of the segment.
// Also adds the Ramachandran and omega terms (functions of the torsions alone) and the total.
__global__ void __launch_bounds__(SEG_THREADS) torsion_grad_kernel(FoldState s)
{
    __shared__ float tot[SEG_WARPS][6][LANES];
    __shared__ double esum[SEG_WARPS][2][LANES];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;   // n = slot
    if (!s.gslot[g]) return;
    const int dec = s.perm[n];
    const bool live = dec >= 0;
    const int dn = live ? dec : 0;
    const int L = s.L, Npad = s.Npad;
    const float *__restrict__ X = s.X + (size_t)g * s.Lpad * NAT3 * LANES + lane;
    const float *__restrict__ G1 = s.gk1 + (size_t)g * s.Lpad * 9 * LANES + lane;
    const float *__restrict__ gn = s.gnat + (size_t)n * L * NATP;
    const size_t dvb = (size_t)(dn / LANES) * s.ndof * LANES + dn % LANES;
    const float *__restrict__ t = s.xt + dvb;
    float *__restrict__ gt = s.gt + dvb;
    const float w_rama = s.wslot[(size_t)TRX_T_RAMA * Npad + n], w_omega = s.wslot[(size_t)TRX_T_OMEGA * Npad + n];
    const bool k1 = s.gneedk1[g] != 0;
    const int Lseg = (L + SEG_WARPS - 1) / SEG_WARPS;
    const int r0 = warp * Lseg, r1 = min(L, r0 + Lseg);
    f3 F1 = {0.f, 0.f, 0.f}, F2 = {0.f, 0.f, 0.f};
    double e_rama = 0.0, e_omega = 0.0;
    auto load = [&](int i, int a) -> f3 { return {X[((size_t)i * NAT3 + a * 3) * LANES], X[((size_t)i * NAT3 + a * 3 + 1) * LANES], X[((size_t)i * NAT3 + a * 3 + 2) * LANES]}; };
    auto add = [&](f3 x, f3 gr) { F1 = F1 + cross(x, gr); F2 = F2 + gr; };
    auto dtor = [&](f3 p, f3 q, f3 A1, f3 A2) -> float { f3 u = unit(q - p); return dot(u, A1 - cross(p, A2)); };
    for (int i = r1 - 1; i >= r0; --i) {
        f3 xa[TRX_NAT], ga[TRX_NAT];
        float gvn[NAT3];
#pragma unroll
        for (int k = 0; k < NAT3; ++k) gvn[k] = 0.f;
        if (live) nat_load(gn + (size_t)i * NATP, gvn);
#pragma unroll
        for (int a = 0; a < TRX_NAT; ++a) {
            xa[a] = load(i, a);
            ga[a] = f3{gvn[a * 3], gvn[a * 3 + 1], gvn[a * 3 + 2]};
        }
        if (k1) {   // restraint gradient lives on N, CA, CB (absent when the group's runs do not score restraints)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                ga[a].x += G1[((size_t)i * 9 + a * 3) * LANES];
                ga[a].y += G1[((size_t)i * 9 + a * 3 + 1) * LANES];
                ga[a].z += G1[((size_t)i * 9 + a * 3 + 2) * LANES];
            }
        }
        const float phi = t[(i * 3 + 0) * LANES], psi = t[(i * 3 + 1) * LANES];
        float gphi = 0.f, gpsi = 0.f;
        add(xa[TRX_AT_O], ga[TRX_AT_O]);
        gpsi = dtor(xa[TRX_AT_CA], xa[TRX_AT_C], F1, F2);
        add(xa[TRX_AT_C], ga[TRX_AT_C]);
        add(xa[TRX_AT_CB], ga[TRX_AT_CB]);
        if (i > 0) gphi = dtor(xa[TRX_AT_N], xa[TRX_AT_CA], F1, F2);
        add(xa[TRX_AT_CA], ga[TRX_AT_CA]);
        if (i > 0) {
            f3 Cp = load(i - 1, TRX_AT_C);
            const float go = dtor(Cp, xa[TRX_AT_N], F1, F2);
            const float om = t[((i - 1) * 3 + 2) * LANES];   // omega(i-1): its tether term lives here
            float dev = om - (float)TRX_PI;
            dev -= 2.0f * (float)TRX_PI * floorf((dev + (float)TRX_PI) / (2.0f * (float)TRX_PI));
            const float deg = dev * (float)(1.0 / TRX_DEG);
            e_omega += (double)((float)TRX_OMEGA_K * deg * deg);
            if (live) gt[((i - 1) * 3 + 2) * LANES] = go + w_omega * 2.0f * (float)TRX_OMEGA_K * deg * (float)(1.0 / TRX_DEG);
        }
        add(xa[TRX_AT_N], ga[TRX_AT_N]);
        if (i > 0 && i < L - 1) {   // Ramachandran, termini skipped
            const int cls = s.aa[i] == TRX_AA_PRO ? 1 : 0;
            float P = (float)TRX_RAMA_FLOOR, dPphi = 0.f, dPpsi = 0.f;
#pragma unroll
            for (int k = 0; k < TRX_RAMA_NB; ++k) {
                const float *b = c_model.rama[cls][k];
                const float dphi = phi - b[0] * (float)TRX_DEG, dpsi = psi - b[1] * (float)TRX_DEG;
                float s1, c1, s2, c2;
                sincosf(dphi, &s1, &c1);
                sincosf(dpsi, &s2, &c2);
                const float e = b[4] * expf(b[2] * (c1 - 1.0f) + b[3] * (c2 - 1.0f));
                P += e;
                dPphi -= e * b[2] * s1;
                dPpsi -= e * b[3] * s2;
            }
            e_rama += (double)(-logf(P) - c_model.rama_off[cls]);
            gphi -= w_rama * dPphi / P;
            gpsi -= w_rama * dPpsi / P;
        }
        if (live) {
            gt[(i * 3 + 0) * LANES] = i == 0 ? 0.f : gphi;
            gt[(i * 3 + 1) * LANES] = gpsi;
        }
    }
    tot[warp][0][lane] = F1.x; tot[warp][1][lane] = F1.y; tot[warp][2][lane] = F1.z;
    tot[warp][3][lane] = F2.x; tot[warp][4][lane] = F2.y; tot[warp][5][lane] = F2.z;
    esum[warp][0][lane] = e_rama; esum[warp][1][lane] = e_omega;
    __syncthreads();
    // everything downstream of this segment
    f3 D1 = {0.f, 0.f, 0.f}, D2 = {0.f, 0.f, 0.f};
    for (int w2 = warp + 1; w2 < SEG_WARPS; ++w2) {
        D1 = D1 + f3{tot[w2][0][lane], tot[w2][1][lane], tot[w2][2][lane]};
        D2 = D2 + f3{tot[w2][3][lane], tot[w2][4][lane], tot[w2][5][lane]};
    }
    if (live && r1 < L) {
        for (int i = r0; i < r1; ++i) {
            const f3 N = load(i, TRX_AT_N), CA = load(i, TRX_AT_CA), C = load(i, TRX_AT_C);
            gt[(i * 3 + 1) * LANES] += dtor(CA, C, D1, D2);
            if (i > 0) {
                gt[(i * 3 + 0) * LANES] += dtor(N, CA, D1, D2);
                gt[((i - 1) * 3 + 2) * LANES] += dtor(load(i - 1, TRX_AT_C), N, D1, D2);
            }
        }
    }
    if (warp == 0 && live) {
        gt[((L - 1) * 3 + 2) * LANES] = 0.f;
        double er = 0.0, eo = 0.0;
        for (int w2 = 0; w2 < SEG_WARPS; ++w2) { er += esum[w2][0][lane]; eo += esum[w2][1][lane]; }
        double term[TRX_NTERM];
        term[TRX_T_APC] = k1 ? s.E3[0 * (size_t)Npad + n] : 0.0;
        term[TRX_T_DIH] = k1 ? s.E3[1 * (size_t)Npad + n] : 0.0;
        term[TRX_T_ANG] = k1 ? s.E3[2 * (size_t)Npad + n] : 0.0;
        term[TRX_T_VDW] = s.Evdw[n];
        term[TRX_T_RAMA] = er;
        term[TRX_T_OMEGA] = eo;
        term[TRX_T_CART] = 0.0;   // ideal internal geometry in torsion space
        term[TRX_T_HB] = s.Ehb[n];
        double total = 0.0;
#pragma unroll
        for (int k = 0; k < TRX_NTERM; ++k) {
            s.terms[(size_t)k * Npad + dec] = term[k];
            total += (double)s.wslot[(size_t)k * Npad + n] * term[k];
        }
        s.ft[dec] = total;
    }
}


// ---- Cartesian stage (min_mover_cart, folding.py:100-102,170) ------------------------------
// The degrees of freedom are the coordinates themselves, x = [Lpad][15] per decoy in the
// layout of X.  Evaluation = gather (below) -> K1 -> vdw -> cart_grad_kernel.
__device__ __forceinline__ void axpy(f3 &a, float s, f3 b) { a.x += s * b.x; a.y += s * b.y; a.z += s * b.z; }

// Trial coordinates of the decoy in each live slot -> X (slot-grouped, for K1) and xnat.
__global__ void __launch_bounds__(256) cart_gather_kernel(FoldState s)
{
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, n = g * LANES + lane;
    if (!s.gslot[g]) return;
    const int dec = s.perm[n];
    const bool live = dec >= 0;
    const int dn = live ? dec : 0;
    const float *__restrict__ xt = s.xt + (size_t)(dn / LANES) * s.ndof * LANES + dn % LANES;
    float *__restrict__ X = s.X + (size_t)g * s.Lpad * NAT3 * LANES + lane;
    float *__restrict__ xn = s.xnat + (size_t)n * s.L * NATP;
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < TRX_NTERM; ++k) s.wslot[(size_t)k * s.Npad + n] = live ? s.wl[(size_t)k * s.Npad + dec] : 0.f;
    }
    for (int i = warp; i < s.L; i += nw) {
        float v[NAT3];
#pragma unroll
        for (int k = 0; k < NAT3; ++k) {
            v[k] = xt[((size_t)i * NAT3 + k) * LANES];
            X[((size_t)i * NAT3 + k) * LANES] = v[k];
        }
        if (live) nat_store(xn + (size_t)i * NATP, v);
    }
}

// |a-b| spring: returns (d-d0)^2, adds f * d(d)/dx with f = wk2 (d-d0)
__device__ __forceinline__ float cart_bond(f3 a, f3 b, float d0, float wk2, f3 &ga, f3 &gb)
{
    const f3 d = a - b;
    const float len = sqrtf(dot(d, d)), dev = len - d0, f = wk2 * dev / len;
    axpy(ga, f, d);
    axpy(gb, -f, d);
    return dev * dev;
}

// angle a-b-c spring (vertex b)
__device__ __forceinline__ float cart_angle(f3 a, f3 b, f3 c, float t0, float wk2, f3 &ga, f3 &gb, f3 &gc)
{
    const f3 v = a - b, w = c - b;
    const float inv = 1.0f / sqrtf(dot(v, v)), inw = 1.0f / sqrtf(dot(w, w));
    const f3 vu = inv * v, wu = inw * w;
    const float cs = dot(vu, wu);
    const f3 cr = cross(vu, wu);
    const float sn = sqrtf(dot(cr, cr));
    const float dev = atan2f(sn, cs) - t0, f = wk2 * dev;
    const f3 g1 = (-inv / sn) * (wu - cs * vu), g3 = (-inw / sn) * (vu - cs * wu);
    axpy(ga, f, g1);
    axpy(gc, f, g3);
    axpy(gb, -f, g1 + g3);
    return dev * dev;
}

// dihedral p1-p2-p3-p4 (IUPAC sign, as oracle trxo_dihedral) and its gradient (Blondel & Karplus)
__device__ __forceinline__ float cart_dihedral(f3 p1, f3 p2, f3 p3, f3 p4, f3 &d1, f3 &d2, f3 &d3, f3 &d4)
{
    const f3 F = p1 - p2, G = p2 - p3, H = p4 - p3;
    const f3 A = cross(F, G), B = cross(H, G);
    const float A2 = dot(A, A), B2 = dot(B, B), Gn = sqrtf(dot(G, G)), FG = dot(F, G), HG = dot(H, G);
    const float iA = 1.0f / A2, iB = 1.0f / B2, iG = 1.0f / Gn;
    d1 = (-Gn * iA) * A;
    d4 = (Gn * iB) * B;
    const f3 t = (FG * iA * iG) * A - (HG * iB * iG) * B;
    d2 = (Gn * iA) * A + t;
    d3 = (-Gn * iB) * B - t;
    return atan2f(Gn * dot(A, H), dot(A, B));
}

// Cartesian-mode gradient assembly.  One CTA per slot group (lane = decoy), SEG_WARPS warps:
// warp w owns residues [r0, r1).  Step i evaluates every term ANCHORED at residue i -- its
// own springs, the peptide link (i, i+1) and Ramachandran(i) -- which touch residues i-1
// (C only), i and i+1 (N, CA only).  A warp also replays the steps r0-1 and r1 of its
// neighbours (without counting their energy), so every residue it owns is completed in
// registers: no atomics, no halo exchange, fixed summation order.
// gt = restraint gradient (K1) + vdw gradient + these terms; also the terms and the total.
constexpr int CART_WARPS = 16;
constexpr int CART_THREADS = CART_WARPS * 32;
__global__ void __launch_bounds__(CART_THREADS) cart_grad_kernel(FoldState s)
{
    __shared__ double esum[CART_WARPS][3][LANES];
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;   // n = slot
    if (!s.gslot[g]) return;
    const int dec = s.perm[n];
    const bool live = dec >= 0;
    const int dn = live ? dec : 0;
    const int L = s.L, Npad = s.Npad;
    const float *__restrict__ X = s.X + (size_t)g * s.Lpad * NAT3 * LANES + lane;
    const float *__restrict__ G1 = s.gk1 + (size_t)g * s.Lpad * 9 * LANES + lane;
    const float *__restrict__ gn = s.gnat + (size_t)n * L * NATP;
    float *__restrict__ gt = s.gt + (size_t)(dn / LANES) * s.ndof * LANES + dn % LANES;
    const float w_cart = s.wslot[(size_t)TRX_T_CART * Npad + n], w_rama = s.wslot[(size_t)TRX_T_RAMA * Npad + n];
    const float w_omega = s.wslot[(size_t)TRX_T_OMEGA * Npad + n];
    const float kb2 = w_cart * 2.0f * (float)TRX_CART_KB, ka2 = w_cart * 2.0f * (float)TRX_CART_KA;
    const bool k1 = s.gneedk1[g] != 0;
    const int Lseg = (L + CART_WARPS - 1) / CART_WARPS;
    const int r0 = warp * Lseg, r1 = min(L, r0 + Lseg);
    auto load = [&](int i, int a) -> f3 { return {X[((size_t)i * NAT3 + a * 3) * LANES], X[((size_t)i * NAT3 + a * 3 + 1) * LANES], X[((size_t)i * NAT3 + a * 3 + 2) * LANES]}; };
    const f3 zero = {0.f, 0.f, 0.f};
    double e_cart = 0.0, e_rama = 0.0, e_omega = 0.0;
    if (r0 < L) {
        f3 gp[TRX_NAT], gc[TRX_NAT], gx[2];   // gradient of residue i-1, i, and of N/CA of i+1
#pragma unroll
        for (int a = 0; a < TRX_NAT; ++a) { gp[a] = zero; gc[a] = zero; }
        gx[0] = zero; gx[1] = zero;
        auto store = [&](int i, const f3 *gr) {
            if (!live) return;
            float v[NAT3];
#pragma unroll
            for (int a = 0; a < TRX_NAT; ++a) { v[a * 3] = gr[a].x; v[a * 3 + 1] = gr[a].y; v[a * 3 + 2] = gr[a].z; }
            float gvn[NAT3];
            nat_load(gn + (size_t)i * NATP, gvn);
#pragma unroll
            for (int k = 0; k < NAT3; ++k) v[k] += gvn[k];
            if (k1) {
#pragma unroll
                for (int k = 0; k < 9; ++k) v[k] += G1[((size_t)i * 9 + k) * LANES];
            }
#pragma unroll
            for (int k = 0; k < NAT3; ++k) gt[((size_t)i * NAT3 + k) * LANES] = v[k];
        };
        const int i0 = max(r0 - 1, 0), i1 = min(r1, L - 1);
        f3 Cp = i0 > 0 ? load(i0 - 1, TRX_AT_C) : zero;
        f3 N = load(i0, TRX_AT_N), CA = load(i0, TRX_AT_CA);
        for (int i = i0; i <= i1; ++i) {
            const bool own = i >= r0 && i < r1;
            const f3 CB = load(i, TRX_AT_CB), C = load(i, TRX_AT_C), O = load(i, TRX_AT_O);
            float ec = 0.f;
            ec += (float)TRX_CART_KB * cart_bond(CA, N, (float)TRX_B_N_CA, kb2, gc[TRX_AT_CA], gc[TRX_AT_N]);
            ec += (float)TRX_CART_KB * cart_bond(C, CA, (float)TRX_B_CA_C, kb2, gc[TRX_AT_C], gc[TRX_AT_CA]);
            ec += (float)TRX_CART_KB * cart_bond(O, C, (float)TRX_B_C_O, kb2, gc[TRX_AT_O], gc[TRX_AT_C]);
            ec += (float)TRX_CART_KA * cart_angle(N, CA, C, (float)TRX_A_N_CA_C, ka2, gc[TRX_AT_N], gc[TRX_AT_CA], gc[TRX_AT_C]);
            ec += (float)TRX_CART_KA * cart_angle(CA, C, O, (float)TRX_A_CA_C_O, ka2, gc[TRX_AT_CA], gc[TRX_AT_C], gc[TRX_AT_O]);
            {   // CB tether to the virtual-CB position
                const f3 b = CA - N, c = C - CA, a = cross(b, c);
                const f3 r = CB - ((float)TRX_CB_A * a + (float)TRX_CB_B * b + (float)TRX_CB_C * c + CA);
                ec += (float)TRX_CART_KCB * dot(r, r);
                const f3 q = (-w_cart * 2.0f * (float)TRX_CART_KCB) * r;   // dE/d vCB
                const f3 gb = (float)TRX_CB_A * cross(c, q) + (float)TRX_CB_B * q;
                const f3 gcv = (float)TRX_CB_A * cross(q, b) + (float)TRX_CB_C * q;
                axpy(gc[TRX_AT_CB], -1.0f, q);
                axpy(gc[TRX_AT_N], -1.0f, gb);
                axpy(gc[TRX_AT_CA], 1.0f, gb - gcv + q);
                axpy(gc[TRX_AT_C], 1.0f, gcv);
            }
            f3 N1 = zero, CA1 = zero;
            float eo = 0.f, er = 0.f;
            if (i < L - 1) {
                N1 = load(i + 1, TRX_AT_N); CA1 = load(i + 1, TRX_AT_CA);
                ec += (float)TRX_CART_KB * cart_bond(N1, C, (float)TRX_B_C_N, kb2, gx[0], gc[TRX_AT_C]);
                ec += (float)TRX_CART_KA * cart_angle(CA, C, N1, (float)TRX_A_CA_C_N, ka2, gc[TRX_AT_CA], gc[TRX_AT_C], gx[0]);
                ec += (float)TRX_CART_KA * cart_angle(O, C, N1, (float)TRX_A_O_C_N, ka2, gc[TRX_AT_O], gc[TRX_AT_C], gx[0]);
                ec += (float)TRX_CART_KA * cart_angle(C, N1, CA1, (float)TRX_A_C_N_CA, ka2, gc[TRX_AT_C], gx[0], gx[1]);
                {   // carbonyl O in the peptide plane
                    const f3 u = CA - C, v = N1 - C, o = O - C;
                    const f3 uv = cross(u, v), vo = cross(v, o), ou = cross(o, u);
                    const float t = dot(uv, o), f = w_cart * 2.0f * (float)TRX_CART_KPL * t;
                    ec += (float)TRX_CART_KPL * t * t;
                    axpy(gc[TRX_AT_CA], f, vo);
                    axpy(gx[0], f, ou);
                    axpy(gc[TRX_AT_O], f, uv);
                    axpy(gc[TRX_AT_C], -f, vo + ou + uv);
                }
                {   // omega tether
                    f3 d1, d2, d3, d4;
                    const float om = cart_dihedral(CA, C, N1, CA1, d1, d2, d3, d4);
                    float dev = om - (float)TRX_PI;
                    dev -= 2.0f * (float)TRX_PI * floorf((dev + (float)TRX_PI) / (2.0f * (float)TRX_PI));
                    const float deg = dev * (float)(1.0 / TRX_DEG);
                    eo = (float)TRX_OMEGA_K * deg * deg;
                    const float f = w_omega * 2.0f * (float)TRX_OMEGA_K * deg * (float)(1.0 / TRX_DEG);
                    axpy(gc[TRX_AT_CA], f, d1); axpy(gc[TRX_AT_C], f, d2); axpy(gx[0], f, d3); axpy(gx[1], f, d4);
                }
                if (i > 0) {   // Ramachandran, termini skipped
                    f3 a1, a2, a3, a4, b1, b2, b3, b4;
                    const float phi = cart_dihedral(Cp, N, CA, C, a1, a2, a3, a4);
                    const float psi = cart_dihedral(N, CA, C, N1, b1, b2, b3, b4);
                    const int cls = s.aa[i] == TRX_AA_PRO ? 1 : 0;
                    float P = (float)TRX_RAMA_FLOOR, dPphi = 0.f, dPpsi = 0.f;
#pragma unroll
                    for (int k = 0; k < TRX_RAMA_NB; ++k) {
                        const float *b = c_model.rama[cls][k];
                        const float dphi = phi - b[0] * (float)TRX_DEG, dpsi = psi - b[1] * (float)TRX_DEG;
                        float s1, c1, s2, c2;
                        sincosf(dphi, &s1, &c1);
                        sincosf(dpsi, &s2, &c2);
                        const float e = b[4] * expf(b[2] * (c1 - 1.0f) + b[3] * (c2 - 1.0f));
                        P += e;
                        dPphi -= e * b[2] * s1;
                        dPpsi -= e * b[3] * s2;
                    }
                    er = -logf(P) - c_model.rama_off[cls];
                    const float fphi = -w_rama * dPphi / P, fpsi = -w_rama * dPpsi / P;
                    axpy(gp[TRX_AT_C], fphi, a1); axpy(gc[TRX_AT_N], fphi, a2); axpy(gc[TRX_AT_CA], fphi, a3); axpy(gc[TRX_AT_C], fphi, a4);
                    axpy(gc[TRX_AT_N], fpsi, b1); axpy(gc[TRX_AT_CA], fpsi, b2); axpy(gc[TRX_AT_C], fpsi, b3); axpy(gx[0], fpsi, b4);
                }
            }
            if (own) { e_cart += (double)ec; e_rama += (double)er; e_omega += (double)eo; }
            if (i - 1 >= r0) store(i - 1, gp);   // residue i-1 is complete (i-1 < r1 holds inside the loop)
#pragma unroll
            for (int a = 0; a < TRX_NAT; ++a) { gp[a] = gc[a]; gc[a] = zero; }
            gc[TRX_AT_N] = gx[0]; gc[TRX_AT_CA] = gx[1];
            gx[0] = zero; gx[1] = zero;
            Cp = C; N = N1; CA = CA1;
        }
        if (i1 >= r0 && i1 < r1) store(i1, gp);   // the chain's last residue
    }
    // padded residues never move
    if (live) for (int k = L * NAT3 + warp; k < s.Lpad * NAT3; k += CART_WARPS) gt[(size_t)k * LANES] = 0.f;
    esum[warp][0][lane] = e_cart; esum[warp][1][lane] = e_rama; esum[warp][2][lane] = e_omega;
    __syncthreads();
    if (warp == 0 && live) {
        double ec = 0.0, er = 0.0, eo = 0.0;
        for (int w2 = 0; w2 < CART_WARPS; ++w2) { ec += esum[w2][0][lane]; er += esum[w2][1][lane]; eo += esum[w2][2][lane]; }
        double term[TRX_NTERM];
        term[TRX_T_APC] = k1 ? s.E3[0 * (size_t)Npad + n] : 0.0;
        term[TRX_T_DIH] = k1 ? s.E3[1 * (size_t)Npad + n] : 0.0;
        term[TRX_T_ANG] = k1 ? s.E3[2 * (size_t)Npad + n] : 0.0;
        term[TRX_T_VDW] = s.Evdw[n];
        term[TRX_T_RAMA] = er;
        term[TRX_T_OMEGA] = eo;
        term[TRX_T_CART] = ec;
        term[TRX_T_HB] = s.Ehb[n];
        double total = 0.0;
#pragma unroll
        for (int k = 0; k < TRX_NTERM; ++k) {
            s.terms[(size_t)k * Npad + dec] = term[k];
            total += (double)s.wslot[(size_t)k * Npad + n] * term[k];
        }
        s.ft[dec] = total;
    }
}

// Backbone torsions of residue i read back from coordinates (load(i, atom) -> f3): what a decoy keeps of its
// Cartesian run when a torsion-space run rebuilds it with ideal bond geometry.
template <typename LoadFn>
__device__ __forceinline__ void readback_torsions(LoadFn load, int i, int L, float &phi, float &psi, float &omg)
{
    const f3 N = load(i, TRX_AT_N), CA = load(i, TRX_AT_CA), C = load(i, TRX_AT_C);
    f3 d1, d2, d3, d4;
    phi = (float)TRX_PI; omg = (float)TRX_PI;
    if (i > 0) phi = cart_dihedral(load(i - 1, TRX_AT_C), N, CA, C, d1, d2, d3, d4);
    if (i < L - 1) {
        const f3 N1 = load(i + 1, TRX_AT_N), CA1 = load(i + 1, TRX_AT_CA);
        psi = cart_dihedral(N, CA, C, N1, d1, d2, d3, d4);
        omg = cart_dihedral(CA, C, N1, CA1, d1, d2, d3, d4);
    } else {
        psi = cart_dihedral(N, CA, C, load(i, TRX_AT_O), d1, d2, d3, d4) - (float)TRX_PI;   // NeRF places O at psi + pi
        if (psi <= -(float)TRX_PI) psi += 2.0f * (float)TRX_PI;
    }
}

// Parity entry (trx_fold_eval_cart): torsions of the coordinates just evaluated (identity slots) -> x.
__global__ void __launch_bounds__(256) cart_readback_kernel(FoldState s)
{
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, n = g * LANES + lane;
    if (n >= s.N) return;
    const int L = s.L;
    const float *__restrict__ X = s.X + (size_t)g * s.Lpad * NAT3 * LANES + lane;
    float *__restrict__ x = s.x + (size_t)g * s.ndof_t * LANES + lane;
    auto load = [&](int i, int a) -> f3 { return {X[((size_t)i * NAT3 + a * 3) * LANES], X[((size_t)i * NAT3 + a * 3 + 1) * LANES], X[((size_t)i * NAT3 + a * 3 + 2) * LANES]}; };
    for (int i = warp; i < L; i += nw) {
        float phi, psi, omg;
        readback_torsions(load, i, L, phi, psi, omg);
        x[(size_t)(i * 3 + 0) * LANES] = phi;
        x[(size_t)(i * 3 + 1) * LANES] = psi;
        x[(size_t)(i * 3 + 2) * LANES] = omg;
    }
}

// ---- Monte-Carlo extension (no reference behaviour: BASELINE config 4 / SURVEY 8a row 16).
// A cycle = perturb a block of consecutive residues' phi/psi, re-minimise through the schedule's last run,
// Metropolis accept/reject on that run's weighted score.  It is part of the per-decoy state machine (a decoy
// that finishes its minimisation is judged, perturbed and restarted in the same round: no batch-wide barrier
// per cycle); the counter-based generator is keyed by (seed, global decoy id, cycle, draw), so a decoy's
// trajectory does not depend on the batch, the position or the GPU it sits in.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(unsigned long long seed, unsigned long long id, unsigned cycle, unsigned draw)
{
    const unsigned long long h = mix64(mix64(seed ^ mix64(id)) ^ ((unsigned long long)cycle << 32 | draw));
    return ((h >> 40) + 0.5f) * (1.0f / 16777216.0f);
}
// index of queue entry q in the caller's arrays (queue ids are block-major with every block 32-aligned)
__device__ __forceinline__ int caller_index(const FoldState &s, int q)
{
    int t = 0;
    while (t + 1 < s.ntab && q >= s.qd0[t + 1]) ++t;
    return s.qc0[t] + (q - s.qd0[t]);
}
// phi/psi element k of a decoy after the perturbation of cycle `cycle`
__device__ __noinline__ float mc_perturb(const McOpts o, int L, unsigned long long id, int cycle, int k, float v)
{
    const int blk = o.block_min + (int)(u01(o.seed, id, cycle, 0) * (o.block_max - o.block_min + 1));
    const int len = min(max(blk, 1), L - 2);
    const int start = 1 + (int)(u01(o.seed, id, cycle, 1) * (L - 1 - len));
    const int res = k / 3, t = k % 3;
    if (t < 2 && res >= start && res < start + len) {
        // Box-Muller from two counter-based uniforms
        const float a = u01(o.seed, id, cycle, 2 + 2 * k), b = u01(o.seed, id, cycle, 3 + 2 * k);
        return v + o.sigma * sqrtf(-2.0f * logf(a)) * cospif(2.0f * b);
    }
    return v;
}

// Start of a segment: every position is empty; the first turnover fills them from the head of the queue.
__global__ void seg_reset_kernel(FoldState s)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 16) { s.qcursor[n] = 0; s.qocc[n] = 0; }
    if (n >= s.Npad) return;
    s.status[n] = ST_DONE;
    s.orig[n] = -1;
    s.held[n] = 0;
}

// K5: batched L-BFGS with non-monotone Armijo back-tracking (lane = decoy).  Consumes the
// evaluation of the trial point (ft, gt) and produces the next trial point.
//
// The two-loop recursion is done on Gram matrices: per decoy we keep SY[i][j] = s_i.y_j and
// YY[i][j] = y_i.y_j of the stored pairs.  One streaming pass over the history computes every
// dot product the step needs (s_j.g, y_j.g and the new pair's products with all stored pairs:
// 5 accumulators per slot, all loads independent), the m-dimensional recursion then runs on
// scalars, and a second streaming pass forms d = c_g g + sum a_j s_j + sum b_j y_j and the
// trial point.  2 bandwidth-bound sweeps instead of 4m latency-bound dependent ones.
//
// Three launches per round so that a decoy group is streamed by SEVERAL CTAs (one CTA per
// group leaves most of the HBM bandwidth idle: a group's history is 4.6 MB per sweep in
// torsion space, 23 MB in Cartesian space):
//   lbfgs_dots_kernel    grid (G, nch): sweep A -> partial sums of the 16 fixed chunks of the vector
//   lbfgs_step_kernel    grid  G      : sums the partials in chunk order (deterministic and the same
//                                       in any batch), step logic on scalars, coefficients of the new direction
//   lbfgs_update_kernel  grid (G, nch): sweep B over a share of the vector
// History layout [G][ndof][m][32]: the m slots of one vector element are contiguous, so a
// sweep reads each group's history as one sequential stream.
constexpr int LB_WARPS = 8;
constexpr int LB_THREADS = LB_WARPS * 32;
constexpr float LS_SIGMA = 0.1f;
constexpr int LS_MAXBACK = 20;
constexpr int LB_NSCAL = 8;   // ss, sy, yy, gg, s.g, y.g (+2 spare)
constexpr int LB_MAXCH = 16;  // most chunks a group's vector is cut into
constexpr int LB_STEP_THREADS = 512;

template <int M>
struct LbSmem {
    static constexpr int NRED = 5 * M + LB_NSCAL;
    static constexpr int NCOEF = 2 * M + 3;   // coefS[M], coefY[M], cg, alpha, mode
    // sweep A: per-warp partial sums; the step kernel keeps the Gram matrices in the same storage
    union {
        float red[LB_WARPS][NRED][LANES];
        float gram[2][M * M][LANES];
    } u;
    float sum[NRED][LANES];      // reduced sums: YG, YYn, YSn, SG, SYn (M each) then the scalars
    float coefS[M][LANES], coefY[M][LANES];
    float cg[LANES], alpha[LANES];
    int mode[LANES];             // what sweep B does for the lane: 0 nothing, 1 new direction, 2 x + alpha d, 3 xt = x
};

// What the evaluation just made means for decoy n: 0 none, 1 start run here (steepest descent),
// 2 accepted step, 3 rejected step, 4 run skipped (clash check), 5 the closing evaluation of a decoy that
// has left the segment.  Pure function of the per-decoy scalars, so the dots kernel and the step kernel
// agree on it.
__device__ __forceinline__ int lb_action(const FoldState &s, int n, int status)
{
    const int Npad = s.Npad;
    if (status == ST_INIT) {
        const Run &r = s.runs[s.run[n]];
        bool skip = false;
        if (r.clash_check) {   // a decoy holding Cartesian coordinates is judged on those
            const double *tv = s.held[n] ? s.theld : s.terms;
            skip = (float)(tv[(size_t)TRX_T_VDW * Npad + n] + tv[(size_t)TRX_T_RAMA * Npad + n]) < r.clash_thr;
        }
        return skip ? 4 : 1;
    }
    if (status == ST_LS) {
        const double ft = s.ft[n];
        const int nmem = s.nmem[n];
        double fref = s.fmem[n];
        for (int q = 1; q < min(nmem, 3); ++q) fref = fmax(fref, s.fmem[(size_t)q * Npad + n]);
        return (isfinite(ft) && ft <= fref + (double)(LS_SIGMA * s.alpha[n] * s.slope[n])) ? 2 : 3;
    }
    if (status == ST_FINAL) return 5;
    return 0;
}

template <int M>
__global__ void __launch_bounds__(LB_THREADS, 1) lbfgs_dots_kernel(FoldState s)
{
    extern __shared__ __align__(16) unsigned char lb_raw[];
    LbSmem<M> &sm = *reinterpret_cast<LbSmem<M> *>(lb_raw);
    constexpr int NRED = LbSmem<M>::NRED;
    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;
    if (!s.gactive[g]) return;
    const int nd = s.ndof, m = s.m;
    const size_t vb = (size_t)g * nd * LANES + lane;
    float *__restrict__ x = s.x + vb, *__restrict__ gv = s.g + vb;
    const float *__restrict__ xt = s.xt + vb, *__restrict__ gt = s.gt + vb;
    float *__restrict__ S = s.S + (size_t)g * m * nd * LANES + lane, *__restrict__ Y = s.Y + (size_t)g * m * nd * LANES + lane;
    const int status = n < s.N ? s.status[n] : ST_DONE;
    const int action = lb_action(s, n, status), head = s.head[n];
    const bool is_acc = action == 2, is_new = action == 1 || action == 2;
    // The vector is ALWAYS cut into LB_MAXCH chunks and each chunk's sums are formed by the 8 warps of one CTA in a
    // fixed order; only the number of CTAs that share the chunks (gridDim.y) follows the batch size.  The sums a
    // decoy sees are therefore the same bits in any batch.
    const int per = (nd + LB_MAXCH - 1) / LB_MAXCH;
    for (int ch = blockIdx.y; ch < LB_MAXCH; ch += gridDim.y) {
        const int k0 = ch * per, k1 = min(nd, k0 + per);
        float aYG[M], aYY[M], aYS[M], aSG[M], aSY[M], sc[LB_NSCAL];
#pragma unroll
        for (int j = 0; j < M; ++j) { aYG[j] = 0.f; aYY[j] = 0.f; aYS[j] = 0.f; aSG[j] = 0.f; aSY[j] = 0.f; }
#pragma unroll
        for (int j = 0; j < LB_NSCAL; ++j) sc[j] = 0.f;
        for (int k = k0 + warp; k < k1; k += LB_WARPS) {
            const float xk = x[(size_t)k * LANES], xtk = xt[(size_t)k * LANES], gk = gv[(size_t)k * LANES], gtk = gt[(size_t)k * LANES];
            const float sn = xtk - xk, yn = gtk - gk, gn = is_new ? gtk : gk;
            float yj[M], sj[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                yj[j] = j < m ? Y[((size_t)k * m + j) * LANES] : 0.f;
                sj[j] = j < m ? S[((size_t)k * m + j) * LANES] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                aYG[j] += yj[j] * gn; aYY[j] += yj[j] * yn; aYS[j] += yj[j] * sn;
                aSG[j] += sj[j] * gn; aSY[j] += sj[j] * yn;
            }
            sc[0] += sn * sn; sc[1] += sn * yn; sc[2] += yn * yn; sc[3] += gn * gn; sc[4] += sn * gn; sc[5] += yn * gn;
            if (is_acc) {
                S[((size_t)k * m + head) * LANES] = sn;
                Y[((size_t)k * m + head) * LANES] = yn;
            }
            if (is_new) {
                x[(size_t)k * LANES] = xtk;
                gv[(size_t)k * LANES] = gtk;
            }
        }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            sm.u.red[warp][j][lane] = aYG[j]; sm.u.red[warp][M + j][lane] = aYY[j]; sm.u.red[warp][2 * M + j][lane] = aYS[j];
            sm.u.red[warp][3 * M + j][lane] = aSG[j]; sm.u.red[warp][4 * M + j][lane] = aSY[j];
        }
#pragma unroll
        for (int j = 0; j < LB_NSCAL; ++j) sm.u.red[warp][5 * M + j][lane] = sc[j];
        __syncthreads();
        float *__restrict__ part = s.lbpart + ((size_t)g * LB_MAXCH + ch) * NRED * LANES;
        for (int e = threadIdx.x; e < NRED * LANES; e += LB_THREADS) {
            const int idx = e / LANES, l = e % LANES;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < LB_WARPS; ++w) t += sm.u.red[w][idx][l];
            part[e] = t;
        }
        __syncthreads();   // the next chunk reuses the buffer
    }
}

template <int M>
__global__ void __launch_bounds__(LB_STEP_THREADS, 1) lbfgs_step_kernel(FoldState s, int nch)
{
    extern __shared__ __align__(16) unsigned char lb_raw[];
    LbSmem<M> &sm = *reinterpret_cast<LbSmem<M> *>(lb_raw);
    constexpr int NRED = LbSmem<M>::NRED;
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;
    if (!s.gactive[g]) return;
    const int m = s.m, Npad = s.Npad;
    float *__restrict__ gram = s.gram + (size_t)g * 2 * M * M * LANES;
    {   // partial sums of the chunks, in chunk order
        const float *__restrict__ part = s.lbpart + (size_t)g * LB_MAXCH * NRED * LANES;
        for (int e = threadIdx.x; e < NRED * LANES; e += LB_STEP_THREADS) {
            float t = 0.f;
            for (int c = 0; c < nch; ++c) t += part[(size_t)c * NRED * LANES + e];
            (&sm.sum[0][0])[e] = t;
        }
    }
    // Gram matrices into shared memory (147 KB per group at M = 24: 16-byte copies by 512 threads; with
    // 4-byte copies by 256 threads this staging was the longest part of the kernel, ~90 us per round)
    {
        const float4 *__restrict__ src = reinterpret_cast<const float4 *>(gram);
        float4 *dst = reinterpret_cast<float4 *>(&sm.u.gram[0][0][0]);
        for (int e = threadIdx.x; e < 2 * M * M * LANES / 4; e += LB_STEP_THREADS) dst[e] = src[e];
    }
    __syncthreads();

    // ---- per-decoy decision
    int status = n < s.N ? s.status[n] : ST_DONE;
    int run = s.run[n], hist = s.hist[n], head = s.head[n], iter = s.iter[n], bt = s.bt[n], restart = s.restart[n], nmem = s.nmem[n];
    double f = s.f[n];
    float alpha = s.alpha[n], slope = s.slope[n];
    const double ft = s.ft[n];
    const bool fin = isfinite(ft);
    const int action = lb_action(s, n, status);

    // ---- per-decoy step logic on scalars (warp 0, one lane per decoy)
    if (warp == 0) {
        float(*SYm)[LANES] = sm.u.gram[0];
        float(*YYm)[LANES] = sm.u.gram[1];
        const float ss = sm.sum[5 * M + 0][lane], sy = sm.sum[5 * M + 1][lane], yy = sm.sum[5 * M + 2][lane];
        const float gg = sm.sum[5 * M + 3][lane], sgn = sm.sum[5 * M + 4][lane], ygn = sm.sum[5 * M + 5][lane];
        int evals = s.evals[n] + ((status == ST_INIT || status == ST_LS) ? 1 : 0), iters = s.iters[n];
        bool run_over = false, need_dir = false;
        int mode = 0, mcmode = 0;
        // the decoy moves on to run r: its weights come into force.  Beyond the segment in progress it
        // leaves the batch: one closing evaluation at its accepted point (ST_FINAL), then it is parked.
        // With Monte-Carlo cycles the end of the schedule is instead the Metropolis test on the cycle just
        // minimised, followed by the next perturbation and a restart of the last run.
        auto enter = [&](int r) {
            if (r < s.nruns)
                for (int k = 0; k < TRX_NTERM; ++k) s.wl[(size_t)k * Npad + n] = s.runs[r].w[k];
            if (r < s.seg_hi) { status = ST_INIT; return; }
            status = ST_FINAL;
            if (s.mc.cycles > 0 && r >= s.nruns && n < s.N) {
                const int c = s.mccyc[n];   // perturbations made so far
                const unsigned long long id = s.mc.id_offset + (unsigned long long)caller_index(s, s.orig[n]);
                bool keep = true;
                if (c > 0) {
                    const double fnew = f, fold = s.fsave[n];
                    keep = isfinite(fnew) && (fnew <= fold || u01(s.mc.seed, id, c - 1, 0x7fffffffu) < expf((float)((fold - fnew) / (double)s.mc.kT)));
                    if (keep) s.naccept[n] += 1;
                    else f = fold;
                }
                if (c < s.mc.cycles) {
                    s.fsave[n] = f;
                    s.mccyc[n] = c + 1;
                    s.held[n] = 0;
                    run = s.mc.mc_run;
                    for (int k = 0; k < TRX_NTERM; ++k) s.wl[(size_t)k * Npad + n] = s.runs[run].w[k];
                    status = ST_INIT; hist = 0; head = 0; iter = 0; bt = 0; restart = 1; nmem = 0;
                    mcmode = keep ? 4 : 5;   // sweep B: (x or the saved x) -> saved x, perturbed -> x, xt
                } else {
                    mcmode = keep ? 3 : 6;   // the last cycle was rejected: back to the saved x
                }
            }
        };
        if (action == 5) status = ST_DONE;   // closing evaluation made: the turnover parks the decoy
        if (action == 4) {
            run = s.runs[run].skip_to;
            enter(run);   // xt stays = x; the next round evaluates it under the new weights
        }
        if (action == 2) {
            const int h = head;
            if (sy > 1e-10f * sqrtf(ss * yy)) {
                // the new pair enters slot h: its row and column of the Gram matrices
                for (int q = 0; q < hist; ++q) {
                    const int j = (head - 1 - q + 2 * m) % m;
                    if (j == h) continue;   // the slot being overwritten (history full)
                    SYm[h * M + j][lane] = sm.sum[2 * M + j][lane];   // s_new . y_j
                    SYm[j * M + h][lane] = sm.sum[4 * M + j][lane];   // s_j . y_new
                    YYm[h * M + j][lane] = sm.sum[M + j][lane];       // y_new . y_j
                    YYm[j * M + h][lane] = sm.sum[M + j][lane];
                }
                SYm[h * M + h][lane] = sy;
                YYm[h * M + h][lane] = yy;
                sm.sum[3 * M + h][lane] = sgn;   // s_new . g_new
                sm.sum[0 * M + h][lane] = ygn;   // y_new . g_new
                head = (head + 1) % m;
                if (hist < m) hist++;
            } else if (hist == m) {
                hist = m - 1;   // the write in pass A clobbered the oldest pair: drop it
            }
            const bool conv = 2.0 * fabs(ft - f) <= (double)s.runs[run].tol * (fabs(ft) + fabs(f) + 1e-10);
            f = ft;
            s.fmem[(size_t)(nmem % 3) * Npad + n] = f;
            nmem++; iter++; iters++;
            restart = 0; bt = 0;
            if (conv || iter >= s.runs[run].max_iter) run_over = true;
            else need_dir = true;
        } else if (action == 1) {
            f = ft;
            hist = 0; head = 0; iter = 0; bt = 0; restart = 1; nmem = 1;
            s.fmem[n] = f;
            if (!s.cart && n < s.N) s.held[n] = 0;   // a torsion-space run rebuilds the chain with ideal geometry
            if (!fin) { run_over = true; if (n < s.N) s.flags[n] |= TRX_DECOY_NONFINITE; }   // cannot start from a non-finite energy
            else need_dir = true;
        } else if (action == 3) {
            bt++;
            if (bt >= LS_MAXBACK) {
                if (hist > 0) { hist = 0; head = 0; restart = 1; bt = 0; need_dir = true; }   // retry from steepest descent
                else { run_over = true; if (n < s.N) s.flags[n] |= TRX_DECOY_LINESEARCH; }
            } else {
                float q = fin ? -0.5f * slope * alpha * alpha / (float)(ft - f - (double)(slope * alpha)) : 0.1f * alpha;
                if (!(q > 0.1f * alpha)) q = 0.1f * alpha;
                if (q > 0.5f * alpha) q = 0.5f * alpha;
                alpha = q;
                mode = 2;
            }
        }
        if (run_over) {
            run++;
            enter(run);
            mode = 3;   // the run ends at x (last accepted point)
        }
        float cgv = -1.f;
#pragma unroll
        for (int j = 0; j < M; ++j) { sm.coefS[j][lane] = 0.f; sm.coefY[j][lane] = 0.f; }
        if (need_dir) {
            // two-loop recursion on scalars; q-th newest pair lives in slot (head-1-q) mod m
            // d = -r,  r = gamma (g - sum a_j y_j) + sum c_j s_j
            float sl = -gg;
            if (hist > 0) {
                const float *SGv = &sm.sum[3 * M][0], *YGv = &sm.sum[0][0];
                float al[M];
                for (int q = 0; q < hist; ++q) {
                    const int i = (head - 1 - q + 2 * m) % m;
                    float t = SGv[i * LANES + lane];
                    for (int q2 = 0; q2 < q; ++q2) {
                        const int j = (head - 1 - q2 + 2 * m) % m;
                        t -= al[q2] * SYm[i * M + j][lane];
                    }
                    al[q] = t / SYm[i * M + i][lane];
                }
                const int h0 = (head - 1 + m) % m;
                const float gamma = SYm[h0 * M + h0][lane] / YYm[h0 * M + h0][lane];
                float cc[M];
                for (int q = hist - 1; q >= 0; --q) {
                    const int i = (head - 1 - q + 2 * m) % m;
                    float t = YGv[i * LANES + lane];
                    for (int q2 = 0; q2 < hist; ++q2) {
                        const int j = (head - 1 - q2 + 2 * m) % m;
                        t -= al[q2] * YYm[i * M + j][lane];
                    }
                    t *= gamma;
                    for (int q2 = hist - 1; q2 > q; --q2) {
                        const int j = (head - 1 - q2 + 2 * m) % m;
                        t += cc[q2] * SYm[j * M + i][lane];
                    }
                    const float beta = t / SYm[i * M + i][lane];
                    cc[q] = al[q] - beta;
                }
                cgv = -gamma;
                sl = -gamma * gg;
                for (int q = 0; q < hist; ++q) {
                    const int i = (head - 1 - q + 2 * m) % m;
                    sm.coefY[i][lane] = gamma * al[q];
                    sm.coefS[i][lane] = -cc[q];
                    sl += gamma * al[q] * YGv[i * LANES + lane] - cc[q] * SGv[i * LANES + lane];
                }
            }
            if (!(sl < 0.f) || !isfinite(sl)) {   // not a descent direction: steepest descent
                hist = 0; head = 0; restart = 1;
                cgv = -1.f;
#pragma unroll
                for (int j = 0; j < M; ++j) { sm.coefS[j][lane] = 0.f; sm.coefY[j][lane] = 0.f; }
                sl = -gg;
            }
            slope = sl;
            const float gnorm = sqrtf(gg);
            alpha = restart ? fminf(1.0f, 1.0f / fmaxf(gnorm, 1e-20f)) : 1.0f;
            status = ST_LS;
            mode = 1;
            if (gg == 0.f) {   // stationary: the run is over
                run++;
                enter(run);
                mode = 3;
            }
        }
        if (mcmode) mode = mcmode;
        sm.cg[lane] = cgv;
        sm.alpha[lane] = alpha;
        sm.mode[lane] = mode;
        if (n < s.N) {
            s.status[n] = status; s.run[n] = run; s.hist[n] = hist; s.head[n] = head; s.iter[n] = iter; s.bt[n] = bt;
            s.restart[n] = restart; s.nmem[n] = nmem; s.f[n] = f; s.alpha[n] = alpha; s.slope[n] = slope;
            s.evals[n] = evals; s.iters[n] = iters;
        }
    }
    __syncthreads();
    // Gram matrices back to global memory; coefficients of sweep B
    {
        float4 *__restrict__ dst = reinterpret_cast<float4 *>(gram);
        const float4 *src = reinterpret_cast<const float4 *>(&sm.u.gram[0][0][0]);
        for (int e = threadIdx.x; e < 2 * M * M * LANES / 4; e += LB_STEP_THREADS) dst[e] = src[e];
    }
    float *__restrict__ coef = s.lbcoef + (size_t)g * LbSmem<M>::NCOEF * LANES;
    for (int e = threadIdx.x; e < M * LANES; e += LB_STEP_THREADS) {
        coef[e] = (&sm.coefS[0][0])[e];
        coef[M * LANES + e] = (&sm.coefY[0][0])[e];
    }
    if (warp == 0) {
        coef[(2 * M) * LANES + lane] = sm.cg[lane];
        coef[(2 * M + 1) * LANES + lane] = sm.alpha[lane];
        coef[(2 * M + 2) * LANES + lane] = __int_as_float(sm.mode[lane]);
    }
}

template <int M>
__global__ void __launch_bounds__(LB_THREADS, 2) lbfgs_update_kernel(FoldState s)   // <= 128 registers: two CTAs per SM keep more of the stream in flight
{
    const int g = blockIdx.x, ch = blockIdx.y, nch = gridDim.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!s.gactive[g]) return;
    const int nd = s.ndof, m = s.m;
    const int per = (nd + nch - 1) / nch, k0 = ch * per, k1 = min(nd, k0 + per);
    const size_t vb = (size_t)g * nd * LANES + lane;
    float *__restrict__ x = s.x + vb;
    const float *__restrict__ gv = s.g + vb;
    float *__restrict__ d = s.d + vb, *__restrict__ xt = s.xt + vb;
    const float *__restrict__ S = s.S + (size_t)g * m * nd * LANES + lane, *__restrict__ Y = s.Y + (size_t)g * m * nd * LANES + lane;
    const float *__restrict__ coef = s.lbcoef + (size_t)g * LbSmem<M>::NCOEF * LANES + lane;
    const int mode = __float_as_int(coef[(2 * M + 2) * LANES]);
    if (__syncthreads_or(mode >= 4)) {   // Monte-Carlo moves (rare rounds): revert and / or perturb, see lbfgs_step_kernel
        const int n = g * LANES + lane;
        if (mode >= 4) {
            float *__restrict__ xs = s.xsave + vb;
            const unsigned long long id = s.mc.id_offset + (unsigned long long)caller_index(s, s.orig[n]);
            const int cycle = s.mccyc[n] - 1;
            for (int k = k0 + warp; k < k1; k += LB_WARPS) {
                const float base = mode == 4 ? x[(size_t)k * LANES] : xs[(size_t)k * LANES];
                if (mode == 4) xs[(size_t)k * LANES] = base;
                const float nv = mode == 6 ? base : mc_perturb(s.mc, s.L, id, cycle, k, base);
                x[(size_t)k * LANES] = nv;
                xt[(size_t)k * LANES] = nv;
            }
        }
    }
    const float cgv = coef[(2 * M) * LANES], al = coef[(2 * M + 1) * LANES];
    float cS[M], cY[M];
#pragma unroll
    for (int j = 0; j < M; ++j) { cS[j] = coef[j * LANES]; cY[j] = coef[(M + j) * LANES]; }
    const int anydir = __syncthreads_or(mode == 1);
    for (int k = k0 + warp; k < k1; k += LB_WARPS) {
        const float xk = x[(size_t)k * LANES];
        if (anydir) {
            float dk = cgv * gv[(size_t)k * LANES];
            float yj[M], sj[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                yj[j] = j < m ? Y[((size_t)k * m + j) * LANES] : 0.f;
                sj[j] = j < m ? S[((size_t)k * m + j) * LANES] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) dk += cS[j] * sj[j] + cY[j] * yj[j];
            if (mode == 1) {
                d[(size_t)k * LANES] = dk;
                xt[(size_t)k * LANES] = xk + al * dk;
            }
        }
        if (mode == 2) xt[(size_t)k * LANES] = xk + al * d[(size_t)k * LANES];
        else if (mode == 3) xt[(size_t)k * LANES] = xk;
    }
}

// ---- L-BFGS sweeps through a bulk-async shared-memory ring (sm_90+/sm_100: cp.async.bulk + mbarrier) -----------
// A group's history [ndof][m][32] is one sequential stream, and with register-staged loads a sweep holds only
// what its registers can (one 255-register CTA per SM: ~40 KB in flight, ~55 % of the copy peak).  Here a
// producer warp streams S and Y -- and in sweep A the four vectors x, xt, g, gt -- into a ring of shared-memory
// stages with bulk copies (SASS UBLKCP) that complete on an mbarrier; eight consumer warps work out of shared
// memory and hand the stage back through a second mbarrier.  Nothing is staged through registers, two CTAs fit an
// SM, and ~170 KB per SM are in flight.
// Sweep A splits the HISTORY SLOTS over the warps (slot j -> warp j % 8; 5 accumulators per slot), not the vector
// elements: every warp owns complete sums over a chunk for its slots, formed in element order, so there is no
// cross-warp reduction and the sums are the same bits in any batch.  Sweep B keeps one warp per vector element
// and the order of its sum over the slots: its output is bit-identical to lbfgs_update_kernel.
constexpr int RING_CONSUMERS = 8;                          // consumer warps
constexpr int RING_THREADS = (RING_CONSUMERS + 1) * 32;    // + the producer warp
constexpr int RING_A_E = 4, RING_A_STAGES = 4;             // sweep A: vector elements per stage, stages
constexpr int RING_B_E = 8, RING_B_STAGES = 2;             // sweep B
__host__ __device__ constexpr size_t ring_a_stage_floats(int m) { return (size_t)RING_A_E * (2 * m + 4) * LANES; }
__host__ __device__ constexpr size_t ring_b_stage_floats(int m) { return (size_t)RING_B_E * 2 * m * LANES; }

template <int M>
__global__ void __launch_bounds__(RING_THREADS, 2) lbfgs_dots_ring_kernel(FoldState s)
{
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ __align__(8) unsigned long long full[RING_A_STAGES], empty[RING_A_STAGES];
    constexpr int NRED = LbSmem<M>::NRED, SPW = (M + RING_CONSUMERS - 1) / RING_CONSUMERS;
    const int g = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n = g * LANES + lane;
    if (!s.gactive[g]) return;
    const int nd = s.ndof, m = s.m;
    float *ring = reinterpret_cast<float *>(ring_raw);
    const size_t stage_f = ring_a_stage_floats(m);
    if (threadIdx.x == 0) {
        for (int k = 0; k < RING_A_STAGES; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], RING_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t vb = (size_t)g * nd * LANES;
    const float *__restrict__ gx = s.x + vb, *__restrict__ gxt = s.xt + vb, *__restrict__ gg = s.g + vb, *__restrict__ ggt = s.gt + vb;
    const float *__restrict__ gS = s.S + (size_t)g * m * nd * LANES, *__restrict__ gY = s.Y + (size_t)g * m * nd * LANES;
    const int per = (nd + LB_MAXCH - 1) / LB_MAXCH;
    if (warp == RING_CONSUMERS) {
        // ---- producer: one lane issues the bulk copies of every stage of every chunk of this CTA, in order
        if (lane == 0) {
            int it = 0;
            for (int ch = blockIdx.y; ch < LB_MAXCH; ch += gridDim.y) {
                const int k0 = ch * per, k1 = min(nd, k0 + per);
                for (int k = k0; k < k1; k += RING_A_E, ++it) {
                    const int st = it % RING_A_STAGES, ne = min(RING_A_E, k1 - k);
                    if (it >= RING_A_STAGES) mbar_wait(&empty[st], ((it / RING_A_STAGES) - 1) & 1);
                    float *dst = ring + (size_t)st * stage_f;
                    const unsigned hb = (unsigned)(ne * m * LANES * sizeof(float)), vbz = (unsigned)(ne * LANES * sizeof(float));
                    mbar_expect_tx(&full[st], 2 * hb + 4 * vbz);
                    bulk_g2s(dst, gS + (size_t)k * m * LANES, hb, &full[st]);
                    bulk_g2s(dst + (size_t)RING_A_E * m * LANES, gY + (size_t)k * m * LANES, hb, &full[st]);
                    float *vec = dst + (size_t)2 * RING_A_E * m * LANES;
                    bulk_g2s(vec, gx + (size_t)k * LANES, vbz, &full[st]);
                    bulk_g2s(vec + RING_A_E * LANES, gxt + (size_t)k * LANES, vbz, &full[st]);
                    bulk_g2s(vec + 2 * RING_A_E * LANES, gg + (size_t)k * LANES, vbz, &full[st]);
                    bulk_g2s(vec + 3 * RING_A_E * LANES, ggt + (size_t)k * LANES, vbz, &full[st]);
                }
            }
        }
        return;
    }
    // ---- consumers
    const int status = n < s.N ? s.status[n] : ST_DONE;
    const int action = lb_action(s, n, status), head = s.head[n];
    const bool is_acc = action == 2, is_new = action == 1 || action == 2;
    float *__restrict__ wx = s.x + vb + lane, *__restrict__ wg = s.g + vb + lane;
    float *__restrict__ wS = s.S + (size_t)g * m * nd * LANES + lane, *__restrict__ wY = s.Y + (size_t)g * m * nd * LANES + lane;
    int it = 0;
    for (int ch = blockIdx.y; ch < LB_MAXCH; ch += gridDim.y) {
        const int k0 = ch * per, k1 = min(nd, k0 + per);
        float aYG[SPW], aYY[SPW], aYS[SPW], aSG[SPW], aSY[SPW], sc[6];
#pragma unroll
        for (int q = 0; q < SPW; ++q) { aYG[q] = 0.f; aYY[q] = 0.f; aYS[q] = 0.f; aSG[q] = 0.f; aSY[q] = 0.f; }
#pragma unroll
        for (int q = 0; q < 6; ++q) sc[q] = 0.f;
        for (int k = k0; k < k1; k += RING_A_E, ++it) {
            const int st = it % RING_A_STAGES, ne = min(RING_A_E, k1 - k);
            mbar_wait(&full[st], (it / RING_A_STAGES) & 1);
            const float *src = ring + (size_t)st * stage_f;
            const float *Ss = src + lane, *Ys = src + (size_t)RING_A_E * m * LANES + lane;
            const float *vec = src + (size_t)2 * RING_A_E * m * LANES + lane;
            for (int e = 0; e < ne; ++e) {
                const float xk = vec[e * LANES], xtk = vec[(RING_A_E + e) * LANES], gk = vec[(2 * RING_A_E + e) * LANES], gtk = vec[(3 * RING_A_E + e) * LANES];
                const float sn = xtk - xk, yn = gtk - gk, gn = is_new ? gtk : gk;
#pragma unroll
                for (int q = 0; q < SPW; ++q) {
                    const int j = warp + RING_CONSUMERS * q;
                    if (j < m) {
                        const float yj = Ys[(e * m + j) * LANES], sj = Ss[(e * m + j) * LANES];
                        aYG[q] += yj * gn; aYY[q] += yj * yn; aYS[q] += yj * sn;
                        aSG[q] += sj * gn; aSY[q] += sj * yn;
                    }
                }
                if (warp == 0) { sc[0] += sn * sn; sc[1] += sn * yn; sc[2] += yn * yn; sc[3] += gn * gn; sc[4] += sn * gn; sc[5] += yn * gn; }
                if (warp == ((k + e) & (RING_CONSUMERS - 1))) {   // the element's write-back, spread over the warps
                    if (is_acc) {
                        wS[((size_t)(k + e) * m + head) * LANES] = sn;
                        wY[((size_t)(k + e) * m + head) * LANES] = yn;
                    }
                    if (is_new) {
                        wx[(size_t)(k + e) * LANES] = xtk;
                        wg[(size_t)(k + e) * LANES] = gtk;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        float *__restrict__ part = s.lbpart + ((size_t)g * LB_MAXCH + ch) * NRED * LANES + lane;
#pragma unroll
        for (int q = 0; q < SPW; ++q) {
            const int j = warp + RING_CONSUMERS * q;
            if (j < M) {
                part[(size_t)j * LANES] = aYG[q]; part[(size_t)(M + j) * LANES] = aYY[q]; part[(size_t)(2 * M + j) * LANES] = aYS[q];
                part[(size_t)(3 * M + j) * LANES] = aSG[q]; part[(size_t)(4 * M + j) * LANES] = aSY[q];
            }
        }
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 6; ++q) part[(size_t)(5 * M + q) * LANES] = sc[q];
            part[(size_t)(5 * M + 6) * LANES] = 0.f; part[(size_t)(5 * M + 7) * LANES] = 0.f;
        }
    }
}

template <int M>
__global__ void __launch_bounds__(RING_THREADS, 2) lbfgs_update_ring_kernel(FoldState s)
{
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ __align__(8) unsigned long long full[RING_B_STAGES], empty[RING_B_STAGES];
    const int g = blockIdx.x, ch = blockIdx.y, nch = gridDim.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (!s.gactive[g]) return;
    const int nd = s.ndof, m = s.m;
    const int per = (nd + nch - 1) / nch, k0 = ch * per, k1 = min(nd, k0 + per);
    const size_t vb = (size_t)g * nd * LANES + lane;
    float *__restrict__ x = s.x + vb;
    const float *__restrict__ gv = s.g + vb;
    float *__restrict__ d = s.d + vb, *__restrict__ xt = s.xt + vb;
    const float *__restrict__ coef = s.lbcoef + (size_t)g * LbSmem<M>::NCOEF * LANES + lane;
    const int mode = __float_as_int(coef[(2 * M + 2) * LANES]);
    float *ring = reinterpret_cast<float *>(ring_raw);
    const size_t stage_f = ring_b_stage_floats(m);
    if (threadIdx.x == 0) {
        for (int k = 0; k < RING_B_STAGES; ++k) { mbar_init(&full[k], 1); mbar_init(&empty[k], RING_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int anymc = __syncthreads_or(mode >= 4);
    const int anydir = __syncthreads_or(mode == 1);
    if (warp == RING_CONSUMERS) {
        if (lane == 0 && anydir) {
            const float *__restrict__ gS = s.S + (size_t)g * m * nd * LANES, *__restrict__ gY = s.Y + (size_t)g * m * nd * LANES;
            int it = 0;
            for (int k = k0; k < k1; k += RING_B_E, ++it) {
                const int st = it % RING_B_STAGES, ne = min(RING_B_E, k1 - k);
                if (it >= RING_B_STAGES) mbar_wait(&empty[st], ((it / RING_B_STAGES) - 1) & 1);
                float *dst = ring + (size_t)st * stage_f;
                const unsigned hb = (unsigned)(ne * m * LANES * sizeof(float));
                mbar_expect_tx(&full[st], 2 * hb);
                bulk_g2s(dst, gS + (size_t)k * m * LANES, hb, &full[st]);
                bulk_g2s(dst + (size_t)RING_B_E * m * LANES, gY + (size_t)k * m * LANES, hb, &full[st]);
            }
        }
        return;
    }
    if (anymc) {   // Monte-Carlo moves (rare rounds): revert and / or perturb, see lbfgs_step_kernel
        const int n = g * LANES + lane;
        if (mode >= 4) {
            float *__restrict__ xs = s.xsave + vb;
            const unsigned long long id = s.mc.id_offset + (unsigned long long)caller_index(s, s.orig[n]);
            const int cycle = s.mccyc[n] - 1;
            for (int k = k0 + warp; k < k1; k += RING_CONSUMERS) {
                const float base = mode == 4 ? x[(size_t)k * LANES] : xs[(size_t)k * LANES];
                if (mode == 4) xs[(size_t)k * LANES] = base;
                const float nv = mode == 6 ? base : mc_perturb(s.mc, s.L, id, cycle, k, base);
                x[(size_t)k * LANES] = nv;
                xt[(size_t)k * LANES] = nv;
            }
        }
    }
    const float cgv = coef[(2 * M) * LANES], al = coef[(2 * M + 1) * LANES];
    if (anydir) {
        float cS[M], cY[M];
#pragma unroll
        for (int j = 0; j < M; ++j) { cS[j] = coef[j * LANES]; cY[j] = coef[(M + j) * LANES]; }
        int it = 0;
        for (int k = k0; k < k1; k += RING_B_E, ++it) {
            const int st = it % RING_B_STAGES, ne = min(RING_B_E, k1 - k);
            mbar_wait(&full[st], (it / RING_B_STAGES) & 1);
            const float *src = ring + (size_t)st * stage_f;
            if (warp < ne) {   // one element per consumer warp and stage
                const int kk = k + warp;
                const float *Ss = src + (size_t)warp * m * LANES + lane, *Ys = src + ((size_t)RING_B_E + warp) * m * LANES + lane;
                const float xk = x[(size_t)kk * LANES];
                float dk = cgv * gv[(size_t)kk * LANES];
                float yj[M], sj[M];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    yj[j] = j < m ? Ys[j * LANES] : 0.f;
                    sj[j] = j < m ? Ss[j * LANES] : 0.f;
                }
#pragma unroll
                for (int j = 0; j < M; ++j) dk += cS[j] * sj[j] + cY[j] * yj[j];
                if (mode == 1) {
                    d[(size_t)kk * LANES] = dk;
                    xt[(size_t)kk * LANES] = xk + al * dk;
                } else if (mode == 2) xt[(size_t)kk * LANES] = xk + al * d[(size_t)kk * LANES];
                else if (mode == 3) xt[(size_t)kk * LANES] = xk;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
    } else {
        for (int k = k0 + warp; k < k1; k += RING_CONSUMERS) {
            const float xk = x[(size_t)k * LANES];
            if (mode == 2) xt[(size_t)k * LANES] = xk + al * d[(size_t)k * LANES];
            else if (mode == 3) xt[(size_t)k * LANES] = xk;
        }
    }
}

// Slot assignment: the unfinished decoys of table block t fill the slots from the start of the
// block (one CTA per block, chunked block-wide exclusive scans): first, in decoy order, those
// whose run in force scores the restraints, then those whose run does not (the vdw-only runs of
// remove_clash, folding.py:119: all restraint weights zero) -- the restraint kernel and its
// reduction skip the slot groups of the second kind (12 % of the evaluations of a fold).
// identity != 0: every decoy gets its own slot and every group is scored (final consistent
// pass / parity entries: the reported terms are unweighted and must be complete).
__global__ void __launch_bounds__(1024) compact_kernel(FoldState s, int identity)
{
    __shared__ int wsum[32];
    __shared__ int base_s, nk1_s;
    const int t = blockIdx.x, d0 = s.tab_d0[t], nt = s.tab_n[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span = (nt + LANES - 1) / LANES * LANES, Npad = s.Npad;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int pass = 0; pass < (identity ? 1 : 2); ++pass) {
        for (int c0 = 0; c0 < span; c0 += 1024) {
            const int i = c0 + threadIdx.x;
            bool pick = false;
            if (pass == 0) {   // which position groups the L-BFGS kernels stream (a warp of this scan is one position group)
                const unsigned ma = __ballot_sync(0xffffffffu, i < nt && s.status[d0 + i] != ST_DONE);
                if (lane == 0 && i < span) s.gactive[(d0 + i) / LANES] = ma != 0;
            }
            if (i < nt) {
                const int n = d0 + i;
                if (identity) pick = true;
                else if (s.status[n] != ST_DONE) {
                    // a closing evaluation reports every term, whatever the weights in force
                    const bool k1 = !s.k1skip || s.status[n] == ST_FINAL || s.wl[n] != 0.f || s.wl[(size_t)Npad + n] != 0.f || s.wl[(size_t)2 * Npad + n] != 0.f;
                    pick = pass == 0 ? k1 : !k1;
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, pick);
            if (lane == 0) wsum[warp] = __popc(m);
            __syncthreads();
            int off = base_s;
            for (int w = 0; w < warp; ++w) off += wsum[w];
            if (pick) {
                const int slot = d0 + off + __popc(m & ((1u << lane) - 1));
                s.perm[slot] = d0 + i;
                s.slot_of[d0 + i] = slot;
            }
            __syncthreads();
            if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < 32; ++w) tot += wsum[w]; base_s += tot; }
            __syncthreads();
        }
        if (pass == 0 && threadIdx.x == 0) nk1_s = base_s;
        __syncthreads();
    }
    const int nlive = base_s, nk1 = identity ? nlive : nk1_s;
    for (int i = nlive + threadIdx.x; i < span; i += 1024) s.perm[d0 + i] = -1;
    for (int g = threadIdx.x; g < span / LANES; g += 1024) {
        s.gslot[d0 / LANES + g] = g * LANES < nlive;
        s.gneedk1[d0 / LANES + g] = g * LANES < nk1;
    }
    if (threadIdx.x == 0) {
        s.nslot[t] = nlive;
        if (!identity) s.k1count[t] += nk1;
    }
}

__global__ void init_state_kernel(FoldState s, const float *__restrict__ tors_nat)
{
    // tors_nat [N][L][3] -> x, xt grouped; scalars reset
    const int g = blockIdx.x, lane = threadIdx.x, n = g * LANES + lane;
    float *x = s.x + (size_t)g * s.ndof * LANES + lane, *xt = s.xt + (size_t)g * s.ndof * LANES + lane;
    for (int k = 0; k < s.ndof; ++k) {
        const float v = n < s.N ? tors_nat[(size_t)n * s.ndof + k] : (float)TRX_PI;
        x[k * LANES] = v;
        xt[k * LANES] = v;
    }
    s.held[n] = 0;
    s.orig[n] = -1;
    s.mccyc[n] = 0;
    s.flags[n] = 0;
    if (s.nbl_ok) s.nbl_ok[n] = 0;
    s.status[n] = n < s.N ? ST_INIT : ST_DONE;
    s.run[n] = 0; s.hist[n] = 0; s.head[n] = 0; s.iter[n] = 0; s.bt[n] = 0; s.restart[n] = 1; s.nmem[n] = 0;
    s.f[n] = 0.0; s.alpha[n] = 0.f; s.slope[n] = 0.f; s.evals[n] = 0; s.iters[n] = 0;
    s.ft[n] = 0.0; s.Evdw[n] = 0.0; s.Ehb[n] = 0.0;
    for (int k = 0; k < TRX_NTERM; ++k) { s.wl[(size_t)k * s.Npad + n] = s.runs[0].w[k]; s.terms[(size_t)k * s.Npad + n] = 0.0; }
    for (int k = 0; k < 3; ++k) { s.fmem[(size_t)k * s.Npad + n] = 0.0; s.E3[(size_t)k * s.Npad + n] = 0.0; }
}

// ---- continuous batching: queue store, turnover (park + refill) ----------------------------------------
// Start of a call: the caller's start torsions [Nq][L][3] (block-major, caller order) -> queue records.
__global__ void __launch_bounds__(256) queue_init_kernel(FoldState s, const float *__restrict__ tors_nat)
{
    const int gq = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, q = gq * LANES + lane;
    int t = 0;
    while (t + 1 < s.ntab && q >= s.qd0[t + 1]) ++t;
    const bool real = q - s.qd0[t] < s.nq_tab[t];
    const int c = s.qc0[t] + (q - s.qd0[t]);
    float *__restrict__ qt = s.q_tors + (size_t)gq * s.ndof_t * LANES + lane;
    for (int k = warp; k < s.ndof_t; k += nw) qt[(size_t)k * LANES] = real ? tors_nat[(size_t)c * s.ndof_t + k] : (float)TRX_PI;
    if (warp == 0) {
        s.q_run[q] = 0; s.q_held[q] = 0; s.q_evals[q] = 0; s.q_iters[q] = 0; s.q_flags[q] = 0;
        for (int k = 0; k < TRX_NTERM; ++k) s.q_terms[(size_t)k * s.Nqpad + q] = 0.0;
    }
}

// Turnover, step 1 (one CTA per table block): the positions whose decoy has left the segment (or that are
// empty) take the next waiting queue entries, in position order.  Deterministic, and immaterial to the
// results: a decoy's trajectory does not depend on the position it occupies.
__global__ void __launch_bounds__(1024) turnover_plan_kernel(FoldState s)
{
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int t = blockIdx.x, d0 = s.tab_d0[t], nt = s.tab_n[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span = (nt + 1023) / 1024 * 1024;
    const int cursor = s.qcursor[t], nq = s.nq_tab[t];
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < span; c0 += 1024) {
        const int i = c0 + threadIdx.x;
        const bool free_ = i < nt && s.status[d0 + i] == ST_DONE;
        const unsigned m = __ballot_sync(0xffffffffu, free_);
        if (lane == 0) wsum[warp] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (i < nt) {
            int nid = -2;
            if (free_) {
                const int k = cursor + off + __popc(m & ((1u << lane) - 1));
                nid = k < nq ? s.qd0[t] + k : -1;
                if (nid == -1 && s.orig[d0 + i] == -1) nid = -2;   // empty stays empty
            }
            s.newid[d0 + i] = nid;
        } else if (i < (nt + LANES - 1) / LANES * LANES) {
            s.newid[d0 + i] = -2;   // padding positions of the block's last group never hold a decoy
        }
        __syncthreads();
        if (threadIdx.x == 0) { int a = 0; for (int w = 0; w < 32; ++w) a += wsum[w]; base_s += a; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int nfree = base_s, taken = min(nfree, nq - cursor);
        s.qcursor[t] = cursor + taken;
        s.qocc[t] = nt - nfree + taken;
    }
}

__device__ void turnover_move_one(const FoldState &s, const int pos);

// Turnover, step 2 (one CTA per position group; a position that changes hands is handled by the whole CTA).  PARK: the decoy's record goes back to the
// queue store -- torsions (read back from the coordinates after a Cartesian segment), the coordinates it
// holds or that the next (Cartesian) segment starts from, the terms of its closing evaluation, counters --
// and, when it has left the LAST segment, its results go to the output arrays in the caller's order.
// LOAD: the next waiting decoy starts the segment from its record.
__global__ void __launch_bounds__(128) turnover_move_kernel(FoldState s)
{
    // one CTA per position group: in most rounds none of its 32 positions changes hands
    const int nid_mine = s.newid[blockIdx.x * LANES + (threadIdx.x & 31)];
    if (!__syncthreads_or(nid_mine != -2)) return;
    for (int lanepos = 0; lanepos < LANES; ++lanepos) turnover_move_one(s, blockIdx.x * LANES + lanepos);
}

__device__ void turnover_move_one(const FoldState &s, const int pos)
{
    const int nid = s.newid[pos];
    if (nid == -2) return;   // uniform over the CTA
    const int L = s.L, Npad = s.Npad, Nqpad = s.Nqpad;
    const int old = s.orig[pos];
    const size_t vb = (size_t)(pos / LANES) * s.ndof * LANES + pos % LANES;
    float *__restrict__ x = s.x + vb, *__restrict__ xt = s.xt + vb;
    if (old >= 0) {
        float *__restrict__ qt = s.q_tors + (size_t)(old / LANES) * s.ndof_t * LANES + old % LANES;
        float *__restrict__ qX = s.has_cart ? s.q_X + (size_t)(old / LANES) * s.ndof_c * LANES + old % LANES : nullptr;
        const int slot = s.slot_of[pos], held = s.held[pos];
        const int c = caller_index(s, old);
        if (s.cart) {
            // the coordinates the decoy holds from now on, and its torsions read back from them
            for (int e = threadIdx.x; e < L * NAT3; e += blockDim.x) {
                const float v = x[(size_t)e * LANES];
                qX[(size_t)e * LANES] = v;
                if (s.seg_last && s.o_xyz) s.o_xyz[(size_t)c * L * NAT3 + e] = v;
            }
            auto load = [&](int i, int a) -> f3 { return {x[((size_t)i * NAT3 + a * 3) * LANES], x[((size_t)i * NAT3 + a * 3 + 1) * LANES], x[((size_t)i * NAT3 + a * 3 + 2) * LANES]}; };
            for (int i = threadIdx.x; i < L; i += blockDim.x) {
                float phi, psi, omg;
                readback_torsions(load, i, L, phi, psi, omg);
                qt[(size_t)(i * 3 + 0) * LANES] = phi; qt[(size_t)(i * 3 + 1) * LANES] = psi; qt[(size_t)(i * 3 + 2) * LANES] = omg;
                if (s.seg_last) { s.o_tors[(size_t)c * s.ndof_t + i * 3] = phi; s.o_tors[(size_t)c * s.ndof_t + i * 3 + 1] = psi; s.o_tors[(size_t)c * s.ndof_t + i * 3 + 2] = omg; }
            }
        } else {
            for (int k = threadIdx.x; k < s.ndof_t; k += blockDim.x) {
                const float v = x[(size_t)k * LANES];
                qt[(size_t)k * LANES] = v;
                if (s.seg_last) s.o_tors[(size_t)c * s.ndof_t + k] = v;
            }
            // coordinates of the closing evaluation (slot space): the start of a Cartesian segment / the result
            const float *__restrict__ X = s.X + (size_t)(slot / LANES) * s.Lpad * NAT3 * LANES + slot % LANES;
            if (s.park_X && !held)
                for (int e = threadIdx.x; e < L * NAT3; e += blockDim.x) qX[(size_t)e * LANES] = X[(size_t)e * LANES];
            if (s.seg_last && s.o_xyz)
                for (int e = threadIdx.x; e < L * NAT3; e += blockDim.x)
                    s.o_xyz[(size_t)c * L * NAT3 + e] = held ? qX[(size_t)e * LANES] : X[(size_t)e * LANES];
        }
        if (threadIdx.x == 0) {
            const bool keep_terms = !s.cart && held;   // still holding its Cartesian coordinates: their terms stand
            s.q_run[old] = s.run[pos]; s.q_evals[old] = s.evals[pos]; s.q_iters[old] = s.iters[pos]; s.q_flags[old] = s.flags[pos];
            s.q_held[old] = s.cart ? 1 : held;
            for (int k = 0; k < TRX_NTERM; ++k) {
                const double v = keep_terms ? s.theld[(size_t)k * Npad + pos] : s.terms[(size_t)k * Npad + pos];
                s.q_terms[(size_t)k * Nqpad + old] = v;
                if (s.seg_last) s.o_terms[(size_t)c * TRX_NTERM + k] = v;
            }
            if (s.seg_last) {
                s.o_stats[(size_t)c * 3] = s.evals[pos];
                s.o_stats[(size_t)c * 3 + 1] = s.iters[pos];
                s.o_stats[(size_t)c * 3 + 2] = s.naccept[pos];
                int fl = s.flags[pos];
                bool finite = true;
                for (int k = 0; k < TRX_NTERM; ++k) finite = finite && isfinite(s.q_terms[(size_t)k * Nqpad + old]);
                if (!finite) fl |= TRX_DECOY_NONFINITE;
                s.o_flags[c] = fl;
            }
        }
    }
    __syncthreads();   // the old record is out before the position is overwritten
    if (nid >= 0) {
        const float *__restrict__ qt = s.q_tors + (size_t)(nid / LANES) * s.ndof_t * LANES + nid % LANES;
        if (s.cart) {
            const float *__restrict__ qX = s.q_X + (size_t)(nid / LANES) * s.ndof_c * LANES + nid % LANES;
            for (int e = threadIdx.x; e < s.ndof_c; e += blockDim.x) {
                const float v = e < L * NAT3 ? qX[(size_t)e * LANES] : 0.f;
                x[(size_t)e * LANES] = v;
                xt[(size_t)e * LANES] = v;
            }
        } else {
            for (int k = threadIdx.x; k < s.ndof_t; k += blockDim.x) {
                const float v = qt[(size_t)k * LANES];
                x[(size_t)k * LANES] = v;
                xt[(size_t)k * LANES] = v;
            }
        }
    }
    if (threadIdx.x == 0) {
        const int n = pos;
        s.orig[n] = nid;
        s.hist[n] = 0; s.head[n] = 0; s.iter[n] = 0; s.bt[n] = 0; s.restart[n] = 1; s.nmem[n] = 0;
        s.f[n] = 0.0; s.alpha[n] = 0.f; s.slope[n] = 0.f; s.ft[n] = 0.0; s.Evdw[n] = 0.0;
        s.mccyc[n] = 0; s.naccept[n] = 0; s.fsave[n] = 0.0;
        if (s.nbl_ok) s.nbl_ok[n] = 0;   // another decoy: its pair list is rebuilt by its first evaluation
        for (int k = 0; k < 3; ++k) { s.fmem[(size_t)k * Npad + n] = 0.0; s.E3[(size_t)k * Npad + n] = 0.0; }
        if (nid >= 0) {
            const int run = s.q_run[nid];
            // a record beyond the segment passes through with one closing evaluation
            s.status[n] = (run >= s.seg_lo && run < s.seg_hi) ? ST_INIT : ST_FINAL;
            s.run[n] = run; s.evals[n] = s.q_evals[nid]; s.iters[n] = s.q_iters[nid]; s.held[n] = s.q_held[nid];
            s.flags[n] = s.q_flags[nid];
            const Run &r = s.runs[min(run, s.nruns - 1)];
            for (int k = 0; k < TRX_NTERM; ++k) {
                s.wl[(size_t)k * Npad + n] = r.w[k];
                s.terms[(size_t)k * Npad + n] = 0.0;
                s.theld[(size_t)k * Npad + n] = s.q_terms[(size_t)k * Nqpad + nid];
            }
        } else {
            s.status[n] = ST_DONE; s.run[n] = s.nruns; s.held[n] = 0; s.evals[n] = 0; s.iters[n] = 0; s.flags[n] = 0;
        }
    }
}

// The round budget ran out: every unfinished decoy stops where it is (its accepted point) and gets its
// closing evaluation; decoys still waiting in the queue are handed out to be closed too.
__global__ void __launch_bounds__(256) force_final_kernel(FoldState s)
{
    const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, n = g * LANES + lane;
    if (n >= s.N) return;
    const int st = s.status[n];
    if (st != ST_INIT && st != ST_LS) return;
    const size_t vb = (size_t)g * s.ndof * LANES + lane;
    for (int k = warp; k < s.ndof; k += nw) s.xt[vb + (size_t)k * LANES] = s.x[vb + (size_t)k * LANES];
    if (warp == 0) { s.status[n] = ST_FINAL; s.run[n] = s.nruns; s.flags[n] |= TRX_DECOY_UNFINISHED; }
}
__global__ void force_final_queue_kernel(FoldState s)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < s.Nqpad && s.q_run[q] < s.nruns) { s.q_run[q] = s.nruns; s.q_flags[q] |= TRX_DECOY_UNFINISHED; }
}

// Which positions swap: in table block t, the k-th finished decoy among the first `nlive` positions
// with the k-th unfinished decoy behind them (nlive = unfinished decoys of the block).  One CTA per block.
__global__ void __launch_bounds__(1024) migrate_plan_kernel(FoldState s)
{
    __shared__ int wsum[32];
    __shared__ int tot_s, base_s;
    const int t = blockIdx.x, d0 = s.tab_d0[t], nt = s.tab_n[t];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int span = (nt + 1023) / 1024 * 1024;
    int cnt = 0;
    for (int i = threadIdx.x; i < nt; i += 1024) cnt += s.status[d0 + i] != ST_DONE;
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) wsum[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { int a = 0; for (int w = 0; w < 32; ++w) a += wsum[w]; tot_s = a; }
    __syncthreads();
    const int nlive = tot_s;
    // pass 0: finished decoys in front of nlive -> mig_a; pass 1: unfinished ones behind it -> mig_b; by position
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
        if (threadIdx.x == 0) base_s = 0;
        __syncthreads();
        for (int c0 = 0; c0 < span; c0 += 1024) {
            const int i = c0 + threadIdx.x;
            bool pick = false;
            if (i < nt) {
                const bool live = s.status[d0 + i] != ST_DONE;
                pick = pass == 0 ? (i < nlive && !live) : (i >= nlive && live);
            }
            const unsigned m = __ballot_sync(0xffffffffu, pick);
            if (lane == 0) wsum[warp] = __popc(m);
            __syncthreads();
            int off = base_s;
            for (int w = 0; w < warp; ++w) off += wsum[w];
            if (pick) (pass == 0 ? s.mig_a : s.mig_b)[d0 + off + __popc(m & ((1u << lane) - 1))] = d0 + i;
            __syncthreads();
            if (threadIdx.x == 0) { int a = 0; for (int w = 0; w < 32; ++w) a += wsum[w]; base_s += a; }
            __syncthreads();
        }
        if (pass == 1 && threadIdx.x == 0) s.mig_n[t] = base_s;
    }
}

// Swaps everything a position carries between a (finished, front) and b (unfinished, back); the
// L-BFGS history and Gram matrices only travel b -> a (a finished decoy's history is dead).
__global__ void __launch_bounds__(256) migrate_swap_kernel(FoldState s)
{
    const int t = blockIdx.y, k = blockIdx.x;
    if (k >= s.mig_n[t]) return;
    const int a = s.mig_a[s.tab_d0[t] + k], b = s.mig_b[s.tab_d0[t] + k];
    const int nd = s.ndof, m = s.m, Npad = s.Npad;
    const size_t va = (size_t)(a / LANES) * nd * LANES + a % LANES, vb = (size_t)(b / LANES) * nd * LANES + b % LANES;
    float *vecs[6] = {s.x, s.g, s.d, s.xt, s.gt, s.xsave};
    for (int e = threadIdx.x; e < nd; e += blockDim.x) {
#pragma unroll
        for (int v = 0; v < 6; ++v) {
            float *p = vecs[v];
            const float ta = p[va + (size_t)e * LANES], tb = p[vb + (size_t)e * LANES];
            p[va + (size_t)e * LANES] = tb;
            p[vb + (size_t)e * LANES] = ta;
        }
    }
    {
        const size_t ha = (size_t)(a / LANES) * nd * m * LANES + a % LANES, hb = (size_t)(b / LANES) * nd * m * LANES + b % LANES;
        for (size_t e = threadIdx.x; e < (size_t)nd * m; e += blockDim.x) {
            s.S[ha + e * LANES] = s.S[hb + e * LANES];
            s.Y[ha + e * LANES] = s.Y[hb + e * LANES];
        }
        const int MM = 2 * s.lb_M * s.lb_M;
        const size_t ga = (size_t)(a / LANES) * MM * LANES + a % LANES, gb = (size_t)(b / LANES) * MM * LANES + b % LANES;
        for (int e = threadIdx.x; e < MM; e += blockDim.x) s.gram[ga + (size_t)e * LANES] = s.gram[gb + (size_t)e * LANES];
    }
    if (threadIdx.x == 0) {
        auto swp_d = [&](double *p, size_t stride, int cnt) { for (int q = 0; q < cnt; ++q) { const double u = p[q * stride + a]; p[q * stride + a] = p[q * stride + b]; p[q * stride + b] = u; } };
        auto swp_f = [&](float *p, size_t stride, int cnt) { for (int q = 0; q < cnt; ++q) { const float u = p[q * stride + a]; p[q * stride + a] = p[q * stride + b]; p[q * stride + b] = u; } };
        auto swp_i = [&](int *p) { const int u = p[a]; p[a] = p[b]; p[b] = u; };
        swp_d(s.f, 0, 1); swp_d(s.fmem, Npad, 3); swp_d(s.fsave, 0, 1); swp_d(s.terms, Npad, TRX_NTERM); swp_d(s.ft, 0, 1); swp_d(s.theld, Npad, TRX_NTERM);
        swp_f(s.alpha, 0, 1); swp_f(s.slope, 0, 1); swp_f(s.wl, Npad, TRX_NTERM);
        swp_i(s.nmem); swp_i(s.hist); swp_i(s.head); swp_i(s.iter); swp_i(s.run); swp_i(s.bt); swp_i(s.status); swp_i(s.restart);
        swp_i(s.evals); swp_i(s.iters); swp_i(s.naccept); swp_i(s.held); swp_i(s.orig); swp_i(s.mccyc); swp_i(s.slot_of); swp_i(s.flags);
        if (s.nbl_ok) { s.nbl_ok[a] = 0; s.nbl_ok[b] = 0; }   // the lists stay behind: rebuilt at the new positions (same results: energies and gradients are exact integer sums)
    }
}

static void upload_model()
{
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done[dev]) return;
    Model mdl;
    for (int k = 0; k < 20; ++k) { mdl.cen_s[k] = (float)TRX_CEN_S[k]; mdl.r_cen[k] = (float)TRX_R_CEN[k]; }
    for (int k = 0; k < 5; ++k) mdl.r_bb[k] = (float)TRX_R_BB[k];
    for (int c = 0; c < 2; ++c) for (int k = 0; k < TRX_RAMA_NB; ++k) for (int j = 0; j < 5; ++j) mdl.rama[c][k][j] = (float)TRX_RAMA[c][k][j];
    mdl.rama_off[0] = (float)TRX_RAMA_OFFSET[0]; mdl.rama_off[1] = (float)TRX_RAMA_OFFSET[1];
    cudaMemcpyToSymbol(c_model, &mdl, sizeof(mdl));
    done[dev] = true;
}

}  // namespace trx

using namespace trx;

struct trx_fold_batch {
    trx_ctx *ctx = nullptr;
    FoldState s{};
    std::vector<trx_tables *> tabs;
    std::vector<int> tab_g0, tab_ng;
    void *arena = nullptr;
    size_t arena_bytes = 0;
    int *d_aa = nullptr;
    Run *d_runs = nullptr;
    int *h_poll = nullptr;       // pinned: live slots, occupied positions and queue cursors of every table block
    size_t vdw_smem = 0, lb_smem = 0;
    struct Segment { int lo, hi, cart; };
    std::vector<Segment> segs;   // maximal stretches of torsion-space / Cartesian runs
    bool has_cart = false;
    bool lb_ring = true;         // L-BFGS sweeps through the bulk-async shared-memory ring (TRX_NO_LB_RING=1: register-staged sweeps)
    bool migrate = true;         // TRX_NO_MIGRATE=1 disables the packing of unfinished decoys (same results, bit for bit)
    int mig_num = 3, mig_den = 4; // pack when unfinished <= mig_num/mig_den of the positions they are spread over (1/2: 1246, 3/4: 1257 decoys/s)
    long long k1_decoy_evals[16] = {0};   // restraint-kernel decoy evaluations of the last call, per table block
    std::vector<int> status;              // TRX_DECOY_* bits of every decoy of the last call (caller order)
};

extern "C" {

int trx_fold_destroy(trx_fold_batch *b)
{
    if (!b) return TRX_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    b->ctx->dev_free(b->arena);
    b->ctx->dev_free(b->d_aa);
    b->ctx->dev_free(b->d_runs);
    b->ctx->pinned_release(b->h_poll);
    trx_ctx *ctx = b->ctx;
    delete b;
    ctx_release(ctx);
    return TRX_OK;
}

int trx_fold_create(trx_ctx *ctx, int ntab, trx_tables *const *tabs, const int *ndecoys, const int32_t *aa,
                    const trx_run *runs, int nruns, int lbfgs_m, trx_fold_batch **out)
{
    TRX_REQUIRE(ctx && tabs && ndecoys && aa && runs && out, "trx_fold_create: NULL argument");
    TRX_REQUIRE(ntab >= 1 && ntab <= 16, "trx_fold_create: ntab=%d out of range [1,16]", ntab);
    TRX_REQUIRE(nruns >= 1 && nruns <= 256, "trx_fold_create: nruns=%d out of range [1,256]", nruns);
    TRX_REQUIRE(lbfgs_m >= 1 && lbfgs_m <= 24, "trx_fold_create: lbfgs_m=%d out of range [1,24]", lbfgs_m);
    TRX_CUDA(cudaSetDevice(ctx->device));
    const int L = tabs[0]->L;
    int G = 0;
    trx_fold_batch *b = new trx_fold_batch();
    struct Guard {   // every early return below releases what has been built so far
        trx_fold_batch *b;
        ~Guard()
        {
            if (!b) return;
            if (!b->ctx) { delete b; return; }
            cudaStreamSynchronize(b->ctx->stream);
            b->ctx->dev_free(b->arena);
            b->ctx->dev_free(b->d_aa);
            b->ctx->dev_free(b->d_runs);
            b->ctx->pinned_release(b->h_poll);
            delete b;
        }
    } guard{b};
    b->ctx = ctx;
    if (const char *ev = getenv("TRX_NO_MIGRATE")) b->migrate = !(ev[0] && ev[0] != '0');
    if (const char *ev = getenv("TRX_MIGRATE_AT")) { int p = atoi(ev); if (p > 0 && p < 100) { b->mig_num = p; b->mig_den = 100; } }   // development knob (percent)
    for (int t = 0; t < ntab; ++t) {
        TRX_REQUIRE(tabs[t] && tabs[t]->ctx == ctx && tabs[t]->L == L, "trx_fold_create: tables %d: NULL, other context or other L", t);
        TRX_REQUIRE(ndecoys[t] > 0, "trx_fold_create: ndecoys[%d] must be positive", t);
        TRX_REQUIRE(t == ntab - 1 || ndecoys[t] % LANES == 0, "trx_fold_create: all but the last decoy block must be multiples of 32");
        b->tabs.push_back(tabs[t]);
        b->tab_g0.push_back(G);
        b->tab_ng.push_back(num_groups(ndecoys[t]));
        G += num_groups(ndecoys[t]);
    }
    for (int i = 0; i < L; ++i) TRX_REQUIRE(aa[i] >= 0 && aa[i] < 20, "trx_fold_create: aa[%d]=%d out of range", i, aa[i]);
    for (int r = 0; r < nruns; ++r) {
        TRX_REQUIRE(runs[r].max_iter >= 0 && (!runs[r].clash_check || (runs[r].skip_to > r && runs[r].skip_to <= nruns)),
                    "trx_fold_create: run %d has a bad skip_to/max_iter", r);
        if (b->segs.empty() || b->segs.back().cart != (runs[r].cartesian ? 1 : 0)) b->segs.push_back({r, r + 1, runs[r].cartesian ? 1 : 0});
        else b->segs.back().hi = r + 1;
        b->has_cart = b->has_cart || runs[r].cartesian;
    }
    // a Cartesian segment starts from the coordinates the torsion-space segment before it parked
    TRX_REQUIRE(!runs[0].cartesian, "trx_fold_create: the schedule may not open with a Cartesian run (put a torsion-space run with max_iter 0 in front)");
    for (int r = 0; r < nruns; ++r)   // a clash check may not jump over a Cartesian run
        if (runs[r].clash_check)
            for (int q = r + 1; q < runs[r].skip_to; ++q)
                TRX_REQUIRE(!runs[q].cartesian || runs[r].cartesian, "trx_fold_create: run %d skips over the Cartesian run %d", r, q);
    int N = 0;
    for (int t = 0; t < ntab; ++t) N += ndecoys[t];
    FoldState &s = b->s;
    s.N = (G - 1) * LANES + ((ndecoys[ntab - 1] - 1) % LANES + 1);
    TRX_REQUIRE(s.N == N, "trx_fold_create: internal decoy count mismatch");
    s.G = G; s.Npad = G * LANES; s.L = L; s.Lpad = padded_length(L); s.ndof = 3 * L; s.m = lbfgs_m; s.nruns = nruns;
    s.ndof_t = 3 * L; s.ndof_c = NAT3 * s.Lpad; s.cart = 0; s.seg_hi = nruns;
    const int ndof_max = b->has_cart ? s.ndof_c : s.ndof_t;
    s.has_cart = b->has_cart ? 1 : 0;
    s.k1skip = 1;
    if (const char *ev = getenv("TRX_NO_K1SKIP")) s.k1skip = !(ev[0] && ev[0] != '0');   // A/B and tests: same results, bit for bit
    b->vdw_smem = sizeof(float4) * 6 * L + ((sizeof(int) * 18 * L + 15) / 16) * 16 + sizeof(float4) * L + sizeof(float2) * L;
    TRX_REQUIRE(b->vdw_smem <= 220 * 1024, "trx_fold_create: L=%d exceeds the shared-memory budget of the vdw kernel", L);
    TRX_REQUIRE(L < 65536, "trx_fold_create: L too large");
    // one arena, carved into aligned pieces
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t vec = (size_t)G * ndof_max * LANES * sizeof(float), np = (size_t)s.Npad;
    size_t o_x = carve(vec), o_g = carve(vec), o_d = carve(vec), o_xt = carve(vec), o_gt = carve(vec);
    const int lbM = lbfgs_m <= 8 ? 8 : (lbfgs_m <= 16 ? 16 : (lbfgs_m <= 20 ? 20 : 24));
    size_t o_S = carve(vec * s.m), o_Y = carve(vec * s.m), o_rho = carve((size_t)G * 2 * lbM * lbM * LANES * sizeof(float));
    size_t o_lbp = carve((size_t)G * 16 * (5 * lbM + 8) * LANES * sizeof(float)), o_lbc = carve((size_t)G * (2 * lbM + 3) * LANES * sizeof(float));
    size_t o_f = carve(np * 8), o_al = carve(np * 4), o_sl = carve(np * 4), o_fm = carve(np * 24);
    size_t o_i[10];
    for (int k = 0; k < 10; ++k) o_i[k] = carve(np * 4);
    size_t o_terms = carve(np * 8 * TRX_NTERM), o_ft = carve(np * 8), o_wl = carve(np * 4 * TRX_NTERM);
    size_t o_X = carve((size_t)G * s.Lpad * NAT3 * LANES * 4), o_xn = carve(np * L * NATP * 4), o_gn = carve(np * L * NATP * 4);
    size_t o_gk = carve((size_t)G * s.Lpad * 9 * LANES * 4), o_E3 = carve(np * 8 * 3), o_Ev = carve(np * 8);
    size_t o_ga = carve((size_t)G * 4), o_na = carve(256);
    size_t o_xs = carve(vec), o_fs = carve(np * 8), o_nacc = carve(np * 4);
    size_t o_perm = carve(np * 4), o_gs = carve((size_t)G * 4), o_ns = carve(256), o_ws = carve(np * 4 * TRX_NTERM);
    size_t o_gk1f = carve((size_t)G * 4);
    size_t o_orig = carve(np * 4), o_ma = carve(np * 4), o_mb = carve(np * 4), o_mn = carve(256);
    size_t o_held = carve(np * 4), o_th = carve(np * 8 * TRX_NTERM);
    size_t o_mcc = carve(np * 4), o_sof = carve(np * 4), o_nid = carve(np * 4), o_qc = carve(256), o_qo = carve(256), o_k1c = carve(256);
    size_t o_flg = carve(np * 4), o_Eh = carve(np * 8);
    // Verlet list of the pair search: measured SLOWER than the full scan (the atom-pair pass over the queued pairs, not the
    // O(L^2) sphere scan, is what the kernel spends its time on: 620 vs 556 us per launch at L=300, 2313 vs 1987 us at
    // L=800), so it is opt-in (TRX_NBL=1); results are the same bits either way.
    const bool nbl = getenv("TRX_NBL") && getenv("TRX_NBL")[0] && getenv("TRX_NBL")[0] != '0';
    size_t o_nok = carve(np * 4), o_nref = carve(nbl ? np * L * sizeof(float4) : 256), o_nj = carve(nbl ? np * L * NBL_W * 2 : 256), o_nc = carve(nbl ? np * L * 2 : 256);
    cudaError_t e = ctx->dev_alloc(&b->arena, off);
    if (e != cudaSuccess) {
        set_error("trx_fold_create: cudaMalloc(%zu bytes) failed: %s", off, cudaGetErrorString(e));
        b->arena = nullptr;
        return TRX_ERR_NOMEM;
    }
    b->arena_bytes = off;
    TRX_CUDA(cudaMemsetAsync(b->arena, 0, off, ctx->stream));
    char *A = (char *)b->arena;
    s.x = (float *)(A + o_x); s.g = (float *)(A + o_g); s.d = (float *)(A + o_d); s.xt = (float *)(A + o_xt); s.gt = (float *)(A + o_gt);
    s.S = (float *)(A + o_S); s.Y = (float *)(A + o_Y); s.gram = (float *)(A + o_rho); s.lb_M = lbM;
    s.lbpart = (float *)(A + o_lbp); s.lbcoef = (float *)(A + o_lbc);
    s.f = (double *)(A + o_f); s.alpha = (float *)(A + o_al); s.slope = (float *)(A + o_sl); s.fmem = (double *)(A + o_fm);
    s.nmem = (int *)(A + o_i[0]); s.hist = (int *)(A + o_i[1]); s.head = (int *)(A + o_i[2]); s.iter = (int *)(A + o_i[3]);
    s.run = (int *)(A + o_i[4]); s.bt = (int *)(A + o_i[5]); s.status = (int *)(A + o_i[6]); s.restart = (int *)(A + o_i[7]);
    s.evals = (int *)(A + o_i[8]); s.iters = (int *)(A + o_i[9]);
    s.terms = (double *)(A + o_terms); s.ft = (double *)(A + o_ft); s.wl = (float *)(A + o_wl);
    s.X = (float *)(A + o_X); s.xnat = (float *)(A + o_xn); s.gnat = (float *)(A + o_gn); s.gk1 = (float *)(A + o_gk);
    s.E3 = (double *)(A + o_E3); s.Evdw = (double *)(A + o_Ev); s.gactive = (int *)(A + o_ga);
    s.xsave = (float *)(A + o_xs); s.fsave = (double *)(A + o_fs); s.naccept = (int *)(A + o_nacc);
    s.perm = (int *)(A + o_perm); s.gslot = (int *)(A + o_gs); s.nslot = (int *)(A + o_ns); s.wslot = (float *)(A + o_ws);
    s.gneedk1 = (int *)(A + o_gk1f);
    s.orig = (int *)(A + o_orig); s.mig_a = (int *)(A + o_ma); s.mig_b = (int *)(A + o_mb); s.mig_n = (int *)(A + o_mn);
    s.held = (int *)(A + o_held); s.theld = (double *)(A + o_th);
    s.mccyc = (int *)(A + o_mcc); s.slot_of = (int *)(A + o_sof); s.newid = (int *)(A + o_nid);
    s.qcursor = (int *)(A + o_qc); s.qocc = (int *)(A + o_qo); s.k1count = (long long *)(A + o_k1c);
    s.flags = (int *)(A + o_flg); s.Ehb = (double *)(A + o_Eh);
    s.nbl_ok = (int *)(A + o_nok); s.nbl_ref = (float4 *)(A + o_nref); s.nbl_j = nbl ? (unsigned short *)(A + o_nj) : nullptr; s.nbl_cnt = (unsigned short *)(A + o_nc);
    s.mc = McOpts{};
    static_assert(64 * sizeof(int) <= trx_ctx::PINNED_BLOCK, "h_poll");
    TRX_CUDA(ctx->pinned_alloc((void **)&b->h_poll));
    s.ntab = ntab;
    for (int t = 0; t < ntab; ++t) { s.tab_d0[t] = b->tab_g0[t] * LANES; s.tab_n[t] = ndecoys[t]; }
    TRX_CUDA(ctx->dev_alloc(&b->d_aa, L * sizeof(int)));
    TRX_CUDA(cudaMemcpyAsync(b->d_aa, aa, L * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<Run> hr(nruns);
    for (int r = 0; r < nruns; ++r) {
        for (int k = 0; k < TRX_NTERM; ++k) hr[r].w[k] = (float)runs[r].w[k];
        hr[r].max_iter = runs[r].max_iter; hr[r].tol = (float)runs[r].tol; hr[r].clash_check = runs[r].clash_check;
        hr[r].clash_thr = (float)runs[r].clash_thr; hr[r].skip_to = runs[r].skip_to; hr[r].cartesian = runs[r].cartesian ? 1 : 0;
    }
    TRX_CUDA(ctx->dev_alloc(&b->d_runs, nruns * sizeof(Run)));
    TRX_CUDA(cudaMemcpyAsync(b->d_runs, hr.data(), nruns * sizeof(Run), cudaMemcpyHostToDevice, ctx->stream));
    s.aa = b->d_aa;
    s.runs = b->d_runs;
    upload_model();
    {   // function attributes are per device and the vdw kernel's need follows L: only ever raise it, so that a
        // batch created later for a shorter chain does not cut the limit of one that is still alive
        // (batches are created from several host threads in batch mode: one lock)
        static std::mutex vdw_attr_lock;
        std::lock_guard<std::mutex> guard_attr(vdw_attr_lock);
        static size_t vdw_attr_dev[64] = {};
        size_t &cur = vdw_attr_dev[ctx->device & 63];
        if (b->vdw_smem > cur) {
            TRX_CUDA(cudaFuncSetAttribute(vdw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->vdw_smem));
            cur = b->vdw_smem;
        }
    }
    static_assert(LB_MAXCH == 16, "lbpart is carved for 16 chunks");
    if (const char *ev = getenv("TRX_NO_LB_RING")) b->lb_ring = !(ev[0] && ev[0] != '0');
    auto lb_attr = [&](auto dots, auto step, auto rdots, auto rupd, size_t bytes) -> int {
        b->lb_smem = bytes;
        TRX_CUDA(cudaFuncSetAttribute(dots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        TRX_CUDA(cudaFuncSetAttribute(step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        // ring sizes for the largest history the template serves (the attribute is per device, batches differ in m)
        const int mmax = lbM;
        TRX_CUDA(cudaFuncSetAttribute(rdots, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RING_A_STAGES * ring_a_stage_floats(mmax) * sizeof(float))));
        TRX_CUDA(cudaFuncSetAttribute(rupd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RING_B_STAGES * ring_b_stage_floats(mmax) * sizeof(float))));
        return TRX_OK;
    };
    int rc_attr;
    if (lbM == 8) rc_attr = lb_attr(lbfgs_dots_kernel<8>, lbfgs_step_kernel<8>, lbfgs_dots_ring_kernel<8>, lbfgs_update_ring_kernel<8>, sizeof(LbSmem<8>));
    else if (lbM == 16) rc_attr = lb_attr(lbfgs_dots_kernel<16>, lbfgs_step_kernel<16>, lbfgs_dots_ring_kernel<16>, lbfgs_update_ring_kernel<16>, sizeof(LbSmem<16>));
    else if (lbM == 20) rc_attr = lb_attr(lbfgs_dots_kernel<20>, lbfgs_step_kernel<20>, lbfgs_dots_ring_kernel<20>, lbfgs_update_ring_kernel<20>, sizeof(LbSmem<20>));
    else rc_attr = lb_attr(lbfgs_dots_kernel<24>, lbfgs_step_kernel<24>, lbfgs_dots_ring_kernel<24>, lbfgs_update_ring_kernel<24>, sizeof(LbSmem<24>));
    if (rc_attr) return rc_attr;
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));   // aa may be pinned memory of the caller: done with it on return
    guard.b = nullptr;
    ctx->retain();
    *out = b;
    return TRX_OK;
}

// One evaluation of the current trial points (xt) of every unfinished decoy:
// coordinates, restraint terms, vdw, torsion gradient.  Fills ft, gt, terms.
static int fold_eval(trx_fold_batch *b, const int *ng_tab, bool identity)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    ctx->time_begin("compact");
    compact_kernel<<<s.ntab, 1024, 0, ctx->stream>>>(s, identity ? 1 : 0);
    ctx->time_end("compact");
    if (s.cart) {
        ctx->time_begin("cart_gather");
        cart_gather_kernel<<<s.G, 256, 0, ctx->stream>>>(s);
        ctx->time_end("cart_gather");
    } else {
        ctx->time_begin("nerf");
        nerf_kernel<<<s.G, SEG_THREADS, 0, ctx->stream>>>(s);
        ctx->time_end("nerf");
    }
    for (size_t t = 0; t < b->tabs.size(); ++t) {
        const int ng = ng_tab ? std::min(ng_tab[t], b->tab_ng[t]) : b->tab_ng[t];
        if (ng <= 0) continue;
        int rc = k1_launch<float>(ctx, b->tabs[t], s.G, b->tab_g0[t], ng, s.X, NAT3, s.wslot, nullptr, s.gneedk1, s.E3, s.gk1);
        if (rc) return rc;
    }
    ctx->time_begin("centroid");
    vdw_kernel<<<s.N, VDW_THREADS, b->vdw_smem, ctx->stream>>>(s);
    ctx->time_end("centroid");
    if (s.cart) {
        ctx->time_begin("cart_grad");
        cart_grad_kernel<<<s.G, CART_THREADS, 0, ctx->stream>>>(s);
        ctx->time_end("cart_grad");
    } else {
        ctx->time_begin("torsion_grad");
        torsion_grad_kernel<<<s.G, SEG_THREADS, 0, ctx->stream>>>(s);
        ctx->time_end("torsion_grad");
    }
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

extern "C++" {
template <int M>
static void lbfgs_launch(trx_fold_batch *b, const dim3 &grid)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    if (b->lb_ring) {
        const size_t ra = RING_A_STAGES * ring_a_stage_floats(s.m) * sizeof(float), rb = RING_B_STAGES * ring_b_stage_floats(s.m) * sizeof(float);
        lbfgs_dots_ring_kernel<M><<<grid, RING_THREADS, ra, ctx->stream>>>(s);
        lbfgs_step_kernel<M><<<s.G, LB_STEP_THREADS, b->lb_smem, ctx->stream>>>(s, LB_MAXCH);
        lbfgs_update_ring_kernel<M><<<grid, RING_THREADS, rb, ctx->stream>>>(s);
        return;
    }
    lbfgs_dots_kernel<M><<<grid, LB_THREADS, b->lb_smem, ctx->stream>>>(s);
    lbfgs_step_kernel<M><<<s.G, LB_STEP_THREADS, b->lb_smem, ctx->stream>>>(s, LB_MAXCH);
    lbfgs_update_kernel<M><<<grid, LB_THREADS, 0, ctx->stream>>>(s);
}
}  // extern "C++"

// Evaluation rounds of the segment in progress until every decoy of the queue has passed through it
// (or max_rounds).  A round = turnover (park the decoys that left the segment, refill their positions from
// the queue) -> compact -> evaluate -> L-BFGS step.
static int run_rounds(trx_fold_batch *b, int max_rounds, int check_every, int *rounds_io)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    int rc, rounds = 0;
    bool busy = true;
    int *h = b->h_poll;          // [0,16) live slots, [16,32) occupied positions, [32,48) queue cursors
    std::vector<int> ng(b->tab_ng);   // live slot groups per table block: an upper bound between polls
    std::vector<int> cap(s.ntab);     // positions of each block its unfinished decoys are spread over (all, until a migration)
    for (int t = 0; t < s.ntab; ++t) cap[t] = s.tab_n[t];
    while (busy && *rounds_io + rounds < max_rounds) {
        for (int k = 0; k < check_every && *rounds_io + rounds < max_rounds; ++k, ++rounds) {
            ctx->time_begin("turnover");
            turnover_plan_kernel<<<s.ntab, 1024, 0, ctx->stream>>>(s);
            ctx->time_end("turnover");
            ctx->time_begin("turnover");
            turnover_move_kernel<<<s.G, 128, 0, ctx->stream>>>(s);
            ctx->time_end("turnover");
            if ((rc = fold_eval(b, ng.data(), false))) return rc;   // (compact_kernel also flags the position groups that are live)
            ctx->time_begin("lbfgs");
            {   // enough CTAs for ~3 per SM, at least 4 vector elements per warp and chunk
                int nch = std::max(1, (3 * 148 + s.G - 1) / s.G);   // CTAs that share a group's LB_MAXCH chunks
                nch = std::min(nch, LB_MAXCH);
                const dim3 grid(s.G, nch);
                if (s.lb_M == 8) lbfgs_launch<8>(b, grid);
                else if (s.lb_M == 16) lbfgs_launch<16>(b, grid);
                else if (s.lb_M == 20) lbfgs_launch<20>(b, grid);
                else lbfgs_launch<24>(b, grid);
                ctx->launches += 2;
            }
            ctx->time_end("lbfgs");
        }
        TRX_CUDA(cudaMemcpyAsync(h, s.nslot, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TRX_CUDA(cudaMemcpyAsync(h + 16, s.qocc, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TRX_CUDA(cudaMemcpyAsync(h + 32, s.qcursor, 16 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        TRX_CUDA(cudaStreamSynchronize(ctx->stream));
        // counts as of the start of the last round: decoys that finished in it are parked by the next turnover
        busy = false;
        bool migrate = false;
        for (int t = 0; t < s.ntab; ++t) {
            const int occ = h[16 + t];
            const bool waiting = h[32 + t] < s.nq_tab[t];
            if (occ > 0 || waiting) busy = true;
            if (waiting) {   // freed positions are refilled: the block stays full
                ng[t] = b->tab_ng[t];
                cap[t] = s.tab_n[t];
                continue;
            }
            ng[t] = num_groups(occ);   // the queue is empty: occupied positions only ever decrease
            const int live = std::min(h[t], occ);
            // unfinished decoys fill less than mig_num/mig_den of the positions they are spread over: pack them
            if (b->migrate && live > 0 && b->mig_den * live <= b->mig_num * cap[t] && cap[t] >= 2 * LANES) { migrate = true; cap[t] = live; }
        }
        if (migrate && busy) {
            int maxn = 0;
            for (int t = 0; t < s.ntab; ++t) maxn = std::max(maxn, s.tab_n[t]);
            ctx->time_begin("migrate");
            migrate_plan_kernel<<<s.ntab, 1024, 0, ctx->stream>>>(s);
            ctx->time_end("migrate");
            ctx->time_begin("migrate");
            migrate_swap_kernel<<<dim3((maxn + 1) / 2, s.ntab), 256, 0, ctx->stream>>>(s);
            ctx->time_end("migrate");
            // a migrated position may hold a finished decoy that is not parked yet: every slot group may be live
            for (int t = 0; t < s.ntab; ++t) ng[t] = std::max(ng[t], num_groups(h[16 + t]));
        }
    }
    *rounds_io += rounds;
    return busy ? 1 : TRX_OK;   // 1: the round budget ran out with decoys still in the segment
}

static void begin_segment(trx_fold_batch *b, int lo, int hi, int cart, int last, int park_X)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    s.cart = cart; s.ndof = cart ? s.ndof_c : s.ndof_t;
    s.seg_lo = lo; s.seg_hi = hi; s.seg_last = last; s.park_X = park_X;
    ctx->time_begin("segment");
    seg_reset_kernel<<<(s.Npad + 255) / 256, 256, 0, ctx->stream>>>(s);
    ctx->time_end("segment");
}

// The whole schedule, segment by segment over the whole queue.
static int run_schedule(trx_fold_batch *b, int max_rounds, int check_every, int *rounds_io)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    int rc;
    for (size_t k = 0; k < b->segs.size(); ++k) {
        const auto &sg = b->segs[k];
        const bool last = k + 1 == b->segs.size();
        begin_segment(b, sg.lo, sg.hi, sg.cart, last ? 1 : 0, (!last && b->segs[k + 1].cart) ? 1 : 0);
        rc = run_rounds(b, max_rounds, check_every, rounds_io);
        if (rc < 0) return rc;
        if (rc == 1) {
            // Round budget exhausted.  Close the segment in progress: unfinished decoys stop at their accepted
            // point, waiting ones pass through; then (unless it was the last) one pass in which every record
            // gets its closing evaluation in torsion space and its results are written.
            ctx->time_begin("segment");
            force_final_kernel<<<s.G, 256, 0, ctx->stream>>>(s);
            ctx->time_end("segment");
            ctx->time_begin("segment");
            force_final_queue_kernel<<<(s.Nqpad + 255) / 256, 256, 0, ctx->stream>>>(s);
            ctx->time_end("segment");
            int extra = 0;
            rc = run_rounds(b, 1 << 30, check_every, &extra);
            if (rc < 0) return rc;
            if (!last) {
                begin_segment(b, s.nruns, s.nruns, 0, 1, 0);
                rc = run_rounds(b, 1 << 30, check_every, &extra);
                if (rc < 0) return rc;
            }
            break;
        }
    }
    s.cart = 0; s.ndof = s.ndof_t; s.seg_lo = 0; s.seg_hi = s.nruns;
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// Folds nq[t] decoys per table block through the batch's positions; MC options in b->s.mc.
static int fold_queue(trx_fold_batch *b, const int *nq, float *tors, float *xyz, double *terms, long long *stats, int stats_cols,
                      int max_rounds, int check_every, int *rounds_out)
{
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    TRX_CUDA(cudaSetDevice(ctx->device));
    if (check_every < 1) check_every = 16;
    int Nq = 0, Gq = 0;
    for (int t = 0; t < s.ntab; ++t) {
        TRX_REQUIRE(nq[t] >= 0, "trx_fold_run: negative decoy count for table block %d", t);
        s.nq_tab[t] = nq[t]; s.qc0[t] = Nq; s.qd0[t] = Gq * LANES;
        Nq += nq[t];
        Gq += num_groups(nq[t]);
    }
    TRX_REQUIRE(Nq > 0, "trx_fold_run: no decoys");
    s.Nqpad = Gq * LANES;
    const size_t np = (size_t)s.Nqpad;
    void *d_in = nullptr, *d_q = nullptr, *d_ot = nullptr, *d_ox = nullptr, *d_oe = nullptr, *d_os = nullptr;
    int rc;
    const size_t tb = (size_t)Nq * s.ndof_t * sizeof(float), xb = (size_t)Nq * s.L * NAT3 * sizeof(float);
    if ((rc = ctx->get_scratch("fold_tors", tb, &d_in))) return rc;
    if ((rc = ctx->get_scratch("fold_tors_out", tb, &d_ot))) return rc;
    if (xyz && (rc = ctx->get_scratch("fold_xyz_out", xb, &d_ox))) return rc;
    if ((rc = ctx->get_scratch("fold_terms", (size_t)Nq * TRX_NTERM * sizeof(double), &d_oe))) return rc;
    if ((rc = ctx->get_scratch("fold_stats", (size_t)Nq * 3 * sizeof(long long), &d_os))) return rc;
    // the queue store, one allocation
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_qt = carve(np * s.ndof_t * 4), o_qx = carve(b->has_cart ? np * s.ndof_c * 4 : 256);
    const size_t o_qe = carve(np * 8 * TRX_NTERM), o_qr = carve(np * 4), o_qh = carve(np * 4), o_qv = carve(np * 4), o_qi = carve(np * 4);
    const size_t o_qf = carve(np * 4), o_of = carve((size_t)Nq * 4);
    if ((rc = ctx->get_scratch("fold_queue", off, &d_q))) return rc;
    char *Q = (char *)d_q;
    s.q_tors = (float *)(Q + o_qt); s.q_X = (float *)(Q + o_qx); s.q_terms = (double *)(Q + o_qe);
    s.q_run = (int *)(Q + o_qr); s.q_held = (int *)(Q + o_qh); s.q_evals = (int *)(Q + o_qv); s.q_iters = (int *)(Q + o_qi);
    s.q_flags = (int *)(Q + o_qf); s.o_flags = (int *)(Q + o_of);
    b->status.assign(Nq, 0);
    s.o_tors = (float *)d_ot; s.o_xyz = (float *)d_ox; s.o_terms = (double *)d_oe; s.o_stats = (long long *)d_os;
    TRX_CUDA(cudaMemcpyAsync(d_in, tors, tb, cudaMemcpyHostToDevice, ctx->stream));
    ctx->time_begin("fold_device");   // device time of the whole fold, inputs resident (H2D done, D2H not started)
    --ctx->launches;
    TRX_CUDA(cudaMemsetAsync(s.k1count, 0, 16 * sizeof(long long), ctx->stream));
    TRX_CUDA(cudaMemsetAsync(d_os, 0, (size_t)Nq * 3 * sizeof(long long), ctx->stream));
    TRX_CUDA(cudaMemsetAsync(s.o_flags, 0xff, (size_t)Nq * sizeof(int), ctx->stream));   // -1 until the decoy's results are written
    ctx->time_begin("fold_init");
    queue_init_kernel<<<Gq, 256, 0, ctx->stream>>>(s, (const float *)d_in);
    ctx->time_end("fold_init");
    int rounds = 0;
    if ((rc = run_schedule(b, max_rounds, check_every, &rounds))) return rc;
    ctx->time_end("fold_device");
    TRX_CUDA(cudaGetLastError());
    TRX_CUDA(cudaMemcpyAsync(tors, d_ot, tb, cudaMemcpyDeviceToHost, ctx->stream));
    if (terms) TRX_CUDA(cudaMemcpyAsync(terms, d_oe, (size_t)Nq * TRX_NTERM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (xyz) TRX_CUDA(cudaMemcpyAsync(xyz, d_ox, xb, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<long long> st3;
    if (stats) {
        if (stats_cols == 3) TRX_CUDA(cudaMemcpyAsync(stats, d_os, (size_t)Nq * 3 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        else {
            st3.resize((size_t)Nq * 3);
            TRX_CUDA(cudaMemcpyAsync(st3.data(), d_os, st3.size() * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    TRX_CUDA(cudaMemcpyAsync(b->status.data(), s.o_flags, (size_t)Nq * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(b->k1_decoy_evals, s.k1count, 16 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    if (stats && stats_cols == 2)
        for (int n = 0; n < Nq; ++n) { stats[(size_t)n * 2] = st3[(size_t)n * 3]; stats[(size_t)n * 2 + 1] = st3[(size_t)n * 3 + 1]; }
    if (rounds_out) *rounds_out = rounds;
    return TRX_OK;
}

/* Runs the schedule to completion (or max_rounds evaluation rounds) for as many decoys as the batch has
 * positions.  tors: host [N][L][3] float, in: start torsions, out: final torsions.  xyz (may be NULL):
 * host [N][L][5][3] float, atoms N,CA,CB,C,O.  terms (may be NULL): [N][8] double.  stats (may be NULL):
 * [N][2] evaluations, accepted iterations.  *rounds_out (may be NULL): evaluation rounds executed. */
int trx_fold_run(trx_fold_batch *b, float *tors, float *xyz, double *terms, long long *stats, int max_rounds,
                 int check_every, int *rounds_out)
{
    TRX_REQUIRE(b && tors, "trx_fold_run: NULL argument");
    b->s.mc = McOpts{};
    return fold_queue(b, b->s.tab_n, tors, xyz, terms, stats, 2, max_rounds, check_every, rounds_out);
}

/* Continuous batching: folds nq[t] decoys against table block t -- any number, more than the batch has
 * positions -- keeping the positions full: a position whose decoy has finished the schedule segment in
 * progress is refilled with the next waiting decoy in the same evaluation round.  Arrays as trx_fold_run
 * with N = sum nq[t], decoys of block 0 first.  A decoy's result does not depend on the position it
 * occupied, nor on nq or the batch size (bit for bit). */
int trx_fold_run_queue(trx_fold_batch *b, const int *nq, float *tors, float *xyz, double *terms, long long *stats,
                       int max_rounds, int check_every, int *rounds_out)
{
    TRX_REQUIRE(b && nq && tors, "trx_fold_run_queue: NULL argument");
    b->s.mc = McOpts{};
    return fold_queue(b, nq, tors, xyz, terms, stats, 2, max_rounds, check_every, rounds_out);
}

/* Restraint-kernel work of the last trx_fold_run* call on this batch: decoy evaluations the kernel made
 * per table block (the evaluations of vdw-only runs skip it).  out: [ntab]. */
int trx_fold_k1_evals(trx_fold_batch *b, long long *out)
{
    TRX_REQUIRE(b && out, "trx_fold_k1_evals: NULL argument");
    for (int t = 0; t < b->s.ntab; ++t) out[t] = b->k1_decoy_evals[t];
    return TRX_OK;
}

/* Failure reporting (the reference has none: a failed child process is a missing PDB file, utils_trX2dy/utils.py:491-498).
 * Per decoy of the last trx_fold_run* / trx_fold_mc* call on this batch, caller order: 0 = clean, else TRX_DECOY_* bits.
 * A caller re-seeds the decoys it does not accept (sampler.fold does).  out: [n], n <= decoys of the call. */
int trx_fold_status(trx_fold_batch *b, int *out, int n)
{
    TRX_REQUIRE(b && out && n >= 0 && (size_t)n <= b->status.size(), "trx_fold_status: bad argument (the last call had %zu decoys)", b ? b->status.size() : (size_t)0);
    for (int k = 0; k < n; ++k) out[k] = b->status[k];
    return TRX_OK;
}

/* Monte-Carlo sampling on top of the fold (extension, BASELINE config 4): minimise through
 * runs [0, mc_run) exactly as trx_fold_run, then `cycles` times { perturb phi/psi of a random
 * block of block_min..block_max residues by N(0, sigma_deg); re-minimise with run mc_run;
 * Metropolis at temperature kT on that run's weighted score }.  The cycles are part of each decoy's own
 * state machine: no batch-wide barrier per cycle.  stats: [N][3] = evaluations, accepted L-BFGS
 * iterations, accepted MC moves.  id_offset: global index of decoy 0 (keeps random streams independent
 * of the sharding).  nq: decoys per table block as trx_fold_run_queue, or NULL = one per position. */
int trx_fold_mc_queue(trx_fold_batch *b, const int *nq, float *tors, float *xyz, double *terms, long long *stats, int mc_run,
                      int cycles, double kT, int block_min, int block_max, double sigma_deg, unsigned long long seed,
                      unsigned long long id_offset, int max_rounds, int check_every, int *rounds_out)
{
    TRX_REQUIRE(b && tors, "trx_fold_mc: NULL argument");
    FoldState &s = b->s;
    TRX_REQUIRE(mc_run >= 1 && mc_run == s.nruns - 1, "trx_fold_mc: mc_run must be the last run of the schedule (got %d of %d)", mc_run, s.nruns);
    TRX_REQUIRE(cycles >= 0 && kT > 0 && block_min >= 1 && block_max >= block_min && sigma_deg >= 0, "trx_fold_mc: bad options");
    TRX_REQUIRE(!b->segs.back().cart, "trx_fold_mc: the Monte-Carlo run must be a torsion-space run");
    McOpts o;
    o.seed = seed; o.id_offset = id_offset; o.block_min = block_min; o.block_max = block_max; o.mc_run = mc_run;
    o.sigma = (float)(sigma_deg * TRX_DEG); o.kT = (float)kT; o.cycles = cycles;
    s.mc = o;
    const int rc = fold_queue(b, nq ? nq : s.tab_n, tors, xyz, terms, stats, 3, max_rounds, check_every, rounds_out);
    s.mc = McOpts{};
    return rc;
}

int trx_fold_mc(trx_fold_batch *b, float *tors, float *xyz, double *terms, long long *stats, int mc_run, int cycles,
                double kT, int block_min, int block_max, double sigma_deg, unsigned long long seed,
                unsigned long long id_offset, int max_rounds, int check_every, int *rounds_out)
{
    return trx_fold_mc_queue(b, nullptr, tors, xyz, terms, stats, mc_run, cycles, kT, block_min, block_max, sigma_deg, seed,
                             id_offset, max_rounds, check_every, rounds_out);
}

/* Single evaluation at given torsions under uniform weights (parity tests of K2-K4):
 * tors host [N][L][3] float -> total [N], terms [N][8], gtors [N][L][3] float, xyz [N][L][5][3] float. */
int trx_fold_eval(trx_fold_batch *b, const float *tors, const double w[TRX_NTERM], double *total, double *terms, float *gtors, float *xyz)
{
    TRX_REQUIRE(b && tors && w, "trx_fold_eval: NULL argument");
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    TRX_CUDA(cudaSetDevice(ctx->device));
    s.cart = 0; s.ndof = s.ndof_t; s.seg_lo = 0; s.seg_hi = s.nruns;
    void *d_tors = nullptr;
    int rc;
    const size_t tb = (size_t)s.N * s.ndof_t * sizeof(float);
    if ((rc = ctx->get_scratch("fold_tors", tb, &d_tors))) return rc;
    TRX_CUDA(cudaMemcpyAsync(d_tors, tors, tb, cudaMemcpyHostToDevice, ctx->stream));
    init_state_kernel<<<s.G, 32, 0, ctx->stream>>>(s, (const float *)d_tors);
    std::vector<float> wl((size_t)TRX_NTERM * s.Npad);
    for (int k = 0; k < TRX_NTERM; ++k) for (int n = 0; n < s.Npad; ++n) wl[(size_t)k * s.Npad + n] = (float)w[k];
    TRX_CUDA(cudaMemcpyAsync(s.wl, wl.data(), wl.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    ctx->launches += 1;
    if ((rc = fold_eval(b, nullptr, true))) return rc;
    std::vector<double> ft(s.Npad), tr((size_t)TRX_NTERM * s.Npad);
    std::vector<float> gt((size_t)s.G * s.ndof * LANES);
    TRX_CUDA(cudaMemcpyAsync(ft.data(), s.ft, ft.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(tr.data(), s.terms, tr.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(gt.data(), s.gt, gt.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (xyz) TRX_CUDA(cudaMemcpy2DAsync(xyz, NAT3 * sizeof(float), s.xnat, NATP * sizeof(float), NAT3 * sizeof(float), (size_t)s.N * s.L, cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int n = 0; n < s.N; ++n) {
        if (total) total[n] = ft[n];
        if (terms) for (int k = 0; k < TRX_NTERM; ++k) terms[(size_t)n * TRX_NTERM + k] = tr[(size_t)k * s.Npad + n];
        if (gtors) for (int k = 0; k < s.ndof; ++k) gtors[(size_t)n * s.ndof + k] = gt[((size_t)(n / LANES) * s.ndof + k) * LANES + n % LANES];
    }
    return TRX_OK;
}

/* Single Cartesian-mode evaluation (parity entry of the Cartesian stage): xyz host
 * [N][L][5][3] float are the degrees of freedom -> total [N], terms [N][8], grad [N][L][5][3]
 * float (gradient of the weighted total w.r.t. every coordinate), tors [N][L][3] float (the
 * torsions read back from the coordinates).  Any output may be NULL.  The batch must have
 * been created with a schedule that contains a Cartesian run. */
int trx_fold_eval_cart(trx_fold_batch *b, const float *xyz, const double w[TRX_NTERM], double *total, double *terms, float *grad,
                       float *tors)
{
    TRX_REQUIRE(b && xyz && w, "trx_fold_eval_cart: NULL argument");
    TRX_REQUIRE(b->has_cart, "trx_fold_eval_cart: the batch's schedule has no Cartesian run (buffers are sized for torsions)");
    trx_ctx *ctx = b->ctx;
    FoldState &s = b->s;
    TRX_CUDA(cudaSetDevice(ctx->device));
    void *d_tors = nullptr, *d_xyz = nullptr;
    int rc;
    const size_t tb = (size_t)s.N * s.ndof_t * sizeof(float), xb = (size_t)s.N * s.L * NAT3 * sizeof(float);
    if ((rc = ctx->get_scratch("fold_tors", tb, &d_tors))) return rc;
    if ((rc = ctx->get_scratch("fold_xyz", xb, &d_xyz))) return rc;
    TRX_CUDA(cudaMemsetAsync(d_tors, 0, tb, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(d_xyz, xyz, xb, cudaMemcpyHostToDevice, ctx->stream));
    s.cart = 0; s.ndof = s.ndof_t; s.seg_lo = 0; s.seg_hi = s.nruns;
    init_state_kernel<<<s.G, 32, 0, ctx->stream>>>(s, (const float *)d_tors);
    std::vector<float> wl((size_t)TRX_NTERM * s.Npad);
    for (int k = 0; k < TRX_NTERM; ++k) for (int n = 0; n < s.Npad; ++n) wl[(size_t)k * s.Npad + n] = (float)w[k];
    TRX_CUDA(cudaMemcpyAsync(s.wl, wl.data(), wl.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    s.cart = 1; s.ndof = s.ndof_c;
    TRX_CUDA(cudaMemsetAsync(s.xt, 0, (size_t)s.G * s.ndof_c * LANES * sizeof(float), ctx->stream));
    if ((rc = trx_to_grouped(ctx, s.N, s.L, TRX_NAT, TRX_F32, d_xyz, s.xt))) return rc;
    ctx->launches += 1;
    rc = fold_eval(b, nullptr, true);
    if (!rc && tors) {
        cart_readback_kernel<<<s.G, 256, 0, ctx->stream>>>(s);
        ctx->launches += 1;
    }
    s.cart = 0; s.ndof = s.ndof_t;
    if (rc) return rc;
    std::vector<double> ft(s.Npad), tr((size_t)TRX_NTERM * s.Npad);
    std::vector<float> gt((size_t)s.G * s.ndof_c * LANES), tx;
    TRX_CUDA(cudaMemcpyAsync(ft.data(), s.ft, ft.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(tr.data(), s.terms, tr.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(gt.data(), s.gt, gt.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tors) {
        tx.resize((size_t)s.G * s.ndof_t * LANES);
        TRX_CUDA(cudaMemcpyAsync(tx.data(), s.x, tx.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int n = 0; n < s.N; ++n) {
        const size_t gb = (size_t)(n / LANES), ln = n % LANES;
        if (total) total[n] = ft[n];
        if (terms) for (int k = 0; k < TRX_NTERM; ++k) terms[(size_t)n * TRX_NTERM + k] = tr[(size_t)k * s.Npad + n];
        if (grad) for (int k = 0; k < s.L * NAT3; ++k) grad[(size_t)n * s.L * NAT3 + k] = gt[(gb * s.ndof_c + k) * LANES + ln];
        if (tors) for (int k = 0; k < s.ndof_t; ++k) tors[(size_t)n * s.ndof_t + k] = tx[(gb * s.ndof_t + k) * LANES + ln];
    }
    return TRX_OK;
}

}  // extern "C"
