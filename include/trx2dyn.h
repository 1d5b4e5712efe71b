/* trx2dyn -- C ABI of the B200-native folding hot path of trRosettaX2-Dynamics.
 *
 * The reference exposes this path as a PROCESS boundary, not an FFI
 * (utils_trX2dy/utils.py:484-505 shells out to folding/folding.py once per decoy);
 * inside that process all arithmetic is PyRosetta.  Each entry point below names the
 * reference call it replaces.  Plain C: opaque handles, caller-owned buffers, every
 * function returns 0 on success and a negative trx_status on failure
 * (trx_last_error() gives the message for the calling thread).  A handle is bound to
 * one device and one stream; handles are thread-compatible, not thread-safe.
 *
 * Layouts.  "natural" host layout of backbone coordinates: [decoy][residue][atom][xyz]
 * with atoms (N, CA, CB) for the restraint entry points and (N, CA, C, O, CB) for
 * folded decoys.  Device-resident ("grouped") layout, used between kernels:
 * [decoy/32][residue (padded to 16)][atom*3+xyz][decoy%32] so that the 32 lanes of a
 * warp are 32 decoys and every load/store is one full line.
 */
#ifndef TRX2DYN_H
#define TRX2DYN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRX_ABI_VERSION 2

typedef enum {
    TRX_OK = 0,
    TRX_ERR_INVALID = -1,   /* bad argument */
    TRX_ERR_CUDA = -2,      /* CUDA runtime error (message has the CUDA string) */
    TRX_ERR_NOMEM = -3,
    TRX_ERR_STATE = -4      /* call order / handle mismatch */
} trx_status;

typedef enum { TRX_F64 = 64, TRX_F32 = 32 } trx_precision;

/* restraint types, in the order every array of 4 below uses */
enum { TRX_DIST = 0, TRX_OMEGA = 1, TRX_THETA = 2, TRX_PHI = 3 };
/* score terms, the order of every E[3] / w[3]:
 * atom_pair_constraint, dihedral_constraint, angle_constraint */
enum { TRX_APC = 0, TRX_DIH = 1, TRX_ANG = 2 };

typedef struct trx_ctx trx_ctx;
typedef struct trx_tables trx_tables;

/* One restraint type of one target: n restraints on residue pairs (a[k], b[k])
 * (0-based), all sharing the K spline knots x[K]; y[n][K] are the knot energies.
 * Replaces: the per-restraint text files gen_rst writes (folding/utils_ros/
 * utils_ros.py:68-73, 91-95, 110-114, 134-138) plus Rosetta's SplineFunc::read_data. */
typedef struct {
    int n;
    const int32_t *a;
    const int32_t *b;
    int K;
    const double *x;
    const double *y;
} trx_rst_set;

int trx_abi_version(void);
const char *trx_last_error(void);

/* device: CUDA ordinal.  stream: a cudaStream_t (e.g. torch's current stream) or NULL
 * for a stream the context creates and owns. */
int trx_ctx_create(int device, void *stream, trx_ctx **out);
/* Tables, fold batches and dynamics states keep their context alive: the context is released with the last of them,
 * whichever order the caller (or its garbage collector) destroys them in; the handle itself is dead after this call.
 * Device blocks of destroyed tables / batches stay with the context for the next create (a dynamics loop rebuilds
 * both every iteration) and go back to the driver with the context. */
int trx_ctx_destroy(trx_ctx *ctx);
int trx_ctx_sync(trx_ctx *ctx);
/* Per-kernel device timing (CUDA events on the context's stream around each launch).  enabled: 0 off, 1 every kernel,
 * 2 only the restraint kernel ("restraints") and the whole fold ("fold_device") -- two event records per launch are
 * not free when a round is a dozen launches of a few microseconds.
 * name: "restraints", "reduce", "nerf", "centroid", "torsion_grad", "lbfgs", "turnover", "fold_device", ... */
int trx_ctx_set_timing(trx_ctx *ctx, int enabled);
int trx_ctx_get_timing(trx_ctx *ctx, const char *name, double *total_ms, long long *launches);
int trx_ctx_reset_timing(trx_ctx *ctx);
/* Number of kernels this library has launched on ctx since creation. */
long long trx_ctx_launch_count(trx_ctx *ctx);

/* Fit the clamped cubic splines (fp64, on device) and pack the tables + pair tiles.
 * Replaces: add_rst -> ConstraintSetMover.apply (utils_ros.py:706-743), i.e. Rosetta
 * parsing the .cst file and fitting one SplineFunc per line.  sets[t].n may be 0. */
int trx_tables_create(trx_ctx *ctx, int L, const trx_rst_set sets[4], trx_tables **out);
int trx_tables_destroy(trx_tables *t);
/* Atom the AtomPair (distance) restraints sit on: TRX_ATOM_CB (default; 'AtomPair CB a CB b',
 * utils_ros.py:73) or TRX_ATOM_CA ('AtomPair CA a CA b' of the -r af2 variant, utils_ros.py:191).
 * CA is only accepted for tables without angular restraints, as gen_rst_af2 produces. */
enum { TRX_ATOM_CA = 1, TRX_ATOM_CB = 2 };
int trx_tables_set_dist_atom(trx_tables *t, int atom);
/* counts[4] = restraints per type; *tiles = active 16x16 residue-pair tiles. */
int trx_tables_info(const trx_tables *t, int *L, int counts[4], int *tiles);
/* Copies the fitted second derivatives of type `type` back: y2[n][K] (for parity tests). */
int trx_tables_get_y2(trx_tables *t, int type, double *y2);

/* Restraint energies + analytic gradient for N decoys, host buffers in natural layout.
 * xyz: [N][L][3 atoms N,CA,CB][3], double (TRX_F64) or float (TRX_F32).
 * E: [N][3] unweighted term sums (always double).  grad (may be NULL): same layout and
 * type as xyz, gradient of w[0]*E_apc + w[1]*E_dih + w[2]*E_ang.
 * Replaces: ScoreFunction(pose) restricted to atom_pair_constraint, dihedral_constraint,
 * angle_constraint and the matching derivative pass inside MinMover (folding.py:164-171). */
int trx_energy_grad(trx_ctx *ctx, trx_tables *t, int N, int precision, const void *xyz,
                    const double w[3], double *E, void *grad);

/* Same, device-resident in the grouped layout (see top).  d_xyz: [ceil(N/32)][Lpad][9][32]
 * with Lpad = trx_padded_length(L); d_E: [3][32*ceil(N/32)] double; d_grad like d_xyz. */
int trx_energy_grad_device(trx_ctx *ctx, trx_tables *t, int N, int precision, const void *d_xyz,
                           const double w[3], double *d_E, void *d_grad);
int trx_padded_length(int L);
/* natural <-> grouped conversion on device (n_atoms atoms per residue, same precision both sides) */
int trx_to_grouped(trx_ctx *ctx, int N, int L, int n_atoms, int precision, const void *d_natural, void *d_grouped);
int trx_from_grouped(trx_ctx *ctx, int N, int L, int n_atoms, int precision, const void *d_grouped, void *d_natural);

/* ------------------------------------------------------------------ the centroid fold
 * Energy terms of the fold, order of every w[8] / terms[8]:
 * atom_pair_constraint, dihedral_constraint, angle_constraint, vdw, rama, omega, cart_bonded, and the backbone
 * hydrogen-bond term that stands in for cen_hb (centroid stages) and hbond_sr_bb + hbond_lr_bb (Cartesian stage).
 * Only the three constraint terms carry a parity claim; the others are stated approximations of Rosetta's
 * database-driven terms (include/trx_centroid_model.h). */
enum { TRX_TERM_APC = 0, TRX_TERM_DIH = 1, TRX_TERM_ANG = 2, TRX_TERM_VDW = 3, TRX_TERM_RAMA = 4, TRX_TERM_OMEGA = 5, TRX_TERM_CART = 6, TRX_TERM_HB = 7, TRX_NTERMS = 8 };

/* One MinMover.apply of the reference's schedule (folding/folding.py:91-104): score
 * weights (data/ *.wts), max_iter, tolerance.  clash_check = 1 restates remove_clash
 * (utils_ros.py:699-703): when rama+vdw (weights 1,1) < clash_thr at the start of the run,
 * execution continues at run skip_to instead. */
typedef struct {
    double w[8];
    int max_iter;
    double tol;
    int clash_check;
    double clash_thr;
    int skip_to;
    int cartesian; /* 1: MinMover.cartesian(True) (min_mover_cart, folding.py:100-102): the coordinates of
                    * N,CA,CB,C,O are the degrees of freedom.  Afterwards the decoy HOLDS those coordinates:
                    * a following clash check scores them, and the next torsion-space run that actually
                    * starts rebuilds the chain from the read-back torsions with ideal bond geometry. */
} trx_run;

typedef struct trx_fold_batch trx_fold_batch;

/* A batch of N = sum(ndecoys) decoys of one target; decoy block t is folded against
 * tabs[t] (two-model mixing: two table sets, half the decoys each).  All blocks but the
 * last must be multiples of 32 decoys.  aa[L]: residue types (index into
 * "ARNDCQEGHILKMFPSTWYV"), Gly already mapped to Ala as folding.py:112-115 does.
 * Replaces: pose_from_sequence + MoveMap/MinMover construction (folding.py:74-115). */
int trx_fold_create(trx_ctx *ctx, int ntab, trx_tables *const *tabs, const int *ndecoys, const int32_t *aa,
                    const trx_run *runs, int nruns, int lbfgs_m, trx_fold_batch **out);
int trx_fold_destroy(trx_fold_batch *b);
/* Minimises every decoy through the schedule, all on device (NeRF, restraint + centroid
 * terms, torsion gradient, L-BFGS / Armijo); the host only polls a counter every
 * check_every evaluation rounds.  tors: host [N][L][3] float (phi,psi,omega radians), in/out.
 * xyz (may be NULL): [N][L][5][3] float, atoms N,CA,CB,C,O.  terms (may be NULL): [N][8].
 * stats (may be NULL): [N][2] = energy evaluations, accepted L-BFGS iterations.
 * Replaces: remove_clash + repeat_mover.apply + remove_clash (folding.py:119,164-171). */
int trx_fold_run(trx_fold_batch *b, float *tors, float *xyz, double *terms, long long *stats, int max_rounds,
                 int check_every, int *rounds_out);
/* Continuous batching: folds nq[t] decoys against table block t -- any number, normally many more than the
 * batch has positions (ndecoys[t] of trx_fold_create) -- keeping the positions full: a position whose decoy
 * has finished the schedule segment in progress (torsion-space runs | min_mover_cart | torsion-space runs)
 * is refilled with the next waiting decoy in the same evaluation round; between segments a decoy lives in a
 * device-side queue record (torsions, held coordinates, terms, counters).  Arrays as trx_fold_run with
 * N = sum nq[t], the decoys of block 0 first.  A decoy's result does not depend on the position it occupied,
 * on nq or on the batch size, bit for bit.
 * Replaces: folding_with_pred_npz(repeat=N) (utils_trX2dy/utils.py:484-505), whose ThreadPoolExecutor keeps
 * min(32, cpu+4) one-decoy processes in flight until the N decoys are done. */
int trx_fold_run_queue(trx_fold_batch *b, const int *nq, float *tors, float *xyz, double *terms, long long *stats,
                       int max_rounds, int check_every, int *rounds_out);
/* Restraint-kernel work of the last trx_fold_run / trx_fold_run_queue / trx_fold_mc call on this batch: decoy
 * evaluations the restraint kernel made per table block (evaluations of vdw-only runs skip it).  out: [ntab].
 * Measurement aid (roofline accounting of bench.py); no reference counterpart. */
int trx_fold_k1_evals(trx_fold_batch *b, long long *out);
/* Failure reporting.  The reference has no error convention on this path: a decoy whose child process failed is a
 * missing PDB file that raises later in the parent (utils_trX2dy/utils.py:491-498).  Here every decoy of
 * the last trx_fold_run / trx_fold_run_queue / trx_fold_mc* call reports, in the caller's order, 0 (clean) or a
 * combination of the bits below; the host driver re-seeds decoys that report TRX_DECOY_NONFINITE. */
enum {
    TRX_DECOY_NONFINITE = 1,  /* a run started from, or the decoy ended with, a non-finite energy term */
    TRX_DECOY_LINESEARCH = 2, /* a line search failed from steepest descent (the run was cut short there) */
    TRX_DECOY_UNFINISHED = 4  /* the round budget (max_rounds) ran out before the decoy finished its schedule */
};
int trx_fold_status(trx_fold_batch *b, int *out, int n);
/* Monte-Carlo sampling on top of the fold -- an EXTENSION with no reference behaviour
 * (BASELINE config 4; the reference's folding/ has no Metropolis step).
 * Minimises through the whole schedule, whose LAST run (index mc_run) defines the MC score;
 * then `cycles` times: perturb phi/psi of a random block of block_min..block_max residues by
 * N(0, sigma_deg), re-minimise with run mc_run, Metropolis accept at temperature kT.
 * Perturbation, minimisation and acceptance run on device; the generator is counter-based
 * and keyed by (seed, id_offset + decoy index), so results do not depend on the sharding.
 * stats (may be NULL): [N][3] = evaluations, accepted L-BFGS iterations, accepted MC moves. */
int trx_fold_mc(trx_fold_batch *b, float *tors, float *xyz, double *terms, long long *stats, int mc_run, int cycles,
                double kT, int block_min, int block_max, double sigma_deg, unsigned long long seed,
                unsigned long long id_offset, int max_rounds, int check_every, int *rounds_out);
/* trx_fold_mc over a queue of nq[t] decoys per table block (see trx_fold_run_queue; nq NULL = one decoy per
 * position).  The Monte-Carlo cycles are part of each decoy's own state machine: a decoy that finishes a
 * minimisation is judged, perturbed and restarted in the same evaluation round. */
int trx_fold_mc_queue(trx_fold_batch *b, const int *nq, float *tors, float *xyz, double *terms, long long *stats, int mc_run,
                      int cycles, double kT, int block_min, int block_max, double sigma_deg, unsigned long long seed,
                      unsigned long long id_offset, int max_rounds, int check_every, int *rounds_out);
/* One evaluation at given torsions under uniform weights (parity entry for the NeRF /
 * vdw / rama / omega / torsion-gradient kernels): total[N], terms[N][8], gtors[N][L][3],
 * xyz[N][L][5][3] (any output may be NULL). */
int trx_fold_eval(trx_fold_batch *b, const float *tors, const double w[8], double *total, double *terms, float *gtors,
                  float *xyz);

/* One Cartesian-mode evaluation under uniform weights (parity entry for the Cartesian stage:
 * cart_bonded springs, rama / omega from coordinates, restraint + vdw gradients on xyz):
 * xyz[N][L][5][3] in; total[N], terms[N][8], grad[N][L][5][3], tors[N][L][3] (torsions read
 * back from the coordinates) out, any of which may be NULL.  The batch's schedule must
 * contain a Cartesian run. */
int trx_fold_eval_cart(trx_fold_batch *b, const float *xyz, const double w[8], double *total, double *terms, float *grad,
                       float *tors);

/* ------------------------------------------------------------------ the outer dynamics loop
 * Device-resident distograms of one chain of run_inference.generate_npz_and_pdb (run_inference.py:97-139).
 * dist [L][L][37] float; omega, theta [L][L][25], phi [L][L][13] (all three or none: --no-angle).
 * Replaces: the npz files the reference rewrites every iteration (run_inference.py:116-133). */
typedef struct trx_dyn trx_dyn;
int trx_dyn_create(trx_ctx *ctx, int L, const float *dist, const float *omega, const float *theta, const float *phi,
                   trx_dyn **out);
int trx_dyn_destroy(trx_dyn *d);
/* One iteration: the decoy just folded (backbone n, ca, c and cb, [L][3] double each; cb is used where use_cb[i] != 0
 * -- the reference takes the file's CB for non-Gly residues, utils_trX2dy/utils.py:145-150 -- and the virtual CB
 * elsewhere) is turned into realised 6D bins, and the four maps and the un-normalised `tmp` map are decayed,
 * renormalised and Gaussian-smoothed in place, bit-identically to the reference's numpy / scipy arithmetic.
 * w9: the nine Gaussian taps (scipy.ndimage, sigma as chosen, truncate 4).  *max_tmp_change: the convergence signal.
 * Replaces: get_neighbors + pros + process_distribution_with_pred_distribution
 * (utils_trX2dy/utils.py:125-249,379-403) behind get_npz_from_pred_pdb (:406-476). */
int trx_dyn_step(trx_dyn *d, const double *n, const double *ca, const double *c, const double *cb, const unsigned char *use_cb,
                 const double *w9, double *max_tmp_change);
/* Current maps to the host (any pointer may be NULL); bins: [4][L][L] realised bins of the last step. */
int trx_dyn_get(trx_dyn *d, float *dist, float *omega, float *theta, float *phi, float *tmp, int *bins);

/* ------------------------------------------------------------------ decoy-set metrics
 * GloCon matrix of M decoys of L residues: out[M][M], out[i][j] = mean over residue pairs a<b of
 * |d_i(a,b) - d_j(a,b)| where that exceeds thr (3 A), d = CB-CB distance or 0 beyond dmax (20 A).
 * cb: host [M][L][3] double (virtual CB for Gly, as get_neighbors builds it).
 * Replaces: get_glocon_matrix (utils_trX2dy/utils.py:543-567). */
int trx_glocon_matrix(trx_ctx *ctx, int M, int L, const double *cb, double dmax, double thr, double *out);
/* TM-score and RMSD of every ordered decoy pair from CA traces with identical residue numbering:
 * tm[i][j] = TM-score of decoy i superposed on decoy j normalised by L, rmsd[i][j] = Kabsch RMSD.
 * ca: host [M][L][3] double.  Replaces: get_tmscore_and_rmsd_matrix (utils_trX2dy/utils.py:514-540),
 * i.e. one `./bin/TMscore a.pdb b.pdb` subprocess per pair. */
int trx_tmscore_matrix(trx_ctx *ctx, int M, int L, const double *ca, double *tm, double *rmsd);

#ifdef __cplusplus
}
#endif
#endif /* TRX2DYN_H */
