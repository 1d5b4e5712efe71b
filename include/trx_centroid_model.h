/* Centroid backbone model: constants shared (as DATA) by the CUDA library and the CPU
 * oracle.  Plain C, header-only.
 *
 * Rosetta's own database tables for the non-restraint centroid terms (atom_vdw, Rama
 * maps, cen_hb) are not part of the reference tree, so the terms
 * below are stated approximations with the same functional role; only the three
 * constraint terms carry a parity claim.
 *
 * Ideal backbone geometry [ROSETTA-RECALL, SURVEY 8a row 12]; the virtual-CB formula is
 * the reference's own (utils_trX2dy/utils.py:132-135), i.e. the CB the network's
 * orientation labels were computed from.
 */
#ifndef TRX_CENTROID_MODEL_H
#define TRX_CENTROID_MODEL_H

#define TRX_PI 3.14159265358979323846
#define TRX_DEG (TRX_PI / 180.0)

/* bond lengths (A) and angles (rad) */
#define TRX_B_N_CA 1.458
#define TRX_B_CA_C 1.523
#define TRX_B_C_N 1.329
#define TRX_B_C_O 1.231
#define TRX_A_N_CA_C (111.2 * TRX_DEG)
#define TRX_A_CA_C_N (116.2 * TRX_DEG)
#define TRX_A_C_N_CA (121.7 * TRX_DEG)
#define TRX_A_CA_C_O (120.8 * TRX_DEG)

/* virtual CB = CB_A*(b x c) + CB_B*b + CB_C*c + CA,  b = CA-N, c = C-CA */
#define TRX_CB_A (-0.58273431)
#define TRX_CB_B (0.56802827)
#define TRX_CB_C (-0.54067466)

/* atoms per residue in every coordinate / gradient array of the fold path; the order
 * makes the atoms moved by omega(i-1), phi(i), psi(i) suffixes of the atom sequence */
enum { TRX_AT_N = 0, TRX_AT_CA = 1, TRX_AT_CB = 2, TRX_AT_C = 3, TRX_AT_O = 4, TRX_NAT = 5 };

/* energy terms of the fold path, order of every terms[] / weights[] array */
enum { TRX_T_APC = 0, TRX_T_DIH = 1, TRX_T_ANG = 2, TRX_T_VDW = 3, TRX_T_RAMA = 4, TRX_T_OMEGA = 5, TRX_T_CART = 6, TRX_T_HB = 7, TRX_NTERM = 8 };

/* amino-acid index: position in "ARNDCQEGHILKMFPSTWYV" */
#define TRX_AA_ORDER "ARNDCQEGHILKMFPSTWYV"
#define TRX_AA_GLY 7
#define TRX_AA_PRO 14
#define TRX_AA_ALA 0

/* centroid pseudo-atom: CEN = CA + TRX_CEN_S[aa]*(CB - CA) (on the CA->CB ray) */
static const double TRX_CEN_S[20] = {
    /* A     R     N     D     C     Q     E     G     H     I  */
    0.95, 2.70, 1.65, 1.65, 1.40, 2.05, 2.05, 0.95, 2.05, 1.50,
    /* L     K     M     F     P     S     T     W     Y     V  */
    1.70, 2.30, 1.95, 2.25, 1.25, 1.25, 1.25, 2.55, 2.50, 1.30};
/* soft-sphere radii (A): r_ij = r_i + r_j; backbone atoms then CEN by residue type */
static const double TRX_R_BB[5] = {1.40, 1.80, 1.80, 1.70, 1.35}; /* N CA CB C O */
static const double TRX_R_CEN[20] = {
    1.90, 2.60, 2.20, 2.20, 2.10, 2.40, 2.40, 1.90, 2.40, 2.30,
    2.40, 2.50, 2.40, 2.60, 2.10, 2.00, 2.10, 2.80, 2.70, 2.20};
#define TRX_VDW_SCALE 0.8   /* Rosetta's vdw term carries this factor */
#define TRX_VDW_MINSEP 2    /* residue pairs closer than this in sequence are skipped */
#define TRX_VDW_CA_CUTOFF 14.0 /* CA-CA distance beyond which no atom pair can touch (max reach 6.72 A per residue) */

/* Ramachandran term: -ln of a von-Mises mixture, classes: 0 general, 1 proline.
 * Each basin: phi0, psi0 (deg), kappa_phi, kappa_psi (1/rad^2), weight. */
#define TRX_RAMA_NB 5
static const double TRX_RAMA[2][TRX_RAMA_NB][5] = {
    {{-63.0, -43.0, 9.0, 9.0, 0.45}, {-120.0, 130.0, 4.0, 4.0, 0.30}, {-70.0, 145.0, 8.0, 8.0, 0.17},
     {57.0, 39.0, 12.0, 12.0, 0.03}, {-90.0, 0.0, 6.0, 6.0, 0.05}},
    {{-65.0, -35.0, 20.0, 9.0, 0.40}, {-65.0, 145.0, 20.0, 8.0, 0.55}, {-85.0, 70.0, 12.0, 6.0, 0.05},
     {-65.0, -35.0, 20.0, 9.0, 0.0}, {-65.0, -35.0, 20.0, 9.0, 0.0}}};
#define TRX_RAMA_FLOOR 1e-4
/* Rosetta's rama energy of a residue is -ln P(bin) MINUS the entropy of the map (Ramachandran.cc:
 * ram_energ = -log(prob) + sum P log P), i.e. negative in favourable regions and zero on average
 * over the map.  The same normalisation for the mixture above, over 10 x 10 degree bins:
 * E = -ln P(phi,psi) - TRX_RAMA_OFFSET[class], offset = ln(bin area / integral of P) + entropy
 * (2.125 general, 1.956 proline; minimum of E = -1.34 at the helix centre).  A constant: it moves
 * no gradient, but it is what lets remove_clash's test rama + vdw < 10 (utils_ros.py:699-703) pass
 * for a clash-free chain, as it does in the reference. */
static const double TRX_RAMA_OFFSET[2] = {2.125029287790165, 1.956450144676099};
/* omega tether: 0.01 * (deviation from 180 in degrees)^2 */
#define TRX_OMEGA_K 0.01


/* Backbone hydrogen-bond term.  Stands in for cen_hb (weight 5 in scorefxn.wts / scorefxn1.wts, with -hb_cen_soft,
 * folding.py:48) in the centroid stages and for hbond_sr_bb + hbond_lr_bb (weight 3 each, scorefxn_cart.wts) in the
 * Cartesian stage; Rosetta's own potentials are database-driven and not in the reference tree, so this is a stated
 * approximation with the same role: a smooth, orientation-dependent attraction between backbone N-H and C=O.
 *   donor residue i (i >= 1, not Pro), acceptor residue j, |i - j| >= TRX_HB_MINSEP
 *   r = O_j - N_i, d = |r|;  v = 2 N_i - C_{i-1} - CA_i (direction of the N-H bond);  q = O_j - C_j
 *   c1 = unit(v).unit(r)   (N-H...O linearity),   c2 = -unit(q).unit(r)   (C=O...N linearity)
 *   E = -TRX_HB_EPS f(d) g(c1) g(c2),   f(d) = (1 - ((d - D0)/W)^2)^2 for |d - D0| < W else 0,   g(c) = max(c, 0)^2
 * C1-continuous with compact support (N...O within D0 +- W = 2.0 .. 3.8 A). */
#define TRX_HB_EPS 1.0
#define TRX_HB_D0 2.9
#define TRX_HB_W 0.9
#define TRX_HB_MINSEP 3

/* Cartesian stage (min_mover_cart, folding.py:100-102,170): a cart_bonded-like term keeps the
 * backbone near ideal geometry while xyz are the degrees of freedom.  Rosetta's cart_bonded
 * parameter database is not in the reference tree: harmonic springs with round constants.
 *   bonds   E = KB (d - d0)^2      N-CA, CA-C, C-O, C-N(+1)
 *   angles  E = KA (t - t0)^2      N-CA-C, CA-C-O, CA-C-N(+1), O-C-N(+1), C-N(+1)-CA(+1)
 *   CB      E = KCB |CB - vCB|^2   tether to the virtual-CB position of (N, CA, C)
 *   planar  E = KPL t^2, t = ((CA-C) x (N(+1)-C)) . (O-C)   carbonyl O in the peptide plane */
#define TRX_CART_KB 300.0
#define TRX_CART_KA 80.0
#define TRX_CART_KCB 300.0
#define TRX_CART_KPL 40.0
#define TRX_A_O_C_N (123.0 * TRX_DEG)

#endif
