/* CPU ORACLE (test infrastructure, NOT the product) -- the centroid fold.
 *
 * fp64, scalar C restatement of what the B200 library computes for one decoy:
 * NeRF backbone building from phi/psi/omega, the non-restraint centroid terms
 * (soft-sphere vdw, Ramachandran, omega tether -- stated APPROXIMATIONS of Rosetta's
 * vdw / rama / omega, whose database tables are not in the reference tree), the
 * reverse-mode torsion gradient (Abe-Go / Rosetta F1,F2 suffix sums), an L-BFGS with
 * non-monotone Armijo back-tracking, the Cartesian stage (min_mover_cart,
 * folding/folding.py:100-102,170: xyz as degrees of freedom, a cart_bonded-like spring
 * term, rama/omega evaluated from coordinates) and the reference's staged schedule
 * (folding/folding.py:86-104,118-119,164-171; utils_ros.py:699-703).
 * PARITY UNPINNED against PyRosetta (absent); the three constraint terms it calls are
 * the ones of restraints_oracle.c.  Only tests/, smoke() and bench.py's CPU-baseline /
 * --impl reference legs may use this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/trx_centroid_model.h"

/* from restraints_oracle.c */
typedef struct {
    int n, K;
    const int *a, *b;
    const double *x, *y, *y2;
} trxo_set;
void trxo_energy_grad(int L, const double *xyz, const trxo_set *dist, const trxo_set *omega, const trxo_set *theta,
                      const trxo_set *phi, const double *w, double *E, double *grad, double **val, double **ener);

static void v_sub(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double v_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v_cross(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static void v_unit(double *a) { double n = sqrt(v_dot(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }

/* NeRF: place d so that |cd| = bond, angle(b,c,d) = ang, dihedral(a,b,c,d) = tor. */
static void place(const double *a, const double *b, const double *c, double bond, double ang, double tor, double *d)
{
    double bc[3], ab[3], n[3], m[3];
    v_sub(c, b, bc); v_unit(bc);
    v_sub(b, a, ab);
    v_cross(ab, bc, n); v_unit(n);
    v_cross(n, bc, m);
    double dx = -bond * cos(ang), dy = bond * sin(ang) * cos(tor), dz = bond * sin(ang) * sin(tor);
    for (int k = 0; k < 3; ++k) d[k] = c[k] + dx * bc[k] + dy * m[k] + dz * n[k];
}

#define XYZ(r, at) (xyz + ((size_t)(r) * TRX_NAT + (at)) * 3)

/* tors[L][3] = (phi, psi, omega) in radians -> xyz[L][5][3] in atom order N,CA,CB,C,O. */
void trxo_nerf(int L, const double *tors, double *xyz)
{
    double *n0 = XYZ(0, TRX_AT_N), *ca0 = XYZ(0, TRX_AT_CA), *c0 = XYZ(0, TRX_AT_C);
    n0[0] = n0[1] = n0[2] = 0.0;
    ca0[0] = TRX_B_N_CA; ca0[1] = ca0[2] = 0.0;
    c0[0] = TRX_B_N_CA - TRX_B_CA_C * cos(TRX_A_N_CA_C); c0[1] = TRX_B_CA_C * sin(TRX_A_N_CA_C); c0[2] = 0.0;
    for (int i = 0; i < L; ++i) {
        const double *t = tors + (size_t)i * 3;
        if (i > 0) {
            const double *tp = tors + (size_t)(i - 1) * 3;
            place(XYZ(i - 1, TRX_AT_N), XYZ(i - 1, TRX_AT_CA), XYZ(i - 1, TRX_AT_C), TRX_B_C_N, TRX_A_CA_C_N, tp[1], XYZ(i, TRX_AT_N));
            place(XYZ(i - 1, TRX_AT_CA), XYZ(i - 1, TRX_AT_C), XYZ(i, TRX_AT_N), TRX_B_N_CA, TRX_A_C_N_CA, tp[2], XYZ(i, TRX_AT_CA));
            place(XYZ(i - 1, TRX_AT_C), XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), TRX_B_CA_C, TRX_A_N_CA_C, t[0], XYZ(i, TRX_AT_C));
        }
        place(XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C), TRX_B_C_O, TRX_A_CA_C_O, t[1] + TRX_PI, XYZ(i, TRX_AT_O));
        double b[3], c[3], a[3];
        v_sub(XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_N), b);
        v_sub(XYZ(i, TRX_AT_C), XYZ(i, TRX_AT_CA), c);
        v_cross(b, c, a);
        for (int k = 0; k < 3; ++k) XYZ(i, TRX_AT_CB)[k] = TRX_CB_A * a[k] + TRX_CB_B * b[k] + TRX_CB_C * c[k] + XYZ(i, TRX_AT_CA)[k];
    }
}

/* Soft-sphere repulsion over N,CA,CB,C,O + CEN of residue pairs |i-j| >= TRX_VDW_MINSEP:
 * E = 0.8 * sum (r^2 - d^2)^2 / r^2 for d < r.  grad[L][5][3] is ACCUMULATED with w*dE/dx
 * (CEN's share is distributed to CA and CB, CEN being CA + s*(CB-CA)). */
double trxo_vdw(int L, const int *aa, const double *xyz, double w, double *grad)
{
    double E = 0.0;
    double *at = (double *)malloc(sizeof(double) * (size_t)L * 6 * 3);
    double *g = (double *)calloc((size_t)L * 6 * 3, sizeof(double));
    for (int i = 0; i < L; ++i) {
        for (int a = 0; a < 5; ++a) memcpy(at + ((size_t)i * 6 + a) * 3, XYZ(i, a), 3 * sizeof(double));
        double s = TRX_CEN_S[aa[i]];
        for (int k = 0; k < 3; ++k)
            at[((size_t)i * 6 + 5) * 3 + k] = XYZ(i, TRX_AT_CA)[k] + s * (XYZ(i, TRX_AT_CB)[k] - XYZ(i, TRX_AT_CA)[k]);
    }
    for (int i = 0; i < L; ++i)
        for (int j = i + TRX_VDW_MINSEP; j < L; ++j) {
            double dca[3];
            v_sub(XYZ(i, TRX_AT_CA), XYZ(j, TRX_AT_CA), dca);
            if (v_dot(dca, dca) > TRX_VDW_CA_CUTOFF * TRX_VDW_CA_CUTOFF) continue;
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) {
                    double ra = a < 5 ? TRX_R_BB[a] : TRX_R_CEN[aa[i]], rb = b < 5 ? TRX_R_BB[b] : TRX_R_CEN[aa[j]];
                    double r2 = (ra + rb) * (ra + rb), d[3];
                    double *pa = at + ((size_t)i * 6 + a) * 3, *pb = at + ((size_t)j * 6 + b) * 3;
                    v_sub(pa, pb, d);
                    double d2 = v_dot(d, d);
                    if (d2 >= r2) continue;
                    double c = r2 - d2;
                    E += TRX_VDW_SCALE * c * c / r2;
                    double f = -4.0 * TRX_VDW_SCALE * c / r2;  /* dE/d(d^2) * 2 */
                    for (int k = 0; k < 3; ++k) {
                        g[((size_t)i * 6 + a) * 3 + k] += f * d[k];
                        g[((size_t)j * 6 + b) * 3 + k] -= f * d[k];
                    }
                }
        }
    if (grad)
        for (int i = 0; i < L; ++i) {
            double s = TRX_CEN_S[aa[i]];
            for (int a = 0; a < 5; ++a)
                for (int k = 0; k < 3; ++k) grad[((size_t)i * TRX_NAT + a) * 3 + k] += w * g[((size_t)i * 6 + a) * 3 + k];
            for (int k = 0; k < 3; ++k) {
                double gc = w * g[((size_t)i * 6 + 5) * 3 + k];
                grad[((size_t)i * TRX_NAT + TRX_AT_CA) * 3 + k] += (1.0 - s) * gc;
                grad[((size_t)i * TRX_NAT + TRX_AT_CB) * 3 + k] += s * gc;
            }
        }
    free(at); free(g);
    return E;
}

/* Backbone hydrogen-bond term (include/trx_centroid_model.h): donors N_i (i >= 1, not Pro), acceptors O_j,
 * |i-j| >= TRX_HB_MINSEP.  grad[L][5][3] is ACCUMULATED with w*dE/dx on N_i, CA_i, C_{i-1}, O_j, C_j. */
double trxo_hbond(int L, const int *aa, const double *xyz, double w, double *grad)
{
    double E = 0.0;
    const double dmax = TRX_HB_D0 + TRX_HB_W, dmin = TRX_HB_D0 - TRX_HB_W;
    for (int i = 1; i < L; ++i) {
        if (aa[i] == TRX_AA_PRO) continue;
        const double *N = XYZ(i, TRX_AT_N), *CA = XYZ(i, TRX_AT_CA), *Cp = XYZ(i - 1, TRX_AT_C);
        double v[3];
        for (int k = 0; k < 3; ++k) v[k] = 2.0 * N[k] - Cp[k] - CA[k];
        const double vn = sqrt(v_dot(v, v));
        for (int j = 0; j < L; ++j) {
            if (abs(i - j) < TRX_HB_MINSEP) continue;
            const double *O = XYZ(j, TRX_AT_O), *C = XYZ(j, TRX_AT_C);
            double r[3], q[3];
            v_sub(O, N, r);
            const double d2 = v_dot(r, r);
            if (d2 >= dmax * dmax || d2 <= dmin * dmin) continue;
            const double d = sqrt(d2);
            v_sub(O, C, q);
            const double qn = sqrt(v_dot(q, q));
            const double c1 = v_dot(v, r) / (vn * d), c2 = -v_dot(q, r) / (qn * d);
            if (c1 <= 0.0 || c2 <= 0.0) continue;
            const double t = (d - TRX_HB_D0) / TRX_HB_W, u = 1.0 - t * t;
            const double F = u * u, dF = -4.0 * u * t / TRX_HB_W;
            const double G1 = c1 * c1, G2 = c2 * c2;
            E += -TRX_HB_EPS * F * G1 * G2;
            if (!grad) continue;
            const double a = -TRX_HB_EPS * w;
            const double kd = a * dF * G1 * G2, k1 = a * F * 2.0 * c1 * G2, k2 = a * F * G1 * 2.0 * c2;
            double gr[3], gv[3], gq[3];
            for (int k = 0; k < 3; ++k) {
                const double rh = r[k] / d, vh = v[k] / vn, qh = q[k] / qn;
                gr[k] = kd * rh + k1 * (vh - c1 * rh) / d - k2 * (qh + c2 * rh) / d;
                gv[k] = k1 * (rh - c1 * vh) / vn;
                gq[k] = -k2 * (rh + c2 * qh) / qn;
            }
            for (int k = 0; k < 3; ++k) {
                grad[((size_t)j * TRX_NAT + TRX_AT_O) * 3 + k] += gr[k] + gq[k];
                grad[((size_t)j * TRX_NAT + TRX_AT_C) * 3 + k] -= gq[k];
                grad[((size_t)i * TRX_NAT + TRX_AT_N) * 3 + k] += -gr[k] + 2.0 * gv[k];
                grad[((size_t)i * TRX_NAT + TRX_AT_CA) * 3 + k] -= gv[k];
                grad[((size_t)(i - 1) * TRX_NAT + TRX_AT_C) * 3 + k] -= gv[k];
            }
        }
    }
    return E;
}

/* Ramachandran (von-Mises mixture, residues 1..L-2 as Rosetta skips termini) and omega
 * tether (residues 0..L-2).  gtors[L][3] is ACCUMULATED with the weighted derivatives. */
void trxo_rama_omega(int L, const int *aa, const double *tors, double w_rama, double w_omega,
                     double *E_rama, double *E_omega, double *gtors)
{
    *E_rama = 0.0; *E_omega = 0.0;
    for (int i = 0; i < L; ++i) {
        const double *t = tors + (size_t)i * 3;
        if (i > 0 && i < L - 1) {
            int cls = aa[i] == TRX_AA_PRO ? 1 : 0;
            double P = TRX_RAMA_FLOOR, dP_dphi = 0.0, dP_dpsi = 0.0;
            for (int k = 0; k < TRX_RAMA_NB; ++k) {
                const double *b = TRX_RAMA[cls][k];
                double dphi = t[0] - b[0] * TRX_DEG, dpsi = t[1] - b[1] * TRX_DEG;
                double e = b[4] * exp(b[2] * (cos(dphi) - 1.0) + b[3] * (cos(dpsi) - 1.0));
                P += e;
                dP_dphi += -e * b[2] * sin(dphi);
                dP_dpsi += -e * b[3] * sin(dpsi);
            }
            *E_rama += -log(P) - TRX_RAMA_OFFSET[cls];
            if (gtors) { gtors[(size_t)i * 3 + 0] += -w_rama * dP_dphi / P; gtors[(size_t)i * 3 + 1] += -w_rama * dP_dpsi / P; }
        }
        if (i < L - 1) {
            double dev = t[2] - TRX_PI;                      /* wrap to (-pi, pi] */
            dev -= 2.0 * TRX_PI * floor((dev + TRX_PI) / (2.0 * TRX_PI));
            double deg = dev / TRX_DEG;
            *E_omega += TRX_OMEGA_K * deg * deg;
            if (gtors) gtors[(size_t)i * 3 + 2] += w_omega * 2.0 * TRX_OMEGA_K * deg / TRX_DEG;
        }
    }
}

/* Reverse mode: Cartesian gradient g[L][5][3] -> torsion gradient gt[L][3] (ACCUMULATED).
 * Atoms moved by a torsion form a suffix of the atom sequence (order N,CA,CB,C,O):
 * omega(i) (axis C_i->N_i+1) moves from CA_i+1 on, phi(i) (axis N_i->CA_i) from CB_i on,
 * psi(i) (axis CA_i->C_i) from O_i on.  dE/dt = u . (F1 - p x F2), F1 = sum x_a x g_a,
 * F2 = sum g_a over the suffix. */
void trxo_torsion_grad(int L, const double *xyz, const double *g, double *gt)
{
    double F1[3] = {0, 0, 0}, F2[3] = {0, 0, 0};
#define ADD(r, at)                                                                  \
    do {                                                                            \
        const double *x_ = XYZ(r, at), *g_ = g + ((size_t)(r) * TRX_NAT + (at)) * 3; \
        double c_[3];                                                               \
        v_cross(x_, g_, c_);                                                        \
        for (int k = 0; k < 3; ++k) { F1[k] += c_[k]; F2[k] += g_[k]; }            \
    } while (0)
#define DTOR(p_, q_, out)                                                  \
    do {                                                                   \
        double u_[3], c_[3];                                               \
        v_sub(q_, p_, u_); v_unit(u_);                                     \
        v_cross(p_, F2, c_);                                               \
        (out) += u_[0] * (F1[0] - c_[0]) + u_[1] * (F1[1] - c_[1]) + u_[2] * (F1[2] - c_[2]); \
    } while (0)
    for (int i = L - 1; i >= 0; --i) {
        ADD(i, TRX_AT_O);
        DTOR(XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C), gt[(size_t)i * 3 + 1]);          /* psi(i) */
        ADD(i, TRX_AT_C);
        ADD(i, TRX_AT_CB);
        if (i > 0) DTOR(XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), gt[(size_t)i * 3 + 0]); /* phi(i) */
        ADD(i, TRX_AT_CA);
        if (i > 0) DTOR(XYZ(i - 1, TRX_AT_C), XYZ(i, TRX_AT_N), gt[(size_t)(i - 1) * 3 + 2]); /* omega(i-1) */
        ADD(i, TRX_AT_N);
    }
#undef ADD
#undef DTOR
}

/* ------------------------------------------------------------------ target + total energy */
typedef struct {
    int L;
    const int *aa;                 /* residue types used for scoring (Gly already mapped to Ala) */
    const trxo_set *sets[4];       /* dist, omega, theta, phi; NULL if absent */
} trxo_target;

/* terms[TRX_NTERM] unweighted; returns the weighted total; gt[L][3] = d total / d torsion;
 * xyz[L][5][3] work/output buffer. */
double trxo_eval(const trxo_target *T, const double *w, const double *tors, double *terms, double *gt, double *xyz)
{
    const int L = T->L;
    trxo_nerf(L, tors, xyz);
    double *x9 = (double *)malloc(sizeof(double) * (size_t)L * 9), *g9 = (double *)malloc(sizeof(double) * (size_t)L * 9);
    double *g = (double *)calloc((size_t)L * TRX_NAT * 3, sizeof(double));
    for (int i = 0; i < L; ++i) memcpy(x9 + (size_t)i * 9, XYZ(i, 0), 9 * sizeof(double));  /* N,CA,CB are atoms 0..2 */
    trxo_energy_grad(L, x9, T->sets[0], T->sets[1], T->sets[2], T->sets[3], w, terms, g9, NULL, NULL);
    for (int i = 0; i < L; ++i) memcpy(g + (size_t)i * TRX_NAT * 3, g9 + (size_t)i * 9, 9 * sizeof(double));
    terms[TRX_T_VDW] = trxo_vdw(L, T->aa, xyz, w[TRX_T_VDW], g);
    terms[TRX_T_HB] = trxo_hbond(L, T->aa, xyz, w[TRX_T_HB], g);
    terms[TRX_T_CART] = 0.0;   /* ideal internal geometry in torsion space */
    memset(gt, 0, sizeof(double) * (size_t)L * 3);
    trxo_rama_omega(L, T->aa, tors, w[TRX_T_RAMA], w[TRX_T_OMEGA], &terms[TRX_T_RAMA], &terms[TRX_T_OMEGA], gt);
    trxo_torsion_grad(L, xyz, g, gt);
    gt[0] = 0.0;                               /* phi(0), psi(L-1), omega(L-1) move nothing scored */
    gt[(size_t)(L - 1) * 3 + 2] = 0.0;
    double tot = 0.0;
    for (int k = 0; k < TRX_NTERM; ++k) tot += w[k] * terms[k];
    free(x9); free(g9); free(g);
    return tot;
}

/* ------------------------------------------------------------------ Cartesian stage */
/* from restraints_oracle.c */
double trxo_dihedral(const double *p1, const double *p2, const double *p3, const double *p4);
void trxo_dihedral_grad(const double *p1, const double *p2, const double *p3, const double *p4, double *g1, double *g2,
                        double *g3, double *g4);
double trxo_angle(const double *p1, const double *p2, const double *p3);
void trxo_angle_grad(const double *p1, const double *p2, const double *p3, double *g1, double *g2, double *g3);

#define GRD(r, at) (grad + ((size_t)(r) * TRX_NAT + (at)) * 3)

static double spring_bond(const double *a, const double *b, double d0, double k, double w, double *ga, double *gb)
{
    double d[3];
    v_sub(a, b, d);
    double len = sqrt(v_dot(d, d)), dev = len - d0;
    if (ga) for (int c = 0; c < 3; ++c) { double f = w * 2.0 * k * dev * d[c] / len; ga[c] += f; gb[c] -= f; }
    return k * dev * dev;
}

static double spring_angle(const double *a, const double *b, const double *c, double t0, double k, double w, double *ga,
                           double *gb, double *gc)
{
    double dev = trxo_angle(a, b, c) - t0;
    if (ga) {
        double g1[3], g2[3], g3[3];
        trxo_angle_grad(a, b, c, g1, g2, g3);
        for (int q = 0; q < 3; ++q) { double f = w * 2.0 * k * dev; ga[q] += f * g1[q]; gb[q] += f * g2[q]; gc[q] += f * g3[q]; }
    }
    return k * dev * dev;
}

static void add_dihedral_grad(const double *p1, const double *p2, const double *p3, const double *p4, double f, double *g1,
                              double *g2, double *g3, double *g4)
{
    double d1[3], d2[3], d3[3], d4[3];
    trxo_dihedral_grad(p1, p2, p3, p4, d1, d2, d3, d4);
    for (int q = 0; q < 3; ++q) { g1[q] += f * d1[q]; g2[q] += f * d2[q]; g3[q] += f * d3[q]; g4[q] += f * d4[q]; }
}

/* phi(i) = dihedral(C_i-1, N_i, CA_i, C_i), psi(i) = dihedral(N_i, CA_i, C_i, N_i+1), omega(i) =
 * dihedral(CA_i, C_i, N_i+1, CA_i+1) read back from coordinates; the torsions that move nothing
 * (phi(0), omega(L-1)) are set to pi and psi(L-1) is taken from the carbonyl O (NeRF places
 * O at psi + pi). */
void trxo_torsions_from_xyz(int L, const double *xyz, double *tors)
{
    for (int i = 0; i < L; ++i) {
        double *t = tors + (size_t)i * 3;
        t[0] = i > 0 ? trxo_dihedral(XYZ(i - 1, TRX_AT_C), XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C)) : TRX_PI;
        if (i < L - 1) {
            t[1] = trxo_dihedral(XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C), XYZ(i + 1, TRX_AT_N));
            t[2] = trxo_dihedral(XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C), XYZ(i + 1, TRX_AT_N), XYZ(i + 1, TRX_AT_CA));
        } else {
            double p = trxo_dihedral(XYZ(i, TRX_AT_N), XYZ(i, TRX_AT_CA), XYZ(i, TRX_AT_C), XYZ(i, TRX_AT_O)) - TRX_PI;
            t[1] = p <= -TRX_PI ? p + 2.0 * TRX_PI : p;
            t[2] = TRX_PI;
        }
    }
}

/* cart_bonded-like springs (include/trx_centroid_model.h) + Ramachandran + omega tether as
 * functions of the coordinates.  E[3] = unweighted (cart, rama, omega); grad[L][5][3] is
 * ACCUMULATED with the weighted derivatives (may be NULL). */
void trxo_cart_terms(int L, const int *aa, const double *xyz, double w_cart, double w_rama, double w_omega, double *E,
                     double *grad)
{
    double ec = 0.0, er = 0.0, eo = 0.0;
    for (int i = 0; i < L; ++i) {
        const double *N = XYZ(i, TRX_AT_N), *CA = XYZ(i, TRX_AT_CA), *CB = XYZ(i, TRX_AT_CB), *Cc = XYZ(i, TRX_AT_C), *O = XYZ(i, TRX_AT_O);
        double *gN = grad ? GRD(i, TRX_AT_N) : NULL, *gCA = grad ? GRD(i, TRX_AT_CA) : NULL, *gCB = grad ? GRD(i, TRX_AT_CB) : NULL;
        double *gC = grad ? GRD(i, TRX_AT_C) : NULL, *gO = grad ? GRD(i, TRX_AT_O) : NULL;
        ec += spring_bond(CA, N, TRX_B_N_CA, TRX_CART_KB, w_cart, gCA, gN);
        ec += spring_bond(Cc, CA, TRX_B_CA_C, TRX_CART_KB, w_cart, gC, gCA);
        ec += spring_bond(O, Cc, TRX_B_C_O, TRX_CART_KB, w_cart, gO, gC);
        ec += spring_angle(N, CA, Cc, TRX_A_N_CA_C, TRX_CART_KA, w_cart, gN, gCA, gC);
        ec += spring_angle(CA, Cc, O, TRX_A_CA_C_O, TRX_CART_KA, w_cart, gCA, gC, gO);
        {   /* CB tether to the virtual-CB position */
            double b[3], c[3], a[3], r[3];
            v_sub(CA, N, b); v_sub(Cc, CA, c); v_cross(b, c, a);
            for (int k = 0; k < 3; ++k) r[k] = CB[k] - (TRX_CB_A * a[k] + TRX_CB_B * b[k] + TRX_CB_C * c[k] + CA[k]);
            ec += TRX_CART_KCB * v_dot(r, r);
            if (grad) {
                double q[3], cq[3], qb[3];
                for (int k = 0; k < 3; ++k) q[k] = -w_cart * 2.0 * TRX_CART_KCB * r[k];   /* dE/d vCB */
                v_cross(c, q, cq); v_cross(q, b, qb);
                for (int k = 0; k < 3; ++k) {
                    double gb = TRX_CB_A * cq[k] + TRX_CB_B * q[k], gc = TRX_CB_A * qb[k] + TRX_CB_C * q[k];
                    gCB[k] -= q[k];
                    gN[k] -= gb;
                    gCA[k] += gb - gc + q[k];
                    gC[k] += gc;
                }
            }
        }
        if (i < L - 1) {
            const double *N1 = XYZ(i + 1, TRX_AT_N), *CA1 = XYZ(i + 1, TRX_AT_CA);
            double *gN1 = grad ? GRD(i + 1, TRX_AT_N) : NULL, *gCA1 = grad ? GRD(i + 1, TRX_AT_CA) : NULL;
            ec += spring_bond(N1, Cc, TRX_B_C_N, TRX_CART_KB, w_cart, gN1, gC);
            ec += spring_angle(CA, Cc, N1, TRX_A_CA_C_N, TRX_CART_KA, w_cart, gCA, gC, gN1);
            ec += spring_angle(O, Cc, N1, TRX_A_O_C_N, TRX_CART_KA, w_cart, gO, gC, gN1);
            ec += spring_angle(Cc, N1, CA1, TRX_A_C_N_CA, TRX_CART_KA, w_cart, gC, gN1, gCA1);
            {   /* carbonyl O in the peptide plane */
                double u[3], v[3], o[3], uv[3], vo[3], ou[3];
                v_sub(CA, Cc, u); v_sub(N1, Cc, v); v_sub(O, Cc, o);
                v_cross(u, v, uv); v_cross(v, o, vo); v_cross(o, u, ou);
                double t = v_dot(uv, o);
                ec += TRX_CART_KPL * t * t;
                if (grad) {
                    double f = w_cart * 2.0 * TRX_CART_KPL * t;
                    for (int k = 0; k < 3; ++k) {
                        gCA[k] += f * vo[k]; gN1[k] += f * ou[k]; gO[k] += f * uv[k];
                        gC[k] -= f * (vo[k] + ou[k] + uv[k]);
                    }
                }
            }
            {   /* omega tether */
                double om = trxo_dihedral(CA, Cc, N1, CA1), dev = om - TRX_PI;
                dev -= 2.0 * TRX_PI * floor((dev + TRX_PI) / (2.0 * TRX_PI));
                double deg = dev / TRX_DEG;
                eo += TRX_OMEGA_K * deg * deg;
                if (grad) add_dihedral_grad(CA, Cc, N1, CA1, w_omega * 2.0 * TRX_OMEGA_K * deg / TRX_DEG, gCA, gC, gN1, gCA1);
            }
        }
        if (i > 0 && i < L - 1) {   /* Ramachandran */
            const double *Cp = XYZ(i - 1, TRX_AT_C), *N1 = XYZ(i + 1, TRX_AT_N);
            double phi = trxo_dihedral(Cp, N, CA, Cc), psi = trxo_dihedral(N, CA, Cc, N1);
            int cls = aa[i] == TRX_AA_PRO ? 1 : 0;
            double P = TRX_RAMA_FLOOR, dP_dphi = 0.0, dP_dpsi = 0.0;
            for (int k = 0; k < TRX_RAMA_NB; ++k) {
                const double *b = TRX_RAMA[cls][k];
                double dphi = phi - b[0] * TRX_DEG, dpsi = psi - b[1] * TRX_DEG;
                double e = b[4] * exp(b[2] * (cos(dphi) - 1.0) + b[3] * (cos(dpsi) - 1.0));
                P += e;
                dP_dphi += -e * b[2] * sin(dphi);
                dP_dpsi += -e * b[3] * sin(dpsi);
            }
            er += -log(P) - TRX_RAMA_OFFSET[cls];
            if (grad) {
                add_dihedral_grad(Cp, N, CA, Cc, -w_rama * dP_dphi / P, GRD(i - 1, TRX_AT_C), gN, gCA, gC);
                add_dihedral_grad(N, CA, Cc, N1, -w_rama * dP_dpsi / P, gN, gCA, gC, GRD(i + 1, TRX_AT_N));
            }
        }
    }
    E[0] = ec; E[1] = er; E[2] = eo;
}

/* Cartesian-mode evaluation: xyz[L][5][3] are the degrees of freedom.  terms[TRX_NTERM]
 * unweighted; g[L][5][3] = d total / d xyz; returns the weighted total. */
double trxo_eval_cart(const trxo_target *T, const double *w, const double *xyz, double *terms, double *g)
{
    const int L = T->L;
    double *x9 = (double *)calloc((size_t)L * 9, sizeof(double)), *g9 = (double *)malloc(sizeof(double) * (size_t)L * 9);
    memset(g, 0, sizeof(double) * (size_t)L * TRX_NAT * 3);
    for (int i = 0; i < L; ++i) memcpy(x9 + (size_t)i * 9, XYZ(i, 0), 9 * sizeof(double));
    trxo_energy_grad(L, x9, T->sets[0], T->sets[1], T->sets[2], T->sets[3], w, terms, g9, NULL, NULL);
    for (int i = 0; i < L; ++i) memcpy(g + (size_t)i * TRX_NAT * 3, g9 + (size_t)i * 9, 9 * sizeof(double));
    terms[TRX_T_VDW] = trxo_vdw(L, T->aa, xyz, w[TRX_T_VDW], g);
    terms[TRX_T_HB] = trxo_hbond(L, T->aa, xyz, w[TRX_T_HB], g);
    double E[3];
    trxo_cart_terms(L, T->aa, xyz, w[TRX_T_CART], w[TRX_T_RAMA], w[TRX_T_OMEGA], E, g);
    terms[TRX_T_CART] = E[0]; terms[TRX_T_RAMA] = E[1]; terms[TRX_T_OMEGA] = E[2];
    double tot = 0.0;
    for (int k = 0; k < TRX_NTERM; ++k) tot += w[k] * terms[k];
    free(x9); free(g9);
    return tot;
}

/* ------------------------------------------------------------------ L-BFGS */
typedef struct {
    double w[TRX_NTERM];
    int max_iter;
    double tol;
    int clash_check;     /* 1: skip to run `skip_to` when rama+vdw (weights 1,1) < clash_thr at run start */
    double clash_thr;
    int skip_to;
    int cartesian;       /* 1: MinMover.cartesian(True): the coordinates are the degrees of freedom */
} trx_run;

typedef struct { long long evals, iters; } trxo_stats;

#define LS_SIGMA 0.1
#define LS_MAXBACK 20
#define NM_MEMORY 3

/* the function a run minimises: torsion space (x = tors, xyz rebuilt by NeRF) or Cartesian */
typedef struct {
    const trxo_target *T;
    const double *w;
    int cartesian;
    double *xyz;   /* torsion mode: coordinates of the last evaluation */
} objective;

static double obj_eval(const objective *o, const double *x, double *terms, double *g)
{
    return o->cartesian ? trxo_eval_cart(o->T, o->w, x, terms, g) : trxo_eval(o->T, o->w, x, terms, g, o->xyz);
}

/* One MinMover.apply: L-BFGS (history m) with non-monotone Armijo back-tracking over the n
 * degrees of freedom x (in/out).  Returns the final weighted energy. */
static double lbfgs_core(const objective *o, int n, const trx_run *run, int m, double *x, double *terms, trxo_stats *st)
{
    double *g = malloc(sizeof(double) * n), *gn = malloc(sizeof(double) * n), *d = malloc(sizeof(double) * n);
    double *xn = malloc(sizeof(double) * n), *S = malloc(sizeof(double) * (size_t)n * m), *Y = malloc(sizeof(double) * (size_t)n * m);
    double *rho = malloc(sizeof(double) * m), *al = malloc(sizeof(double) * m);
    double fmem[NM_MEMORY];
    int hist = 0, head = 0, nmem = 0;
    double f = obj_eval(o, x, terms, g);
    st->evals++;
    fmem[nmem++ % NM_MEMORY] = f;
    int restart = 1;
    for (int it = 0; it < run->max_iter; ++it) {
        /* direction */
        double gnorm = 0.0;
        for (int k = 0; k < n; ++k) gnorm += g[k] * g[k];
        gnorm = sqrt(gnorm);
        if (gnorm == 0.0) break;
        for (int k = 0; k < n; ++k) d[k] = -g[k];
        if (hist > 0) {
            for (int q = 0; q < hist; ++q) {
                int h = (head - 1 - q + 2 * m) % m;
                double a = 0.0;
                for (int k = 0; k < n; ++k) a += S[(size_t)h * n + k] * d[k];
                al[h] = rho[h] * a;
                for (int k = 0; k < n; ++k) d[k] -= al[h] * Y[(size_t)h * n + k];
            }
            int h0 = (head - 1 + m) % m;
            double yy = 0.0;
            for (int k = 0; k < n; ++k) yy += Y[(size_t)h0 * n + k] * Y[(size_t)h0 * n + k];
            double gamma = 1.0 / (rho[h0] * yy);
            for (int k = 0; k < n; ++k) d[k] *= gamma;
            for (int q = hist - 1; q >= 0; --q) {
                int h = (head - 1 - q + 2 * m) % m;
                double b = 0.0;
                for (int k = 0; k < n; ++k) b += Y[(size_t)h * n + k] * d[k];
                b *= rho[h];
                for (int k = 0; k < n; ++k) d[k] += (al[h] - b) * S[(size_t)h * n + k];
            }
        }
        double slope = 0.0;
        for (int k = 0; k < n; ++k) slope += g[k] * d[k];
        if (slope >= 0.0) {                     /* not a descent direction: steepest descent */
            hist = 0; restart = 1;
            for (int k = 0; k < n; ++k) d[k] = -g[k];
            slope = -gnorm * gnorm;
        }
        double alpha = restart ? fmin(1.0, 1.0 / gnorm) : 1.0;
        double fref = fmem[0];
        for (int q = 1; q < (nmem < NM_MEMORY ? nmem : NM_MEMORY); ++q) fref = fmax(fref, fmem[q]);
        double fn = 0.0;
        int ok = 0;
        for (int bt = 0; bt < LS_MAXBACK; ++bt) {
            for (int k = 0; k < n; ++k) xn[k] = x[k] + alpha * d[k];
            fn = obj_eval(o, xn, terms, gn);
            st->evals++;
            if (isfinite(fn) && fn <= fref + LS_SIGMA * alpha * slope) { ok = 1; break; }
            double q = isfinite(fn) ? -0.5 * slope * alpha * alpha / (fn - f - slope * alpha) : 0.1 * alpha;
            if (!(q > 0.1 * alpha)) q = 0.1 * alpha;
            if (q > 0.5 * alpha) q = 0.5 * alpha;
            alpha = q;
        }
        if (!ok) {
            if (hist > 0) { hist = 0; restart = 1; continue; }  /* retry once from steepest descent */
            break;
        }
        st->iters++;
        double sy = 0.0, ss = 0.0, yy = 0.0;
        double *s = S + (size_t)head * n, *y = Y + (size_t)head * n;
        for (int k = 0; k < n; ++k) {
            s[k] = xn[k] - x[k]; y[k] = gn[k] - g[k];
            sy += s[k] * y[k]; ss += s[k] * s[k]; yy += y[k] * y[k];
        }
        if (sy > 1e-10 * sqrt(ss * yy)) { rho[head] = 1.0 / sy; head = (head + 1) % m; if (hist < m) hist++; }
        int converged = 2.0 * fabs(fn - f) <= run->tol * (fabs(fn) + fabs(f) + 1e-10);
        memcpy(x, xn, sizeof(double) * n);
        memcpy(g, gn, sizeof(double) * n);
        f = fn;
        fmem[nmem++ % NM_MEMORY] = f;
        restart = 0;
        if (converged) break;
    }
    f = obj_eval(o, x, terms, g);   /* leave terms (and xyz) consistent with x */
    st->evals++;
    free(g); free(gn); free(d); free(xn); free(S); free(Y); free(rho); free(al);
    return f;
}

double trxo_lbfgs(const trxo_target *T, const trx_run *run, int m, double *tors, double *terms, double *xyz, trxo_stats *st)
{
    objective o = {T, run->w, 0, xyz};
    return lbfgs_core(&o, T->L * 3, run, m, tors, terms, st);
}

/* The staged schedule: runs[] in order; a run with clash_check evaluates rama+vdw first.
 * A Cartesian run minimises the coordinates the decoy currently has; afterwards the decoy
 * HOLDS those coordinates (xyz, terms) and its torsions are read back from them.  A clash
 * check on a holding decoy scores the held coordinates; the next torsion-space run that
 * actually starts rebuilds the chain from the torsions with ideal bond geometry. */
double trxo_fold(const trxo_target *T, const trx_run *runs, int nruns, int m, double *tors, double *terms, double *xyz,
                 trxo_stats *st)
{
    double f = 0.0;
    const int L = T->L;
    double *gt = malloc(sizeof(double) * (size_t)L * TRX_NAT * 3);
    int held = 0;
    {   /* coordinates of the start point (a schedule may open with a Cartesian run) */
        double w0[TRX_NTERM] = {0};
        trxo_eval(T, w0, tors, terms, gt, xyz);
    }
    for (int r = 0; r < nruns;) {
        if (runs[r].clash_check) {
            double e;
            if (held) e = terms[TRX_T_VDW] + terms[TRX_T_RAMA];
            else {
                double wv[TRX_NTERM] = {0, 0, 0, 1.0, 1.0, 0, 0, 0};
                e = trxo_eval(T, wv, tors, terms, gt, xyz);
                st->evals++;
            }
            if (e < runs[r].clash_thr) { r = runs[r].skip_to; continue; }
        }
        if (runs[r].cartesian) {
            objective o = {T, runs[r].w, 1, NULL};
            f = lbfgs_core(&o, L * TRX_NAT * 3, &runs[r], m, xyz, terms, st);
            trxo_torsions_from_xyz(L, xyz, tors);
            held = 1;
        } else {
            f = trxo_lbfgs(T, &runs[r], m, tors, terms, xyz, st);
            held = 0;
        }
        ++r;
    }
    free(gt);
    return f;
}

/* ------------------------------------------------------------------ flat wrappers for ctypes */
static void mk_sets(trxo_set *s, const trxo_set **p, const int *n, const int *const *a, const int *const *b, const int *K,
                    const double *const *x, const double *const *y, const double *const *y2)
{
    for (int t = 0; t < 4; ++t) {
        s[t].n = n[t]; s[t].K = K[t]; s[t].a = a[t]; s[t].b = b[t]; s[t].x = x[t]; s[t].y = y[t]; s[t].y2 = y2[t];
        p[t] = n[t] ? &s[t] : NULL;
    }
}

double trxo_eval_flat(int L, const int *aa, const int *n, const int *const *a, const int *const *b, const int *K,
                      const double *const *x, const double *const *y, const double *const *y2, const double *w,
                      const double *tors, double *terms, double *gt, double *xyz)
{
    trxo_set s[4]; trxo_target T; T.L = L; T.aa = aa;
    mk_sets(s, T.sets, n, a, b, K, x, y, y2);
    return trxo_eval(&T, w, tors, terms, gt, xyz);
}

double trxo_eval_cart_flat(int L, const int *aa, const int *n, const int *const *a, const int *const *b, const int *K,
                           const double *const *x, const double *const *y, const double *const *y2, const double *w,
                           const double *xyz, double *terms, double *g)
{
    trxo_set s[4]; trxo_target T; T.L = L; T.aa = aa;
    mk_sets(s, T.sets, n, a, b, K, x, y, y2);
    return trxo_eval_cart(&T, w, xyz, terms, g);
}

/* Folds N decoys (tors[N][L][3] in/out) on nthreads host threads. */
#include <pthread.h>
typedef struct {
    const trxo_target *T; const trx_run *runs; int nruns, m, n0, n1, L;
    double *tors, *terms, *xyz, *ftot; long long *evals, *iters;
} fold_job;

static void *fold_worker(void *arg)
{
    fold_job *j = (fold_job *)arg;
    for (int n = j->n0; n < j->n1; ++n) {
        trxo_stats st = {0, 0};
        j->ftot[n] = trxo_fold(j->T, j->runs, j->nruns, j->m, j->tors + (size_t)n * j->L * 3, j->terms + (size_t)n * TRX_NTERM,
                               j->xyz + (size_t)n * j->L * TRX_NAT * 3, &st);
        j->evals[n] = st.evals; j->iters[n] = st.iters;
    }
    return NULL;
}

void trxo_fold_batch(int nthreads, int N, int L, const int *aa, const int *n, const int *const *a, const int *const *b,
                     const int *K, const double *const *x, const double *const *y, const double *const *y2,
                     const trx_run *runs, int nruns, int m, double *tors, double *terms, double *xyz, double *ftot,
                     long long *evals, long long *iters)
{
    trxo_set s[4]; trxo_target T; T.L = L; T.aa = aa;
    mk_sets(s, T.sets, n, a, b, K, x, y, y2);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > N) nthreads = N;
    pthread_t th[256]; fold_job jobs[256];
    for (int k = 0; k < nthreads; ++k) {
        fold_job jb = {&T, runs, nruns, m, (int)((long long)N * k / nthreads), (int)((long long)N * (k + 1) / nthreads), L,
                       tors, terms, xyz, ftot, evals, iters};
        jobs[k] = jb;
        pthread_create(&th[k], NULL, fold_worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
}

int trxo_fold_abi_version(void) { return 3; }
