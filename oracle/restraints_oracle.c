/* CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain C, fp64 restatement of the arithmetic PyRosetta performs for the three
 * constraint score terms on the folding hot path:
 *   atom_pair_constraint (CB-CB SPLINE), dihedral_constraint (omega, theta SPLINE),
 *   angle_constraint (phi SPLINE).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * PARITY UNPINNED against PyRosetta: PyRosetta 2024.39 (environment.yml:16) is an
 * un-vendored binary dependency, absent from /root/reference and from this image.
 * What IS pinned: the knot tables (bytes-equal to the reference's gen_rst, see
 * tests/test_oracle_tables.py), the dihedral / angle conventions (equal to the
 * reference's own numpy get_dihedrals / get_angles, utils_trX2dy/utils.py:97-122,
 * golden vectors in tests/golden/geometry_random.npz), the spline algebra
 * (== scipy CubicSpline clamped, tests/test_oracle_restraints.py) and the analytic
 * gradients (== central differences).
 *
 * Restated published algorithm [ROSETTA-RECALL, SURVEY.md section 8a rows 9-10]:
 *  - core/scoring/func/SplineFunc: cubic spline through the listed knots with ZERO
 *    first derivative at both ends (numeric::interpolation::spline::SplineGenerator
 *    -> SimpleInterpolator), value weight*S(x) inside [lbx,ubx], constant
 *    weight*lby / weight*uby outside, derivative 0 outside.
 *    End-knot rule H1 (default): lbx = x_1 - bin_size, ubx = x_n + bin_size,
 *    lby = y_1, uby = y_n and all n listed points are interior knots (n+2 knots).
 *    Rule H2: lbx = x_1, ubx = x_n (n knots).  The rule is applied by the CALLER
 *    (oracle/restraints_oracle.py: apply_end_rule); this file sees final knots.
 *  - second derivatives: Numerical-Recipes `spline` (clamped form), evaluation:
 *    NR `splint` with bisection (numeric/interpolation/spline/spline_functions.cc).
 *  - AtomPairConstraint: f(|x_CBa - x_CBb|);  DihedralConstraint: f(dihedral in
 *    radians, IUPAC sign, (-pi,pi]) with omega = (CA_a,CB_a,CB_b,CA_b), theta =
 *    (N_a,CA_a,CB_a,CB_b);  AngleConstraint: f(angle(CA_a,CB_a,CB_b)) in [0,pi].
 *    Raw value goes straight into the spline (no periodic wrap).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ATOM_N 0
#define ATOM_CA 1
#define ATOM_CB 2

/* NR `spline`, clamped both ends: yp1, ypn first derivatives. 0-based arrays. */
void trxo_spline_fit(int n, const double *x, const double *y, double yp1, double ypn, double *y2)
{
    double *u = (double *)malloc(sizeof(double) * (size_t)n);
    y2[0] = -0.5;
    u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
    for (int i = 1; i < n - 1; ++i) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        u[i] = (y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1]);
        u[i] = (6.0 * u[i] / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    double qn = 0.5;
    double un = (3.0 / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
    y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
    for (int k = n - 2; k >= 0; --k) y2[k] = y2[k] * y2[k + 1] + u[k];
    free(u);
}

/* NR `splint` with bisection; also the first derivative. */
void trxo_spline_eval(int n, const double *xa, const double *ya, const double *y2a, double x,
                      double *y, double *dy)
{
    int klo = 0, khi = n - 1;
    while (khi - klo > 1) {
        int k = (khi + klo) >> 1;
        if (xa[k] > x) khi = k; else klo = k;
    }
    double h = xa[khi] - xa[klo];
    double a = (xa[khi] - x) / h, b = (x - xa[klo]) / h;
    *y = a * ya[klo] + b * ya[khi] + ((a * a * a - a) * y2a[klo] + (b * b * b - b) * y2a[khi]) * (h * h) / 6.0;
    *dy = (ya[khi] - ya[klo]) / h - ((3.0 * a * a - 1.0) * y2a[klo] - (3.0 * b * b - 1.0) * y2a[khi]) * h / 6.0;
}

/* SplineFunc::func / dfunc with weight 1: flat outside [x_0, x_{n-1}]. */
void trxo_splinefunc(int n, const double *xa, const double *ya, const double *y2a, double x,
                     double *f, double *df)
{
    if (x < xa[0]) { *f = ya[0]; *df = 0.0; return; }
    if (x > xa[n - 1]) { *f = ya[n - 1]; *df = 0.0; return; }
    trxo_spline_eval(n, xa, ya, y2a, x, f, df);
}

static void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* Dihedral p1-p2-p3-p4 exactly as the reference's numpy get_dihedrals
 * (utils_trX2dy/utils.py:97-110): project on the plane normal to b1, atan2. */
double trxo_dihedral(const double *p1, const double *p2, const double *p3, const double *p4)
{
    double b0[3], b1[3], b2[3], v[3], w[3], c[3];
    sub3(p1, p2, b0);            /* -(b - a) */
    sub3(p3, p2, b1);
    sub3(p4, p3, b2);
    double n1 = sqrt(dot3(b1, b1));
    b1[0] /= n1; b1[1] /= n1; b1[2] /= n1;
    double d0 = dot3(b0, b1), d2 = dot3(b2, b1);
    for (int k = 0; k < 3; ++k) { v[k] = b0[k] - d0 * b1[k]; w[k] = b2[k] - d2 * b1[k]; }
    cross3(b1, v, c);
    return atan2(dot3(c, w), dot3(v, w));
}

/* Gradient of the dihedral w.r.t. the four points (Blondel & Karplus 1996). */
void trxo_dihedral_grad(const double *p1, const double *p2, const double *p3, const double *p4,
                        double *g1, double *g2, double *g3, double *g4)
{
    double F[3], G[3], H[3], A[3], B[3];
    sub3(p1, p2, F); sub3(p2, p3, G); sub3(p4, p3, H);
    cross3(F, G, A); cross3(H, G, B);
    double A2 = dot3(A, A), B2 = dot3(B, B), Gn = sqrt(dot3(G, G));
    double FG = dot3(F, G), HG = dot3(H, G);
    for (int k = 0; k < 3; ++k) {
        g1[k] = -Gn / A2 * A[k];
        g4[k] = Gn / B2 * B[k];
        double t = FG / (A2 * Gn) * A[k] - HG / (B2 * Gn) * B[k];
        g2[k] = Gn / A2 * A[k] + t;
        g3[k] = -Gn / B2 * B[k] - t;
    }
}

/* Angle p1-p2-p3 at vertex p2, as the reference's get_angles (utils.py:113-122). */
double trxo_angle(const double *p1, const double *p2, const double *p3)
{
    double v[3], w[3];
    sub3(p1, p2, v); sub3(p3, p2, w);
    double nv = sqrt(dot3(v, v)), nw = sqrt(dot3(w, w));
    for (int k = 0; k < 3; ++k) { v[k] /= nv; w[k] /= nw; }
    return acos(dot3(v, w));
}

void trxo_angle_grad(const double *p1, const double *p2, const double *p3,
                     double *g1, double *g2, double *g3)
{
    double v[3], w[3];
    sub3(p1, p2, v); sub3(p3, p2, w);
    double nv = sqrt(dot3(v, v)), nw = sqrt(dot3(w, w));
    for (int k = 0; k < 3; ++k) { v[k] /= nv; w[k] /= nw; }
    double c = dot3(v, w);
    double s = sqrt(1.0 - c * c);
    for (int k = 0; k < 3; ++k) {
        g1[k] = -(w[k] - c * v[k]) / (nv * s);
        g3[k] = -(v[k] - c * w[k]) / (nw * s);
        g2[k] = -(g1[k] + g3[k]);
    }
}

/* One restraint type: n restraints, residue indices a[], b[], K shared knots x[K],
 * per-restraint y[n][K] and fitted y2[n][K]. */
typedef struct {
    int n, K;
    const int *a, *b;
    const double *x, *y, *y2;
} trxo_set;

/* Energies (unweighted) of the three terms and gradient of
 * w[0]*E_apc + w[1]*E_dih + w[2]*E_ang w.r.t. xyz[L][3 atoms: N,CA,CB][3].
 * val[t] (may be NULL) receives the raw geometric value of every restraint of
 * type t, ener[t] its energy, in table order. */
void trxo_energy_grad(int L, const double *xyz,
                      const trxo_set *dist, const trxo_set *omega, const trxo_set *theta, const trxo_set *phi,
                      const double *w, double *E, double *grad, double **val, double **ener)
{
    E[0] = E[1] = E[2] = 0.0;
    if (grad) memset(grad, 0, sizeof(double) * (size_t)L * 9);
#define AT(r, at) (xyz + ((size_t)(r) * 3 + (at)) * 3)
#define GR(r, at) (grad + ((size_t)(r) * 3 + (at)) * 3)
    if (dist) for (int r = 0; r < dist->n; ++r) {
        int a = dist->a[r], b = dist->b[r];
        double d[3]; sub3(AT(a, ATOM_CB), AT(b, ATOM_CB), d);
        double len = sqrt(dot3(d, d)), f, df;
        trxo_splinefunc(dist->K, dist->x, dist->y + (size_t)r * dist->K, dist->y2 + (size_t)r * dist->K, len, &f, &df);
        E[0] += f;
        if (val && val[0]) val[0][r] = len;
        if (ener && ener[0]) ener[0][r] = f;
        if (grad && df != 0.0 && len != 0.0)
            for (int k = 0; k < 3; ++k) {
                GR(a, ATOM_CB)[k] += w[0] * df * d[k] / len;
                GR(b, ATOM_CB)[k] -= w[0] * df * d[k] / len;
            }
    }
    for (int pass = 0; pass < 2; ++pass) {
        const trxo_set *s = pass ? theta : omega;
        if (!s) continue;
        for (int r = 0; r < s->n; ++r) {
            int a = s->a[r], b = s->b[r];
            const double *p1, *p2, *p3, *p4; double *q1, *q2, *q3, *q4;
            if (pass == 0) { p1 = AT(a, ATOM_CA); p2 = AT(a, ATOM_CB); p3 = AT(b, ATOM_CB); p4 = AT(b, ATOM_CA); }
            else           { p1 = AT(a, ATOM_N);  p2 = AT(a, ATOM_CA); p3 = AT(a, ATOM_CB); p4 = AT(b, ATOM_CB); }
            double ang = trxo_dihedral(p1, p2, p3, p4), f, df;
            trxo_splinefunc(s->K, s->x, s->y + (size_t)r * s->K, s->y2 + (size_t)r * s->K, ang, &f, &df);
            E[1] += f;
            if (val && val[1 + pass]) val[1 + pass][r] = ang;
            if (ener && ener[1 + pass]) ener[1 + pass][r] = f;
            if (grad && df != 0.0) {
                double g1[3], g2[3], g3[3], g4[3];
                trxo_dihedral_grad(p1, p2, p3, p4, g1, g2, g3, g4);
                if (pass == 0) { q1 = GR(a, ATOM_CA); q2 = GR(a, ATOM_CB); q3 = GR(b, ATOM_CB); q4 = GR(b, ATOM_CA); }
                else           { q1 = GR(a, ATOM_N);  q2 = GR(a, ATOM_CA); q3 = GR(a, ATOM_CB); q4 = GR(b, ATOM_CB); }
                for (int k = 0; k < 3; ++k) {
                    q1[k] += w[1] * df * g1[k]; q2[k] += w[1] * df * g2[k];
                    q3[k] += w[1] * df * g3[k]; q4[k] += w[1] * df * g4[k];
                }
            }
        }
    }
    if (phi) for (int r = 0; r < phi->n; ++r) {
        int a = phi->a[r], b = phi->b[r];
        const double *p1 = AT(a, ATOM_CA), *p2 = AT(a, ATOM_CB), *p3 = AT(b, ATOM_CB);
        double ang = trxo_angle(p1, p2, p3), f, df;
        trxo_splinefunc(phi->K, phi->x, phi->y + (size_t)r * phi->K, phi->y2 + (size_t)r * phi->K, ang, &f, &df);
        E[2] += f;
        if (val && val[3]) val[3][r] = ang;
        if (ener && ener[3]) ener[3][r] = f;
        if (grad && df != 0.0) {
            double g1[3], g2[3], g3[3];
            trxo_angle_grad(p1, p2, p3, g1, g2, g3);
            for (int k = 0; k < 3; ++k) {
                GR(a, ATOM_CA)[k] += w[2] * df * g1[k];
                GR(a, ATOM_CB)[k] += w[2] * df * g2[k];
                GR(b, ATOM_CB)[k] += w[2] * df * g3[k];
            }
        }
    }
#undef AT
#undef GR
}

/* Flat-argument wrapper for ctypes: sets with n == 0 are skipped. */
void trxo_energy_grad_flat(int L, const double *xyz,
                           int nd, const int *ad, const int *bd, int Kd, const double *xd, const double *yd, const double *y2d,
                           int no, const int *ao, const int *bo, int Ko, const double *xo, const double *yo, const double *y2o,
                           int nt, const int *at, const int *bt, int Kt, const double *xt, const double *yt, const double *y2t,
                           int np_, const int *ap, const int *bp, int Kp, const double *xp, const double *yp, const double *y2p,
                           const double *w, double *E, double *grad)
{
    trxo_set d = {nd, Kd, ad, bd, xd, yd, y2d}, o = {no, Ko, ao, bo, xo, yo, y2o};
    trxo_set t = {nt, Kt, at, bt, xt, yt, y2t}, p = {np_, Kp, ap, bp, xp, yp, y2p};
    trxo_energy_grad(L, xyz, nd ? &d : NULL, no ? &o : NULL, nt ? &t : NULL, np_ ? &p : NULL, w, E, grad, NULL, NULL);
}

/* ---- batch over decoys on several host threads (CPU baseline leg of bench.py) ---- */
#include <pthread.h>

typedef struct {
    int L, n0, n1;
    const double *xyz; const trxo_set *s[4]; const double *w; double *E; double *grad;
} trxo_job;

static void *trxo_worker(void *arg)
{
    trxo_job *j = (trxo_job *)arg;
    for (int n = j->n0; n < j->n1; ++n)
        trxo_energy_grad(j->L, j->xyz + (size_t)n * j->L * 9, j->s[0], j->s[1], j->s[2], j->s[3], j->w,
                         j->E + (size_t)n * 3, j->grad ? j->grad + (size_t)n * j->L * 9 : NULL, NULL, NULL);
    return NULL;
}

/* xyz[N][L][3][3], E[N][3], grad[N][L][3][3] (may be NULL); nthreads >= 1. */
void trxo_energy_grad_batch(int nthreads, int N, int L, const double *xyz,
                            int nd, const int *ad, const int *bd, int Kd, const double *xd, const double *yd, const double *y2d,
                            int no, const int *ao, const int *bo, int Ko, const double *xo, const double *yo, const double *y2o,
                            int nt, const int *at, const int *bt, int Kt, const double *xt, const double *yt, const double *y2t,
                            int np_, const int *ap, const int *bp, int Kp, const double *xp, const double *yp, const double *y2p,
                            const double *w, double *E, double *grad)
{
    trxo_set d = {nd, Kd, ad, bd, xd, yd, y2d}, o = {no, Ko, ao, bo, xo, yo, y2o};
    trxo_set t = {nt, Kt, at, bt, xt, yt, y2t}, p = {np_, Kp, ap, bp, xp, yp, y2p};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    trxo_job jobs[256];
    for (int k = 0; k < nthreads; ++k) {
        jobs[k].L = L; jobs[k].xyz = xyz; jobs[k].w = w; jobs[k].E = E; jobs[k].grad = grad;
        jobs[k].s[0] = nd ? &d : NULL; jobs[k].s[1] = no ? &o : NULL; jobs[k].s[2] = nt ? &t : NULL; jobs[k].s[3] = np_ ? &p : NULL;
        jobs[k].n0 = (int)((long long)N * k / nthreads);
        jobs[k].n1 = (int)((long long)N * (k + 1) / nthreads);
        pthread_create(&th[k], NULL, trxo_worker, &jobs[k]);
    }
    for (int k = 0; k < nthreads; ++k) pthread_join(th[k], NULL);
}
