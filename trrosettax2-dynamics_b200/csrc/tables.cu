// Restraint tables: clamped cubic spline fit on device (fp64), knot geometry, and the
// 16x16 residue-pair tiles the restraint kernel walks.
//
// What this replaces in the reference: add_rst writing minimize.cst and Rosetta's
// ConstraintSetMover parsing one text file per restraint and fitting one SplineFunc
// each (folding/utils_ros/utils_ros.py:706-743).
#include <algorithm>
#include <cmath>
#include <cstring>

#include <cstdlib>

#include "internal.cuh"

namespace trx {

// Numerical-Recipes `spline`, clamped with zero slope at both ends (what Rosetta's
// SplineGenerator(lbx,lby,0, ubx,uby,0) -> SimpleInterpolator computes).  One thread
// per restraint; the scratch vector u lives in registers/local memory (K <= MAXK).
__global__ void spline_fit_kernel(int n, int K, const double *__restrict__ x, const double *__restrict__ y,
                                  double *__restrict__ y2out, Coef<double> *__restrict__ tab64, Coef<float> *__restrict__ tab32)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const double *yr = y + (size_t)r * K;
    double u[MAXK], y2[MAXK];
    y2[0] = -0.5;
    u[0] = (3.0 / (x[1] - x[0])) * ((yr[1] - yr[0]) / (x[1] - x[0]) - 0.0);
    for (int i = 1; i < K - 1; ++i) {
        double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
        double p = sig * y2[i - 1] + 2.0;
        y2[i] = (sig - 1.0) / p;
        double t = (yr[i + 1] - yr[i]) / (x[i + 1] - x[i]) - (yr[i] - yr[i - 1]) / (x[i] - x[i - 1]);
        u[i] = (6.0 * t / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
    }
    double un = (3.0 / (x[K - 1] - x[K - 2])) * (0.0 - (yr[K - 1] - yr[K - 2]) / (x[K - 1] - x[K - 2]));
    y2[K - 1] = (un - 0.5 * u[K - 2]) / (0.5 * y2[K - 2] + 1.0);
    for (int k = K - 2; k >= 0; --k) y2[k] = y2[k] * y2[k + 1] + u[k];
    for (int k = 0; k < K; ++k) y2out[(size_t)r * K + k] = y2[k];
    // the same cubic per interval in powers of (x - x_k): 4 scalars = one 16 B (fp32) load.
    // Entry K-1 carries the flat value beyond the last knot.
    for (int k = 0; k < K; ++k) {
        Coef<double> c;
        if (k < K - 1) {
            const double h = x[k + 1] - x[k];
            c.c0 = yr[k];
            c.c1 = (yr[k + 1] - yr[k]) / h - h * (2.0 * y2[k] + y2[k + 1]) / 6.0;
            c.c2 = 0.5 * y2[k];
            c.c3 = (y2[k + 1] - y2[k]) / (6.0 * h);
        } else {
            c.c0 = yr[K - 1]; c.c1 = 0.0; c.c2 = 0.0; c.c3 = 0.0;
        }
        tab64[(size_t)r * K + k] = c;
        Coef<float> f;
        f.c0 = (float)c.c0; f.c1 = (float)c.c1; f.c2 = (float)c.c2; f.c3 = (float)c.c3;
        tab32[(size_t)r * K + k] = f;
    }
}

template <typename T>
static void fill_geom(KnotGeom<T> &g, int K, const double *x)
{
    memset(&g, 0, sizeof(g));
    g.K = K;
    if (K < 2) return;
    for (int k = 0; k < K; ++k) g.x[k] = (T)x[k];
    // interval guess anchored on the longest run of (nearly) equal spacings -- the
    // uniform part of the grid; the kernel corrects the guess against the true knots
    int best = 0, bestlen = 0;
    for (int k = 0; k + 1 < K;) {
        double hk = x[k + 1] - x[k];
        int m = k + 1;
        while (m + 1 < K && std::fabs((x[m + 1] - x[m]) - hk) < 1e-2 * hk) ++m;
        if (m - k > bestlen) { bestlen = m - k; best = k; }
        k = m;
    }
    // mean spacing of the run: the reference's %.3f rounding makes single intervals uneven at the 5e-4
    // level (0.3125 -> 0.312 / 0.313); taking one of them as THE spacing would let the guess drift
    double h = (x[best + bestlen] - x[best]) / bestlen;
    g.gx0 = (T)x[best];
    g.ginv = (T)(1.0 / h);
    g.goff = best;
    g.urun0 = best;
    g.urun1 = best + bestlen;
}

}  // namespace trx

using namespace trx;

int trx_tables::get_plan(int groups, Plan **out)
{
    // A work item is `chunk` consecutive tiles of one block row.  The chunk is a constant of the tables, NOT a
    // function of the number of decoy groups: a CTA adds the row gradients of its tiles in shared memory before
    // writing one row record, so the chunk fixes the order of those fp32 sums -- with a chunk that followed the
    // live-decoy count (as an earlier version did, to keep a few waves of CTAs in flight) a decoy's gradient
    // bits, and then its whole trajectory, depended on how many other decoys were still running.
    // One tile per CTA (145 CTAs per decoy group on the L=300 bench tables): measured best inside a fold, where most
    // launches carry a few hundred live decoys (chunk 1 / 2 / 3: 1379 / 1347 / 1294 decoys/s; standalone at 4096
    // decoys 1.27 / 1.26 / 1.23 ms).  TRX_K1_CHUNK: development knob.
    (void)groups;
    static const int fixed_chunk = [] { const char *ev = getenv("TRX_K1_CHUNK"); const int v = ev ? atoi(ev) : 0; return v > 0 ? v : 1; }();
    long long chunk = std::min<long long>(fixed_chunk, std::max(1, nb));
    auto it = plans.find((int)chunk);
    if (it != plans.end()) { *out = &it->second; return TRX_OK; }
    Plan p;
    p.groups = groups;
    std::vector<int> work;               // I, first tile, count, row record
    std::vector<std::vector<int>> blk(nb);
    int nrec = 0;
    std::vector<int> tile_rec(ntiles);
    for (int t = 0; t < ntiles; ++t) { tile_rec[t] = nrec++; blk[tileJ[t]].push_back(tile_rec[t]); }
    int t0 = 0;
    while (t0 < ntiles) {
        int I = tileI[t0], t1 = t0;
        while (t1 < ntiles && tileI[t1] == I) ++t1;
        for (int s = t0; s < t1; s += (int)chunk) {
            int cnt = std::min<int>((int)chunk, t1 - s);
            int rr = nrec++;
            blk[I].push_back(rr);
            work.insert(work.end(), {I, s, cnt, rr});
        }
        t0 = t1;
    }
    p.nwork = (int)work.size() / 4;
    p.nrec = nrec;
    // longest work items first (they were emitted row by row; sort by tile count desc)
    std::vector<int> order(p.nwork);
    for (int i = 0; i < p.nwork; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return work[a * 4 + 2] > work[b * 4 + 2]; });
    std::vector<int> sorted(work.size());
    for (int i = 0; i < p.nwork; ++i) memcpy(&sorted[i * 4], &work[order[i] * 4], 4 * sizeof(int));
    std::vector<int> ptr(nb + 1, 0), recs;
    for (int b = 0; b < nb; ++b) { ptr[b + 1] = ptr[b] + (int)blk[b].size(); recs.insert(recs.end(), blk[b].begin(), blk[b].end()); }
    if (recs.empty()) recs.push_back(0);
    if (sorted.empty()) sorted.assign(4, 0);
    // tile -> col record id is the identity by construction (tile_rec[t] == t)
    TRX_CUDA(ctx->dev_alloc(&p.d_work, sorted.size() * sizeof(int)));
    TRX_CUDA(ctx->dev_alloc(&p.d_blk_ptr, ptr.size() * sizeof(int)));
    TRX_CUDA(ctx->dev_alloc(&p.d_blk_rec, recs.size() * sizeof(int)));
    TRX_CUDA(cudaMemcpyAsync(p.d_work, sorted.data(), sorted.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(p.d_blk_ptr, ptr.data(), ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(p.d_blk_rec, recs.data(), recs.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    plans[(int)chunk] = p;
    *out = &plans[(int)chunk];
    return TRX_OK;
}

extern "C" {

int trx_tables_create(trx_ctx *ctx, int L, const trx_rst_set sets[4], trx_tables **out)
{
    TRX_REQUIRE(ctx && sets && out, "trx_tables_create: NULL argument");
    TRX_REQUIRE(L >= 2 && L <= 4096, "trx_tables_create: L=%d out of range [2,4096]", L);
    TRX_CUDA(cudaSetDevice(ctx->device));
    for (int t = 0; t < 4; ++t) {
        const trx_rst_set &s = sets[t];
        TRX_REQUIRE(s.n >= 0, "trx_tables_create: sets[%d].n < 0", t);
        if (s.n == 0) continue;
        TRX_REQUIRE(s.a && s.b && s.x && s.y, "trx_tables_create: sets[%d] has NULL arrays", t);
        const int kmax = t == TRX_DIST ? MAXK : MAXK_ANG;
        TRX_REQUIRE(s.K >= 3 && s.K <= kmax, "trx_tables_create: sets[%d].K=%d out of range [3,%d]", t, s.K, kmax);
        for (int k = 0; k + 1 < s.K; ++k)
            TRX_REQUIRE(s.x[k + 1] > s.x[k], "trx_tables_create: sets[%d].x not strictly increasing at %d", t, k);
        for (int r = 0; r < s.n; ++r)
            TRX_REQUIRE(s.a[r] >= 0 && s.a[r] < L && s.b[r] >= 0 && s.b[r] < L && s.a[r] != s.b[r],
                        "trx_tables_create: sets[%d] restraint %d has bad residues (%d,%d)", t, r, s.a[r], s.b[r]);
    }
    trx_tables *T = new trx_tables();
    struct Guard {   // a failing CUDA call below returns early: release what has been built
        trx_tables *T;
        ~Guard() { if (T) trx_tables_destroy(T); }
    } guard{T};
    T->ctx = ctx;
    ctx->retain();   // given back by trx_tables_destroy (the guard goes through it too)
    T->L = L;
    T->Lpad = padded_length(L);
    T->nb = T->Lpad / TILE;
    const int nb = T->nb;

    // ---- spline fit on device
    KnotGeom<double> g64[4];
    KnotGeom<float> g32[4];
    for (int t = 0; t < 4; ++t) {
        const trx_rst_set &s = sets[t];
        T->n[t] = s.n;
        T->K[t] = s.n ? s.K : 0;
        static const double dummy[2] = {0.0, 1.0};
        fill_geom(g64[t], s.n ? s.K : 2, s.n ? s.x : dummy);
        fill_geom(g32[t], s.n ? s.K : 2, s.n ? s.x : dummy);
        if (s.n && (g32[t].urun0 > 6 || g32[t].urun1 < s.K - 1)) {
            set_error("trx_tables_create: sets[%d].x must be uniformly spaced after at most 6 leading uneven intervals "
                      "(uniform run found: intervals [%d,%d) of %d)", t, g32[t].urun0, g32[t].urun1, s.K - 1);
            return TRX_ERR_INVALID;
        }
        if (s.n && t != TRX_DIST && g32[t].urun0 != 0) {   // the fp32 kernel looks angular intervals up without a head search
            set_error("trx_tables_create: sets[%d].x (angular grid) must be uniformly spaced from its first knot", t);
            return TRX_ERR_INVALID;
        }
        if (s.n == 0) continue;
        double *d_x = nullptr, *d_y = nullptr;
        size_t ny = (size_t)s.n * s.K;
        TRX_CUDA(ctx->dev_alloc(&d_x, s.K * sizeof(double)));
        TRX_CUDA(ctx->dev_alloc(&d_y, ny * sizeof(double)));
        TRX_CUDA(ctx->dev_alloc(&T->d_tab64[t], ny * sizeof(Coef<double>)));
        TRX_CUDA(ctx->dev_alloc(&T->d_tab32[t], ny * sizeof(Coef<float>)));
        TRX_CUDA(ctx->dev_alloc(&T->d_y2[t], ny * sizeof(double)));
        TRX_CUDA(cudaMemcpyAsync(d_x, s.x, s.K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        TRX_CUDA(cudaMemcpyAsync(d_y, s.y, ny * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        ctx->time_begin("spline_fit");
        spline_fit_kernel<<<(s.n + 127) / 128, 128, 0, ctx->stream>>>(s.n, s.K, d_x, d_y, T->d_y2[t], (Coef<double> *)T->d_tab64[t],
                                                                     (Coef<float> *)T->d_tab32[t]);
        ctx->time_end("spline_fit");
        TRX_CUDA(cudaGetLastError());
        ctx->dev_free(d_x);   // reuse is ordered on ctx->stream, no synchronisation needed
        ctx->dev_free(d_y);
    }
    TRX_CUDA(ctx->dev_alloc(&T->d_geom64, sizeof(g64)));
    TRX_CUDA(ctx->dev_alloc(&T->d_geom32, sizeof(g32)));
    TRX_CUDA(cudaMemcpyAsync(T->d_geom64, g64, sizeof(g64), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(T->d_geom32, g32, sizeof(g32), cudaMemcpyHostToDevice, ctx->stream));

    // ---- pair slots: for the unordered pair i<j the six restraints
    //   0 dist(i,j) 1 omega(i,j) 2 theta(i,j) 3 theta(j,i) 4 phi(i,j) 5 phi(j,i)
    // kept per active 16x16 tile as [r][c][8] ints: mask, 6 indices, pad
    std::vector<int> tile_of((size_t)nb * nb, -1);
    auto slot_of = [&](int type, int a, int b, int &i, int &j) -> int {
        i = std::min(a, b);
        j = std::max(a, b);
        if (type == TRX_DIST) return 0;
        if (type == TRX_OMEGA) return 1;
        if (type == TRX_THETA) return a < b ? 2 : 3;
        return a < b ? 4 : 5;
    };
    // pass 1: which tiles are active
    for (int t = 0; t < 4; ++t)
        for (int r = 0; r < sets[t].n; ++r) {
            int i, j;
            slot_of(t, sets[t].a[r], sets[t].b[r], i, j);
            tile_of[(size_t)(i / TILE) * nb + j / TILE] = 0;
        }
    for (int I = 0; I < nb; ++I)
        for (int J = I; J < nb; ++J)
            if (tile_of[(size_t)I * nb + J] == 0) {
                tile_of[(size_t)I * nb + J] = T->ntiles++;
                T->tileI.push_back(I);
                T->tileJ.push_back(J);
            }
    std::vector<int> rec((size_t)std::max(1, T->ntiles) * TILE * TILE * 8, -1);
    for (size_t p = 0; p < rec.size(); p += 8) { rec[p] = 0; rec[p + 7] = 0; }
    for (int t = 0; t < 4; ++t)
        for (int r = 0; r < sets[t].n; ++r) {
            int i, j;
            int slot = slot_of(t, sets[t].a[r], sets[t].b[r], i, j);
            int tile = tile_of[(size_t)(i / TILE) * nb + j / TILE];
            size_t p = (((size_t)tile * TILE + i % TILE) * TILE + j % TILE) * 8;
            if (rec[p + 1 + slot] != -1) {
                set_error("trx_tables_create: duplicate restraint of type %d on pair (%d,%d)", t, sets[t].a[r], sets[t].b[r]);
                return TRX_ERR_INVALID;
            }
            if (rec[p] == 0) T->active_pairs++;
            rec[p] |= 1 << slot;
            rec[p + 1 + slot] = r * sets[t].K;   // element offset of the restraint's first interval
        }
    // ---- per-tile step schedule.  In one step each of the K1_WARPS warps of the restraint kernel
    // evaluates one residue pair; row and column gradients are accumulated in shared memory
    // without atomics, so the pairs of a step must have distinct rows and distinct columns
    // (a matching of the tile's rows x columns graph).  Greedy list scheduling, busiest
    // rows/columns first: sparse tiles take about max(pairs/K1_WARPS, max degree) steps.
    std::vector<unsigned short> sched;
    std::vector<int> step_ptr(std::max(1, T->ntiles) + 1, 0);
    for (int t = 0; t < T->ntiles; ++t) {
        struct P { int r, c, key; };
        std::vector<P> pairs;
        int degr[TILE] = {0}, degc[TILE] = {0};
        for (int r = 0; r < TILE; ++r)
            for (int c = 0; c < TILE; ++c)
                if (rec[(((size_t)t * TILE + r) * TILE + c) * 8]) { degr[r]++; degc[c]++; pairs.push_back({r, c, 0}); }
        for (auto &q : pairs) q.key = std::max(degr[q.r], degc[q.c]);
        std::stable_sort(pairs.begin(), pairs.end(), [](const P &a, const P &b) { return a.key > b.key; });
        std::vector<unsigned> rowmask, colmask;   // per step: rows / columns in use
        std::vector<int> fill;
        std::vector<std::vector<unsigned short>> steps;
        for (auto &q : pairs) {
            size_t sidx = 0;
            for (; sidx < steps.size(); ++sidx)
                if (fill[sidx] < K1_WARPS && !(rowmask[sidx] >> q.r & 1u) && !(colmask[sidx] >> q.c & 1u)) break;
            if (sidx == steps.size()) { steps.emplace_back(); rowmask.push_back(0); colmask.push_back(0); fill.push_back(0); }
            steps[sidx].push_back((unsigned short)(q.r | (q.c << 4)));
            rowmask[sidx] |= 1u << q.r;
            colmask[sidx] |= 1u << q.c;
            fill[sidx]++;
        }
        for (auto &st : steps) {
            for (int w = 0; w < K1_WARPS; ++w) sched.push_back(w < (int)st.size() ? st[w] : (unsigned short)0xffff);
        }
        step_ptr[t + 1] = step_ptr[t] + (int)steps.size();
    }
    if (sched.empty()) sched.assign(K1_WARPS, 0xffff);
    T->total_steps = step_ptr[std::max(1, T->ntiles)];
    TRX_CUDA(ctx->dev_alloc(&T->d_sched, sched.size() * sizeof(unsigned short)));
    TRX_CUDA(ctx->dev_alloc(&T->d_nsteps, step_ptr.size() * sizeof(int)));
    TRX_CUDA(cudaMemcpyAsync(T->d_sched, sched.data(), sched.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(T->d_nsteps, step_ptr.data(), step_ptr.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int> tj = T->tileJ;
    if (tj.empty()) tj.push_back(0);
    TRX_CUDA(ctx->dev_alloc(&T->d_tileJ, tj.size() * sizeof(int)));
    TRX_CUDA(ctx->dev_alloc(&T->d_pairrec, rec.size() * sizeof(int)));
    TRX_CUDA(cudaMemcpyAsync(T->d_tileJ, tj.data(), tj.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(T->d_pairrec, rec.data(), rec.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));   // the caller's knot arrays may be pinned: done with them on return
    guard.T = nullptr;
    *out = T;
    return TRX_OK;
}

int trx_tables_destroy(trx_tables *T)
{
    if (!T) return TRX_OK;
    cudaSetDevice(T->ctx->device);
    cudaStreamSynchronize(T->ctx->stream);
    for (int t = 0; t < 4; ++t) {
        T->ctx->dev_free(T->d_tab64[t]);
        T->ctx->dev_free(T->d_tab32[t]);
        T->ctx->dev_free(T->d_y2[t]);
    }
    T->ctx->dev_free(T->d_geom64);
    T->ctx->dev_free(T->d_geom32);
    T->ctx->dev_free(T->d_tileJ);
    T->ctx->dev_free(T->d_pairrec);
    T->ctx->dev_free(T->d_sched);
    T->ctx->dev_free(T->d_nsteps);
    for (auto &kv : T->plans) {
        T->ctx->dev_free(kv.second.d_work);
        T->ctx->dev_free(kv.second.d_blk_ptr);
        T->ctx->dev_free(kv.second.d_blk_rec);
    }
    trx_ctx *ctx = T->ctx;
    delete T;
    trx::ctx_release(ctx);
    return TRX_OK;
}

int trx_tables_set_dist_atom(trx_tables *t, int atom)
{
    TRX_REQUIRE(t, "trx_tables_set_dist_atom: NULL tables");
    TRX_REQUIRE(atom == TRX_ATOM_CA || atom == TRX_ATOM_CB, "trx_tables_set_dist_atom: atom must be TRX_ATOM_CA or TRX_ATOM_CB");
    TRX_REQUIRE(atom == TRX_ATOM_CB || (t->n[1] == 0 && t->n[2] == 0 && t->n[3] == 0),
                "trx_tables_set_dist_atom: CA-CA distance restraints cannot be combined with angular restraints");
    t->dist_ca = atom == TRX_ATOM_CA;
    return TRX_OK;
}

int trx_tables_info(const trx_tables *T, int *L, int counts[4], int *tiles)
{
    TRX_REQUIRE(T, "trx_tables_info: tables is NULL");
    if (L) *L = T->L;
    if (counts) for (int t = 0; t < 4; ++t) counts[t] = T->n[t];
    if (tiles) *tiles = T->ntiles;
    return TRX_OK;
}

int trx_tables_get_y2(trx_tables *T, int type, double *y2)
{
    TRX_REQUIRE(T && y2 && type >= 0 && type < 4, "trx_tables_get_y2: bad argument");
    size_t n = (size_t)T->n[type] * T->K[type];
    if (n == 0) return TRX_OK;
    TRX_CUDA(cudaMemcpyAsync(y2, T->d_y2[type], n * sizeof(double), cudaMemcpyDeviceToHost, T->ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(T->ctx->stream));
    return TRX_OK;
}

}  // extern "C"
