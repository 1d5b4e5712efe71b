// Batched structure-comparison metrics of a decoy set, all pairs at once on the device.
//
// Replaces the reference's analysis step, which compares decoys pair by pair on the host:
//   get_tmscore_and_rmsd_matrix (utils_trX2dy/utils.py:527-540): one `./bin/TMscore a.pdb b.pdb`
//     subprocess per pair, TM-score and RMSD parsed from its text output;
//   get_glocon_matrix (utils.py:543-567): CB distance maps (get_neighbors, 20 A cut-off), mean
//     over residue pairs of |d1 - d2| where it exceeds 3 A.
// fp64 throughout: these are small reductions (M^2 pairs of L <= a few hundred residues) whose
// results feed thresholds (cluster assignment, convergence), so they follow the reference's
// double arithmetic rather than the fold's fp32.
#include <cmath>

#include "internal.cuh"

namespace trx {

// ---- GloCon ------------------------------------------------------------------------------
// dmap[m][p]: CB-CB distance of residue pair p = (a < b) of decoy m, 0 beyond dmax (the
// reference's dist6d is zero where the KD-tree found no neighbour, utils.py:160-163).
__global__ void dmap_kernel(int M, int L, const double *__restrict__ cb, double dmax, double *__restrict__ dmap)
{
    const int m = blockIdx.y;
    const long long P = (long long)L * (L - 1) / 2;
    const double *x = cb + (size_t)m * L * 3;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        // p -> (a, b), a < b, row-major over the upper triangle
        int a = (int)((2.0 * L - 1.0 - sqrt((2.0 * L - 1.0) * (2.0 * L - 1.0) - 8.0 * (double)p)) * 0.5);
        while ((long long)a * (2 * L - a - 1) / 2 > p) --a;
        while ((long long)(a + 1) * (2 * L - a - 2) / 2 <= p) ++a;
        const int b = (int)(p - (long long)a * (2 * L - a - 1) / 2) + a + 1;
        const double dx = x[b * 3] - x[a * 3], dy = x[b * 3 + 1] - x[a * 3 + 1], dz = x[b * 3 + 2] - x[a * 3 + 2];
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        dmap[(size_t)m * P + p] = d <= dmax ? d : 0.0;
    }
}

// out[i][j] = sum_p g(|dmap[i][p] - dmap[j][p]|) / P with g(x) = x if x > thr else 0.
// A CTA owns a 16x16 tile of decoy pairs and streams the residue pairs through shared memory.
constexpr int GT = 16, GP = 64;
__global__ void __launch_bounds__(GT * GT) glocon_kernel(int M, long long P, const double *__restrict__ dmap, double thr,
                                                         double *__restrict__ out)
{
    __shared__ double A[GT][GP + 1], B[GT][GP + 1];
    const int ti = threadIdx.x / GT, tj = threadIdx.x % GT;
    const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
    if (j0 > i0) return;   // symmetric: the lower triangle of tiles is enough
    double acc = 0.0;
    for (long long p0 = 0; p0 < P; p0 += GP) {
        for (int e = threadIdx.x; e < GT * GP; e += GT * GT) {
            const int r = e / GP, c = e % GP;
            const long long p = p0 + c;
            A[r][c] = (i0 + r < M && p < P) ? dmap[(size_t)(i0 + r) * P + p] : 0.0;
            B[r][c] = (j0 + r < M && p < P) ? dmap[(size_t)(j0 + r) * P + p] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < GP; ++c) {
            const double d = fabs(A[ti][c] - B[tj][c]);
            acc += d > thr ? d : 0.0;
        }
        __syncthreads();
    }
    const int i = i0 + ti, j = j0 + tj;
    if (i < M && j < M && i != j) {
        const double v = acc / (double)P;
        out[(size_t)i * M + j] = v;
        out[(size_t)j * M + i] = v;
    }
    if (i < M && i == j) out[(size_t)i * M + i] = 0.0;
}

// ---- superposition: RMSD and TM-score ------------------------------------------------------
// Optimal rotation by Horn's quaternion method: the rotation maximising sum q.(R p) is the
// eigenvector of the largest eigenvalue of a symmetric 4x4 matrix built from the covariance;
// cyclic Jacobi on 4x4 converges in a few sweeps.  Same optimum as Kabsch/SVD with the
// reflection fix.  Every lane runs it redundantly on identical inputs (no broadcast needed).
__device__ void horn_rotation(const double S[9], double R[9])
{
    double N[4][4] = {{S[0] + S[4] + S[8], S[5] - S[7], S[6] - S[2], S[1] - S[3]},
                      {S[5] - S[7], S[0] - S[4] - S[8], S[1] + S[3], S[6] + S[2]},
                      {S[6] - S[2], S[1] + S[3], -S[0] + S[4] - S[8], S[5] + S[7]},
                      {S[1] - S[3], S[6] + S[2], S[5] + S[7], -S[0] - S[4] + S[8]}};
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) off += N[p][q] * N[p][q];
        if (off < 1e-30) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (fabs(N[p][q]) < 1e-300) continue;
                const double theta = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = N[k][p], akq = N[k][q];
                    N[k][p] = c * akp - s * akq;
                    N[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = N[p][k], aqk = N[q][k];
                    N[p][k] = c * apk - s * aqk;
                    N[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; ++k) if (N[k][k] > N[best][best]) best = k;
    const double q0 = V[0][best], q1 = V[1][best], q2 = V[2][best], q3 = V[3][best];
    R[0] = q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3; R[1] = 2 * (q1 * q2 - q0 * q3); R[2] = 2 * (q1 * q3 + q0 * q2);
    R[3] = 2 * (q1 * q2 + q0 * q3); R[4] = q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3; R[5] = 2 * (q2 * q3 - q0 * q1);
    R[6] = 2 * (q1 * q3 - q0 * q2); R[7] = 2 * (q2 * q3 + q0 * q1); R[8] = q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3;
}

__device__ __forceinline__ double warp_sum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per ordered decoy pair (i = model, j = native).  The TM-score search restates the
// published heuristic (Zhang & Skolnick 2004) the way the host metric of this package does
// (metrics.tm_score): seeds = contiguous fragments of length L, L/2, ... >= 4 at half-fragment
// strides; from each seed iterate { superpose on the current subset; score all residues; keep
// the residues closer than a growing cut-off } until the subset stops changing (<= 20 times).
// rmsd[i][j]: Kabsch RMSD over all residues.  tm[i][j]: normalised by L (both have L residues).
constexpr int TM_WARPS = 4;
__global__ void __launch_bounds__(TM_WARPS * 32) tm_kernel(int M, int L, const double *__restrict__ ca, double *__restrict__ tm,
                                                           double *__restrict__ rmsd)
{
    extern __shared__ double tm_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long pair = (long long)blockIdx.x * TM_WARPS + warp;
    if (pair >= (long long)M * M) return;
    const int i = (int)(pair / M), j = (int)(pair % M);
    if (i == j) {
        if (lane == 0) { tm[pair] = 1.0; rmsd[pair] = 0.0; }
        return;
    }
    double *X = tm_smem + (size_t)warp * (7 * L);    // model xyz [L][3]
    double *Y = X + 3 * L;                            // native xyz [L][3]
    double *d2 = Y + 3 * L;                           // squared deviation per residue
    unsigned char *in = reinterpret_cast<unsigned char *>(tm_smem + (size_t)TM_WARPS * 7 * L) + (size_t)warp * 2 * L;   // subset flags, old and new
    unsigned char *nw = in + L;
    for (int k = lane; k < 3 * L; k += 32) { X[k] = ca[(size_t)i * L * 3 + k]; Y[k] = ca[(size_t)j * L * 3 + k]; }
    __syncwarp();
    const double d0 = L > 21 ? fmax(0.5, 1.24 * cbrt((double)(L - 15)) - 1.8) : 0.5;
    double R[9], T[3];
    // superpose X onto Y over the flagged subset; fills R, T; returns the subset size
    auto superpose = [&](const unsigned char *flag) -> int {
        double n = 0, sx[3] = {0, 0, 0}, sy[3] = {0, 0, 0};
        for (int k = lane; k < L; k += 32)
            if (flag[k]) { n += 1; for (int c = 0; c < 3; ++c) { sx[c] += X[k * 3 + c]; sy[c] += Y[k * 3 + c]; } }
        n = warp_sum(n);
        for (int c = 0; c < 3; ++c) { sx[c] = warp_sum(sx[c]) / n; sy[c] = warp_sum(sy[c]) / n; }
        double S[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = lane; k < L; k += 32)
            if (flag[k])
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) S[a * 3 + b] += (X[k * 3 + a] - sx[a]) * (Y[k * 3 + b] - sy[b]);
        for (int e = 0; e < 9; ++e) S[e] = warp_sum(S[e]);
        horn_rotation(S, R);
        for (int a = 0; a < 3; ++a) T[a] = sy[a] - (R[a * 3] * sx[0] + R[a * 3 + 1] * sx[1] + R[a * 3 + 2] * sx[2]);
        return (int)(n + 0.5);
    };
    // deviations of all residues under (R, T); returns sum 1/(1+d2/d0^2) and sum d2
    auto deviations = [&](double &sum_d2) -> double {
        double s = 0, q = 0;
        for (int k = lane; k < L; k += 32) {
            double e = 0;
            for (int a = 0; a < 3; ++a) {
                const double v = R[a * 3] * X[k * 3] + R[a * 3 + 1] * X[k * 3 + 1] + R[a * 3 + 2] * X[k * 3 + 2] + T[a] - Y[k * 3 + a];
                e += v * v;
            }
            d2[k] = e;
            s += 1.0 / (1.0 + e / (d0 * d0));
            q += e;
        }
        sum_d2 = warp_sum(q);
        return warp_sum(s);
    };
    double best = 0.0;
    for (int frag = L; frag >= 4; frag /= 2) {
        const int step = max(1, frag / 2);
        for (int start = 0; start + frag <= L; start += step) {
            for (int k = lane; k < L; k += 32) in[k] = (k >= start && k < start + frag) ? 1 : 0;
            __syncwarp();
            for (int it = 0; it < 20; ++it) {
                superpose(in);
                double q;
                const double s = deviations(q) / L;
                __syncwarp();
                if (frag == L && start == 0 && it == 0 && lane == 0) rmsd[pair] = sqrt(q / L);
                best = fmax(best, s);
                double cut = it == 0 ? d0 + 1.0 : fmin(d0 + 1.0 + 0.5 * it, 8.0);
                int cnt, same;
                for (;;) {
                    int c = 0, sm = 1;
                    for (int k = lane; k < L; k += 32) {
                        const unsigned char f = d2[k] < cut * cut ? 1 : 0;
                        nw[k] = f;
                        c += f;
                        sm &= (f == in[k]);
                    }
                    cnt = (int)(warp_sum((double)c) + 0.5);
                    same = __all_sync(0xffffffffu, sm);
                    if (cnt >= 3) break;
                    cut += 0.5;
                }
                __syncwarp();
                if (same) break;
                for (int k = lane; k < L; k += 32) in[k] = nw[k];
                __syncwarp();
            }
        }
    }
    if (lane == 0) tm[pair] = best;
}

}  // namespace trx

using namespace trx;

extern "C" {

int trx_glocon_matrix(trx_ctx *ctx, int M, int L, const double *cb, double dmax, double thr, double *out)
{
    TRX_REQUIRE(ctx && cb && out, "trx_glocon_matrix: NULL argument");
    TRX_REQUIRE(M >= 1 && L >= 2, "trx_glocon_matrix: need M >= 1 decoys of L >= 2 residues");
    TRX_CUDA(cudaSetDevice(ctx->device));
    const long long P = (long long)L * (L - 1) / 2;
    void *d_cb, *d_map, *d_out;
    int rc;
    if ((rc = ctx->get_scratch("metric_xyz", (size_t)M * L * 3 * sizeof(double), &d_cb))) return rc;
    if ((rc = ctx->get_scratch("metric_dmap", (size_t)M * P * sizeof(double), &d_map))) return rc;
    if ((rc = ctx->get_scratch("metric_out", (size_t)M * M * sizeof(double), &d_out))) return rc;
    TRX_CUDA(cudaMemcpyAsync(d_cb, cb, (size_t)M * L * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->time_begin("glocon");
    dmap_kernel<<<dim3((unsigned)std::min<long long>((P + 255) / 256, 1024), M), 256, 0, ctx->stream>>>(M, L, (const double *)d_cb, dmax, (double *)d_map);
    ctx->time_end("glocon");
    ctx->time_begin("glocon");
    const int nt = (M + GT - 1) / GT;
    glocon_kernel<<<dim3(nt, nt), GT * GT, 0, ctx->stream>>>(M, P, (const double *)d_map, thr, (double *)d_out);
    ctx->time_end("glocon");
    TRX_CUDA(cudaGetLastError());
    TRX_CUDA(cudaMemcpyAsync(out, d_out, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRX_OK;
}

int trx_tmscore_matrix(trx_ctx *ctx, int M, int L, const double *ca, double *tm, double *rmsd)
{
    TRX_REQUIRE(ctx && ca && tm && rmsd, "trx_tmscore_matrix: NULL argument");
    TRX_REQUIRE(M >= 1 && L >= 4, "trx_tmscore_matrix: need M >= 1 decoys of L >= 4 residues");
    const size_t smem = (size_t)TM_WARPS * 7 * L * sizeof(double) + (size_t)TM_WARPS * 2 * L;
    TRX_REQUIRE(smem <= 220 * 1024, "trx_tmscore_matrix: L=%d exceeds the shared-memory budget", L);
    TRX_CUDA(cudaSetDevice(ctx->device));
    void *d_ca, *d_tm, *d_rm;
    int rc;
    if ((rc = ctx->get_scratch("metric_xyz", (size_t)M * L * 3 * sizeof(double), &d_ca))) return rc;
    if ((rc = ctx->get_scratch("metric_out", (size_t)M * M * sizeof(double), &d_tm))) return rc;
    if ((rc = ctx->get_scratch("metric_out2", (size_t)M * M * sizeof(double), &d_rm))) return rc;
    TRX_CUDA(cudaMemcpyAsync(d_ca, ca, (size_t)M * L * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    TRX_CUDA(cudaFuncSetAttribute(tm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long pairs = (long long)M * M;
    ctx->time_begin("tmscore");
    tm_kernel<<<(unsigned)((pairs + TM_WARPS - 1) / TM_WARPS), TM_WARPS * 32, smem, ctx->stream>>>(M, L, (const double *)d_ca, (double *)d_tm, (double *)d_rm);
    ctx->time_end("tmscore");
    TRX_CUDA(cudaGetLastError());
    TRX_CUDA(cudaMemcpyAsync(tm, d_tm, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaMemcpyAsync(rmsd, d_rm, (size_t)M * M * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    TRX_CUDA(cudaStreamSynchronize(ctx->stream));
    return TRX_OK;
}

}  // extern "C"
