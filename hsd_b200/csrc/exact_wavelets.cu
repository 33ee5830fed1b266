// Exact heat-kernel wavelets from an eigendecomposition, fused with the reference's threshold:
//   Psi = U diag(exp(-s lambda)) U^T ;  Psi[i][j] = Psi[i][j] > thr ? Psi[i][j] : 0
// Reference: model/HSD.py:61-66 (and model/GraphWave.py:42-49) — two dense np.dot products and a
// Python-level np.vectorize over N^2 entries.  Here one FP64 kernel: the scaling by exp(-s lambda_k)
// is applied while the A tile is staged, only tiles on or above the diagonal are computed (Psi is
// symmetric) and each is stored thresholded and mirrored, so the un-thresholded N x N product is
// never materialised.  The eigendecomposition itself stays on the vendor solver (cuSOLVER through
// torch.linalg.eigh): it is the step before the hot path (SURVEY.md §8 f rank 1).
//
// Shape: 64 x 64 output tile per CTA, 256 threads x (4 x 4) FP64 register tiles, K in chunks of 16
// staged k-major in shared memory (conflict-free broadcast reads).  FP64 CUDA cores; the summation
// runs k ascending inside each thread, so results are bit-reproducible run to run.
#include <math.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int XT = 64;    // tile edge
constexpr int XK = 16;    // K chunk

__global__ void __launch_bounds__(256)
exact_wavelets_kernel(const double* __restrict__ U, int64_t ldu, const double* __restrict__ lam, int n,
                      double scale, double thr, int apply_thr, double* __restrict__ out, int64_t ldo,
                      int tiles) {
    __shared__ double sa[XK][XT + 1];
    __shared__ double sb[XK][XT + 1];
    // linear index over tiles (I, J >= I)
    int t = blockIdx.x, I = 0;
    while (t >= tiles - I) { t -= tiles - I; ++I; }
    const int J = I + t;
    const int i0 = I * XT, j0 = J * XT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads, 4 x 4 outputs each
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4; // loader: row lr (0..63), k offset lk (0,4,8,12)
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    for (int k0 = 0; k0 < n; k0 += XK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lk + q;
            const int ia = i0 + lr, jb = j0 + lr;
            const double w = (k < n) ? exp(-scale * __ldg(lam + k)) : 0.0;
            sa[lk + q][lr] = (k < n && ia < n) ? __ldg(U + (int64_t)ia * ldu + k) * w : 0.0;
            sb[lk + q][lr] = (k < n && jb < n) ? __ldg(U + (int64_t)jb * ldu + k) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < XK; ++kk) {
            double a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = sa[kk][ty * 4 + r];
#pragma unroll
            for (int c = 0; c < 4; ++c) b[c] = sb[kk][tx * 4 + c];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        if (i >= n) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx * 4 + c;
            if (j >= n) continue;
            if (I == J && j < i) continue;      // diagonal tile: the lower half comes from the mirror below
            double v = acc[r][c];
            if (apply_thr) v = (v > thr) ? v : 0.0;   // model/HSD.py:65
            out[(int64_t)i * ldo + j] = v;
            if (i != j) out[(int64_t)j * ldo + i] = v;
        }
    }
}

}  // namespace hsd

extern "C" int hsd_exact_wavelets(const double* U, int64_t ldu, const double* lam, int32_t n_nodes,
                                  double scale, double threshold, int32_t apply_threshold, double* out,
                                  int64_t ld_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(U && lam && out, "null pointer");
    HSD_REQUIRE(n_nodes > 0 && ldu >= n_nodes && ld_out >= n_nodes, "bad sizes");
    const int tiles = (n_nodes + XT - 1) / XT;
    const long long n_tiles = (long long)tiles * (tiles + 1) / 2;
    HSD_REQUIRE(n_tiles < (1ll << 31), "matrix too large for one launch");
    exact_wavelets_kernel<<<(unsigned)n_tiles, 256, 0, (cudaStream_t)stream>>>(
        U, ldu, lam, n_nodes, scale, threshold, apply_threshold ? 1 : 0, out, ld_out, tiles);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
