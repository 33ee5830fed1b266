// K6 — GraphWave characteristic-function embedding (the component next to the hot path,
// SURVEY.md §8 f rank 1).  Replaces model/GraphWave.py:53-69:
//   emb[i][2k], emb[i][2k+1] = Re, Im of mean_j exp(i * t_k * Psi[i][j]).
// One CTA per (row, sample point): a fused sincos reduction over the wavelet row, FP64,
// fixed reduction order (deterministic).
#include <math.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int CF_THREADS = 256;

__global__ void __launch_bounds__(CF_THREADS)
characteristic_function_kernel(const double* __restrict__ psi, int64_t ld, int n_cols,
                               const double* __restrict__ t, int n_t, double* __restrict__ out) {
    __shared__ double red[2][CF_THREADS / 32];
    const int i = blockIdx.y, k = blockIdx.x;
    const double tk = t[k];
    const double* row = psi + (int64_t)i * ld;
    double sc = 0.0, ss = 0.0;
    for (int j = threadIdx.x; j < n_cols; j += CF_THREADS) {
        double s, c;
        sincos(tk * row[j], &s, &c);
        sc += c;
        ss += s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sc += __shfl_down_sync(0xffffffffu, sc, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = sc; red[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0, s = 0.0;
#pragma unroll
        for (int w = 0; w < CF_THREADS / 32; ++w) { c += red[0][w]; s += red[1][w]; }
        out[((int64_t)i * n_t + k) * 2] = c / (double)n_cols;
        out[((int64_t)i * n_t + k) * 2 + 1] = s / (double)n_cols;
    }
}

}  // namespace hsd

extern "C" int hsd_characteristic_function(const double* psi, int64_t psi_ld, int32_t n_rows,
                                           int32_t n_cols, const double* sample_points, int32_t n_points,
                                           double* out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(psi && sample_points && out, "null pointer");
    HSD_REQUIRE(n_rows >= 0 && n_cols > 0 && n_points >= 0 && psi_ld >= n_cols, "bad sizes");
    if (n_rows == 0 || n_points == 0) return HSD_OK;
    HSD_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    dim3 grid(n_points, n_rows);
    characteristic_function_kernel<<<grid, CF_THREADS, 0, (cudaStream_t)stream>>>(psi, psi_ld, n_cols,
                                                                                 sample_points, n_points, out);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
