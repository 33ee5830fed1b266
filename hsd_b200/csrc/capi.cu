// C-ABI plumbing: error string, version, signature transpose, symmetric row/column scatter.
#include <stdarg.h>
#include <string.h>
#include "hsd_common.cuh"

namespace hsd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// sig[r][k] (row-major, ld) -> sigT[k][col0 + r] (ld n_pad); 32x32 tiles through shared memory
__global__ void __launch_bounds__(256)
signature_transpose_kernel(const float* __restrict__ sig, int64_t sig_ld, int n_rows, int k_used,
                           float* __restrict__ sigT, int64_t n_pad, int col0,
                           const int32_t* __restrict__ src_rows) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + ty + q * 8, k = k0 + tx;
        // with src_rows, output column col0 + r is fed from table row src_rows[r] (a gather of
        // whole rows: each read is still a contiguous 128-byte segment)
        const int64_t src = (r < n_rows) ? (src_rows ? (int64_t)__ldg(src_rows + r) : (int64_t)r) : 0;
        tile[ty + q * 8][tx] = (r < n_rows && k < k_used) ? sig[src * sig_ld + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = k0 + ty + q * 8, r = r0 + tx;
        if (k < k_used && r < n_rows) sigT[(int64_t)k * n_pad + col0 + r] = tile[tx][ty + q * 8];
    }
}

// Incremental update (config 5): blk[a][c] = D(idx[a], c) for m recomputed rows.  Stores each value as
// D[idx[a]][c] (rows: coalesced) and, mirrored, as D[c][idx[a]] (a column of a row-major matrix: the
// tile is turned through shared memory so that a warp writes 32 ascending affected columns of ONE row).
// Entries with both indices affected are written twice with the same bits (|a - b| == |b - a|).
__global__ void __launch_bounds__(256)
scatter_symmetric_kernel(const float* __restrict__ blk, int64_t blk_ld, int m, int n,
                         const int64_t* __restrict__ idx, float* __restrict__ D, int64_t d_ld, int mirror) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int a = a0 + ty + q * 8, c = c0 + tx;
        float v = 0.f;
        if (a < m && c < n) {
            v = blk[(int64_t)a * blk_ld + c];
            D[__ldg(idx + a) * d_ld + c] = v;
        }
        tile[ty + q * 8][tx] = v;
    }
    if (!mirror) return;
    __syncthreads();
    const int a = a0 + tx;
    const int64_t col = (a < m) ? __ldg(idx + a) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int c = c0 + ty + q * 8;
        if (a < m && c < n) D[(int64_t)c * d_ld + col] = tile[tx][ty + q * 8];
    }
}

}  // namespace hsd

extern "C" int hsd_version(void) { return 100; }

extern "C" const char* hsd_last_error_string(void) { return hsd::g_err; }

extern "C" int hsd_signature_transpose(const float* sig, int64_t sig_ld, int32_t n_rows,
                                       int32_t k_used, float* sigT, int64_t n_pad, int32_t col0,
                                       const int32_t* src_rows, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sig && sigT, "null pointer");
    HSD_REQUIRE(n_rows >= 0 && k_used >= 0 && sig_ld >= k_used, "bad sizes");
    HSD_REQUIRE(col0 >= 0 && col0 + (int64_t)n_rows <= n_pad, "column range exceeds n_pad");
    if (n_rows == 0 || k_used == 0) return HSD_OK;
    dim3 grid((n_rows + 31) / 32, (k_used + 31) / 32);
    signature_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sig, sig_ld, n_rows, k_used,
                                                                      sigT, n_pad, col0, src_rows);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_scatter_symmetric(const float* blk, int64_t blk_ld, int32_t m, int32_t n,
                                     const int64_t* idx, float* D, int64_t d_ld, int32_t mirror,
                                     void* stream) {
    using namespace hsd;
    HSD_REQUIRE(blk && idx && D, "null pointer");
    HSD_REQUIRE(m >= 0 && n >= 0 && blk_ld >= n && d_ld >= n, "bad sizes");
    if (m == 0 || n == 0) return HSD_OK;
    dim3 grid((n + 31) / 32, (m + 31) / 32);
    HSD_REQUIRE(grid.y <= 65535u, "too many rows for one launch");
    scatter_symmetric_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(blk, blk_ld, m, n, idx, D, d_ld, mirror ? 1 : 0);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_copy2d_to_host(void* dst_host, int64_t dst_pitch_bytes, const void* src_dev,
                                  int64_t src_pitch_bytes, int64_t width_bytes, int64_t rows, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(dst_host && src_dev, "null pointer");
    HSD_REQUIRE(width_bytes >= 0 && rows >= 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
                "bad window");
    if (width_bytes == 0 || rows == 0) return HSD_OK;
    HSD_CUDA_TRY(cudaMemcpy2DAsync(dst_host, (size_t)dst_pitch_bytes, src_dev, (size_t)src_pitch_bytes,
                                   (size_t)width_bytes, (size_t)rows, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return HSD_OK;
}
