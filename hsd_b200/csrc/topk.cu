// K7 — k nearest neighbours per row of the (device-resident, possibly sharded) distance matrix:
// the step right after the hot path (SURVEY.md §8 f rank 4).  The reference hands the full N x N
// ndarray to sklearn (tools/evaluate.py:61-69 via main.py:29); at N = 100k that is 40 GB, so the
// selection runs where the matrix lives and only N x k pairs leave the GPU.
//
// One CTA per row: exact radix select on the float bit pattern (distances are >= 0, so the
// unsigned order equals the numeric order) — three histogram passes (11 + 11 + 10 bits) find the
// k-th smallest value, a fourth pass collects everything below it plus ties in ascending column
// order, and a small bitonic sort orders the k results by (distance, column).
#include "hsd_common.cuh"

namespace hsd {

constexpr int TK_THREADS = 256;
constexpr int TK_MAX = 64;

__device__ __forceinline__ bool tk_allowed(const uint32_t* __restrict__ mask, int j) {
    return !mask || ((mask[j >> 5] >> (j & 31)) & 1u);
}

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float* __restrict__ D, int64_t ld, int n_cols, int k, int self_col0,
                 const uint32_t* __restrict__ mask, int32_t* __restrict__ idx_out,
                 float* __restrict__ val_out) {
    __shared__ int hist[2048];
    __shared__ uint32_t sel_prefix, sel_mask;
    __shared__ int sel_need, n_less, n_tie;
    __shared__ float res_v[TK_MAX];
    __shared__ int res_i[TK_MAX];
    __shared__ int warp_tot[TK_THREADS / 32];

    const int row = blockIdx.x;
    const float* d = D + (int64_t)row * ld;
    const int self_col = self_col0 + row;      // excluded: a node is not its own neighbour
    const int tid = threadIdx.x;

    if (tid == 0) { sel_prefix = 0u; sel_mask = 0u; sel_need = k; }
    __syncthreads();
    // ---- three radix passes: bits [31:21], [20:10], [9:0] ----
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        const int nb = 1 << widths[pass];
        for (int b = tid; b < nb; b += TK_THREADS) hist[b] = 0;
        __syncthreads();
        const uint32_t pre = sel_prefix, msk = sel_mask;
        for (int j = tid; j < n_cols; j += TK_THREADS) {
            if (j == self_col || !tk_allowed(mask, j)) continue;
            const uint32_t u = __float_as_uint(d[j]);
            if ((u & msk) == pre) atomicAdd(&hist[(u >> shifts[pass]) & (nb - 1)], 1);
        }
        __syncthreads();
        if (tid == 0) {   // nb <= 2048 sequential adds per row per pass: negligible next to the scan of the row
            int need = sel_need, b = 0;
            while (b < nb - 1 && hist[b] < need) { need -= hist[b]; ++b; }
            sel_need = need;
            sel_prefix = pre | ((uint32_t)b << shifts[pass]);
            sel_mask = msk | ((uint32_t)(nb - 1) << shifts[pass]);
        }
        __syncthreads();
    }
    const uint32_t kth = sel_prefix;           // bit pattern of the k-th smallest distance
    const int need_ties = sel_need;            // how many elements equal to it belong to the result
    if (tid == 0) { n_less = 0; n_tie = 0; }
    __syncthreads();
    // ---- collect: strictly smaller in any order; ties in ascending column order ----
    // Fast path: when exactly `need_ties` elements equal the k-th value they all belong to the
    // result and their order does not matter (the final sort fixes it).  Only when there are more
    // ties than needed does the block fall back to the ordered placement (a scan per 256 columns).
    for (int j = tid; j < n_cols; j += TK_THREADS) {
        if (j == self_col || !tk_allowed(mask, j)) continue;
        const uint32_t u = __float_as_uint(d[j]);
        if (u < kth) {
            const int p = atomicAdd(&n_less, 1);
            res_v[p] = __uint_as_float(u);
            res_i[p] = j;
        } else if (u == kth) {
            const int t = atomicAdd(&n_tie, 1);
            if (t < need_ties) {                       // tentative: valid iff n_tie ends == need_ties
                res_v[(k - need_ties) + t] = __uint_as_float(u);
                res_i[(k - need_ties) + t] = j;
            }
        }
    }
    __syncthreads();
    if (n_tie > need_ties) {                            // uniform across the block
        __syncthreads();
        if (tid == 0) n_tie = 0;
        __syncthreads();
        for (int base = 0; base < n_cols; base += TK_THREADS) {
            const int j = base + tid;
            const bool ok = j < n_cols && j != self_col && tk_allowed(mask, j);
            const int is_tie = (ok && __float_as_uint(d[j]) == kth) ? 1 : 0;
            int total;
            const int before = block_exclusive_scan<TK_THREADS>(is_tie, warp_tot, &total);
            const int t0 = n_tie;                       // read before anyone updates it (scan synced)
            __syncthreads();
            if (is_tie && t0 + before < need_ties) {
                const int p = (k - need_ties) + t0 + before;
                res_v[p] = __uint_as_float(kth);
                res_i[p] = j;
            }
            if (tid == 0) n_tie = t0 + total;
            __syncthreads();
            if (t0 + total >= need_ties) break;         // uniform: every needed tie is placed
        }
    }
    // ---- sort the k results by (distance, column): bitonic over 64 slots ----
    const int avail = min(k, n_less + min(n_tie, need_ties));
    if (tid < TK_MAX && tid >= avail) { res_v[tid] = INFINITY; res_i[tid] = 0x7fffffff; }
    __syncthreads();
    for (int sz = 2; sz <= TK_MAX; sz <<= 1) {
        for (int st = sz >> 1; st > 0; st >>= 1) {
            if (tid < TK_MAX) {
                const int o = tid ^ st;
                if (o > tid) {
                    const bool up = ((tid & sz) == 0);
                    const float a = res_v[tid], b = res_v[o];
                    const int ia = res_i[tid], ib = res_i[o];
                    const bool gt = (a > b) || (a == b && ia > ib);
                    if (gt == up) { res_v[tid] = b; res_v[o] = a; res_i[tid] = ib; res_i[o] = ia; }
                }
            }
            __syncthreads();
        }
    }
    if (tid < k) {
        val_out[(int64_t)row * k + tid] = (tid < avail) ? res_v[tid] : INFINITY;
        idx_out[(int64_t)row * k + tid] = (tid < avail) ? res_i[tid] : -1;
    }
}

}  // namespace hsd

extern "C" int hsd_topk_rows(const float* D, int64_t ld, int32_t n_rows, int32_t n_cols, int32_t k,
                             int32_t self_col0, const uint32_t* col_mask, int32_t* idx_out,
                             float* val_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(D && idx_out && val_out, "null pointer");
    HSD_REQUIRE(n_rows >= 0 && n_cols > 0 && ld >= n_cols, "bad sizes");
    HSD_REQUIRE(k >= 1 && k <= TK_MAX, "k must be in 1..64");
    if (n_rows == 0) return HSD_OK;
    topk_rows_kernel<<<n_rows, TK_THREADS, 0, (cudaStream_t)stream>>>(D, ld, n_cols, k, self_col0, col_mask,
                                                                     idx_out, val_out);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
