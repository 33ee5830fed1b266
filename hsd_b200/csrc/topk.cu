// K7 — k nearest neighbours per row of the (device-resident, possibly sharded) distance matrix:
// the step right after the hot path (SURVEY.md §8 f rank 4).  The reference hands the full N x N
// ndarray to sklearn (tools/evaluate.py:61-69 via main.py:29); at N = 100k that is 40 GB, so the
// selection runs where the matrix lives and only N x k pairs leave the GPU.
//
// One CTA per row, normally ONE read of the row:
//   pass 1  every thread keeps the 4 smallest of the elements it streams (16-byte loads, no atomics);
//           the k-th smallest of those THREADS minima is an upper bound T of the row's k-th smallest
//           value (the minima are distinct elements of the row), and a tight one: the row's k
//           smallest elements mostly sit in different threads, so about k + k^2/THREADS elements are <= T;
//   collect everything <= T goes to a shared candidate list — straight from the threads' registers when
//           no thread can have dropped an element <= T (its 4th smallest is above T), else by pass 2,
//           a second read of the row (heavy ties, very short rows);
//   sort    the candidates by (distance, column) — a single warp when there are <= 64 — and the
//           first k are the answer, ties at the k-th value resolved by ascending column.
// A row with more than TK_CAND candidates (a large class of nodes at exactly the same distance)
// takes the exact radix select instead: three histogram passes over the float bit pattern
// (11 + 11 + 10 bits; distances are >= 0, so the unsigned order equals the numeric order) find the
// k-th smallest value, a fourth pass collects everything below it plus ties in ascending column
// order.  That path alone measured 4.1 ms at N = 20k (shared-atomic contention on the few exponent
// buckets the distances share); HSD_TOPK_RADIX=1 forces it (tests).
#include <stdlib.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int TK_THREADS = 256;
constexpr int TK_MAX = 64;

__device__ __forceinline__ bool tk_allowed(const uint32_t* __restrict__ mask, int j) {
    return !mask || ((mask[j >> 5] >> (j & 31)) & 1u);
}

__device__ __noinline__ void
topk_row_radix(const float* __restrict__ D, int64_t ld, int n_cols, int k, int self_col0,
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

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_radix_kernel(const float* __restrict__ D, int64_t ld, int n_cols, int k, int self_col0,
                       const uint32_t* __restrict__ mask, int32_t* __restrict__ idx_out,
                       float* __restrict__ val_out) {
    topk_row_radix(D, ld, n_cols, k, self_col0, mask, idx_out, val_out);
}

constexpr int TK_CAND = 1024;   // candidate capacity of the two-pass kernel
constexpr int TK_DEEP_MIN_COLS = 32768;   // rows at least this long use the 4-deep register buffers

// (distance, column) "a sorts after b"
__device__ __forceinline__ bool tk_after(float a, int ia, float b, int ib) {
    return (a > b) || (a == b && ia > ib);
}

// Streams the row once: f(value, column) for every admissible column.  VEC4: 16-byte loads (row
// start 16-byte aligned); the 4 mask bits of a group sit in one bitmap word.  PLAIN: no column mask
// and the self column is NOT filtered here (the caller accounts for it), so the loop is loads + f.
template <bool VEC4, bool PLAIN, typename F>
__device__ __forceinline__ void tk_stream_row(const float* __restrict__ d, int n_cols, int self_col,
                                              const uint32_t* __restrict__ mask, F f) {
    const int tid = threadIdx.x;
    if (PLAIN) {
        if (VEC4) {
            const int n4 = n_cols >> 2;
            const float4* d4 = reinterpret_cast<const float4*>(d);
#pragma unroll 4
            for (int g = tid; g < n4; g += TK_THREADS) {
                const float4 v = __ldg(d4 + g);
                const int j = g << 2;
                f(v.x, j); f(v.y, j + 1); f(v.z, j + 2); f(v.w, j + 3);
            }
            const int j = (n4 << 2) + tid;
            if (j < n_cols) f(__ldg(d + j), j);
        } else {
#pragma unroll 4
            for (int j = tid; j < n_cols; j += TK_THREADS) f(__ldg(d + j), j);
        }
        return;
    }
    if (VEC4) {
        const int n4 = n_cols >> 2;
        const float4* d4 = reinterpret_cast<const float4*>(d);
#pragma unroll 4
        for (int g = tid; g < n4; g += TK_THREADS) {
            const float4 v = __ldg(d4 + g);
            const int j = g << 2;
            uint32_t ok = mask ? ((__ldg(mask + (j >> 5)) >> (j & 31)) & 0xfu) : 0xfu;
            if ((unsigned)(self_col - j) < 4u) ok &= ~(1u << (self_col - j));
            if (ok & 1u) f(v.x, j);
            if (ok & 2u) f(v.y, j + 1);
            if (ok & 4u) f(v.z, j + 2);
            if (ok & 8u) f(v.w, j + 3);
        }
        const int j = (n4 << 2) + tid;
        if (j < n_cols && j != self_col && tk_allowed(mask, j)) f(__ldg(d + j), j);
    } else {
#pragma unroll 4
        for (int j = tid; j < n_cols; j += TK_THREADS)
            if (j != self_col && tk_allowed(mask, j)) f(__ldg(d + j), j);
    }
}

// TK_KEEP = smallest elements a thread keeps in registers during pass 1.  4: long rows (the row is
// read once; measured 11.1 -> 8.2 ms at N = 100k).  1: short rows, where the insertions of a 4-deep
// buffer cost more instructions than the second (L2-resident) read saves (N = 20k: 0.36 vs 0.53 ms).
template <bool VEC4, bool PLAIN, int TK_KEEP>
__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const float* __restrict__ D, int64_t ld, int n_cols, int k, int self_col0,
                 const uint32_t* __restrict__ mask, int32_t* __restrict__ idx_out,
                 float* __restrict__ val_out) {
    __shared__ __align__(16) float tmin[TK_THREADS];
    __shared__ float cand_v[TK_CAND];
    __shared__ int cand_i[TK_CAND];
    __shared__ float bound;
    __shared__ int n_cand;

    const int row = blockIdx.x;
    const float* d = D + (int64_t)row * ld;
    const int self_col = self_col0 + row;      // excluded: a node is not its own neighbour
    const int tid = threadIdx.x;

    // ---- pass 1: per-thread TK_KEEP smallest (ascending, in registers); bv[0] is the minimum ----
    float bv[TK_KEEP];
    int bi[TK_KEEP];
#pragma unroll
    for (int q = 0; q < TK_KEEP; ++q) { bv[q] = INFINITY; bi[q] = -1; }
    tk_stream_row<VEC4, PLAIN>(d, n_cols, self_col, mask, [&](float v, int j) {
        if (v < bv[TK_KEEP - 1]) {              // rarely true once the buffer has warmed up
            bv[TK_KEEP - 1] = v; bi[TK_KEEP - 1] = j;
#pragma unroll
            for (int q = TK_KEEP - 1; q > 0; --q) {
                if (bv[q] < bv[q - 1]) {
                    const float tv = bv[q]; bv[q] = bv[q - 1]; bv[q - 1] = tv;
                    const int ti = bi[q]; bi[q] = bi[q - 1]; bi[q - 1] = ti;
                }
            }
        }
    });
    const float mine = bv[0];
    // PLAIN: the self column took part in the minima, so the bound is the (k+1)-th smallest of them —
    // at most one of the k+1 smallest minima is the self element, the other k are neighbours <= bound
    const int kb = PLAIN ? k : k - 1;
    // ---- bound = k-th smallest of the THREADS minima ----
    // Each warp sorts its 32 minima with shuffles; a thread then ranks its own value against the
    // 8 sorted lists by binary search: lt = #minima below it, le = #minima not above it.  The value
    // with lt <= k-1 < le is the k-th smallest (every thread holding it writes the same bound).
    // Threads without an element hold +inf; if fewer than k threads have one, the bound is +inf
    // and the whole (then short) row becomes the candidate list.
    {
        const int lane = tid & 31;
        float x = mine;
#pragma unroll
        for (int sz = 2; sz <= 32; sz <<= 1) {
#pragma unroll
            for (int st = sz >> 1; st > 0; st >>= 1) {
                const float y = __shfl_xor_sync(0xffffffffu, x, st);
                const bool up = ((lane & sz) == 0);
                const bool lower = ((lane & st) == 0);
                x = (lower == up) ? fminf(x, y) : fmaxf(x, y);
            }
        }
        tmin[tid] = x;                          // warp w's minima, ascending, at tmin[32 w ..]
        if (tid == 0) { n_cand = 0; bound = INFINITY; }
        __syncthreads();
        int lt = 0, le = 0;
#pragma unroll
        for (int w = 0; w < TK_THREADS / 32; ++w) {
            const float* a = tmin + 32 * w;
            int p = 0, q = 0;
#pragma unroll
            for (int s2 = 16; s2 >= 1; s2 >>= 1) {
                if (a[p + s2 - 1] < x) p += s2;
                if (a[q + s2 - 1] <= x) q += s2;
            }
            p += (a[p] < x) ? 1 : 0;
            q += (a[q] <= x) ? 1 : 0;
            lt += p;
            le += q;
        }
        if (lt <= kb && kb < le) bound = x;
    }
    __syncthreads();
    // ---- candidates = everything <= bound ----
    // A thread dropped an element only when it was >= its TK_KEEP-th smallest; if that one is above the
    // bound, every element of the thread that is <= bound still sits in its registers and the row is
    // NOT read again.  Otherwise (uniform vote; heavy ties or a very short row) pass 2 re-streams the row.
    const float T = bound;
    auto push = [&](float v, int j) {
        const int p = atomicAdd(&n_cand, 1);
        if (p < TK_CAND) { cand_v[p] = v; cand_i[p] = j; }
    };
    if (TK_KEEP == 1 || __syncthreads_or(bv[TK_KEEP - 1] <= T)) {
        tk_stream_row<VEC4, PLAIN>(d, n_cols, self_col, mask, [&](float v, int j) {
            if (v <= T && (!PLAIN || j != self_col)) push(v, j);
        });
    } else {
#pragma unroll
        for (int q = 0; q < TK_KEEP; ++q)
            if (bv[q] <= T && (!PLAIN || bi[q] != self_col)) push(bv[q], bi[q]);
    }
    __syncthreads();
    const int nc = n_cand;
    if (nc > TK_CAND) {                         // uniform: a huge tie class -> exact radix select
        topk_row_radix(D, ld, n_cols, k, self_col0, mask, idx_out, val_out);
        return;
    }
    // ---- sort the candidates by (distance, column) ----
    int sz_all = 64;
    while (sz_all < nc) sz_all <<= 1;
    for (int i = nc + tid; i < sz_all; i += TK_THREADS) { cand_v[i] = INFINITY; cand_i[i] = 0x7fffffff; }
    __syncthreads();
    if (sz_all == 64) {                         // one warp, one compare-exchange per lane and stage
        if (tid < 32) {
            for (int sz = 2; sz <= 64; sz <<= 1) {
                for (int st = sz >> 1; st > 0; st >>= 1) {
                    const int i = ((tid & ~(st - 1)) << 1) | (tid & (st - 1));
                    const int o = i | st;
                    const bool up = ((i & sz) == 0);
                    const float a = cand_v[i], b = cand_v[o];
                    const int ia = cand_i[i], ib = cand_i[o];
                    if (tk_after(a, ia, b, ib) == up) { cand_v[i] = b; cand_v[o] = a; cand_i[i] = ib; cand_i[o] = ia; }
                    __syncwarp();
                }
            }
        }
    } else {
        for (int sz = 2; sz <= sz_all; sz <<= 1) {
            for (int st = sz >> 1; st > 0; st >>= 1) {
                for (int t = tid; t < (sz_all >> 1); t += TK_THREADS) {
                    const int i = ((t & ~(st - 1)) << 1) | (t & (st - 1));
                    const int o = i | st;
                    const bool up = ((i & sz) == 0);
                    const float a = cand_v[i], b = cand_v[o];
                    const int ia = cand_i[i], ib = cand_i[o];
                    if (tk_after(a, ia, b, ib) == up) { cand_v[i] = b; cand_v[o] = a; cand_i[i] = ib; cand_i[o] = ia; }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    if (tid < k) {
        val_out[(int64_t)row * k + tid] = (tid < nc) ? cand_v[tid] : INFINITY;
        idx_out[(int64_t)row * k + tid] = (tid < nc) ? cand_i[tid] : -1;
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
    static int force_radix = -1;   // test knob: HSD_TOPK_RADIX=1 sends every row through the radix select
    if (force_radix < 0) { const char* e = getenv("HSD_TOPK_RADIX"); force_radix = (e && atoi(e)) ? 1 : 0; }
    cudaStream_t st = (cudaStream_t)stream;
    if (force_radix)
        topk_rows_radix_kernel<<<n_rows, TK_THREADS, 0, st>>>(D, ld, n_cols, k, self_col0, col_mask, idx_out, val_out);
    else {
        const bool vec4 = (reinterpret_cast<uintptr_t>(D) & 15) == 0 && (ld & 3) == 0;
        auto go = [&](auto kern) { kern<<<n_rows, TK_THREADS, 0, st>>>(D, ld, n_cols, k, self_col0, col_mask, idx_out, val_out); };
        static int keep_force = -1;   // tuning knob: HSD_TOPK_KEEP in {1, 4}
        if (keep_force < 0) { const char* e = getenv("HSD_TOPK_KEEP"); keep_force = e ? atoi(e) : 0; }
        const bool deep = keep_force ? keep_force == 4 : n_cols >= TK_DEEP_MIN_COLS;
        if (deep) {
            if (vec4) { if (col_mask) go(topk_rows_kernel<true, false, 4>); else go(topk_rows_kernel<true, true, 4>); }
            else      { if (col_mask) go(topk_rows_kernel<false, false, 4>); else go(topk_rows_kernel<false, true, 4>); }
        } else {
            if (vec4) { if (col_mask) go(topk_rows_kernel<true, false, 1>); else go(topk_rows_kernel<true, true, 1>); }
            else      { if (col_mask) go(topk_rows_kernel<false, false, 1>); else go(topk_rows_kernel<false, true, 1>); }
        }
    }
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
