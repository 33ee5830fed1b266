// Value-mode ring signals (the reference-faithful wavelet path, config 1):
//   K2v  hsd_ring_signature_values — gather Psi[i, ring_h(i)] and sort ascending
//        (model/HSD.py:71-83 + the argsort inside scipy's _cdf_distance);
//   K3v  hsd_pairwise_w1_merge     — exact ragged 1-D Wasserstein by two-pointer
//        merge of two ascending lists (scipy.stats.wasserstein_distance as called
//        at model/HSD.py:111);
//   K3a  hsd_pairwise_aligned      — the zero-pad + sort variant of
//        tools/metrics.py:18-36,151-192 ('wasserstein' and 'hellinger').
// All three run in FP64: these graphs are reference-sized (N ~ 1e3) and the
// kernels are latency/divergence bound, so FP64 costs nothing and gives parity
// to ~1e-13 instead of 1e-6.
#include <math.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int SORT_THREADS = 256;
constexpr int MAX_SORT = 16384;

__global__ void __launch_bounds__(SORT_THREADS)
ring_values_kernel(const double* __restrict__ psi, int64_t psi_ld,
                   const uint32_t* __restrict__ bitmaps, const int32_t* __restrict__ sizes,
                   const int64_t* __restrict__ offsets, const int32_t* __restrict__ orig_of,
                   int hops1, int n_words, int cap, double* __restrict__ vals) {
    extern __shared__ double sv[];
    __shared__ int counter;
    const int seg = blockIdx.x;  // row * hops1 + h
    const int row = seg / hops1;
    const int n = sizes[seg];
    if (n == 0) return;
    int m = 1;
    while (m < n) m <<= 1;  // bitonic size, <= cap
    if (threadIdx.x == 0) counter = 0;
    for (int i = threadIdx.x; i < m; i += SORT_THREADS) sv[i] = INFINITY;
    __syncthreads();
    const uint32_t* bm = bitmaps + (int64_t)seg * n_words;
    const double* prow = psi + (int64_t)row * psi_ld;
    for (int w = threadIdx.x; w < n_words; w += SORT_THREADS) {
        uint32_t bits = bm[w];
        while (bits) {
            const int j = (w << 5) + __ffs(bits) - 1;
            bits &= bits - 1;
            const int pos = atomicAdd(&counter, 1);
            sv[pos] = prow[orig_of ? orig_of[j] : j];
        }
    }
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += SORT_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const double a = sv[i], b = sv[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { sv[i] = b; sv[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    double* dst = vals + offsets[seg];
    for (int i = threadIdx.x; i < n; i += SORT_THREADS) dst[i] = sv[i];
}

// W1 between two ascending lists (scipy _cdf_distance, p = 1): walk the merged
// sequence, between consecutive values x_k < x_{k+1} the CDFs are iu/nu, iv/nv.
__device__ __forceinline__ double w1_merge(const double* __restrict__ u, int nu,
                                           const double* __restrict__ v, int nv) {
    int iu = 0, iv = 0;
    const double inu = 1.0 / (double)nu, inv = 1.0 / (double)nv;
    double prev, acc = 0.0;
    {
        const double a = u[0], b = v[0];
        if (a <= b) { prev = a; iu = 1; } else { prev = b; iv = 1; }
    }
    while (iu < nu || iv < nv) {
        double x;
        bool take_u;
        if (iu < nu && iv < nv) {
            const double a = u[iu], b = v[iv];
            take_u = (a <= b);
            x = take_u ? a : b;
        } else if (iu < nu) {
            take_u = true; x = u[iu];
        } else {
            take_u = false; x = v[iv];
        }
        acc += fabs((double)iu * inu - (double)iv * inv) * (x - prev);
        prev = x;
        if (take_u) ++iu; else ++iv;
    }
    return acc;
}

__global__ void __launch_bounds__(256)
pairwise_w1_merge_kernel(const double* __restrict__ vals, const int64_t* __restrict__ offsets,
                         const int32_t* __restrict__ sizes, int n_total, int hops1, int hop_begin,
                         int hop_end, int row0, int n_rows, double* __restrict__ out, int64_t ld,
                         int32_t* status) {
    const int i = row0 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row0 + n_rows || j >= n_total || j < i) return;
    if (j == i) { out[(int64_t)i * ld + i] = 0.0; return; }
    double d = 0.0;
    for (int h = hop_begin; h < hop_end; ++h) {
        const int si = i * hops1 + h, sj = j * hops1 + h;
        const int ni = sizes[si], nj = sizes[sj];
        if (ni == 0 || nj == 0) { atomicOr(status, 1); continue; }
        d += w1_merge(vals + offsets[si], ni, vals + offsets[sj], nj);
    }
    out[(int64_t)i * ld + j] = d;
    out[(int64_t)j * ld + i] = d;
}

// ascending zero-padded sequence generator (tools/metrics.py:27-30: pad with zeros, np.sort)
struct PaddedAsc {
    const double* p; int n; int idx; int zeros;
    __device__ PaddedAsc(const double* p_, int n_, int L) : p(p_), n(n_), idx(0), zeros(L - n_) {}
    __device__ double next() {
        if (idx < n && (p[idx] < 0.0 || zeros == 0)) return p[idx++];
        --zeros;
        return 0.0;
    }
};

// The three metrics of tools/metrics.py::calculate_distance that the HSD path reaches, on two
// zero-padded ascending sequences of equal length L.  MakeP / MakeQ build a fresh stream (the
// Gaussian metric needs two passes).
//   0 'wasserstein'        scipy on equal-length samples == mean |p_(k) - q_(k)|   (:186-190)
//   1 'hellinger'          sqrt(max(1 - sum sqrt(p q), 0)) with 1e-6 snapping        (:117-138)
//   2 'wasserstein_guass'  (u1-u2)^2 + s1 + s2 - 2 sqrt(s1 s2), s = population var   (:54-71)
template <typename MakeP, typename MakeQ>
__device__ __forceinline__ double aligned_metric(int metric, int L, MakeP make_p, MakeQ make_q) {
    auto P = make_p();
    auto Q = make_q();
    if (metric == 0) {
        double s = 0.0;
        for (int k = 0; k < L; ++k) s += fabs(P.next() - Q.next());
        return s / (double)L;
    }
    if (metric == 1) {
        double bc = 0.0;
        for (int k = 0; k < L; ++k) {
            const double px = P.next(), qx = Q.next();
            if (px < 0.0 || qx < 0.0) continue;
            bc += sqrt(fmax(px * qx, 0.0));
        }
        if (fabs(bc) <= 1e-6) bc = 0.0;
        else if (fabs(bc - 1.0) <= fmax(1e-9 * fmax(fabs(bc), 1.0), 1e-6)) bc = 1.0;
        return sqrt(fmax(1.0 - bc, 0.0));
    }
    double sp = 0.0, sq = 0.0;
    for (int k = 0; k < L; ++k) { sp += P.next(); sq += Q.next(); }
    const double u1 = sp / (double)L, u2 = sq / (double)L;
    auto P2 = make_p();
    auto Q2 = make_q();
    double vp = 0.0, vq = 0.0;
    for (int k = 0; k < L; ++k) {
        const double a = P2.next() - u1, b = Q2.next() - u2;
        vp += a * a;
        vq += b * b;
    }
    const double s1 = vp / (double)L, s2 = vq / (double)L;
    return (u1 - u2) * (u1 - u2) + s1 + s2 - 2.0 * sqrt(s1 * s2);
}

__global__ void __launch_bounds__(256)
pairwise_aligned_kernel(const double* __restrict__ vals, const int64_t* __restrict__ offsets,
                        const int32_t* __restrict__ sizes, int n_total, int hops1, int hop_begin,
                        int hop_end, int metric, int row0, int n_rows, double* __restrict__ out,
                        int64_t ld) {
    const int i = row0 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row0 + n_rows || j >= n_total || j < i) return;
    if (j == i) { out[(int64_t)i * ld + i] = 0.0; return; }
    double d = 0.0;
    for (int h = hop_begin; h < hop_end; ++h) {
        const int si = i * hops1 + h, sj = j * hops1 + h;
        const int ni = sizes[si], nj = sizes[sj];
        const int L = max(ni, nj);
        if (L == 0) continue;  // tools/metrics.py:170-171
        const double* pi = vals + offsets[si];
        const double* pj = vals + offsets[sj];
        d += aligned_metric(metric, L, [&] { return PaddedAsc(pi, ni, L); }, [&] { return PaddedAsc(pj, nj, L); });
    }
    out[(int64_t)i * ld + j] = d;
    out[(int64_t)j * ld + i] = d;
}


// ---- K3w: the reference's row worker exactly as written (model/HSD.py:140-161) ----
// d(i, j) = sum_{h < hop_end} aligned(Psi[i, ring_h(i)], Psi[i, ring_h(j)]) — BOTH signals
// are read from wavelet row i (the reference indexes q with startIndex, :155).
// order[i][t] = node with the t-th smallest Psi[i, .]; walking it while testing ring
// membership yields both ascending sequences without any per-pair sort.
struct RingStream {
    const double* sv; const int* ord; const int32_t* bit_of; const uint32_t* bm;
    int n_nodes, t, zeros;
    double head; bool has;
    __device__ RingStream(const double* sv_, const int* ord_, const int32_t* bit_of_, const uint32_t* bm_,
                          int n_nodes_, int n_members, int L)
        : sv(sv_), ord(ord_), bit_of(bit_of_), bm(bm_), n_nodes(n_nodes_), t(0), zeros(L - n_members),
          head(0.0), has(false) { advance(); }
    __device__ void advance() {
        has = false;
        while (t < n_nodes) {
            const int node = ord[t];
            const int b = bit_of ? bit_of[node] : node;
            if ((bm[b >> 5] >> (b & 31)) & 1u) { head = sv[t]; has = true; ++t; return; }
            ++t;
        }
    }
    __device__ double next() {  // ascending zero-padded sequence (tools/metrics.py:27-30)
        if (has && (head < 0.0 || zeros == 0)) { const double x = head; advance(); return x; }
        --zeros;
        return 0.0;
    }
};

__global__ void __launch_bounds__(128)
pairwise_worker_kernel(const double* __restrict__ sorted_vals, const int32_t* __restrict__ order,
                       const uint32_t* __restrict__ bitmaps, const int32_t* __restrict__ sizes,
                       const int32_t* __restrict__ bit_of, int n, int hops1, int n_words, int hop_end,
                       int metric, int row0, double* __restrict__ out, int64_t ld) {
    const int i = row0 + blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j <= i) { out[(int64_t)i * ld + j] = 0.0; return; }  // the worker leaves columns <= startIndex at 0
    const double* sv = sorted_vals + (int64_t)i * n;
    const int* ord = order + (int64_t)i * n;
    double d = 0.0;
    for (int h = 0; h < hop_end; ++h) {
        const int ni = sizes[i * hops1 + h], nj = sizes[j * hops1 + h];
        const int L = max(ni, nj);
        if (L == 0) continue;
        const uint32_t* bi = bitmaps + ((int64_t)i * hops1 + h) * n_words;
        const uint32_t* bj = bitmaps + ((int64_t)j * hops1 + h) * n_words;
        d += aligned_metric(metric, L, [&] { return RingStream(sv, ord, bit_of, bi, n, ni, L); },
                            [&] { return RingStream(sv, ord, bit_of, bj, n, nj, L); });
    }
    out[(int64_t)i * ld + j] = d;
}

}  // namespace hsd

extern "C" int hsd_ring_signature_values(const double* psi, int64_t psi_ld,
                                         const uint32_t* ring_bitmaps, const int32_t* ring_sizes,
                                         const int64_t* offsets, const int32_t* orig_of,
                                         int32_t n_rows, int32_t hops, int32_t n_nodes,
                                         int32_t max_ring_size, double* vals, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(psi && ring_bitmaps && ring_sizes && offsets && vals, "null pointer");
    HSD_REQUIRE(n_rows >= 0 && hops >= 0 && n_nodes > 0 && max_ring_size >= 0, "bad sizes");
    if (n_rows == 0) return HSD_OK;
    int cap = 1;
    while (cap < max_ring_size) cap <<= 1;
    if (cap > MAX_SORT) {
        set_error("hsd_ring_signature_values: ring of %d members exceeds the %d-element "
                  "shared-memory sort", max_ring_size, MAX_SORT);
        return HSD_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)cap * sizeof(double);
    HSD_CUDA_TRY(cudaFuncSetAttribute(ring_values_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    ring_values_kernel<<<n_rows * (hops + 1), SORT_THREADS, smem, (cudaStream_t)stream>>>(
        psi, psi_ld, ring_bitmaps, ring_sizes, offsets, orig_of, hops + 1, (n_nodes + 31) / 32, cap, vals);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_pairwise_w1_merge(const double* vals, const int64_t* offsets,
                                     const int32_t* ring_sizes, int32_t n_total, int32_t hops,
                                     int32_t hop_begin, int32_t hop_end, int32_t row0,
                                     int32_t n_rows, double* out, int64_t ld_out, int32_t* status,
                                     void* stream) {
    using namespace hsd;
    HSD_REQUIRE(vals && offsets && ring_sizes && out && status, "null pointer");
    HSD_REQUIRE(n_total > 0 && hops >= 0 && 0 <= hop_begin && hop_begin <= hop_end && hop_end <= hops + 1,
                "bad hop range");
    HSD_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= n_total && ld_out >= n_total, "bad row range");
    if (n_rows == 0) return HSD_OK;
    HSD_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    dim3 grid((n_total + 255) / 256, n_rows);
    pairwise_w1_merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        vals, offsets, ring_sizes, n_total, hops + 1, hop_begin, hop_end, row0, n_rows, out, ld_out, status);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_pairwise_aligned(const double* vals, const int64_t* offsets,
                                    const int32_t* ring_sizes, int32_t n_total, int32_t hops,
                                    int32_t hop_begin, int32_t hop_end, int32_t metric, int32_t row0,
                                    int32_t n_rows, double* out, int64_t ld_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(vals && offsets && ring_sizes && out, "null pointer");
    HSD_REQUIRE(metric >= 0 && metric <= 2, "metric must be 0 (wasserstein), 1 (hellinger) or 2 (wasserstein_guass)");
    HSD_REQUIRE(n_total > 0 && hops >= 0 && 0 <= hop_begin && hop_begin <= hop_end && hop_end <= hops + 1,
                "bad hop range");
    HSD_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= n_total && ld_out >= n_total, "bad row range");
    if (n_rows == 0) return HSD_OK;
    HSD_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    dim3 grid((n_total + 255) / 256, n_rows);
    pairwise_aligned_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        vals, offsets, ring_sizes, n_total, hops + 1, hop_begin, hop_end, metric, row0, n_rows, out, ld_out);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_pairwise_worker(const double* sorted_vals, const int32_t* order,
                                   const uint32_t* ring_bitmaps, const int32_t* ring_sizes,
                                   const int32_t* bit_of, int32_t n_nodes, int32_t hops,
                                   int32_t hop_end, int32_t metric, int32_t row0, int32_t n_rows,
                                   double* out, int64_t ld_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sorted_vals && order && ring_bitmaps && ring_sizes && out, "null pointer");
    HSD_REQUIRE(metric >= 0 && metric <= 2, "metric must be 0 (wasserstein), 1 (hellinger) or 2 (wasserstein_guass)");
    HSD_REQUIRE(n_nodes > 0 && hops >= 0 && hop_end >= 0 && hop_end <= hops + 1, "bad hop range");
    HSD_REQUIRE(row0 >= 0 && n_rows >= 0 && row0 + n_rows <= n_nodes && ld_out >= n_nodes, "bad row range");
    if (n_rows == 0) return HSD_OK;
    HSD_REQUIRE(n_rows <= 65535, "at most 65535 rows per call");
    dim3 grid((n_nodes + 127) / 128, n_rows);
    pairwise_worker_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        sorted_vals, order, ring_bitmaps, ring_sizes, bit_of, n_nodes, hops + 1, (n_nodes + 31) / 32,
        hop_end, metric, row0, out, ld_out);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
