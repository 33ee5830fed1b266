// K4 — Chebyshev heat-kernel wavelets as a CSR SpMM, and K5 — the ring
// gather-reduce that turns wavelet columns into MultiHSD embeddings.
//
// Reference loops replaced: model/HSD.py:50-59 (pygsp cheby_op applied to one
// unit impulse at a time: N x order SpMVs) and model/multiscale_HSD.py:45-61
// (Python gather of Psi[i, ring_h(i)] -> [sum, mean]).
//
// The recurrence runs on a block of impulse columns at once, node-major
// ([node][column], columns contiguous) so neighbour rows are read with
// coalesced 128-byte requests; all scales share T_k and are accumulated in the
// same pass.  FP64 throughout: the SpMM is bandwidth-bound and FP32 cannot hold
// 1e-5 relative on coefficients that span nine orders of magnitude.
#include <stdlib.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int MAX_SCALES = 8;

struct ChebArgs {
    const int32_t* rowptr;
    const int32_t* col;
    int n_nodes, n_cols, col0;
    int k, order, n_scales;
    double a;          // lmax / 2
    double threshold;
    // Accumulation is DEFERRED: a thread holds T_k, T_{k-1} and T_{k-2} of its element anyway, so the
    // S accumulators are read-modify-written only every third step (acc != 0) with three terms at
    // once: out_s += w2[s] T_{k-2} + w1[s] T_{k-1} + w0[s] T_k (weights of terms already added are 0;
    // the k = 0 term carries c_{s,0}/2).  acc_first: nothing accumulated yet, do not read out.
    int acc, acc_first;
    double w0[MAX_SCALES], w1[MAX_SCALES], w2[MAX_SCALES];
    const double* t_prev;    // T_{k-1}
    const double* t_prev2;   // T_{k-2} (k >= 2)
    double* t_new;
    double* out;             // [n_scales][n_nodes][n_cols]
};

__global__ void cheb_init_kernel(double* t0, int n_nodes, int n_cols, int col0) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_nodes * n_cols) return;
    const int v = (int)(idx / n_cols), c = (int)(idx - (int64_t)v * n_cols);
    t0[idx] = (v == col0 + c) ? 1.0 : 0.0;
}

// one thread per (node v, column PAIR): 16-byte loads/stores, blockDim.x spans column pairs,
// blockDim.y nodes.  The neighbour loop is unrolled 4x with the column indices loaded first so
// four independent row gathers are in flight per thread (the first version, one element per
// thread with a dependent col -> row chain, sat at 46 % of DRAM bandwidth with 26 % issue-active:
// latency-bound).  Accumulators and T_{k-2} are streamed (ld/st .cs) so that T_{k-1}, which is
// re-read once per neighbour, keeps its place in L2.
__device__ __forceinline__ double2 ld_cs(const double* p) {
    double2 v;
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs(double* p, double2 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__global__ void __launch_bounds__(256) cheb_step_kernel(const ChebArgs p) {
    const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 2;     // grid.x carries the nodes (no 65 535 limit)
    const int v = blockIdx.x * blockDim.y + threadIdx.y;
    if (c >= p.n_cols || v >= p.n_nodes) return;
    const int e0 = __ldg(p.rowptr + v), e1 = __ldg(p.rowptr + v + 1);
    double2 nb = make_double2(0.0, 0.0);
    int deg = 0;
    const double* tp_base = p.t_prev + c;
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
        const int u0 = __ldg(p.col + e), u1 = __ldg(p.col + e + 1), u2 = __ldg(p.col + e + 2), u3 = __ldg(p.col + e + 3);
        const double2 x0 = *reinterpret_cast<const double2*>(tp_base + (int64_t)u0 * p.n_cols);
        const double2 x1 = *reinterpret_cast<const double2*>(tp_base + (int64_t)u1 * p.n_cols);
        const double2 x2 = *reinterpret_cast<const double2*>(tp_base + (int64_t)u2 * p.n_cols);
        const double2 x3 = *reinterpret_cast<const double2*>(tp_base + (int64_t)u3 * p.n_cols);
        // nx.laplacian_matrix: a self-loop cancels out of D - A
        if (u0 != v) { nb.x += x0.x; nb.y += x0.y; ++deg; }
        if (u1 != v) { nb.x += x1.x; nb.y += x1.y; ++deg; }
        if (u2 != v) { nb.x += x2.x; nb.y += x2.y; ++deg; }
        if (u3 != v) { nb.x += x3.x; nb.y += x3.y; ++deg; }
    }
    for (; e < e1; ++e) {
        const int u = __ldg(p.col + e);
        if (u == v) continue;
        const double2 x = *reinterpret_cast<const double2*>(tp_base + (int64_t)u * p.n_cols);
        nb.x += x.x; nb.y += x.y;
        ++deg;
    }
    const int64_t idx = (int64_t)v * p.n_cols + c;
    const double2 tp = *reinterpret_cast<const double2*>(p.t_prev + idx);
    const double dm = (double)deg - p.a;
    double2 t;  // ((L - a I) T_{k-1})[v][c..c+1], then the recurrence
    t.x = dm * tp.x - nb.x;
    t.y = dm * tp.y - nb.y;
    double2 t2 = make_double2(0.0, 0.0);
    if (p.k == 1) {
        t.x /= p.a; t.y /= p.a;
    } else {
        t2 = ld_cs(p.t_prev2 + idx);
        t.x = (2.0 / p.a) * t.x - t2.x;
        t.y = (2.0 / p.a) * t.y - t2.y;
    }
    *reinterpret_cast<double2*>(p.t_new + idx) = t;
    if (!p.acc) return;
    const int64_t plane = (int64_t)p.n_nodes * p.n_cols;
    const bool last = (p.k == p.order);
#pragma unroll
    for (int s = 0; s < MAX_SCALES; ++s) {
        if (s >= p.n_scales) break;
        double2 r = make_double2(0.0, 0.0);
        if (!p.acc_first) r = ld_cs(p.out + s * plane + idx);
        r.x += p.w2[s] * t2.x; r.y += p.w2[s] * t2.y;      // ascending k, like the oracle's running sum
        r.x += p.w1[s] * tp.x; r.y += p.w1[s] * tp.y;
        r.x += p.w0[s] * t.x;  r.y += p.w0[s] * t.y;
        if (last) {  // model/HSD.py:65
            r.x = (r.x > p.threshold) ? r.x : 0.0;
            r.y = (r.y > p.threshold) ? r.y : 0.0;
        }
        st_cs(p.out + s * plane + idx, r);
    }
}

// y = L x for one vector (L = D - A from the CSR, self-loops cancel): one warp per row.  Used by the
// lmax estimate (power iteration), the device counterpart of pygsp's Graph.estimate_lmax.
__global__ void __launch_bounds__(256)
laplacian_spmv_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
                      const double* __restrict__ x, double* __restrict__ y) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const int e0 = __ldg(rowptr + v), e1 = __ldg(rowptr + v + 1);
    double nb = 0.0;
    int deg = 0;
    for (int e = e0 + lane; e < e1; e += 32) {
        const int u = __ldg(col + e);
        if (u != v) { nb += x[u]; ++deg; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nb += __shfl_down_sync(0xffffffffu, nb, o);
        deg += __shfl_down_sync(0xffffffffu, deg, o);
    }
    if (lane == 0) y[v] = (double)deg * x[v] - nb;
}

// order == 0 degenerate case: R = c0/2 * E (+ threshold)
__global__ void cheb_order0_kernel(const double* t0, double* out, int64_t n, double c0, double thr) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double r = 0.5 * c0 * t0[idx];
    out[idx] = (r > thr) ? r : 0.0;
}

// ---------------- K5 ----------------
constexpr int RR_MAX_HOPS1 = 8;

__global__ void __launch_bounds__(128)
ring_reduce_kernel(const double* __restrict__ psiT, int n_nodes, int n_cols,
                   const uint32_t* __restrict__ bitmaps, const int32_t* __restrict__ orig_of,
                   int hops1, int n_words, int col0, int v_chunk, int n_scales,
                   double* __restrict__ partial) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.z;
    if (c >= n_cols) return;
    const int v_lo = blockIdx.y * v_chunk, v_hi = min(v_lo + v_chunk, n_nodes);  // multiples of 32
    const uint32_t* bm = bitmaps + (int64_t)(col0 + c) * hops1 * n_words;
    const double* ps = psiT + (int64_t)s * n_nodes * n_cols + c;
    double acc[RR_MAX_HOPS1];
#pragma unroll
    for (int h = 0; h < RR_MAX_HOPS1; ++h) acc[h] = 0.0;
    for (int w = v_lo >> 5; w < (v_hi + 31) >> 5; ++w) {
        uint32_t wh[RR_MAX_HOPS1];
        uint32_t any = 0;
#pragma unroll
        for (int h = 0; h < RR_MAX_HOPS1; ++h) {
            wh[h] = (h < hops1) ? bm[(int64_t)h * n_words + w] : 0u;
            any |= wh[h];
        }
        // every thread walks the same 32 node slots so psiT reads stay coalesced across c
        const int nv = min(32, n_nodes - (w << 5));
        for (int b = 0; b < nv; ++b) {
            const int vid = (w << 5) + b;
            const int vrow = orig_of ? __ldg(orig_of + vid) : vid;
            const double x = ps[(int64_t)vrow * n_cols];
            const uint32_t m = 1u << b;
            if (any & m) {
#pragma unroll
                for (int h = 0; h < RR_MAX_HOPS1; ++h)
                    if (wh[h] & m) acc[h] += x;
            }
        }
    }
    // partial[chunk][c][s][h]: plain stores, summed in chunk order by ring_mean_kernel, so the
    // result does not depend on scheduling (FP64 atomics would make it vary in the last bits)
    double* dst = partial + (((int64_t)blockIdx.y * n_cols + c) * n_scales + s) * hops1;
#pragma unroll
    for (int h = 0; h < RR_MAX_HOPS1; ++h)
        if (h < hops1) dst[h] = acc[h];
}

__global__ void ring_mean_kernel(double* emb, const double* __restrict__ partial, int chunks,
                                 const int32_t* __restrict__ sizes, int n_cols,
                                 int col0, int n_scales, int hops1) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = n_cols * n_scales * hops1;
    if (idx >= total) return;
    const int h = idx % hops1;
    const int c = idx / (hops1 * n_scales);
    const int s = (idx / hops1) % n_scales;
    const int n = sizes[(int64_t)(col0 + c) * hops1 + h];
    double* e = emb + (((int64_t)(col0 + c) * n_scales + s) * hops1 + h) * 2;
    double sum = 0.0;
    for (int k = 0; k < chunks; ++k) sum += partial[(((int64_t)k * n_cols + c) * n_scales + s) * hops1 + h];
    if (n > 0) { e[0] = sum; e[1] = sum / (double)n; }
    else { e[0] = 0.0; e[1] = 0.0; }
}

}  // namespace hsd

extern "C" int hsd_cheb_spmm(const int32_t* rowptr, const int32_t* col, int32_t n_nodes, double lmax,
                             const double* coeff_host, int32_t n_scales, int32_t order, int32_t col0,
                             int32_t n_cols, double threshold, double* work, double* out,
                             void* stream_) {
    using namespace hsd;
    cudaStream_t stream = (cudaStream_t)stream_;
    HSD_REQUIRE(rowptr && col && coeff_host && work && out, "null pointer");
    // one column past the last node is allowed (an all-zero impulse) so odd N can be padded to a pair
    HSD_REQUIRE(n_nodes > 0 && n_cols > 0 && col0 >= 0 && col0 + n_cols <= n_nodes + 1, "bad column block");
    HSD_REQUIRE(n_scales >= 1 && n_scales <= MAX_SCALES, "1..8 scales per call");
    HSD_REQUIRE(order >= 0 && lmax > 0.0, "bad order / lmax");
    const int64_t plane = (int64_t)n_nodes * n_cols;
    double* T[3] = {work, work + plane, work + 2 * plane};
    {
        const int64_t nblk = (plane + 255) / 256;
        HSD_REQUIRE(nblk < (1ll << 31), "column block too large");
        cheb_init_kernel<<<(unsigned)nblk, 256, 0, stream>>>(T[0], n_nodes, n_cols, col0);
        HSD_CUDA_TRY(cudaGetLastError());
        if (order == 0) {
            for (int s = 0; s < n_scales; ++s)
                cheb_order0_kernel<<<(unsigned)nblk, 256, 0, stream>>>(T[0], out + s * plane, plane,
                                                                      coeff_host[s], threshold);
            HSD_CUDA_TRY(cudaGetLastError());
            return HSD_OK;
        }
    }
    HSD_REQUIRE(n_cols % 2 == 0, "n_cols must be even (16-byte column pairs)");
    const int pairs = n_cols / 2;
    // 256 columns and more: one node per 128-thread CTA, so that a CTA's warps all walk the same adjacency
    // list and retire together (a hub next to a leaf in one CTA left the leaf's warps idle: 5.56 -> 5.35 ms
    // per block at C4); narrower blocks keep 256 threads (several nodes per CTA measured faster there)
    static int cta_force = -1;   // tuning knob: HSD_CHEB_CTA = threads per CTA (64, 128, 256)
    if (cta_force < 0) { const char* e = getenv("HSD_CHEB_CTA"); cta_force = e ? atoi(e) : 0; }
    int cta = (cta_force == 64 || cta_force == 128 || cta_force == 256) ? cta_force : (pairs >= 128 ? 128 : 256);
    int bx = 32;
    while (bx < pairs && bx < cta) bx <<= 1;
    const int by = cta / bx;
    dim3 block(bx, by), grid((n_nodes + by - 1) / by, (pairs + bx - 1) / bx);
    HSD_REQUIRE(grid.y <= 65535u, "column block too wide for one launch dimension");
    ChebArgs a;
    a.rowptr = rowptr; a.col = col; a.n_nodes = n_nodes; a.n_cols = n_cols; a.col0 = col0;
    a.order = order; a.n_scales = n_scales; a.a = lmax / 2.0; a.threshold = threshold; a.out = out;
    static int every = -1;   // tuning knob: HSD_CHEB_ACC_EVERY=1 accumulates at every step (the first version)
    if (every < 0) { const char* e = getenv("HSD_CHEB_ACC_EVERY"); every = e ? atoi(e) : 3; if (every < 1 || every > 3) every = 3; }
    int next_term = 0;       // lowest k whose c_k T_k has not been added to the accumulators yet
    for (int k = 1; k <= order; ++k) {
        a.k = k;
        // accumulate at k = 2, 5, 8, ... (terms k-2..k, all in registers at that step) and at the last step
        a.acc = (k == order || k % every == every - 1) ? 1 : 0;
        a.acc_first = (next_term == 0) ? 1 : 0;
        if (a.acc) {
            for (int s = 0; s < n_scales; ++s) {
                const double* cs = coeff_host + (int64_t)s * (order + 1);
                auto w = [&](int j) { return j < next_term ? 0.0 : (j == 0 ? 0.5 * cs[0] : cs[j]); };
                a.w0[s] = w(k);
                a.w1[s] = w(k - 1);
                a.w2[s] = (k >= 2) ? w(k - 2) : 0.0;
            }
            next_term = k + 1;
        }
        a.t_prev = T[(k - 1) % 3];
        a.t_prev2 = T[(k + 1) % 3];  // == (k-2) mod 3
        a.t_new = T[k % 3];
        cheb_step_kernel<<<grid, block, 0, stream>>>(a);
    }
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_ring_reduce(const double* psiT, int32_t n_scales, int32_t n_nodes, int32_t n_cols,
                               const uint32_t* ring_bitmaps, const int32_t* ring_sizes,
                               const int32_t* orig_of, int32_t hops, int32_t col0, double* emb,
                               double* scratch, int64_t scratch_elems, void* stream_) {
    using namespace hsd;
    cudaStream_t stream = (cudaStream_t)stream_;
    HSD_REQUIRE(psiT && ring_bitmaps && ring_sizes && emb && scratch, "null pointer");
    HSD_REQUIRE(n_scales >= 1 && n_nodes > 0 && n_cols > 0 && col0 >= 0 && col0 + n_cols <= n_nodes, "bad sizes");
    HSD_REQUIRE(hops >= 0 && hops + 1 <= RR_MAX_HOPS1, "hops must be <= 7");
    HSD_REQUIRE(n_scales <= 65535, "too many scales");
    const int hops1 = hops + 1, n_words = (n_nodes + 31) / 32;
    // split the node range so the grid has enough CTAs to pull HBM bandwidth
    int chunks = (148 * 8) / (((n_cols + 127) / 128) * n_scales);
    const int64_t per_chunk = (int64_t)n_cols * n_scales * hops1;
    HSD_REQUIRE(scratch_elems >= per_chunk, "scratch must hold at least n_cols * n_scales * (hops+1) doubles");
    if ((int64_t)chunks * per_chunk > scratch_elems) chunks = (int)(scratch_elems / per_chunk);
    chunks = chunks < 1 ? 1 : chunks;
    int v_chunk = ((n_nodes + chunks - 1) / chunks + 31) / 32 * 32;
    chunks = (n_nodes + v_chunk - 1) / v_chunk;
    dim3 grid((n_cols + 127) / 128, chunks, n_scales);
    ring_reduce_kernel<<<grid, 128, 0, stream>>>(psiT, n_nodes, n_cols, ring_bitmaps, orig_of, hops1,
                                                 n_words, col0, v_chunk, n_scales, scratch);
    HSD_CUDA_TRY(cudaGetLastError());
    const int total = n_cols * n_scales * hops1;
    ring_mean_kernel<<<(total + 255) / 256, 256, 0, stream>>>(emb, scratch, chunks, ring_sizes, n_cols, col0,
                                                              n_scales, hops1);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

extern "C" int hsd_laplacian_spmv(const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                                  const double* x, double* y, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(rowptr && col && x && y && n_nodes > 0, "bad arguments");
    const int64_t threads = (int64_t)n_nodes * 32;
    laplacian_spmv_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, n_nodes, x, y);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
