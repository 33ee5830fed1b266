// K1/K2, dense variant — k-hop rings of MANY sources by bitmap dynamic programming.
//
// Reference loop replaced: tools/hierarchy.py:25-38 run for every node (tools/hierarchy.py:16-22) and,
// for degree-valued ring signals, the sort + searchsorted inside scipy.stats.wasserstein_distance
// (model/HSD.py:103-112) — the same outputs as bfs_rings.cu, bit for bit.
//
// The frontier-expansion kernel (bfs_rings.cu) walks every adjacency list incident to ball_{H-1}(s)
// one 4-byte column index at a time: ~10 instructions and one random shared-memory bitmap lookup per
// edge visit, level barriers, scans.  When (almost) ALL sources are wanted, the balls obey
//     ball_h(s) = {s}  U  OR_{u in N(s)} ball_{h-1}(u),
// so level h of every source is deg(s) coalesced ORs of N-bit rows of the previous level's table:
// 2E row reads of N/8 bytes per level, streamed with 16-byte loads, no atomics, no barriers, and the
// rows of the hubs (read deg times) stay in L2.  At C3 (100k nodes, 4 hops) that is 3 x 12.5 GB of
// row reads instead of 1.5e10 edge visits.  ring_h = ball_h & ~ball_{h-1}; the degree CDF is the same
// prefix popcount over the degree-ordered ids as in the frontier kernel.
//
// Cost is O(E N / 32) per level whatever the ball sizes, and the two tables take 2 N^2 / 8 bytes
// (2.5 GB at N = 100k), so the host side picks this variant when most sources are requested and the
// workspace fits; few sources or huge sparse graphs stay on the frontier kernel.
#include <algorithm>
#include "hsd_common.cuh"

namespace hsd {

// ball_1: T[s] = {s} U N(s).  One thread per CSR entry / per node; the table was zeroed before.
__global__ void __launch_bounds__(256)
ball1_scatter_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, int64_t nnz,
                     int64_t row_words, uint32_t* __restrict__ T) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) atomicOr(T + idx * row_words + (idx >> 5), 1u << (idx & 31));
    if (idx >= nnz) return;
    int lo = 0, hi = n;                       // row of entry idx: largest s with rowptr[s] <= idx
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) <= idx) lo = mid; else hi = mid;
    }
    const int u = __ldg(col + idx);
    atomicOr(T + (int64_t)lo * row_words + (u >> 5), 1u << (u & 31));
}

// T_next[s][chunk] = T_prev[s][chunk] | OR_{u in N(s)} T_prev[u][chunk]; a CTA owns 256 x 16 bytes of a row.
// srcs == nullptr: all nodes, biggest degree first (ids are degree-ascending) so the hub rows, whose
// CTAs read the most, start early.
__global__ void __launch_bounds__(256, 4)
ball_or_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
               const int32_t* __restrict__ srcs, int64_t row_words, const uint32_t* __restrict__ Tp,
               uint32_t* __restrict__ Tn) {
    const int s = srcs ? __ldg(srcs + blockIdx.x) : n - 1 - (int)blockIdx.x;
    const int64_t w4 = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * 4;    // first word of this thread's uint4
    if (w4 >= row_words) return;
    const uint32_t* base = Tp + w4;
    uint4 acc = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)s * row_words));
    const int e0 = __ldg(rowptr + s), e1 = __ldg(rowptr + s + 1);
    int e = e0;
    for (; e + 8 <= e1; e += 8) {
        int u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) u[q] = __ldg(col + e + q);
        uint4 x[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)u[q] * row_words));
#pragma unroll
        for (int q = 0; q < 8; ++q) { acc.x |= x[q].x; acc.y |= x[q].y; acc.z |= x[q].z; acc.w |= x[q].w; }
    }
    for (; e < e1; ++e) {
        const int u = __ldg(col + e);
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(base + (int64_t)u * row_words));
        acc.x |= x.x; acc.y |= x.y; acc.z |= x.z; acc.w |= x.w;
    }
    *reinterpret_cast<uint4*>(Tn + (int64_t)s * row_words + w4) = acc;
}

struct RingCdfArgs {
    const int32_t* rowptr;
    const int32_t* src_nodes;
    const int32_t* out_rows;
    int32_t n_words, hops, h;                     // h = hop of this launch (1..hops)
    int64_t row_words;
    const uint32_t* cur;                          // ball_h table
    const uint32_t* prev;                         // ball_{h-1} table (h >= 2)
    const int32_t* bin_end;
    const float* delta;
    int32_t n_bins;
    float* sig;
    int64_t sig_ld;
    float* const* sig_peers;
    int32_t n_peers;
    int32_t* ring_sizes;
    uint32_t* ring_bitmaps;
    int32_t empty_as_zero;
    int32_t* status;
};

// Outputs of one (source, hop): ring size, ring bitmap (optional) and the delta-scaled degree CDF of
// the ring whose words sit in shared memory (Fn) — the same arithmetic as the epilogue of
// bfs_ring_signature_kernel (integer prefix popcount at the bin boundaries, one IEEE divide), so both
// variants agree bit for bit.  Hop 0 rides on the hop-1 call.
template <int THREADS>
__device__ __forceinline__ void ring_outputs(const RingCdfArgs& p, const int s, const int64_t row, const int h,
                                             uint32_t* __restrict__ Fn, uint32_t* __restrict__ P, int* warp_tot) {
    const int tid = threadIdx.x, nw = p.n_words, hops1 = p.hops + 1, nb1 = p.n_bins - 1;
    if (h == 1) {
        if (tid == 0) {
            if (p.ring_sizes) p.ring_sizes[row * hops1] = 1;
            if (p.sig) {
                const float d0 = (float)(p.rowptr[s + 1] - p.rowptr[s]);
                p.sig[row * p.sig_ld] = d0;
                for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][row * p.sig_ld] = d0;
            }
        }
        if (p.ring_bitmaps) {
            uint32_t* dst = p.ring_bitmaps + (row * hops1) * (int64_t)nw;
            for (int w = tid; w < nw; w += THREADS) dst[w] = (w == (s >> 5)) ? (1u << (s & 31)) : 0u;
        }
    }
    const int cpt = (nw + THREADS - 1) / THREADS;
    const int w_lo = min(tid * cpt, nw), w_hi = min(w_lo + cpt, nw);
    int local = 0;
    for (int w = w_lo; w < w_hi; ++w) local += __popc(Fn[w]);
    int n_ring;
    int run = block_exclusive_scan<THREADS>(local, warp_tot, &n_ring);
    for (int w = w_lo; w < w_hi; ++w) {
        P[w] = (uint32_t)run;
        run += __popc(Fn[w]);
    }
    __syncthreads();
    if (tid == 0 && p.ring_sizes) p.ring_sizes[row * hops1 + h] = n_ring;
    if (p.ring_bitmaps) {
        uint32_t* dst = p.ring_bitmaps + (row * hops1 + h) * (int64_t)nw;
        for (int w = tid; w < nw; w += THREADS) dst[w] = Fn[w];
    }
    if (p.sig) {
        const int64_t dst_off = row * p.sig_ld + 1 + (int64_t)(h - 1) * nb1;
        float* dst = p.sig + dst_off;
        if (n_ring > 0) {
            const float n_f = (float)n_ring;
            for (int b = tid; b < nb1; b += THREADS) {
                const int e = __ldg(p.bin_end + b);  // < n_nodes for b < n_bins-1
                const int cnt = (int)P[e >> 5] + __popc(Fn[e >> 5] & ((1u << (e & 31)) - 1u));
                const float val = __fdiv_rn((float)cnt * __ldg(p.delta + b), n_f);
                dst[b] = val;
                for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][dst_off + b] = val;
            }
        } else {
            if (!p.empty_as_zero && tid == 0) atomicOr(p.status, 1);
            for (int b = tid; b < nb1; b += THREADS) {
                const float val = p.empty_as_zero ? __ldg(p.delta + b) : 0.f;
                dst[b] = val;
                for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][dst_off + b] = val;
            }
        }
    }
}

// One CTA per source: ring_h = ball_h & ~ball_{h-1} from the two tables (hop 1, and rows too long for the
// fused kernel below).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
ring_cdf_kernel(const RingCdfArgs p) {
    extern __shared__ __align__(16) uint32_t rc_smem[];
    __shared__ int warp_tot[THREADS / 32];
    uint32_t* Fn = rc_smem;                       // row_words words (16-byte granules)
    uint32_t* P = rc_smem + p.row_words;          // n_words words
    const int tid = threadIdx.x;
    const int s = p.src_nodes[blockIdx.x];
    const int64_t row = p.out_rows[blockIdx.x];
    // ring words into shared memory with coalesced 16-byte loads (rows are 16-byte aligned and padded
    // with zero bits)
    const uint4* cur4 = reinterpret_cast<const uint4*>(p.cur + (int64_t)s * p.row_words);
    const uint4* prv4 = p.prev ? reinterpret_cast<const uint4*>(p.prev + (int64_t)s * p.row_words) : nullptr;
    const int n4 = (int)(p.row_words / 4);
    for (int q = tid; q < n4; q += THREADS) {
        uint4 c = __ldg(cur4 + q);
        if (prv4) {
            const uint4 b = __ldg(prv4 + q);
            c.x &= ~b.x; c.y &= ~b.y; c.z &= ~b.z; c.w &= ~b.w;
        }
        *reinterpret_cast<uint4*>(Fn + 4 * q) = c;
    }
    __syncthreads();
    if (!prv4 && tid == 0) Fn[s >> 5] &= ~(1u << (s & 31));     // hop 1: ball_1 minus the source itself
    __syncthreads();
    ring_outputs<THREADS>(p, s, row, p.h, Fn, P, warp_tot);
}

// Levels h >= 2 for rows of at most CHUNKS x 1024 words, FUSED: one CTA owns the whole row of node s,
// ORs the neighbours' previous-level rows into registers (CHUNKS uint4 per thread), writes ball_h(s) and —
// the new and the old row being in registers already — emits ring_h = new & ~old and its CDF in place,
// instead of a second pass that re-reads both rows from the tables.  row_of_node (all-nodes launches):
// output row of node s, -1 if s is not a requested source.
template <int CHUNKS>
__global__ void __launch_bounds__(256, CHUNKS >= 4 ? 3 : 4)
ball_or_cdf_kernel(const int32_t* __restrict__ col, int n, const int32_t* __restrict__ srcs,
                   const int32_t* __restrict__ row_of_node, const uint32_t* __restrict__ Tp,
                   uint32_t* __restrict__ Tn, const RingCdfArgs p) {
    extern __shared__ __align__(16) uint32_t rc_smem[];
    __shared__ int warp_tot[256 / 32];
    const int tid = threadIdx.x;
    const int s = srcs ? __ldg(srcs + blockIdx.x) : n - 1 - (int)blockIdx.x;
    const int64_t rw = p.row_words;
    const int n4 = (int)(rw / 4);
    uint4 old[CHUNKS], acc[CHUNKS];
    const uint4* own = reinterpret_cast<const uint4*>(Tp + (int64_t)s * rw);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int q = c * 256 + tid;
        old[c] = (q < n4) ? __ldg(own + q) : make_uint4(0u, 0u, 0u, 0u);
        acc[c] = old[c];
    }
    constexpr int NB = (CHUNKS >= 4) ? 2 : (CHUNKS == 2 ? 4 : 8);     // neighbours in flight: 8 uint4 loads per thread
    const int e0 = __ldg(p.rowptr + s), e1 = __ldg(p.rowptr + s + 1);
    int e = e0;
    for (; e + NB <= e1; e += NB) {
        const uint4* nbr[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) nbr[b] = reinterpret_cast<const uint4*>(Tp + (int64_t)__ldg(col + e + b) * rw);
        uint4 x[NB][CHUNKS];
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int q = c * 256 + tid;
                x[b][c] = (q < n4) ? __ldg(nbr[b] + q) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                acc[c].x |= x[b][c].x; acc[c].y |= x[b][c].y; acc[c].z |= x[b][c].z; acc[c].w |= x[b][c].w;
            }
    }
    for (; e < e1; ++e) {
        const uint4* nb = reinterpret_cast<const uint4*>(Tp + (int64_t)__ldg(col + e) * rw);
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
            const int q = c * 256 + tid;
            if (q < n4) {
                const uint4 x = __ldg(nb + q);
                acc[c].x |= x.x; acc[c].y |= x.y; acc[c].z |= x.z; acc[c].w |= x.w;
            }
        }
    }
    uint4* dst = reinterpret_cast<uint4*>(Tn + (int64_t)s * rw);
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int q = c * 256 + tid;
        if (q < n4) dst[q] = acc[c];
    }
    const int64_t row = row_of_node ? (int64_t)__ldg(row_of_node + s) : (int64_t)__ldg(p.out_rows + blockIdx.x);
    if (row < 0) return;                          // not a requested source (uniform over the CTA)
    uint32_t* Fn = rc_smem;
    uint32_t* P = rc_smem + rw;
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
        const int q = c * 256 + tid;
        if (q < n4)
            *reinterpret_cast<uint4*>(Fn + 4 * q) =
                make_uint4(acc[c].x & ~old[c].x, acc[c].y & ~old[c].y, acc[c].z & ~old[c].z, acc[c].w & ~old[c].w);
    }
    __syncthreads();
    ring_outputs<256>(p, s, row, p.h, Fn, P, warp_tot);
}

// row_of_node[s] = output row of source s (the table is pre-filled with -1)
__global__ void row_of_node_kernel(const int32_t* __restrict__ srcs, const int32_t* __restrict__ out_rows, int n_src,
                                   int32_t* __restrict__ row_of_node) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_src) row_of_node[srcs[i]] = out_rows[i];
}

// ---- column-split dense variant (several GPUs) -----------------------------------------------------
// The OR recursion is independent per bitmap word, so rank r of W can run EVERY level for ALL nodes on
// its own range of bitmap words [w_begin, w_begin + lw) with no exchange at all (tables of N x lw words,
// 1/W of the work and of the memory).  What a rank cannot do alone is the prefix popcount of a ring over
// the whole row; it emits the PARTIAL integer counts of its word range,
//   counts[s][(h-1)*(B-1) + b] = #{ j in ring_h(s) : w_begin*32 <= j < min(bin_end[b], (w_begin+lw)*32) }
//   counts[s][hops*(B-1) + h-1] = #{ j in ring_h(s) in the word range }              (partial ring size)
// the host sums them over the ranks (one integer all-reduce: exact, order-free) and
// signature_from_counts_kernel applies the same float arithmetic as the other variants to the sums, so
// the signatures are bit-identical to the single-GPU ones.
__global__ void __launch_bounds__(256)
ball1_scatter_cols_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, int64_t nnz,
                          int lw, int w_begin, uint32_t* __restrict__ T) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) {
        const int w = (int)(idx >> 5) - w_begin;
        if (w >= 0 && w < lw) atomicOr(T + idx * lw + w, 1u << (idx & 31));
    }
    if (idx >= nnz) return;
    const int u = __ldg(col + idx);
    const int w = (u >> 5) - w_begin;
    if (w < 0 || w >= lw) return;
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(rowptr + mid) <= idx) lo = mid; else hi = mid;
    }
    atomicOr(T + (int64_t)lo * lw + w, 1u << (u & 31));
}

struct RingCountArgs {
    int32_t n_nodes, lw, w_begin, hops, h, n_bins;
    const uint32_t* cur;
    const uint32_t* prev;      // nullptr at h = 1 (the ring is ball_1 minus the node itself)
    const int32_t* bin_end;
    int32_t* counts;
    int64_t ld_c;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
ring_count_cols_kernel(const RingCountArgs p) {
    extern __shared__ __align__(16) uint32_t rc_smem[];
    __shared__ int warp_tot[THREADS / 32];
    uint32_t* Fn = rc_smem;
    uint32_t* P = rc_smem + p.lw;
    const int tid = threadIdx.x, lw = p.lw, nb1 = p.n_bins - 1;
    const int s = blockIdx.x;
    const uint4* cur4 = reinterpret_cast<const uint4*>(p.cur + (int64_t)s * lw);
    const uint4* prv4 = p.prev ? reinterpret_cast<const uint4*>(p.prev + (int64_t)s * lw) : nullptr;
    for (int q = tid; q < lw / 4; q += THREADS) {
        uint4 c = __ldg(cur4 + q);
        if (prv4) {
            const uint4 b = __ldg(prv4 + q);
            c.x &= ~b.x; c.y &= ~b.y; c.z &= ~b.z; c.w &= ~b.w;
        }
        *reinterpret_cast<uint4*>(Fn + 4 * q) = c;
    }
    __syncthreads();
    if (!prv4 && tid == 0) {
        const int w = (s >> 5) - p.w_begin;
        if (w >= 0 && w < lw) Fn[w] &= ~(1u << (s & 31));
    }
    __syncthreads();
    const int cpt = (lw + THREADS - 1) / THREADS;
    const int w_lo = min(tid * cpt, lw), w_hi = min(w_lo + cpt, lw);
    int local = 0;
    for (int w = w_lo; w < w_hi; ++w) local += __popc(Fn[w]);
    int n_part;
    int run = block_exclusive_scan<THREADS>(local, warp_tot, &n_part);
    for (int w = w_lo; w < w_hi; ++w) {
        P[w] = (uint32_t)run;
        run += __popc(Fn[w]);
    }
    __syncthreads();
    int32_t* dst = p.counts + (int64_t)s * p.ld_c;
    if (tid == 0) dst[p.hops * nb1 + (p.h - 1)] = n_part;
    const int bit_lo = p.w_begin * 32, bit_hi = bit_lo + lw * 32;
    for (int b = tid; b < nb1; b += THREADS) {
        const int e = min(max(__ldg(p.bin_end + b), bit_lo), bit_hi) - bit_lo;     // bits of the local range below bin_end[b]
        int cnt = n_part;
        if (e < lw * 32) cnt = (int)P[e >> 5] + __popc(Fn[e >> 5] & ((1u << (e & 31)) - 1u));
        dst[(p.h - 1) * nb1 + b] = cnt;
    }
}

// counts (summed over the word ranges) -> signature rows; one warp per (source, hop) would do, a CTA per
// source keeps it simple: same arithmetic as ring_outputs.
__global__ void __launch_bounds__(128)
signature_from_counts_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ counts, int64_t ld_c,
                             const int32_t* __restrict__ src_nodes, const int32_t* __restrict__ out_rows,
                             int hops, const float* __restrict__ delta, int n_bins, float* __restrict__ sig,
                             int64_t sig_ld, int32_t* __restrict__ ring_sizes, int empty_as_zero,
                             int32_t* __restrict__ status) {
    const int s = src_nodes[blockIdx.x];
    const int64_t row = out_rows[blockIdx.x];
    const int nb1 = n_bins - 1, hops1 = hops + 1, tid = threadIdx.x;
    const int32_t* c = counts + (int64_t)s * ld_c;
    if (tid == 0) {
        if (sig) sig[row * sig_ld] = (float)(rowptr[s + 1] - rowptr[s]);
        if (ring_sizes) ring_sizes[row * hops1] = 1;
    }
    for (int h = 1; h <= hops; ++h) {
        const int n_ring = c[hops * nb1 + (h - 1)];
        if (tid == 0 && ring_sizes) ring_sizes[row * hops1 + h] = n_ring;
        if (!sig) continue;
        float* dst = sig + row * sig_ld + 1 + (int64_t)(h - 1) * nb1;
        if (n_ring > 0) {
            const float n_f = (float)n_ring;
            for (int b = tid; b < nb1; b += 128)
                dst[b] = __fdiv_rn((float)c[(h - 1) * nb1 + b] * __ldg(delta + b), n_f);
        } else {
            if (!empty_as_zero && tid == 0) atomicOr(status, 1);
            for (int b = tid; b < nb1; b += 128) dst[b] = empty_as_zero ? __ldg(delta + b) : 0.f;
        }
    }
}

static inline int64_t dense_row_words(int32_t n_nodes) { return ((int64_t)(n_nodes + 31) / 32 + 3) / 4 * 4; }

}  // namespace hsd

extern "C" int64_t hsd_ring_dense_workspace_words(int32_t n_nodes) {
    // two N x N-bit tables + the node -> output row map of the fused level kernel
    return 2 * (int64_t)n_nodes * hsd::dense_row_words(n_nodes) + ((int64_t)n_nodes + 3) / 4 * 4;
}

extern "C" int hsd_ring_signature_degree_dense(const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                                               const int32_t* src_nodes, const int32_t* out_rows, int32_t n_src,
                                               int32_t hops, const int32_t* bin_end, const float* delta,
                                               int32_t n_bins, float* sig, int64_t sig_ld,
                                               float* const* sig_peers, int32_t n_peers, int32_t* ring_sizes,
                                               uint32_t* ring_bitmaps, int32_t empty_as_zero, int32_t* status,
                                               uint32_t* workspace, int64_t workspace_words, int64_t nnz,
                                               void* stream_) {
    using namespace hsd;
    cudaStream_t stream = (cudaStream_t)stream_;
    HSD_REQUIRE(rowptr && col && src_nodes && out_rows && workspace, "null pointer");
    HSD_REQUIRE(n_nodes > 0 && n_src >= 0 && hops >= 0 && nnz >= 0, "bad sizes");
    HSD_REQUIRE(hops >= 1, "the dense variant needs hops >= 1 (hop 0 alone: use hsd_ring_signature_degree)");
    HSD_REQUIRE(n_peers >= 0 && n_peers <= 64 && (n_peers == 0 || (sig && sig_peers)), "bad peer list");
    if (sig) {
        HSD_REQUIRE(bin_end && delta && n_bins >= 1 && status, "sig requested without support tables");
        HSD_REQUIRE(sig_ld >= 1 + (int64_t)hops * (n_bins - 1), "sig_ld too small");
    }
    const int64_t rw = dense_row_words(n_nodes);
    HSD_REQUIRE(workspace_words >= hsd_ring_dense_workspace_words(n_nodes), "workspace smaller than hsd_ring_dense_workspace_words");
    HSD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    if (n_src == 0) return HSD_OK;
    const int nw = (n_nodes + 31) / 32;
    const size_t smem = (size_t)(rw + nw) * sizeof(uint32_t);
    HSD_REQUIRE(smem <= 200 * 1024, "graph too large for the dense variant's shared-memory CDF pass");
    uint32_t* T[2] = {workspace, workspace + (int64_t)n_nodes * rw};
    RingCdfArgs a;
    a.rowptr = rowptr; a.src_nodes = src_nodes; a.out_rows = out_rows; a.n_words = nw; a.hops = hops;
    a.row_words = rw; a.bin_end = bin_end; a.delta = delta; a.n_bins = sig ? n_bins : 1; a.sig = sig; a.sig_ld = sig_ld;
    a.sig_peers = sig_peers; a.n_peers = n_peers; a.ring_sizes = ring_sizes; a.ring_bitmaps = ring_bitmaps;
    a.empty_as_zero = empty_as_zero; a.status = status;
    HSD_CUDA_TRY(cudaFuncSetAttribute(ring_cdf_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HSD_CUDA_TRY(cudaFuncSetAttribute(ring_cdf_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // short rows (<= 1024 words): the CDF pass is bound by per-CTA fixed instruction work (block scan, bin loop:
    // 80 % issue-active at C2), which 128-thread CTAs halve
    auto launch_cdf = [&]() -> int {
        if (nw <= 1024) ring_cdf_kernel<128><<<n_src, 128, smem, stream>>>(a);
        else ring_cdf_kernel<256><<<n_src, 256, smem, stream>>>(a);
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    };
    // ---- ball_1 ----
    HSD_CUDA_TRY(cudaMemsetAsync(T[0], 0, (size_t)n_nodes * rw * sizeof(uint32_t), stream));
    {
        const int64_t items = std::max<int64_t>(nnz, n_nodes);
        ball1_scatter_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(rowptr, col, n_nodes, nnz, rw, T[0]);
        HSD_CUDA_TRY(cudaGetLastError());
    }
    a.h = 1; a.cur = T[0]; a.prev = nullptr;
    { const int rc = launch_cdf(); if (rc != HSD_OK) return rc; }
    const unsigned chunks = (unsigned)((rw / 4 + 255) / 256);
    // rows of 1025..4096 words (32k < N <= 131k nodes): the OR pass and the CDF pass of a level are one kernel,
    // which saves re-reading both rows from HBM (C3: 8.7 -> 7.7 ms).  Shorter rows are L2-resident and measured
    // faster as two lean kernels (C2: 0.36 vs 0.40 ms fused); longer rows do not fit the register budget.
    const bool fused = chunks >= 2 && chunks <= 4;
    int32_t* row_of_node = reinterpret_cast<int32_t*>(workspace + 2 * (int64_t)n_nodes * rw);
    if (fused && hops > 2) {
        HSD_CUDA_TRY(cudaMemsetAsync(row_of_node, 0xff, (size_t)n_nodes * sizeof(int32_t), stream));
        row_of_node_kernel<<<(n_src + 255) / 256, 256, 0, stream>>>(src_nodes, out_rows, n_src, row_of_node);
        HSD_CUDA_TRY(cudaGetLastError());
    }
    auto launch_fused = [&](int grid, const int32_t* srcs, const int32_t* rmap, const uint32_t* Tp, uint32_t* Tn) -> int {
        if (chunks == 2) {
            HSD_CUDA_TRY(cudaFuncSetAttribute(ball_or_cdf_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ball_or_cdf_kernel<2><<<grid, 256, smem, stream>>>(col, n_nodes, srcs, rmap, Tp, Tn, a);
        } else {
            HSD_CUDA_TRY(cudaFuncSetAttribute(ball_or_cdf_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ball_or_cdf_kernel<4><<<grid, 256, smem, stream>>>(col, n_nodes, srcs, rmap, Tp, Tn, a);
        }
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    };
    for (int h = 2; h <= hops; ++h) {
        const uint32_t* Tp = T[h & 1];        // h = 2 reads T[0]
        uint32_t* Tn = T[(h + 1) & 1];
        a.h = h; a.cur = Tn; a.prev = Tp;
        if (fused) {
            const int rc = (h < hops) ? launch_fused(n_nodes, nullptr, row_of_node, Tp, Tn)
                                      : launch_fused(n_src, src_nodes, nullptr, Tp, Tn);
            if (rc != HSD_OK) return rc;
            continue;
        }
        // a thread owns one 16-byte piece of the row: rows shorter than 1024 words get narrower CTAs (no idle warps)
        const int or_threads = chunks == 1 ? (int)std::min<int64_t>(256, (rw / 4 + 31) / 32 * 32) : 256;
        if (h < hops) {
            ball_or_kernel<<<dim3(n_nodes, chunks), or_threads, 0, stream>>>(rowptr, col, n_nodes, nullptr, rw, Tp, Tn);
        } else {
            ball_or_kernel<<<dim3(n_src, chunks), or_threads, 0, stream>>>(rowptr, col, n_nodes, src_nodes, rw, Tp, Tn);
        }
        HSD_CUDA_TRY(cudaGetLastError());
        { const int rc = launch_cdf(); if (rc != HSD_OK) return rc; }
    }
    return HSD_OK;
}

// ---- column-split dense variant: entry points -------------------------------------------------------
extern "C" int64_t hsd_ring_cols_workspace_words(int32_t n_nodes, int32_t word4_begin, int32_t word4_end) {
    return 2 * (int64_t)n_nodes * 4 * (int64_t)(word4_end - word4_begin);
}

extern "C" int hsd_ring_counts_dense_cols(const int32_t* rowptr, const int32_t* col, int32_t n_nodes, int64_t nnz,
                                          int32_t hops, const int32_t* bin_end, int32_t n_bins,
                                          int32_t word4_begin, int32_t word4_end, int32_t* counts, int64_t ld_c,
                                          uint32_t* workspace, int64_t workspace_words, void* stream_) {
    using namespace hsd;
    cudaStream_t stream = (cudaStream_t)stream_;
    HSD_REQUIRE(rowptr && col && bin_end && counts && workspace, "null pointer");
    HSD_REQUIRE(n_nodes > 0 && hops >= 1 && n_bins >= 1 && nnz >= 0, "bad sizes");
    const int64_t rw = dense_row_words(n_nodes);
    HSD_REQUIRE(word4_begin >= 0 && word4_begin <= word4_end && (int64_t)word4_end * 4 <= rw, "bad word range");
    HSD_REQUIRE(ld_c >= (int64_t)hops * (n_bins - 1) + hops, "ld_c too small");
    HSD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    const int lw = 4 * (word4_end - word4_begin);
    HSD_REQUIRE(workspace_words >= 2 * (int64_t)n_nodes * lw, "workspace smaller than hsd_ring_cols_workspace_words");
    if (lw == 0) {      // a rank without columns (more ranks than 16-byte pieces): all counts are zero
        HSD_CUDA_TRY(cudaMemsetAsync(counts, 0, (size_t)n_nodes * ld_c * sizeof(int32_t), stream));
        return HSD_OK;
    }
    const size_t smem = (size_t)2 * lw * sizeof(uint32_t);
    HSD_REQUIRE(smem <= 200 * 1024, "word range too wide for the shared-memory count pass");
    uint32_t* T[2] = {workspace, workspace + (int64_t)n_nodes * lw};
    HSD_CUDA_TRY(cudaFuncSetAttribute(ring_count_cols_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HSD_CUDA_TRY(cudaMemsetAsync(T[0], 0, (size_t)n_nodes * lw * sizeof(uint32_t), stream));
    const int64_t items = std::max<int64_t>(nnz, n_nodes);
    ball1_scatter_cols_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(rowptr, col, n_nodes, nnz, lw,
                                                                                 4 * word4_begin, T[0]);
    HSD_CUDA_TRY(cudaGetLastError());
    RingCountArgs a;
    a.n_nodes = n_nodes; a.lw = lw; a.w_begin = 4 * word4_begin; a.hops = hops; a.n_bins = n_bins;
    a.bin_end = bin_end; a.counts = counts; a.ld_c = ld_c;
    a.h = 1; a.cur = T[0]; a.prev = nullptr;
    ring_count_cols_kernel<128><<<n_nodes, 128, smem, stream>>>(a);
    HSD_CUDA_TRY(cudaGetLastError());
    const unsigned chunks = (unsigned)((lw / 4 + 255) / 256);
    const int or_threads = chunks == 1 ? (int)std::min<int64_t>(256, (lw / 4 + 31) / 32 * 32) : 256;
    for (int h = 2; h <= hops; ++h) {
        const uint32_t* Tp = T[h & 1];
        uint32_t* Tn = T[(h + 1) & 1];
        ball_or_kernel<<<dim3(n_nodes, chunks), or_threads, 0, stream>>>(rowptr, col, n_nodes, nullptr, lw, Tp, Tn);
        HSD_CUDA_TRY(cudaGetLastError());
        a.h = h; a.cur = Tn; a.prev = Tp;
        ring_count_cols_kernel<128><<<n_nodes, 128, smem, stream>>>(a);
        HSD_CUDA_TRY(cudaGetLastError());
    }
    return HSD_OK;
}

extern "C" int hsd_ring_signature_from_counts(const int32_t* rowptr, const int32_t* counts, int64_t ld_c,
                                              const int32_t* src_nodes, const int32_t* out_rows, int32_t n_src,
                                              int32_t hops, const float* delta, int32_t n_bins, float* sig,
                                              int64_t sig_ld, int32_t* ring_sizes, int32_t empty_as_zero,
                                              int32_t* status, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(rowptr && counts && src_nodes && out_rows && status, "null pointer");
    HSD_REQUIRE(n_src >= 0 && hops >= 1 && n_bins >= 1, "bad sizes");
    HSD_REQUIRE(ld_c >= (int64_t)hops * (n_bins - 1) + hops, "ld_c too small");
    if (sig) {
        HSD_REQUIRE(delta, "sig requested without the support gaps");
        HSD_REQUIRE(sig_ld >= 1 + (int64_t)hops * (n_bins - 1), "sig_ld too small");
    }
    if (n_src == 0) return HSD_OK;
    signature_from_counts_kernel<<<n_src, 128, 0, (cudaStream_t)stream>>>(rowptr, counts, ld_c, src_nodes, out_rows,
                                                                        hops, delta, n_bins, sig, sig_ld, ring_sizes,
                                                                        empty_as_zero, status);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}
