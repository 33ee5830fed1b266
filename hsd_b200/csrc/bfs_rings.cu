// K1/K2 — k-hop ring extraction fused with the per-ring degree CDF.
//
// Reference loops replaced: tools/hierarchy.py:25-38 (level-synchronous BFS
// with a Python `visited` set, one source at a time) and, for degree-valued
// ring signals, the sort + searchsorted inside scipy.stats.wasserstein_distance
// as called from model/HSD.py:103-112.
//
// Design: one CTA per source; the `seen` set, two ring bitmaps and the ring's
// prefix popcount live in shared memory (4 * N/8 bytes: 50 KB at N = 100k).  Node ids are in
// degree-ascending order, so the ring's degree CDF at support[b] is
// popcount(ring bitmap[0 : bin_end[b])) -> one block prefix-popcount per ring and NO histogram
// atomics (the skewed degree distribution would serialise them).  The same prefix popcount gives
// every ring member its slot in a compact frontier list, which is expanded edge-balanced.
#include <stdlib.h>
#include <algorithm>
#include "hsd_common.cuh"

namespace hsd {

// size (uint32 words) of the global bitmap workspace registered with hsd_bfs_set_workspace
static thread_local long long g_ws_words = 0;
static thread_local uint32_t* g_ws = nullptr;

// test knob: HSD_BFS_FORCE_GLOBAL=1 sends every graph through the global-workspace variant
static bool force_global() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("HSD_BFS_FORCE_GLOBAL"); v = (e && atoi(e)) ? 1 : 0; }
    return v == 1;
}

constexpr int FL_CAP = 2048;                       // frontier nodes expanded per round

struct BfsArgs {
    const int32_t* rowptr;
    const int32_t* col;                            // readable up to the next multiple of 4 entries
    int32_t n_nodes;
    int32_t n_words;
    const int32_t* src_nodes;
    const int32_t* out_rows;
    int32_t n_src;
    int32_t hops;
    const int32_t* bin_end;
    const float* delta;
    int32_t n_bins;
    float* sig;
    int64_t sig_ld;
    // fused all-gather: when set, every signature row is ALSO stored into these peer-mapped
    // copies of the table (same layout), over NVLink, straight from the kernel that produced it
    float* const* sig_peers;
    int32_t n_peers;
    int32_t* ring_sizes;
    uint32_t* ring_bitmaps;
    int32_t empty_as_zero;
    int32_t* status;
    int32_t cta_threads;   // 0 = choose from the graph size
    uint32_t* ws;          // global bitmap workspace (only for graphs too large for shared memory)
};

// `seen` = every node discovered so far (all earlier rings + the part of the current ring found
// so far).  One plain read filters the common case; a new node costs ONE fire-and-forget
// shared-memory atomic.  The ring itself is recovered after the level as seen & ~snapshot, where
// the snapshot of `seen` taken at the start of the level sits in the ring buffer.
__device__ __forceinline__ void visit_neighbor(int u, uint32_t* __restrict__ seen) {
    const uint32_t m = 1u << (u & 31);
    const int w = u >> 5;
    if (!(*((volatile uint32_t*)&seen[w]) & m)) atomicOr(&seen[w], m);
}

// One CTA per source.  Shared memory: the `seen` bitmap, two ring bitmaps (ping-pong), the prefix
// popcount P of the current ring, and a FL_CAP-entry frontier list (CSR start + edge prefix).
//
// Expansion is EDGE-balanced: the frontier ring is compacted (its prefix popcount gives every
// member its slot), the degrees are scanned, and each thread walks an equal, contiguous share
// of the concatenated adjacency lists with 16-byte loads (4 column indices per LDG.128, masked at
// the ends of a node's list).  A node-per-thread / node-per-warp split left most of the CTA
// waiting at the level barrier behind the few threads that drew the hubs (ncu r1: 50 % of all
// stall samples on that barrier); scalar 4-byte loads relied on L1 keeping each 32-byte sector
// alive across 8 iterations and re-fetched it 4x from L2 instead (59 % L1 hit rate).
template <int THREADS>
__device__ __forceinline__ void bfs_one_source(const BfsArgs& p, const int sidx, uint32_t* __restrict__ S,
                                               uint32_t* __restrict__ R0, uint32_t* __restrict__ R1,
                                               uint32_t* __restrict__ P, int* warp_tot, int* fl_start,
                                               int* fl_eo) {
    constexpr int FL_PER_THREAD = FL_CAP / THREADS;
    const int nw = p.n_words;
    const int tid = threadIdx.x;
    const int hops1 = p.hops + 1;
    const int nb1 = p.n_bins - 1;
    const int s = p.src_nodes[sidx];
    const int64_t row = p.out_rows[sidx];

    // ---- hop 0: the ring is the source alone ----
    for (int w = tid; w < nw; w += THREADS) {
        const uint32_t b = (w == (s >> 5)) ? (1u << (s & 31)) : 0u;
        S[w] = b;
        R0[w] = b;
        P[w] = (w > (s >> 5)) ? 1u : 0u;
    }
    if (tid == 0) {
        if (p.ring_sizes) p.ring_sizes[row * hops1] = 1;
        // its W1 term is |deg_i - deg_j|: one scalar instead of a CDF
        if (p.sig) {
            const float d0 = (float)(p.rowptr[s + 1] - p.rowptr[s]);
            p.sig[row * p.sig_ld] = d0;
            for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][row * p.sig_ld] = d0;
        }
    }
    __syncthreads();
    if (p.ring_bitmaps) {
        uint32_t* dst = p.ring_bitmaps + (row * hops1) * (int64_t)nw;
        for (int w = tid; w < nw; w += THREADS) dst[w] = R0[w];
    }

    const int cpt = (nw + THREADS - 1) / THREADS;  // words per thread, prefix-popcount pass
    int n_cur = 1;

    for (int h = 1; h <= p.hops; ++h) {
        uint32_t* F = (h & 1) ? R0 : R1;    // ring h-1 (prefix popcounts in P)
        uint32_t* Fn = (h & 1) ? R1 : R0;   // ring h
        for (int w = tid; w < nw; w += THREADS) Fn[w] = S[w];   // snapshot of `seen` before this level
        __syncthreads();

        for (int r0 = 0; r0 < n_cur; r0 += FL_CAP) {
            const int m = min(FL_CAP, n_cur - r0);
            // ---- compact ring members with rank in [r0, r0 + m) into the frontier list ----
            // RANK-balanced: every thread takes an equal, contiguous run of ranks, finds the word
            // that holds its first rank by binary search on the prefix popcounts and walks the bits
            // from there.  (Word-balanced compaction — round 1 — gave the threads that own the top
            // words of the degree-ordered bitmap, where every ring has all its hubs, up to 32
            // dependent row-pointer loads while the rest had none: 26 % of all stall samples at C3
            // were on the barrier behind them.)  The run's row pointers are loaded back to back and
            // the degrees stay in registers for the scan: one barrier less per round.
            const int cpr = (m + THREADS - 1) / THREADS;          // ranks per thread, <= FL_PER_THREAD
            const int p0 = min(tid * cpr, m), p1 = min(p0 + cpr, m);
            int st[FL_PER_THREAD], loc[FL_PER_THREAD];
            int sum = 0;
            {
                int vs[FL_PER_THREAD];
                if (p0 < p1) {
                    const int r = r0 + p0;
                    int lo = 0, hi = nw;                         // last w with P[w] <= r holds rank r
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if ((int)P[mid] <= r) lo = mid; else hi = mid;
                    }
                    int w = lo;
                    uint32_t bits = F[w];
                    for (int skip = r - (int)P[w]; skip > 0; --skip) bits &= bits - 1;
#pragma unroll
                    for (int q = 0; q < FL_PER_THREAD; ++q) {
                        if (p0 + q < p1) {
                            while (!bits) bits = F[++w];
                            vs[q] = (w << 5) + __ffs(bits) - 1;
                            bits &= bits - 1;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < FL_PER_THREAD; ++q) {        // independent loads: all in flight together
                    st[q] = 0;
                    loc[q] = 0;
                    if (p0 + q < p1) {
                        st[q] = __ldg(p.rowptr + vs[q]);
                        loc[q] = __ldg(p.rowptr + vs[q] + 1);
                    }
                }
#pragma unroll
                for (int q = 0; q < FL_PER_THREAD; ++q) {
                    loc[q] -= st[q];                              // degree
                    sum += loc[q];
                }
            }
            // ---- exclusive scan of the degrees -> edge offsets ----
            int total;
            int run = block_exclusive_scan<THREADS>(sum, warp_tot, &total);   // syncs inside
#pragma unroll
            for (int q = 0; q < FL_PER_THREAD; ++q) {
                if (p0 + q < p1) {
                    fl_start[p0 + q] = st[q];
                    fl_eo[p0 + q] = run;
                }
                run += loc[q];
            }
            if (tid == 0) fl_eo[m] = total;
            __syncthreads();
            // ---- each thread walks an equal contiguous share of the `total` edges ----
            // 16-byte groups of column indices, software-pipelined: the next group of this list (or
            // the first group of the next frontier node's list) is requested before the current
            // group's four bitmap tests, so a thread always has one L2 round trip in flight.
            const int share = (total + THREADS - 1) / THREADS;
            int e = min(tid * share, total);
            const int e_hi = min(e + share, total);
            if (e < e_hi) {
                int lo = 0, hi = m;             // largest k in [0, m) with fl_eo[k] <= e
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (fl_eo[mid] <= e) lo = mid; else hi = mid;
                }
                int k = lo;
                int k_end = min(fl_eo[k + 1], e_hi);
                int i = fl_start[k] + (e - fl_eo[k]);            // CSR index range [i, i_end) of this list
                int i_end = i + (k_end - e);
                int g = i & ~3;
                int4 c = __ldg(reinterpret_cast<const int4*>(p.col + g));
                for (;;) {
                    // what comes after group g: the next group of this list, else the next list
                    int ng = g + 4, ni = i, ni_end = i_end, nk_end = k_end;
                    bool more = true;
                    if (ng >= i_end) {
                        if (k_end < e_hi) {
                            ++k;
                            nk_end = min(fl_eo[k + 1], e_hi);
                            ni = fl_start[k];
                            ni_end = ni + (nk_end - k_end);
                            ng = ni & ~3;
                        } else {
                            more = false;
                        }
                    }
                    int4 cn = c;
                    if (more) cn = __ldg(reinterpret_cast<const int4*>(p.col + ng));
                    if (g >= i && g < i_end) visit_neighbor(c.x, S);
                    if (g + 1 >= i && g + 1 < i_end) visit_neighbor(c.y, S);
                    if (g + 2 >= i && g + 2 < i_end) visit_neighbor(c.z, S);
                    if (g + 3 >= i && g + 3 < i_end) visit_neighbor(c.w, S);
                    if (!more) break;
                    c = cn; g = ng; i = ni; i_end = ni_end; k_end = nk_end;
                }
            }
            __syncthreads();
        }

        // ---- ring h = seen & ~snapshot (into Fn); prefix popcount into P ----
        const int w_lo = min(tid * cpt, nw), w_hi = min(w_lo + cpt, nw);
        int local = 0;
        for (int w = w_lo; w < w_hi; ++w) {
            const uint32_t r = S[w] & ~Fn[w];
            Fn[w] = r;
            local += __popc(r);
        }
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
                    // integer count times integer gap, then one IEEE divide (<= 1.5 ulp total)
                    const float val = __fdiv_rn((float)cnt * __ldg(p.delta + b), n_f);
                    dst[b] = val;
                    for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][dst_off + b] = val;
                }
            } else {
                if (!p.empty_as_zero && tid == 0) atomicOr(p.status, 1);
                // empty ring == point mass at 0 (zero padding of tools/metrics.py:18-36): CDF = 1
                for (int b = tid; b < nb1; b += THREADS) {
                    const float val = p.empty_as_zero ? __ldg(p.delta + b) : 0.f;
                    dst[b] = val;
                    for (int r = 0; r < p.n_peers; ++r) p.sig_peers[r][dst_off + b] = val;
                }
            }
        }
        n_cur = n_ring;
        // no barrier needed here: the next level's first barrier orders these reads of P / Fn
        // before anything overwrites them (Fn of level h+1 is the buffer F of this level)
    }
}

// One CTA per source (bitmaps in shared memory), or — GLOBAL_BM, graphs whose four bitmaps exceed
// shared memory — persistent CTAs that keep their bitmaps in an L2-resident slice of a
// caller-provided global workspace and walk several sources each.
//
// MINB (min CTAs per SM for __launch_bounds__; 0 = unspecified): left to itself ptxas targets full
// occupancy (32 registers at 512 threads, with a 16-byte spill); the global variant runs 2 CTAs per
// SM and gets the 56 registers it wants.  (A 40-register, spill-free build of the 512-thread
// shared-memory variant measured the same 44.9 ms at C3, so that one is left alone.)
template <int THREADS, bool GLOBAL_BM, int MINB = 0>
__global__ void __launch_bounds__(THREADS, MINB)
bfs_ring_signature_kernel(const BfsArgs p) {
    extern __shared__ uint32_t bfs_smem[];
    __shared__ int warp_tot[THREADS / 32];
    __shared__ int fl_start[FL_CAP];
    __shared__ int fl_eo[FL_CAP + 1];
    const int nw = p.n_words;
    if (!GLOBAL_BM) {
        if ((int)blockIdx.x >= p.n_src) return;
        bfs_one_source<THREADS>(p, blockIdx.x, bfs_smem, bfs_smem + nw, bfs_smem + 2 * nw, bfs_smem + 3 * nw,
                                warp_tot, fl_start, fl_eo);
    } else {
        uint32_t* base = p.ws + (size_t)blockIdx.x * 4 * nw;
        for (int sidx = blockIdx.x; sidx < p.n_src; sidx += gridDim.x) {
            bfs_one_source<THREADS>(p, sidx, base, base + nw, base + 2 * nw, base + 3 * nw, warp_tot, fl_start, fl_eo);
            __syncthreads();   // the next source re-initialises the bitmaps
        }
    }
}

static int launch_bfs(const BfsArgs& a_in, cudaStream_t stream) {
    BfsArgs a = a_in;
    if (a.n_src == 0) return HSD_OK;
    const size_t smem = (size_t)4 * a.n_words * sizeof(uint32_t);
    if (smem > 200 * 1024 || force_global()) {
        // bitmaps do not fit shared memory (> ~400k nodes): global workspace, persistent CTAs
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = std::min<long long>(a.n_src, (long long)sms * 2);
        const long long need = (long long)grid * 4 * a.n_words;
        if (!a.ws || need > g_ws_words) {
            set_error("hsd_bfs: %d nodes need a global bitmap workspace of %lld uint32 words "
                      "(hsd_bfs_workspace_words); got %lld", a.n_nodes, need, a.ws ? g_ws_words : 0ll);
            return HSD_ERR_UNSUPPORTED;
        }
        bfs_ring_signature_kernel<512, true, 2><<<grid, 512, 0, stream>>>(a);
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    }
    a.ws = nullptr;
    // large graphs: the bitmaps limit an SM to a few CTAs, so use 512-thread CTAs to keep warps resident;
    // small graphs: per-level fixed costs (barriers, scans) dominate, so use small CTAs and more of them
    static int force = -1;   // tuning knob: HSD_BFS_THREADS in {128, 256, 512, 1024}
    if (force < 0) { const char* e = getenv("HSD_BFS_THREADS"); force = e ? atoi(e) : 0; }
    const int threads = a.cta_threads ? a.cta_threads : (force ? force : (a.n_nodes > 48 * 1024 ? 512 : 256));
    auto launch = [&](auto kern, int t) -> int {
        HSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<a.n_src, t, smem, stream>>>(a);
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    };
    switch (threads) {
        case 128: return launch(bfs_ring_signature_kernel<128, false>, 128);
        case 512: return launch(bfs_ring_signature_kernel<512, false>, 512);
        case 1024: return launch(bfs_ring_signature_kernel<1024, false>, 1024);
        default: return launch(bfs_ring_signature_kernel<256, false>, 256);
    }
}

}  // namespace hsd

static int ring_signature_degree_impl(float* const* sig_peers, int32_t n_peers,
                                      const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                                         const int32_t* src_nodes, const int32_t* out_rows,
                                         int32_t n_src, int32_t hops,
                                         const int32_t* bin_end, const float* delta, int32_t n_bins,
                                         float* sig, int64_t sig_ld, int32_t* ring_sizes,
                                         uint32_t* ring_bitmaps, int32_t empty_as_zero,
                                         int32_t* status, int32_t cta_threads, void* stream) {
    HSD_REQUIRE(rowptr && col && src_nodes && out_rows, "null graph/source pointer");
    HSD_REQUIRE(cta_threads == 0 || cta_threads == 128 || cta_threads == 256 || cta_threads == 512 ||
                    cta_threads == 1024, "cta_threads must be 0 (auto), 128, 256, 512 or 1024");
    HSD_REQUIRE(n_nodes > 0 && n_src >= 0 && hops >= 0, "bad sizes");
    if (sig) {
        HSD_REQUIRE(bin_end && delta && n_bins >= 1 && status, "sig requested without support tables");
        HSD_REQUIRE(sig_ld >= 1 + (int64_t)hops * (n_bins - 1), "sig_ld too small");
    }
    hsd::BfsArgs a;
    a.rowptr = rowptr; a.col = col; a.n_nodes = n_nodes; a.n_words = (n_nodes + 31) / 32;
    a.src_nodes = src_nodes; a.out_rows = out_rows; a.n_src = n_src; a.hops = hops;
    a.bin_end = bin_end; a.delta = delta; a.n_bins = sig ? n_bins : 1;
    a.sig = sig; a.sig_ld = sig_ld; a.sig_peers = sig_peers; a.n_peers = n_peers; a.ring_sizes = ring_sizes; a.ring_bitmaps = ring_bitmaps;
    a.empty_as_zero = empty_as_zero; a.status = status; a.cta_threads = cta_threads; a.ws = hsd::g_ws;
    return hsd::launch_bfs(a, (cudaStream_t)stream);
}

extern "C" int hsd_ring_signature_degree(const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                                         const int32_t* src_nodes, const int32_t* out_rows,
                                         int32_t n_src, int32_t hops,
                                         const int32_t* bin_end, const float* delta, int32_t n_bins,
                                         float* sig, int64_t sig_ld, int32_t* ring_sizes,
                                         uint32_t* ring_bitmaps, int32_t empty_as_zero,
                                         int32_t* status, int32_t cta_threads, void* stream) {
    return ring_signature_degree_impl(nullptr, 0, rowptr, col, n_nodes, src_nodes, out_rows, n_src, hops, bin_end,
                                      delta, n_bins, sig, sig_ld, ring_sizes, ring_bitmaps, empty_as_zero,
                                      status, cta_threads, stream);
}

extern "C" int hsd_ring_signature_degree_allgather(const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                                                   const int32_t* src_nodes, const int32_t* out_rows,
                                                   int32_t n_src, int32_t hops, const int32_t* bin_end,
                                                   const float* delta, int32_t n_bins, float* sig,
                                                   int64_t sig_ld, float* const* sig_peers, int32_t n_peers,
                                                   int32_t* ring_sizes, int32_t empty_as_zero,
                                                   int32_t* status, int32_t cta_threads, void* stream) {
    HSD_REQUIRE(sig && (n_peers == 0 || sig_peers), "fused all-gather needs the local table and the peer pointer array");
    HSD_REQUIRE(n_peers >= 0 && n_peers <= 64, "bad n_peers");
    return ring_signature_degree_impl(sig_peers, n_peers, rowptr, col, n_nodes, src_nodes, out_rows, n_src, hops,
                                      bin_end, delta, n_bins, sig, sig_ld, ring_sizes, nullptr, empty_as_zero,
                                      status, cta_threads, stream);
}

extern "C" int hsd_bfs_rings(const int32_t* rowptr, const int32_t* col, int32_t n_nodes,
                             const int32_t* src_nodes, const int32_t* out_rows, int32_t n_src,
                             int32_t hops, int32_t* ring_sizes,
                             uint32_t* ring_bitmaps, void* stream) {
    return hsd_ring_signature_degree(rowptr, col, n_nodes, src_nodes, out_rows, n_src, hops,
                                     nullptr, nullptr, 1, nullptr, 0, ring_sizes,
                                     ring_bitmaps, 1, nullptr, 0, stream);
}

// Graphs whose four N-bit bitmaps exceed shared memory (> ~400k nodes) need a global workspace.
extern "C" int64_t hsd_bfs_workspace_words(int32_t n_nodes) {
    const long long nw = ((long long)n_nodes + 31) / 32;
    if (4 * nw * (long long)sizeof(uint32_t) <= 200 * 1024 && !hsd::force_global()) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (long long)sms * 2 * 4 * nw;
}

extern "C" int hsd_bfs_set_workspace(uint32_t* workspace, int64_t words) {
    hsd::g_ws = workspace;
    hsd::g_ws_words = workspace ? words : 0;
    return HSD_OK;
}
