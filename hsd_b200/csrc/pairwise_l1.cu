// K3 — pairwise L1 between signature columns: D[i][j] = sum_k |S[k][i] - S[k][j]|.
//
// Reference loop replaced: model/HSD.py:103-112 — N(N-1)/2 * (H+1) calls of
// scipy.stats.wasserstein_distance.  With every ring signal living on one
// shared support, W1 is the L1 distance between (delta-scaled) CDF vectors, so
// the whole O(N^2 * H * B) core is this one kernel.
//
// Shape: SGEMM-like 128x128 output tile per CTA, K streamed in chunks of 16
// signature rows through a 4-stage TMA (cp.async.bulk.tensor.2d) + mbarrier
// ring (thread 0 issues two chunks ahead); 8 warps hold 8x8 register tiles and
// run   d = a - b ; acc += |d|   (FADD + FADD with |.| source modifier) on the
// FP32 CUDA cores.  |a-b| is not a contraction, so tensor cores do not apply.
// The table is K-major ([k][node]) so a TMA box {128 nodes, 16 k} lands in
// shared memory already in the layout the register tiles read with
// conflict-free LDS.128.
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>
#include <cudaTypedefs.h>
#include "hsd_common.cuh"

namespace hsd {

constexpr int TILE = HSD_PAIR_TILE;
constexpr int KC = HSD_PAIR_KCHUNK;
constexpr int STAGES = 4;
constexpr int LOOKAHEAD = 2;   // chunks in flight ahead of the one being consumed
constexpr int PAIR_THREADS = 256;
#ifndef HSD_PAIR_LANE_MAP
#define HSD_PAIR_LANE_MAP 0
#endif
#ifndef HSD_PAIR_V2_DEFAULT
#define HSD_PAIR_V2_DEFAULT 32, 3, 1, 0, 2, 8   // kc (0 = round-1 kernel), stages, packed, (unused), lookahead, unroll
#endif
constexpr uint32_t STAGE_BYTES = 2u * KC * TILE * sizeof(float);

struct __align__(128) PairSmem {
    float a[STAGES][KC][TILE];
    float b[STAGES][KC][TILE];
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

// linear index over the tiles (I, J >= I) of a tiles_r x tiles_c trapezoid (row-major);
// tile row i starts at first(i) = i*tiles_c - i*(i-1)/2
__device__ __forceinline__ void tri_decode(int t, int tiles_r, int tiles_c, int& I, int& J) {
    const float b = 2.f * tiles_c + 1.f;
    int i = (int)((b - sqrtf(fmaxf(b * b - 8.f * (float)t, 0.f))) * 0.5f);
    i = max(0, min(i, tiles_r - 1));
    while (i > 0 && i * tiles_c - (i * (i - 1)) / 2 > t) --i;
    while (i + 1 < tiles_r && (i + 1) * tiles_c - ((i + 1) * i) / 2 <= t) ++i;
    I = i;
    J = i + (t - (i * tiles_c - (i * (i - 1)) / 2));
}

struct PairArgs {
    int k_chunks;
    int row0, n_rows, col0, n_cols;
    int tiles_r, tiles_c;
    int symmetric;
    float* out;
    int64_t ld;
    int vec_ok;  // float4 stores legal (ld % 4 == 0, base and col0 aligned)
    // sharded symmetric mode: the logical N x N result lives as `world` row blocks of `per`
    // rows, one per GPU, shard_ptrs[r] = base of block r (peer-mapped for r != this rank);
    // this launch computes tiles tile_offset, tile_offset + tile_stride, ...
    float* const* shard_ptrs;
    int per;
    int tile_stride, tile_offset;
    // v2 (persistent) kernel: tiles of this launch, valid rows of the last K chunk rounded up to 4,
    // chunks the elected producer thread runs ahead when there is no producer warp
    int n_tiles, k_last, lookahead;
    // explicit tile list (sharded runs): entry t = {first row, first column, mirror flag}; nullptr: tiles are
    // enumerated from the launch rectangle / triangle
    const int32_t* tile_list;
    int list_tile_n;              // 128 or 64: column width of the listed tiles
    unsigned int* tile_counter;   // dynamic tile scheduler: zeroed before the launch
};

// Thread -> (tx, ty) position in the 16 x 16 thread grid of a tile (thread (tx, ty) owns columns tx*4.. and
// rows ty*4.. of each 64-wide half).  The map decides which lanes of a warp sit side by side, i.e. how long the
// contiguous runs of a warp's result stores are — which matters when the stores cross NVLink (peer-mapped blocks):
//   0: warp = 16 tx x 2 ty  -> direct stores 2 rows x 256 B, mirrored stores 16 rows x 32 B   (round 1)
//   1: warp =  4 tx x 8 ty  -> direct 8 rows x 64 B,  mirrored 4 rows x 128 B
//   2: warp =  8 tx x 4 ty  -> direct 4 rows x 128 B, mirrored 8 rows x 64 B
// Operand reads stay conflict-free broadcasts in every map (a warp reads 8/4/2 distinct 16-byte pieces of A
// and 4/8/16 of B per LDS.128).
__device__ __forceinline__ void tile_thread_pos(int tid, int& tx, int& ty) {
    constexpr int lane_map = HSD_PAIR_LANE_MAP;      // compile-time: a run-time choice costs the v2 kernel registers (spills)
    const int lane = tid & 31, w = tid >> 5;
    if (lane_map == 1) {
        tx = (w & 3) * 4 + (lane & 3);
        ty = (w >> 2) * 8 + (lane >> 2);
    } else if (lane_map == 2) {
        tx = (w & 1) * 8 + (lane & 7);
        ty = (w >> 1) * 4 + (lane >> 3);
    } else {
        tx = tid & 15;
        ty = tid >> 4;
    }
}

// pointer to logical element (i, 0)
__device__ __forceinline__ float* row_ptr(const PairArgs& p, int i) {
    if (p.shard_ptrs) {
        const int o = i / p.per;
        return p.shard_ptrs[o] + (int64_t)(i - o * p.per) * p.ld;
    }
    return p.out + (int64_t)(i - p.row0) * p.ld - p.col0;
}

// Row pointers of the 128 consecutive rows [base, base + TILE) without a division per row: in the
// sharded layout the owner changes at most once inside a tile when per >= TILE.
struct TileRows {
    float* p0;
    float* p1;
    int base, split;
    int64_t ld;
    __device__ __forceinline__ float* row(int i) const {
        return i < split ? p0 + (int64_t)(i - base) * ld : p1 + (int64_t)(i - split) * ld;
    }
};
__device__ __forceinline__ TileRows tile_rows(const PairArgs& p, int base) {
    TileRows t;
    t.base = base;
    t.ld = p.ld;
    if (p.shard_ptrs) {
        const int o = base / p.per;
        t.split = (o + 1) * p.per;
        t.p0 = p.shard_ptrs[o] + (int64_t)(base - o * p.per) * p.ld;
        t.p1 = (t.split < base + TILE) ? p.shard_ptrs[o + 1] : t.p0;
    } else {
        t.split = 0x7fffffff;
        t.p0 = t.p1 = p.out + (int64_t)(base - p.row0) * p.ld - p.col0;
    }
    return t;
}

#ifndef HSD_PAIR_MINB
#define HSD_PAIR_MINB 2
#endif
template <int UNROLL, int KLAST>
__global__ void __launch_bounds__(PAIR_THREADS, HSD_PAIR_MINB)
pairwise_l1_kernel(const __grid_constant__ CUtensorMap tmap, const PairArgs p) {
    extern __shared__ __align__(128) unsigned char pair_smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(pair_smem_raw);

    int I, J;
    if (p.symmetric) {
        tri_decode(blockIdx.x * p.tile_stride + p.tile_offset, p.tiles_r, p.tiles_c, I, J);
    } else {
        I = blockIdx.x / p.tiles_c;
        J = blockIdx.x - I * p.tiles_c;
    }
    const int i_base = p.row0 + I * TILE;  // node index of tile row 0
    const int j_base = p.col0 + J * TILE;

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), PAIR_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // Producer duty rides on thread 0: before consuming chunk c it issues the TMA
    // loads of chunk c+LOOKAHEAD into the stage chunk c+LOOKAHEAD-STAGES used, so
    // it only ever waits on a stage every warp released two chunks ago.
    auto issue_chunk = [&](int n) {
        const int s = n % STAGES;
        const uint32_t ph = (n / STAGES) & 1;
        mbar_wait(smem_u32(&sm.empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&sm.full[s]);
        mbar_expect_tx(full, STAGE_BYTES);
        tma_load_2d(smem_u32(&sm.a[s][0][0]), &tmap, i_base, n * KC, full);
        tma_load_2d(smem_u32(&sm.b[s][0][0]), &tmap, j_base, n * KC, full);
    };
    if (tid == 0)
        for (int n = 0; n < LOOKAHEAD && n < p.k_chunks; ++n) issue_chunk(n);

    // ===== 16 x 16 threads, each 8 x 8 outputs (2 x 2 blocks of 4 x 4) =====
    int tx, ty;
    tile_thread_pos(tid, tx, ty);
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    uint32_t ready = 0;   // phase of the next chunk already observed complete
    // KLAST > 0: the last chunk holds only KLAST signature rows (the rest is zero padding) and is
    // peeled off as straight-line code; KLAST == 0: every chunk is full
    const int full_chunks = KLAST ? p.k_chunks - 1 : p.k_chunks;
    for (int c = 0; c < full_chunks; ++c) {
        if (tid == 0 && c + LOOKAHEAD < p.k_chunks) issue_chunk(c + LOOKAHEAD);
        const int s = c % STAGES;
        const uint32_t ph = (c / STAGES) & 1;
        if (!ready) mbar_wait(smem_u32(&sm.full[s]), ph);
        // probe the NEXT chunk's barrier now (its TMA was issued a chunk ago): the try_wait's
        // ~100-cycle latency then hides under this chunk's 2048 FADDs instead of heading the
        // next chunk's dependency chain
        ready = (c + 1 < p.k_chunks)
                    ? mbar_test(smem_u32(&sm.full[(c + 1) % STAGES]), ((c + 1) / STAGES) & 1) : 0u;
        // Partial unroll on purpose: a fully unrolled 16-step chunk is ~34 KB of SASS, larger
        // than the 32 KB instruction cache, and the kernel then stalls on instruction fetch
        // (ncu: stall_no_instruction 0.82 -> 0.09 per issue, profiles/r1_pairwise_notes.md).
#pragma unroll UNROLL
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[s][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[s][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[s][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[s][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            // 64 subtracts, then 64 |.|-accumulates: issuing the subtracts as a block lets ptxas
            // place d[][] and acc[][] in different register banks
            float d[8][8];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) d[r][q] = av[r] - bv[q];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[r][q] += fabsf(d[r][q]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(&sm.empty[s]));
    }
    if (KLAST) {
        const int c = full_chunks;
        const int s = c % STAGES;
        if (!ready) mbar_wait(smem_u32(&sm.full[s]), (c / STAGES) & 1);
#pragma unroll
        for (int kk = 0; kk < KLAST; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[s][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[s][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[s][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sm.b[s][kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[r][q] += fabsf(av[r] - bv[q]);
        }
    }

    // ===== epilogue: direct store (+ mirrored store for off-diagonal symmetric tiles) =====
    const int row_end = p.row0 + p.n_rows, col_end = p.col0 + p.n_cols;
    const bool full_tile = (i_base + TILE <= row_end) && (j_base + TILE <= col_end) && p.vec_ok;
    if (full_tile) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
            float* o = row_ptr(p, i) + j_base;
            *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
            *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        }
        if (p.symmetric && I != J) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = j_base + (q < 4 ? tx * 4 + q : 64 + tx * 4 + (q - 4));
                float* o = row_ptr(p, j) + i_base;
                *reinterpret_cast<float4*>(o + ty * 4) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
                *reinterpret_cast<float4*>(o + 64 + ty * 4) = make_float4(acc[4][q], acc[5][q], acc[6][q], acc[7][q]);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
            if (i >= row_end) continue;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = j_base + (q < 4 ? tx * 4 + q : 64 + tx * 4 + (q - 4));
                if (j >= col_end) continue;
                row_ptr(p, i)[j] = acc[r][q];
                if (p.symmetric && I != J) row_ptr(p, j)[i] = acc[r][q];
            }
        }
    }
}

// ---- 128 x 64 tile variant -------------------------------------------------------------------
// Same pipeline with a 64-column B tile: 8 x 4 register tiles (32 accumulators) fit 3 CTAs per SM
// (24 warps instead of 16) and halve the tile granularity, which is what matters when a launch has
// only a few tiles per CTA slot (an 8-GPU run of the 20k-node graph: 5.2 waves of 128 x 128 tiles).
// Symmetric mode: tile row I (128 rows) owns column tiles J = 2I .. tiles_c-1 (64 columns each);
// J in {2I, 2I+1} together cover the diagonal 128 x 128 block in full, J >= 2I+2 is mirrored.
constexpr int TILE_N64 = 64;
constexpr uint32_t STAGE_BYTES_N64 = (uint32_t)KC * (TILE + TILE_N64) * sizeof(float);

struct __align__(128) PairSmemN64 {
    float a[STAGES][KC][TILE];
    float b[STAGES][KC][TILE_N64];
    unsigned long long full[STAGES];
    unsigned long long empty[STAGES];
};

// tile row i starts at first(i) = i * (tiles_c + 1 - i)
__device__ __forceinline__ void tri_decode_n64(int t, int tiles_r, int tiles_c, int& I, int& J) {
    const float b = (float)tiles_c + 1.f;
    int i = (int)((b - sqrtf(fmaxf(b * b - 4.f * (float)t, 0.f))) * 0.5f);
    i = max(0, min(i, tiles_r - 1));
    while (i > 0 && i * (tiles_c + 1 - i) > t) --i;
    while (i + 1 < tiles_r && (i + 1) * (tiles_c - i) <= t) ++i;
    I = i;
    J = 2 * i + (t - i * (tiles_c + 1 - i));
}

__global__ void __launch_bounds__(PAIR_THREADS, 3)
pairwise_l1_n64_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                       const PairArgs p) {
    extern __shared__ __align__(128) unsigned char pair_smem_raw[];
    PairSmemN64& sm = *reinterpret_cast<PairSmemN64*>(pair_smem_raw);

    int i_base, j_base;
    bool mirror;
    if (p.tile_list) {
        const int32_t* e = p.tile_list + 3 * (int64_t)blockIdx.x;
        i_base = __ldg(e);
        j_base = __ldg(e + 1);
        mirror = __ldg(e + 2) != 0;
    } else {
        int I, J;
        if (p.symmetric) {
            tri_decode_n64(blockIdx.x * p.tile_stride + p.tile_offset, p.tiles_r, p.tiles_c, I, J);
        } else {
            I = blockIdx.x / p.tiles_c;
            J = blockIdx.x - I * p.tiles_c;
        }
        i_base = p.row0 + I * TILE;
        j_base = p.col0 + J * TILE_N64;
        mirror = p.symmetric && (J >= 2 * I + 2);
    }

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), PAIR_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue_chunk = [&](int n) {
        const int s = n % STAGES;
        const uint32_t ph = (n / STAGES) & 1;
        mbar_wait(smem_u32(&sm.empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&sm.full[s]);
        mbar_expect_tx(full, STAGE_BYTES_N64);
        tma_load_2d(smem_u32(&sm.a[s][0][0]), &tmap_a, i_base, n * KC, full);
        tma_load_2d(smem_u32(&sm.b[s][0][0]), &tmap_b, j_base, n * KC, full);
    };
    if (tid == 0)
        for (int n = 0; n < LOOKAHEAD && n < p.k_chunks; ++n) issue_chunk(n);

    int tx, ty;
    tile_thread_pos(tid, tx, ty);
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    uint32_t ready = 0;
    for (int c = 0; c < p.k_chunks; ++c) {
        if (tid == 0 && c + LOOKAHEAD < p.k_chunks) issue_chunk(c + LOOKAHEAD);
        const int s = c % STAGES;
        const uint32_t ph = (c / STAGES) & 1;
        if (!ready) mbar_wait(smem_u32(&sm.full[s]), ph);
        ready = (c + 1 < p.k_chunks)
                    ? mbar_test(smem_u32(&sm.full[(c + 1) % STAGES]), ((c + 1) / STAGES) & 1) : 0u;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sm.a[s][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sm.a[s][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sm.b[s][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
            float d[8][4];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) d[r][q] = av[r] - bv[q];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] += fabsf(d[r][q]);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(smem_u32(&sm.empty[s]));
    }

    const int row_end = p.row0 + p.n_rows, col_end = p.col0 + p.n_cols;
    const bool full_tile = (i_base + TILE <= row_end) && (j_base + TILE_N64 <= col_end) && p.vec_ok;
    if (full_tile) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
            *reinterpret_cast<float4*>(row_ptr(p, i) + j_base + tx * 4) =
                make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
        if (mirror) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float* o = row_ptr(p, j_base + tx * 4 + q) + i_base;
                *reinterpret_cast<float4*>(o + ty * 4) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
                *reinterpret_cast<float4*>(o + 64 + ty * 4) = make_float4(acc[4][q], acc[5][q], acc[6][q], acc[7][q]);
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
            if (i >= row_end) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j_base + tx * 4 + q;
                if (j >= col_end) continue;
                row_ptr(p, i)[j] = acc[r][q];
                if (mirror) row_ptr(p, j)[i] = acc[r][q];
            }
        }
    }
}

// ---- v2: persistent CTAs, dynamic tile scheduler, packed subtract ---------------------------------
// (round 2) Same 128 x 128 tile and 8 x 8 register tiles, restructured around what ncu showed in
// round 1 (profiles/r1_pairwise_notes.md): the kernel was ISSUE-bound — 7.7 % of the issue slots
// went to instructions that are not FADD and the FMA pipe idled while they issued.
//  * PACKED: the 64 subtracts of a k-step are 32 `sub.f32x2` (SASS FADD2 with a broadcast .F32
//    operand: {a,a} - {b0,b1}); same IEEE result per lane, same FMA-pipe cycles, half the issue
//    slots -> 96 + 4 LDS slots per 128 pipe cycles, so operand loads and chunk bookkeeping issue
//    in the shadow of the pipe instead of displacing FADDs;
//  * persistent CTAs (grid = 2 per SM) walk the tile list; the TMA ring runs ACROSS tiles, so the
//    next tile's first chunks land while the current tile's 64 + 64 results are being stored;
//  * a dedicated producer warp (288 threads) was built and dropped: at 2 CTAs/SM ptxas then has 96
//    registers per thread, the 8 x 8 tile spills, and the variant measured 6.5 ms against 5.7 ms;
//  * the last K chunk runs only its valid rows (groups of 4 k-steps).
template <int KC_, int STAGES_>
struct __align__(128) PairSmemV2 {
    float a[STAGES_][KC_][TILE];
    float b[STAGES_][KC_][TILE];
    unsigned long long full[STAGES_];
    unsigned long long empty[STAGES_];
    int tile_of_stage[STAGES_];   // tile whose FIRST chunk sits in this stage (-1: no more tiles)
};

__device__ __forceinline__ uint32_t mbar_probe(uint32_t bar, uint32_t parity) {
    uint32_t ok;   // non-blocking test (try_wait may suspend the thread for a while)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}

template <bool PACKED>
__device__ __forceinline__ void pair_kstep(const float* __restrict__ arow, const float* __restrict__ brow,
                                           const int ty, const int tx, float (&acc)[8][8]) {
    const float4 a0 = *reinterpret_cast<const float4*>(arow + ty * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(arow + 64 + ty * 4);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    if (PACKED) {
        const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(brow + tx * 4);
        const ulonglong2 b1 = *reinterpret_cast<const ulonglong2*>(brow + 64 + tx * 4);
        const unsigned long long bp[4] = {b0.x, b0.y, b1.x, b1.y};
        unsigned long long d[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            unsigned long long aa;
            asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(av[r]));
#pragma unroll
            for (int q = 0; q < 4; ++q) asm("sub.f32x2 %0, %1, %2;" : "=l"(d[r][q]) : "l"(aa), "l"(bp[q]));
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d0, d1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d[r][q]));
                acc[r][2 * q] += fabsf(d0);
                acc[r][2 * q + 1] += fabsf(d1);
            }
    } else {
        const float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(brow + 64 + tx * 4);
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float d[8][8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) d[r][q] = av[r] - bv[q];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[r][q] += fabsf(d[r][q]);
    }
}

template <int KC_, int STAGES_, int UNROLL, bool PACKED>
__global__ void __launch_bounds__(PAIR_THREADS, 2)
pairwise_l1_v2_kernel(const __grid_constant__ CUtensorMap tmap, const PairArgs p) {
    using Smem = PairSmemV2<KC_, STAGES_>;
    constexpr uint32_t BYTES = 2u * KC_ * TILE * sizeof(float);
    extern __shared__ __align__(128) unsigned char pair_smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(pair_smem_raw);

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES_; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), PAIR_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int k_chunks = p.k_chunks;
    auto tile_origin = [&](int t, int& i_base, int& j_base, bool& mirror) {
        if (p.tile_list) {
            const int32_t* e = p.tile_list + 3 * (int64_t)t;
            i_base = __ldg(e);
            j_base = __ldg(e + 1);
            mirror = __ldg(e + 2) != 0;
            return;
        }
        int I, J;
        if (p.symmetric) {
            tri_decode(t * p.tile_stride + p.tile_offset, p.tiles_r, p.tiles_c, I, J);
        } else {
            I = t / p.tiles_c;
            J = t - I * p.tiles_c;
        }
        i_base = p.row0 + I * TILE;
        j_base = p.col0 + J * TILE;
        mirror = p.symmetric && I != J;
    };

    // ---- producer cursor: (tile, chunk), one stage per chunk.  Tiles come from a global counter
    // (dynamic scheduling): a static round-robin list left the SMs that finish early idle at the
    // end — ncu showed 3.69 of 4 warps per scheduler active on average, and more time blocked on
    // the full barriers, against 3.88 for one-tile CTAs scheduled by the hardware.  The tile id
    // travels to the consumers through shared memory, published by the barrier of its first chunk.
    int pt = 0, pc = 0, ps = 0, pi_base = 0, pj_base = 0;
    uint32_t pph = 0;
    bool pdone = false;
    auto produce_one = [&]() {
        if (pdone) return;
        mbar_wait(smem_u32(&sm.empty[ps]), pph ^ 1u);
        const uint32_t full = smem_u32(&sm.full[ps]);
        if (pc == 0) {
            pt = (int)atomicAdd(p.tile_counter, 1u);
            if (pt >= p.n_tiles) {
                sm.tile_of_stage[ps] = -1;
                mbar_arrive(full);          // completes the phase: the consumers wake up and leave
                pdone = true;
                return;
            }
            sm.tile_of_stage[ps] = pt;
            bool unused;
            tile_origin(pt, pi_base, pj_base, unused);
        }
        mbar_expect_tx(full, BYTES);
        tma_load_2d(smem_u32(&sm.a[ps][0][0]), &tmap, pi_base, pc * KC_, full);
        tma_load_2d(smem_u32(&sm.b[ps][0][0]), &tmap, pj_base, pc * KC_, full);
        if (++pc == k_chunks) pc = 0;
        if (++ps == STAGES_) { ps = 0; pph ^= 1u; }
    };

    if (tid == 0)
        for (int n = 0; n < p.lookahead; ++n) produce_one();

    // ===== consumers: 16 x 16 threads, each 8 x 8 outputs (2 x 2 blocks of 4 x 4) =====
    int tx, ty;
    tile_thread_pos(tid, tx, ty);
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    int s = 0;
    uint32_t ph = 0, ready = 0;
    const int row_end = p.row0 + p.n_rows, col_end = p.col0 + p.n_cols;
    for (;;) {
        int t = 0;
        for (int c = 0; c < k_chunks; ++c) {
            if (tid == 0) produce_one();
            if (!ready) mbar_wait(smem_u32(&sm.full[s]), ph);
            if (c == 0) {
                t = *((volatile int*)&sm.tile_of_stage[s]);
                if (t < 0) return;
            }
            const int ns = (s + 1 == STAGES_) ? 0 : s + 1;
            const uint32_t nph = (ns == 0) ? ph ^ 1u : ph;
            // look at the NEXT chunk's barrier now (its TMA was issued long ago): the round trip
            // hides under this chunk's FADDs instead of heading the next chunk's dependency chain
            ready = mbar_probe(smem_u32(&sm.full[ns]), nph);
            // One rolled loop of 8-k-step bodies serves full chunks and the (shorter) last chunk alike.
            // The body must stay ~17 KB of SASS: when ptxas sees a constant trip count it peels the
            // first iteration, the chunk becomes 34 KB of straight-line code (> 32 KB instruction
            // cache) and the warps stall on instruction fetch (measured: 69 % instead of 84 %).
            const int kn = (c + 1 < k_chunks) ? KC_ : p.k_last;   // valid rows of this chunk, multiple of 4
            int kk = 0;
#pragma unroll 1
            for (; kk + UNROLL <= kn; kk += UNROLL) {
#pragma unroll
                for (int j = 0; j < UNROLL; ++j)
                    pair_kstep<PACKED>(&sm.a[s][kk + j][0], &sm.b[s][kk + j][0], ty, tx, acc);
            }
#pragma unroll 1
            for (; kk < kn; kk += 4) {      // remainder of the last chunk, groups of 4 rows
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    pair_kstep<PACKED>(&sm.a[s][kk + j][0], &sm.b[s][kk + j][0], ty, tx, acc);
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(smem_u32(&sm.empty[s]));
            s = ns;
            ph = nph;
        }

        // ---- epilogue: direct store (+ mirrored store for off-diagonal symmetric tiles) ----
        int i_base, j_base;
        bool mirror;
        tile_origin(t, i_base, j_base, mirror);
        const bool full_tile = (i_base + TILE <= row_end) && (j_base + TILE <= col_end) && p.vec_ok &&
                               (!p.shard_ptrs || p.per >= TILE);
        if (full_tile) {
            const TileRows ri = tile_rows(p, i_base);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
                float* o = ri.row(i) + j_base;
                *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
            }
            if (mirror) {
                const TileRows rj = tile_rows(p, j_base);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = j_base + (q < 4 ? tx * 4 + q : 64 + tx * 4 + (q - 4));
                    float* o = rj.row(j) + i_base;
                    *reinterpret_cast<float4*>(o + ty * 4) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
                    *reinterpret_cast<float4*>(o + 64 + ty * 4) = make_float4(acc[4][q], acc[5][q], acc[6][q], acc[7][q]);
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int i = i_base + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
                if (i >= row_end) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = j_base + (q < 4 ? tx * 4 + q : 64 + tx * 4 + (q - 4));
                    if (j >= col_end) continue;
                    row_ptr(p, i)[j] = acc[r][q];
                    if (mirror) row_ptr(p, j)[i] = acc[r][q];
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    }
}

// ---- FP32 issue-peak probe: same instruction mix as the inner loop, no memory ----
__global__ void __launch_bounds__(256, 2) fp32_peak_probe_kernel(float* sink, int iters) {
    float a[8], b[8], acc[8][8];
    const float seed = (float)(threadIdx.x & 7) * 0.125f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        a[r] = seed + r;
        b[r] = seed * 0.5f - r;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = 0.f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[r][q] += fabsf(a[r] - b[q]);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a[r] += 0.001f;
            b[r] += 0.002f;
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) s += acc[r][q];
    if (s == -1.f) sink[0] = s;  // never true; keeps the loop alive
}

// ---- tensor map creation through the runtime's driver entry point ----
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

}  // namespace hsd

namespace hsd {

// ---- v2 dispatch ----
struct V2Config { int kc, stages, packed, prodw, lookahead, unroll; };

static V2Config v2_config() {
    // HSD_PAIR_V2="kc,stages,packed,prodw,lookahead,unroll" (tuning knob, read once); "0" = round-1 kernels
    static V2Config cfg = {-1, 0, 0, 0, 0, 0};
    if (cfg.kc < 0) {
        V2Config c = {HSD_PAIR_V2_DEFAULT};
        const char* e = getenv("HSD_PAIR_V2");
        if (e) {
            int v[6] = {c.kc, c.stages, c.packed, c.prodw, c.lookahead, c.unroll};
            int n = 0;
            const char* q = e;
            while (n < 6 && *q) {
                v[n++] = atoi(q);
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
            c = {v[0], v[1], v[2], v[3], v[4], v[5]};
        }
        cfg = c;
    }
    return cfg;
}

// Tile counters of the dynamic scheduler: a small ring in static device memory (no allocation);
// every launch takes the next one and zeroes it in-stream first, so launches that overlap on
// different streams (up to 64 in flight per device) do not share a counter.
__device__ unsigned int g_tile_counters[64];

template <int KC_, int STAGES_, bool PACKED, int UNROLL>
static int launch_v2(const CUtensorMap& tmap, const PairArgs& a, cudaStream_t stream, int sms) {
    auto kern = pairwise_l1_v2_kernel<KC_, STAGES_, UNROLL, PACKED>;
    const int smem = (int)sizeof(PairSmemV2<KC_, STAGES_>);
    HSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)std::min<long long>(a.n_tiles, 2ll * sms);
    kern<<<grid, PAIR_THREADS, smem, stream>>>(tmap, a);
    HSD_CUDA_TRY(cudaGetLastError());
    return HSD_OK;
}

static int encode_table(CUtensorMap* tmap, const float* sigT, int32_t k_rows, int64_t n_pad, int box_n, int box_k) {
    auto encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return HSD_ERR_NO_DEVICE;
    }
    // rows >= k_rows and columns >= n_pad are out of bounds: TMA fills them with zeros
    const cuuint64_t gdim[2] = {(cuuint64_t)n_pad, (cuuint64_t)k_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)n_pad * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_n, (cuuint32_t)box_k};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sigT), gdim,
                         gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return HSD_ERR_CUDA;
    }
    return HSD_OK;
}

static int launch_pairwise_v2(const V2Config& c, const float* sigT, int32_t k_used, int64_t n_pad, PairArgs a,
                              long long n_tiles, cudaStream_t stream) {
    HSD_REQUIRE(n_tiles < (1ll << 31), "too many tiles for one launch");
    if (n_tiles <= 0) return HSD_OK;
    CUtensorMap tmap;
    const int rc = encode_table(&tmap, sigT, k_used, n_pad, TILE, c.kc);
    if (rc != HSD_OK) return rc;
    a.k_chunks = (k_used + c.kc - 1) / c.kc;
    a.k_last = ((k_used - (a.k_chunks - 1) * c.kc) + 3) / 4 * 4;
    a.n_tiles = (int)n_tiles;
    a.lookahead = std::max(1, std::min(c.lookahead, c.stages - 1));
    {
        static thread_local unsigned int next = 0;
        unsigned int* base = nullptr;
        HSD_CUDA_TRY(cudaGetSymbolAddress((void**)&base, g_tile_counters));
        a.tile_counter = base + (next++ & 63u);
        HSD_CUDA_TRY(cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned int), stream));
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define HSD_V2_CASE(KC_, ST_)                                                                        \
    if (c.kc == KC_ && c.stages == ST_) {                                                            \
        if (c.packed && c.unroll == 16) return launch_v2<KC_, ST_, true, 16>(tmap, a, stream, sms);  \
        if (c.packed) return launch_v2<KC_, ST_, true, 8>(tmap, a, stream, sms);                     \
        return launch_v2<KC_, ST_, false, 8>(tmap, a, stream, sms);                                  \
    }
    HSD_V2_CASE(16, 4)
    HSD_V2_CASE(32, 3)
#undef HSD_V2_CASE
    set_error("HSD_PAIR_V2: unsupported kc/stages %d/%d (16/4, 32/3)", c.kc, c.stages);
    return HSD_ERR_INVALID;
}

static int launch_pairwise(const float* sigT, int32_t k_used, int64_t n_pad, PairArgs a, long long n_tiles,
                           cudaStream_t stream) {
    const int32_t k_pad = (k_used + KC - 1) / KC * KC;
    auto encode = get_encode();
    if (!encode) {
        set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
        return HSD_ERR_NO_DEVICE;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)n_pad, (cuuint64_t)k_pad};
    const cuuint64_t gstride[1] = {(cuuint64_t)n_pad * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TILE, (cuuint32_t)KC};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sigT), gdim,
                         gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return HSD_ERR_CUDA;
    }
    a.k_chunks = k_pad / KC;
    HSD_REQUIRE(n_tiles < (1ll << 31), "too many tiles for one launch");
    if (n_tiles <= 0) return HSD_OK;

    static int tile_n = -1;   // HSD_PAIR_TILE_N=64 / 128 forces a variant; default: by tiles per CTA slot
    if (tile_n < 0) { const char* e = getenv("HSD_PAIR_TILE_N"); tile_n = e ? atoi(e) : 0; }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    bool use_n64 = tile_n == 64 || (tile_n == 0 && n_tiles < (long long)sms * 2 * 8);
    if (a.tile_list) use_n64 = (a.list_tile_n == 64);     // an explicit tile list fixes the tile shape
    if (!use_n64) {
        // large launches: persistent, dynamically scheduled 128 x 128 kernel (v2); the 128 x 64 kernel
        // below keeps the small launches (few tiles per CTA slot: its finer tiles quantise better)
        const V2Config c = v2_config();
        if (c.kc > 0) return launch_pairwise_v2(c, sigT, k_used, n_pad, a, n_tiles, stream);
    }
    if (use_n64) {
        // re-derive the tile grid for 64-column tiles
        const int tr = a.tiles_r;
        const int tc = (a.n_cols + TILE_N64 - 1) / TILE_N64;
        const long long total = a.symmetric ? (long long)tr * tc - (long long)tr * (tr - 1) : (long long)tr * tc;
        const long long mine = a.tile_list ? n_tiles
                               : (total > a.tile_offset ? (total - a.tile_offset + a.tile_stride - 1) / a.tile_stride : 0);
        if (mine <= 0) return HSD_OK;
        HSD_REQUIRE(mine < (1ll << 31), "too many tiles for one launch");
        a.tiles_c = tc;
        CUtensorMap tmap_b;
        const cuuint32_t box_b[2] = {(cuuint32_t)TILE_N64, (cuuint32_t)KC};
        cr = encode(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(sigT), gdim, gstride, box_b,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled (B tile) failed with CUresult %d", (int)cr);
            return HSD_ERR_CUDA;
        }
        const int smem64 = (int)sizeof(PairSmemN64);
        HSD_CUDA_TRY(cudaFuncSetAttribute(pairwise_l1_n64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64));
        pairwise_l1_n64_kernel<<<(unsigned)mine, PAIR_THREADS, smem64, stream>>>(tmap, tmap_b, a);
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    }

    if (a.tile_list) {
        set_error("explicit tile lists need the v2 kernel (HSD_PAIR_V2 must not be 0)");
        return HSD_ERR_UNSUPPORTED;
    }
    const int smem = (int)sizeof(PairSmem);
    static int unroll = 0;   // tuning knob, read once: HSD_PAIR_UNROLL in {4, 8, 16}
    if (!unroll) {
        const char* e = getenv("HSD_PAIR_UNROLL");
        unroll = e ? atoi(e) : 8;
    }
    auto launch = [&](auto kern) -> int {
        HSD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<(unsigned)n_tiles, PAIR_THREADS, smem, stream>>>(tmap, a);
        HSD_CUDA_TRY(cudaGetLastError());
        return HSD_OK;
    };
    static int use_tail = -1;   // HSD_PAIR_TAIL=0 disables the peeled last chunk
    if (use_tail < 0) { const char* e = getenv("HSD_PAIR_TAIL"); use_tail = e ? atoi(e) : 1; }
    const int k_last = k_used - (a.k_chunks - 1) * KC;   // 1..KC valid rows in the last chunk
    if (unroll == 4) return launch(pairwise_l1_kernel<4, 0>);
    if (unroll == 16) return launch(pairwise_l1_kernel<16, 0>);
    if (use_tail && k_last <= 4) return launch(pairwise_l1_kernel<8, 4>);
    if (use_tail && k_last <= 8) return launch(pairwise_l1_kernel<8, 8>);
    return launch(pairwise_l1_kernel<8, 0>);
}

}  // namespace hsd

extern "C" int hsd_pairwise_l1(const float* sigT, int32_t k_used, int64_t n_pad, int32_t row0,
                               int32_t n_rows, int32_t col0, int32_t n_cols, int32_t symmetric,
                               float* out, int64_t ld_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sigT && out, "null pointer");
    HSD_REQUIRE(k_used > 0, "k_used must be positive");
    HSD_REQUIRE(n_pad > 0 && n_pad % 4 == 0, "n_pad must be a multiple of 4");
    HSD_REQUIRE((reinterpret_cast<uintptr_t>(sigT) & 15) == 0, "sigT must be 16-byte aligned");
    HSD_REQUIRE(row0 >= 0 && col0 >= 0 && n_rows >= 0 && n_cols >= 0, "negative range");
    // TMA tile loads start at element row0 / col0 of a K-major row: the address must be 16-byte aligned
    HSD_REQUIRE(row0 % 4 == 0 && col0 % 4 == 0, "row0 and col0 must be multiples of 4 (16-byte TMA alignment)");
    HSD_REQUIRE(row0 + (int64_t)n_rows <= n_pad && col0 + (int64_t)n_cols <= n_pad, "range exceeds n_pad");
    HSD_REQUIRE(!symmetric || (row0 == col0 && n_cols >= n_rows), "symmetric needs row0 == col0 and n_cols >= n_rows");
    if (n_rows == 0 || n_cols == 0) return HSD_OK;
    PairArgs a = {};
    a.row0 = row0; a.n_rows = n_rows; a.col0 = col0; a.n_cols = n_cols;
    a.tiles_r = (n_rows + TILE - 1) / TILE;
    a.tiles_c = (n_cols + TILE - 1) / TILE;
    a.symmetric = symmetric ? 1 : 0;
    a.out = out; a.ld = ld_out;
    a.vec_ok = (ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    a.shard_ptrs = nullptr; a.per = 1; a.tile_stride = 1; a.tile_offset = 0;
    const long long n_tiles = symmetric ? (long long)a.tiles_r * a.tiles_c - (long long)a.tiles_r * (a.tiles_r - 1) / 2
                                        : (long long)a.tiles_r * a.tiles_c;
    return launch_pairwise(sigT, k_used, n_pad, a, n_tiles, (cudaStream_t)stream);
}

extern "C" int hsd_pairwise_l1_sharded(const float* sigT, int32_t k_used, int64_t n_pad, int32_t n_nodes,
                                       int32_t rank, int32_t world, int32_t rows_per_rank,
                                       float* const* shard_ptrs, int64_t ld_out, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sigT && shard_ptrs, "null pointer");
    HSD_REQUIRE(k_used > 0, "k_used must be positive");
    HSD_REQUIRE(n_pad > 0 && n_pad % 4 == 0 && n_nodes > 0 && n_nodes <= n_pad, "bad n_pad / n_nodes");
    HSD_REQUIRE((reinterpret_cast<uintptr_t>(sigT) & 15) == 0, "sigT must be 16-byte aligned");
    HSD_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    HSD_REQUIRE(rows_per_rank > 0 && (int64_t)rows_per_rank * world >= n_nodes, "row blocks do not cover the matrix");
    HSD_REQUIRE(ld_out >= n_nodes && ld_out % 4 == 0, "ld_out must be >= n_nodes and a multiple of 4");
    PairArgs a = {};
    a.row0 = 0; a.n_rows = n_nodes; a.col0 = 0; a.n_cols = n_nodes;
    a.tiles_r = a.tiles_c = (n_nodes + TILE - 1) / TILE;
    a.symmetric = 1;
    a.out = nullptr; a.ld = ld_out;
    a.vec_ok = 1;   // block bases come from the allocator (>= 256-byte aligned), ld_out % 4 == 0
    a.shard_ptrs = shard_ptrs; a.per = rows_per_rank; a.tile_stride = world; a.tile_offset = rank;
    const long long total = (long long)a.tiles_r * (a.tiles_r + 1) / 2;
    const long long mine = total > rank ? (total - rank + world - 1) / world : 0;
    return launch_pairwise(sigT, k_used, n_pad, a, mine, (cudaStream_t)stream);
}

extern "C" int hsd_fp32_peak_probe(float* sink, int32_t iters, int64_t* lane_ops_host, void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sink && iters > 0, "bad arguments");
    int dev = 0, sms = 0;
    HSD_CUDA_TRY(cudaGetDevice(&dev));
    HSD_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 2 * 4;
    fp32_peak_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(sink, iters);
    HSD_CUDA_TRY(cudaGetLastError());
    // 64 sub + 64 |.|-accumulate + 16 operand updates = 144 FADD per thread-iteration
    if (lane_ops_host) *lane_ops_host = (int64_t)blocks * 256 * (int64_t)iters * (128 + 16);
    return HSD_OK;
}

extern "C" int hsd_pairwise_l1_tile_list(const float* sigT, int32_t k_used, int64_t n_pad, int32_t n_nodes,
                                         const int32_t* tile_list, int32_t n_tiles, int32_t tile_n,
                                         int32_t rows_per_rank, float* const* shard_ptrs, int64_t ld_out,
                                         void* stream) {
    using namespace hsd;
    HSD_REQUIRE(sigT && shard_ptrs && (tile_list || n_tiles == 0), "null pointer");
    HSD_REQUIRE(k_used > 0, "k_used must be positive");
    HSD_REQUIRE(n_pad > 0 && n_pad % 4 == 0 && n_nodes > 0 && n_nodes <= n_pad, "bad n_pad / n_nodes");
    HSD_REQUIRE((reinterpret_cast<uintptr_t>(sigT) & 15) == 0, "sigT must be 16-byte aligned");
    HSD_REQUIRE(tile_n == 128 || tile_n == 64, "tile_n must be 128 or 64");
    HSD_REQUIRE(rows_per_rank > 0 && n_tiles >= 0, "bad sizes");
    HSD_REQUIRE(ld_out >= n_nodes && ld_out % 4 == 0, "ld_out must be >= n_nodes and a multiple of 4");
    if (n_tiles == 0) return HSD_OK;
    PairArgs a = {};
    a.row0 = 0; a.n_rows = n_nodes; a.col0 = 0; a.n_cols = n_nodes;
    a.tiles_r = a.tiles_c = (n_nodes + TILE - 1) / TILE;
    a.symmetric = 1;
    a.out = nullptr; a.ld = ld_out;
    a.vec_ok = 1;
    a.shard_ptrs = shard_ptrs; a.per = rows_per_rank; a.tile_stride = 1; a.tile_offset = 0;
    a.tile_list = tile_list; a.list_tile_n = tile_n;
    return launch_pairwise(sigT, k_used, n_pad, a, n_tiles, (cudaStream_t)stream);
}
