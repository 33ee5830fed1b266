// Host side of the e2e path: complete a symmetric matrix whose UPPER trapezoid panels arrive from
// the GPU.  The reference fills dist_mat[i, j] = dist_mat[j, i] pair by pair (model/HSD.py:112); here
// the GPU ships each unordered pair once over PCIe (the link is the e2e bottleneck: 1.6 GB at
// ~55 GB/s for the 20k-node matrix) and the host cores mirror the panel into the rows below it while
// the next panels are still in flight.  Blocked 64 x 64 transposes, 8 x 8 AVX2 micro-kernel with
// streaming stores when the CPU has AVX2 (runtime check), plain loops otherwise.
#include <stdint.h>
#include <algorithm>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "../../include/hsd_b200.h"

namespace {

constexpr int BLK = 64;

// dst[c][r] = src[r][c] for an R x C block (any sizes), scalar
inline void transpose_scalar(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int R, int C) {
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) dst[(int64_t)c * ld_dst + r] = src[(int64_t)r * ld_src + c];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline void transpose8x8_avx2(const float* src, int64_t ld_src, float* dst,
                                                              int64_t ld_dst, bool stream) {
    __m256 r0 = _mm256_loadu_ps(src + 0 * ld_src), r1 = _mm256_loadu_ps(src + 1 * ld_src);
    __m256 r2 = _mm256_loadu_ps(src + 2 * ld_src), r3 = _mm256_loadu_ps(src + 3 * ld_src);
    __m256 r4 = _mm256_loadu_ps(src + 4 * ld_src), r5 = _mm256_loadu_ps(src + 5 * ld_src);
    __m256 r6 = _mm256_loadu_ps(src + 6 * ld_src), r7 = _mm256_loadu_ps(src + 7 * ld_src);
    __m256 t0 = _mm256_unpacklo_ps(r0, r1), t1 = _mm256_unpackhi_ps(r0, r1);
    __m256 t2 = _mm256_unpacklo_ps(r2, r3), t3 = _mm256_unpackhi_ps(r2, r3);
    __m256 t4 = _mm256_unpacklo_ps(r4, r5), t5 = _mm256_unpackhi_ps(r4, r5);
    __m256 t6 = _mm256_unpacklo_ps(r6, r7), t7 = _mm256_unpackhi_ps(r6, r7);
    __m256 u0 = _mm256_shuffle_ps(t0, t2, 0x44), u1 = _mm256_shuffle_ps(t0, t2, 0xEE);
    __m256 u2 = _mm256_shuffle_ps(t1, t3, 0x44), u3 = _mm256_shuffle_ps(t1, t3, 0xEE);
    __m256 u4 = _mm256_shuffle_ps(t4, t6, 0x44), u5 = _mm256_shuffle_ps(t4, t6, 0xEE);
    __m256 u6 = _mm256_shuffle_ps(t5, t7, 0x44), u7 = _mm256_shuffle_ps(t5, t7, 0xEE);
    __m256 o[8];
    o[0] = _mm256_permute2f128_ps(u0, u4, 0x20); o[1] = _mm256_permute2f128_ps(u1, u5, 0x20);
    o[2] = _mm256_permute2f128_ps(u2, u6, 0x20); o[3] = _mm256_permute2f128_ps(u3, u7, 0x20);
    o[4] = _mm256_permute2f128_ps(u0, u4, 0x31); o[5] = _mm256_permute2f128_ps(u1, u5, 0x31);
    o[6] = _mm256_permute2f128_ps(u2, u6, 0x31); o[7] = _mm256_permute2f128_ps(u3, u7, 0x31);
    if (stream) {
        for (int k = 0; k < 8; ++k) _mm256_stream_ps(dst + k * ld_dst, o[k]);
    } else {
        for (int k = 0; k < 8; ++k) _mm256_storeu_ps(dst + k * ld_dst, o[k]);
    }
}

// dst[c][r] = src[r][c], R and C multiples of 8; for a fixed group of 8 destination rows the 32-byte
// pieces are written left to right so the write-combining buffers fill whole cache lines
__attribute__((target("avx2"))) void transpose_block_avx2(const float* src, int64_t ld_src, float* dst,
                                                          int64_t ld_dst, int R, int C, bool stream) {
    for (int c = 0; c < C; c += 8)
        for (int r = 0; r < R; r += 8)
            transpose8x8_avx2(src + (int64_t)r * ld_src + c, ld_src, dst + (int64_t)c * ld_dst + r, ld_dst, stream);
}
#endif

struct MirrorJob {
    float* D;
    int64_t ld;
    int n, rb, re;
    bool avx2, stream;
};

// all destination blocks of block-row Jb: rows [Jb*BLK, ..) of the lower part, columns in [rb, re)
void mirror_block_row(const MirrorJob& m, int Jb) {
    const int j0 = Jb * BLK, j1 = std::min(j0 + BLK, m.n);
    for (int i0 = m.rb; i0 < m.re && i0 <= j0; i0 += BLK) {   // rb is a multiple of BLK
        const int i1 = std::min(std::min(i0 + BLK, m.re), m.n);
        const float* src = m.D + (int64_t)i0 * m.ld + j0;      // upper block: rows i, columns j
        float* dst = m.D + (int64_t)j0 * m.ld + i0;            // lower block: rows j, columns i
        if (i0 == j0) {                                       // diagonal block: strictly-lower entries only
            for (int j = j0; j < j1; ++j)
                for (int i = i0; i < std::min(i1, j); ++i) m.D[(int64_t)j * m.ld + i] = m.D[(int64_t)i * m.ld + j];
            continue;
        }
        const int R = i1 - i0, C = j1 - j0;
#if defined(__x86_64__)
        if (m.avx2 && R % 8 == 0 && C % 8 == 0) {
            transpose_block_avx2(src, m.ld, dst, m.ld, R, C, m.stream);
            continue;
        }
#endif
        transpose_scalar(src, m.ld, dst, m.ld, R, C);
    }
}

}  // namespace

extern "C" int hsd_mirror_upper_to_lower_host(float* D_host, int64_t ld, int32_t n, int32_t row_begin,
                                              int32_t row_end, int32_t n_threads) {
    if (!D_host || ld < n || n < 0 || row_begin < 0 || row_end > n || row_begin > row_end) return HSD_ERR_INVALID;
    if (row_begin % BLK != 0) return HSD_ERR_INVALID;          // panels start on tile boundaries (128)
    if (row_begin == row_end) return HSD_OK;
    MirrorJob m;
    m.D = D_host; m.ld = ld; m.n = n; m.rb = row_begin; m.re = row_end;
#if defined(__x86_64__)
    m.avx2 = __builtin_cpu_supports("avx2");
#else
    m.avx2 = false;
#endif
    // streaming stores need 32-byte aligned destinations: base and leading dimension
    m.stream = m.avx2 && (reinterpret_cast<uintptr_t>(D_host) % 32 == 0) && (ld % 8 == 0);
    const int jb0 = row_begin / BLK, jb1 = (n + BLK - 1) / BLK;
    const int T = std::max(1, std::min<int>(n_threads, jb1 - jb0));
    auto work = [&](int t) {
        for (int Jb = jb0 + t; Jb < jb1; Jb += T) mirror_block_row(m, Jb);
#if defined(__x86_64__)
        if (m.stream) _mm_sfence();
#endif
    };
    if (T == 1) {
        work(0);
        return HSD_OK;
    }
    std::vector<std::thread> pool;
    pool.reserve(T - 1);
    for (int t = 1; t < T; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    return HSD_OK;
}
