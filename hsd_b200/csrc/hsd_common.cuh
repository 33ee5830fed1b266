// Shared helpers for the HSD sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hsd_b200.h"

namespace hsd {

void set_error(const char* fmt, ...);

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return HSD_ERR_CUDA;
    }
    return HSD_OK;
}

#define HSD_CUDA_TRY(expr)                                         \
    do {                                                           \
        int _rc = ::hsd::check_cuda((expr), #expr);                \
        if (_rc != HSD_OK) return _rc;                             \
    } while (0)

#define HSD_REQUIRE(cond, msg)                                     \
    do {                                                           \
        if (!(cond)) {                                             \
            ::hsd::set_error("%s: %s", __func__, msg);             \
            return HSD_ERR_INVALID;                                \
        }                                                          \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Block-wide exclusive scan of one int per thread. THREADS must be a multiple
// of 32 and <= 1024. `warp_tot` is THREADS/32 ints of shared scratch.
// Returns the exclusive prefix; *total receives the block sum.
template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) {
        int t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    *total = tot;
    __syncthreads();  // warp_tot reusable after return
    return base + inc - v;
}

}  // namespace hsd
