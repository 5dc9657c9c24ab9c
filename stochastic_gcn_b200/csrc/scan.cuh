// Device-wide exclusive scan of int32 with a host-side length BOUND and a device-side length.
// Three launches for large inputs (tile scan -> scan of tile sums -> add), one launch when the
// bound fits one tile.  Lengths are read from device memory so no host round-trip is needed for
// data-dependent sizes.
#pragma once

#include "common.cuh"

namespace sgcn {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;   // 2048

// Exclusive scan of one value per thread across a 256-thread block; returns the exclusive prefix
// and writes the block total to *total (all threads).
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_tot[kScanThreads / 32];
    __shared__ int block_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanThreads / 32 ? warp_tot[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < kScanThreads / 32) warp_tot[lane] = winc - w;   // exclusive warp offsets
        if (lane == kScanThreads / 32 - 1) block_tot = winc;
    }
    __syncthreads();
    int excl = inc - v + warp_tot[warp];
    *total = block_tot;
    __syncthreads();   // shared arrays are reused by the next call
    return excl;
}

// One tile per block.  out[i] = exclusive prefix within the tile; tile_sums[b] = tile total.
// When single_tile != 0 the block also finalises: *total_out = total and, if end_slot, out[n] = total.
static __global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ n_dev,
                  int n_bound, int* __restrict__ tile_sums, int single_tile,
                  int* __restrict__ total_out, int end_slot) {
    const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int v[kScanItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int total;
    int excl = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 0) {
        if (single_tile) {
            if (total_out) *total_out = total;
            if (end_slot) out[n] = total;
        } else {
            tile_sums[blockIdx.x] = total;
        }
    }
}

// Single block: exclusive scan of the tile sums in place; writes the grand total.
static __global__ void __launch_bounds__(kScanThreads)
scan_sums_kernel(int* __restrict__ tile_sums, int n_tiles, int* __restrict__ out,
                 const int* __restrict__ n_dev, int n_bound, int* __restrict__ total_out,
                 int end_slot) {
    int carry = 0;
    for (int base = 0; base < n_tiles; base += kScanThreads) {
        int i = base + threadIdx.x;
        int v = i < n_tiles ? tile_sums[i] : 0;
        int total;
        int excl = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = carry;
        if (end_slot) {
            const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
            out[n] = carry;
        }
    }
}

static __global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(int* __restrict__ out, const int* __restrict__ tile_sums,
                const int* __restrict__ n_dev, int n_bound) {
    const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
    const int add = tile_sums[blockIdx.x];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) out[base + k] += add;
}

// out[0..n) = exclusive scan of in[0..n); optionally out[n] = total and *total_out = total.
// `out` may alias `in`.  tile_sums must hold div_up(n_bound, kScanTile) ints.
inline int launch_exclusive_scan(const int* in, int* out, const int* n_dev, int n_bound,
                                 int* tile_sums, int* total_out, bool end_slot,
                                 cudaStream_t stream) {
    if (n_bound <= 0) n_bound = 1;   // still writes total = 0 / out[0] = 0
    const int n_tiles = div_up(n_bound, kScanTile);
    if (n_tiles == 1) {
        scan_tiles_kernel<<<1, kScanThreads, 0, stream>>>(in, out, n_dev, n_bound, tile_sums, 1,
                                                          total_out, end_slot ? 1 : 0);
        SGCN_LAUNCHED();
        return SGCN_OK;
    }
    scan_tiles_kernel<<<n_tiles, kScanThreads, 0, stream>>>(in, out, n_dev, n_bound, tile_sums, 0,
                                                            nullptr, 0);
    SGCN_LAUNCHED();
    scan_sums_kernel<<<1, kScanThreads, 0, stream>>>(tile_sums, n_tiles, out, n_dev, n_bound,
                                                     total_out, end_slot ? 1 : 0);
    SGCN_LAUNCHED();
    scan_add_kernel<<<n_tiles, kScanThreads, 0, stream>>>(out, tile_sums, n_dev, n_bound);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

}  // namespace sgcn
