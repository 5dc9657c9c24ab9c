// Write-back exchange payload layout and the pack / push body (shared with the aggregate kernels, which
// can carry the push as extra blocks of the sampled-aggregate launch).
#pragma once

#include "common.cuh"

namespace sgcn {

constexpr int kWbHeaderInts = 4;   // {count, step, 0, 0}: keeps ids 16-byte aligned

__host__ __device__ inline int64_t wb_ids_offset() { return kWbHeaderInts * 4; }
__host__ __device__ inline int64_t wb_rows_offset(int n_bound) {
    return wb_ids_offset() + (((int64_t)n_bound * 4 + 15) & ~int64_t(15));
}
__host__ __device__ inline int64_t wb_payload_bytes(int n_bound, int D) {
    return (wb_rows_offset(n_bound) + (int64_t)n_bound * D * 4 + 255) & ~int64_t(255);
}

constexpr int kMaxPeers = 16;
struct PeerPtrs { char* p[kMaxPeers]; };

struct WbPushArgs {
    const int32_t* field; const int32_t* n_dev; int n_bound;
    const float* rows; int64_t ld_rows; int D;
    PeerPtrs dst_even, dst_odd; int n_dst; int step; int32_t* epoch; PeerPtrs flags; int my_rank;
    int32_t* block_counter;
    // ring > 0: epoch e lands in receive area e % ring, at dst_even.p[k] + (e % ring) * ring_stride (dst_odd
    // unused); ring == 0: the two-area (even / odd) protocol
    int ring; int64_t ring_stride;
};

// pack {count, ids, rows} into up to `n_dst` destinations (own buffer and/or peers' receive slots);
// executed by `nblocks` CTAs of 256 threads, this one being number `bid`
__device__ __forceinline__ void wb_pack_body(const WbPushArgs& a, int bid, int nblocks) {
    // peer transport: this push belongs to epoch *epoch + 1 and lands in the slot set of its parity
    int step = a.step;
    if (a.epoch) step = *a.epoch + 1;
    PeerPtrs dst = (a.epoch && a.ring == 0 && (step & 1)) ? a.dst_odd : a.dst_even;
    if (a.ring > 0) {
        const int64_t off = (int64_t)(step % a.ring) * a.ring_stride;
        for (int k = 0; k < a.n_dst; ++k) dst.p[k] += off;
    }
    const int n = min(*a.n_dev, a.n_bound);
    const int D = a.D, n_dst = a.n_dst;
    const int64_t t0 = (int64_t)bid * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)nblocks * blockDim.x;
    const int64_t ids_off = wb_ids_offset(), rows_off = wb_rows_offset(a.n_bound);
    if (t0 == 0)
        for (int k = 0; k < n_dst; ++k) {
            int32_t* h = (int32_t*)dst.p[k];
            h[0] = n; h[1] = step; h[2] = 0; h[3] = 0;
        }
    for (int64_t i = t0; i < n; i += stride) {
        const int32_t id = a.field[i];
        for (int k = 0; k < n_dst; ++k) ((int32_t*)(dst.p[k] + ids_off))[i] = id;
    }
    if ((D & 3) == 0 && (a.ld_rows & 3) == 0 && (((uintptr_t)a.rows) & 15) == 0) {
        const int d4 = D >> 2;
        const int64_t total = (int64_t)n * d4;
        for (int64_t t = t0; t < total; t += stride) {
            const int64_t r = t / d4;
            const int c = (int)(t - r * d4) * 4;
            const float4 v = ldg_stream4(a.rows + r * a.ld_rows + c);
            for (int k = 0; k < n_dst; ++k) *(float4*)(dst.p[k] + rows_off + (r * D + c) * 4) = v;
        }
    } else {
        const int64_t total = (int64_t)n * D;
        for (int64_t t = t0; t < total; t += stride) {
            const int64_t r = t / D;
            const int c = (int)(t - r * D);
            const float v = a.rows[r * a.ld_rows + c];
            for (int k = 0; k < n_dst; ++k) *(float*)(dst.p[k] + rows_off + (r * D + c) * 4) = v;
        }
    }
    if (a.block_counter) {
        // fused signal: the last block to finish advances the epoch and publishes it to every rank
        __shared__ int s_last;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(a.block_counter, 1) == nblocks - 1;
        __syncthreads();
        if (s_last && threadIdx.x == 0) {
            *a.block_counter = 0;
            *a.epoch = step;
            __threadfence_system();
            for (int k = 0; k < n_dst; ++k) {
                volatile int32_t* f = (volatile int32_t*)a.flags.p[k];
                f[a.my_rank] = step;
            }
            __threadfence_system();
        }
    }
}

}  // namespace sgcn
