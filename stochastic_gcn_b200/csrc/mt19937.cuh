// std::mt19937 on the device: block-parallel state regeneration + tempering, and the float
// conversion std::uniform_real_distribution<float> performs in libstdc++ (generate_canonical).
// The sampler must consume exactly the stream the reference's `generator` / Mult::generator
// produce (gcn/scheduler.h:11, gcn/mult.h:26), one 32-bit output per draw.
#pragma once

#include "common.cuh"

namespace sgcn {

constexpr int kMtN = 624;
constexpr int kMtM = 397;
constexpr int kMtThreads = 256;
// device layout of an engine: 624 state words followed by the cursor
constexpr int kMtWords = kMtN + 1;

__host__ __device__ inline uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__host__ __device__ inline uint32_t mt_temper(uint32_t z) {
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}

// std::mt19937::seed(s): cursor at 624 so that the first draw regenerates the block
inline void mt_seed_host(uint32_t seed, uint32_t out[kMtWords]) {
    out[0] = seed;
    for (int i = 1; i < kMtN; ++i) out[i] = 1812433253u * (out[i - 1] ^ (out[i - 1] >> 30)) + (uint32_t)i;
    out[kMtN] = kMtN;
}

// One 32-bit engine output -> uniform_real_distribution<float>(0,1): float(r) * 2^-32 with the
// libstdc++ clamp to nextafterf(1, 0) when the int->float conversion rounds up to 2^32.
__device__ __forceinline__ float mt_canonical(uint32_t r) {
    float u = __fmul_rn(__uint2float_rn(r), 2.3283064365386963e-10f);
    if (u >= 1.0f) u = __uint_as_float(0x3f7fffffu);
    return u;
}

// Regenerate all 624 words held in shared memory.  The recurrence x[k] <- f(x[k], x[k+1],
// x[k+397]) only ever reaches 227 words back, so three sweeps of <= 227 independent words each
// (reads, barrier, writes, barrier) reproduce the sequential order exactly.
__device__ __forceinline__ void mt_regen_block(uint32_t* x) {
    const int t = threadIdx.x;
    {   // words 0..226 use old x[k+397]
        uint32_t v = 0;
        if (t < 227) v = mt_twist(x[t], x[t + 1], x[t + kMtM]);
        __syncthreads();
        if (t < 227) x[t] = v;
        __syncthreads();
    }
    {   // words 227..453 use new x[k-227]
        uint32_t v = 0;
        const int k = 227 + t;
        if (t < 227) v = mt_twist(x[k], x[k + 1], x[k - 227]);
        __syncthreads();
        if (t < 227) x[k] = v;
        __syncthreads();
    }
    {   // words 454..623 use new x[k-227]; word 623 wraps to the new x[0]
        uint32_t v = 0;
        const int k = 454 + t;
        if (k < kMtN) v = mt_twist(x[k], x[(k + 1) % kMtN], x[k - 227]);
        __syncthreads();
        if (k < kMtN) x[k] = v;
        __syncthreads();
    }
}

// Append the next n engine outputs to out[0..n) and advance the engine.  Single CTA of kMtThreads.
// n is read from *n_dev (clamped to n_bound) so that data-dependent draw counts need no host trip.
static __global__ void __launch_bounds__(kMtThreads)
mt_draw_kernel(uint32_t* __restrict__ engine, const int32_t* __restrict__ n_dev, int n_bound,
               uint32_t* __restrict__ out) {
    __shared__ uint32_t x[kMtN];
    __shared__ int s_pos;
    for (int i = threadIdx.x; i < kMtN; i += kMtThreads) x[i] = engine[i];
    if (threadIdx.x == 0) s_pos = (int)engine[kMtN];
    __syncthreads();
    const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
    int pos = s_pos;
    int produced = 0;
    while (produced < n) {
        if (pos >= kMtN) {
            mt_regen_block(x);
            pos = 0;
        }
        const int avail = min(kMtN - pos, n - produced);
        for (int i = threadIdx.x; i < avail; i += kMtThreads) out[produced + i] = mt_temper(x[pos + i]);
        pos += avail;
        produced += avail;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kMtN; i += kMtThreads) engine[i] = x[i];
    if (threadIdx.x == 0) engine[kMtN] = (uint32_t)pos;
}

}  // namespace sgcn
