// Row-wise pieces of the dense layers around the aggregate: layer norm + activation
// (gcn/layers.py:87-97,130-138,404-412), dropout (layers.py:396,415-433), the loss (gcn/models.py:68-83)
// and the Adam update (models.py:50-51,187).  The matrix products X @ W themselves are plain library
// GEMMs (cuBLAS through the host framework); everything here is HBM-bound elementwise / row-reduction
// work: one warp per row, 128-bit accesses where the layout allows, statistics kept in registers.
#include "common.cuh"

namespace sgcn {

constexpr int kDenseThreads = 256;
constexpr int kLnMaxPerLane = 32;           // rows up to 1024 columns stay in registers

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// y = act(((x - mean) * rsqrt(var + eps)) * scale + offset), mean/var over the row (tf.nn.moments,
// biased variance; tf.nn.batch_normalization, gcn/layers.py:87-97).  scale/offset NULL = ones/zeros.
// stats[r] = {mean, rstd} kept for the backward.
template <int PER>
__global__ void __launch_bounds__(kDenseThreads)
ln_act_fwd_kernel(const float* __restrict__ x, int64_t ld_x, int n, const int32_t* __restrict__ n_dev, int D,
                  const float* __restrict__ scale, const float* __restrict__ offset, float eps, int relu,
                  float* __restrict__ y, int64_t ld_y, float2* __restrict__ stats) {
    const int rows = dev_count(n_dev, n);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * kDenseThreads) >> 5;
    for (int r = (blockIdx.x * kDenseThreads + threadIdx.x) >> 5; r < rows; r += warps) {
        const float* xr = x + (int64_t)r * ld_x;
        float v[PER];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + k * 32;
            v[k] = c < D ? xr[c] : 0.f;
            s += v[k];
        }
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + k * 32;
            const float d = c < D ? v[k] - mean : 0.f;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
        if (stats && lane == 0) stats[r] = make_float2(mean, rstd);
        float* yr = y + (int64_t)r * ld_y;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + k * 32;
            if (c < D) {
                float o = (v[k] - mean) * rstd;
                o = o * (scale ? scale[c] : 1.f) + (offset ? offset[c] : 0.f);
                yr[c] = relu ? fmaxf(o, 0.f) : o;
            }
        }
    }
}

// backward of the above.  g = dy * [y > 0] (relu) ; xhat = (x - mean) rstd ; dxhat = g * scale
//   dx = rstd * (dxhat - mean_c(dxhat) - xhat * mean_c(dxhat * xhat))
//   dscale[c] += sum_r g * xhat ; doffset[c] += sum_r g        (block partials in shared memory, then atomics)
template <int PER>
__global__ void __launch_bounds__(kDenseThreads)
ln_act_bwd_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ y, int64_t ld_y,
                  const float* __restrict__ dy, int64_t ld_dy, int n, const int32_t* __restrict__ n_dev, int D,
                  const float* __restrict__ scale, const float2* __restrict__ stats, int relu,
                  float* __restrict__ dx, int64_t ld_dx, float* __restrict__ dscale, float* __restrict__ doffset) {
    extern __shared__ float s_part[];                 // [2][D] when dscale/doffset are wanted
    const bool want_param = dscale != nullptr || doffset != nullptr;
    if (want_param) {
        for (int i = threadIdx.x; i < 2 * D; i += kDenseThreads) s_part[i] = 0.f;
        __syncthreads();
    }
    const int rows = dev_count(n_dev, n);
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * kDenseThreads) >> 5;
    float ps[PER], po[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { ps[k] = 0.f; po[k] = 0.f; }
    for (int r = (blockIdx.x * kDenseThreads + threadIdx.x) >> 5; r < rows; r += warps) {
        const float2 st = stats[r];
        float xh[PER], dh[PER];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + k * 32;
            xh[k] = 0.f; dh[k] = 0.f;
            if (c < D) {
                float g = dy[(int64_t)r * ld_dy + c];
                if (relu && !(y[(int64_t)r * ld_y + c] > 0.f)) g = 0.f;
                xh[k] = (x[(int64_t)r * ld_x + c] - st.x) * st.y;
                dh[k] = g * (scale ? scale[c] : 1.f);
                ps[k] += g * xh[k];
                po[k] += g;
                s1 += dh[k];
                s2 += dh[k] * xh[k];
            }
        }
        const float m1 = warp_sum(s1) / (float)D, m2 = warp_sum(s2) / (float)D;
        if (dx) {
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                const int c = lane + k * 32;
                if (c < D) dx[(int64_t)r * ld_dx + c] = st.y * (dh[k] - m1 - xh[k] * m2);
            }
        }
    }
    if (want_param) {
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            const int c = lane + k * 32;
            if (c < D) {
                atomicAdd(s_part + c, ps[k]);
                atomicAdd(s_part + D + c, po[k]);
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += kDenseThreads) {
            if (dscale) atomicAdd(dscale + c, s_part[c]);
            if (doffset) atomicAdd(doffset + c, s_part[D + c]);
        }
    }
}

// ---- dropout (tf.nn.dropout: keep with probability keep_prob, scale kept entries by 1/keep_prob) ----
// The reference's TensorFlow RNG cannot be reproduced (TF is not installed): masks come from a
// counter-based generator (Philox-4x32-10, element index as the counter) or are injected by the caller.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

__global__ void __launch_bounds__(kDenseThreads)
dropout_kernel(const float* __restrict__ x, int64_t ld_x, int n, const int32_t* __restrict__ n_dev, int D,
               float keep, uint64_t seed, uint64_t offset, const uint8_t* __restrict__ mask_in,
               uint8_t* __restrict__ mask_out, float* __restrict__ y, int64_t ld_y) {
    const int rows = dev_count(n_dev, n);
    const int64_t total = (int64_t)rows * D;
    const float inv = 1.f / keep;
    for (int64_t t4 = ((int64_t)blockIdx.x * kDenseThreads + threadIdx.x) * 4; t4 < total;
         t4 += (int64_t)gridDim.x * kDenseThreads * 4) {
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (!mask_in) {
            const uint64_t c = offset + (uint64_t)(t4 >> 2);
            rnd = philox4x32(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        }
        const uint32_t rr[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t t = t4 + q;
            if (t >= total) break;
            const int64_t r = t / D;
            const int c = (int)(t - r * D);
            // keep iff u < keep with u = rr * 2^-32 in [0, 1)   (tf: floor(keep + u) == 1  <=>  u >= 1 - keep)
            const bool kept = mask_in ? mask_in[t] != 0 : ((float)rr[q] * 2.3283064365386963e-10f) < keep;
            if (mask_out) mask_out[t] = kept ? 1 : 0;
            y[r * ld_y + c] = kept ? x[r * ld_x + c] * inv : 0.f;
        }
    }
}

// ---- loss: mean over rows of softmax (or sigmoid, multitask) cross entropy with logits -----------
// (gcn/models.py:76-83).  One warp per row; dlogits = d(mean loss)/d(logits) = (p - labels) / n.
__global__ void __launch_bounds__(kDenseThreads)
xent_kernel(const float* __restrict__ logits, int64_t ld_l, const float* __restrict__ labels, int64_t ld_t,
            int n, int C, int sigmoid, float* __restrict__ loss_sum, float* __restrict__ dlogits, int64_t ld_d) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * kDenseThreads) >> 5;
    float local = 0.f;
    for (int r = (blockIdx.x * kDenseThreads + threadIdx.x) >> 5; r < n; r += warps) {
        const float* z = logits + (int64_t)r * ld_l;
        const float* t = labels + (int64_t)r * ld_t;
        if (sigmoid) {
            // max(z,0) - z t + log(1 + exp(-|z|)), averaged over ALL entries (tf.reduce_mean)
            for (int c = lane; c < C; c += 32) {
                const float zz = z[c], tt = t[c];
                local += fmaxf(zz, 0.f) - zz * tt + log1pf(expf(-fabsf(zz)));
                if (dlogits) dlogits[(int64_t)r * ld_d + c] = (1.f / (1.f + expf(-zz)) - tt) / ((float)n * (float)C);
            }
        } else {
            float m = -INFINITY;
            for (int c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
            m = warp_max(m);
            float s = 0.f, tz = 0.f, ts = 0.f;
            for (int c = lane; c < C; c += 32) {
                s += expf(z[c] - m);
                tz += t[c] * (z[c] - m);
                ts += t[c];
            }
            s = warp_sum(s); tz = warp_sum(tz); ts = warp_sum(ts);
            const float lse = logf(s);
            if (lane == 0) local += ts * lse - tz;                 // -sum_c t_c log softmax_c
            if (dlogits)
                for (int c = lane; c < C; c += 32)
                    dlogits[(int64_t)r * ld_d + c] = (ts * expf(z[c] - m) / s - t[c]) / (float)n;
        }
    }
    local = warp_sum(local);
    if (lane == 0 && local != 0.f) atomicAdd(loss_sum, sigmoid ? local / ((float)n * (float)C) : local / (float)n);
}

// ---- Adam (tf.train.AdamOptimizer: epsilon outside the square root, bias correction in the step size) --
//   g' = g + wd * p      (gradient of weight_decay * l2_loss(p), gcn/models.py:68-74; wd = 0 elsewhere)
//   m = b1 m + (1 - b1) g' ; v = b2 v + (1 - b2) g'^2 ; p -= lr_t * m / (sqrt(v) + eps)
__global__ void __launch_bounds__(kDenseThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr_t, float b1, float b2, float eps, float wd) {
    for (int64_t i = (int64_t)blockIdx.x * kDenseThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kDenseThreads) {
        const float gg = g[i] + wd * p[i];
        const float mm = b1 * m[i] + (1.f - b1) * gg;
        const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
        m[i] = mm;
        v[i] = vv;
        p[i] -= lr_t * mm / (sqrtf(vv) + eps);
    }
}

static int rows_grid(int n) {
    const int per_block = kDenseThreads / 32;
    return std::max(1, std::min((n + per_block - 1) / per_block, kNumSMs * 8));
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

int sgcn_ln_act_fwd(const float* x, int64_t ld_x, int32_t n, const int32_t* n_dev, int32_t D,
                    const float* scale, const float* offset, float eps, int32_t relu, float* y,
                    int64_t ld_y, float* stats /* [n, 2] or NULL */, void* stream) {
    SGCN_REQUIRE(n >= 0 && D >= 0, "ln_act_fwd: negative size");
    if (n == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(x && y, "ln_act_fwd: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_y >= D, "ln_act_fwd: row stride smaller than width");
    SGCN_REQUIRE(D <= 32 * kLnMaxPerLane, "ln_act_fwd: rows wider than 1024 columns are not supported");
    SGCN_REQUIRE(!stats || ((uintptr_t)stats & 7) == 0, "ln_act_fwd: stats must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int per = (D + 31) / 32;
#define LN(P) ln_act_fwd_kernel<P><<<rows_grid(n), kDenseThreads, 0, st>>>(x, ld_x, n, n_dev, D, scale, offset, \
                                                                             eps, relu, y, ld_y, (float2*)stats)
    if (per <= 1) LN(1); else if (per <= 2) LN(2); else if (per <= 4) LN(4); else if (per <= 8) LN(8);
    else if (per <= 16) LN(16); else LN(32);
#undef LN
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_ln_act_bwd(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy,
                    int64_t ld_dy, int32_t n, const int32_t* n_dev, int32_t D, const float* scale,
                    const float* stats, int32_t relu, float* dx, int64_t ld_dx, float* dscale,
                    float* doffset, void* stream) {
    SGCN_REQUIRE(n >= 0 && D >= 0, "ln_act_bwd: negative size");
    if (n == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(x && y && dy && stats, "ln_act_bwd: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_y >= D && ld_dy >= D && (!dx || ld_dx >= D),
                 "ln_act_bwd: row stride smaller than width");
    SGCN_REQUIRE(D <= 32 * kLnMaxPerLane, "ln_act_bwd: rows wider than 1024 columns are not supported");
    cudaStream_t st = (cudaStream_t)stream;
    const int per = (D + 31) / 32;
    const size_t dyn = (dscale || doffset) ? sizeof(float) * 2 * (size_t)D : 0;
    const int grid = std::min(rows_grid(n), kNumSMs * 2);
#define LN(P) ln_act_bwd_kernel<P><<<grid, kDenseThreads, dyn, st>>>(x, ld_x, y, ld_y, dy, ld_dy, n, n_dev, D, scale, \
                                                                    (const float2*)stats, relu, dx, ld_dx, dscale, doffset)
    if (per <= 1) LN(1); else if (per <= 2) LN(2); else if (per <= 4) LN(4); else if (per <= 8) LN(8);
    else if (per <= 16) LN(16); else LN(32);
#undef LN
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_dropout(const float* x, int64_t ld_x, int32_t n, const int32_t* n_dev, int32_t D, float keep_prob,
                 uint64_t seed, uint64_t offset, const uint8_t* mask_in, uint8_t* mask_out, float* y,
                 int64_t ld_y, void* stream) {
    SGCN_REQUIRE(n >= 0 && D >= 0, "dropout: negative size");
    SGCN_REQUIRE(keep_prob > 0.f && keep_prob <= 1.f, "dropout: keep_prob must be in (0, 1]");
    if (n == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(x && y, "dropout: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_y >= D, "dropout: row stride smaller than width");
    const int64_t total = (int64_t)n * D;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total / 4 + kDenseThreads - 1) / kDenseThreads,
                                                                kNumSMs * 8));
    dropout_kernel<<<grid, kDenseThreads, 0, (cudaStream_t)stream>>>(x, ld_x, n, n_dev, D, keep_prob, seed, offset,
                                                                    mask_in, mask_out, y, ld_y);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_xent(const float* logits, int64_t ld_l, const float* labels, int64_t ld_t, int32_t n, int32_t C,
              int32_t sigmoid, float* loss /* DEVICE scalar, += mean loss */, float* dlogits, int64_t ld_d,
              void* stream) {
    SGCN_REQUIRE(n >= 0 && C >= 0, "xent: negative size");
    if (n == 0 || C == 0) return SGCN_OK;
    SGCN_REQUIRE(logits && labels && loss, "xent: null pointer");
    SGCN_REQUIRE(ld_l >= C && ld_t >= C && (!dlogits || ld_d >= C), "xent: row stride smaller than width");
    xent_kernel<<<rows_grid(n), kDenseThreads, 0, (cudaStream_t)stream>>>(logits, ld_l, labels, ld_t, n, C, sigmoid,
                                                                         loss, dlogits, ld_d);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, float weight_decay, void* stream) {
    SGCN_REQUIRE(n >= 0, "adam_step: negative size");
    if (n == 0) return SGCN_OK;
    SGCN_REQUIRE(p && g && m && v, "adam_step: null pointer");
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + kDenseThreads - 1) / kDenseThreads, kNumSMs * 8));
    adam_kernel<<<grid, kDenseThreads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, beta1, beta2, eps, weight_decay);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

}  // extern "C"
