// Aggregation kernels: sampled-adjacency SpMM forward (plain / CV / CVD), the edge-balanced
// full-neighbour history mean, and the SpMM backward scatter-add.
//
// Mapping.  A dense row of D floats is covered by a GROUP of LPR lanes (8, 16 or 32), each lane
// holding VPL 16-byte vectors, so one group load is one fully coalesced D*4-byte burst
// (D=128 -> one warp, one 512 B row).  Sampled rows are short (<= degree), so the sampled kernels
// are group-per-output-row.  The full-neighbour term is power-law skewed (mean 492, max ~22k
// at Reddit shape), so it is EDGE-balanced: the concatenated neighbour list of the whole output
// field is cut into fixed chunks, one warp per chunk, partial row sums are reduced in registers
// and flushed with one 128-bit RED per lane at row / chunk boundaries.
//
// All kernels are HBM-bound (0.5 flop/B): no tensor cores here by design.
#include "common.cuh"
#include "exchange.cuh"

namespace sgcn {

// ---- vector abstraction (float4 fast path, float fallback for odd widths / alignments) -------
template <typename V> struct VT;
template <> struct VT<float4> {
    static constexpr int W = 4;
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 ld(const float* p) { return __ldg((const float4*)p); }
    static __device__ __forceinline__ float4 ld_stream(const float* p) { return ldg_stream4(p); }
    static __device__ __forceinline__ float4 ld_hist(const float* p, uint64_t pol) { return ld_hist4(p, pol); }
    static __device__ __forceinline__ void st(float* p, float4 v) { *(float4*)p = v; }
    static __device__ __forceinline__ void red(float* p, float4 v) { red_add4(p, v); }
    static __device__ __forceinline__ void fma(float4& a, float w, float4 v) { fma4(a, w, v); }
    static __device__ __forceinline__ float4 sub(float4 a, float4 b) {
        return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    }
    static __device__ __forceinline__ float4 add(float4 a, float4 b) {
        return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
    static __device__ __forceinline__ float4 mul(float4 a, float s) {
        return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
    }
    static __device__ __forceinline__ bool nonzero(float4 a) {
        return a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f;
    }
};
template <> struct VT<float> {
    static constexpr int W = 1;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ld_stream(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ld_hist(const float* p, uint64_t pol) { return ld_hist1(p, pol); }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ void red(float* p, float v) { atomicAdd(p, v); }
    static __device__ __forceinline__ void fma(float& a, float w, float v) { a = fmaf(w, v, a); }
    static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float mul(float a, float s) { return a * s; }
    static __device__ __forceinline__ bool nonzero(float a) { return a != 0.f; }
};

constexpr int kAggThreads = 256;
constexpr int kWbWorkers = 64;         // thread blocks that carry the fused write-back (the last ones to finish)

// counters of the fused write-back (device ints, zero before the first pass of a run: sgcn_wb_counters_reset)
enum { WB_TICKET = 0, WB_FINISHED = 1, WB_SAMPLED = 2, WB_FULL = 3, WB_ERROR = 4, WB_STICKET = 5, WB_COUNTERS = 8 };

enum { MODE_PLAIN = 0, MODE_CV = 1, MODE_CVD = 2, MODE_DET = 3 };

// programmatic dependent launch: no-ops unless the kernel was launched with the attribute
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename V> __device__ __forceinline__ V vsqrt(V a);
template <> __device__ __forceinline__ float vsqrt<float>(float a) { return sqrtf(a); }
template <> __device__ __forceinline__ float4 vsqrt<float4>(float4 a) {
    return make_float4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w));
}
template <typename V> __device__ __forceinline__ V vmulv(V a, V b);
template <> __device__ __forceinline__ float vmulv<float>(float a, float b) { return a * b; }
template <> __device__ __forceinline__ float4 vmulv<float4>(float4 a, float4 b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
// relu(a) + 1e-10 (gcn/layers.py:341) ; gate(g, pre) = pre > 0 ? g : 0 (ReluGrad)
template <typename V> __device__ __forceinline__ V vrelu_eps(V a);
template <> __device__ __forceinline__ float vrelu_eps<float>(float a) { return fmaxf(a, 0.f) + 1e-10f; }
template <> __device__ __forceinline__ float4 vrelu_eps<float4>(float4 a) {
    return make_float4(fmaxf(a.x, 0.f) + 1e-10f, fmaxf(a.y, 0.f) + 1e-10f, fmaxf(a.z, 0.f) + 1e-10f,
                       fmaxf(a.w, 0.f) + 1e-10f);
}
template <typename V> __device__ __forceinline__ V vgate(V g, V pre);
template <> __device__ __forceinline__ float vgate<float>(float g, float pre) { return pre > 0.f ? g : 0.f; }
template <> __device__ __forceinline__ float4 vgate<float4>(float4 g, float4 p) {
    return make_float4(p.x > 0.f ? g.x : 0.f, p.y > 0.f ? g.y : 0.f, p.z > 0.f ? g.z : 0.f,
                       p.w > 0.f ? g.w : 0.f);
}
template <typename V> __device__ __forceinline__ V vdivv(V a, V b);
template <> __device__ __forceinline__ float vdivv<float>(float a, float b) { return a / b; }
template <> __device__ __forceinline__ float4 vdivv<float4>(float4 a, float4 b) {
    return make_float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w);
}

struct SampledArgs {
    const int32_t* rowptr; const int32_t* cols; const float* vals;
    const int32_t* map;      // PLAIN: optional row map of x;  CV/CVD/DET: tgt (global ids into hist)
    const float* scale;      // CVD
    const float* vals2;      // DET: madj weights (medg_w of the sampler, gcn/scheduler.cpp:164)
    int square;              // PLAIN: use vals^2 (tf.square(adj), gcn/layers.py:242)
    float* pre; int64_t ld_pre;   // DET: optional pre-relu values (the backward's gate)
    int n_out; const int32_t* n_out_dev;
    const float* x; int64_t ld_x;      // PLAIN/CV: x ; CVD: h
    const float* mu; int64_t ld_mu;    // CVD
    const float* hist; int64_t ld_h;   // CV/CVD
    int D;                             // width of this column tile, in floats
    float* y; int64_t ld_y;            // PLAIN/CV: y ; CVD: yh
    float* ymu; int64_t ld_ymu;        // CVD
    float* self0; int64_t ld_s0;       // CV: copy of x[:n_out]; CVD: copy of h[:n_out]
    float* self1; int64_t ld_s1;       // CVD: copy of mu[:n_out]
    int accumulate;                    // 0: overwrite y; 1: y += (PLAIN: read-modify-write by the row's
                                       // owner; CV/CVD: 128-bit RED, commutes with full_mean_kernel)
    unsigned long long* trace;
    // optional fused backward of the same sampled adjacency: dx[cols[e]] += vals[e] * bscale[r] * dy[r]
    const float* dy; int64_t ld_dy; float* dx; int64_t ld_dx; const float* bscale;
    // optional: the multi-GPU write-back push rides on this launch as the blocks with blockIdx.y == 1
    // (both only need the gathered input rows; one launch instead of two on the step's side branch)
    int has_push; WbPushArgs push;
    ShardMap hmap;   // row-sharded history (CV / CVD): hist rows are addressed by GLOBAL node id through it
    // optional (sgcn_sampled_done_attach): the last thread block to finish adds 1 to done_ctr[WB_SAMPLED] -- the
    // fused write-back in the tail of the pass's full-neighbour mean waits for it (this kernel reads hist[tgt])
    int32_t* done_ctr;
    int hist_l2, hist_l2_pct;          // L2 eviction policy of the history rows (l2_policy)
};

// every thread of the block calls it on its way out; the last block of the grid publishes "one more sampled pass done"
__device__ __forceinline__ void sampled_done_signal(int32_t* ctr, unsigned long long* trace) {
    if (!ctr) return;
    __syncthreads();                                   // every history row this block reads has been consumed
    if (threadIdx.x == 0) {
        __threadfence();
        const int total = (int)(gridDim.x * gridDim.y);
        if (atomicAdd(ctr + WB_STICKET, 1) == total - 1) {
            ctr[WB_STICKET] = 0;
            __threadfence();
            atomicAdd(ctr + WB_SAMPLED, 1);
            trace_stamp(trace, TR_SAMPLED_END);
        }
    }
}

template <typename V, int LPR, int VPL, int MODE>
__global__ void __launch_bounds__(kAggThreads)
sampled_rows_kernel(const SampledArgs a) {
    using T = VT<V>;
    TraceScope ts(a.trace, TR_SAMPLED);
    grid_dep_wait();      // (PDL) launched early behind the gather that writes x: nothing to do before it
    if (a.has_push && blockIdx.y == 1) {
        wb_pack_body(a.push, blockIdx.x, gridDim.x);
        sampled_done_signal(a.done_ctr, a.trace);
        return;
    }
    const uint64_t hpol = l2_policy(a.hist_l2, a.hist_l2_pct);
    const int n_out = dev_count(a.n_out_dev, a.n_out);
    const int gl = threadIdx.x % LPR;                      // lane within the group
    const int groups = (gridDim.x * kAggThreads) / LPR;
    int off[VPL];
    bool ok[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        off[k] = (gl + k * LPR) * T::W;
        ok[k] = off[k] < a.D;
    }
    for (int r = (blockIdx.x * kAggThreads + threadIdx.x) / LPR; r < n_out; r += groups) {
        const int e0 = a.rowptr[r], e1 = a.rowptr[r + 1];
        V acc[VPL], acc2[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) { acc[k] = T::zero(); acc2[k] = T::zero(); }
#pragma unroll 2
        for (int e = e0; e < e1; ++e) {
            const int c = __ldg(a.cols + e);
            float w = __ldg(a.vals + e);
            if (MODE == MODE_PLAIN) {
                if (a.square) w *= w;
                const int64_t src = a.map ? (int64_t)__ldg(a.map + c) : (int64_t)c;
#pragma unroll
                for (int k = 0; k < VPL; ++k)
                    if (ok[k]) T::fma(acc[k], w, T::ld(a.x + src * a.ld_x + off[k]));
            } else {
                const int64_t t = __ldg(a.map + e);
#pragma unroll
                for (int k = 0; k < VPL; ++k) {
                    if (!ok[k]) continue;
                    const V hv = T::ld_hist(shard_row(a.hmap, a.hist, t, a.ld_h) + off[k], hpol);
                    if (MODE == MODE_CV) {
                        const V xv = T::ld(a.x + (int64_t)c * a.ld_x + off[k]);
                        T::fma(acc[k], w, T::sub(xv, hv));
                    } else if (MODE == MODE_DET) {
                        // adj^2 @ dsigma^2 + 2 madj @ (dsigma * sigma_bar)   (gcn/layers.py:331-339)
                        const V sb = vsqrt<V>(hv);
                        const V ds = T::sub(vsqrt<V>(T::ld(a.x + (int64_t)c * a.ld_x + off[k])), sb);
                        T::fma(acc[k], w * w, vmulv<V>(ds, ds));
                        T::fma(acc[k], 2.f * __ldg(a.vals2 + e), vmulv<V>(ds, sb));
                    } else {
                        const V hh = T::ld(a.x + (int64_t)c * a.ld_x + off[k]);
                        const V mv = T::ld(a.mu + (int64_t)c * a.ld_mu + off[k]);
                        T::fma(acc[k], w, T::sub(mv, hv));     // adj @ (mu - hist[ifield])
                        T::fma(acc2[k], w, T::sub(hh, mv));    // adj @ (h - mu)
                    }
                }
            }
        }
        const float sc = (MODE == MODE_CVD) ? a.scale[r] : 1.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            if (!ok[k]) continue;
            if (MODE == MODE_PLAIN) {
                float* yp = a.y + (int64_t)r * a.ld_y + off[k];
                if (a.accumulate) acc[k] = T::add(acc[k], *(const V*)yp);
                T::st(yp, acc[k]);
            } else if (MODE == MODE_DET) {
                // y holds fadj^2 @ var_history[ffield] already (accumulate) ; the row's owner finishes:
                // var_neighbour = relu(sum) + 1e-10   (gcn/layers.py:341)
                float* yp = a.y + (int64_t)r * a.ld_y + off[k];
                if (a.accumulate) acc[k] = T::add(acc[k], *(const V*)yp);
                if (a.pre) T::st(a.pre + (int64_t)r * a.ld_pre + off[k], acc[k]);
                T::st(yp, vrelu_eps<V>(acc[k]));
                if (a.self0) T::st(a.self0 + (int64_t)r * a.ld_s0 + off[k],
                                   T::ld(a.x + (int64_t)r * a.ld_x + off[k]));
            } else if (MODE == MODE_CV) {
                if (a.accumulate) T::red(a.y + (int64_t)r * a.ld_y + off[k], acc[k]);
                else T::st(a.y + (int64_t)r * a.ld_y + off[k], acc[k]);
                if (a.self0) T::st(a.self0 + (int64_t)r * a.ld_s0 + off[k],
                                   T::ld(a.x + (int64_t)r * a.ld_x + off[k]));
            } else {
                const V yh = T::add(T::mul(acc2[k], sc), acc[k]);
                if (a.accumulate) {
                    T::red(a.ymu + (int64_t)r * a.ld_ymu + off[k], acc[k]);
                    T::red(a.y + (int64_t)r * a.ld_y + off[k], yh);
                } else {
                    T::st(a.ymu + (int64_t)r * a.ld_ymu + off[k], acc[k]);
                    T::st(a.y + (int64_t)r * a.ld_y + off[k], yh);
                }
                if (a.self0) T::st(a.self0 + (int64_t)r * a.ld_s0 + off[k],
                                   T::ld(a.x + (int64_t)r * a.ld_x + off[k]));
                if (a.self1) T::st(a.self1 + (int64_t)r * a.ld_s1 + off[k],
                                   T::ld(a.mu + (int64_t)r * a.ld_mu + off[k]));
            }
        }
        if (a.dx && e1 > e0) {
            // backward of this row while its edge list is hot (dx was initialised by the caller)
            const float s = a.bscale ? a.bscale[r] : 1.f;
            V g[VPL];
#pragma unroll
            for (int k = 0; k < VPL; ++k)
                g[k] = ok[k] ? T::mul(T::ld(a.dy + (int64_t)r * a.ld_dy + off[k]), s) : T::zero();
            for (int e = e0; e < e1; ++e) {
                const int64_t c = __ldg(a.cols + e);
                const float w = __ldg(a.vals + e);
#pragma unroll
                for (int k = 0; k < VPL; ++k)
                    if (ok[k]) T::red(a.dx + c * a.ld_dx + off[k], T::mul(g[k], w));
            }
        }
    }
    sampled_done_signal(a.done_ctr, a.trace);
}

// ---- SpMM backward: dx[cols[e]] += vals[e] * rscale[r] * dy[r] --------------------------------
struct BwdArgs {
    const int32_t* rowptr; const int32_t* cols; const float* vals; const float* rscale;
    int n_out; const int32_t* n_out_dev;
    const float* dy; int64_t ld_dy; int D; float* dx; int64_t ld_dx;
    unsigned long long* trace;
    int square;              // use vals^2 (backward of tf.square(adj) @ var)
};

template <typename V, int LPR, int VPL>
__global__ void __launch_bounds__(kAggThreads)
spmm_bwd_kernel(const BwdArgs a) {
    using T = VT<V>;
    TraceScope ts(a.trace, TR_BWD);
    const int n_out = dev_count(a.n_out_dev, a.n_out);
    const int gl = threadIdx.x % LPR;
    const int groups = (gridDim.x * kAggThreads) / LPR;
    for (int r = (blockIdx.x * kAggThreads + threadIdx.x) / LPR; r < n_out; r += groups) {
        const int e0 = a.rowptr[r], e1 = a.rowptr[r + 1];
        if (e0 == e1) continue;
        const float s = a.rscale ? a.rscale[r] : 1.f;
        V g[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int off = (gl + k * LPR) * T::W;
            g[k] = off < a.D ? T::mul(T::ld(a.dy + (int64_t)r * a.ld_dy + off), s) : T::zero();
        }
        for (int e = e0; e < e1; ++e) {
            const int64_t c = __ldg(a.cols + e);
            float w = __ldg(a.vals + e);
            if (a.square) w *= w;
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int off = (gl + k * LPR) * T::W;
                if (off < a.D) T::red(a.dx + c * a.ld_dx + off, T::mul(g[k], w));
            }
        }
    }
}

// ---- det-dropout variance backward (gcn/layers.py:331-341 under TF autodiff) --------------------
// var_nb = relu(pre) + 1e-10, pre = sum_e w^2 ds_c^2 + 2 mw ds_c sb_c + (history term), with
// ds_c = sqrt(var[c]) - sb_c, sb_c = sqrt(var_history[tgt]).  d var[c] += gate(dy[r], pre[r]) *
// (w^2 ds_c + mw sb_c) / sqrt(var[c])    (= (2 w^2 ds + 2 mw sb) * 0.5 / sigma)
struct DetBwdArgs {
    const int32_t* rowptr; const int32_t* cols; const float* vals; const float* vals2;
    const int32_t* tgt; int n_out; const int32_t* n_out_dev;
    const float* var; int64_t ld_v; const float* hvar; int64_t ld_h; int D;
    const float* dy; int64_t ld_dy; const float* pre; int64_t ld_pre; float* dvar; int64_t ld_dv;
};

template <typename V, int LPR, int VPL>
__global__ void __launch_bounds__(kAggThreads)
det_var_bwd_kernel(const DetBwdArgs a) {
    using T = VT<V>;
    const int n_out = dev_count(a.n_out_dev, a.n_out);
    const int gl = threadIdx.x % LPR;
    const int groups = (gridDim.x * kAggThreads) / LPR;
    for (int r = (blockIdx.x * kAggThreads + threadIdx.x) / LPR; r < n_out; r += groups) {
        const int e0 = a.rowptr[r], e1 = a.rowptr[r + 1];
        if (e0 == e1) continue;
        V g[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int off = (gl + k * LPR) * T::W;
            g[k] = off < a.D ? vgate<V>(T::ld(a.dy + (int64_t)r * a.ld_dy + off),
                                        T::ld(a.pre + (int64_t)r * a.ld_pre + off)) : T::zero();
        }
        for (int e = e0; e < e1; ++e) {
            const int64_t c = __ldg(a.cols + e);
            const int64_t t = __ldg(a.tgt + e);
            const float w = __ldg(a.vals + e);
            const float mw = __ldg(a.vals2 + e);
#pragma unroll
            for (int k = 0; k < VPL; ++k) {
                const int off = (gl + k * LPR) * T::W;
                if (off >= a.D) continue;
                const V sg = vsqrt<V>(T::ld(a.var + c * a.ld_v + off));
                const V sb = vsqrt<V>(T::ld_stream(a.hvar + t * a.ld_h + off));
                V coef = T::mul(T::sub(sg, sb), w * w);
                T::fma(coef, mw, sb);
                T::red(a.dvar + c * a.ld_dv + off, vmulv<V>(g[k], vdivv<V>(coef, sg)));
            }
        }
    }
}

// ---- unsorted COO product with atomics (reference-format adjacency triples) -------------------
template <typename V, int LPR, int VPL>
__global__ void __launch_bounds__(kAggThreads)
spmm_coo_kernel(const int2* __restrict__ idx2, const float* __restrict__ vals, int nnz,
                const float* __restrict__ x, int64_t ld_x, int D, float* __restrict__ y,
                int64_t ld_y, int transpose) {
    using T = VT<V>;
    const int gl = threadIdx.x % LPR;
    const int groups = (gridDim.x * kAggThreads) / LPR;
    for (int e = (blockIdx.x * kAggThreads + threadIdx.x) / LPR; e < nnz; e += groups) {
        const int2 rc = __ldg(idx2 + e);
        const float w = __ldg(vals + e);
        const int64_t r = transpose ? rc.y : rc.x;
        const int64_t c = transpose ? rc.x : rc.y;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int off = (gl + k * LPR) * T::W;
            if (off < D) T::red(y + r * ld_y + off, T::mul(T::ld(x + c * ld_x + off), w));
        }
    }
}

// ---- edge-balanced full-neighbour history mean -------------------------------------------------
// Position p of the concatenated neighbour list belongs to output row r = upper_bound(rowptr_f, p)-1
// and is entry adj[adj_p[nodes[r]] + p - rowptr_f[r]] of the sampler's CSR.
//
// Each warp owns one contiguous span of positions (nnz_f / #warps, rounded to 32) and walks it in
// macro-chunks of 64:
//   stage   the lanes resolve (row, column id, weight) for 64 positions in parallel -- row pointers
//           come from a shared-memory copy made once per CTA, column ids / weights are coalesced
//           loads -- and park {history-row offset, weight, row} in shared memory;
//   stream  the 64 edges are consumed in groups of UN with TWO register buffers: the loads of group
//           i+1 are in flight while group i is multiplied in, so every warp keeps UN..2UN
//           independent 16-byte-per-lane row loads outstanding (the kernel is latency-bound, not
//           issue-bound: bytes in flight per SM are what buys HBM bandwidth);
//   reduce  per-edge operands are warp-uniform shared-memory broadcasts (no shuffles, no
//           divergence bookkeeping); a group that lies inside one output row takes a branch-free
//           path; partial sums stay in registers across the span and leave through one 128-bit RED
//           per lane per row segment.
constexpr int kFullStageRows = 4096;   // row pointers are staged in shared memory up to this many rows
constexpr int kFullMacro = 64;         // positions staged per warp per macro-chunk
constexpr int kFullWarps = kAggThreads / 32;

struct FullArgs {
    const int32_t* nodes; const int32_t* rowptr_f; int n_out; const int32_t* n_out_dev;
    const int32_t* adj_p; const int32_t* adj_i; const float* adj_w;
    const float* hist; int64_t ld_h; int D;
    float* y0; int64_t ld_y0; float* y1; int64_t ld_y1;
    int32_t* work;   // optional: device counter (0 on entry) for dynamic 64-position chunk scheduling
    int stage_rows;  // row pointers of up to this many output rows are staged in shared memory
    unsigned long long* trace;
    int square;      // use adj_w^2 (tf.square(fadj) @ var_history, gcn/layers.py:338)
    // fused write-back (sgcn_full_history_mean_wb): hist[wb_ids[i], :] = wb_rows[i, :] for i < *wb_n_dev, stored by
    // the last thread blocks to finish once EVERY block of this launch has read its history rows and the pass's
    // sampled aggregate has read its own (wb_ctr[WB_SAMPLED] > wb_ctr[WB_FULL]); then *consumed += 1
    const int32_t* wb_ids; const int32_t* wb_n_dev; int wb_bound;
    const float* wb_rows; int64_t ld_wb; int wb_D;
    float* wb_hist;          // == hist of the first column tile, writable
    int32_t* wb_ctr; int32_t* consumed;
    int hist_l2, hist_l2_pct;          // L2 eviction policy of the history rows (l2_policy)
    int late_trigger;                  // (PDL) launch_dependents after this block's positions instead of at entry
    ShardMap hmap;   // row-sharded history (world <= 1: hist is one local table)
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename V, int LPR, int VPL>
__device__ __forceinline__ void full_flush(const FullArgs& a, float* y0, float* y1, int row, int gl, V (&acc)[VPL]) {
    using T = VT<V>;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int off = (gl + k * LPR) * T::W;
        if (off < a.D && T::nonzero(acc[k])) {
            T::red(y0 + (int64_t)row * a.ld_y0 + off, acc[k]);
            if (y1) T::red(y1 + (int64_t)row * a.ld_y1 + off, acc[k]);
        }
        acc[k] = T::zero();
    }
}

// The full-neighbour mean by this thread block (its warps' spans of the pass's positions); the block orders
// itself against its stream predecessor with griddepcontrol (programmatic dependent launch).
template <typename V, int LPR, int VPL>
__device__ __forceinline__ void full_mean_body(const FullArgs& a, const uint64_t hpol) {
    using T = VT<V>;
    const int32_t* const nodes = a.nodes;
    const int32_t* const rowptr_f = a.rowptr_f;
    float* const y0 = a.y0;
    float* const y1 = a.y1;
    constexpr int G = 32 / LPR;                                  // groups per warp
    constexpr int UN = (VPL >= 8) ? 1 : ((VPL == 4) ? 2 : ((VPL == 2) ? 4 : 8));   // row loads per buffer
    constexpr int STEP = G * UN;                                 // positions per group-iteration
    extern __shared__ int32_t s_dyn[];                           // sized to the launch's row bound
    int32_t* s_ptr = s_dyn;                                      // rowptr_f            [stage_rows + 1]
    int32_t* s_base = s_dyn + a.stage_rows + 1;                  // adj_p[nodes[r]] - rowptr_f[r]
    __shared__ int64_t s_off[kFullWarps][kFullMacro];            // adj_i * ld_h (element offset of the row)
    __shared__ float s_w[kFullWarps][kFullMacro];
    __shared__ int32_t s_r[kFullWarps][kFullMacro];
    __shared__ int32_t s_lcol[kFullWarps][kFullMacro];           // landing buffers of the NEXT chunk's metadata
    __shared__ float s_lw[kFullWarps][kFullMacro];
    __shared__ int32_t s_lr[kFullWarps][kFullMacro];
    const int n_out = dev_count(a.n_out_dev, a.n_out);
    if (n_out <= 0) return;
    const bool staged = n_out <= a.stage_rows;
    if (staged) {
        for (int i = threadIdx.x; i <= n_out; i += kAggThreads) s_ptr[i] = __ldg(rowptr_f + i);
        for (int i = threadIdx.x; i < n_out; i += kAggThreads)
            s_base[i] = __ldg(a.adj_p + __ldg(nodes + i)) - __ldg(rowptr_f + i);
        __syncthreads();
    }
    const int32_t* ptr = staged ? s_ptr : rowptr_f;
    const int nnz = ptr[n_out];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gl = lane % LPR, g = lane / LPR;
    const int warp = (blockIdx.x * kAggThreads + threadIdx.x) >> 5;
    const int warps = (gridDim.x * kAggThreads) >> 5;
    // static schedule: one contiguous span per warp.  dynamic schedule (a.work): warps pull
    // 64-position chunks from a device counter, which keeps every SM busy when some SMs are held
    // by a concurrently running kernel (the next batch's sampler in the pipelined step)
    const int span = max(32, (((nnz + warps - 1) / warps) + 7) & ~7);        // every warp of the wave gets positions
    int p0 = warp * span;
    int p1 = min(p0 + span, nnz);
    if (a.work) {
        int c = 0;
        if (lane == 0) c = atomicAdd(a.work, 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        p0 = c * kFullMacro;
        p1 = min(p0 + kFullMacro, nnz);
    }
    if (p0 >= p1) return;
    int64_t* my_off = s_off[wib];
    float* my_w = s_w[wib];
    int32_t* my_r = s_r[wib];

    bool ok[VPL];
    const float* hist_lane[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int off = (gl + k * LPR) * T::W;
        ok[k] = off < a.D;
        hist_lane[k] = a.hist + (ok[k] ? off : 0);
    }
    V acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = T::zero();
    int cur = -1;                                                 // row this group is accumulating

    // Metadata of a macro-chunk (row, column id, weight of 64 positions, two per lane) is FETCHED one
    // macro-chunk ahead: the lanes resolve their positions' rows, then the column ids and weights go
    // global -> shared memory as 4-byte asynchronous copies (cp.async / LDGSTS: no registers held while
    // they fly) into a landing buffer, under the previous chunk's streaming history rows.  It is COMMITTED
    // -- element offsets computed -- right before its own rows are issued.
    // ncu on the round-1 kernel: 22 % of the stall samples sat in this staging when it ran synchronously
    // in front of every chunk (profiles/r01_full_mean_ncu.json).
    int32_t* l_col = s_lcol[wib];
    float* l_w = s_lw[wib];
    int32_t* l_r = s_lr[wib];
    auto fetch = [&](int pm, int pend) {
#pragma unroll
        for (int t = 0; t < kFullMacro / 32; ++t) {
            const int j = t * 32 + lane;
            const int p = pm + j;
            if (p < pend) {
                int lo = 0, hi = n_out;                           // last r with ptr[r] <= p
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (ptr[mid] <= p) lo = mid; else hi = mid;
                }
                l_r[j] = lo;
                const int q = staged ? (s_base[lo] + p)
                                     : (__ldg(a.adj_p + __ldg(nodes + lo)) + (p - ptr[lo]));
                cp_async4(l_col + j, a.adj_i + q);
                cp_async4(l_w + j, a.adj_w + q);
            } else {
                l_r[j] = -1;
                l_col[j] = 0;
                l_w[j] = 0.f;
            }
        }
        cp_async_commit();
    };
    auto commit = [&]() {
        cp_async_wait_all();
        __syncwarp();
#pragma unroll
        for (int t = 0; t < kFullMacro / 32; ++t) {
            const int j = t * 32 + lane;
            const int col = l_col[j];
            const float w = l_w[j];
            const int r = l_r[j];
            int64_t off = (int64_t)col * a.ld_h;
            if (a.hmap.world > 1)          // a remote shard: the offset still is relative to a.hist
                off = (int64_t)(((intptr_t)shard_row(a.hmap, a.hist, col, a.ld_h) - (intptr_t)a.hist) /
                                (intptr_t)sizeof(float));
            my_off[j] = off;
            my_w[j] = a.square ? w * w : w;
            my_r[j] = r;
        }
    };
    fetch(p0, p1);                                                // (PDL) still under the stream predecessor
    grid_dep_wait();                                              // (PDL) history rows: after the write-back

  for (;;) {
    int next_chunk = 0;
    if (a.work) {                                                 // claim the next chunk early
        if (lane == 0) next_chunk = atomicAdd(a.work, 1);
    }
    for (int pm = p0; pm < p1; pm += kFullMacro) {
        __syncwarp();                                             // the previous chunk's rows are consumed
        commit();
        __syncwarp();
        const int cnt = min(kFullMacro, p1 - pm);
        const int ng = (cnt + STEP - 1) / STEP;

        auto issue = [&](V (&buf)[UN][VPL], int gi) {
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int64_t off = my_off[gi * STEP + u * G + g];
#pragma unroll
                for (int k = 0; k < VPL; ++k)
                    buf[u][k] = ok[k] ? T::ld_hist(hist_lane[k] + off, hpol) : T::zero();
            }
        };
        auto consume = [&](V (&buf)[UN][VPL], int gi) {
            const int jf = gi * STEP + g;
            const int rf = my_r[jf], rl = my_r[jf + (UN - 1) * G];
            if (rf == rl && rf >= 0) {                            // whole group inside one output row
                if (rf != cur) {
                    if (cur >= 0) full_flush<V, LPR, VPL>(a, y0, y1, cur, gl, acc);
                    cur = rf;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const float w = my_w[jf + u * G];
#pragma unroll
                    for (int k = 0; k < VPL; ++k) T::fma(acc[k], w, buf[u][k]);
                }
            } else {
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int r = my_r[jf + u * G];
                    if (r < 0) continue;
                    if (r != cur) {
                        if (cur >= 0) full_flush<V, LPR, VPL>(a, y0, y1, cur, gl, acc);
                        cur = r;
                    }
                    const float w = my_w[jf + u * G];
#pragma unroll
                    for (int k = 0; k < VPL; ++k) T::fma(acc[k], w, buf[u][k]);
                }
            }
        };

        // ---- stream: two register buffers, loads of group i+1 in flight while group i is consumed ----
        V bufA[UN][VPL], bufB[UN][VPL];
        issue(bufA, 0);
        if (ng > 1) issue(bufB, 1);
        if (pm + kFullMacro < p1) fetch(pm + kFullMacro, p1);     // next chunk's metadata behind the first rows
        for (int gi = 0; gi < ng; gi += 2) {
            consume(bufA, gi);
            if (gi + 2 < ng) issue(bufA, gi + 2);
            if (gi + 1 < ng) consume(bufB, gi + 1);
            if (gi + 3 < ng) issue(bufB, gi + 3);
        }
    }
    if (!a.work) break;
    next_chunk = __shfl_sync(0xffffffffu, next_chunk, 0);
    p0 = next_chunk * kFullMacro;
    if (p0 >= nnz) break;
    p1 = min(p0 + kFullMacro, nnz);
    fetch(p0, p1);
  }
    if (cur >= 0) full_flush<V, LPR, VPL>(a, y0, y1, cur, gl, acc);
}

// ---- the history write-back fused into the tail of the full-neighbour mean -----------------------------------
// As kernels of their own the write-back and its launch cost ~7 us between two 19 us means (device timelines,
// profiles/r02_timeline_trains_ov0.txt: 4.5 us from the end of the mean to the end of the write-back -- a
// programmatic launch that sits resident in griddepcontrol.wait and keeps the next pass's gather off the SMs --
// and the gather behind it).  Here the LAST kWbWorkers thread blocks to finish (ticket order) wait until every
// block of the launch has consumed its history rows and the pass's sampled aggregate has consumed its own
// (device counters), then store the pass's new rows: hist[wb_ids[i]] = wb_rows[i].  Every read of the table
// precedes every write as in the reference (gcn/models.py:186-194); the next mean follows with a programmatic
// launch and nothing else on the chain.  The blocks that wait are those that finish last anyway; blocks not
// yet resident need no slot of theirs (at most 64 of a few hundred are held), so the wait cannot deadlock;
// it is bounded all the same (wb_ctr[WB_ERROR], sgcn_step_status).
template <typename V>
__device__ __forceinline__ void full_write_back_tail(const FullArgs& a, uint64_t hpol) {
    __shared__ int s_ticket, s_go;
    __syncthreads();                                   // every warp of the block has consumed its history rows
    int32_t* const c = a.wb_ctr;
    const int grid = (int)gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        s_ticket = atomicAdd(c + WB_TICKET, 1);
        if (s_ticket == grid - 1) trace_stamp(a.trace, TR_FULL_BODY_END);
    }
    __syncthreads();
    const int workers = min(grid, kWbWorkers);
    const int wi = s_ticket - (grid - workers);
    if (wi < 0) return;
    // the new rows were gathered a pass ago: every warp loads its (first) rows BEFORE it waits, only the
    // stores are left behind the wait
    const int n = min(*a.wb_n_dev, a.wb_bound);
    const int lane = threadIdx.x & 31;
    constexpr int wpb = kAggThreads / 32;
    constexpr int RU = 4;                              // rows in flight per warp
    const int first = (wi * wpb + (threadIdx.x >> 5)) * RU;
    const bool one_shot = sizeof(V) == 16 && a.wb_D <= 128;      // a row = one float4 per lane
    int id[RU];
    float4 v[RU];
    if (one_shot) {
        const bool in_row = lane * 4 < a.wb_D;
#pragma unroll
        for (int u = 0; u < RU; ++u) id[u] = (first + u < n && in_row) ? a.wb_ids[first + u] : -1;
#pragma unroll
        for (int u = 0; u < RU; ++u)
            if (id[u] >= 0) v[u] = *(const float4*)(a.wb_rows + (int64_t)(first + u) * a.ld_wb + lane * 4);
    }
    if (threadIdx.x == 0) {
        bool go = spin_until_ge(c + WB_TICKET, grid, c + WB_ERROR);
        // sampled passes finished > full passes finished: this pass's sampled aggregate is through
        if (go) go = spin_until_ge(c + WB_SAMPLED, (int)ld_acquire_u32((const unsigned*)(c + WB_FULL)) + 1, c + WB_ERROR);
        s_go = go;
    }
    __syncthreads();
    if (s_go) {
        for (int r0 = first; r0 < n; r0 += workers * wpb * RU) {
            if (one_shot) {
                if (r0 != first) {
                    const bool in_row = lane * 4 < a.wb_D;
#pragma unroll
                    for (int u = 0; u < RU; ++u) id[u] = (r0 + u < n && in_row) ? a.wb_ids[r0 + u] : -1;
#pragma unroll
                    for (int u = 0; u < RU; ++u)
                        if (id[u] >= 0) v[u] = *(const float4*)(a.wb_rows + (int64_t)(r0 + u) * a.ld_wb + lane * 4);
                }
#pragma unroll
                for (int u = 0; u < RU; ++u)
                    if (id[u] >= 0) st_hist4(a.wb_hist + (int64_t)id[u] * a.ld_h + lane * 4, v[u], hpol);
                continue;
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) id[u] = r0 + u < n ? a.wb_ids[r0 + u] : -1;
            if (sizeof(V) == 16) {
                for (int c0 = lane * 4; c0 < a.wb_D; c0 += 128) {
#pragma unroll
                    for (int u = 0; u < RU; ++u)
                        if (id[u] >= 0) v[u] = *(const float4*)(a.wb_rows + (int64_t)(r0 + u) * a.ld_wb + c0);
#pragma unroll
                    for (int u = 0; u < RU; ++u)
                        if (id[u] >= 0) st_hist4(a.wb_hist + (int64_t)id[u] * a.ld_h + c0, v[u], hpol);
                }
            } else {
                for (int c0 = lane; c0 < a.wb_D; c0 += 32)
#pragma unroll
                    for (int u = 0; u < RU; ++u)
                        if (id[u] >= 0) a.wb_hist[(int64_t)id[u] * a.ld_h + c0] = a.wb_rows[(int64_t)(r0 + u) * a.ld_wb + c0];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(c + WB_FINISHED, 1) == workers - 1) {       // the last worker: counters ready for the next pass
            c[WB_TICKET] = 0;
            c[WB_FINISHED] = 0;
            if (a.consumed) atomicAdd(a.consumed, 1);             // the sampler's guard: this pass's reads are over
            __threadfence();
            atomicAdd(c + WB_FULL, 1);
            trace_stamp(a.trace, TR_FULL_TAIL_END);
        }
    }
}

// REGS = 96: 2 CTAs x 256 threads x 96 regs leave 16K registers per SM free for the kernels that run beside it;
// REGS = 80 (sgcn_tune_set SGCN_TUNE_FULL_REGS): 3 CTAs per SM, shorter spans per warp, some spills
template <typename V, int LPR, int VPL, int REGS>
__global__ void __maxnreg__(REGS)
full_mean_kernel(const FullArgs a) {
    TraceScope ts(a.trace, TR_FULL);
    // (PDL) when may the stream successor become resident?  At once (late_trigger = 0: right for a small
    // write-back kernel that slips in beside this one), or when this block has finished its positions: the
    // successor of the fused form is the NEXT full-neighbour mean, whose blocks cannot fit before these leave --
    // and a launch that waits for room holds up every launch behind it, whatever its stream (the pass's sampled
    // aggregate and the next gather started only after this kernel had drained: profiles/r02_timeline_fused_*)
    if (!a.late_trigger) grid_dep_launch();
    const uint64_t hpol = l2_policy(a.hist_l2, a.hist_l2_pct);
    full_mean_body<V, LPR, VPL>(a, hpol);
    if (a.late_trigger) grid_dep_launch();
    if (a.wb_ids) {
        grid_dep_wait();                                         // (blocks without positions skipped it above)
        full_write_back_tail<V>(a, hpol);
    }
}

__global__ void wb_counters_reset_kernel(int32_t* ctr) {
    if (threadIdx.x < WB_COUNTERS) ctr[threadIdx.x] = 0;
}

// ---- full-neighbour history mean, bulk-copy (TMA engine) variant ---------------------------------
// Same arithmetic as full_mean_kernel; different data movement (sgcn_tune_set SGCN_TUNE_FULL_VARIANT
// = 1; NOT the default).  One persistent CTA per SM
//   1. resolves (row, column, weight) of its WHOLE contiguous span of positions in one parallel
//      pass by all threads,
//   2. then every warp walks its own contiguous slice of that span as a private ring of shared-
//      memory stages filled by `cp.async.bulk` row copies (one 16-byte-aligned D*4-byte copy per
//      history row, completion counted on an mbarrier): rows in flight cost no registers; the warp
//      that drains a stage re-arms it itself (no producer warp, no empty barriers), reduces the rows
//      from shared memory with warp-uniform (row, weight) broadcasts and leaves through one 128-bit
//      RED per lane per row segment.
// Measured at Reddit shape (profiles/r01_full_mean_variants_sweep.jsonl): best ring 16 warps x 16
// rows x 1 stage = 25.3 us per launch against 22.4 us for the register variant, and the time falls
// with the number of issuing warps (8 warps 37 us, 12 warps 29 us): a 512-byte row per UBLKCP is too
// small a unit for the copy engine -- ptxas serialises the per-lane issues (ELECT + R2UR + UBLKCP per
// row) and the engine's per-request cost, not bytes in flight, bounds the kernel.  Kept, memcheck-
// clean and parity-tested, as the measured alternative; rows of >= 2 KB would be its regime.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a copy that never lands (bad pointer) traps after ~1 s instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr int kTmaMeta = 2048;        // positions resolved per pass (24 KB of shared memory)
constexpr int kTmaMaxWarps = 16;
constexpr int kTmaMaxDepth = 4;

struct FullTmaCfg { int warps, rows, depth; };   // warps per CTA, rows per stage (<= 32), stages per warp

__global__ void __launch_bounds__(kTmaMaxWarps * 32, 1)
full_mean_tma_kernel(const FullArgs a, const FullTmaCfg cfg) {
    using T = VT<float4>;
    TraceScope ts(a.trace, TR_FULL);
    grid_dep_launch();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int n_out = dev_count(a.n_out_dev, a.n_out);
    if (n_out <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x;
    const uint32_t row_bytes = (uint32_t)a.D * 4u;
    const uint32_t stage_bytes = (uint32_t)cfg.rows * row_bytes;
    const int n_stages = cfg.warps * cfg.depth;
    unsigned char* s_data = smem_raw;
    uint64_t* s_full = (uint64_t*)(smem_raw + (size_t)n_stages * stage_bytes);
    int32_t* m_col = (int32_t*)(s_full + n_stages);
    float* m_w = (float*)(m_col + kTmaMeta);
    int32_t* m_row = (int32_t*)(m_w + kTmaMeta);
    int32_t* s_ptr = m_row + kTmaMeta;                            // rowptr_f            [stage_rows + 1]
    int32_t* s_base = s_ptr + a.stage_rows + 1;                   // adj_p[nodes[r]] - rowptr_f[r]

    for (int i = tid; i <= n_out; i += nthreads) s_ptr[i] = __ldg(a.rowptr_f + i);
    for (int i = tid; i < n_out; i += nthreads)
        s_base[i] = __ldg(a.adj_p + __ldg(a.nodes + i)) - __ldg(a.rowptr_f + i);
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) mbar_init(smem_u32(s_full + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async_smem();
    __syncthreads();
    const int nnz = s_ptr[n_out];
    const int span = max(32, (((nnz + (int)gridDim.x - 1) / (int)gridDim.x) + 31) & ~31);
    const int p0 = blockIdx.x * span;
    const int p1 = min(p0 + span, nnz);
    if (p0 >= p1) return;

    const int gl = lane;
    const bool ok = gl * 4 < a.D;
    float4 acc[1] = {T::zero()};
    int cur = -1;
    uint32_t phase = 0;                                           // bit d: parity to wait for on my stage d
    const uint32_t my_data = smem_u32(s_data) + (uint32_t)(warp * cfg.depth) * stage_bytes;
    const uint32_t my_full = smem_u32(s_full + warp * cfg.depth);
    const unsigned char* my_rows = s_data + (size_t)(warp * cfg.depth) * stage_bytes;

    for (int sc = p0; sc < p1; sc += kTmaMeta) {
        const int cnt = min(kTmaMeta, p1 - sc);
        // ---- pass 1: all threads resolve the metadata of cnt positions ----
        for (int i = tid; i < cnt; i += nthreads) {
            const int p = sc + i;
            int lo = 0, hi = n_out;                               // last r with ptr[r] <= p
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_ptr[mid] <= p) lo = mid; else hi = mid;
            }
            const int q = s_base[lo] + p;
            float w = __ldg(a.adj_w + q);
            if (a.square) w *= w;
            m_col[i] = __ldg(a.adj_i + q);
            m_w[i] = w;
            m_row[i] = lo;
        }
        __syncthreads();
        if (sc == p0) grid_dep_wait();
        // ---- pass 2: each warp streams its contiguous slice through its private ring ----
        const int per = (((cnt + cfg.warps - 1) / cfg.warps) + cfg.rows - 1) / cfg.rows * cfg.rows;
        const int w0 = min(warp * per, cnt), w1 = min(w0 + per, cnt);
        const int nch = (w1 - w0 + cfg.rows - 1) / cfg.rows;
        auto issue = [&](int j) {                                 // chunk j of my slice -> stage j % depth
            const int d = j % cfg.depth;
            const int b = w0 + j * cfg.rows;
            const int rows = min(cfg.rows, w1 - b);
            if (lane == 0) mbar_expect_tx(my_full + 8u * d, (uint32_t)rows * row_bytes);
            __syncwarp();
            if (lane < rows)
                bulk_g2s(my_data + (uint32_t)d * stage_bytes + (uint32_t)lane * row_bytes,
                         a.hist + (int64_t)m_col[b + lane] * a.ld_h, row_bytes, my_full + 8u * d);
        };
        if (warp < cfg.warps) {
            for (int j = 0; j < min(nch, cfg.depth); ++j) issue(j);
            for (int j = 0; j < nch; ++j) {
                const int d = j % cfg.depth;
                const int b = w0 + j * cfg.rows;
                const int rows = min(cfg.rows, w1 - b);
                mbar_wait(my_full + 8u * d, (phase >> d) & 1u);
                phase ^= 1u << d;
                const unsigned char* base = my_rows + (size_t)d * stage_bytes + (size_t)gl * 16;
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    const int row = m_row[b + r];
                    const float w = m_w[b + r];
                    if (row != cur) {
                        if (cur >= 0) full_flush<float4, 32, 1>(a, a.y0, a.y1, cur, gl, acc);
                        cur = row;
                    }
                    if (ok) T::fma(acc[0], w, *(const float4*)(base + (size_t)r * row_bytes));
                }
                if (j + cfg.depth < nch) {
                    __syncwarp();
                    fence_proxy_async_smem();                     // my reads of the stage before its refill
                    issue(j + cfg.depth);
                }
            }
        }
        __syncthreads();                                          // metadata is rewritten by the next pass
    }
    if (cur >= 0) full_flush<float4, 32, 1>(a, a.y0, a.y1, cur, gl, acc);
}

// runtime tunables (sgcn_tune_set): which full-mean variant runs and the shape of its ring
static int g_full_variant = 0;        // 0 = register variant, 1 = bulk-copy variant
static int g_full_regs = 96;          // register cap of full_mean_kernel: 96 (2 CTAs / SM) or 80 (3 CTAs / SM)
static int g_full_trigger = 1;        // (PDL) launch_dependents of full_mean_kernel: 0 at entry, 1 late when fused, 2 late
static FullTmaCfg g_tma_cfg = {12, 16, 2};
static int g_tma_grid = kNumSMs;      // CTAs (one per SM); 147 leaves an SM to a concurrently running sampler

// ---- dispatch -----------------------------------------------------------------------------------

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

struct Shape { bool vec; int lpr, vpl, tile; };

// pick the lane mapping for a row of D floats; tile = floats covered per launch
static Shape pick_shape(int D, bool vec_ok) {
    if (!vec_ok) return {false, 32, 8, 256};
    const int d4 = D / 4;
    if (d4 <= 8) return {true, 8, 1, 32};
    if (d4 <= 16) return {true, 16, 1, 64};
    if (d4 <= 32) return {true, 32, 1, 128};
    if (d4 <= 64) return {true, 32, 2, 256};
    if (d4 <= 128) return {true, 32, 4, 512};
    return {true, 32, 8, 1024};
}

#define SGCN_DISPATCH_SHAPE(shape, CALL)                                         \
    do {                                                                         \
        if (!(shape).vec) { CALL(float, 32, 8); }                                \
        else if ((shape).lpr == 8) { CALL(float4, 8, 1); }                       \
        else if ((shape).lpr == 16) { CALL(float4, 16, 1); }                     \
        else if ((shape).vpl == 1) { CALL(float4, 32, 1); }                      \
        else if ((shape).vpl == 2) { CALL(float4, 32, 2); }                      \
        else if ((shape).vpl == 4) { CALL(float4, 32, 4); }                      \
        else { CALL(float4, 32, 8); }                                            \
    } while (0)

static int grid_for_groups(int64_t n_groups, int lpr) {
    const int per_block = kAggThreads / lpr;
    int64_t blocks = (n_groups + per_block - 1) / per_block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, kNumSMs * 8));
}

struct PendingPush { bool valid = false; WbPushArgs args; };
static thread_local PendingPush t_pending_push;
static thread_local int32_t* t_pending_done = nullptr;      // sgcn_sampled_done_attach

template <int MODE>
static int launch_sampled(SampledArgs a, int D, bool vec_ok, cudaStream_t st) {
    if (a.n_out <= 0 || D <= 0) return SGCN_OK;
    const Shape sh = pick_shape(D, vec_ok);
    if ((MODE == MODE_CV || MODE == MODE_CVD) && t_hist_map.world > 1) {
        SGCN_REQUIRE(D <= sh.tile, "sharded history: the aggregated width must fit one column tile");
        a.hmap = t_hist_map;
    }
    for (int c0 = 0; c0 < D; c0 += sh.tile) {
        SampledArgs t = a;
        t.trace = g_trace;
        t.D = std::min(sh.tile, D - c0);
        t.x = a.x + c0;
        if (a.mu) t.mu = a.mu + c0;
        if (a.hist) t.hist = a.hist + c0;
        t.y = a.y + c0;
        if (a.ymu) t.ymu = a.ymu + c0;
        if (a.pre) t.pre = a.pre + c0;
        if (a.self0) t.self0 = a.self0 + c0;
        if (a.self1) t.self1 = a.self1 + c0;
        if (a.dx) { t.dy = a.dy + c0; t.dx = a.dx + c0; }
        int gx = grid_for_groups(a.n_out, sh.lpr);
        t.hist_l2 = g_hist_l2[0];
        t.hist_l2_pct = g_hist_l2[1];
        t.done_ctr = nullptr;
        if (c0 + sh.tile >= D && t_pending_done) {        // the last column tile's launch carries the signal
            t.done_ctr = t_pending_done;
            t_pending_done = nullptr;
        }
        t.has_push = 0;
        if (c0 == 0 && t_pending_push.valid) {            // see sgcn_wb_push_attach
            t.has_push = 1;
            t.push = t_pending_push.args;
            t_pending_push.valid = false;
            gx = std::max(gx, 64);
        }
        const dim3 grid(gx, t.has_push ? 2 : 1);
#define CALL(V, L, P)                                                       \
    do {                                                                    \
        SGCN_MATCH_CARVEOUT((sampled_rows_kernel<V, L, P, MODE>));          \
        SGCN_CUDA(launch_pdl(sampled_rows_kernel<V, L, P, MODE>, grid, dim3(kAggThreads), 0, st, t)); \
    } while (0)
        SGCN_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

static int spmm_csr_impl(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const int32_t* map, int32_t n_out, const int32_t* n_out_dev, const float* x,
                         int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t accumulate,
                         int square, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "spmm_csr: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && y, "spmm_csr: null pointer");
    SGCN_REQUIRE(ld_y >= D && ld_x >= D, "spmm_csr: row stride smaller than width");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.map = map;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = x; a.ld_x = ld_x;
    a.y = y; a.ld_y = ld_y; a.accumulate = accumulate; a.square = square;
    const bool vec_ok = D % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0 && aligned16(x) && aligned16(y);
    return launch_sampled<MODE_PLAIN>(a, D, vec_ok, (cudaStream_t)stream);
}

int sgcn_spmm_csr(const int32_t* rowptr, const int32_t* cols, const float* vals,
                  const int32_t* map, int32_t n_out, const int32_t* n_out_dev, const float* x,
                  int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t accumulate,
                  void* stream) {
    return spmm_csr_impl(rowptr, cols, vals, map, n_out, n_out_dev, x, ld_x, D, y, ld_y, accumulate, 0, stream);
}

int sgcn_spmm_csr_sq(const int32_t* rowptr, const int32_t* cols, const float* vals,
                     const int32_t* map, int32_t n_out, const int32_t* n_out_dev, const float* x,
                     int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t accumulate,
                     void* stream) {
    return spmm_csr_impl(rowptr, cols, vals, map, n_out, n_out_dev, x, ld_x, D, y, ld_y, accumulate, 1, stream);
}

int sgcn_det_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* mvals, const int32_t* tgt, int32_t n_out,
                         const int32_t* n_out_dev, const float* var, int64_t ld_v,
                         const float* hvar, int64_t ld_h, int32_t D, float* y, int64_t ld_y,
                         float* pre, int64_t ld_pre, float* self, int64_t ld_self,
                         int32_t accumulate, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "det_sampled_fwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && var && hvar && y, "det_sampled_fwd: null pointer");
    SGCN_REQUIRE(ld_v >= D && ld_h >= D && ld_y >= D && (!pre || ld_pre >= D) && (!self || ld_self >= D),
                 "det_sampled_fwd: row stride smaller than width");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.vals2 = mvals; a.map = tgt;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = var; a.ld_x = ld_v; a.hist = hvar; a.ld_h = ld_h;
    a.y = y; a.ld_y = ld_y; a.pre = pre; a.ld_pre = ld_pre; a.self0 = self; a.ld_s0 = ld_self;
    a.accumulate = accumulate;
    const bool vec_ok = D % 4 == 0 && ld_v % 4 == 0 && ld_h % 4 == 0 && ld_y % 4 == 0 &&
                        aligned16(var) && aligned16(hvar) && aligned16(y) &&
                        (!pre || (ld_pre % 4 == 0 && aligned16(pre))) &&
                        (!self || (ld_self % 4 == 0 && aligned16(self)));
    return launch_sampled<MODE_DET>(a, D, vec_ok, (cudaStream_t)stream);
}

int sgcn_det_sampled_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* mvals, const int32_t* tgt, int32_t n_out,
                         const int32_t* n_out_dev, const float* var, int64_t ld_v,
                         const float* hvar, int64_t ld_h, int32_t D, const float* dy, int64_t ld_dy,
                         const float* pre, int64_t ld_pre, float* dvar, int64_t ld_dv, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "det_sampled_bwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && cols && vals && mvals && tgt && var && hvar && dy && pre && dvar,
                 "det_sampled_bwd: null pointer");
    SGCN_REQUIRE(ld_v >= D && ld_h >= D && ld_dy >= D && ld_pre >= D && ld_dv >= D,
                 "det_sampled_bwd: row stride smaller than width");
    const bool vec_ok = D % 4 == 0 && ld_v % 4 == 0 && ld_h % 4 == 0 && ld_dy % 4 == 0 &&
                        ld_pre % 4 == 0 && ld_dv % 4 == 0 && aligned16(var) && aligned16(hvar) &&
                        aligned16(dy) && aligned16(pre) && aligned16(dvar);
    const Shape sh = pick_shape(D, vec_ok);
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < D; c0 += sh.tile) {
        DetBwdArgs a{rowptr, cols, vals, mvals, tgt, n_out, n_out_dev, var + c0, ld_v, hvar + c0, ld_h,
                     std::min(sh.tile, D - c0), dy + c0, ld_dy, pre + c0, ld_pre, dvar + c0, ld_dv};
        const int grid = grid_for_groups(n_out, sh.lpr);
#define CALL(V, L, P) det_var_bwd_kernel<V, L, P><<<grid, kAggThreads, 0, st>>>(a)
        SGCN_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

int sgcn_cv_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                        const int32_t* tgt, int32_t n_out, const int32_t* n_out_dev,
                        const float* x, int64_t ld_x, const float* hist, int64_t ld_h, int32_t D,
                        float* y, int64_t ld_y, float* self, int64_t ld_self, int32_t accumulate,
                        void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "cv_sampled_fwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && x && hist && y, "cv_sampled_fwd: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_h >= D && ld_y >= D && (!self || ld_self >= D),
                 "cv_sampled_fwd: row stride smaller than width");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.map = tgt;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = x; a.ld_x = ld_x; a.hist = hist; a.ld_h = ld_h;
    a.y = y; a.ld_y = ld_y; a.self0 = self; a.ld_s0 = ld_self; a.accumulate = accumulate;
    const bool vec_ok = D % 4 == 0 && ld_x % 4 == 0 && ld_h % 4 == 0 && ld_y % 4 == 0 &&
                        aligned16(x) && aligned16(hist) && aligned16(y) &&
                        (!self || (ld_self % 4 == 0 && aligned16(self)));
    return launch_sampled<MODE_CV>(a, D, vec_ok, (cudaStream_t)stream);
}

int sgcn_cvd_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const int32_t* tgt, const float* scale, int32_t n_out,
                         const int32_t* n_out_dev, const float* h, int64_t ld_hh,
                         const float* mu, int64_t ld_mu, const float* hist, int64_t ld_h,
                         int32_t D, float* yh, int64_t ld_yh, float* ymu, int64_t ld_ymu,
                         float* self_h, int64_t ld_sh, float* self_mu, int64_t ld_sm,
                         int32_t accumulate, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "cvd_sampled_fwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && scale && h && mu && hist && yh && ymu, "cvd_sampled_fwd: null pointer");
    SGCN_REQUIRE(ld_hh >= D && ld_mu >= D && ld_h >= D && ld_yh >= D && ld_ymu >= D &&
                     (!self_h || ld_sh >= D) && (!self_mu || ld_sm >= D),
                 "cvd_sampled_fwd: row stride smaller than width");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.map = tgt; a.scale = scale;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = h; a.ld_x = ld_hh; a.mu = mu; a.ld_mu = ld_mu;
    a.hist = hist; a.ld_h = ld_h; a.y = yh; a.ld_y = ld_yh; a.ymu = ymu; a.ld_ymu = ld_ymu;
    a.self0 = self_h; a.ld_s0 = ld_sh; a.self1 = self_mu; a.ld_s1 = ld_sm; a.accumulate = accumulate;
    const bool vec_ok = D % 4 == 0 && ld_hh % 4 == 0 && ld_mu % 4 == 0 && ld_h % 4 == 0 &&
                        ld_yh % 4 == 0 && ld_ymu % 4 == 0 && aligned16(h) && aligned16(mu) &&
                        aligned16(hist) && aligned16(yh) && aligned16(ymu) &&
                        (!self_h || (ld_sh % 4 == 0 && aligned16(self_h))) &&
                        (!self_mu || (ld_sm % 4 == 0 && aligned16(self_mu)));
    return launch_sampled<MODE_CVD>(a, D, vec_ok, (cudaStream_t)stream);
}

static bool bwd_ok(const float* dy, int64_t ld_dy, const float* dx, int64_t ld_dx, int32_t D) {
    return dy && dx && ld_dy >= D && ld_dx >= D;
}
static bool bwd_vec(const float* dy, int64_t ld_dy, const float* dx, int64_t ld_dx) {
    return ld_dy % 4 == 0 && ld_dx % 4 == 0 && aligned16(dy) && aligned16(dx);
}

int sgcn_cv_sampled_fwd_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                            const int32_t* tgt, int32_t n_out, const int32_t* n_out_dev,
                            const float* x, int64_t ld_x, const float* hist, int64_t ld_h, int32_t D,
                            float* y, int64_t ld_y, float* self, int64_t ld_self, int32_t accumulate,
                            const float* dy, int64_t ld_dy, float* dx, int64_t ld_dx, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "cv_sampled_fwd_bwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && x && hist && y, "cv_sampled_fwd_bwd: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_h >= D && ld_y >= D && (!self || ld_self >= D) && bwd_ok(dy, ld_dy, dx, ld_dx, D),
                 "cv_sampled_fwd_bwd: bad row stride or null gradient buffers");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.map = tgt;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = x; a.ld_x = ld_x; a.hist = hist; a.ld_h = ld_h;
    a.y = y; a.ld_y = ld_y; a.self0 = self; a.ld_s0 = ld_self; a.accumulate = accumulate;
    a.dy = dy; a.ld_dy = ld_dy; a.dx = dx; a.ld_dx = ld_dx;
    const bool vec_ok = D % 4 == 0 && ld_x % 4 == 0 && ld_h % 4 == 0 && ld_y % 4 == 0 &&
                        aligned16(x) && aligned16(hist) && aligned16(y) &&
                        (!self || (ld_self % 4 == 0 && aligned16(self))) && bwd_vec(dy, ld_dy, dx, ld_dx);
    return launch_sampled<MODE_CV>(a, D, vec_ok, (cudaStream_t)stream);
}

int sgcn_cvd_sampled_fwd_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                             const int32_t* tgt, const float* scale, int32_t n_out,
                             const int32_t* n_out_dev, const float* h, int64_t ld_hh,
                             const float* mu, int64_t ld_mu, const float* hist, int64_t ld_h,
                             int32_t D, float* yh, int64_t ld_yh, float* ymu, int64_t ld_ymu,
                             float* self_h, int64_t ld_sh, float* self_mu, int64_t ld_sm,
                             int32_t accumulate, const float* dy, int64_t ld_dy, float* dx,
                             int64_t ld_dx, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "cvd_sampled_fwd_bwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && scale && h && mu && hist && yh && ymu, "cvd_sampled_fwd_bwd: null pointer");
    SGCN_REQUIRE(ld_hh >= D && ld_mu >= D && ld_h >= D && ld_yh >= D && ld_ymu >= D &&
                     (!self_h || ld_sh >= D) && (!self_mu || ld_sm >= D) && bwd_ok(dy, ld_dy, dx, ld_dx, D),
                 "cvd_sampled_fwd_bwd: bad row stride or null gradient buffers");
    SampledArgs a{};
    a.rowptr = rowptr; a.cols = cols; a.vals = vals; a.map = tgt; a.scale = scale;
    a.n_out = n_out; a.n_out_dev = n_out_dev; a.x = h; a.ld_x = ld_hh; a.mu = mu; a.ld_mu = ld_mu;
    a.hist = hist; a.ld_h = ld_h; a.y = yh; a.ld_y = ld_yh; a.ymu = ymu; a.ld_ymu = ld_ymu;
    a.self0 = self_h; a.ld_s0 = ld_sh; a.self1 = self_mu; a.ld_s1 = ld_sm; a.accumulate = accumulate;
    a.dy = dy; a.ld_dy = ld_dy; a.dx = dx; a.ld_dx = ld_dx; a.bscale = scale;
    const bool vec_ok = D % 4 == 0 && ld_hh % 4 == 0 && ld_mu % 4 == 0 && ld_h % 4 == 0 &&
                        ld_yh % 4 == 0 && ld_ymu % 4 == 0 && aligned16(h) && aligned16(mu) &&
                        aligned16(hist) && aligned16(yh) && aligned16(ymu) &&
                        (!self_h || (ld_sh % 4 == 0 && aligned16(self_h))) &&
                        (!self_mu || (ld_sm % 4 == 0 && aligned16(self_mu))) && bwd_vec(dy, ld_dy, dx, ld_dx);
    return launch_sampled<MODE_CVD>(a, D, vec_ok, (cudaStream_t)stream);
}

static int spmm_csr_bwd_impl(const int32_t* rowptr, const int32_t* cols, const float* vals,
                             const float* rscale, int32_t n_out, const int32_t* n_out_dev,
                             const float* dy, int64_t ld_dy, int32_t D, float* dx, int64_t ld_dx,
                             int square, void* stream) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "spmm_csr_bwd: negative size");
    if (n_out == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(rowptr && dy && dx, "spmm_csr_bwd: null pointer");
    SGCN_REQUIRE(ld_dy >= D && ld_dx >= D, "spmm_csr_bwd: row stride smaller than width");
    const bool vec_ok = D % 4 == 0 && ld_dy % 4 == 0 && ld_dx % 4 == 0 && aligned16(dy) && aligned16(dx);
    const Shape sh = pick_shape(D, vec_ok);
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < D; c0 += sh.tile) {
        BwdArgs a{rowptr, cols, vals, rscale, n_out, n_out_dev, dy + c0, ld_dy,
                  std::min(sh.tile, D - c0), dx + c0, ld_dx, g_trace, square};
        const int grid = grid_for_groups(n_out, sh.lpr);
#define CALL(V, L, P)                                                  \
    do {                                                               \
        SGCN_MATCH_CARVEOUT((spmm_bwd_kernel<V, L, P>));               \
        spmm_bwd_kernel<V, L, P><<<grid, kAggThreads, 0, st>>>(a);     \
    } while (0)
        SGCN_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

int sgcn_spmm_csr_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                      const float* rscale, int32_t n_out, const int32_t* n_out_dev,
                      const float* dy, int64_t ld_dy, int32_t D, float* dx, int64_t ld_dx,
                      void* stream) {
    return spmm_csr_bwd_impl(rowptr, cols, vals, rscale, n_out, n_out_dev, dy, ld_dy, D, dx, ld_dx, 0, stream);
}

int sgcn_spmm_csr_bwd_sq(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* rscale, int32_t n_out, const int32_t* n_out_dev,
                         const float* dy, int64_t ld_dy, int32_t D, float* dx, int64_t ld_dx,
                         void* stream) {
    return spmm_csr_bwd_impl(rowptr, cols, vals, rscale, n_out, n_out_dev, dy, ld_dy, D, dx, ld_dx, 1, stream);
}

int sgcn_spmm_coo(const int32_t* idx2, const float* vals, int32_t nnz, const float* x,
                  int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t transpose,
                  void* stream) {
    SGCN_REQUIRE(nnz >= 0 && D >= 0, "spmm_coo: negative size");
    if (nnz == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(idx2 && vals && x && y, "spmm_coo: null pointer");
    SGCN_REQUIRE(((uintptr_t)idx2 & 7) == 0, "spmm_coo: idx2 must be 8-byte aligned");
    SGCN_REQUIRE(ld_x >= D && ld_y >= D, "spmm_coo: row stride smaller than width");
    const bool vec_ok = D % 4 == 0 && ld_x % 4 == 0 && ld_y % 4 == 0 && aligned16(x) && aligned16(y);
    const Shape sh = pick_shape(D, vec_ok);
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < D; c0 += sh.tile) {
        const int Dt = std::min(sh.tile, D - c0);
        const int grid = grid_for_groups(nnz, sh.lpr);
#define CALL(V, L, P)                                                                         \
    spmm_coo_kernel<V, L, P><<<grid, kAggThreads, 0, st>>>((const int2*)idx2, vals, nnz, x + c0, \
                                                           ld_x, Dt, y + c0, ld_y, transpose)
        SGCN_DISPATCH_SHAPE(sh, CALL);
#undef CALL
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

struct FullWb {          // fused write-back of sgcn_full_history_mean_wb
    const int32_t* ids = nullptr; const int32_t* n_dev = nullptr; int bound = 0;
    const float* rows = nullptr; int64_t ld = 0; int32_t* ctr = nullptr; int32_t* consumed = nullptr;
};

static int full_history_mean_impl(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                                  const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                                  const float* adj_w, const float* hist, int64_t ld_h, int32_t D,
                                  float* y0, int64_t ld_y0, float* y1, int64_t ld_y1,
                                  int32_t* work_counter, int square, void* stream, const FullWb wb = FullWb{}) {
    SGCN_REQUIRE(n_out >= 0 && D >= 0, "full_history_mean: negative size");
    if (n_out == 0 || D == 0) {
        SGCN_REQUIRE(!wb.ids, "full_history_mean_wb: empty launch");
        return SGCN_OK;
    }
    SGCN_REQUIRE(nodes && rowptr_f && adj_p && adj_i && adj_w && hist && y0,
                 "full_history_mean: null pointer");
    SGCN_REQUIRE(ld_h >= D && ld_y0 >= D && (!y1 || ld_y1 >= D),
                 "full_history_mean: row stride smaller than width");
    const bool vec_ok = D % 4 == 0 && ld_h % 4 == 0 && ld_y0 % 4 == 0 && aligned16(hist) &&
                        aligned16(y0) && (!y1 || (ld_y1 % 4 == 0 && aligned16(y1))) &&
                        (!wb.ids || (wb.ld % 4 == 0 && aligned16(wb.rows)));
    const Shape sh = pick_shape(D, vec_ok);
    cudaStream_t st = (cudaStream_t)stream;
    if (wb.ids)
        SGCN_REQUIRE(wb.n_dev && wb.rows && wb.ctr && wb.bound > 0 && wb.ld >= D,
                     "full_history_mean_wb: bad write-back arguments");
    const bool sharded = t_hist_map.world > 1;
    if (sharded) SGCN_REQUIRE(D <= sh.tile, "sharded history: the aggregated width must fit one column tile");
    if (g_full_variant == 1 && vec_ok && D <= 128 && n_out <= kFullStageRows && !wb.ids && !sharded) {
        FullTmaCfg cfg = g_tma_cfg;
        FullArgs a{nodes, rowptr_f, n_out, n_out_dev, adj_p, adj_i, adj_w, hist, ld_h, D, y0, ld_y0, y1, ld_y1,
                   nullptr, n_out, g_trace, square, nullptr, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, nullptr,
                   g_hist_l2[0], g_hist_l2[1], 0, ShardMap{}};
        const size_t fixed = 12 * (size_t)kTmaMeta + sizeof(int32_t) * (2 * (size_t)n_out + 2) + 128;
        const size_t stage = (size_t)cfg.rows * D * 4;
        while (cfg.depth > 1 && fixed + (size_t)cfg.warps * cfg.depth * (stage + 8) > 226 * 1024) --cfg.depth;
        const size_t dyn = fixed + (size_t)cfg.warps * cfg.depth * (stage + 8);
        SGCN_REQUIRE(dyn <= 227 * 1024, "full_history_mean: bulk-copy ring does not fit in shared memory");
        static bool attr_set = false;
        if (!attr_set) {
            SGCN_CUDA(cudaFuncSetAttribute(full_mean_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024));
            attr_set = true;
        }
        SGCN_CUDA(launch_pdl(full_mean_tma_kernel, g_tma_grid, cfg.warps * 32, dyn, st, a, cfg));
        SGCN_LAUNCHED();
        return SGCN_OK;
    }
    for (int c0 = 0; c0 < D; c0 += sh.tile) {
        // several column tiles: the write-back (whole rows) rides on the LAST tile's launch
        const bool with_wb = wb.ids && c0 + sh.tile >= D;
        FullArgs a{nodes, rowptr_f, n_out, n_out_dev, adj_p, adj_i, adj_w, hist + c0, ld_h,
                   std::min(sh.tile, D - c0), y0 + c0, ld_y0, y1 ? y1 + c0 : nullptr, ld_y1,
                   D <= sh.tile ? work_counter : nullptr,    // one launch per counter reset
                   std::min(n_out, kFullStageRows), g_trace, square,
                   with_wb ? wb.ids : nullptr, wb.n_dev, wb.bound, wb.rows, wb.ld, D, const_cast<float*>(hist),
                   wb.ctr, wb.consumed, g_hist_l2[0], g_hist_l2[1],
                   g_full_trigger == 2 || (g_full_trigger == 1 && with_wb) ? 1 : 0,
                   sharded ? t_hist_map : ShardMap{}};
        const size_t dyn = sizeof(int32_t) * (2 * (size_t)a.stage_rows + 2);
        // one resident wave: every CTA the SMs can hold at once, spans cut accordingly
#define CALL_R(V, L, P, R)                                                                   \
    do {                                                                                     \
        static int per_sm = 0;                                                               \
        static size_t per_sm_dyn = 0;                                                        \
        if (per_sm == 0 || dyn != per_sm_dyn) {                                              \
            /* ask for a 132 KB shared-memory carve-out although the CTAs need far less: a train-sampler \
               CTA (~47 KB) can then join an SM that runs this kernel without waiting for the SM  \
               to drain and re-partition its L1 / shared memory.  Occupancy is asked for THIS    \
               launch's dynamic shared memory (a worst-case figure halved it: 35 us per launch). */ \
            SGCN_CUDA(cudaFuncSetAttribute(full_mean_kernel<V, L, P, R>,                     \
                                           cudaFuncAttributePreferredSharedMemoryCarveout, kStepCarveout)); \
            SGCN_CUDA(cudaFuncSetAttribute(full_mean_kernel<V, L, P, R>,                     \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
            SGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                         \
                &per_sm, full_mean_kernel<V, L, P, R>, kAggThreads, dyn));                   \
            if (per_sm < 1) per_sm = 1;                                                      \
            per_sm_dyn = dyn;                                                                \
        }                                                                                    \
        SGCN_CUDA(launch_pdl(full_mean_kernel<V, L, P, R>, kNumSMs * per_sm, kAggThreads, dyn, st, a)); \
    } while (0)
#define CALL(V, L, P)                                                                        \
    do {                                                                                     \
        if (g_full_regs == 80) CALL_R(V, L, P, 80); else CALL_R(V, L, P, 96);                \
    } while (0)
        SGCN_DISPATCH_SHAPE(sh, CALL);
#undef CALL
#undef CALL_R
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

int sgcn_wb_counters_reset(int32_t* counters, void* stream) {
    SGCN_REQUIRE(counters, "wb_counters_reset: null pointer");
    wb_counters_reset_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counters);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_sampled_done_attach(int32_t* counters) {
    SGCN_REQUIRE(counters, "sampled_done_attach: null pointer");
    t_pending_done = counters;
    return SGCN_OK;
}

int sgcn_tune_set(int32_t key, int32_t value) {
    switch (key) {
        case SGCN_TUNE_FULL_VARIANT:
            SGCN_REQUIRE(value == 0 || value == 1, "tune: full-mean variant is 0 (register) or 1 (bulk copy)");
            g_full_variant = value; return SGCN_OK;
        case SGCN_TUNE_TMA_WARPS:
            SGCN_REQUIRE(value >= 1 && value <= kTmaMaxWarps, "tune: 1..16 warps");
            g_tma_cfg.warps = value; return SGCN_OK;
        case SGCN_TUNE_TMA_ROWS:
            SGCN_REQUIRE(value >= 1 && value <= 32, "tune: 1..32 rows per stage");
            g_tma_cfg.rows = value; return SGCN_OK;
        case SGCN_TUNE_TMA_DEPTH:
            SGCN_REQUIRE(value >= 1 && value <= kTmaMaxDepth, "tune: 1..4 stages per warp");
            g_tma_cfg.depth = value; return SGCN_OK;
        case SGCN_TUNE_PDL:
            g_pdl = value != 0; return SGCN_OK;
        case SGCN_TUNE_WB_TRIGGER:
            g_wb_late_trigger = value != 0; return SGCN_OK;
        case SGCN_TUNE_FULL_REGS:
            SGCN_REQUIRE(value == 96 || value == 80, "tune: full-mean register cap is 96 or 80");
            g_full_regs = value; return SGCN_OK;
        case SGCN_TUNE_FULL_TRIGGER:
            SGCN_REQUIRE(value >= 0 && value <= 2, "tune: full-mean trigger is 0, 1 or 2");
            g_full_trigger = value; return SGCN_OK;
        case SGCN_TUNE_HIST_L2:
            SGCN_REQUIRE(value >= 0 && value <= 100, "tune: history L2 policy is 0 (normal) or 1..100 (% evict_last)");
            g_hist_l2[0] = value > 0 ? 1 : 0; g_hist_l2[1] = value; return SGCN_OK;
        case SGCN_TUNE_STREAM_L2:
            SGCN_REQUIRE(value >= 0 && value <= 100, "tune: stream L2 policy is 0 (normal) or 1..100 (% evict_first)");
            g_stream_l2[0] = value > 0 ? 2 : 0; g_stream_l2[1] = value; return SGCN_OK;
        case SGCN_TUNE_TMA_GRID:
            SGCN_REQUIRE(value >= 1 && value <= 4 * kNumSMs, "tune: 1..592 CTAs");
            g_tma_grid = value; return SGCN_OK;
        default:
            SGCN_REQUIRE(false, "tune: unknown key");
    }
}

int sgcn_full_history_mean(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                           const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                           const float* adj_w, const float* hist, int64_t ld_h, int32_t D,
                           float* y0, int64_t ld_y0, float* y1, int64_t ld_y1, int32_t* work_counter,
                           void* stream) {
    return full_history_mean_impl(nodes, rowptr_f, n_out, n_out_dev, adj_p, adj_i, adj_w, hist, ld_h, D, y0,
                                  ld_y0, y1, ld_y1, work_counter, 0, stream);
}

int sgcn_full_history_mean_wb(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                              const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                              const float* adj_w, float* hist, int64_t ld_h, int32_t D,
                              float* y0, int64_t ld_y0, float* y1, int64_t ld_y1,
                              const int32_t* wb_ids, const int32_t* wb_n_dev, int32_t wb_bound,
                              const float* wb_rows, int64_t ld_wb, int32_t* counters, int32_t* consumed,
                              void* stream) {
    SGCN_REQUIRE(wb_ids, "full_history_mean_wb: null id list");
    SGCN_REQUIRE(t_hist_map.world <= 1, "full_history_mean_wb: not for row-sharded history tables");
    FullWb wb;
    wb.ids = wb_ids; wb.n_dev = wb_n_dev; wb.bound = wb_bound; wb.rows = wb_rows; wb.ld = ld_wb;
    wb.ctr = counters; wb.consumed = consumed;
    return full_history_mean_impl(nodes, rowptr_f, n_out, n_out_dev, adj_p, adj_i, adj_w, hist, ld_h, D, y0,
                                  ld_y0, y1, ld_y1, nullptr, 0, stream, wb);
}

int sgcn_full_history_mean_sq(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                              const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                              const float* adj_w, const float* hist, int64_t ld_h, int32_t D,
                              float* y0, int64_t ld_y0, float* y1, int64_t ld_y1,
                              int32_t* work_counter, void* stream) {
    return full_history_mean_impl(nodes, rowptr_f, n_out, n_out_dev, adj_p, adj_i, adj_w, hist, ld_h, D, y0,
                                  ld_y0, y1, ld_y1, work_counter, 1, stream);
}

int sgcn_wb_push_attach(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                        int64_t ld_rows, int32_t D, void* const* dst_even, void* const* dst_odd,
                        int32_t n_dst, void* const* peer_flags, int32_t my_rank, int32_t* epoch,
                        int32_t* block_counter) {
    SGCN_REQUIRE(field && n_dev && rows && dst_even && dst_odd && peer_flags && epoch && block_counter,
                 "wb_push_attach: null pointer");
    SGCN_REQUIRE(n_bound >= 0 && D > 0 && ld_rows >= D, "wb_push_attach: bad size");
    SGCN_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers && my_rank >= 0 && my_rank < kMaxPeers,
                 "wb_push_attach: 1..16 ranks");
    WbPushArgs a{};
    a.field = field; a.n_dev = n_dev; a.n_bound = n_bound; a.rows = rows; a.ld_rows = ld_rows; a.D = D;
    for (int k = 0; k < n_dst; ++k) {
        SGCN_REQUIRE(dst_even[k] && dst_odd[k] && peer_flags[k], "wb_push_attach: null peer pointer");
        SGCN_REQUIRE((((uintptr_t)dst_even[k]) & 15) == 0 && (((uintptr_t)dst_odd[k]) & 15) == 0,
                     "wb_push_attach: receive slots must be 16-byte aligned");
        a.dst_even.p[k] = (char*)dst_even[k];
        a.dst_odd.p[k] = (char*)dst_odd[k];
        a.flags.p[k] = (char*)peer_flags[k];
    }
    a.n_dst = n_dst; a.step = 0; a.epoch = epoch; a.my_rank = my_rank; a.block_counter = block_counter;
    t_pending_push.args = a;
    t_pending_push.valid = true;
    return SGCN_OK;
}

}  // extern "C"
