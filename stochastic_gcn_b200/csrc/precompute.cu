// Pre-processing ("PP") product over the WHOLE graph:  Y = A_hat @ X  with A_hat the normalised
// adjacency in CSR and X the dense node features -- the reference's `train_adj.dot(feats)` /
// `full_adj.dot(feats)` (gcn/utils.py:168-169,321-322), whose result is stacked next to the self
// features as the model input (gcn/models.py:235-239).  Same arithmetic as the per-step aggregate
// (y[r] += w * x[c]) at 10^8 edges x 602 columns: ~250 GB of row-segment reads at Reddit shape,
// re-reading a 561 MB matrix ~445 times, so the kernel lives on L2:
//   * column tiles: one launch per tile of T columns (T*4 bytes of every X row), so that the tile of
//     X (N x T x 4 B) stays L2-resident while the adjacency (8 B per edge) streams through once per
//     tile;
//   * edge-balanced: the edge list is cut into equal contiguous spans, one per warp, independent
//     of the power-law row lengths; a warp finds its first row by one binary search, then walks
//     rows in order;
//   * per 32 edges: (column, weight) loaded coalesced, one edge per lane, and broadcast with
//     shuffles; row segments of 8 edges are fetched together (8 x VPL vector loads in flight per
//     lane) before the FMAs;
//   * partial row sums stay in registers and leave through one vector RED per lane per row segment
//     (rows may straddle spans), into a Y tile the launch zeroed first.
// Vector width follows the layout: [N, 602] rows are 8-byte aligned (float2), padded layouts take
// float4, anything else the scalar path.
#include "common.cuh"

namespace sgcn {

template <typename V> struct PV;
template <> struct PV<float4> {
    static constexpr int W = 4;
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ float4 ld(const float* p) { return __ldg((const float4*)p); }
    static __device__ __forceinline__ void fma(float4& a, float w, const float4& v) { fma4(a, w, v); }
    static __device__ __forceinline__ void red(float* p, const float4& v) { red_add4(p, v); }
};
template <> struct PV<float2> {
    static constexpr int W = 2;
    static __device__ __forceinline__ float2 zero() { return make_float2(0.f, 0.f); }
    static __device__ __forceinline__ float2 ld(const float* p) { return __ldg((const float2*)p); }
    static __device__ __forceinline__ void fma(float2& a, float w, const float2& v) {
        a.x = fmaf(w, v.x, a.x);
        a.y = fmaf(w, v.y, a.y);
    }
    static __device__ __forceinline__ void red(float* p, const float2& v) {
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
    }
};
template <> struct PV<float> {
    static constexpr int W = 1;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void fma(float& a, float w, const float& v) { a = fmaf(w, v, a); }
    static __device__ __forceinline__ void red(float* p, const float& v) { atomicAdd(p, v); }
};

constexpr int kPpThreads = 256;
constexpr int kPpUnroll = 8;          // edges whose row segments are fetched together

struct PpArgs {
    const int32_t* adj_p; const int32_t* adj_i; const float* adj_w; int n_rows; int64_t nnz;
    const float* x; int64_t ld_x; float* y; int64_t ld_y; int T;     // T = columns of this tile
};

template <typename V, int VPL>
__global__ void __launch_bounds__(kPpThreads)
csr_spmm_tile_kernel(const PpArgs a) {
    using P = PV<V>;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kPpThreads + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * kPpThreads) >> 5;
    const int64_t span = ((a.nnz + warps - 1) / warps + 31) & ~(int64_t)31;
    const int64_t e_begin = warp * span;
    const int64_t e_end = min(e_begin + span, a.nnz);
    if (e_begin >= e_end) return;

    int off[VPL];
    bool ok[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        off[k] = (lane + k * 32) * P::W;
        ok[k] = off[k] < a.T;
    }
    // first row of the span: last r with adj_p[r] <= e_begin
    int row;
    {
        int lo = 0, hi = a.n_rows;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(a.adj_p + mid) <= e_begin) lo = mid; else hi = mid;
        }
        row = lo;
    }
    int64_t row_end = __ldg(a.adj_p + row + 1);
    V acc[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) acc[k] = P::zero();
    bool dirty = false;

    auto flush = [&]() {
        if (dirty) {
#pragma unroll
            for (int k = 0; k < VPL; ++k)
                if (ok[k]) P::red(a.y + (int64_t)row * a.ld_y + off[k], acc[k]);
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) acc[k] = P::zero();
        dirty = false;
    };
    auto advance_to = [&](int64_t e) {          // make `row` the row that owns edge e
        while (e >= row_end) {
            flush();
            ++row;
            row_end = __ldg(a.adj_p + row + 1);
        }
    };

    for (int64_t base = e_begin; base < e_end; base += 32) {
        const int64_t mine = base + lane;
        int col = 0;
        float w = 0.f;
        if (mine < e_end) {
            col = __ldg(a.adj_i + mine);
            w = __ldg(a.adj_w + mine);
        }
        const int cnt = (int)min((int64_t)32, e_end - base);
        for (int j0 = 0; j0 < cnt; j0 += kPpUnroll) {
            const int m = min(kPpUnroll, cnt - j0);
            V buf[kPpUnroll][VPL];
            float wj[kPpUnroll];
#pragma unroll
            for (int u = 0; u < kPpUnroll; ++u) {
                const int cj = __shfl_sync(0xffffffffu, col, (j0 + u) & 31);
                wj[u] = __shfl_sync(0xffffffffu, w, (j0 + u) & 31);
                const float* src = a.x + (int64_t)cj * a.ld_x;
#pragma unroll
                for (int k = 0; k < VPL; ++k)
                    buf[u][k] = (u < m && ok[k]) ? P::ld(src + off[k]) : P::zero();
            }
            const int64_t e0 = base + j0;
            advance_to(e0);
            if (e0 + m <= row_end) {                       // the whole group lies in one row
#pragma unroll
                for (int u = 0; u < kPpUnroll; ++u) {
                    if (u < m) {
#pragma unroll
                        for (int k = 0; k < VPL; ++k) P::fma(acc[k], wj[u], buf[u][k]);
                    }
                }
                dirty = true;
            } else {
#pragma unroll
                for (int u = 0; u < kPpUnroll; ++u) {
                    if (u < m) {
                        advance_to(e0 + u);
#pragma unroll
                        for (int k = 0; k < VPL; ++k) P::fma(acc[k], wj[u], buf[u][k]);
                        dirty = true;
                    }
                }
            }
        }
    }
    flush();
}

__global__ void __launch_bounds__(256)
zero_tile_kernel(float* y, int64_t ld_y, int n_rows, int T) {
    const int64_t total = (int64_t)n_rows * T;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t r = t / T;
        y[r * ld_y + (t - r * T)] = 0.f;
    }
}

static int g_pp_tile = 0;     // 0 = auto

template <typename V, int VPL>
static int launch_pp_tile(const PpArgs& a, cudaStream_t st) {
    static int per_sm = 0;
    if (per_sm == 0) {
        SGCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_tile_kernel<V, VPL>,
                                                                kPpThreads, 0));
        if (per_sm < 1) per_sm = 1;
    }
    // more, shorter spans than resident warps (x4): the power-law tail of a span's rows evens out
    const int grid = kNumSMs * per_sm * 4;
    zero_tile_kernel<<<kNumSMs * 8, 256, 0, st>>>(a.y, a.ld_y, a.n_rows, a.T);
    SGCN_LAUNCHED();
    csr_spmm_tile_kernel<V, VPL><<<grid, kPpThreads, 0, st>>>(a);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

int sgcn_csr_spmm(const int32_t* adj_p, const int32_t* adj_i, const float* adj_w, int32_t n_rows,
                  const float* x, int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t tile_cols,
                  void* stream) {
    SGCN_REQUIRE(n_rows >= 0 && D >= 0 && tile_cols >= 0, "csr_spmm: negative size");
    if (n_rows == 0 || D == 0) return SGCN_OK;
    SGCN_REQUIRE(adj_p && x && y, "csr_spmm: null pointer");
    SGCN_REQUIRE(ld_x >= D && ld_y >= D, "csr_spmm: row stride smaller than width");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t ends[2];
    SGCN_CUDA(cudaMemcpyAsync(&ends[0], adj_p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SGCN_CUDA(cudaMemcpyAsync(&ends[1], adj_p + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SGCN_CUDA(cudaStreamSynchronize(st));      // one-off pre-processing call: the edge count sizes the spans
    SGCN_REQUIRE(ends[0] == 0 && ends[1] >= 0, "csr_spmm: adj_p must start at 0");
    const int64_t nnz = ends[1];
    SGCN_REQUIRE(nnz == 0 || (adj_i && adj_w), "csr_spmm: null pointer");
    const auto al = [](const void* p, int64_t ld, int bytes) {
        return ((uintptr_t)p % bytes) == 0 && (ld * 4) % bytes == 0;
    };
    const int vec = (D % 4 == 0 && al(x, ld_x, 16) && al(y, ld_y, 16)) ? 4
                    : (D % 2 == 0 && al(x, ld_x, 8) && al(y, ld_y, 8)) ? 2 : 1;
    int T = tile_cols ? tile_cols : g_pp_tile;
    if (T <= 0) T = 128;                                   // N x 512 B of X per tile: ~L2-sized at Reddit shape
    T = std::max(32 * vec, (T / (32 * vec)) * (32 * vec)); // a whole number of warp-wide vectors
    T = std::min(T, 32 * vec * 4);
    const int vpl = T / (32 * vec);
    for (int c0 = 0; c0 < D; c0 += T) {
        PpArgs a{adj_p, adj_i, adj_w, n_rows, nnz, x + c0, ld_x, y + c0, ld_y, std::min(T, D - c0)};
        int rc;
#define PP(V, N) rc = launch_pp_tile<V, N>(a, st)
        if (vec == 4) { if (vpl == 1) PP(float4, 1); else if (vpl == 2) PP(float4, 2); else if (vpl == 3) PP(float4, 3); else PP(float4, 4); }
        else if (vec == 2) { if (vpl == 1) PP(float2, 1); else if (vpl == 2) PP(float2, 2); else if (vpl == 3) PP(float2, 3); else PP(float2, 4); }
        else { if (vpl == 1) PP(float, 1); else if (vpl == 2) PP(float, 2); else if (vpl == 3) PP(float, 3); else PP(float, 4); }
#undef PP
        if (rc != SGCN_OK) return rc;
    }
    return SGCN_OK;
}

}  // extern "C"
