// Row movers: dense row gather (history.dense_slice / tf.gather), row scatter-store
// (tf.scatter_update), strided row copy with zero padding, and the CSR row slicer (history.slice).
//
// All of these are pure HBM copies: the roofline is bytes moved / HBM bandwidth.  The kernels use
// a flattened (row, 16-byte vector) index space so that a short list of long rows (n0 ~ 1.5k rows
// of 4.8 KB at Reddit shape) still spreads over every SM, with 4 independent 128-bit loads in
// flight per thread.
#include "common.cuh"
#include "scan.cuh"

namespace sgcn {

constexpr int kRowThreads = 256;
constexpr int kRowUnroll = 4;

__device__ __forceinline__ void stg_stream4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// MODE 0: dst[i] = src[idx[i]]   (gather)
// MODE 1: dst[idx[i]] = src[i]   (scatter-store, idx distinct)
// MODE 2: dst[i] = i < n ? src[i] : 0   for i < n_total (copy + zero pad)
template <int MODE>
__global__ void __launch_bounds__(kRowThreads)
move_rows_vec4_kernel(const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ idx,
                      int n_host, const int32_t* __restrict__ n_dev, int n_total, int c4,
                      float* __restrict__ dst, int64_t ld_dst, unsigned long long* trace,
                      int32_t* done_counter, const ShardMap smap, const int late_trigger) {
    TraceScope ts(trace, MODE == 0 ? TR_GATHER : (MODE == 1 ? TR_UPDATE : TR_PAD));
    // (PDL) a dependent launched with programmatic stream serialization may start its preamble now;
    // it still waits for this whole grid (griddepcontrol.wait) before touching what is written here.
    // late_trigger (the history write-back): only when this block's stores are issued, see g_wb_late_trigger
    if (!late_trigger) asm volatile("griddepcontrol.launch_dependents;");
    // ... and if THIS kernel was launched that way (the write-back behind the full-neighbour mean),
    // nothing below may run before the predecessor grid has finished reading the history table
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // optional: count this launch as "everything stream-ordered before it has finished"
    if (done_counter && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(done_counter, 1);
    const int n = dev_count(n_dev, n_host);
    const int rows = MODE == 2 ? n_total : n;
    const int64_t total = (int64_t)rows * c4;
    const int64_t stride = (int64_t)gridDim.x * kRowThreads;
    for (int64_t base = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; base < total;
         base += stride * kRowUnroll) {
        float4 v[kRowUnroll];
        int64_t off_dst[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
            const int64_t t = base + (int64_t)u * stride;
            off_dst[u] = -1;
            if (t < total) {
                const int r = (int)(t / c4);
                const int c = (int)(t - (int64_t)r * c4) * 4;
                int64_t rs = r, rd = r;
                if (MODE == 0) rs = idx[r];
                if (MODE == 1) rd = idx[r];
                off_dst[u] = rd * ld_dst + c;
                if (MODE == 2 && r >= n) v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                else v[u] = ldg_stream4((MODE == 0 ? shard_row(smap, src, rs, ld_src) : src + rs * ld_src) + c);
            }
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u)
            if (off_dst[u] >= 0) stg_stream4(dst + off_dst[u], v[u]);
    }
    if (late_trigger) asm volatile("griddepcontrol.launch_dependents;");
}

// scalar fallback for widths / strides / pointers that are not 16-byte friendly (e.g. C = 1433)
template <int MODE>
__global__ void __launch_bounds__(kRowThreads)
move_rows_scalar_kernel(const float* __restrict__ src, int64_t ld_src, const int32_t* __restrict__ idx,
                        int n_host, const int32_t* __restrict__ n_dev, int n_total, int C,
                        float* __restrict__ dst, int64_t ld_dst, int32_t* done_counter) {
    if (done_counter && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(done_counter, 1);
    const int n = dev_count(n_dev, n_host);
    const int rows = MODE == 2 ? n_total : n;
    const int64_t total = (int64_t)rows * C;
    const int64_t stride = (int64_t)gridDim.x * kRowThreads;
    for (int64_t t = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; t < total; t += stride) {
        const int r = (int)(t / C);
        const int c = (int)(t - (int64_t)r * C);
        int64_t rs = r, rd = r;
        if (MODE == 0) rs = idx[r];
        if (MODE == 1) rd = idx[r];
        float v = 0.f;
        if (!(MODE == 2 && r >= n)) v = __ldg(src + rs * ld_src + c);
        dst[rd * ld_dst + c] = v;
    }
}

static bool vec4_ok(const void* a, int64_t lda, const void* b, int64_t ldb, int C) {
    return C % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)a & 15) == 0 &&
           ((uintptr_t)b & 15) == 0;
}

template <int MODE>
static int launch_move_rows(const float* src, int64_t ld_src, const int32_t* idx, int n,
                            const int32_t* n_dev, int n_total, int C, float* dst, int64_t ld_dst,
                            cudaStream_t st, int32_t* done_counter = nullptr) {
    const int rows = MODE == 2 ? n_total : n;
    if (rows <= 0 || C <= 0) return SGCN_OK;
    const int max_blocks = kNumSMs * 8;
    if (vec4_ok(src, ld_src, dst, ld_dst, C)) {
        const int c4 = C / 4;
        const int64_t total = (int64_t)rows * c4;
        int blocks = (int)std::min<int64_t>((total + (int64_t)kRowThreads * kRowUnroll - 1) /
                                                ((int64_t)kRowThreads * kRowUnroll), max_blocks);
        if (blocks < 1) blocks = 1;
        SGCN_MATCH_CARVEOUT(move_rows_vec4_kernel<MODE>);
        if (MODE == 1) {      // history write-back: second link of the step's main chain (PDL when enabled)
            SGCN_CUDA(launch_pdl(move_rows_vec4_kernel<MODE>, dim3(blocks), dim3(kRowThreads), 0, st, src, ld_src,
                                 idx, n, n_dev, n_total, c4, dst, ld_dst, g_trace, done_counter, ShardMap{}, g_wb_late_trigger));
        } else {
            move_rows_vec4_kernel<MODE><<<blocks, kRowThreads, 0, st>>>(src, ld_src, idx, n, n_dev,
                                                                       n_total, c4, dst, ld_dst, g_trace,
                                                                       done_counter, MODE == 0 ? t_feat_map : ShardMap{}, 0);
        }
    } else {
        SGCN_REQUIRE(!(MODE == 0 && t_feat_map.world > 1), "sharded features need 16-byte aligned rows");
        const int64_t total = (int64_t)rows * C;
        int blocks = (int)std::min<int64_t>((total + kRowThreads - 1) / kRowThreads, max_blocks);
        move_rows_scalar_kernel<MODE><<<blocks, kRowThreads, 0, st>>>(src, ld_src, idx, n, n_dev,
                                                                     n_total, C, dst, ld_dst, done_counter);
    }
    SGCN_LAUNCHED();
    return SGCN_OK;
}

// up to three independent row jobs in one launch (blockIdx.y picks the job): the step driver
// initialises dX, pre-zeroes the next step's output and gathers the step's input rows with a single
// graph node -- every dependent launch that leaves a step's side branch is worth ~4 us while the
// full-neighbour mean saturates the memory system (profiles/r01_timeline_*.txt).
//   dst[r] = r < n ? src[idx ? idx[r] : r] : 0     for r < n_total   (pad = 1)
//   dst[r] = src[idx ? idx[r] : r]                 for r < n         (pad = 0; rows >= n untouched)
struct PadJob { const float* src; int64_t ld_src; const int32_t* idx; int n; const int32_t* n_dev; int n_total;
                int C; float* dst; int64_t ld_dst; int pad; };
constexpr int kPadJobs = 3;
struct PadJobs { PadJob j[kPadJobs]; ShardMap smap; /* of job 0's gather source (world <= 1: off) */
                 int stream_l2, stream_l2_pct; /* L2 policy of the gathered (read-once) source rows */ };

__global__ void __launch_bounds__(kRowThreads)
pad_jobs_kernel(const PadJobs p, unsigned long long* trace) {
    const PadJob& a = p.j[blockIdx.y];
    const bool sharded_src = blockIdx.y == 0 && p.smap.world > 1;
    TraceScope ts(trace, a.idx ? TR_GATHER : TR_PAD);
    asm volatile("griddepcontrol.launch_dependents;");          // (PDL) see move_rows_vec4_kernel
    if (!a.dst || a.n_total <= 0 || a.C <= 0) return;
    const int n = dev_count(a.n_dev, a.n);
    const int rows = a.pad ? a.n_total : min(n, a.n_total);
    const uint64_t spol = l2_policy(a.idx ? p.stream_l2 : 0, p.stream_l2_pct);
    const bool vec = (a.C & 3) == 0 && (a.ld_dst & 3) == 0 && (((uintptr_t)a.dst) & 15) == 0 &&
                     (a.n == 0 || ((a.ld_src & 3) == 0 && (((uintptr_t)a.src) & 15) == 0));
    const int64_t stride = (int64_t)gridDim.x * kRowThreads;
    if (vec) {
        const int c4 = a.C >> 2;
        const int64_t total = (int64_t)rows * c4;
        for (int64_t base = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; base < total;
             base += stride * kRowUnroll) {
            float4 v[kRowUnroll];
            int64_t off_dst[kRowUnroll];
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u) {
                const int64_t t = base + (int64_t)u * stride;
                off_dst[u] = -1;
                if (t < total) {
                    const int r = (int)(t / c4);
                    const int c = (int)(t - (int64_t)r * c4) * 4;
                    off_dst[u] = (int64_t)r * a.ld_dst + c;
                    if (r < n) {
                        const int64_t rs = a.idx ? (int64_t)a.idx[r] : (int64_t)r;
                        v[u] = ldg_stream4_hint((sharded_src ? shard_row(p.smap, a.src, rs, a.ld_src) : a.src + rs * a.ld_src) + c, spol);
                    } else {
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kRowUnroll; ++u)
                if (off_dst[u] >= 0) stg_stream4(a.dst + off_dst[u], v[u]);
        }
    } else {
        const int64_t total = (int64_t)rows * a.C;
        for (int64_t t = (int64_t)blockIdx.x * kRowThreads + threadIdx.x; t < total; t += stride) {
            const int r = (int)(t / a.C);
            const int c = (int)(t - (int64_t)r * a.C);
            float v = 0.f;
            if (r < n) v = a.src[(a.idx ? (int64_t)a.idx[r] : (int64_t)r) * a.ld_src + c];
            a.dst[(int64_t)r * a.ld_dst + c] = v;
        }
    }
}

static int launch_pad_jobs(const PadJobs& p, int n_jobs, cudaStream_t st) {
    int64_t work = 1;
    for (int k = 0; k < n_jobs; ++k)
        if (p.j[k].dst) work = std::max<int64_t>(work, (int64_t)p.j[k].n_total * p.j[k].C / 4);
    const int64_t per_block = (int64_t)kRowThreads * kRowUnroll;
    static int mult = 0;                     // thread blocks per SM and job (SGCN_PAD_GRID_MULT: A/B, default 4)
    if (mult == 0) {
        const char* e = getenv("SGCN_PAD_GRID_MULT");
        mult = e ? std::max(1, std::min(16, atoi(e))) : 4;
    }
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((work + per_block - 1) / per_block, kNumSMs * mult)),
              (unsigned)n_jobs);
    SGCN_MATCH_CARVEOUT(pad_jobs_kernel);
    pad_jobs_kernel<<<grid, kRowThreads, 0, st>>>(p, g_trace);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

// ---- CSR row slicer -------------------------------------------------------------------------

__global__ void __launch_bounds__(kScanThreads)
slice_row_len_kernel(const int32_t* __restrict__ a_p, const int32_t* __restrict__ r, int n,
                     int32_t* __restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int row = r[i];
        len[i] = a_p[row + 1] - a_p[row];
    }
}

// one warp per sliced row: coalesced copy of values and (local row, column) pairs
__global__ void __launch_bounds__(256)
slice_copy_kernel(const float* __restrict__ a_d, const int32_t* __restrict__ a_i,
                  const int32_t* __restrict__ a_p, const int32_t* __restrict__ r, int n,
                  const int32_t* __restrict__ o_p, float* __restrict__ o_d,
                  int2* __restrict__ o_i2) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const int src = a_p[r[i]];
        const int dst = o_p[i];
        const int len = o_p[i + 1] - dst;
        for (int k = lane; k < len; k += 32) {
            o_d[dst + k] = __ldg(a_d + src + k);
            o_i2[dst + k] = make_int2(i, __ldg(a_i + src + k));
        }
    }
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

int sgcn_gather_rows(const float* src, int64_t ld_src, const int32_t* idx, int32_t n,
                     const int32_t* n_dev, int32_t C, float* dst, int64_t ld_dst, void* stream) {
    SGCN_REQUIRE(n >= 0 && C >= 0, "gather_rows: negative size");
    SGCN_REQUIRE(n == 0 || C == 0 || (src && idx && dst), "gather_rows: null pointer");
    SGCN_REQUIRE(ld_src >= C && ld_dst >= C, "gather_rows: row stride smaller than width");
    return launch_move_rows<0>(src, ld_src, idx, n, n_dev, 0, C, dst, ld_dst, (cudaStream_t)stream);
}

int sgcn_history_update(float* hist, int64_t ld_h, const int32_t* idx, int32_t n,
                        const int32_t* n_dev, const float* rows, int64_t ld_rows, int32_t D,
                        int32_t* done_counter, void* stream) {
    SGCN_REQUIRE(n >= 0 && D >= 0, "history_update: negative size");
    SGCN_REQUIRE(n == 0 || D == 0 || (hist && idx && rows), "history_update: null pointer");
    SGCN_REQUIRE(ld_h >= D && ld_rows >= D, "history_update: row stride smaller than width");
    return launch_move_rows<1>(rows, ld_rows, idx, n, n_dev, 0, D, hist, ld_h, (cudaStream_t)stream,
                               done_counter);
}

int sgcn_copy_rows_pad(const float* src, int64_t ld_src, int32_t n, const int32_t* n_dev,
                       int32_t n_total, int32_t D, float* dst, int64_t ld_dst, void* stream) {
    SGCN_REQUIRE(n >= 0 && n_total >= 0 && D >= 0, "copy_rows_pad: negative size");
    SGCN_REQUIRE(n_total == 0 || D == 0 || dst, "copy_rows_pad: null dst");
    SGCN_REQUIRE(n == 0 || D == 0 || src, "copy_rows_pad: null src");
    SGCN_REQUIRE(ld_dst >= D && (n == 0 || ld_src >= D), "copy_rows_pad: row stride smaller than width");
    if (n == 0) { src = dst; ld_src = ld_dst; }   // never dereferenced: every row is padding
    return launch_move_rows<2>(src, ld_src, nullptr, n, n_dev, n_total, D, dst, ld_dst,
                               (cudaStream_t)stream);
}

int sgcn_copy_rows_pad_pair(const float* src0, int64_t ld_src0, int32_t n0, const int32_t* n0_dev,
                            int32_t n_total0, int32_t D0, float* dst0, int64_t ld_dst0,
                            const float* src1, int64_t ld_src1, int32_t n1, const int32_t* n1_dev,
                            int32_t n_total1, int32_t D1, float* dst1, int64_t ld_dst1, void* stream) {
    SGCN_REQUIRE(n0 >= 0 && n1 >= 0 && n_total0 >= 0 && n_total1 >= 0 && D0 >= 0 && D1 >= 0,
                 "copy_rows_pad_pair: negative size");
    SGCN_REQUIRE((n0 == 0 || src0) && (n1 == 0 || src1), "copy_rows_pad_pair: null src");
    SGCN_REQUIRE((!dst0 || ld_dst0 >= D0) && (!dst1 || ld_dst1 >= D1) && (n0 == 0 || ld_src0 >= D0) &&
                     (n1 == 0 || ld_src1 >= D1), "copy_rows_pad_pair: row stride smaller than width");
    PadJobs p{};
    p.j[0] = PadJob{src0, ld_src0, nullptr, n0, n0_dev, n_total0, D0, dst0, ld_dst0, 1};
    p.j[1] = PadJob{src1, ld_src1, nullptr, n1, n1_dev, n_total1, D1, dst1, ld_dst1, 1};
    return launch_pad_jobs(p, 2, (cudaStream_t)stream);
}

int sgcn_gather_pad_pair(const float* src, int64_t ld_src, const int32_t* idx, int32_t n,
                         const int32_t* n_dev, int32_t C, float* dst, int64_t ld_dst,
                         const float* src0, int64_t ld_src0, int32_t n0, const int32_t* n0_dev,
                         int32_t n_total0, int32_t D0, float* dst0, int64_t ld_dst0,
                         const float* src1, int64_t ld_src1, int32_t n1, const int32_t* n1_dev,
                         int32_t n_total1, int32_t D1, float* dst1, int64_t ld_dst1, void* stream) {
    SGCN_REQUIRE(n >= 0 && C >= 0 && n0 >= 0 && n1 >= 0 && n_total0 >= 0 && n_total1 >= 0 && D0 >= 0 && D1 >= 0,
                 "gather_pad_pair: negative size");
    SGCN_REQUIRE(n == 0 || C == 0 || (src && idx && dst), "gather_pad_pair: null gather pointer");
    SGCN_REQUIRE(ld_src >= C && ld_dst >= C, "gather_pad_pair: row stride smaller than width");
    SGCN_REQUIRE((n0 == 0 || src0) && (n1 == 0 || src1), "gather_pad_pair: null src");
    SGCN_REQUIRE((!dst0 || ld_dst0 >= D0) && (!dst1 || ld_dst1 >= D1) && (n0 == 0 || ld_src0 >= D0) &&
                     (n1 == 0 || ld_src1 >= D1), "gather_pad_pair: row stride smaller than width");
    PadJobs p{};
    p.j[0] = PadJob{src, ld_src, idx, n, n_dev, n, C, n > 0 && C > 0 ? dst : nullptr, ld_dst, 0};
    p.j[1] = PadJob{src0, ld_src0, nullptr, n0, n0_dev, n_total0, D0, dst0, ld_dst0, 1};
    p.j[2] = PadJob{src1, ld_src1, nullptr, n1, n1_dev, n_total1, D1, dst1, ld_dst1, 1};
    if (t_feat_map.world > 1) {
        SGCN_REQUIRE(vec4_ok(src, ld_src, dst, ld_dst, C), "sharded features need 16-byte aligned rows");
        p.smap = t_feat_map;
    }
    p.stream_l2 = g_stream_l2[0];
    p.stream_l2_pct = g_stream_l2[1];
    return launch_pad_jobs(p, 3, (cudaStream_t)stream);
}

int sgcn_csr_slice_indptr(const int32_t* a_p, const int32_t* r, int32_t n, int32_t* o_p,
                          void* stream) {
    SGCN_REQUIRE(n >= 0, "csr_slice_indptr: negative n");
    SGCN_REQUIRE(o_p && (n == 0 || (a_p && r)), "csr_slice_indptr: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        SGCN_CUDA(cudaMemsetAsync(o_p, 0, sizeof(int32_t), st));
        return SGCN_OK;
    }
    slice_row_len_kernel<<<div_up(n, kScanThreads), kScanThreads, 0, st>>>(a_p, r, n, o_p);
    SGCN_LAUNCHED();
    int* tile_sums = nullptr;
    const int n_tiles = div_up(n, kScanTile);
    if (n_tiles > 1) SGCN_CUDA(cudaMallocAsync(&tile_sums, sizeof(int) * n_tiles, st));
    int rc = launch_exclusive_scan(o_p, o_p, nullptr, n, tile_sums, nullptr, true, st);
    if (tile_sums) cudaFreeAsync(tile_sums, st);
    return rc;
}

int sgcn_csr_slice(const float* a_d, const int32_t* a_i, const int32_t* a_p, const int32_t* r,
                   int32_t n, const int32_t* o_p, float* o_d, int32_t* o_i2, void* stream) {
    SGCN_REQUIRE(n >= 0, "csr_slice: negative n");
    if (n == 0) return SGCN_OK;
    SGCN_REQUIRE(a_d && a_i && a_p && r && o_p && o_d && o_i2, "csr_slice: null pointer");
    SGCN_REQUIRE(((uintptr_t)o_i2 & 7) == 0, "csr_slice: o_i2 must be 8-byte aligned");
    const int blocks = std::min(div_up(n, 8), kNumSMs * 8);
    slice_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a_d, a_i, a_p, r, n, o_p, o_d,
                                                               (int2*)o_i2);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

}  // extern "C"
