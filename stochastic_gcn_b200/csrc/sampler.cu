// Device-resident neighbour sampler: the B200 replacement of `class Scheduler`
// (gcn/scheduler.h:6-28, gcn/scheduler.cpp:11-189) and `struct Mult` (gcn/mult.h, gcn/mult.cpp).
//
// The reference walks the field sequentially, consuming one mt19937 output per draw, permuting
// the stored rows in place and numbering newly met nodes by first occurrence.  Here the same
// result is produced by data-parallel passes whose only sequential element is the order encoded
// in prefix sums:
//   rows     per field row: degree, sample count, scale                    (row_setup_kernel)
//   scan     sample-count / degree prefix sums = CSR row pointers + each draw's stream offset
//   draws    the next nnz_s engine outputs, block-parallel MT19937          (mt_draw_kernel)
//   shuffle  per row: the <= degree Fisher-Yates swaps (rows are disjoint), edge weights,
//            atomicMin of the edge position into slot[target]               (fisher_yates_kernel)
//   number   first-occurrence flags -> prefix sum = index in the grown field (scan + assign)
// Everything stays in HBM; sizes are carried in a device meta block so that no pass needs a host
// round-trip (buffers are sized from upper bounds: |field| <= n_out (1+degree), nnz_s <= n_out
// degree).  Only the optional reference-format full-neighbour COO (ffield / fedg_*) and the
// importance branch synchronise, because their sizes (sum of full degrees) have no useful bound.
#include <cstdlib>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "mt19937.cuh"
#include "scan.cuh"

namespace sgcn {

constexpr int kUnseen = 0x7fffffff;
constexpr int kExactSizingDegree = 32;   // above this, nnz_s is read back instead of bounded
constexpr int kMetaInts = 8;
enum { M_NOUT = 0, M_NIN = 1, M_NNZS = 2, M_NNZF = 3, M_NFF = 4, M_STATUS = 5, M_AUX = 6 };
enum { ST_DUPLICATE = 1, ST_RANGE = 2, ST_NAN = 4, ST_EMPTY = 8, ST_OVERFLOW = 16 };

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return SGCN_OK;
        size_t want = std::max(bytes, cap + cap / 2);
        want = (want + 255) & ~size_t(255);
        if (p) SGCN_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        SGCN_CUDA(cudaMalloc(&p, want));
        cap = want;
        return SGCN_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T* as() const { return (T*)p; }
};

struct Level {
    int n_out_bound = 0, n_in_bound = 0, s_bound = 0;
    bool done = false, full_materialized = false;
    DevBuf field, rowptr_s, rowptr_f, edg_s, edg_t, tgt, edg_w, medg_w, scales, meta;
    DevBuf ffield, fedg_s, fedg_t, fedg_w;
    void release() {
        for (DevBuf* b : {&field, &rowptr_s, &rowptr_f, &edg_s, &edg_t, &tgt, &edg_w, &medg_w,
                          &scales, &meta, &ffield, &fedg_s, &fedg_t, &fedg_w})
            b->release();
    }
};

}  // namespace sgcn

using namespace sgcn;

struct sgcn_sampler {
    int device = 0;
    int N = 0, E = 0, L = 1;
    bool cv = false, is = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // private mutable CSR + per-node state
    float* adj_w = nullptr;
    int32_t* adj_i = nullptr;
    int32_t* adj_p = nullptr;
    int32_t* slot = nullptr;     // "visited"
    int32_t* fslot = nullptr;    // "fvisited"
    float* importance = nullptr;
    uint32_t* engine = nullptr;  // kMtWords
    uint32_t* engine_is = nullptr;
    // per-batch state, double-buffered: slot p can be filled for batch i+1 while the kernels of
    // batch i still read slot 1-p (cross-step pipelining, sgcn_sampler_set_slot)
    struct Slot {
        DevBuf batch_ids;
        const int32_t* batch_src = nullptr;   // level-0 field: batch_ids (host path) or the caller's buffer
        int batch_n = -1;
        int cur = 0;                          // number of expands since start_batch
        std::vector<Level> levels;
    };
    static constexpr int kSlots = 128;    // buffer sets: 3 for one- / two-batch lookahead, 2 x train for trains
    Slot slots[kSlots];
    int cur_slot = 0;
    Slot& sl() { return slots[cur_slot]; }
    DevBuf batch_meta;
    bool pipeline = false;
    // pipelining guards (device): {number of expands finished, number of consumer passes finished}
    int32_t* pipe_counters = nullptr;
    // trains of batches (sgcn_sampler_expand_train): device copies of the sets' level-0 pointers, the raw
    // engine stream and the control block
    DevBuf train_sets, train_raw, train_ctl;
    int train_n_sets = 0, train_batch = 0, train_degree = 0;
    std::vector<void*> train_field_ptrs;     // host copy of the field pointers (stale-pointer check)
    // scratch
    DevBuf take, deg, draws, rank, tile_sums, pool_mass, tree, hits;
    int32_t* host_meta = nullptr;   // pinned
};

namespace sgcn {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- single-CTA exclusive scan with a carry over tiles -------------------------------------------
template <typename F>
__device__ __forceinline__ int cta_exclusive_scan(int n, F load, int* __restrict__ out) {
    int carry = 0;
    for (int base = 0; base < n; base += kScanTile) {
        const int b = base + threadIdx.x * kScanItems;
        int v[kScanItems];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (b + k < n) ? load(b + k) : 0;
            sum += v[k];
        }
        int total;
        int excl = block_exclusive_scan(sum, &total) + carry;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (b + k < n) out[b + k] = excl;
            excl += v[k];
        }
        carry += total;
    }
    return carry;
}

__global__ void fill_int_kernel(int32_t* p, int64_t n, int32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}
__global__ void fill_float_kernel(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = v;
}

// importance[c] = 1e-6 + sum_r w_rc^2 accumulated in CSR order (scheduler.cpp:22-25).  fp32 sums are
// order dependent, so the entries are first grouped by column WITHOUT changing their relative order
// (stable radix sort of (column, w^2) pairs) and each column is then summed front to back by one
// thread: the per-column order is exactly the reference's, the columns run in parallel.
__global__ void square_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fmul_rn(w[i], w[i]);
}

__global__ void importance_columns_kernel(const int32_t* __restrict__ cols_sorted,
                                          const float* __restrict__ sq_sorted, int E, int N,
                                          float* __restrict__ imp) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    int lo = 0, hi = E;                       // first entry with column >= c
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cols_sorted[mid] < c) lo = mid + 1; else hi = mid;
    }
    float acc = (float)1e-6;
    for (int e = lo; e < E && cols_sorted[e] == c; ++e) acc = __fadd_rn(acc, sq_sorted[e]);
    imp[c] = acc;
}

// ---- pass 1: per-row degree / sample count / scale; old field becomes the prefix -----------------
__global__ void __launch_bounds__(256)
row_setup_kernel(const int32_t* __restrict__ field_in, const int32_t* __restrict__ n_ptr, int n_host,
                 int nb, const int32_t* __restrict__ adj_p, int N, int degree, int want_deg, int want_scales,
                 int32_t* __restrict__ take, int32_t* __restrict__ deg_out,
                 float* __restrict__ scales, int32_t* __restrict__ field_out,
                 int32_t* __restrict__ slot, int32_t* __restrict__ meta) {
    const int n_raw = n_ptr ? *n_ptr : n_host;
    const int n_out = min(n_raw, nb);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {   // meta was zeroed by the memset that precedes this launch
        meta[M_NOUT] = n_out;
        if (n_raw > nb) atomicOr(meta + M_STATUS, ST_OVERFLOW);
    }
    if (i >= n_out) return;
    const int node = field_in[i];
    if (node < 0 || node >= N) {
        atomicOr(meta + M_STATUS, ST_RANGE);
        take[i] = 0;
        deg_out[i] = 0;
        field_out[i] = node;
        if (want_scales) scales[i] = 1.f;
        return;
    }
    const int d = adj_p[node + 1] - adj_p[node];
    const int t = min(d, degree);
    take[i] = t;
    deg_out[i] = want_deg ? d : 0;
    if (want_scales) {
        // scale = (float)deg / take; scales = 1.0 / sqrt(scale) in double  (scheduler.cpp:132-134)
        float scale = (d == 0) ? 1.f : __fdiv_rn((float)d, (float)t);
        scales[i] = (float)(1.0 / (double)__fsqrt_rn(scale));
    }
    field_out[i] = node;
    slot[node] = i;
}

// ---- pass 2: row pointers of the sampled and full-neighbour adjacencies --------------------------
__global__ void __launch_bounds__(kScanThreads)
scan_take_deg_kernel(const int32_t* __restrict__ take, const int32_t* __restrict__ deg,
                     int32_t* __restrict__ rowptr_s, int32_t* __restrict__ rowptr_f,
                     int32_t* __restrict__ meta) {
    const int n = meta[M_NOUT];
    const int tot_s = cta_exclusive_scan(n, [&](int i) { return take[i]; }, rowptr_s);
    const int tot_f = cta_exclusive_scan(n, [&](int i) { return deg[i]; }, rowptr_f);
    if (threadIdx.x == 0) {
        rowptr_s[n] = tot_s;
        rowptr_f[n] = tot_f;
        meta[M_NNZS] = tot_s;
        meta[M_NNZF] = tot_f;
    }
}

// ---- pass 4: per-row partial Fisher-Yates on the stored row + edge emission ----------------------
__global__ void __launch_bounds__(128)
fisher_yates_kernel(const int32_t* __restrict__ field, const int32_t* __restrict__ meta,
                    const int32_t* __restrict__ adj_p, int32_t* __restrict__ adj_i,
                    float* __restrict__ adj_w, const int32_t* __restrict__ rowptr_s,
                    const uint32_t* __restrict__ draws, int s_bound, int cv,
                    int32_t* __restrict__ edg_s, int32_t* __restrict__ tgt,
                    float* __restrict__ edg_w, float* __restrict__ medg_w,
                    int32_t* __restrict__ slot, int32_t* __restrict__ status) {
    const int n_out = meta[M_NOUT];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const int e0 = rowptr_s[i];
    const int take = rowptr_s[i + 1] - e0;
    if (take <= 0) return;
    if (e0 + take > s_bound) {
        atomicOr(status, ST_OVERFLOW);
        return;
    }
    const int node = field[i];
    const int base = adj_p[node];
    const int d = adj_p[node + 1] - base;
    int32_t* rc = adj_i + base;
    float* rw = adj_w + base;
    const float scale = __fdiv_rn((float)d, (float)take);
    for (int k = 0; k < take; ++k) {
        // idx = min((int)(it + num_remaining * u01(generator)), adj_range-1)  (scheduler.cpp:141-143)
        const float u = mt_canonical(draws[e0 + k]);
        const float where = __fadd_rn((float)k, __fmul_rn((float)(d - k), u));
        int j = (int)where;
        if (j > d - 1) j = d - 1;
        const int ck = rc[k], cj = rc[j];
        const float wk = rw[k], wj = rw[j];
        rc[k] = cj; rc[j] = ck;
        rw[k] = wj; rw[j] = wk;
        const int t = (j == k) ? ck : cj;
        const float wv = (j == k) ? wk : wj;
        const float w = __fmul_rn(wv, scale);
        const int e = e0 + k;
        edg_s[e] = i;
        tgt[e] = t;
        edg_w[e] = w;
        if (cv) medg_w[e] = __fmul_rn(wv, w);
        if (slot[t] >= n_out) atomicMin(slot + t, n_out + e);
    }
}

// ---- pass 5: first-occurrence flags -> rank in the grown field ------------------------------------
__global__ void __launch_bounds__(kScanThreads)
scan_first_kernel(const int32_t* __restrict__ tgt, const int32_t* __restrict__ slot, int s_bound,
                  int32_t* __restrict__ rank, int32_t* __restrict__ meta) {
    const int n_out = meta[M_NOUT];
    const int nnz = min(meta[M_NNZS], s_bound);
    const int n_new = cta_exclusive_scan(
        nnz, [&](int e) { return slot[tgt[e]] == n_out + e ? 1 : 0; }, rank);
    if (threadIdx.x == 0) meta[M_NIN] = n_out + n_new;
}

__global__ void __launch_bounds__(256)
assign_kernel(const int32_t* __restrict__ tgt, const int32_t* __restrict__ slot,
              const int32_t* __restrict__ rank, const int32_t* __restrict__ meta, int s_bound,
              int32_t* __restrict__ edg_t, int32_t* __restrict__ field) {
    const int n_out = meta[M_NOUT];
    const int nnz = min(meta[M_NNZS], s_bound);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int t = tgt[e];
    const int s = slot[t];
    if (s < n_out) {
        edg_t[e] = s;
    } else {
        const int e_first = s - n_out;
        const int pos = n_out + rank[e_first];
        edg_t[e] = pos;
        if (e_first == e) field[pos] = t;
    }
}

// visited[s] = -1 for s in field (scheduler.cpp:183-184); also detects duplicate batch ids
__global__ void __launch_bounds__(256)
reset_slots_kernel(const int32_t* __restrict__ field, int32_t* __restrict__ meta, int n_in_bound,
                   int N, int32_t* __restrict__ slot) {
    const int n_out = meta[M_NOUT];
    const int n_in = min(meta[M_NIN], n_in_bound);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_in) return;
    const int node = field[j];
    if (node < 0 || node >= N) return;   // range error already flagged by row_setup_kernel
    if (j < n_out && slot[node] != j) atomicOr(meta + M_STATUS, ST_DUPLICATE);
    slot[node] = kUnseen;
}

// ---- reference-format full-neighbour COO (cv) / neighbour pool (importance) ----------------------
// position p of the concatenated rows -> (row r, target t, weight); atomicMin of p into fslot[t]
__global__ void __launch_bounds__(256)
full_expand_kernel(const int32_t* __restrict__ field, const int32_t* __restrict__ rowptr_f, int n_out,
                   int nnz_f, const int32_t* __restrict__ adj_p, const int32_t* __restrict__ adj_i,
                   const float* __restrict__ adj_w, int32_t* __restrict__ fedg_s,
                   int32_t* __restrict__ fedg_t, float* __restrict__ fedg_w,
                   int32_t* __restrict__ fslot) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz_f) return;
    int lo = 0, hi = n_out;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (rowptr_f[mid] <= p) lo = mid; else hi = mid;
    }
    const int q = adj_p[field[lo]] + (p - rowptr_f[lo]);
    const int t = adj_i[q];
    fedg_s[p] = lo;
    fedg_t[p] = t;          // global id for now; renumbered by full_assign_kernel
    fedg_w[p] = adj_w[q];
    atomicMin(fslot + t, p);
}

__global__ void __launch_bounds__(kScanThreads)
scan_full_first_kernel(const int32_t* __restrict__ fedg_t, const int32_t* __restrict__ fslot,
                       int nnz_f, int32_t* __restrict__ rank, int32_t* __restrict__ meta,
                       int meta_slot) {
    const int n = cta_exclusive_scan(
        nnz_f, [&](int p) { return fslot[fedg_t[p]] == p ? 1 : 0; }, rank);
    if (threadIdx.x == 0) meta[meta_slot] = n;
}

__global__ void __launch_bounds__(256)
full_assign_kernel(int32_t* __restrict__ fedg_t, const int32_t* __restrict__ fslot,
                   const int32_t* __restrict__ rank, int nnz_f, int32_t* __restrict__ ffield) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz_f) return;
    const int t = fedg_t[p];
    const int p_first = fslot[t];
    const int pos = rank[p_first];
    fedg_t[p] = pos;
    if (p_first == p) ffield[pos] = t;
}

__global__ void __launch_bounds__(256)
reset_list_kernel(const int32_t* __restrict__ list, const int32_t* __restrict__ n_ptr, int n_bound,
                  int32_t* __restrict__ table) {
    const int n = min(*n_ptr, n_bound);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) table[list[j]] = kUnseen;
}

// ---- importance branch (scheduler.cpp:63-123 + Mult) ----------------------------------------------
// pool = distinct neighbours in first-seen order (= ffield of the un-permuted rows); mass[j] =
// importance[pool[j]].  The Fenwick array is what Mult::Mult builds by n left-to-right Add()s
// (mult.cpp:7-28): bit[k] is the in-order fp32 sum of prob over (k - lowbit(k), k].
__global__ void __launch_bounds__(256)
pool_mass_kernel(const int32_t* __restrict__ pool, int n_pool, const float* __restrict__ imp,
                 float* __restrict__ mass) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_pool) mass[j] = imp[pool[j]];
}

__global__ void __launch_bounds__(256)
fenwick_build_kernel(const float* __restrict__ mass, int n_pool, int cap, float* __restrict__ tree) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x + 1;   // 1-based node
    if (k > cap) return;
    const int lb = k & (-k);
    float s = 0.f;
    const int hi = min(k, n_pool);
    for (int i = k - lb; i < hi; ++i) s = __fadd_rn(s, mass[i]);
    tree[k] = s;
    if (k == 1) tree[0] = 0.f;
}

// One thread replays the draws in order (each draw removes mass, so the sequence is inherently
// serial): sum = running in-order total, Query() = descent + clamp + Add(-p) (mult.cpp:30-51).
// New targets are numbered by first draw (scheduler.cpp:92-99).
__global__ void importance_draw_kernel(const int32_t* __restrict__ pool, float* __restrict__ mass,
                                       float* __restrict__ tree, int n_pool, int cap, int n_draw,
                                       const uint32_t* __restrict__ draws,
                                       int32_t* __restrict__ hits, int32_t* __restrict__ slot,
                                       int32_t* __restrict__ field, int32_t* __restrict__ meta,
                                       float* __restrict__ mass_total_out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    // sum_importance is accumulated in pool order (scheduler.cpp:74) and Mult::sum by the Adds in
    // the same order (mult.cpp:27): identical sequences, one running sum serves both
    float total = 0.f;
    for (int j = 0; j < n_pool; ++j) total = __fadd_rn(total, mass[j]);
    *mass_total_out = total;
    float sum = total;
    int n_in = meta[M_NOUT];
    for (int d = 0; d < n_draw; ++d) {
        float u = __fmul_rn(mt_canonical(draws[d]), sum);
        int at = 0;
        for (int span = cap; span > 0; span >>= 1) {
            const int nxt = at + span;
            if (nxt <= cap) {
                const float tv = tree[nxt];
                if (!(tv > u)) { u = __fsub_rn(u, tv); at = nxt; }
            }
        }
        int r = at;
        if (r > n_pool - 1) r = n_pool - 1;
        const float delta = -mass[r];
        for (int k = r + 1; k <= cap; k += k & (-k)) tree[k] = __fadd_rn(tree[k], delta);
        sum = __fadd_rn(sum, delta);
        mass[r] = 0.f;
        const int t = pool[r];
        hits[t] += 1;
        if (slot[t] == kUnseen) {
            slot[t] = n_in;
            field[n_in] = t;
            ++n_in;
        }
    }
    meta[M_NIN] = n_in;
}

// per field row: how many stored entries point at a drawn target (warp per row)
__global__ void __launch_bounds__(256)
importance_count_kernel(const int32_t* __restrict__ field, int n_out,
                        const int32_t* __restrict__ adj_p, const int32_t* __restrict__ adj_i,
                        const int32_t* __restrict__ hits, int32_t* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_out; i += warps) {
        const int node = field[i];
        const int b = adj_p[node], e = adj_p[node + 1];
        int c = 0;
        for (int q = b + lane; q < e; q += 32) c += hits[adj_i[q]] > 0 ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) count[i] = c;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_counts_kernel(const int32_t* __restrict__ count, int n, int32_t* __restrict__ rowptr,
                   int32_t* __restrict__ meta) {
    const int tot = cta_exclusive_scan(n, [&](int i) { return count[i]; }, rowptr);
    if (threadIdx.x == 0) {
        rowptr[n] = tot;
        meta[M_NNZS] = tot;
    }
}

// w = times * w * sum_importance / (importance[t] * num_samples), evaluated left to right in fp32
// (scheduler.cpp:108-113); NaN raises the reference's runtime_error("nan") -> status
__global__ void __launch_bounds__(256)
importance_emit_kernel(const int32_t* __restrict__ field, int n_out,
                       const int32_t* __restrict__ adj_p, const int32_t* __restrict__ adj_i,
                       const float* __restrict__ adj_w, const int32_t* __restrict__ hits,
                       const int32_t* __restrict__ slot, const float* __restrict__ imp,
                       const float* __restrict__ mass_total, int n_draw,
                       const int32_t* __restrict__ rowptr, int32_t* __restrict__ edg_s,
                       int32_t* __restrict__ edg_t, int32_t* __restrict__ tgt,
                       float* __restrict__ edg_w, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float total = *mass_total;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_out; i += warps) {
        const int node = field[i];
        const int b = adj_p[node], e = adj_p[node + 1];
        int out = rowptr[i];
        for (int q0 = b; q0 < e; q0 += 32) {
            const int q = q0 + lane;
            int t = -1, h = 0;
            if (q < e) {
                t = adj_i[q];
                h = hits[t];
            }
            const unsigned m = __ballot_sync(0xffffffffu, h > 0);
            if (h > 0) {
                const int pos = out + __popc(m & ((1u << lane) - 1u));
                float num = __fmul_rn((float)h, adj_w[q]);
                num = __fmul_rn(num, total);
                const float den = __fmul_rn(imp[t], (float)n_draw);
                const float w = __fdiv_rn(num, den);
                edg_s[pos] = i;
                edg_t[pos] = slot[t];
                tgt[pos] = t;
                edg_w[pos] = w;
                if (w != w) atomicOr(status, ST_NAN);
            }
            out += __popc(m);
        }
    }
}

__global__ void __launch_bounds__(256)
reset_hits_kernel(const int32_t* __restrict__ pool, int n_pool, int32_t* __restrict__ hits) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_pool) hits[pool[j]] = 0;
}

__global__ void set_meta_kernel(int32_t* meta, int n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < kMetaInts; ++k) meta[k] = 0;
        meta[M_NOUT] = n;
        meta[M_NIN] = n;
    }
}


// ---- fused single-CTA expand (uniform branch) ------------------------------------------------------
// At training batch sizes (B = 512, degree <= 2) every pass above is a few hundred threads of
// work and the chain is pure launch + dependent-load latency.  One CTA of 1024 threads runs the
// whole expand with block barriers in place of kernel boundaries: rows -> scans -> MT19937 draws
// (into shared memory) -> Fisher-Yates -> first-occurrence numbering.  The visited[] table of the
// reference becomes an open-addressing hash table in SHARED memory (node id -> smallest position
// that claimed it), so the numbering needs no global round-trips and leaves nothing to reset; the
// only global traffic on the critical path is ids -> row pointers -> the touched row entries.
constexpr int kFusedThreads = 1024;
constexpr int kFusedMaxRows = 4096;
constexpr int kFusedMaxEdges = 8192;
constexpr int kFusedHashMax = 16384;                       // >= (rows + edges) / 0.75
constexpr size_t kFusedSmemBytes =
    sizeof(int32_t) * ((size_t)kFusedMaxRows + 1 + 2 * (size_t)kFusedMaxEdges + 2 * (size_t)kFusedHashMax +
                       kMtN + 64);   // worst case, the opt-in limit set on the kernel

static int fused_hash_bits(int nb, int sb) {
    int hbits = 10;   // table holds the old field + one entry per sampled edge at load <= 0.75
    while ((1 << hbits) < 2 * (nb + sb) && (1 << hbits) < kFusedHashMax) ++hbits;
    return hbits;
}
static size_t fused_smem_bytes(int nb, int sb, int hbits) {
    return sizeof(int32_t) * ((size_t)nb + 1 + 2 * (size_t)sb + 2 * ((size_t)1 << hbits) + kMtN + 64);
}

struct FusedArgs {
    const int32_t* field_in; const int32_t* n_ptr; int n_host, nb, sb;
    const int32_t* adj_p; int32_t* adj_i; float* adj_w; int N, degree, cv;
    int hbits;               // log2 of the shared-memory hash table size (>= 2 (nb + sb) entries)
    const int32_t* prev_ids; int prev_n;   // pipelining: the batches whose consumer passes may still run
    const int32_t* prev_ids2; int prev_n2;
    int32_t* pipe;                         // {expands finished, consumer passes finished} or NULL
    unsigned long long* trace;
    uint32_t* engine;
    int32_t* field; int32_t* rowptr_s; int32_t* rowptr_f; int32_t* edg_s; int32_t* edg_t;
    int32_t* tgt; float* edg_w; float* medg_w; float* scales; int32_t* meta;
};

template <int NT>
__device__ __forceinline__ int block_scan_excl_nt(int v, int* total, int* s_warp /*33 ints*/) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < NW ? s_warp[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        if (lane < NW) s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const int excl = inc - v + s_warp[warp];
    *total = s_warp[32];
    __syncthreads();
    return excl;
}

// find-or-insert `node` in the shared-memory table; returns its slot
__device__ __forceinline__ int hash_claim(int32_t* keys, int mask, int shift, int node) {
    unsigned h = ((unsigned)node * 2654435761u) >> shift;
    for (;;) {
        const int old = atomicCAS(keys + h, -1, node);
        if (old == -1 || old == node) return (int)h;
        h = (h + 1) & (unsigned)mask;
    }
}

__device__ __forceinline__ bool hash_has(const int32_t* keys, int mask, int shift, int node) {
    unsigned h = ((unsigned)node * 2654435761u) >> shift;
    for (;;) {
        const int k = keys[h];
        if (k == node) return true;
        if (k == -1) return false;
        h = (h + 1) & (unsigned)mask;
    }
}

// The fused expand is written as three block-wide phases over a context of shared-memory scratch,
// inputs and outputs, so that the one-batch kernel (expand_fused_kernel) and the many-batch kernel
// (expand_train_kernel: one CTA per batch of a whole train of batches) run the very same code.
struct FusedCtx {
    // shared-memory scratch
    int32_t* s_rowptr;       // nb + 1
    int32_t* s_eslot;        // sb: hash slot of each edge's target
    uint32_t* s_u;           // sb: draws, later the first-occurrence ranks
    int32_t* s_keys;         // hsize: node id or -1
    int32_t* s_vals;         // hsize: smallest claiming position
    int32_t* s_warp;         // 33
    int* s_status;
    int hmask, hshift;
    // inputs
    const int32_t* field_in; int n_out, sb;
    const int32_t* adj_p; int32_t* adj_i; float* adj_w; int N, degree, cv;
    // outputs
    int32_t* field; int32_t* rowptr_s; int32_t* rowptr_f; int32_t* edg_s; int32_t* edg_t;
    int32_t* tgt; float* edg_w; float* medg_w; float* scales;
};

// rows + both prefix sums, NT * IPT rows per sweep with a carry; old field -> table (value =
// position).  Every thread owns IPT CONSECUTIVE rows and issues their dependent loads together
// (ids -> row pointers), so the narrow co-resident variant pays one round trip per sweep, not IPT.
template <int NT, int IPT>
__device__ __forceinline__ void fused_rows_phase(const FusedCtx& c, int& carry_s, int& carry_f) {
    const int tid = threadIdx.x;
    carry_s = 0; carry_f = 0;
    for (int base = 0; base < c.n_out; base += NT * IPT) {
        int node[IPT], take[IPT], d[IPT], b0[IPT], b1[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = base + tid * IPT + q;
            node[q] = i < c.n_out ? c.field_in[i] : -1;
        }
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const bool in_range = node[q] >= 0 && node[q] < c.N;
            b0[q] = in_range ? c.adj_p[node[q]] : 0;
            b1[q] = in_range ? c.adj_p[node[q] + 1] : 0;
        }
        int sum_s = 0, sum_f = 0;
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = base + tid * IPT + q;
            take[q] = 0; d[q] = 0;
            if (i < c.n_out) {
                c.field[i] = node[q];
                if (node[q] < 0 || node[q] >= c.N) {
                    atomicOr(c.s_status, ST_RANGE);
                    c.scales[i] = 1.f;
                } else {
                    d[q] = b1[q] - b0[q];
                    take[q] = min(d[q], c.degree);
                    const float scale = (d[q] == 0) ? 1.f : __fdiv_rn((float)d[q], (float)take[q]);
                    c.scales[i] = (float)(1.0 / (double)__fsqrt_rn(scale));
                    const int h = hash_claim(c.s_keys, c.hmask, c.hshift, node[q]);
                    if (atomicMin(c.s_vals + h, i) != kUnseen) atomicOr(c.s_status, ST_DUPLICATE);
                }
            }
            sum_s += take[q];
            sum_f += c.cv ? d[q] : 0;
        }
        int tot_s, tot_f;
        int ex_s = block_scan_excl_nt<NT>(sum_s, &tot_s, c.s_warp);
        int ex_f = block_scan_excl_nt<NT>(sum_f, &tot_f, c.s_warp);
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = base + tid * IPT + q;
            if (i < c.n_out) {
                c.s_rowptr[i] = carry_s + ex_s;
                c.rowptr_s[i] = carry_s + ex_s;
                c.rowptr_f[i] = carry_f + ex_f;
            }
            ex_s += take[q];
            ex_f += c.cv ? d[q] : 0;
        }
        carry_s += tot_s;
        carry_f += tot_f;
    }
    if (tid == 0) {
        c.s_rowptr[c.n_out] = carry_s;
        c.rowptr_s[c.n_out] = carry_s;
        c.rowptr_f[c.n_out] = carry_f;
        if (carry_s > c.sb) *c.s_status |= ST_OVERFLOW;
    }
}

// per-row partial Fisher-Yates on the stored row (rows of one field are disjoint).  A thread walks
// IPT consecutive rows in three phases -- ids + row pointers, the touched row entries, swap +
// emit -- so the dependent global round trips of its rows overlap instead of queueing.
// (the draws of this batch are in c.s_u[0 .. nnz))
template <int NT, int IPT>
__device__ __forceinline__ void fused_shuffle_phase(const FusedCtx& c) {
    const int tid = threadIdx.x;
    const int n_out = c.n_out;
    for (int base = 0; base < n_out; base += NT * IPT) {
        int e0[IPT], take[IPT], d[IPT], rb[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const int i = base + tid * IPT + q;
            take[q] = 0; e0[q] = 0; d[q] = 0; rb[q] = 0;
            if (i < n_out) {
                e0[q] = c.s_rowptr[i];
                take[q] = c.s_rowptr[i + 1] - e0[q];
                if (take[q] <= 0 || e0[q] + take[q] > c.sb) take[q] = 0;
            }
        }
        int node[IPT];
#pragma unroll
        for (int q = 0; q < IPT; ++q) node[q] = take[q] > 0 ? c.field_in[base + tid * IPT + q] : 0;
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            if (take[q] > 0) {
                rb[q] = c.adj_p[node[q]];
                d[q] = c.adj_p[node[q] + 1] - rb[q];
            }
        }
        // draws of the rows that take <= 2 entries: every touched position is known from the draws
        // alone -- fetch them all at once, replay the <= 2 swaps on the local copy, store once
        int pos[IPT][4], cv4[IPT][4];
        float wv4[IPT][4];
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            if (take[q] > 0 && take[q] <= 2) {
                auto draw_index = [&](int k) {
                    // idx = min((int)(it + num_remaining * u01(generator)), adj_range-1)  (scheduler.cpp:141-143)
                    const float u = mt_canonical(c.s_u[e0[q] + k]);
                    const int j = (int)__fadd_rn((float)k, __fmul_rn((float)(d[q] - k), u));
                    return min(j, d[q] - 1);
                };
                const int j0 = draw_index(0);
                const int j1 = take[q] == 2 ? draw_index(1) : j0;
                pos[q][0] = 0; pos[q][1] = j0; pos[q][2] = take[q] == 2 ? 1 : 0; pos[q][3] = j1;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    cv4[q][x] = c.adj_i[rb[q] + pos[q][x]];
                    wv4[q][x] = c.adj_w[rb[q] + pos[q][x]];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            if (take[q] <= 0) continue;
            const int i = base + tid * IPT + q;
            int32_t* rc = c.adj_i + rb[q];
            float* rw = c.adj_w + rb[q];
            const float scale = __fdiv_rn((float)d[q], (float)take[q]);
            auto emit = [&](int k, int t, float wv) {
                const float w = __fmul_rn(wv, scale);
                const int e = e0[q] + k;
                c.edg_s[e] = i;
                c.tgt[e] = t;
                c.edg_w[e] = w;
                if (c.cv) c.medg_w[e] = __fmul_rn(wv, w);
                const int h = hash_claim(c.s_keys, c.hmask, c.hshift, t);
                atomicMin(c.s_vals + h, n_out + e);
                c.s_eslot[e] = h;
            };
            if (take[q] <= 2) {
                auto rd = [&](int x, int& cc, float& w) {
#pragma unroll
                    for (int y = 3; y >= 0; --y)
                        if (pos[q][y] == x) { cc = cv4[q][y]; w = wv4[q][y]; }
                };
                auto wr = [&](int x, int cc, float w) {
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        if (pos[q][y] == x) { cv4[q][y] = cc; wv4[q][y] = w; }
                };
                for (int k = 0; k < take[q]; ++k) {
                    const int j = k == 0 ? pos[q][1] : pos[q][3];
                    int ck = 0, cj = 0;
                    float wk = 0.f, wj = 0.f;
                    rd(k, ck, wk);
                    rd(j, cj, wj);
                    wr(k, cj, wj);
                    wr(j, ck, wk);
                    emit(k, cj, wj);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    rc[pos[q][x]] = cv4[q][x];
                    rw[pos[q][x]] = wv4[q][x];
                }
            } else {
                for (int k = 0; k < take[q]; ++k) {
                    const float u = mt_canonical(c.s_u[e0[q] + k]);
                    const int j = min((int)__fadd_rn((float)k, __fmul_rn((float)(d[q] - k), u)), d[q] - 1);
                    const int ck = rc[k], cj = rc[j];
                    const float wk = rw[k], wj = rw[j];
                    rc[k] = cj; rc[j] = ck;
                    rw[k] = wj; rw[j] = wk;
                    emit(k, cj, wj);
                }
            }
        }
    }
}

// first-occurrence flags -> ranks (reuse the draw buffer) -> column indices + the grown field.
// Returns the number of newly met nodes.
template <int NT>
__device__ __forceinline__ int fused_number_phase(const FusedCtx& c, int nnz) {
    const int tid = threadIdx.x;
    const int n_out = c.n_out;
    int32_t* s_rank = (int32_t*)c.s_u;
    int n_new = 0;
    constexpr int EPT = 4;                                     // consecutive edges per thread per sweep
    for (int base = 0; base < nnz; base += NT * EPT) {
        int f[EPT], sum = 0;
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int e = base + tid * EPT + q;
            f[q] = (e < nnz && c.s_vals[c.s_eslot[e]] == n_out + e) ? 1 : 0;
            sum += f[q];
        }
        int tot;
        int ex = block_scan_excl_nt<NT>(sum, &tot, c.s_warp);
#pragma unroll
        for (int q = 0; q < EPT; ++q) {
            const int e = base + tid * EPT + q;
            if (e < nnz) s_rank[e] = n_new + ex;
            ex += f[q];
        }
        n_new += tot;
    }
    __syncthreads();
    for (int e = tid; e < nnz; e += NT) {
        const int h = c.s_eslot[e];
        const int sl = c.s_vals[h];
        if (sl < n_out) {
            c.edg_t[e] = sl;
        } else {
            const int e_first = sl - n_out;
            const int pos = n_out + s_rank[e_first];
            c.edg_t[e] = pos;
            if (e_first == e) c.field[pos] = c.s_keys[h];
        }
    }
    return n_new;
}

// NT = 1024: fastest stand-alone.  NT = 256 (<= 48 registers): small enough to sit in the registers
// two resident full_mean_kernel CTAs leave free on an SM, so that in the pipelined step the next
// batch's sampler really runs BESIDE the aggregate instead of waiting for an empty SM.
template <int NT, int MIN_BLOCKS>
__global__ void __launch_bounds__(NT, MIN_BLOCKS)
expand_fused_kernel(const FusedArgs a) {
    constexpr int IPT = NT >= 1024 ? 1 : 2;                    // rows per thread per sweep
    extern __shared__ int32_t smem[];
    // carved to the launch's own bounds (fused_smem_bytes): a 512 x 2 batch needs ~46 KB, not the
    // 211 KB worst case, which keeps the SM's L1/shared carve-out where the neighbouring kernels want it
    const int hsize = 1 << a.hbits, hmask = hsize - 1, hshift = 32 - a.hbits;
    int32_t* s_rowptr = smem;                                  // nb + 1
    int32_t* s_eslot = s_rowptr + a.nb + 1;                    // sb: hash slot of each edge's target
    uint32_t* s_u = (uint32_t*)(s_eslot + a.sb);               // sb: draws, later the first-occurrence ranks
    int32_t* s_keys = (int32_t*)(s_u + a.sb);                  // hsize: node id or -1
    int32_t* s_vals = s_keys + hsize;                          // hsize: smallest claiming position
    uint32_t* s_mt = (uint32_t*)(s_vals + hsize);              // kMtN
    int32_t* s_warp = (int32_t*)(s_mt + kMtN);                 // 33
    __shared__ int s_status, s_pos, s_conflict;
    const int tid = threadIdx.x;
    TraceScope ts(a.trace, TR_SAMPLER);

    const int n_raw = a.n_ptr ? *a.n_ptr : a.n_host;
    const int n_out = min(n_raw, a.nb);
    if (tid == 0) {
        s_status = n_raw > a.nb ? ST_OVERFLOW : 0;
        s_pos = (int)a.engine[kMtN];
        s_conflict = 0;
    }
    for (int i = tid; i < kMtN; i += NT) s_mt[i] = a.engine[i];
    for (int i = tid; i < hsize; i += NT) {
        s_keys[i] = -1;
        s_vals[i] = kUnseen;
    }
    __syncthreads();

    const FusedCtx c{s_rowptr, s_eslot, s_u, s_keys, s_vals, s_warp, &s_status, hmask, hshift,
                     a.field_in, n_out, a.sb, a.adj_p, a.adj_i, a.adj_w, a.N, a.degree, a.cv,
                     a.field, a.rowptr_s, a.rowptr_f, a.edg_s, a.edg_t, a.tgt, a.edg_w, a.medg_w, a.scales};
    int carry_s, carry_f;
    fused_rows_phase<NT, IPT>(c, carry_s, carry_f);
    const int nnz = min(carry_s, a.sb);

    // the next nnz engine outputs -> shared memory
    {
        int pos = s_pos, produced = 0;
        while (produced < nnz) {
            if (pos >= kMtN) {
                mt_regen_block(s_mt);
                pos = 0;
            }
            const int avail = min(kMtN - pos, nnz - produced);
            for (int i = tid; i < avail; i += NT) s_u[produced + i] = mt_temper(s_mt[pos + i]);
            pos += avail;
            produced += avail;
        }
        __syncthreads();
        for (int i = tid; i < kMtN; i += NT) a.engine[i] = s_mt[i];
        if (tid == 0) a.engine[kMtN] = (uint32_t)pos;
    }

    // Pipelined steps: the previous batch's consumer pass may still be reading adjacency rows.  The
    // rows this expand permutes are those of ITS batch, so only a node shared with the previous
    // batch can race: if one exists, wait until every earlier consumer pass has finished.
    if (a.pipe) {
        const int seq = a.pipe[0];
        for (int j = tid; j < a.prev_n; j += NT)
            if (hash_has(s_keys, hmask, hshift, a.prev_ids[j])) s_conflict = 1;
        for (int j = tid; j < a.prev_n2; j += NT)
            if (hash_has(s_keys, hmask, hshift, a.prev_ids2[j])) s_conflict = 1;
        __syncthreads();
        if (s_conflict && tid == 0) {
            volatile int32_t* done = (volatile int32_t*)(a.pipe + 1);
            long long spins = 0;
            while (*done < seq) {
                if (++spins > 20000000LL) {       // ~2 s: report instead of hanging the GPU
                    s_status |= ST_OVERFLOW;
                    break;
                }
                __nanosleep(100);
            }
            __threadfence();
        }
        __syncthreads();
    }

    fused_shuffle_phase<NT, IPT>(c);
    __syncthreads();
    const int n_new = fused_number_phase<NT>(c, nnz);
    if (tid == 0) {
        a.meta[M_NOUT] = n_out;
        a.meta[M_NIN] = n_out + n_new;
        a.meta[M_NNZS] = nnz;
        a.meta[M_NNZF] = carry_f;
        a.meta[M_NFF] = 0;
        a.meta[M_STATUS] = s_status;
        a.meta[M_AUX] = 0;
        if (a.pipe) a.pipe[0] = a.pipe[0] + 1;
    }
}

// ---- a whole TRAIN of batches in one launch ---------------------------------------------------------
// The reference's Scheduler is sequential: batch after batch consumes one mt19937 stream and permutes
// the stored rows in place.  Between consecutive batches there are only two real dependencies: (1) the
// stream offset of batch j = the number of draws of batches < j, known as soon as their row degrees are
// (one load round trip), and (2) the stored row of a node that occurs in two batches.  So a train of n
// batches is sampled by ONE launch of n CTAs (one batch each, the same three phases as above):
//   * mt_stream_kernel first lays the next n * B * degree engine outputs out in HBM as raw state
//     blocks (block 0 = the engine's current state, block b = its b-th regeneration);
//   * a CTA takes a ticket j (order of arrival: it only ever waits for tickets < j, which are held by
//     CTAs already running -- no co-residency assumption), runs the rows phase of batch j, publishes
//     its draw count, sums the counts of tickets < j into its stream offset, tempers its draws out of
//     the raw blocks, shuffles and numbers;
//   * a batch that shares a node with an EARLIER batch of the train waits for those CTAs to finish
//     (order of the in-place permutations); one that shares a node with a batch of the PREVIOUS train
//     (which may still be read by its consumer passes) waits until pipe[1] says they are consumed;
//   * the last CTA to finish commits the engine: state block + cursor after the train's total draws.
// Results are bit-identical to n sequential expand calls.
constexpr int kTrainMax = 64;                 // batches per train
struct TrainSet {
    int32_t* field; int32_t* rowptr_s; int32_t* rowptr_f; int32_t* edg_s; int32_t* edg_t; int32_t* tgt;
    float* edg_w; float* medg_w; float* scales; int32_t* meta;
};
struct TrainCtl {
    int ticket, done;
    unsigned epoch;          // launches finished (the tag of cnt[] / fin[] entries is epoch + 1)
    int pos0;                // engine cursor when the raw stream was laid out
    int status;              // stream buffer too short etc.
    int pad[3];
    unsigned long long cnt[kTrainMax];     // (tag << 32) | draws of ticket j
    unsigned fin[kTrainMax];               // tag once ticket j has finished
};
struct TrainArgs {
    const int32_t* ids; int n, nb, sb;
    const int32_t* adj_p; int32_t* adj_i; float* adj_w; int N, degree, cv, hbits;
    const TrainSet* sets; int first_set, n_sets;
    const uint32_t* raw; int raw_words;
    TrainCtl* ctl; uint32_t* engine;
    const int32_t* prev_ids; int prev_n;      // batches of the previous train (consumer passes may still run)
    int32_t* pipe;                            // {batches sampled, consumer passes finished} or NULL
    unsigned long long* trace;
};

// raw[b * 624 + i] = word i of the engine's state after b regenerations (b = 0: the current state), for
// as many blocks as `need` further draws can touch.  One CTA.
static __global__ void __launch_bounds__(kMtThreads)
mt_stream_kernel(const uint32_t* __restrict__ engine, int need, uint32_t* __restrict__ raw, int raw_words,
                 TrainCtl* __restrict__ ctl) {
    __shared__ uint32_t x[kMtN];
    for (int i = threadIdx.x; i < kMtN; i += kMtThreads) {
        x[i] = engine[i];
        raw[i] = x[i];
    }
    const int pos0 = (int)engine[kMtN];
    __syncthreads();
    const int blocks = (pos0 + need + kMtN - 1) / kMtN;        // blocks 0 .. blocks-1 are touched
    int b = 1;
    for (; b < blocks && (b + 1) * kMtN <= raw_words; ++b) {
        mt_regen_block(x);
        for (int i = threadIdx.x; i < kMtN; i += kMtThreads) raw[b * kMtN + i] = x[i];
    }
    if (threadIdx.x == 0) {
        ctl->pos0 = pos0;
        ctl->status = b < blocks ? ST_OVERFLOW : 0;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT, 4)
expand_train_kernel(const TrainArgs a) {
    constexpr int IPT = 2;
    extern __shared__ int32_t smem[];
    const int hsize = 1 << a.hbits, hmask = hsize - 1, hshift = 32 - a.hbits;
    int32_t* s_rowptr = smem;
    int32_t* s_eslot = s_rowptr + a.nb + 1;
    uint32_t* s_u = (uint32_t*)(s_eslot + a.sb);
    int32_t* s_keys = (int32_t*)(s_u + a.sb);
    int32_t* s_vals = s_keys + hsize;
    int32_t* s_warp = s_vals + hsize;                          // 33
    __shared__ int s_status, s_conflict, s_j, s_last, s_seq;
    __shared__ unsigned s_tag;
    const int tid = threadIdx.x;
    TraceScope ts(a.trace, TR_SAMPLER);

    if (tid == 0) {
        s_tag = ld_acquire_u32(&a.ctl->epoch) + 1u;
        s_j = atomicAdd(&a.ctl->ticket, 1);
        s_status = a.ctl->status;
        s_conflict = 0;
        s_last = 0;
        s_seq = a.pipe ? a.pipe[0] : 0;        // batches sampled before this train (stable until the commit)
    }
    for (int i = tid; i < hsize; i += NT) {
        s_keys[i] = -1;
        s_vals[i] = kUnseen;
    }
    __syncthreads();
    const int j = s_j;
    const unsigned tag = s_tag;
    if (j >= a.n) return;
    const TrainSet set = a.sets[(a.first_set + j) % a.n_sets];
    const int32_t* field_in = a.ids + (size_t)j * a.nb;
    const int n_out = a.nb;

    const FusedCtx c{s_rowptr, s_eslot, s_u, s_keys, s_vals, s_warp, &s_status, hmask, hshift,
                     field_in, n_out, a.sb, a.adj_p, a.adj_i, a.adj_w, a.N, a.degree, a.cv,
                     set.field, set.rowptr_s, set.rowptr_f, set.edg_s, set.edg_t, set.tgt, set.edg_w,
                     set.medg_w, set.scales};
    int carry_s, carry_f;
    fused_rows_phase<NT, IPT>(c, carry_s, carry_f);
    const int nnz = min(carry_s, a.sb);
    if (tid == 0) st_release_u64(&a.ctl->cnt[j], ((unsigned long long)tag << 32) | (unsigned)nnz);

    // stream offset = draws of the tickets before mine
    int part = 0;
    for (int i = tid; i < j; i += NT) {
        unsigned long long v;
        long long spins = 0;
        while ((unsigned)((v = ld_acquire_u64(&a.ctl->cnt[i])) >> 32) != tag) {
            if (++spins > 20000000LL) { atomicOr(&s_status, ST_OVERFLOW); break; }
            __nanosleep(40);
        }
        part += (int)(unsigned)(v & 0xffffffffull);
    }
    int off;
    (void)block_scan_excl_nt<NT>(part, &off, s_warp);
    const int pos0 = a.ctl->pos0;
    if (pos0 + off + nnz > a.raw_words) {
        if (tid == 0) s_status |= ST_OVERFLOW;
    } else {
        for (int e = tid; e < nnz; e += NT) s_u[e] = mt_temper(a.raw[pos0 + off + e]);
    }

    // shared nodes: with an earlier batch of this train (bit 0), with the previous train (bit 1)
    {
        const int n_before = j * a.nb;
        int hit = 0;
        for (int q = tid; q < n_before; q += NT)
            if (hash_has(s_keys, hmask, hshift, a.ids[q])) hit |= 1;
        if (a.pipe)
            for (int q = tid; q < a.prev_n; q += NT)
                if (hash_has(s_keys, hmask, hshift, a.prev_ids[q])) hit |= 2;
        if (hit) atomicOr(&s_conflict, hit);
    }
    __syncthreads();
    if (s_conflict & 1) {
        for (int i = tid; i < j; i += NT) {
            long long spins = 0;
            while (ld_acquire_u32(&a.ctl->fin[i]) != tag) {
                if (++spins > 20000000LL) { atomicOr(&s_status, ST_OVERFLOW); break; }
                __nanosleep(100);
            }
        }
    }
    if ((s_conflict & 2) && tid == 0) {
        const unsigned* done = (const unsigned*)(a.pipe + 1);
        long long spins = 0;
        while ((int)ld_acquire_u32(done) < s_seq) {
            if (++spins > 20000000LL) { s_status |= ST_OVERFLOW; break; }
            __nanosleep(200);
        }
    }
    __syncthreads();

    fused_shuffle_phase<NT, IPT>(c);
    __syncthreads();
    const int n_new = fused_number_phase<NT>(c, nnz);
    __syncthreads();
    if (tid == 0) {
        set.meta[M_NOUT] = n_out;
        set.meta[M_NIN] = n_out + n_new;
        set.meta[M_NNZS] = nnz;
        set.meta[M_NNZF] = carry_f;
        set.meta[M_NFF] = 0;
        set.meta[M_STATUS] = s_status;
        set.meta[M_AUX] = 0;
        __threadfence();                                       // my rows / outputs before my flag
        st_release_u32(&a.ctl->fin[j], tag);
        s_last = atomicAdd(&a.ctl->done, 1) == a.n - 1;
    }
    __syncthreads();
    if (!s_last) return;
    // commit: every ticket has published its count before it bumped `done`
    __threadfence();
    int mine = 0;
    for (int i = tid; i < a.n; i += NT) mine += (int)(unsigned)(ld_acquire_u64(&a.ctl->cnt[i]) & 0xffffffffull);
    int total;
    (void)block_scan_excl_nt<NT>(mine, &total, s_warp);
    if (total > 0) {
        const int p = pos0 + total;
        int blk = p / kMtN, cur = p % kMtN;
        if (cur == 0) { blk -= 1; cur = kMtN; }              // libstdc++ regenerates lazily: cursor stays at 624
        if ((blk + 1) * kMtN <= a.raw_words) {
            for (int i = tid; i < kMtN; i += NT) a.engine[i] = a.raw[blk * kMtN + i];
            if (tid == 0) a.engine[kMtN] = (uint32_t)cur;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (a.pipe) a.pipe[0] = s_seq + a.n;
        a.ctl->ticket = 0;
        a.ctl->done = 0;
        __threadfence();
        st_release_u32(&a.ctl->epoch, tag);
    }
}

__global__ void mark_consumed_kernel(int32_t* pipe) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        __threadfence();
        atomicAdd(pipe + 1, 1);
    }
}

static bool fused_sampler_enabled() {
    const char* e = getenv("SGCN_NO_FUSED_SAMPLER");
    return !(e && e[0] == '1');
}

static int fill_int(int32_t* p, int64_t n, int32_t v, cudaStream_t st) {
    if (n <= 0) return SGCN_OK;
    fill_int_kernel<<<(int)std::min<int64_t>((n + 255) / 256, kNumSMs * 8), 256, 0, st>>>(p, n, v);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

static int read_meta(sgcn_sampler* s, const int32_t* meta_dev) {
    SGCN_CUDA(cudaMemcpyAsync(s->host_meta, meta_dev, sizeof(int32_t) * kMetaInts,
                              cudaMemcpyDeviceToHost, s->stream));
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    return SGCN_OK;
}

static int status_to_error(int status) {
    if (status == 0) return SGCN_OK;
    std::string msg = "sampler data error:";
    if (status & ST_DUPLICATE) msg += " duplicate ids in the batch;";
    if (status & ST_RANGE) msg += " node id out of range;";
    if (status & ST_NAN) msg += " nan edge weight (reference: runtime_error(\"nan\"));";
    if (status & ST_EMPTY) msg += " empty neighbour pool (reference: \"Prob is empty\");";
    if (status & ST_OVERFLOW) msg += " buffer bound exceeded;";
    set_error(msg);
    return SGCN_EDATA;
}

#define SGCN_TRY(expr)                     \
    do {                                   \
        int rc__ = (expr);                 \
        if (rc__ != SGCN_OK) return rc__;  \
    } while (0)

static int ensure_level(sgcn_sampler* s, Level& lv, int nb, int64_t sb, bool uniform) {
    const int64_t n_in = std::min<int64_t>((int64_t)nb + sb, std::max(s->N, nb));
    SGCN_TRY(lv.meta.ensure(sizeof(int32_t) * kMetaInts));
    SGCN_TRY(lv.field.ensure(sizeof(int32_t) * (size_t)std::max<int64_t>(n_in, 1)));
    SGCN_TRY(lv.rowptr_s.ensure(sizeof(int32_t) * ((size_t)nb + 1)));
    SGCN_TRY(lv.rowptr_f.ensure(sizeof(int32_t) * ((size_t)nb + 1)));
    SGCN_TRY(lv.scales.ensure(sizeof(float) * (size_t)std::max(nb, 1)));
    const size_t se = (size_t)std::max<int64_t>(sb, 1);
    SGCN_TRY(lv.edg_s.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.edg_t.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.tgt.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.edg_w.ensure(sizeof(float) * se));
    if (s->cv && uniform) SGCN_TRY(lv.medg_w.ensure(sizeof(float) * se));
    SGCN_TRY(s->draws.ensure(sizeof(uint32_t) * se));
    SGCN_TRY(s->rank.ensure(sizeof(int32_t) * se));
    SGCN_TRY(s->take.ensure(sizeof(int32_t) * (size_t)std::max(nb, 1)));
    SGCN_TRY(s->deg.ensure(sizeof(int32_t) * (size_t)std::max(nb, 1)));
    lv.n_out_bound = nb;
    lv.s_bound = (int)sb;
    lv.n_in_bound = (int)n_in;
    return SGCN_OK;
}

static int64_t sample_bound(const sgcn_sampler* s, int nb, int degree) {
    return std::min<int64_t>((int64_t)nb * std::max(degree, 0), s->E);
}

// cv: materialise ffield / fedg_* in the reference's format.  Synchronises (size = sum of degrees).
static int materialize_full(sgcn_sampler* s, Level& lv, int n_out, int nnz_f) {
    cudaStream_t st = s->stream;
    const size_t fe = (size_t)std::max(nnz_f, 1);
    SGCN_TRY(lv.fedg_s.ensure(sizeof(int32_t) * fe));
    SGCN_TRY(lv.fedg_t.ensure(sizeof(int32_t) * fe));
    SGCN_TRY(lv.fedg_w.ensure(sizeof(float) * fe));
    SGCN_TRY(lv.ffield.ensure(sizeof(int32_t) * (size_t)std::max(std::min(nnz_f, s->N), 1)));
    SGCN_TRY(s->rank.ensure(sizeof(int32_t) * fe));
    int32_t* meta = lv.meta.as<int32_t>();
    if (nnz_f > 0) {
        full_expand_kernel<<<div_up(nnz_f, 256), 256, 0, st>>>(
            lv.field.as<int32_t>(), lv.rowptr_f.as<int32_t>(), n_out, nnz_f, s->adj_p, s->adj_i,
            s->adj_w, lv.fedg_s.as<int32_t>(), lv.fedg_t.as<int32_t>(), lv.fedg_w.as<float>(),
            s->fslot);
        SGCN_LAUNCHED();
    }
    scan_full_first_kernel<<<1, kScanThreads, 0, st>>>(lv.fedg_t.as<int32_t>(), s->fslot, nnz_f,
                                                      s->rank.as<int32_t>(), meta, M_NFF);
    SGCN_LAUNCHED();
    if (nnz_f > 0) {
        full_assign_kernel<<<div_up(nnz_f, 256), 256, 0, st>>>(
            lv.fedg_t.as<int32_t>(), s->fslot, s->rank.as<int32_t>(), nnz_f, lv.ffield.as<int32_t>());
        SGCN_LAUNCHED();
        const int ffb = std::min(nnz_f, s->N);
        reset_list_kernel<<<div_up(ffb, 256), 256, 0, st>>>(lv.ffield.as<int32_t>(), meta + M_NFF,
                                                          ffb, s->fslot);
        SGCN_LAUNCHED();
    }
    lv.full_materialized = true;
    return SGCN_OK;
}

static int expand_uniform(sgcn_sampler* s, Level& lv, const int32_t* field_in,
                          const int32_t* n_ptr, int nb, int degree, int materialize) {
    cudaStream_t st = s->stream;
    int64_t sb = sample_bound(s, nb, degree);
    const bool exact = degree > kExactSizingDegree;
    SGCN_TRY(ensure_level(s, lv, nb, exact ? 0 : sb, true));
    int32_t* meta = lv.meta.as<int32_t>();
    const int nbl = std::max(nb, 1);

    if (!exact && nb <= kFusedMaxRows && sb <= kFusedMaxEdges && fused_sampler_enabled()) {
        static bool attr_set = false;
        if (!attr_set) {
            SGCN_CUDA(cudaFuncSetAttribute(expand_fused_kernel<1024, 1>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes));
            SGCN_CUDA(cudaFuncSetAttribute(expand_fused_kernel<256, 4>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes));
            // same ~100 KB carve-out as full_mean_kernel: whichever of the two reaches an SM first,
            // the other can join it without a re-partition
            SGCN_CUDA(cudaFuncSetAttribute(expand_fused_kernel<256, 4>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, kStepCarveout));
            attr_set = true;
        }
        const int hbits = fused_hash_bits(nb, (int)sb);
        // the one- / two-batch lookahead drivers rotate over buffer sets 0..2
        const sgcn_sampler::Slot& other = s->slots[(s->cur_slot + 2) % 3];
        const sgcn_sampler::Slot& other2 = s->slots[(s->cur_slot + 1) % 3];
        const bool piped = s->pipeline && s->sl().cur == 0;
        FusedArgs fa{field_in, n_ptr, nb, nb, (int)sb, s->adj_p, s->adj_i, s->adj_w, s->N, degree,
                     s->cv ? 1 : 0, hbits,
                     piped && other.batch_n > 0 ? other.batch_src : nullptr,
                     piped && other.batch_n > 0 ? other.batch_n : 0,
                     piped && other2.batch_n > 0 ? other2.batch_src : nullptr,
                     piped && other2.batch_n > 0 ? other2.batch_n : 0,
                     piped ? s->pipe_counters : nullptr, g_trace,
                     s->engine, lv.field.as<int32_t>(), lv.rowptr_s.as<int32_t>(),
                     lv.rowptr_f.as<int32_t>(), lv.edg_s.as<int32_t>(), lv.edg_t.as<int32_t>(),
                     lv.tgt.as<int32_t>(), lv.edg_w.as<float>(), lv.medg_w.as<float>(),
                     lv.scales.as<float>(), meta};
        if (s->pipeline)
            expand_fused_kernel<256, 4><<<1, 256, fused_smem_bytes(nb, (int)sb, hbits), st>>>(fa);
        else
            expand_fused_kernel<1024, 1><<<1, 1024, fused_smem_bytes(nb, (int)sb, hbits), st>>>(fa);
        SGCN_LAUNCHED();
        lv.full_materialized = false;
        if (s->cv && materialize) {
            SGCN_TRY(read_meta(s, meta));
            SGCN_TRY(materialize_full(s, lv, s->host_meta[M_NOUT], s->host_meta[M_NNZF]));
        }
        return SGCN_OK;
    }

    SGCN_CUDA(cudaMemsetAsync(meta, 0, sizeof(int32_t) * kMetaInts, st));
    row_setup_kernel<<<div_up(nbl, 256), 256, 0, st>>>(
        field_in, n_ptr, nb, nb, s->adj_p, s->N, degree, s->cv ? 1 : 0, 1, s->take.as<int32_t>(),
        s->deg.as<int32_t>(), lv.scales.as<float>(), lv.field.as<int32_t>(), s->slot, meta);
    SGCN_LAUNCHED();
    scan_take_deg_kernel<<<1, kScanThreads, 0, st>>>(s->take.as<int32_t>(), s->deg.as<int32_t>(),
                                                    lv.rowptr_s.as<int32_t>(),
                                                    lv.rowptr_f.as<int32_t>(), meta);
    SGCN_LAUNCHED();
    if (exact) {
        // large per-row degree ("Exact" runs): size the edge buffers from the true count.  The old
        // field prefix written by row_setup_kernel must survive the re-allocation.
        SGCN_TRY(read_meta(s, meta));
        sb = s->host_meta[M_NNZS];
        const int n_out = s->host_meta[M_NOUT];
        DevBuf keep;
        SGCN_TRY(keep.ensure(sizeof(int32_t) * (size_t)std::max(n_out, 1)));
        SGCN_CUDA(cudaMemcpyAsync(keep.p, lv.field.p, sizeof(int32_t) * (size_t)n_out,
                                  cudaMemcpyDeviceToDevice, st));
        SGCN_CUDA(cudaStreamSynchronize(st));
        int rc = ensure_level(s, lv, nb, sb, true);
        if (rc == SGCN_OK && n_out > 0 &&
            cudaMemcpyAsync(lv.field.p, keep.p, sizeof(int32_t) * (size_t)n_out,
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = SGCN_ECUDA;
        cudaStreamSynchronize(st);
        keep.release();
        SGCN_TRY(rc);
        meta = lv.meta.as<int32_t>();
    }
    const int sbl = (int)std::max<int64_t>(sb, 1);
    mt_draw_kernel<<<1, kMtThreads, 0, st>>>(s->engine, meta + M_NNZS, (int)sb,
                                            s->draws.as<uint32_t>());
    SGCN_LAUNCHED();
    fisher_yates_kernel<<<div_up(nbl, 128), 128, 0, st>>>(
        lv.field.as<int32_t>(), meta, s->adj_p, s->adj_i, s->adj_w, lv.rowptr_s.as<int32_t>(),
        s->draws.as<uint32_t>(), (int)sb, s->cv ? 1 : 0, lv.edg_s.as<int32_t>(),
        lv.tgt.as<int32_t>(), lv.edg_w.as<float>(), lv.medg_w.as<float>(), s->slot,
        meta + M_STATUS);
    SGCN_LAUNCHED();
    scan_first_kernel<<<1, kScanThreads, 0, st>>>(lv.tgt.as<int32_t>(), s->slot, (int)sb,
                                                 s->rank.as<int32_t>(), meta);
    SGCN_LAUNCHED();
    assign_kernel<<<div_up(sbl, 256), 256, 0, st>>>(lv.tgt.as<int32_t>(), s->slot,
                                                   s->rank.as<int32_t>(), meta, (int)sb,
                                                   lv.edg_t.as<int32_t>(), lv.field.as<int32_t>());
    SGCN_LAUNCHED();
    reset_slots_kernel<<<div_up(std::max(lv.n_in_bound, 1), 256), 256, 0, st>>>(
        lv.field.as<int32_t>(), meta, lv.n_in_bound, s->N, s->slot);
    SGCN_LAUNCHED();

    lv.full_materialized = false;
    if (s->cv && materialize) {
        SGCN_TRY(read_meta(s, meta));
        SGCN_TRY(materialize_full(s, lv, s->host_meta[M_NOUT], s->host_meta[M_NNZF]));
    }
    return SGCN_OK;
}

static int expand_importance(sgcn_sampler* s, Level& lv, const int32_t* field_in,
                             const int32_t* n_ptr, int nb, int degree) {
    cudaStream_t st = s->stream;
    SGCN_TRY(ensure_level(s, lv, nb, 0, false));
    int32_t* meta = lv.meta.as<int32_t>();
    const int nbl = std::max(nb, 1);
    // degrees of the field rows -> rowptr_f (take = 0: nothing is drawn per row here).  The
    // reference leaves `scales` empty in this branch (scheduler.cpp:63-123 never pushes).
    SGCN_CUDA(cudaMemsetAsync(meta, 0, sizeof(int32_t) * kMetaInts, st));
    row_setup_kernel<<<div_up(nbl, 256), 256, 0, st>>>(
        field_in, n_ptr, nb, nb, s->adj_p, s->N, 0, 1, 0, s->take.as<int32_t>(), s->deg.as<int32_t>(),
        lv.scales.as<float>(), lv.field.as<int32_t>(), s->slot, meta);
    SGCN_LAUNCHED();
    scan_take_deg_kernel<<<1, kScanThreads, 0, st>>>(s->take.as<int32_t>(), s->deg.as<int32_t>(),
                                                    lv.rowptr_s.as<int32_t>(),
                                                    lv.rowptr_f.as<int32_t>(), meta);
    SGCN_LAUNCHED();
    SGCN_TRY(read_meta(s, meta));
    const int n_out = s->host_meta[M_NOUT];
    const int nnz_f = s->host_meta[M_NNZF];
    if (s->host_meta[M_STATUS]) return status_to_error(s->host_meta[M_STATUS]);

    // neighbour pool in first-seen order == ffield of the rows (scheduler.cpp:66-78)
    SGCN_TRY(materialize_full(s, lv, n_out, nnz_f));
    lv.full_materialized = false;   // the pool is scratch, not an output of this branch
    SGCN_TRY(read_meta(s, meta));
    const int n_pool = s->host_meta[M_NFF];
    if (n_pool == 0) {
        // Mult::Mult throws "Prob is empty" (mult.cpp:17-18); the reference process would abort
        set_meta_kernel<<<1, 32, 0, st>>>(meta, n_out);
        SGCN_LAUNCHED();
        reset_slots_kernel<<<div_up(nbl, 256), 256, 0, st>>>(lv.field.as<int32_t>(), meta,
                                                            lv.n_in_bound, s->N, s->slot);
        SGCN_LAUNCHED();
        SGCN_CUDA(cudaStreamSynchronize(st));
        return status_to_error(ST_EMPTY);
    }
    // num_samples = min(field.size()*degree, neighbors.size()) in size_t (scheduler.cpp:83)
    const int64_t want = (int64_t)n_out * (int64_t)degree;
    const int n_draw = (int)std::min<int64_t>(want < 0 ? 0 : want, n_pool);
    int cap = n_pool;
    while (cap != (cap & (-cap))) cap += cap & (-cap);

    // the grown field can gain at most n_draw nodes; edges are counted exactly below
    {
        DevBuf keep;
        SGCN_TRY(keep.ensure(sizeof(int32_t) * (size_t)std::max(n_out, 1)));
        SGCN_CUDA(cudaMemcpyAsync(keep.p, lv.field.p, sizeof(int32_t) * (size_t)n_out,
                                  cudaMemcpyDeviceToDevice, st));
        SGCN_CUDA(cudaStreamSynchronize(st));
        int rc = lv.field.ensure(sizeof(int32_t) * (size_t)std::max(n_out + n_draw, 1));
        if (rc == SGCN_OK && n_out > 0 &&
            cudaMemcpyAsync(lv.field.p, keep.p, sizeof(int32_t) * (size_t)n_out,
                            cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = SGCN_ECUDA;
        cudaStreamSynchronize(st);
        keep.release();
        SGCN_TRY(rc);
        lv.n_in_bound = n_out + n_draw;
    }
    SGCN_TRY(s->pool_mass.ensure(sizeof(float) * ((size_t)n_pool + 1)));
    SGCN_TRY(s->tree.ensure(sizeof(float) * ((size_t)cap + 1)));
    SGCN_TRY(s->draws.ensure(sizeof(uint32_t) * (size_t)std::max(n_draw, 1)));
    float* mass = s->pool_mass.as<float>();
    float* mass_total = mass + n_pool;
    const int32_t* pool = lv.ffield.as<int32_t>();

    pool_mass_kernel<<<div_up(n_pool, 256), 256, 0, st>>>(pool, n_pool, s->importance, mass);
    SGCN_LAUNCHED();
    fenwick_build_kernel<<<div_up(cap, 256), 256, 0, st>>>(mass, n_pool, cap, s->tree.as<float>());
    SGCN_LAUNCHED();
    // Mult owns a default-constructed std::mt19937 (seed 5489), re-created by every expand
    // (mult.h:26, scheduler.cpp:82): the draws always restart that stream
    {
        uint32_t init[kMtWords];
        mt_seed_host(5489u, init);
        SGCN_CUDA(cudaMemcpyAsync(s->engine_is, init, sizeof(init), cudaMemcpyHostToDevice, st));
        SGCN_CUDA(cudaStreamSynchronize(st));   // `init` is a stack buffer
    }
    mt_draw_kernel<<<1, kMtThreads, 0, st>>>(s->engine_is, nullptr, n_draw, s->draws.as<uint32_t>());
    SGCN_LAUNCHED();
    importance_draw_kernel<<<1, 32, 0, st>>>(pool, mass, s->tree.as<float>(), n_pool, cap, n_draw,
                                            s->draws.as<uint32_t>(), s->hits.as<int32_t>(), s->slot,
                                            lv.field.as<int32_t>(), meta, mass_total);
    SGCN_LAUNCHED();
    const int row_blocks = std::min(div_up(std::max(n_out, 1), 8), kNumSMs * 8);
    importance_count_kernel<<<row_blocks, 256, 0, st>>>(lv.field.as<int32_t>(), n_out, s->adj_p,
                                                       s->adj_i, s->hits.as<int32_t>(),
                                                       s->take.as<int32_t>());
    SGCN_LAUNCHED();
    scan_counts_kernel<<<1, kScanThreads, 0, st>>>(s->take.as<int32_t>(), n_out,
                                                  lv.rowptr_s.as<int32_t>(), meta);
    SGCN_LAUNCHED();
    SGCN_TRY(read_meta(s, meta));
    const int nnz_s = s->host_meta[M_NNZS];
    const size_t se = (size_t)std::max(nnz_s, 1);
    SGCN_TRY(lv.edg_s.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.edg_t.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.tgt.ensure(sizeof(int32_t) * se));
    SGCN_TRY(lv.edg_w.ensure(sizeof(float) * se));
    lv.s_bound = nnz_s;
    importance_emit_kernel<<<row_blocks, 256, 0, st>>>(
        lv.field.as<int32_t>(), n_out, s->adj_p, s->adj_i, s->adj_w, s->hits.as<int32_t>(), s->slot,
        s->importance, mass_total, n_draw, lv.rowptr_s.as<int32_t>(), lv.edg_s.as<int32_t>(),
        lv.edg_t.as<int32_t>(), lv.tgt.as<int32_t>(), lv.edg_w.as<float>(), meta + M_STATUS);
    SGCN_LAUNCHED();
    reset_hits_kernel<<<div_up(n_pool, 256), 256, 0, st>>>(pool, n_pool, s->hits.as<int32_t>());
    SGCN_LAUNCHED();
    reset_slots_kernel<<<div_up(std::max(lv.n_in_bound, 1), 256), 256, 0, st>>>(
        lv.field.as<int32_t>(), meta, lv.n_in_bound, s->N, s->slot);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

static int create_common(sgcn_sampler** out, const float* adj_w, const int32_t* adj_i,
                         const int32_t* adj_p, int32_t num_data, int32_t num_edges, int32_t L,
                         int32_t cv, int32_t is, int32_t device, bool src_on_device) {
    SGCN_REQUIRE(out, "sampler_create: null out");
    *out = nullptr;
    SGCN_REQUIRE(num_data >= 0 && num_edges >= 0, "sampler_create: negative size");
    SGCN_REQUIRE(num_edges == 0 || (adj_w && adj_i), "sampler_create: null adjacency");
    SGCN_REQUIRE(num_data == 0 || adj_p, "sampler_create: null adj_p");
    int n_dev = 0;
    SGCN_CUDA(cudaGetDeviceCount(&n_dev));
    SGCN_REQUIRE(device >= 0 && device < n_dev, "sampler_create: no such CUDA device");
    DeviceGuard guard(device);
    sgcn_sampler* s = new sgcn_sampler();
    s->device = device;
    s->N = num_data;
    s->E = num_edges;
    s->L = std::max(L, 1);
    s->cv = cv != 0;
    s->is = is != 0;
    const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    auto fail = [&](int rc) {
        sgcn_sampler_destroy(s);
        return rc;
    };
#define CK(call)                                                                      \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #call, __FILE__, __LINE__)); \
    } while (0)
    CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    s->own_stream = true;
    const size_t ne = (size_t)std::max(num_edges, 1), nn = (size_t)std::max(num_data, 1);
    CK(cudaMalloc(&s->adj_w, sizeof(float) * ne));
    CK(cudaMalloc(&s->adj_i, sizeof(int32_t) * ne));
    CK(cudaMalloc(&s->adj_p, sizeof(int32_t) * (nn + 1)));
    CK(cudaMalloc(&s->slot, sizeof(int32_t) * nn));
    CK(cudaMalloc(&s->fslot, sizeof(int32_t) * nn));
    CK(cudaMalloc(&s->importance, sizeof(float) * nn));
    CK(cudaMalloc(&s->engine, sizeof(uint32_t) * kMtWords));
    CK(cudaMalloc(&s->engine_is, sizeof(uint32_t) * kMtWords));
    CK(cudaMallocHost(&s->host_meta, sizeof(int32_t) * kMetaInts));
    CK(cudaMalloc(&s->pipe_counters, sizeof(int32_t) * 4));
    CK(cudaMemset(s->pipe_counters, 0, sizeof(int32_t) * 4));
    if (num_edges > 0) {
        CK(cudaMemcpyAsync(s->adj_w, adj_w, sizeof(float) * (size_t)num_edges, kind, s->stream));
        CK(cudaMemcpyAsync(s->adj_i, adj_i, sizeof(int32_t) * (size_t)num_edges, kind, s->stream));
    }
    // the reference takes indptr[0..num_data) and appends num_edges itself (scheduler.cpp:16,20)
    if (num_data > 0)
        CK(cudaMemcpyAsync(s->adj_p, adj_p, sizeof(int32_t) * (size_t)num_data, kind, s->stream));
    const int32_t e32 = num_edges;
    CK(cudaMemcpyAsync(s->adj_p + num_data, &e32, sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));   // e32 is a stack variable
    int rc = fill_int(s->slot, num_data, kUnseen, s->stream);
    if (rc == SGCN_OK) rc = fill_int(s->fslot, num_data, kUnseen, s->stream);
    if (rc != SGCN_OK) return fail(rc);
    if (num_data > 0) {
        fill_float_kernel<<<std::min(div_up(num_data, 256), kNumSMs * 8), 256, 0, s->stream>>>(
            s->importance, num_data, s->is ? (float)1e-6 : 1.0f);
        g_launches.fetch_add(1);
        if (s->is) {
            if (num_edges > 0) {
                DevBuf sq, keys_out, vals_out, tmp;
                int irc = sq.ensure(sizeof(float) * ne);
                if (irc == SGCN_OK) irc = keys_out.ensure(sizeof(int32_t) * ne);
                if (irc == SGCN_OK) irc = vals_out.ensure(sizeof(float) * ne);
                if (irc != SGCN_OK) return fail(irc);
                square_kernel<<<std::min(div_up(num_edges, 256), kNumSMs * 8), 256, 0, s->stream>>>(
                    s->adj_w, num_edges, sq.as<float>());
                g_launches.fetch_add(1);
                size_t tmp_bytes = 0;
                int bits = 1;
                while ((1ll << bits) < (long long)std::max(num_data, 2)) ++bits;
                CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s->adj_i, keys_out.as<int32_t>(),
                                                   sq.as<float>(), vals_out.as<float>(), num_edges, 0, bits,
                                                   s->stream));
                irc = tmp.ensure(std::max<size_t>(tmp_bytes, 16));
                if (irc != SGCN_OK) return fail(irc);
                CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, s->adj_i, keys_out.as<int32_t>(),
                                                   sq.as<float>(), vals_out.as<float>(), num_edges, 0, bits,
                                                   s->stream));
                importance_columns_kernel<<<div_up(num_data, 256), 256, 0, s->stream>>>(
                    keys_out.as<int32_t>(), vals_out.as<float>(), num_edges, num_data, s->importance);
                g_launches.fetch_add(1);
                CK(cudaStreamSynchronize(s->stream));
                sq.release(); keys_out.release(); vals_out.release(); tmp.release();
            }
            rc = s->hits.ensure(sizeof(int32_t) * nn);
            if (rc != SGCN_OK) return fail(rc);
            CK(cudaMemsetAsync(s->hits.p, 0, sizeof(int32_t) * nn, s->stream));
        }
    }
    // default-constructed std::mt19937 until seed() is called
    uint32_t init[kMtWords];
    mt_seed_host(5489u, init);
    CK(cudaMemcpyAsync(s->engine, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaGetLastError());
#undef CK
    for (auto& slt : s->slots) slt.levels.resize((size_t)s->L);
    *out = s;
    return SGCN_OK;
}

static Level* level_at(sgcn_sampler* s, int32_t level) {
    if (level < 0) level = s->sl().cur - 1;
    if (level < 0 || level >= s->sl().cur || level >= (int)s->sl().levels.size()) return nullptr;
    return &s->sl().levels[(size_t)level];
}

}  // namespace sgcn

extern "C" {

int sgcn_sampler_create(sgcn_sampler** out, const float* adj_w, const int32_t* adj_i,
                        const int32_t* adj_p, int32_t num_data, int32_t num_edges, int32_t L,
                        int32_t cv, int32_t is, int32_t device) {
    return create_common(out, adj_w, adj_i, adj_p, num_data, num_edges, L, cv, is, device, false);
}

int sgcn_sampler_create_device(sgcn_sampler** out, const float* adj_w, const int32_t* adj_i,
                               const int32_t* adj_p, int32_t num_data, int32_t num_edges,
                               int32_t L, int32_t cv, int32_t is, int32_t device) {
    return create_common(out, adj_w, adj_i, adj_p, num_data, num_edges, L, cv, is, device, true);
}

void sgcn_sampler_destroy(sgcn_sampler* s) {
    if (!s) return;
    DeviceGuard guard(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& slt : s->slots) {
        for (Level& lv : slt.levels) lv.release();
        slt.batch_ids.release();
    }
    cudaFree(s->pipe_counters);
    for (DevBuf* b : {&s->batch_meta, &s->take, &s->deg, &s->draws, &s->rank,
                      &s->tile_sums, &s->pool_mass, &s->tree, &s->hits, &s->train_sets, &s->train_raw, &s->train_ctl})
        b->release();
    cudaFree(s->adj_w);
    cudaFree(s->adj_i);
    cudaFree(s->adj_p);
    cudaFree(s->slot);
    cudaFree(s->fslot);
    cudaFree(s->importance);
    cudaFree(s->engine);
    cudaFree(s->engine_is);
    if (s->host_meta) cudaFreeHost(s->host_meta);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int sgcn_sampler_seed(sgcn_sampler* s, int32_t seed) {
    SGCN_REQUIRE(s, "sampler_seed: null sampler");
    DeviceGuard guard(s->device);
    uint32_t init[kMtWords];
    mt_seed_host((uint32_t)seed, init);
    SGCN_CUDA(cudaMemcpyAsync(s->engine, init, sizeof(init), cudaMemcpyHostToDevice, s->stream));
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    return SGCN_OK;
}

int sgcn_sampler_set_slot(sgcn_sampler* s, int32_t slot) {
    SGCN_REQUIRE(s && slot >= 0 && slot < sgcn_sampler::kSlots, "sampler_set_slot: no such buffer set");
    s->cur_slot = slot;
    return SGCN_OK;
}

int sgcn_sampler_pipeline(sgcn_sampler* s, int32_t enable) {
    SGCN_REQUIRE(s, "sampler_pipeline: null sampler");
    DeviceGuard guard(s->device);
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    SGCN_CUDA(cudaMemset(s->pipe_counters, 0, sizeof(int32_t) * 4));
    s->pipeline = enable != 0;
    return SGCN_OK;
}

int sgcn_sampler_mark_consumed(sgcn_sampler* s, void* stream) {
    SGCN_REQUIRE(s, "sampler_mark_consumed: null sampler");
    DeviceGuard guard(s->device);
    mark_consumed_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(s->pipe_counters);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_sampler_set_stream_async(sgcn_sampler* s, void* stream) {
    SGCN_REQUIRE(s, "sampler_set_stream_async: null sampler");
    if (s->own_stream && s->stream) {
        set_error("sampler_set_stream_async: call sgcn_sampler_set_stream once first (the private stream must be retired with a synchronise)");
        return SGCN_ESTATE;
    }
    s->stream = (cudaStream_t)stream;
    return SGCN_OK;
}

int sgcn_sampler_set_stream(sgcn_sampler* s, void* stream) {
    SGCN_REQUIRE(s, "sampler_set_stream: null sampler");
    DeviceGuard guard(s->device);
    if (s->stream) SGCN_CUDA(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)stream;
    s->own_stream = false;
    return SGCN_OK;
}

int sgcn_sampler_reserve(sgcn_sampler* s, int32_t max_batch, const int32_t* degrees,
                         int32_t n_degrees, int32_t materialize_full) {
    SGCN_REQUIRE(s && max_batch >= 0 && n_degrees >= 0 && (n_degrees == 0 || degrees),
                 "sampler_reserve: bad argument");
    (void)materialize_full;   // the full COO has no useful size bound; it is grown on demand
    DeviceGuard guard(s->device);
    SGCN_TRY(s->sl().batch_ids.ensure(sizeof(int32_t) * (size_t)std::max(max_batch, 1)));
    SGCN_TRY(s->batch_meta.ensure(sizeof(int32_t) * kMetaInts));
    if ((int)s->sl().levels.size() < n_degrees) s->sl().levels.resize((size_t)n_degrees);
    int nb = max_batch;
    for (int k = 0; k < n_degrees; ++k) {
        Level& lv = s->sl().levels[(size_t)k];
        const bool exact = s->is || degrees[k] > kExactSizingDegree;
        SGCN_TRY(ensure_level(s, lv, nb, exact ? 0 : sample_bound(s, nb, degrees[k]), !s->is));
        if (exact) break;   // deeper bounds are unknown without running
        nb = lv.n_in_bound;
    }
    return SGCN_OK;
}

static int start_batch_common(sgcn_sampler* s, int32_t n, const int32_t* ids, cudaMemcpyKind kind) {
    SGCN_REQUIRE(s, "sampler_start_batch: null sampler");
    SGCN_REQUIRE(n >= 0 && (n == 0 || ids), "sampler_start_batch: bad ids");
    DeviceGuard guard(s->device);
    if (kind == cudaMemcpyHostToDevice) {
        SGCN_TRY(s->sl().batch_ids.ensure(sizeof(int32_t) * (size_t)std::max(n, 1)));
        if (n > 0) {
            SGCN_CUDA(cudaMemcpyAsync(s->sl().batch_ids.p, ids, sizeof(int32_t) * (size_t)n, kind, s->stream));
            SGCN_CUDA(cudaStreamSynchronize(s->stream));   // ids may be pageable host memory
        }
        s->sl().batch_src = s->sl().batch_ids.as<int32_t>();
    } else {
        // device ids are BORROWED, not copied: one launch less on the per-step critical path
        s->sl().batch_src = ids;
    }
    s->sl().batch_n = n;
    s->sl().cur = 0;
    for (Level& lv : s->sl().levels) lv.done = false;
    return SGCN_OK;
}

int sgcn_sampler_start_batch(sgcn_sampler* s, int32_t n, const int32_t* ids) {
    return start_batch_common(s, n, ids, cudaMemcpyHostToDevice);
}

int sgcn_sampler_start_batch_device(sgcn_sampler* s, int32_t n, const int32_t* ids) {
    return start_batch_common(s, n, ids, cudaMemcpyDeviceToDevice);
}

int sgcn_sampler_expand(sgcn_sampler* s, int32_t degree, int32_t materialize_full) {
    SGCN_REQUIRE(s, "sampler_expand: null sampler");
    if (s->sl().batch_n < 0) {
        set_error("sampler_expand called before sampler_start_batch");
        return SGCN_ESTATE;
    }
    SGCN_REQUIRE(degree >= 0, "sampler_expand: negative degree");
    DeviceGuard guard(s->device);
    const int k = s->sl().cur;
    if ((int)s->sl().levels.size() <= k) s->sl().levels.resize((size_t)k + 1);
    Level& lv = s->sl().levels[(size_t)k];
    const int32_t* field_in = k == 0 ? s->sl().batch_src : s->sl().levels[(size_t)k - 1].field.as<int32_t>();
    // level 0: the batch size is a host value (n_ptr == NULL, count = nb); deeper levels read the
    // previous level's |field| from its device meta block
    const int32_t* n_ptr = k == 0 ? nullptr : s->sl().levels[(size_t)k - 1].meta.as<int32_t>() + M_NIN;
    const int nb = k == 0 ? s->sl().batch_n : s->sl().levels[(size_t)k - 1].n_in_bound;
    int rc = s->is ? expand_importance(s, lv, field_in, n_ptr, nb, degree)
                   : expand_uniform(s, lv, field_in, n_ptr, nb, degree, materialize_full);
    if (rc != SGCN_OK) return rc;
    lv.done = true;
    s->sl().cur = k + 1;
    return SGCN_OK;
}


int sgcn_sampler_reserve_sets(sgcn_sampler* s, int32_t n_sets, int32_t batch, int32_t degree) {
    SGCN_REQUIRE(s && n_sets >= 1 && n_sets <= sgcn_sampler::kSlots && batch >= 1 && degree >= 0,
                 "sampler_reserve_sets: bad argument");
    SGCN_REQUIRE(!s->is, "sampler_reserve_sets: trains of batches exist for the uniform branch only");
    const int64_t sb = sample_bound(s, batch, degree);
    SGCN_REQUIRE(degree <= kExactSizingDegree && batch <= kFusedMaxRows && sb <= kFusedMaxEdges,
                 "sampler_reserve_sets: batch / degree beyond the fused sampler's bounds");
    DeviceGuard guard(s->device);
    const int keep = s->cur_slot;
    std::vector<TrainSet> sets((size_t)n_sets);
    s->train_field_ptrs.assign((size_t)n_sets, nullptr);
    for (int k = 0; k < n_sets; ++k) {
        s->cur_slot = k;
        if (s->sl().levels.empty()) s->sl().levels.resize(1);
        Level& lv = s->sl().levels[0];
        SGCN_TRY(s->sl().batch_ids.ensure(sizeof(int32_t) * (size_t)batch));
        SGCN_TRY(ensure_level(s, lv, batch, sb, true));
        SGCN_TRY(lv.medg_w.ensure(sizeof(float) * (size_t)std::max<int64_t>(sb, 1)));
        sets[(size_t)k] = TrainSet{lv.field.as<int32_t>(), lv.rowptr_s.as<int32_t>(), lv.rowptr_f.as<int32_t>(),
                                   lv.edg_s.as<int32_t>(), lv.edg_t.as<int32_t>(), lv.tgt.as<int32_t>(),
                                   lv.edg_w.as<float>(), lv.medg_w.as<float>(), lv.scales.as<float>(),
                                   lv.meta.as<int32_t>()};
        s->train_field_ptrs[(size_t)k] = lv.field.p;
    }
    s->cur_slot = keep;
    SGCN_TRY(s->batch_meta.ensure(sizeof(int32_t) * kMetaInts));
    SGCN_TRY(s->train_sets.ensure(sizeof(TrainSet) * (size_t)n_sets));
    SGCN_TRY(s->train_raw.ensure(sizeof(uint32_t) * ((size_t)kTrainMax * (size_t)std::max<int64_t>(sb, 1) + 3 * kMtN)));
    if (!s->train_ctl.p) {
        SGCN_TRY(s->train_ctl.ensure(sizeof(TrainCtl)));
        SGCN_CUDA(cudaMemsetAsync(s->train_ctl.p, 0, sizeof(TrainCtl), s->stream));
    }
    SGCN_CUDA(cudaMemcpyAsync(s->train_sets.p, sets.data(), sizeof(TrainSet) * (size_t)n_sets, cudaMemcpyHostToDevice,
                              s->stream));
    SGCN_CUDA(cudaStreamSynchronize(s->stream));       // `sets` is a host temporary
    s->train_n_sets = n_sets;
    s->train_batch = batch;
    s->train_degree = degree;
    return SGCN_OK;
}

int sgcn_sampler_expand_train(sgcn_sampler* s, const int32_t* ids, int32_t n, int32_t first_set,
                              const int32_t* prev_ids, int32_t prev_n, void* stream) {
    SGCN_REQUIRE(s && n >= 0 && (n == 0 || ids) && prev_n >= 0 && (prev_n == 0 || prev_ids),
                 "sampler_expand_train: bad argument");
    if (n == 0) return SGCN_OK;
    if (s->train_n_sets <= 0) {
        set_error("sampler_expand_train: call sgcn_sampler_reserve_sets first");
        return SGCN_ESTATE;
    }
    SGCN_REQUIRE(n <= kTrainMax && n <= s->train_n_sets && first_set >= 0 && first_set < s->train_n_sets,
                 "sampler_expand_train: train longer than the reserved buffer sets (or than 64)");
    DeviceGuard guard(s->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int B = s->train_batch, degree = s->train_degree;
    const int sb = (int)sample_bound(s, B, degree);
    for (int j = 0; j < n; ++j) {                      // buffers re-allocated since reserve_sets?
        const int k = (first_set + j) % s->train_n_sets;
        if (s->slots[k].levels.empty() || s->slots[k].levels[0].field.p != s->train_field_ptrs[(size_t)k]) {
            set_error("sampler_expand_train: a buffer set was re-allocated; call sgcn_sampler_reserve_sets again");
            return SGCN_ESTATE;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        SGCN_CUDA(cudaFuncSetAttribute(expand_train_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kFusedSmemBytes));
        SGCN_CUDA(cudaFuncSetAttribute(expand_train_kernel<256>, cudaFuncAttributePreferredSharedMemoryCarveout, kStepCarveout));
        attr_set = true;
    }
    const int hbits = fused_hash_bits(B, sb);
    const int raw_words = (int)(s->train_raw.cap / sizeof(uint32_t));
    mt_stream_kernel<<<1, kMtThreads, 0, st>>>(s->engine, n * sb, s->train_raw.as<uint32_t>(), raw_words,
                                              s->train_ctl.as<TrainCtl>());
    SGCN_LAUNCHED();
    TrainArgs ta{ids, n, B, sb, s->adj_p, s->adj_i, s->adj_w, s->N, degree, s->cv ? 1 : 0, hbits,
                 s->train_sets.as<TrainSet>(), first_set, s->train_n_sets, s->train_raw.as<uint32_t>(), raw_words,
                 s->train_ctl.as<TrainCtl>(), s->engine, s->pipeline ? prev_ids : nullptr, s->pipeline ? prev_n : 0,
                 s->pipeline ? s->pipe_counters : nullptr, g_trace};
    expand_train_kernel<256><<<n, 256, fused_smem_bytes(B, sb, hbits), st>>>(ta);
    SGCN_LAUNCHED();
    for (int j = 0; j < n; ++j) {                      // host bookkeeping: set k now holds batch j, level 0
        sgcn_sampler::Slot& sl = s->slots[(first_set + j) % s->train_n_sets];
        sl.batch_src = ids + (size_t)j * (size_t)B;
        sl.batch_n = B;
        sl.cur = 1;
        sl.levels[0].done = true;
        sl.levels[0].full_materialized = false;
    }
    return SGCN_OK;
}

int sgcn_sampler_sizes(sgcn_sampler* s, int32_t level, int32_t out[6]) {
    SGCN_REQUIRE(s && out, "sampler_sizes: null argument");
    DeviceGuard guard(s->device);
    Level* lv = level_at(s, level);
    if (!lv) {
        set_error("sampler_sizes: no such level (expand not called yet?)");
        return SGCN_ESTATE;
    }
    SGCN_TRY(read_meta(s, lv->meta.as<int32_t>()));
    for (int k = 0; k < 6; ++k) out[k] = s->host_meta[k];
    if (!lv->full_materialized) out[M_NFF] = 0;
    return status_to_error(out[M_STATUS]);
}

int sgcn_sampler_vec(sgcn_sampler* s, int32_t level, int32_t which, void** ptr, int64_t* len) {
    SGCN_REQUIRE(s && ptr && len, "sampler_vec: null argument");
    *ptr = nullptr;
    *len = 0;
    switch (which) {
        case SGCN_VEC_ADJ_I: *ptr = s->adj_i; *len = s->E; return SGCN_OK;
        case SGCN_VEC_ADJ_P: *ptr = s->adj_p; *len = (int64_t)s->N + 1; return SGCN_OK;
        case SGCN_VEC_ADJ_W: *ptr = s->adj_w; *len = s->E; return SGCN_OK;
        case SGCN_VEC_IMPORTANCE: *ptr = s->importance; *len = s->N; return SGCN_OK;
        case SGCN_VEC_PIPE: *ptr = s->pipe_counters; *len = 2; return SGCN_OK;
        default: break;
    }
    Level* lv = level_at(s, level);
    if (!lv) {
        set_error("sampler_vec: no such level (expand not called yet?)");
        return SGCN_ESTATE;
    }
    const DevBuf* b = nullptr;
    int64_t n = 0;
    switch (which) {
        case SGCN_VEC_FIELD: b = &lv->field; n = lv->n_in_bound; break;
        case SGCN_VEC_FFIELD: b = &lv->ffield; n = (int64_t)(lv->ffield.cap / 4); break;
        case SGCN_VEC_EDG_S: b = &lv->edg_s; n = lv->s_bound; break;
        case SGCN_VEC_EDG_T: b = &lv->edg_t; n = lv->s_bound; break;
        case SGCN_VEC_TGT: b = &lv->tgt; n = lv->s_bound; break;
        case SGCN_VEC_EDG_W: b = &lv->edg_w; n = lv->s_bound; break;
        case SGCN_VEC_MEDG_W: b = &lv->medg_w; n = lv->s_bound; break;
        case SGCN_VEC_FEDG_S: b = &lv->fedg_s; n = (int64_t)(lv->fedg_s.cap / 4); break;
        case SGCN_VEC_FEDG_T: b = &lv->fedg_t; n = (int64_t)(lv->fedg_t.cap / 4); break;
        case SGCN_VEC_FEDG_W: b = &lv->fedg_w; n = (int64_t)(lv->fedg_w.cap / 4); break;
        case SGCN_VEC_ROWPTR_S: b = &lv->rowptr_s; n = (int64_t)lv->n_out_bound + 1; break;
        case SGCN_VEC_ROWPTR_F: b = &lv->rowptr_f; n = (int64_t)lv->n_out_bound + 1; break;
        case SGCN_VEC_SCALES: b = &lv->scales; n = lv->n_out_bound; break;
        case SGCN_VEC_META: b = &lv->meta; n = kMetaInts; break;
        default:
            set_error("sampler_vec: unknown vector id");
            return SGCN_EINVAL;
    }
    *ptr = b->p;
    *len = b->p ? n : 0;
    return SGCN_OK;
}

int sgcn_sampler_slot_vec(sgcn_sampler* s, int32_t slot, int32_t which, void** ptr) {
    SGCN_REQUIRE(s && ptr && slot >= 0 && slot < sgcn_sampler::kSlots, "sampler_slot_vec: bad argument");
    *ptr = nullptr;
    auto& sl = s->slots[slot];
    if (sl.levels.empty()) {
        set_error("sampler_slot_vec: slot has no reserved level (call sgcn_sampler_reserve first)");
        return SGCN_ESTATE;
    }
    Level& lv = sl.levels[0];
    switch (which) {
        case SGCN_VEC_FIELD: *ptr = lv.field.p; break;
        case SGCN_VEC_EDG_S: *ptr = lv.edg_s.p; break;
        case SGCN_VEC_EDG_T: *ptr = lv.edg_t.p; break;
        case SGCN_VEC_TGT: *ptr = lv.tgt.p; break;
        case SGCN_VEC_EDG_W: *ptr = lv.edg_w.p; break;
        case SGCN_VEC_MEDG_W: *ptr = lv.medg_w.p; break;
        case SGCN_VEC_ROWPTR_S: *ptr = lv.rowptr_s.p; break;
        case SGCN_VEC_ROWPTR_F: *ptr = lv.rowptr_f.p; break;
        case SGCN_VEC_SCALES: *ptr = lv.scales.p; break;
        case SGCN_VEC_META: *ptr = lv.meta.p; break;
        default:
            set_error("sampler_slot_vec: unknown vector id");
            return SGCN_EINVAL;
    }
    if (!*ptr) {
        set_error("sampler_slot_vec: buffer not allocated (call sgcn_sampler_reserve for this slot first)");
        return SGCN_ESTATE;
    }
    return SGCN_OK;
}

int sgcn_sampler_copy_vec(sgcn_sampler* s, int32_t level, int32_t which, void* dst, int64_t count) {
    SGCN_REQUIRE(s && (dst || count == 0) && count >= 0, "sampler_copy_vec: bad argument");
    void* p = nullptr;
    int64_t len = 0;
    SGCN_TRY(sgcn_sampler_vec(s, level, which, &p, &len));
    DeviceGuard guard(s->device);
    if (count > 0) {
        SGCN_REQUIRE(p && count <= len, "sampler_copy_vec: count exceeds the vector's capacity");
        SGCN_CUDA(cudaMemcpyAsync(dst, p, (size_t)count * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    return SGCN_OK;
}

int sgcn_sampler_get_rng(sgcn_sampler* s, uint32_t state[624], int32_t* pos) {
    SGCN_REQUIRE(s && state && pos, "sampler_get_rng: null argument");
    DeviceGuard guard(s->device);
    uint32_t tmp[kMtWords];
    SGCN_CUDA(cudaMemcpyAsync(tmp, s->engine, sizeof(tmp), cudaMemcpyDeviceToHost, s->stream));
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < kMtN; ++i) state[i] = tmp[i];
    *pos = (int32_t)tmp[kMtN];
    return SGCN_OK;
}

int sgcn_sampler_set_rng(sgcn_sampler* s, const uint32_t state[624], int32_t pos) {
    SGCN_REQUIRE(s && state && pos >= 0 && pos <= kMtN, "sampler_set_rng: bad argument");
    DeviceGuard guard(s->device);
    uint32_t tmp[kMtWords];
    for (int i = 0; i < kMtN; ++i) tmp[i] = state[i];
    tmp[kMtN] = (uint32_t)pos;
    SGCN_CUDA(cudaMemcpyAsync(s->engine, tmp, sizeof(tmp), cudaMemcpyHostToDevice, s->stream));
    SGCN_CUDA(cudaStreamSynchronize(s->stream));
    return SGCN_OK;
}

}  // extern "C"
