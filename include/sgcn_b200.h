/* sgcn_b200.h -- C ABI of libsgcn_b200.so: the B200 (sm_100a) implementation of the
 * variance-reduced GCN training hot path of thu-ml/stochastic_gcn.
 *
 * This is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to the reference checkout, e.g. gcn/scheduler.h:6-28).  Signatures use plain
 * pointers and sizes only (no torch / CUDA types): `stream` arguments are a cudaStream_t passed
 * as void* (NULL = the legacy default stream).  Unless a parameter is marked HOST, every pointer
 * is a DEVICE pointer into the HBM of the sampler's / caller's current device.
 *
 * All functions return 0 on success and a negative SGCN_E* code on failure; the message for the
 * last failure on the calling thread is available from sgcn_last_error().
 * There is NO CPU fallback anywhere in this library.
 */
#ifndef SGCN_B200_H
#define SGCN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGCN_OK 0
#define SGCN_EINVAL (-1)   /* bad argument */
#define SGCN_ECUDA (-2)    /* CUDA runtime error (message has the CUDA string) */
#define SGCN_ESTATE (-3)   /* call out of order (e.g. expand before start_batch) */
#define SGCN_EDATA (-4)    /* data-dependent failure reported by a kernel (duplicate batch ids,
                              capacity overflow, NaN edge weight = the reference's
                              runtime_error("nan"), empty neighbour pool = "Prob is empty") */

int sgcn_abi_version(void);
const char* sgcn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t sgcn_launch_count(void);
/* Device timeline (profiling aid).  buf = device uint64[17 + 2*1024] or NULL (off).  Kernels launched
 * (or captured into a graph) while it is set stamp %globaltimer: buf[2k] = min block start
 * (initialise to ~0), buf[2k+1] = max block end (initialise to 0) for kernel class k: 0 sampler,
 * 1 full-neighbour mean, 2 row gather, 3 sampled aggregate, 4 SpMM backward, 5 history write-back,
 * 6 copy/zero-pad, 7 write-back exchange; buf[16] = event-log cursor (initialise to 0), then pairs
 * (class << 1 | is_end, time) stamped by block 0 of every launch. */
int sgcn_trace_set(void* buf);

/* ------------------------------------------------------------------------------------------
 * Neighbour sampler.  Replaces `class Scheduler` (gcn/scheduler.h:6-28, gcn/scheduler.cpp:11-189)
 * and `struct Mult` (gcn/mult.h:8-27) as bound by gcn/_scheduler.pyx:10-19.
 *
 * The sampler owns a private, mutable copy of the CSR adjacency in HBM (the reference deep-copies
 * it too, scheduler.cpp:14-16): the uniform branch permutes row entries in place and that
 * permutation, like the mt19937 state, persists across batches.  All outputs of expand() stay in
 * HBM; nothing is copied to the host unless a *_copy_* accessor is called.
 * ------------------------------------------------------------------------------------------ */
typedef struct sgcn_sampler sgcn_sampler;

/* Scheduler::Scheduler(adj_w, adj_i, adj_p, num_data, num_edges, L, cv, is)  scheduler.cpp:11-35.
 * adj_w/adj_i/adj_p are HOST arrays (adj_p needs num_data entries; entry num_data is taken to be
 * num_edges as the reference does).  `device` is the CUDA ordinal that will hold the state.
 * `L` is the number of expand() levels kept alive per batch (results of the k-th expand after
 * start_batch are addressed as level k). */
int sgcn_sampler_create(sgcn_sampler** out, const float* adj_w, const int32_t* adj_i,
                        const int32_t* adj_p, int32_t num_data, int32_t num_edges, int32_t L,
                        int32_t cv, int32_t is, int32_t device);
/* same, with adj_w/adj_i/adj_p already in the HBM of `device` (copied device-to-device: the sampler
 * still owns a private mutable copy) */
int sgcn_sampler_create_device(sgcn_sampler** out, const float* adj_w, const int32_t* adj_i,
                               const int32_t* adj_p, int32_t num_data, int32_t num_edges,
                               int32_t L, int32_t cv, int32_t is, int32_t device);
void sgcn_sampler_destroy(sgcn_sampler* s);

/* Scheduler::seed  scheduler.cpp:37-39 (std::mt19937::seed) */
int sgcn_sampler_seed(sgcn_sampler* s, int32_t seed);

/* Pre-size every per-batch buffer for batches of up to max_batch ids expanded with degrees[0..L)
 * (degrees[k] = degree of the k-th expand call).  Optional: expand() grows buffers on demand, but
 * growing cannot happen inside CUDA-graph capture.  materialize_full != 0 also sizes the
 * reference-format full-neighbour outputs (ffield / fedg_*). */
int sgcn_sampler_reserve(sgcn_sampler* s, int32_t max_batch, const int32_t* degrees /*HOST*/,
                         int32_t n_degrees, int32_t materialize_full);

/* Scheduler::start_batch  scheduler.cpp:41-44.  ids must be distinct node ids. */
int sgcn_sampler_start_batch(sgcn_sampler* s, int32_t n, const int32_t* ids /*HOST*/);
int sgcn_sampler_start_batch_device(sgcn_sampler* s, int32_t n, const int32_t* ids /*DEVICE*/);

/* Scheduler::expand(degree)  scheduler.cpp:46-189 (uniform / CV branch 125-180, importance branch
 * 63-123).  Asynchronous on the sampler's stream.  materialize_full selects whether the
 * reference-format full-neighbour COO (fedg_s/fedg_t/fedg_w + ffield) is written (cv only); the
 * fused aggregate kernels below do not need it -- they read the rows in place. */
int sgcn_sampler_expand(sgcn_sampler* s, int32_t degree, int32_t materialize_full);

/* sizes of level `level` (-1 = most recent expand): blocks until the level is complete.
 * out[0]=n_out |field before|, out[1]=n_in |field after|, out[2]=nnz_s sampled edges,
 * out[3]=nnz_f full-neighbour edges (cv), out[4]=n_ff |ffield| (cv, materialized), out[5]=status */
int sgcn_sampler_sizes(sgcn_sampler* s, int32_t level, int32_t out[6] /*HOST*/);

/* Device views of the reference's public vectors (scheduler.h:17-27) for one level.  *ptr is a
 * device pointer owned by the sampler, valid until the next start_batch; *len is the capacity
 * bound -- the exact length is in sgcn_sampler_sizes / the device meta block. */
enum {
    SGCN_VEC_FIELD = 0,    /* int32  field after expand (old field is its prefix)   [n_in]  */
    SGCN_VEC_FFIELD = 1,   /* int32  ffield                                          [n_ff]  */
    SGCN_VEC_EDG_S = 2,    /* int32  sampled edge row (index into old field)          [nnz_s] */
    SGCN_VEC_EDG_T = 3,    /* int32  sampled edge col (index into new field)          [nnz_s] */
    SGCN_VEC_FEDG_S = 4,   /* int32                                                   [nnz_f] */
    SGCN_VEC_FEDG_T = 5,   /* int32                                                   [nnz_f] */
    SGCN_VEC_ADJ_I = 6,    /* int32  the sampler's (permuted) CSR column ids          [E]     */
    SGCN_VEC_ADJ_P = 7,    /* int32  CSR row pointers                                 [N+1]   */
    SGCN_VEC_ROWPTR_S = 10,/* int32  CSR row pointers of the sampled adjacency        [n_out+1] */
    SGCN_VEC_ROWPTR_F = 11,/* int32  CSR row pointers of the full-neighbour adjacency [n_out+1] */
    SGCN_VEC_TGT = 12,     /* int32  sampled edge col as GLOBAL node id               [nnz_s] */
    SGCN_VEC_META = 13,    /* int32  device copy of the 6 sizes of sgcn_sampler_sizes [6]     */
    SGCN_VEC_PIPE = 14,    /* int32  pipelining counters {expands finished, consumer passes} [2] */
    SGCN_VEC_SCALES = 100, /* float  scales                                           [n_out] */
    SGCN_VEC_EDG_W = 101,  /* float                                                   [nnz_s] */
    SGCN_VEC_MEDG_W = 102, /* float                                                   [nnz_s] */
    SGCN_VEC_FEDG_W = 103, /* float                                                   [nnz_f] */
    SGCN_VEC_ADJ_W = 104,  /* float  the sampler's (permuted) CSR values              [E]     */
    SGCN_VEC_IMPORTANCE = 105 /* float importance                                     [N]     */
};
int sgcn_sampler_vec(sgcn_sampler* s, int32_t level, int32_t which, void** ptr /*HOST out*/,
                     int64_t* len /*HOST out*/);
/* copy the first `count` elements (4 bytes each) of a vector to HOST memory (synchronises) */
int sgcn_sampler_copy_vec(sgcn_sampler* s, int32_t level, int32_t which, void* dst /*HOST*/,
                          int64_t count);
/* Cross-step pipelining.  The sampler keeps three sets of per-batch buffers; set_slot selects the set
 * that start_batch / expand / sizes / vec address, so batches i+1 and i+2 can be sampled (on another
 * stream) while the kernels of batch i still read their own set.  With pipeline enabled, an expand
 * whose batch shares a node with a batch held by another set first waits (on the device, bounded)
 * until every earlier consumer pass has been marked finished (sgcn_sampler_mark_consumed, or the
 * done_counter of sgcn_history_update / sgcn_wb_wait_apply) -- the in-place row permutation must
 * not race the full-neighbour reads of an earlier batch.  Disjoint batches never wait. */
int sgcn_sampler_set_slot(sgcn_sampler* s, int32_t slot /*0 .. 127*/);
int sgcn_sampler_pipeline(sgcn_sampler* s, int32_t enable);
int sgcn_sampler_mark_consumed(sgcn_sampler* s, void* stream);
/* Trains of batches.  `Scheduler::expand` is sequential (one mt19937 stream, rows permuted in place,
 * scheduler.cpp:125-189), but consecutive batches depend on each other only through the stream OFFSET
 * (the number of draws of the earlier batches) and through the stored row of a node that occurs in two
 * of them -- so when the ids of the next n batches are known (the reference shuffles the whole epoch up
 * front, _scheduler.pyx:50-53,129-135) they are sampled by ONE launch of n thread blocks with exactly
 * the results of n sequential start_batch + expand(degree) calls.
 *   reserve_sets: sizes buffer sets 0 .. n_sets-1 (sgcn_sampler_set_slot ids) for batches of exactly
 *                 `batch` ids expanded once with `degree` (uniform branch, batch <= 4096, degree <= 32).
 *   expand_train: ids = DEVICE int32 [n][batch] (borrowed until the consumer passes have run); batch j
 *                 goes to buffer set (first_set + j) % n_sets, level 0.  n <= min(64, n_sets).
 *                 prev_ids / prev_n: the ids of batches sampled earlier whose consumer passes may still be
 *                 reading the adjacency (pipeline mode): a batch sharing a node with them waits on the
 *                 device until SGCN_VEC_PIPE[1] has caught up with the batches sampled before this train.
 *                 Asynchronous on `stream`; afterwards sizes / vec / slot_vec address the sets as usual. */
int sgcn_sampler_reserve_sets(sgcn_sampler* s, int32_t n_sets, int32_t batch, int32_t degree);
int sgcn_sampler_expand_train(sgcn_sampler* s, const int32_t* ids, int32_t n, int32_t first_set,
                              const int32_t* prev_ids, int32_t prev_n, void* stream);
/* the stream expand() runs on (cudaStream_t as void*); set to share the caller's stream */
int sgcn_sampler_set_stream(sgcn_sampler* s, void* stream);
/* same without synchronising the previous stream (legal during CUDA-graph capture; the caller
 * orders the streams with events) */
int sgcn_sampler_set_stream_async(sgcn_sampler* s, void* stream);
/* std::mt19937 state (624 words) + cursor, for checkpointing (absent in the reference) */
int sgcn_sampler_get_rng(sgcn_sampler* s, uint32_t state[624] /*HOST*/, int32_t* pos /*HOST*/);
int sgcn_sampler_set_rng(sgcn_sampler* s, const uint32_t state[624] /*HOST*/, int32_t pos);

/* ------------------------------------------------------------------------------------------
 * Row slicers.  Replace c_indptr / c_slice / c_dense_slice (gcn/history.h:6-9,
 * gcn/history.cpp:50-88) as bound by gcn/_history.pyx:25-62.
 * n_dev: optional device pointer to the row count (data-dependent sizes without a host
 * round-trip); when NULL the host value n is used, otherwise n is the launch bound.
 * ------------------------------------------------------------------------------------------ */
/* c_dense_slice: dst[i, 0:C] = src[idx[i], 0:C];  ld_* are row strides in elements */
int sgcn_gather_rows(const float* src, int64_t ld_src, const int32_t* idx, int32_t n,
                     const int32_t* n_dev, int32_t C, float* dst, int64_t ld_dst, void* stream);
/* c_indptr: o_p[i] = sum_{j<i} (a_p[r[j]+1]-a_p[r[j]]), o_p[n] = nnz */
int sgcn_csr_slice_indptr(const int32_t* a_p, const int32_t* r, int32_t n, int32_t* o_p,
                          void* stream);
/* c_slice: values + (local row, column) index pairs, rows in the order of r */
int sgcn_csr_slice(const float* a_d, const int32_t* a_i, const int32_t* a_p, const int32_t* r,
                   int32_t n, const int32_t* o_p, float* o_d, int32_t* o_i2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Aggregation.  Replaces the TensorFlow ops the reference's aggregators call:
 * tf.sparse_tensor_dense_matmul (gcn/layers.py:31-37), tf.gather (layers.py:304-305,354-355),
 * tf.concat of the self rows (layers.py:254-257,318-319,361-362), the implicit SpMM gradient
 * (gcn/models.py:187) and tf.scatter_update (gcn/models.py:160-166).
 *
 * Sparse operands are row-sorted (CSR): rowptr[n_out+1], cols[nnz], vals[nnz] -- exactly what
 * the sampler emits (edg_s ascending).  Dense operands are row-major fp32 with explicit row
 * strides (ld, in elements) so that outputs can be written straight into one half of a
 * concatenated [self | neighbour] buffer.
 * ------------------------------------------------------------------------------------------ */

/* y[r, :] (+)= sum_e vals[e] * x[map ? map[cols[e]] : cols[e], :]   for e in row r.
 * accumulate=0 overwrites y, 1 adds into it.   PlainAggregator: layers.py:249-257. */
int sgcn_spmm_csr(const int32_t* rowptr, const int32_t* cols, const float* vals,
                  const int32_t* map, int32_t n_out, const int32_t* n_out_dev, const float* x,
                  int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t accumulate,
                  void* stream);

/* General (unsorted) COO product with atomics: y[rows[e], :] += vals[e] * x[cols[e], :].
 * idx2 is the reference's interleaved int32[nnz,2] index array (_scheduler.pyx:80-84). */
int sgcn_spmm_coo(const int32_t* idx2, const float* vals, int32_t nnz, const float* x,
                  int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t transpose,
                  void* stream);

/* dx[cols[e], :] += vals[e] * (rscale ? rscale[r] : 1) * dy[r, :]  (SpMM backward, scatter-add).
 * dx must be initialised by the caller (zeros, or the self-row gradient). */
int sgcn_spmm_csr_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                      const float* rscale, int32_t n_out, const int32_t* n_out_dev,
                      const float* dy, int64_t ld_dy, int32_t D, float* dx, int64_t ld_dx,
                      void* stream);

/* Full-neighbour history mean, edge-balanced:  for each output row r with node s = nodes[r]:
 *   y0[r,:] += sum_{q in [adj_p[s], adj_p[s+1])} adj_w[q] * hist[adj_i[q], :]   (and y1 if given)
 * = dot(fadj, gather(history, ffield)) of layers.py:305,309,354,357 without materialising
 * ffield / fadj / the gathered rows.  rowptr_f[n_out+1] is the exclusive scan of the row
 * lengths (SGCN_VEC_ROWPTR_F).  work_counter: NULL = one contiguous span per warp; otherwise a
 * device int32 that is 0 on entry (the sampler zeroes meta[6] of every level for this purpose) from
 * which warps pull 64-edge chunks -- keeps all SMs busy when another kernel runs concurrently. */
int sgcn_full_history_mean(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                           const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                           const float* adj_w, const float* hist, int64_t ld_h, int32_t D,
                           float* y0, int64_t ld_y0, float* y1, int64_t ld_y1, int32_t* work_counter,
                           void* stream);

/* VRAggregator CV branch, sampled part (layers.py:350-362):
 *   y[r,:] (+)= sum_e vals[e] * (x[cols[e],:] - hist[tgt[e],:])
 *   self != NULL:  self[r,:] = x[r,:]   (the concat's left half, graphsage normalisation)
 * The full-neighbour term is added by sgcn_full_history_mean.  accumulate=0 overwrites y (run it
 * BEFORE sgcn_full_history_mean); accumulate=1 adds with 128-bit reductions into a y the caller
 * zeroed, so the two kernels commute and may run concurrently on different streams. */
int sgcn_cv_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                        const int32_t* tgt, int32_t n_out, const int32_t* n_out_dev,
                        const float* x, int64_t ld_x, const float* hist, int64_t ld_h, int32_t D,
                        float* y, int64_t ld_y, float* self, int64_t ld_self, int32_t accumulate,
                        void* stream);

/* VRAggregator CVD branch, sampled part (layers.py:298-319):
 *   ymu[r,:] = sum_e vals[e] * (mu[cols[e],:] - hist[tgt[e],:])
 *   yh[r,:]  = (sum_e vals[e] * (h[cols[e],:] - mu[cols[e],:])) * scale[r] + ymu[r,:]
 *   self_h / self_mu (optional): copies of the first n_out rows of h / mu. */
int sgcn_cvd_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const int32_t* tgt, const float* scale, int32_t n_out,
                         const int32_t* n_out_dev, const float* h, int64_t ld_hh,
                         const float* mu, int64_t ld_mu, const float* hist, int64_t ld_h,
                         int32_t D, float* yh, int64_t ld_yh, float* ymu, int64_t ld_ymu,
                         float* self_h, int64_t ld_sh, float* self_mu, int64_t ld_sm,
                         int32_t accumulate, void* stream);

/* Pre-processing product over the whole graph (gcn/utils.py:168-169,321-322: train_adj.dot(feats),
 * full_adj.dot(feats); stacked beside the self features by gcn/models.py:235-239):
 *   y[r, 0:D] = sum over the stored row r of adj_w[e] * x[adj_i[e], 0:D]        (y overwritten)
 * CSR (adj_p[n_rows+1] starting at 0, adj_i, adj_w) and x, y are DEVICE pointers; x and y may be
 * column blocks of wider matrices (ld_x, ld_y in floats), e.g. y = columns [602, 1204) of the
 * [N, 1204] model input whose columns [0, 602) hold x.  tile_cols: columns per launch (0 = 128):
 * the [N, tile] block of x is what has to stay L2-resident while the adjacency streams through.
 * Synchronises the stream once (reads the edge count): this is a one-off pre-processing call. */
int sgcn_csr_spmm(const int32_t* adj_p, const int32_t* adj_i, const float* adj_w, int32_t n_rows,
                  const float* x, int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t tile_cols,
                  void* stream);

/* ---- dense layers around the aggregate --------------------------------------
 * The products X @ W are library GEMMs issued by the host; these are the row-wise pieces.
 *
 * MyLayerNorm / MyLayerNorm2 + activation (layers.py:87-97,130-138,404-412):
 *   mean, var = moments over the row (biased);  y = act((x - mean) * rsqrt(var + eps) * scale + offset)
 * scale / offset: [D] device vectors or NULL (ones / zeros); relu != 0 applies max(., 0).
 * stats: optional [n, 2] device buffer (8-byte aligned) receiving {mean, rstd} for the backward. */
int sgcn_ln_act_fwd(const float* x, int64_t ld_x, int32_t n, const int32_t* n_dev, int32_t D,
                    const float* scale, const float* offset, float eps, int32_t relu, float* y,
                    int64_t ld_y, float* stats, void* stream);
/* gradient of the above: dx (optional) is overwritten; dscale / doffset (optional, [D]) are ADDED to. */
int sgcn_ln_act_bwd(const float* x, int64_t ld_x, const float* y, int64_t ld_y, const float* dy,
                    int64_t ld_dy, int32_t n, const int32_t* n_dev, int32_t D, const float* scale,
                    const float* stats, int32_t relu, float* dx, int64_t ld_dx, float* dscale,
                    float* doffset, void* stream);
/* tf.nn.dropout(x, keep_prob) (layers.py:396,415-433): y = kept ? x / keep_prob : 0.  The mask is
 * either injected (mask_in, one byte per element, row-major [n, D]) or drawn from Philox-4x32-10
 * keyed by `seed` with the element index + `offset` as counter; mask_out (optional) receives it.
 * TensorFlow's own random stream is not reproducible here (TF absent): parity uses injected masks. */
int sgcn_dropout(const float* x, int64_t ld_x, int32_t n, const int32_t* n_dev, int32_t D, float keep_prob,
                 uint64_t seed, uint64_t offset, const uint8_t* mask_in, uint8_t* mask_out, float* y,
                 int64_t ld_y, void* stream);
/* loss (models.py:76-83): *loss += mean over rows of softmax_cross_entropy_with_logits (sigmoid != 0:
 * mean over all entries of sigmoid_cross_entropy_with_logits);  dlogits (optional) = its gradient. */
int sgcn_xent(const float* logits, int64_t ld_l, const float* labels, int64_t ld_t, int32_t n, int32_t C,
              int32_t sigmoid, float* loss, float* dlogits, int64_t ld_d, void* stream);
/* tf.train.AdamOptimizer step (models.py:50-51,187) on one flat parameter: g' = g + weight_decay * p
 * (the gradient of weight_decay * l2_loss(p), models.py:68-74), m, v updated in place,
 * p -= lr_t * m / (sqrt(v) + eps) with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) computed by the caller. */
int sgcn_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1,
                   float beta2, float eps, float weight_decay, void* stream);

/* ---- first dense layer with the feature-row gather fused into its A-operand load (gcn/layers.py:100-138
 * `Dense` on the rows history.dense_slice picks, gcn/train.py:190) -- tcgen05 tensor cores, TF32 x 3 -------------
 *   out[i, :] = act(LN(src[idx[i], :K] @ W)),  W: [K, 128] row-major, i < min(n, *n_dev)
 * idx NULL: rows 0 .. n-1.  epilogue 0: the raw product; 1: MyLayerNorm (unit scale, zero offset, eps) + relu;
 * 2: MyLayerNorm.  pre (optional): the raw product as well (input of sgcn_ln_act_bwd); stats (optional): {mean,
 * rstd} per row as sgcn_ln_act_fwd writes them.  Both operands are split into tf32 hi + lo parts and three
 * tensor-core products accumulate in fp32 (hi.hi + lo.hi + hi.lo): the result matches the fp32 product to ~1e-6.
 * W is split and laid out once by sgcn_gemm_pack_w into `packed` (sgcn_gemm_packed_floats(K, 128) floats, 16-byte
 * aligned); repack after every update of W.  Source rows: 16-byte aligned, K and ld_src multiples of 4. */
int64_t sgcn_gemm_packed_floats(int32_t K, int32_t N);
int sgcn_gemm_pack_w(const float* w, int64_t ld_w, int32_t K, int32_t N, float* packed, void* stream);
int sgcn_gather_gemm_tf32x3(const float* src, int64_t ld_src, const int32_t* idx, int32_t n, const int32_t* n_dev,
                            int32_t K, const float* w_packed, int32_t N, float* out, int64_t ld_out, float* pre,
                            int64_t ld_pre, float* stats, int32_t epilogue, float eps, void* stream);

/* Runtime tunables (process-wide, not thread-safe against concurrent launches).
 *   SGCN_TUNE_FULL_VARIANT  0 = register-pipelined full_mean_kernel (default, fastest measured),
 *                           1 = bulk-copy (cp.async.bulk + mbarrier ring) full_mean_tma_kernel
 *                           (D <= 128, 16-byte aligned rows, n_out <= 4096; other shapes take 0)
 *   SGCN_TUNE_PDL           1 (default) = the step's chain kernels (full-neighbour mean, history
 *                           write-back, sampled aggregate) are launched with programmatic stream
 *                           serialization: each becomes resident while its stream predecessor still
 *                           runs, does what does not depend on it (the full mean: row pointers and
 *                           the first chunk's metadata) and orders the rest with griddepcontrol.wait;
 *                           0 = plain stream-ordered launches
 *   SGCN_TUNE_TMA_WARPS / _ROWS / _DEPTH  ring shape of variant 1: warps per CTA (1..16), history
 *                           rows per stage (1..32), stages per warp (1..4; clipped to fit 226 KB)
 *   SGCN_TUNE_TMA_GRID      CTAs of variant 1 (default 148 = one per SM; 147 leaves one SM to a kernel
 *                           that runs beside it, e.g. the next batch's sampler)
 *   SGCN_TUNE_HIST_L2       L2 eviction priority of history-row accesses: 0 = normal, p in 1..100 = evict_last
 *                           for p % of the accesses (the 119 MB Reddit-shaped table is about the size of L2;
 *                           everything else the step streams through L2 should not push it out)
 *   SGCN_TUNE_STREAM_L2     same for the read-once feature rows of the gather: p % evict_first
 *   SGCN_TUNE_FULL_TRIGGER  when full_mean_kernel lets its programmatic stream successor become resident:
 *                           0 = at entry, 1 (default) = after its positions when the write-back is fused into
 *                           its tail (the successor is the next full-neighbour mean), 2 = always after
 *   SGCN_TUNE_WB_TRIGGER    when the write-back kernels (sgcn_history_update, the exchange's copy pass) let their
 *                           programmatic stream successor -- the next full-neighbour mean -- launch: 1 (default) =
 *                           once their rows are stored, 0 = at entry (the mean then waits for room during the whole
 *                           previous mean and holds up every launch issued after it)
 *   SGCN_TUNE_FULL_REGS     register cap of full_mean_kernel: 96 (default; 2 thread blocks per SM) or 80 (3 per SM) */
enum { SGCN_TUNE_FULL_VARIANT = 0, SGCN_TUNE_TMA_WARPS = 1, SGCN_TUNE_TMA_ROWS = 2, SGCN_TUNE_TMA_DEPTH = 3,
       SGCN_TUNE_TMA_GRID = 4, SGCN_TUNE_PDL = 5, SGCN_TUNE_HIST_L2 = 6, SGCN_TUNE_STREAM_L2 = 7,
       SGCN_TUNE_FULL_TRIGGER = 8, SGCN_TUNE_FULL_REGS = 9, SGCN_TUNE_WB_TRIGGER = 10 };
int sgcn_tune_set(int32_t key, int32_t value);

/* ---- det-dropout (mu, var) aggregation: PlainAggregator tuple branch layers.py:238-247 and
 * VRAggregator tuple branch layers.py:320-349 ------------------------------
 * The mean stream reuses sgcn_spmm_csr / sgcn_cv_sampled_fwd with the mean history.  The variance
 * stream needs the element-wise squared adjacencies tf.square(adj), tf.square(fadj): the *_sq entry
 * points are the same kernels with vals[e]^2 in place of vals[e] (arguments as the plain versions). */
int sgcn_spmm_csr_sq(const int32_t* rowptr, const int32_t* cols, const float* vals,
                     const int32_t* map, int32_t n_out, const int32_t* n_out_dev, const float* x,
                     int64_t ld_x, int32_t D, float* y, int64_t ld_y, int32_t accumulate,
                     void* stream);
int sgcn_spmm_csr_bwd_sq(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* rscale, int32_t n_out, const int32_t* n_out_dev,
                         const float* dy, int64_t ld_dy, int32_t D, float* dx, int64_t ld_dx,
                         void* stream);
int sgcn_full_history_mean_sq(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                              const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                              const float* adj_w, const float* hist, int64_t ld_h, int32_t D,
                              float* y0, int64_t ld_y0, float* y1, int64_t ld_y1,
                              int32_t* work_counter, void* stream);

/* VRAggregator det-dropout branch, variance stream, sampled part + finish (layers.py:331-341):
 *   ds = sqrt(var[cols[e]]) - sqrt(hvar[tgt[e]]),  sb = sqrt(hvar[tgt[e]])
 *   pre[r,:] = (accumulate ? y[r,:] : 0) + sum_e vals[e]^2 ds^2 + 2 mvals[e] ds sb
 *   y[r,:]   = relu(pre[r,:]) + 1e-10 ;  self != NULL: self[r,:] = var[r,:]
 * Run sgcn_full_history_mean_sq(hvar) into a zeroed y FIRST, then this with accumulate=1 (each
 * output row is finished by its owner, no atomics).  pre (optional) keeps the pre-relu values, the
 * gate of the backward.  mvals = the sampler's medg_w (madj, _scheduler.pyx:116). */
int sgcn_det_sampled_fwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* mvals, const int32_t* tgt, int32_t n_out,
                         const int32_t* n_out_dev, const float* var, int64_t ld_v,
                         const float* hvar, int64_t ld_h, int32_t D, float* y, int64_t ld_y,
                         float* pre, int64_t ld_pre, float* self, int64_t ld_self,
                         int32_t accumulate, void* stream);
/* its gradient w.r.t. var (TF autodiff of the lines above; var_history is not trainable):
 *   dvar[cols[e],:] += (pre[r,:] > 0 ? dy[r,:] : 0) * (vals[e]^2 ds + mvals[e] sb) / sqrt(var[cols[e],:])
 * dvar is initialised by the caller (zeros, or the gradient of the self half). */
int sgcn_det_sampled_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                         const float* mvals, const int32_t* tgt, int32_t n_out,
                         const int32_t* n_out_dev, const float* var, int64_t ld_v,
                         const float* hvar, int64_t ld_h, int32_t D, const float* dy, int64_t ld_dy,
                         const float* pre, int64_t ld_pre, float* dvar, int64_t ld_dv, void* stream);

/* tf.scatter_update(history, fields[l], new_history)  models.py:160-166:
 *   hist[idx[i], :] = rows[i, :]   (idx distinct) */
int sgcn_history_update(float* hist, int64_t ld_h, const int32_t* idx, int32_t n,
                        const int32_t* n_dev, const float* rows, int64_t ld_rows, int32_t D,
                        int32_t* done_counter, void* stream);
/* done_counter (optional device int32): incremented once when the kernel starts, i.e. when everything
 * stream-ordered before the write-back has finished -- the pipelined step passes the sampler's
 * consumer counter (SGCN_VEC_PIPE[1]) instead of a separate sgcn_sampler_mark_consumed launch. */

/* dst[i, :] = src[i, :] for i < n (strided 2-D copy; self-row gradient / concat halves) and
 * dst[i, :] = 0 for n <= i < n_total */
int sgcn_copy_rows_pad(const float* src, int64_t ld_src, int32_t n, const int32_t* n_dev,
                       int32_t n_total, int32_t D, float* dst, int64_t ld_dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU write-back exchange (no counterpart in the single-process reference).
 * Every rank holds a full history replica; per step each rank publishes the rows it refreshed
 * (tf.scatter_update's operands, gcn/models.py:160-166) as a payload
 *     int32 header[4] = {count, step, 0, 0} | int32 ids[n_bound] (16-byte padded) | float rows[n_bound*D]
 * and every rank applies all payloads.  When several ranks refreshed the same node in one step the
 * highest (rank, position) wins as a whole row (deterministic).
 * ------------------------------------------------------------------------------------------ */
int64_t sgcn_wb_payload_bytes(int32_t n_bound, int32_t D);
/* pack field[0..*n_dev) and rows into n_dst destination buffers (own send buffer and / or the
 * peers' NVLink-mapped receive slots): dst is a HOST array of n_dst device pointers */
int sgcn_wb_pack(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                 int64_t ld_rows, int32_t D, void* const* dst /*HOST*/, int32_t n_dst, int32_t step,
                 void* stream);
/* Peer transport (NVLink-mapped memory, no collective launch; capturable in a CUDA graph).
 * Every rank owns, in cudaIpc-exported memory: two receive areas (even / odd epochs) of `world`
 * slots, a flag array int32[world] and a device epoch counter.
 * push: packs this rank's payload into slot `my_rank` of the (epoch+1)-parity receive area of EVERY
 * rank (dst_even / dst_odd: HOST arrays of n_dst device pointers, own slot included), then advances
 * *epoch and publishes it as flags[my_rank] in every rank (peer_flags: HOST array of n_dst pointers
 * to the ranks' flag arrays).  Two areas suffice: a rank can only push epoch e+1 after every rank
 * pushed e, i.e. after every rank finished applying e-1. */
int sgcn_wb_push(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                 int64_t ld_rows, int32_t D, void* const* dst_even /*HOST*/, void* const* dst_odd /*HOST*/,
                 int32_t n_dst, void* const* peer_flags /*HOST*/, int32_t my_rank, int32_t* epoch,
                 int32_t* block_counter /*device scratch int32 = 0, or NULL: signal from a 2nd launch*/,
                 void* stream);
/* Fuse the next write-back push into the next sampled-aggregate launch: the arguments of sgcn_wb_push
 * are remembered (per host thread) and the NEXT sgcn_cv_sampled_fwd[_bwd] / sgcn_cvd_sampled_fwd[_bwd]
 * call on this thread carries the push as extra thread blocks of its own kernel -- both only need the
 * gathered input rows, and every launch that leaves the step's side branch costs a dependent-launch
 * latency while the full-neighbour mean saturates the GPU.  block_counter is required (the last push
 * block publishes the epoch).  Same effect on memory as sgcn_wb_push issued on that call's stream. */
int sgcn_wb_push_attach(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                        int64_t ld_rows, int32_t D, void* const* dst_even, void* const* dst_odd,
                        int32_t n_dst, void* const* peer_flags, int32_t my_rank, int32_t* epoch,
                        int32_t* block_counter);

/* wait_apply: spins (bounded, ~2 s: sets *timeout_flag != 0 instead of hanging) until this rank's
 * flags[0..world) >= *epoch, then merges the `world` payloads of the epoch's receive area */
int sgcn_wb_wait_apply(float* hist, int64_t ld_h, int32_t D, const void* recv_even, const void* recv_odd,
                       int64_t slot_bytes, int32_t world, int32_t n_bound, int32_t* owner,
                       const int32_t* flags, const int32_t* epoch, int32_t* timeout_flag,
                       int32_t* done_counter /*optional, as sgcn_history_update*/, void* stream);
/* Ring form of the peer transport (what sgcn_step_run_trains uses): `ring` receive areas instead of two, and
 * separate counters for the epochs PUSHED (push_epoch, advanced by push_ring) and APPLIED (apply_epoch, advanced
 * by wait_apply_ring), so that a rank can publish the rows of pass k as soon as they are gathered -- one pass
 * ahead of the pass that applies them -- and fast ranks can run ahead of slow ones.  Epoch e lands in area
 * e % ring: dst_base[k] (HOST array) = rank k's receive base + my_rank * slot_bytes, area a at + a * ring_stride;
 * recv_base = this rank's receive base.  With the dependencies of sgcn_step_run_trains a rank is never more
 * than 6 epochs ahead of the slowest rank's applies: ring = 8. */
int sgcn_wb_push_ring(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                      int64_t ld_rows, int32_t D, void* const* dst_base /*HOST*/, int32_t n_dst, int32_t ring,
                      int64_t ring_stride, void* const* peer_flags /*HOST*/, int32_t my_rank, int32_t* push_epoch,
                      int32_t* block_counter, void* stream);
int sgcn_wb_wait_apply_ring(float* hist, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes,
                            int32_t world, int32_t n_bound, int32_t* owner, const int32_t* flags, int32_t ring,
                            int64_t ring_stride, int32_t* apply_epoch, int32_t* apply_counter /*scratch = 0*/,
                            int32_t* timeout_flag, int32_t* done_counter, void* stream);
/* The same epoch applied in two launches that need not share a stream: sgcn_wb_claim_ring (waits for every rank's
 * ring flag, takes the claims; a plain launch -- order it after the previous epoch's copy, e.g. by an event) and
 * sgcn_wb_copy_ring (the winners' rows into the table; launched programmatically behind the full-neighbour mean it
 * loads ids, claims and rows BEFORE griddepcontrol.wait and only stores behind it; adds 1 to *done_counter).  The
 * claim must have finished before the copy is launched.  Rows: 16-byte aligned, D in {4, 8, 16, 32, 64, 128}
 * (other shapes: sgcn_wb_wait_apply_ring). */
int sgcn_wb_claim_ring(const void* recv_base, int64_t slot_bytes, int32_t world, int32_t n_bound, int32_t* owner,
                       const int32_t* flags, int32_t ring, int64_t ring_stride, const int32_t* apply_epoch,
                       int32_t* apply_stash, int32_t* timeout_flag, void* stream);
int sgcn_wb_copy_ring(float* hist, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes, int32_t world,
                      int32_t n_bound, int32_t* owner, int32_t ring, int64_t ring_stride, int32_t* apply_epoch,
                      const int32_t* apply_stash, int32_t* done_counter, void* stream);
/* Row-sharded tables ("boundary fetch").  Instead of replicating the history table and the PP
 * feature matrix on every GPU, rank r keeps rows [r * rows_per_shard, (r + 1) * rows_per_shard) and maps every
 * other rank's shard over NVLink (cudaIpc); the kernels that read history rows (full-neighbour mean, CV / CVD
 * sampled aggregate) or gather feature rows then address row i at bases[i / rows_per_shard] + (i %
 * rows_per_shard) * ld -- the rows a batch's receptive field needs from the other side of a cut cross NVLink as
 * plain loads, exactly once each, and nothing else moves.  sgcn_shard_set arms that addressing for the calling
 * host thread (which = 0 history, 1 features; bases = HOST array of `world` device pointers; world <= 1 turns it
 * off); the `hist` / `src` pointer passed to those entry points is then this rank's own shard. */
int sgcn_shard_set(int32_t which, int32_t world, int32_t rows_per_shard, const void* const* bases /*HOST*/);
/* write-back with a sharded history: every rank still publishes its rows to every rank (sgcn_wb_push_ring), but
 * applies only the rows it OWNS, to its shard.  Two handshakes replace the stream order a single table gave:
 * nobody overwrites a row before every rank has finished the reads of the pass (reads flags), nobody reads a
 * shard before its owner has applied the epoch (applied flags; the call returns on the device only then). */
int sgcn_wb_wait_apply_sharded(float* hist_shard, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes,
                               int32_t world, int32_t rank, int32_t rows_per_shard, int32_t n_bound, int32_t* owner,
                               const int32_t* flags, int32_t ring, int64_t ring_stride, int32_t* apply_epoch,
                               int32_t* apply_stash, const int32_t* reads_flags, void* const* reads_peer_flags,
                               const int32_t* applied_flags, void* const* applied_peer_flags, int32_t* shard_counter,
                               int32_t* timeout_flag, int32_t* done_counter, void* stream);
/* merge `world` payloads (slot r at gathered + r*slot_bytes) into hist; owner is an int32[N]
 * scratch table that must hold -1 everywhere on entry and does again on exit */
int sgcn_wb_apply(float* hist, int64_t ld_h, int32_t D, const void* gathered, int64_t slot_bytes,
                  int32_t world, int32_t n_bound, int32_t* owner, void* stream);
/* cudaMalloc'd buffers that other processes can map over NVLink (cudaIpc*): handle64 = 64 bytes */
int sgcn_ipc_alloc(void** ptr /*HOST out*/, int64_t bytes, int32_t zero);
int sgcn_ipc_free(void* ptr);
int sgcn_ipc_export(void* ptr, void* handle64 /*HOST out*/);
int sgcn_ipc_open(const void* handle64 /*HOST*/, void** ptr /*HOST out*/);
int sgcn_ipc_close(void* ptr);

/* ------------------------------------------------------------------------------------------
 * Step-level fusions (fewer CUDA-graph nodes per training step; same arithmetic as the pieces).
 * ------------------------------------------------------------------------------------------ */
/* two independent sgcn_copy_rows_pad jobs in one launch (a NULL dst skips a job) */
int sgcn_copy_rows_pad_pair(const float* src0, int64_t ld_src0, int32_t n0, const int32_t* n0_dev,
                            int32_t n_total0, int32_t D0, float* dst0, int64_t ld_dst0,
                            const float* src1, int64_t ld_src1, int32_t n1, const int32_t* n1_dev,
                            int32_t n_total1, int32_t D1, float* dst1, int64_t ld_dst1, void* stream);

/* sgcn_gather_rows + sgcn_copy_rows_pad_pair in ONE launch (the step's side branch starts with these
 * three independent row jobs; one graph node instead of two saves a dependent launch per step). */
int sgcn_gather_pad_pair(const float* src, int64_t ld_src, const int32_t* idx, int32_t n,
                         const int32_t* n_dev, int32_t C, float* dst, int64_t ld_dst,
                         const float* src0, int64_t ld_src0, int32_t n0, const int32_t* n0_dev,
                         int32_t n_total0, int32_t D0, float* dst0, int64_t ld_dst0,
                         const float* src1, int64_t ld_src1, int32_t n1, const int32_t* n1_dev,
                         int32_t n_total1, int32_t D1, float* dst1, int64_t ld_dst1, void* stream);
/* sgcn_cv_sampled_fwd / sgcn_cvd_sampled_fwd followed, row by row in the same kernel, by the
 * backward scatter of sgcn_spmm_csr_bwd (dx[cols[e]] += vals[e] * (scale[r]) * dy[r]; dx initialised
 * by the caller).  The forward never reads dx and the backward never reads y, so fusing them only
 * saves a launch. */
int sgcn_cv_sampled_fwd_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                            const int32_t* tgt, int32_t n_out, const int32_t* n_out_dev,
                            const float* x, int64_t ld_x, const float* hist, int64_t ld_h, int32_t D,
                            float* y, int64_t ld_y, float* self, int64_t ld_self, int32_t accumulate,
                            const float* dy, int64_t ld_dy, float* dx, int64_t ld_dx, void* stream);
int sgcn_cvd_sampled_fwd_bwd(const int32_t* rowptr, const int32_t* cols, const float* vals,
                             const int32_t* tgt, const float* scale, int32_t n_out,
                             const int32_t* n_out_dev, const float* h, int64_t ld_hh,
                             const float* mu, int64_t ld_mu, const float* hist, int64_t ld_h,
                             int32_t D, float* yh, int64_t ld_yh, float* ymu, int64_t ld_ymu,
                             float* self_h, int64_t ld_sh, float* self_mu, int64_t ld_sm,
                             int32_t accumulate, const float* dy, int64_t ld_dy, float* dx,
                             int64_t ld_dx, void* stream);

/* level-0 buffer of one of the sampler's two buffer sets, valid once sgcn_sampler_reserve has run
 * for that slot (before any expand) -- which: SGCN_VEC_FIELD / EDG_* / TGT / ROWPTR_* / SCALES / META */
int sgcn_sampler_slot_vec(sgcn_sampler* s, int32_t slot, int32_t which, void** ptr /*HOST out*/);

/* ------------------------------------------------------------------------------------------
 * Native step driver: n consecutive passes (gcn/train.py:187-209 inner loop: minibatch ->
 * run_one_step) issued from C++ onto three internal streams, the sampler of batch k+1 running
 * beside the aggregate of batch k.  Same results as n sequential passes.
 * ------------------------------------------------------------------------------------------ */
typedef struct sgcn_step sgcn_step;
typedef struct {
    int32_t mode;              /* 0 = NS / plain, 1 = CV, 2 = CVD  (gcn/layers.py:214-362) */
    int32_t concat;            /* graphsage normalisation: output rows = [self | neighbour] */
    int32_t batch, degree, hidden, feat_dim;
    int32_t x0_rows;           /* rows of x0 / dx: >= batch * (1 + degree) */
    int32_t world, rank, wb_bound;   /* multi-GPU peer exchange (world <= 1: single GPU) */
    const float* features; int64_t ld_feat;      /* [N, feat_dim] input (PP) features */
    float* history; int64_t ld_hist;             /* [N, hidden] (CV / CVD) */
    float* x0; int64_t ld_x0;                    /* [x0_rows, feat_dim] gathered input rows */
    float* out[2]; float* out_mu[2]; int64_t ld_out;   /* [batch, hidden * (1 + concat)] per buffer set */
    const float* d_out; int64_t ld_dout;         /* upstream gradient, same shape as out */
    float* dx; int64_t ld_dx;                    /* [x0_rows, hidden] */
    int64_t slot_bytes;
    void* dst_even[16]; void* dst_odd[16]; void* peer_flags[16];
    void* recv_even; void* recv_odd; const int32_t* flags;
    int32_t* epoch; int32_t* timeout_flag; int32_t* block_counter; int32_t* owner;
    /* sgcn_step_run_trains only (may be NULL / 0 otherwise): two more copies of x0 and one more of dx (same
     * shapes and strides), the number of batches sampled per launch (2 .. 32, 0 = 16) and whether the
     * history write-back rides on the full-neighbour mean's launch (single GPU, CV / CVD) */
    float* x0_alt[2]; float* dx_alt;
    int32_t train; int32_t fuse_write_back;
    /* ring form of the peer exchange, used by sgcn_step_run_trains when ring > 0 (sgcn_wb_push_ring /
     * sgcn_wb_wait_apply_ring: the rows of pass k are pushed as soon as they are gathered, one pass ahead):
     * ring_dst[k] = rank k's receive base + rank * slot_bytes, ring_recv = this rank's receive base,
     * ring_peer_flags[k] = rank k's ring flag array, ring_flags = this rank's */
    int32_t ring; int32_t pad0; int64_t ring_stride;
    int32_t* push_epoch; int32_t* apply_epoch; int32_t* apply_stash; const int32_t* ring_flags;
    void* ring_dst[16]; void* ring_peer_flags[16]; void* ring_recv;
    /* row-sharded history and features (sgcn_step_run_trains, ring form only; shard_rows = 0: replicated):
     * `history` / `features` above are then THIS rank's shards (rows [rank * shard_rows, ...)), hist_shards /
     * feat_shards every rank's shard as mapped into this process (sgcn_shard_set), and the two flag arrays
     * carry the "reads done" / "applied" handshakes of sgcn_wb_wait_apply_sharded */
    int32_t shard_rows; int32_t pad1;
    const void* hist_shards[16]; const void* feat_shards[16];
    const int32_t* reads_flags; void* reads_peer_flags[16];
    const int32_t* applied_flags; void* applied_peer_flags[16];
    int32_t* shard_counter;
} sgcn_step_desc;
int sgcn_step_create(sgcn_step** out, sgcn_sampler* sampler, const sgcn_step_desc* desc /*HOST*/);
void sgcn_step_destroy(sgcn_step* st);
/* ids: int32 [n][batch], device memory (ids_on_host = 0; borrowed until the run has finished) or
 * PINNED host memory (1: copied H2D step by step on the sampler stream).  out_host: NULL, or pinned
 * float [n][batch][hidden * (1 + concat)] receiving every pass's aggregated rows (D2H per pass).
 * The run is ordered after the work already on `stream`, and `stream` waits for its completion.
 * The sampler must be in pipeline mode (sgcn_sampler_pipeline) and have had sgcn_sampler_set_stream
 * called once. */
int sgcn_step_run(sgcn_step* st, const int32_t* ids, int32_t ids_on_host, int32_t n, float* out_host,
                  void* stream);

/* sgcn_full_history_mean with the pass's history write-back (gcn/models.py:160-166,186-194) fused into its tail:
 * once EVERY thread block of the launch has consumed its history rows and the pass's sampled aggregate has
 * consumed its own (counters[2] > counters[3], see sgcn_sampled_done_attach), the last thread blocks to finish
 * store hist[wb_ids[i], :] = wb_rows[i, :D] for i < min(*wb_n_dev, wb_bound) (ids distinct), then add 1 to
 * *consumed (optional: the sampler's consumer counter) and to counters[3].  The result is what
 * sgcn_full_history_mean followed by sgcn_history_update gives, without a second launch on the step's chain.
 * counters: device int32[8], zeroed by sgcn_wb_counters_reset before the first pass of a run; counters[4] != 0
 * afterwards = a wait gave up (bounded spins).  Not for row-sharded tables. */
int sgcn_full_history_mean_wb(const int32_t* nodes, const int32_t* rowptr_f, int32_t n_out,
                              const int32_t* n_out_dev, const int32_t* adj_p, const int32_t* adj_i,
                              const float* adj_w, float* hist, int64_t ld_h, int32_t D,
                              float* y0, int64_t ld_y0, float* y1, int64_t ld_y1,
                              const int32_t* wb_ids, const int32_t* wb_n_dev, int32_t wb_bound,
                              const float* wb_rows, int64_t ld_wb, int32_t* counters, int32_t* consumed,
                              void* stream);
int sgcn_wb_counters_reset(int32_t* counters, void* stream);
/* The NEXT sgcn_cv_sampled_fwd[_bwd] / sgcn_cvd_sampled_fwd[_bwd] launch of this host thread adds 1 to
 * counters[2] when its last thread block has finished (its reads of hist[tgt] are over). */
int sgcn_sampled_done_attach(int32_t* counters);

/* The schedule bench.py times: n passes with
 *   samp  : trains of `train` batches sampled by ONE launch each (sgcn_sampler_expand_train), one train ahead
 *           of the passes that consume them (first_train: length of the first train, 0 = `train`; a short
 *           first train shortens the start-up bubble)
 *   pre   : gather + dX init + output zeroing of pass k+1 while pass k runs (three x0 copies, two dx copies)
 *   chain : full-neighbour mean(k) -> write-back(k) -> full-neighbour mean(k+1) ...
 *   side  : [write-back(k-1)] sampled aggregate + backward(k)
 * With desc.fuse_write_back (single GPU, CV / CVD) the write-back of pass k is carried by the tail of the pass's
 * own full-neighbour-mean launch (sgcn_full_history_mean_wb): the chain is mean(k) -> mean(k+1) and nothing else.
 * Same results as n sequential passes.  ids / ids_on_host / out_host as
 * sgcn_step_run; every internal stream forks from and joins `stream` (capturable into a CUDA graph).  Pass k:
 * aggregated rows in desc.out[k & 1], gathered rows in x0 copy k % 3 (desc.x0, x0_alt[0], x0_alt[1]), dX in
 * (k & 1 ? dx_alt : desc.dx), sampler buffer set  ((train index & 1) * train + position in the train). */
int sgcn_step_run_trains(sgcn_step* st, const int32_t* ids, int32_t ids_on_host, int32_t n, float* out_host,
                         int32_t first_train, void* stream);

/* *timed_out (HOST) != 0: a device-side wait of the fused write-back gave up (bounded spins) since the last run
 * started -- its results are not to be trusted.  Synchronises the device. */
int sgcn_step_status(sgcn_step* st, int32_t* timed_out);

#ifdef __cplusplus
}
#endif
#endif /* SGCN_B200_H */
