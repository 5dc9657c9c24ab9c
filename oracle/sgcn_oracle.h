/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the thu-ml/stochastic_gcn hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (stochastic_gcn_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Parity pin: the sampler / Mult / slicer parts are checked bit-for-bit against the compiled,
 * unmodified reference (oracle/_ref/libsgcn_ref.so) and against the golden vectors under
 * tests/golden/.  The numeric aggregate (SpMM, gather, scatter) restates TensorFlow-1.x ops that
 * are NOT in /root/reference (tf.sparse_tensor_dense_matmul, tf.gather, tf.scatter_update; TF is
 * unpinned, README.md:8) -- for those the reference holds no test or golden vector, so that part
 * is "parity unpinned" (see DESIGN.md).
 */
#ifndef SGCN_ORACLE_H
#define SGCN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- std::mt19937 + std::uniform_real_distribution<float> as libstdc++ 13 evaluates them ---- */
typedef struct {
    uint32_t x[624];
    int pos;
} orc_mt19937;
void orc_mt_seed(orc_mt19937* g, uint32_t seed);   /* std::mt19937::seed, gcn/scheduler.cpp:37-39 */
uint32_t orc_mt_next(orc_mt19937* g);
float orc_mt_canonical(orc_mt19937* g);            /* u01(generator), gcn/scheduler.cpp:8,142 */
float orc_u32_to_canonical(uint32_t r);

/* ---- struct Mult, gcn/mult.h:8-27, gcn/mult.cpp:7-51 ---- */
typedef struct orc_mult orc_mult;
orc_mult* orc_mult_create(const float* prob, int n);   /* NULL when n == 0 ("Prob is empty") */
void orc_mult_destroy(orc_mult* m);
int orc_mult_draw(orc_mult* m);                    /* Mult::Query()   gcn/mult.cpp:30-36 */
int orc_mult_descend(const orc_mult* m, float u);  /* Mult::Query(u)  gcn/mult.cpp:38-51 */
int orc_mult_tree(const orc_mult* m, const float** out);

/* ---- class Scheduler, gcn/scheduler.h:6-28, gcn/scheduler.cpp:11-189 ---- */
typedef struct orc_sampler orc_sampler;
orc_sampler* orc_sampler_create(const float* adj_w, const int* adj_i, const int* adj_p,
                                int num_data, int num_edges, int cv, int is);
void orc_sampler_destroy(orc_sampler* s);
void orc_sampler_seed(orc_sampler* s, int seed);
void orc_sampler_start_batch(orc_sampler* s, int n, const int* ids);
/* returns 0, or -1 for the reference's runtime_error("nan") / "Prob is empty" */
int orc_sampler_expand(orc_sampler* s, int degree);
/* which: 0 field, 1 ffield, 2 edg_s, 3 edg_t, 4 fedg_s, 5 fedg_t, 6 adj_i, 7 adj_p, 8 visited, 9 fvisited */
int orc_sampler_int_vec(orc_sampler* s, int which, const int** out);
/* which: 0 scales, 1 edg_w, 2 medg_w, 3 fedg_w, 4 adj_w, 5 importance */
int orc_sampler_float_vec(orc_sampler* s, int which, const float** out);

/* ---- row slicers, gcn/history.cpp:50-88 ---- */
void orc_slice_indptr(int n, const int* rows, const int* a_p, int* o_p);
void orc_slice_rows(int n, const int* rows, const float* a_d, const int* a_i, const int* a_p,
                    float* o_d, int* o_i2, const int* o_p);
void orc_dense_slice(int n, int c, const int* rows, const float* src, float* dst);

/* ---- numeric aggregate (restates TF ops called at gcn/layers.py:34,211,304-305,354-355 and
 *      gcn/models.py:165; fp32, sequential in nnz order like the TF CPU kernel) ---- */
void orc_spmm_coo(int nnz, const int* rows, const int* cols, const float* vals,
                  const float* x, int d, float* y, int n_rows);
void orc_spmm_coo_t(int nnz, const int* rows, const int* cols, const float* vals,
                    const float* dy, int d, float* dx, int n_cols);
void orc_gather_rows(int n, int d, const int* idx, const float* table, float* out);
void orc_scatter_rows(int n, int d, const int* idx, const float* rows, float* table);
/* row-parallel (OpenMP) CSR form of the same product for the timed CPU baseline; the row order is
 * the authors' own commented-out compute_history loop, gcn/history.cpp:10-37 */
void orc_spmm_csr_omp(int n_rows, const int* rowptr, const int* cols, const float* vals,
                      const float* x, int d, float* y, int threads);
/* fused CV forward for the timed CPU baseline: z = A(x - H[ifield]) + Af H[ffield-as-global-ids]
 * with Af given as the global CSR rows of the output field (gcn/layers.py:350-358) */
void orc_cv_forward_omp(int n_out, const int* rowptr_s, const int* cols_s, const float* vals_s,
                        const float* x, const int* ifield, const float* hist, int d,
                        const int* out_nodes, const int* adj_p, const int* adj_i, const float* adj_w,
                        float* z, int threads);

#ifdef __cplusplus
}
#endif
#endif
