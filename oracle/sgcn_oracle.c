/* TEST INFRASTRUCTURE ONLY -- see sgcn_oracle.h.
 *
 * Plain-C restatement of the hot path of thu-ml/stochastic_gcn.  It is written from the
 * behaviour of the reference (cited per function), not transcribed from it: the data structures
 * here are flat growable arrays and an explicit Mersenne-Twister, where the reference uses
 * std::vector and <random>.  Float arithmetic is kept operation-for-operation identical
 * (compile with -ffp-contract=off and without -march=native, as gcn/setup.py:12 does for the
 * `scheduler` extension) so that every output can be compared bit for bit.
 */
#include "sgcn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* =============================== MT19937 ======================================= */

/* std::mt19937::seed(value): x0 = value mod 2^32, x_i = 1812433253 (x_{i-1} ^ (x_{i-1}>>30)) + i.
 * After seeding the engine's cursor sits at 624, i.e. the first draw regenerates the block. */
void orc_mt_seed(orc_mt19937* g, uint32_t seed) {
    g->x[0] = seed;
    for (int i = 1; i < 624; ++i) {
        uint32_t prev = g->x[i - 1];
        g->x[i] = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
    }
    g->pos = 624;
}

static void mt_refill(orc_mt19937* g) {
    uint32_t* x = g->x;
    for (int k = 0; k < 624; ++k) {
        uint32_t y = (x[k] & 0x80000000u) | (x[(k + 1) % 624] & 0x7fffffffu);
        uint32_t v = x[(k + 397) % 624] ^ (y >> 1);
        if (y & 1u) v ^= 0x9908b0dfu;
        x[k] = v;
    }
    g->pos = 0;
}

uint32_t orc_mt_next(orc_mt19937* g) {
    if (g->pos >= 624) mt_refill(g);
    uint32_t z = g->x[g->pos++];
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}

/* std::generate_canonical<float,24>(mt19937) as libstdc++ 13 computes it
 * (/usr/include/c++/13/bits/random.tcc:3349-3378): one 32-bit draw, converted to float
 * (round to nearest even), divided by 2^32 in float, and clamped to nextafterf(1,0) if the
 * conversion rounded up to 2^32.  uniform_real_distribution<float>(0,1) then returns u*1+0. */
float orc_u32_to_canonical(uint32_t r) {
    float u = (float)r / 4294967296.0f;
    if (u >= 1.0f) u = nextafterf(1.0f, 0.0f);
    return u;
}

float orc_mt_canonical(orc_mt19937* g) { return orc_u32_to_canonical(orc_mt_next(g)); }

/* =============================== Mult ========================================= */

struct orc_mult {
    int n;          /* number of categories */
    int cap;        /* tree size: n rounded up by repeatedly adding the low bit (gcn/mult.cpp:8-9) */
    float* p;       /* remaining probability mass per category */
    float* tree;    /* Fenwick array, 1-based, cap+1 entries */
    float total;
    orc_mt19937 gen; /* default-constructed std::mt19937 => seed 5489 (gcn/mult.h:26) */
};

static int low_bit(int v) { return v & (-v); }

/* Mult::Add, gcn/mult.cpp:22-28 */
static void mult_bump(orc_mult* m, int pos1, float delta) {
    for (int k = pos1; k <= m->cap; k += low_bit(k)) m->tree[k] += delta;
    m->total += delta;
}

/* Mult::Mult, gcn/mult.cpp:7-20.  Insertion is left to right, one Add per category, so the
 * float partial sums in the tree are order dependent and reproduced exactly. */
orc_mult* orc_mult_create(const float* prob, int n) {
    if (n <= 0) return NULL; /* reference throws runtime_error("Prob is empty") */
    orc_mult* m = (orc_mult*)calloc(1, sizeof(orc_mult));
    m->n = n;
    int cap = n;
    while (cap != low_bit(cap)) cap += low_bit(cap);
    m->cap = cap;
    m->p = (float*)malloc(sizeof(float) * (size_t)n);
    memcpy(m->p, prob, sizeof(float) * (size_t)n);
    m->tree = (float*)calloc((size_t)cap + 1, sizeof(float));
    m->total = 0.0f;
    for (int i = 0; i < n; ++i) mult_bump(m, i + 1, prob[i]);
    orc_mt_seed(&m->gen, 5489u);
    return m;
}

void orc_mult_destroy(orc_mult* m) {
    if (!m) return;
    free(m->p);
    free(m->tree);
    free(m);
}

/* Mult::Query(float), gcn/mult.cpp:38-51: binary descent over the implicit tree.  The result
 * may equal n..cap for u at/above the total mass (the reference's own test prints 8 for
 * Query(14) on 5 categories, gcn/test_mult.cpp:28); Query() clamps it. */
int orc_mult_descend(const orc_mult* m, float u) {
    int at = 0;
    for (int span = m->cap; span > 0; span /= 2) {
        int nxt = at + span;
        if (nxt <= m->cap && !(m->tree[nxt] > u)) {
            u -= m->tree[nxt];
            at = nxt;
        }
    }
    return at;
}

/* Mult::Query(), gcn/mult.cpp:30-36: draw, clamp to n-1, remove the drawn mass. */
int orc_mult_draw(orc_mult* m) {
    float u = orc_mt_canonical(&m->gen) * m->total;
    int r = orc_mult_descend(m, u);
    if (r > m->n - 1) r = m->n - 1;
    mult_bump(m, r + 1, -m->p[r]);
    m->p[r] = 0.0f;
    return r;
}

int orc_mult_tree(const orc_mult* m, const float** out) {
    *out = m->tree;
    return m->cap + 1;
}

/* =============================== Scheduler ==================================== */

typedef struct { int* v; int n, cap; } ivec;
typedef struct { float* v; int n, cap; } fvec;

static void ipush(ivec* a, int x) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 64;
        a->v = (int*)realloc(a->v, sizeof(int) * (size_t)a->cap);
    }
    a->v[a->n++] = x;
}
static void fpush(fvec* a, float x) {
    if (a->n == a->cap) {
        a->cap = a->cap ? a->cap * 2 : 64;
        a->v = (float*)realloc(a->v, sizeof(float) * (size_t)a->cap);
    }
    a->v[a->n++] = x;
}

struct orc_sampler {
    int n_nodes, n_edges, cv, is;
    /* private, MUTABLE copy of the CSR (gcn/scheduler.cpp:14-16,20): the uniform sampler
     * permutes row entries in place and the permutation persists across batches. */
    int* ptr;     /* n_nodes + 1 */
    int* col;     /* n_edges */
    float* val;   /* n_edges */
    int* slot;    /* "visited": position of a node in the growing field, -1 if absent */
    int* fslot;   /* "fvisited": position of a node in ffield, -1 if absent */
    float* imp;   /* importance per node */
    ivec field, grown, ffield, es, et, fes, fet;
    fvec scales, ew, mew, few;
    orc_mt19937 gen;
};

/* Scheduler::Scheduler, gcn/scheduler.cpp:11-35 */
orc_sampler* orc_sampler_create(const float* adj_w, const int* adj_i, const int* adj_p,
                                int num_data, int num_edges, int cv, int is) {
    orc_sampler* s = (orc_sampler*)calloc(1, sizeof(orc_sampler));
    s->n_nodes = num_data;
    s->n_edges = num_edges;
    s->cv = cv;
    s->is = is;
    s->ptr = (int*)malloc(sizeof(int) * ((size_t)num_data + 1));
    memcpy(s->ptr, adj_p, sizeof(int) * (size_t)num_data);
    s->ptr[num_data] = num_edges;
    s->col = (int*)malloc(sizeof(int) * (size_t)(num_edges > 0 ? num_edges : 1));
    s->val = (float*)malloc(sizeof(float) * (size_t)(num_edges > 0 ? num_edges : 1));
    memcpy(s->col, adj_i, sizeof(int) * (size_t)num_edges);
    memcpy(s->val, adj_w, sizeof(float) * (size_t)num_edges);
    s->slot = (int*)malloc(sizeof(int) * (size_t)num_data);
    s->fslot = (int*)malloc(sizeof(int) * (size_t)num_data);
    s->imp = (float*)malloc(sizeof(float) * (size_t)num_data);
    for (int i = 0; i < num_data; ++i) {
        s->slot[i] = -1;
        s->fslot[i] = -1;
        s->imp[i] = (float)1e-6;
    }
    if (is) {
        /* column sums of squared weights, accumulated in CSR order (scheduler.cpp:22-25) */
        for (int r = 0; r < num_data; ++r)
            for (int e = s->ptr[r]; e < s->ptr[r + 1]; ++e) {
                float sq = s->val[e] * s->val[e];
                s->imp[s->col[e]] += sq;
            }
    } else {
        for (int i = 0; i < num_data; ++i) s->imp[i] = 1.0f;
    }
    orc_mt_seed(&s->gen, 5489u); /* default-constructed engine until seed() is called */
    return s;
}

void orc_sampler_destroy(orc_sampler* s) {
    if (!s) return;
    free(s->ptr); free(s->col); free(s->val); free(s->slot); free(s->fslot); free(s->imp);
    free(s->field.v); free(s->grown.v); free(s->ffield.v);
    free(s->es.v); free(s->et.v); free(s->fes.v); free(s->fet.v);
    free(s->scales.v); free(s->ew.v); free(s->mew.v); free(s->few.v);
    free(s);
}

/* Scheduler::seed, gcn/scheduler.cpp:37-39 */
void orc_sampler_seed(orc_sampler* s, int seed) { orc_mt_seed(&s->gen, (uint32_t)seed); }

/* Scheduler::start_batch, gcn/scheduler.cpp:41-44 */
void orc_sampler_start_batch(orc_sampler* s, int n, const int* ids) {
    s->field.n = 0;
    for (int i = 0; i < n; ++i) ipush(&s->field, ids[i]);
}

/* Importance-sampling branch of Scheduler::expand, gcn/scheduler.cpp:63-123 */
static int expand_importance(orc_sampler* s, int degree) {
    const int n_out = s->field.n;
    int* pool = NULL;        /* distinct neighbours of the field in first-seen order */
    float* mass = NULL;
    int n_pool = 0, pool_cap = 0;
    unsigned char* seen = (unsigned char*)calloc((size_t)s->n_nodes, 1);
    int* hits = (int*)calloc((size_t)s->n_nodes, sizeof(int));
    float mass_total = 0.0f;

    for (int i = 0; i < n_out; ++i) {
        int node = s->field.v[i];
        for (int e = s->ptr[node]; e < s->ptr[node + 1]; ++e) {
            int t = s->col[e];
            if (seen[t]) continue;
            seen[t] = 1;
            if (n_pool == pool_cap) {
                pool_cap = pool_cap ? pool_cap * 2 : 64;
                pool = (int*)realloc(pool, sizeof(int) * (size_t)pool_cap);
                mass = (float*)realloc(mass, sizeof(float) * (size_t)pool_cap);
            }
            pool[n_pool] = t;
            mass_total += s->imp[t];
            mass[n_pool] = s->imp[t];
            ++n_pool;
        }
    }

    int status = 0;
    orc_mult* m = orc_mult_create(mass, n_pool);
    if (!m) {
        status = -1; /* "Prob is empty": the reference would terminate */
    } else {
        /* min(field.size()*degree, neighbors.size()) evaluated in size_t, stored in an int */
        size_t want = (size_t)n_out * (size_t)degree;
        int n_draw = (int)(want < (size_t)n_pool ? want : (size_t)n_pool);
        for (int k = 0; k < n_draw; ++k) {
            int t = pool[orc_mult_draw(m)];
            hits[t] += 1;
            if (s->slot[t] == -1) {
                s->slot[t] = s->grown.n;
                ipush(&s->grown, t);
            }
        }
        for (int i = 0; i < n_out && status == 0; ++i) {
            int node = s->field.v[i];
            for (int e = s->ptr[node]; e < s->ptr[node + 1]; ++e) {
                int t = s->col[e];
                if (!hits[t]) continue;
                /* times*w*total / (importance*num_samples), all in float, this order */
                float num = (float)hits[t] * s->val[e];
                num = num * mass_total;
                float den = s->imp[t] * (float)n_draw;
                float w = num / den;
                ipush(&s->es, i);
                ipush(&s->et, s->slot[t]);
                fpush(&s->ew, w);
                if (isnan(w)) { status = -1; break; }
            }
        }
        orc_mult_destroy(m);
    }
    free(pool); free(mass); free(seen); free(hits);
    return status;
}

/* Uniform (NS / CV) branch of Scheduler::expand, gcn/scheduler.cpp:125-180 */
static void expand_uniform(orc_sampler* s, int degree) {
    const int n_out = s->field.n;
    for (int i = 0; i < n_out; ++i) {
        const int node = s->field.v[i];
        int* rc = s->col + s->ptr[node];
        float* rw = s->val + s->ptr[node];
        const int deg = s->ptr[node + 1] - s->ptr[node];
        const int take = deg < degree ? deg : degree;
        float scale = (float)deg / (float)take;     /* 0/0 -> NaN, replaced below */
        if (deg == 0) scale = 1.0f;
        fpush(&s->scales, (float)(1.0 / (double)sqrtf(scale)));

        /* partial Fisher-Yates: position k receives a uniformly chosen element of [k, deg) */
        for (int k = 0; k < take; ++k) {
            float u = orc_mt_canonical(&s->gen);
            float span = (float)(deg - k) * u;
            float where = (float)k + span;
            int j = (int)where;
            if (j > deg - 1) j = deg - 1;
            int tc = rc[k]; rc[k] = rc[j]; rc[j] = tc;
            float tw = rw[k]; rw[k] = rw[j]; rw[j] = tw;

            int t = rc[k];
            float w = rw[k] * scale;
            if (s->slot[t] == -1) {
                s->slot[t] = s->grown.n;
                ipush(&s->grown, t);
            }
            ipush(&s->es, i);
            ipush(&s->et, s->slot[t]);
            fpush(&s->ew, w);
            if (s->cv) fpush(&s->mew, rw[k] * w);
        }

        if (s->cv) {
            /* the whole row, in its post-swap order, goes to the full-neighbour adjacency */
            for (int k = 0; k < deg; ++k) {
                int t = rc[k];
                if (s->fslot[t] == -1) {
                    s->fslot[t] = s->ffield.n;
                    ipush(&s->ffield, t);
                }
                ipush(&s->fes, i);
                ipush(&s->fet, s->fslot[t]);
                fpush(&s->few, rw[k]);
            }
        }
    }
}

/* Scheduler::expand, gcn/scheduler.cpp:46-61,182-189 */
int orc_sampler_expand(orc_sampler* s, int degree) {
    s->grown.n = 0;
    s->ffield.n = 0;
    for (int i = 0; i < s->field.n; ++i) ipush(&s->grown, s->field.v[i]);
    for (int i = 0; i < s->grown.n; ++i) s->slot[s->grown.v[i]] = i;
    s->es.n = s->et.n = s->fes.n = s->fet.n = 0;
    s->ew.n = s->mew.n = s->few.n = s->scales.n = 0;

    int status = 0;
    if (s->is) status = expand_importance(s, degree);
    else expand_uniform(s, degree);

    /* the grown field (old field as prefix) becomes the current field */
    ivec tmp = s->field; s->field = s->grown; s->grown = tmp;
    for (int i = 0; i < s->field.n; ++i) s->slot[s->field.v[i]] = -1;
    if (!s->is && s->cv)
        for (int i = 0; i < s->ffield.n; ++i) s->fslot[s->ffield.v[i]] = -1;
    return status;
}

int orc_sampler_int_vec(orc_sampler* s, int which, const int** out) {
    switch (which) {
        case 0: *out = s->field.v; return s->field.n;
        case 1: *out = s->ffield.v; return s->ffield.n;
        case 2: *out = s->es.v; return s->es.n;
        case 3: *out = s->et.v; return s->et.n;
        case 4: *out = s->fes.v; return s->fes.n;
        case 5: *out = s->fet.v; return s->fet.n;
        case 6: *out = s->col; return s->n_edges;
        case 7: *out = s->ptr; return s->n_nodes + 1;
        case 8: *out = s->slot; return s->n_nodes;
        case 9: *out = s->fslot; return s->n_nodes;
    }
    *out = NULL;
    return -1;
}

int orc_sampler_float_vec(orc_sampler* s, int which, const float** out) {
    switch (which) {
        case 0: *out = s->scales.v; return s->scales.n;
        case 1: *out = s->ew.v; return s->ew.n;
        case 2: *out = s->mew.v; return s->mew.n;
        case 3: *out = s->few.v; return s->few.n;
        case 4: *out = s->val; return s->n_edges;
        case 5: *out = s->imp; return s->n_nodes;
    }
    *out = NULL;
    return -1;
}

/* =============================== row slicers =================================== */

/* c_indptr, gcn/history.cpp:50-57 */
void orc_slice_indptr(int n, const int* rows, const int* a_p, int* o_p) {
    int run = 0;
    for (int i = 0; i < n; ++i) {
        o_p[i] = run;
        run += a_p[rows[i] + 1] - a_p[rows[i]];
    }
    o_p[n] = run;
}

/* c_slice, gcn/history.cpp:59-72: values copied, indices written as (local row, column) pairs */
void orc_slice_rows(int n, const int* rows, const float* a_d, const int* a_i, const int* a_p,
                    float* o_d, int* o_i2, const int* o_p) {
    for (int i = 0; i < n; ++i) {
        int len = o_p[i + 1] - o_p[i];
        int src = a_p[rows[i]];
        for (int k = 0; k < len; ++k) {
            o_d[o_p[i] + k] = a_d[src + k];
            o_i2[2 * (o_p[i] + k)] = i;
            o_i2[2 * (o_p[i] + k) + 1] = a_i[src + k];
        }
    }
}

/* c_dense_slice, gcn/history.cpp:74-88 */
void orc_dense_slice(int n, int c, const int* rows, const float* src, float* dst) {
    for (int i = 0; i < n; ++i)
        memcpy(dst + (size_t)i * c, src + (size_t)rows[i] * c, sizeof(float) * (size_t)c);
}

/* =============================== numeric aggregate ============================= */

/* tf.sparse_tensor_dense_matmul(A, X) (gcn/layers.py:31-37): y[r,:] += v * x[c,:] for each
 * stored entry, in storage order, fp32 multiply then add. */
void orc_spmm_coo(int nnz, const int* rows, const int* cols, const float* vals,
                  const float* x, int d, float* y, int n_rows) {
    memset(y, 0, sizeof(float) * (size_t)n_rows * (size_t)d);
    for (int e = 0; e < nnz; ++e) {
        const float v = vals[e];
        const float* xs = x + (size_t)cols[e] * d;
        float* yd = y + (size_t)rows[e] * d;
        for (int k = 0; k < d; ++k) yd[k] += v * xs[k];
    }
}

/* gradient of the above w.r.t. X (TF autodiff, gcn/models.py:187): dx[c,:] += v * dy[r,:] */
void orc_spmm_coo_t(int nnz, const int* rows, const int* cols, const float* vals,
                    const float* dy, int d, float* dx, int n_cols) {
    memset(dx, 0, sizeof(float) * (size_t)n_cols * (size_t)d);
    for (int e = 0; e < nnz; ++e) {
        const float v = vals[e];
        const float* g = dy + (size_t)rows[e] * d;
        float* o = dx + (size_t)cols[e] * d;
        for (int k = 0; k < d; ++k) o[k] += v * g[k];
    }
}

/* tf.gather(table, idx), gcn/layers.py:211,304-305,354-355 */
void orc_gather_rows(int n, int d, const int* idx, const float* table, float* out) {
    for (int i = 0; i < n; ++i)
        memcpy(out + (size_t)i * d, table + (size_t)idx[i] * d, sizeof(float) * (size_t)d);
}

/* tf.scatter_update(table, idx, rows), gcn/models.py:160-166 (overwrite; idx unique) */
void orc_scatter_rows(int n, int d, const int* idx, const float* rows, float* table) {
    for (int i = 0; i < n; ++i)
        memcpy(table + (size_t)idx[i] * d, rows + (size_t)i * d, sizeof(float) * (size_t)d);
}

void orc_spmm_csr_omp(int n_rows, const int* rowptr, const int* cols, const float* vals,
                      const float* x, int d, float* y, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
    for (int r = 0; r < n_rows; ++r) {
        float* yd = y + (size_t)r * d;
        for (int k = 0; k < d; ++k) yd[k] = 0.0f;
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
            const float v = vals[e];
            const float* xs = x + (size_t)cols[e] * d;
            for (int k = 0; k < d; ++k) yd[k] += v * xs[k];
        }
    }
}

void orc_cv_forward_omp(int n_out, const int* rowptr_s, const int* cols_s, const float* vals_s,
                        const float* x, const int* ifield, const float* hist, int d,
                        const int* out_nodes, const int* adj_p, const int* adj_i, const float* adj_w,
                        float* z, int threads) {
    (void)threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
    for (int r = 0; r < n_out; ++r) {
        float* zd = z + (size_t)r * d;
        for (int k = 0; k < d; ++k) zd[k] = 0.0f;
        for (int e = rowptr_s[r]; e < rowptr_s[r + 1]; ++e) {
            const float v = vals_s[e];
            const float* xs = x + (size_t)cols_s[e] * d;
            const float* hs = hist + (size_t)ifield[cols_s[e]] * d;
            for (int k = 0; k < d; ++k) zd[k] += v * (xs[k] - hs[k]);
        }
        const int node = out_nodes[r];
        for (int e = adj_p[node]; e < adj_p[node + 1]; ++e) {
            const float v = adj_w[e];
            const float* hs = hist + (size_t)adj_i[e] * d;
            for (int k = 0; k < d; ++k) zd[k] += v * hs[k];
        }
    }
}
