// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C-ABI shim around the UNMODIFIED reference sources under /root/reference/gcn
// (scheduler.{h,cpp}, mult.{h,cpp}, history.{h,cpp}).  The reference sources are compiled where
// they lie (see oracle/Makefile); nothing from them is copied into this repository.  The shim
// only forwards calls and exposes the public std::vector members of the reference `Scheduler`
// (gcn/scheduler.h:17-27) as (pointer, length) pairs so that tests can compare them bit for bit
// with the oracle restatement (oracle/sgcn_oracle.c) and with the CUDA path.
//
// Built into oracle/_ref/libsgcn_ref.so (git-ignored, NOT gpurun-ignored).
#include "scheduler.h"   // -I/root/reference/gcn
#include "history.h"
#include "mult.h"

#include <cstring>

extern "C" {

// ---- Scheduler (gcn/scheduler.cpp:11-189) --------------------------------------------------
void* ref_sched_create(float* adj_w, int* adj_i, int* adj_p, int num_data, int num_edges,
                       int L, int cv, int is) {
    return new Scheduler(adj_w, adj_i, adj_p, num_data, num_edges, L, cv != 0, is != 0);
}
void ref_sched_destroy(void* h) { delete static_cast<Scheduler*>(h); }
void ref_sched_seed(void* h, int seed) { static_cast<Scheduler*>(h)->seed(seed); }
void ref_sched_start_batch(void* h, int n, int* data) {
    static_cast<Scheduler*>(h)->start_batch(n, data);
}
// The reference throws std::runtime_error from expand ("nan", scheduler.cpp:114-115; "Prob is empty",
// mult.cpp:17-18), which would std::terminate through Cython (_scheduler.pyx:12-13 declares `except +`
// on the constructors only).  The shim turns it into a status so that a fuzzer can walk over such inputs:
// 0 = ok, -4 = the reference threw (the Scheduler is then in an unspecified state).
int ref_sched_expand(void* h, int degree) {
    try {
        static_cast<Scheduler*>(h)->expand(degree);
        return 0;
    } catch (...) {
        return -4;
    }
}

// which: 0 field, 1 ffield, 2 edg_s, 3 edg_t, 4 fedg_s, 5 fedg_t, 6 adj_i, 7 adj_p, 8 visited, 9 fvisited
int ref_sched_int_vec(void* h, int which, const int** out) {
    Scheduler* s = static_cast<Scheduler*>(h);
    const std::vector<int>* v = nullptr;
    switch (which) {
        case 0: v = &s->field; break;
        case 1: v = &s->ffield; break;
        case 2: v = &s->edg_s; break;
        case 3: v = &s->edg_t; break;
        case 4: v = &s->fedg_s; break;
        case 5: v = &s->fedg_t; break;
        case 6: v = &s->adj_i; break;
        case 7: v = &s->adj_p; break;
        case 8: v = &s->visited; break;
        case 9: v = &s->fvisited; break;
        default: *out = nullptr; return -1;
    }
    *out = v->data();
    return (int)v->size();
}
// which: 0 scales, 1 edg_w, 2 medg_w, 3 fedg_w, 4 adj_w, 5 importance
int ref_sched_float_vec(void* h, int which, const float** out) {
    Scheduler* s = static_cast<Scheduler*>(h);
    const std::vector<float>* v = nullptr;
    switch (which) {
        case 0: v = &s->scales; break;
        case 1: v = &s->edg_w; break;
        case 2: v = &s->medg_w; break;
        case 3: v = &s->fedg_w; break;
        case 4: v = &s->adj_w; break;
        case 5: v = &s->importance; break;
        default: *out = nullptr; return -1;
    }
    *out = v->data();
    return (int)v->size();
}

// ---- Mult (gcn/mult.cpp:7-51) -----------------------------------------------------------------
void* ref_mult_create(const float* prob, int n) {
    try {
        return new Mult(std::vector<float>(prob, prob + n));
    } catch (...) {
        return nullptr;   // "Prob is empty" (gcn/mult.cpp:17-18)
    }
}
void ref_mult_destroy(void* h) { delete static_cast<Mult*>(h); }
int ref_mult_query(void* h) { return static_cast<Mult*>(h)->Query(); }
int ref_mult_query_u(void* h, float u) { return static_cast<Mult*>(h)->Query(u); }
int ref_mult_bit(void* h, const float** out) {
    Mult* m = static_cast<Mult*>(h);
    *out = m->bit.data();
    return (int)m->bit.size();
}

// ---- row slicers (gcn/history.cpp:50-88) --------------------------------------------------------
void ref_c_indptr(int N, int* r, int* a_p, int* o_p) { c_indptr(N, r, a_p, o_p); }
void ref_c_slice(int N, int* r, float* a_d, int* a_i, int* a_p, float* o_d, int* o_i, int* o_p) {
    c_slice(N, r, a_d, a_i, a_p, o_d, o_i, o_p);
}
void ref_c_dense_slice(int N, int C, int* r, float* i_data, float* o_data) {
    c_dense_slice(N, C, r, i_data, o_data);
}

}  // extern "C"
