// Native step driver: n consecutive passes of the hot path issued from C++ onto three CUDA streams,
// with the samplers of batches k+1 / k+2 running beside the aggregate of batch k.
//
// Why not only CUDA graphs: on this driver a graph launch expands its nodes on the device at ~1.8 us
// per kernel node, branch by branch, AFTER the previous launch on the stream has finished
// (tools/graph_overhead.py, tools/timeline.py) -- for a step of six short kernels next to one 20 us
// kernel that is ~10 us of pure turnaround per step.  Plain stream launches keep every queue fed in
// step order: the CPU costs ~15 us per step and runs ahead of the ~25 us the GPU needs.
//
//   chain : [wait sampler k] full_mean(k) ─► [wait fwd k] history_update(k) ─► ...
//   side  : [wait sampler k, rest k-1] dX init + zero(out k+1) ─► gather(k) ─► (publish) ─► fwd+bwd(k) ─► D2H(k)
//   samp  : [wait rest k-1] (H2D ids k+2) ─► expand(k+2) into the buffer set pass k-1 used
//
// Semantics are exactly those of n sequential passes (same sampler order and RNG stream, every
// forward read of history before the write-back, write-back k before any read of pass k+1); the
// in-place row permutation is guarded on the device (sgcn_sampler_pipeline).
#include <vector>

#include "common.cuh"

struct sgcn_step {
    sgcn_sampler* sampler = nullptr;
    sgcn_step_desc d{};
    cudaStream_t chain = nullptr, side = nullptr, samp = nullptr, copy = nullptr, pre = nullptr;
    static constexpr int kSlots = 3;               // sampler buffer sets: two batches of lookahead
    int32_t* ids_dev[kSlots] = {nullptr, nullptr, nullptr};      // staging of host ids
    // level-0 buffers of the sampler's slots
    struct Lv { int32_t *field, *rowptr_s, *rowptr_f, *edg_t, *tgt, *meta; float *edg_w, *scales; } lv[kSlots]{};
    const int32_t* adj_p = nullptr; const int32_t* adj_i = nullptr; const float* adj_w = nullptr;
    int32_t* pipe = nullptr;
    static constexpr int kRing = 4;
    cudaEvent_t ev_samp[kRing]{}, ev_full[kRing]{}, ev_fwd[kRing]{}, ev_rest[kRing]{}, ev_d2h[kRing]{}, ev_pre[kRing]{},
                ev_pre_end = nullptr, ev_begin = nullptr,
                ev_side_end = nullptr, ev_samp_end = nullptr, ev_zero0 = nullptr;
    int device = 0;
    // sgcn_step_run_trains
    int train = 0;                                  // batches per train (0: trains unavailable for this sampler)
    std::vector<Lv> tlv;                            // level-0 buffers of sampler sets 0 .. 2*train-1
    int32_t* ids_stage[2] = {nullptr, nullptr};     // staging of host ids, one per train parity
    static constexpr int kRing2 = 8;
    cudaEvent_t t_pre[kRing2]{}, t_full[kRing2]{}, t_fwd[kRing2]{}, t_rest[kRing2]{}, t_d2h[kRing2]{}, t_train[4]{};
    int32_t* flags = nullptr;                       // counters of the fused write-back (8 ints, sgcn_full_history_mean_wb)
    bool warmed = false;                            // a first run has loaded every kernel (see sgcn_step_run_trains)
    bool gather_after_sampled = false;              // A/B: gather(k+1) ordered behind sampled(k)
    bool split_apply = false;                       // ring exchange: claims off the chain (SGCN_WB_SPLIT=1; measured slower)
    cudaEvent_t t_claim[8]{};
};

namespace sgcn {
#define STEP_TRY(expr)                     \
    do {                                   \
        int rc__ = (expr);                 \
        if (rc__ != SGCN_OK) return rc__;  \
    } while (0)

static int get_slot_vec(sgcn_sampler* s, int slot, int which, void** p) { return sgcn_sampler_slot_vec(s, slot, which, p); }
}  // namespace sgcn

using namespace sgcn;

extern "C" {

int sgcn_step_create(sgcn_step** out, sgcn_sampler* sampler, const sgcn_step_desc* desc) {
    SGCN_REQUIRE(out && sampler && desc, "step_create: null argument");
    *out = nullptr;
    const sgcn_step_desc& d = *desc;
    SGCN_REQUIRE(d.mode >= 0 && d.mode <= 2 && d.batch > 0 && d.degree >= 0 && d.hidden > 0 && d.feat_dim > 0,
                 "step_create: bad sizes");
    SGCN_REQUIRE(d.features && d.x0 && d.out[0] && d.out[1] && d.d_out && d.dx, "step_create: null buffer");
    SGCN_REQUIRE(d.mode == 0 || d.history, "step_create: CV / CVD need a history table");
    SGCN_REQUIRE(d.mode != 2 || (d.out_mu[0] && d.out_mu[1] && d.feat_dim >= 2 * d.hidden),
                 "step_create: CVD needs out_mu buffers and 2*hidden feature columns");
    SGCN_REQUIRE(d.world <= 16, "step_create: at most 16 ranks");
    sgcn_step* st = new sgcn_step();
    st->sampler = sampler;
    st->d = d;
    SGCN_CUDA(cudaGetDevice(&st->device));
    auto fail = [&](int rc) {
        sgcn_step_destroy(st);
        return rc;
    };
    // both buffer sets of the sampler, sized once; their level-0 pointers never move afterwards
    for (int slot = 0; slot < sgcn_step::kSlots; ++slot) {
        int rc = sgcn_sampler_set_slot(sampler, slot);
        if (rc == SGCN_OK) rc = sgcn_sampler_reserve(sampler, d.batch, &d.degree, 1, 0);
        void* p = nullptr;
#define GET(which, field, T)                                                     \
    if (rc == SGCN_OK) { rc = get_slot_vec(sampler, slot, which, &p); st->lv[slot].field = (T*)p; }
        GET(SGCN_VEC_FIELD, field, int32_t)
        GET(SGCN_VEC_ROWPTR_S, rowptr_s, int32_t)
        GET(SGCN_VEC_ROWPTR_F, rowptr_f, int32_t)
        GET(SGCN_VEC_EDG_T, edg_t, int32_t)
        GET(SGCN_VEC_TGT, tgt, int32_t)
        GET(SGCN_VEC_META, meta, int32_t)
        GET(SGCN_VEC_EDG_W, edg_w, float)
        GET(SGCN_VEC_SCALES, scales, float)
#undef GET
        if (rc != SGCN_OK) return fail(rc);
    }
    sgcn_sampler_set_slot(sampler, 0);
    void* p = nullptr;
    int64_t len = 0;
    int rc = sgcn_sampler_vec(sampler, 0, SGCN_VEC_ADJ_P, &p, &len); st->adj_p = (const int32_t*)p;
    if (rc == SGCN_OK) { rc = sgcn_sampler_vec(sampler, 0, SGCN_VEC_ADJ_I, &p, &len); st->adj_i = (const int32_t*)p; }
    if (rc == SGCN_OK) { rc = sgcn_sampler_vec(sampler, 0, SGCN_VEC_ADJ_W, &p, &len); st->adj_w = (const float*)p; }
    if (rc == SGCN_OK) { rc = sgcn_sampler_vec(sampler, 0, SGCN_VEC_PIPE, &p, &len); st->pipe = (int32_t*)p; }
    if (rc != SGCN_OK) return fail(rc);
#define CK(call)                                                                       \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #call, __FILE__, __LINE__)); \
    } while (0)
    // The chain carries the long kernels (a full-neighbour mean fills every SM); everything beside it is short and
    // must slip in whenever its inputs are ready.  The hardware places thread blocks launch by launch within one
    // priority: a launch that waits for room holds up every later one.  So the streams get three priorities:
    // side (the sampled aggregate, which the fused write-back waits for) > pre / samp / copy (gather, sampler) >
    // chain -- when the write-back is fused into the mean (the sampled aggregate is then waited for on the device);
    // otherwise one priority for all (measured 0.3 us per pass faster).  SGCN_STEP_PRIORITY = 0 / 1 / 2 overrides.
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));      // numerically lower = more urgent
    const char* pe = getenv("SGCN_STEP_PRIORITY");
    const int levels = pe ? atoi(pe) : (d.fuse_write_back ? 2 : 0);
    const int p_chain = prio_lo;
    const int p_mid = levels >= 1 ? std::max(prio_hi, prio_lo - 1) : prio_lo;
    const int p_side = levels >= 2 ? std::max(prio_hi, prio_lo - 2) : p_mid;
    CK(cudaStreamCreateWithPriority(&st->chain, cudaStreamNonBlocking, p_chain));
    CK(cudaStreamCreateWithPriority(&st->side, cudaStreamNonBlocking, p_side));
    CK(cudaStreamCreateWithPriority(&st->samp, cudaStreamNonBlocking, p_mid));
    CK(cudaStreamCreateWithPriority(&st->copy, cudaStreamNonBlocking, p_mid));
    CK(cudaStreamCreateWithPriority(&st->pre, cudaStreamNonBlocking, p_mid));
    {
        const char* ge = getenv("SGCN_GATHER_AFTER_SAMPLED");
        st->gather_after_sampled = ge && ge[0] == '1';
        const char* se = getenv("SGCN_WB_SPLIT");
        const int c4 = d.hidden / 4;
        st->split_apply = se && se[0] == '1' && d.hidden % 4 == 0 && c4 <= 32 && 32 % c4 == 0 && d.ld_hist % 4 == 0;
    }
    for (int i = 0; i < sgcn_step::kSlots; ++i) CK(cudaMalloc(&st->ids_dev[i], sizeof(int32_t) * (size_t)d.batch));
    for (int i = 0; i < sgcn_step::kRing; ++i) {
        CK(cudaEventCreateWithFlags(&st->ev_samp[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st->ev_full[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st->ev_fwd[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st->ev_rest[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st->ev_d2h[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st->ev_pre[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < sgcn_step::kRing2; ++i)
        for (cudaEvent_t* e : {&st->t_pre[i], &st->t_full[i], &st->t_fwd[i], &st->t_rest[i], &st->t_d2h[i], &st->t_claim[i]})
            CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < 4; ++i) CK(cudaEventCreateWithFlags(&st->t_train[i], cudaEventDisableTiming));
    CK(cudaMalloc(&st->flags, sizeof(int32_t) * 8));
    CK(cudaMemset(st->flags, 0, sizeof(int32_t) * 8));
    {
        // trains of batches: 2 x train buffer sets (the passes of one train run while the next is sampled)
        const int T = d.train > 0 ? d.train : 16;
        if (T >= 2 && T <= 32 && sgcn_sampler_reserve_sets(sampler, 2 * T, d.batch, d.degree) == SGCN_OK) {
            st->train = T;
            st->tlv.resize((size_t)2 * T);
            for (int slot = 0; slot < 2 * T; ++slot) {
                int rc2 = SGCN_OK;
                void* q = nullptr;
#define GET2(which, field, TY)                                                                   \
    if (rc2 == SGCN_OK) { rc2 = get_slot_vec(sampler, slot, which, &q); st->tlv[(size_t)slot].field = (TY*)q; }
                GET2(SGCN_VEC_FIELD, field, int32_t)
                GET2(SGCN_VEC_ROWPTR_S, rowptr_s, int32_t)
                GET2(SGCN_VEC_ROWPTR_F, rowptr_f, int32_t)
                GET2(SGCN_VEC_EDG_T, edg_t, int32_t)
                GET2(SGCN_VEC_TGT, tgt, int32_t)
                GET2(SGCN_VEC_META, meta, int32_t)
                GET2(SGCN_VEC_EDG_W, edg_w, float)
                GET2(SGCN_VEC_SCALES, scales, float)
#undef GET2
                if (rc2 != SGCN_OK) return fail(rc2);
            }
            // the three-set drivers address sets 0..2, which reserve_sets may have re-sized
            for (int slot = 0; slot < sgcn_step::kSlots; ++slot) st->lv[slot] = st->tlv[(size_t)slot];
            for (int i = 0; i < 2; ++i) CK(cudaMalloc(&st->ids_stage[i], sizeof(int32_t) * (size_t)T * (size_t)d.batch));
        }
    }
    CK(cudaEventCreateWithFlags(&st->ev_begin, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&st->ev_pre_end, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&st->ev_side_end, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&st->ev_samp_end, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&st->ev_zero0, cudaEventDisableTiming));
#undef CK
    *out = st;
    return SGCN_OK;
}

void sgcn_step_destroy(sgcn_step* st) {
    if (!st) return;
    for (cudaStream_t s : {st->chain, st->side, st->samp, st->copy, st->pre})
        if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
    for (int i = 0; i < sgcn_step::kSlots; ++i) cudaFree(st->ids_dev[i]);
    for (int i = 0; i < sgcn_step::kRing; ++i)
        for (cudaEvent_t e : {st->ev_samp[i], st->ev_full[i], st->ev_fwd[i], st->ev_rest[i], st->ev_d2h[i], st->ev_pre[i]})
            if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {st->ev_begin, st->ev_side_end, st->ev_samp_end, st->ev_zero0, st->ev_pre_end})
        if (e) cudaEventDestroy(e);
    for (int i = 0; i < sgcn_step::kRing2; ++i)
        for (cudaEvent_t e : {st->t_pre[i], st->t_full[i], st->t_fwd[i], st->t_rest[i], st->t_d2h[i], st->t_claim[i]})
            if (e) cudaEventDestroy(e);
    for (int i = 0; i < 4; ++i) if (st->t_train[i]) cudaEventDestroy(st->t_train[i]);
    for (int i = 0; i < 2; ++i) cudaFree(st->ids_stage[i]);
    cudaFree(st->flags);
    delete st;
}

int sgcn_step_status(sgcn_step* st, int32_t* timed_out) {
    SGCN_REQUIRE(st && timed_out, "step_status: null argument");
    SGCN_CUDA(cudaDeviceSynchronize());
    SGCN_CUDA(cudaMemcpy(timed_out, st->flags + 4, sizeof(int32_t), cudaMemcpyDeviceToHost));
    return SGCN_OK;
}

int sgcn_step_run(sgcn_step* st, const int32_t* ids, int32_t ids_on_host, int32_t n, float* out_host,
                  void* stream) {
    SGCN_REQUIRE(st && n >= 0 && (n == 0 || ids), "step_run: bad argument");
    if (n == 0) return SGCN_OK;
    const sgcn_step_desc& d = st->d;
    sgcn_sampler* smp = st->sampler;
    const int B = d.batch, H = d.hidden, R = sgcn_step::kRing;
    const bool cv = d.mode != 0, cvd = d.mode == 2, concat = d.concat != 0, multi = d.world > 1 && cv;
    const int width = H * (concat ? 2 : 1);
    cudaStream_t user = (cudaStream_t)stream, chain = st->chain, side = st->side, samp = st->samp,
                 copy = st->copy;

    // the neighbour half of out[slot] (and of out_mu[slot]); the self half when concatenating
    auto nb = [&](float* base) { return base + (concat ? H : 0); };

    SGCN_CUDA(cudaEventRecord(st->ev_begin, user));
    for (cudaStream_t s : {chain, side, samp, copy}) SGCN_CUDA(cudaStreamWaitEvent(s, st->ev_begin, 0));

    constexpr int NS = sgcn_step::kSlots;
    auto sample = [&](int k) -> int {         // sampler of batch k into buffer set k % 3, on the samp stream
        const int slot = k % NS;
        const int32_t* src = ids + (int64_t)k * B;
        if (ids_on_host) {
            SGCN_CUDA(cudaMemcpyAsync(st->ids_dev[slot], src, sizeof(int32_t) * (size_t)B,
                                      cudaMemcpyHostToDevice, samp));
            src = st->ids_dev[slot];
        }
        STEP_TRY(sgcn_sampler_set_slot(smp, slot));
        STEP_TRY(sgcn_sampler_start_batch_device(smp, B, src));
        STEP_TRY(sgcn_sampler_expand(smp, d.degree, 0));
        SGCN_CUDA(cudaEventRecord(st->ev_samp[k % R], samp));
        return SGCN_OK;
    };

    STEP_TRY(sgcn_sampler_set_stream_async(smp, samp));
    // forget the batches of earlier runs: the device guard must not chase ids buffers that are gone
    for (int slot = 0; slot < sgcn_step::kSlots; ++slot) {
        STEP_TRY(sgcn_sampler_set_slot(smp, slot));
        STEP_TRY(sgcn_sampler_start_batch_device(smp, 0, nullptr));
    }
    // outputs of pass 0 start from zero (later passes: zeroed one step ahead on the side stream)
    if (cv) {
        STEP_TRY(sgcn_copy_rows_pad_pair(nullptr, 0, 0, nullptr, B, H, nb(d.out[0]), d.ld_out, nullptr, 0, 0, nullptr,
                                         cvd ? B : 0, H, cvd ? nb(d.out_mu[0]) : nullptr, d.ld_out, chain));
        SGCN_CUDA(cudaEventRecord(st->ev_zero0, chain));
        SGCN_CUDA(cudaStreamWaitEvent(side, st->ev_zero0, 0));     // pass 0's sampled part adds into it too
    }
    STEP_TRY(sample(0));
    if (n > 1) STEP_TRY(sample(1));

    for (int k = 0; k < n; ++k) {
        const int r = k & 1, s = 1 - r;                      // output buffers alternate
        const sgcn_step::Lv& v = st->lv[k % NS];             // sampler buffer sets rotate over three
        const int32_t* n_out_dev = v.meta + 0;
        const int32_t* n_in_dev = v.meta + 1;
        float* out_r = d.out[r];
        float* outmu_r = d.out_mu[r];

        // ---- chain: the full-neighbour history mean of pass k ----
        SGCN_CUDA(cudaStreamWaitEvent(chain, st->ev_samp[k % R], 0));
        if (cv) {
            STEP_TRY(sgcn_full_history_mean(v.field, v.rowptr_f, B, n_out_dev, st->adj_p, st->adj_i, st->adj_w,
                                            d.history, d.ld_hist, H, cvd ? nb(outmu_r) : nb(out_r), d.ld_out,
                                            cvd ? nb(out_r) : nullptr, d.ld_out, nullptr, chain));
        }
        if (out_host) SGCN_CUDA(cudaEventRecord(st->ev_full[k % R], chain));

        // ---- samp: sampler of batch k+2 into the buffer set pass k-1 used (free once it has finished):
        //      two batches of lookahead keep the sampler's latency off the chain entirely ----
        if (k + 2 < n) {
            if (k >= 1) SGCN_CUDA(cudaStreamWaitEvent(samp, st->ev_rest[(k - 1) % R], 0));
            STEP_TRY(sample(k + 2));
        }

        // ---- side: dX init + next output zeroing, gather, (publish), sampled aggregate + backward ----
        SGCN_CUDA(cudaStreamWaitEvent(side, st->ev_samp[k % R], 0));
        if (k >= 1) SGCN_CUDA(cudaStreamWaitEvent(side, st->ev_rest[(k - 1) % R], 0));   // x0 / out[s] are free
        const bool zero_next = cv && k + 1 < n;
        // with host output the other buffer is still being copied out (pass k-1): zero it later, below
        const bool zero_early = zero_next && !out_host;
        STEP_TRY(sgcn_copy_rows_pad_pair(concat ? d.d_out : nullptr, d.ld_dout, concat ? B : 0,
                                         concat ? n_out_dev : nullptr, d.x0_rows, H, d.dx, d.ld_dx,
                                         nullptr, 0, 0, nullptr, zero_early ? B : 0, H,
                                         zero_early ? nb(d.out[s]) : nullptr, d.ld_out, side));
        if (zero_early && cvd)
            STEP_TRY(sgcn_copy_rows_pad(nullptr, 0, 0, nullptr, B, H, nb(d.out_mu[s]), d.ld_out, side));
        STEP_TRY(sgcn_gather_rows(d.features, d.ld_feat, v.field, d.x0_rows, n_in_dev, d.feat_dim, d.x0, d.ld_x0, side));
        const float* x = d.x0;
        const float* mu = d.x0 + H;
        const float* new_hist = cvd ? mu : x;
        const float* d_nb = d.d_out + (concat ? H : 0);
        if (multi)
            STEP_TRY(sgcn_wb_push(v.field, n_in_dev, d.wb_bound, new_hist, d.ld_x0, H, d.dst_even, d.dst_odd,
                                  d.world, d.peer_flags, d.rank, d.epoch, d.block_counter, side));
        if (d.mode == 0) {
            STEP_TRY(sgcn_spmm_csr(v.rowptr_s, v.edg_t, v.edg_w, nullptr, B, n_out_dev, x, d.ld_x0, H, nb(out_r),
                                   d.ld_out, 0, side));
            if (concat)
                STEP_TRY(sgcn_copy_rows_pad(x, d.ld_x0, B, n_out_dev, B, H, out_r, d.ld_out, side));
            STEP_TRY(sgcn_spmm_csr_bwd(v.rowptr_s, v.edg_t, v.edg_w, nullptr, B, n_out_dev, d_nb, d.ld_dout, H, d.dx,
                                       d.ld_dx, side));
        } else if (!cvd) {
            STEP_TRY(sgcn_cv_sampled_fwd_bwd(v.rowptr_s, v.edg_t, v.edg_w, v.tgt, B, n_out_dev, x, d.ld_x0, d.history,
                                             d.ld_hist, H, nb(out_r), d.ld_out, concat ? out_r : nullptr, d.ld_out, 1,
                                             d_nb, d.ld_dout, d.dx, d.ld_dx, side));
        } else {
            STEP_TRY(sgcn_cvd_sampled_fwd_bwd(v.rowptr_s, v.edg_t, v.edg_w, v.tgt, v.scales, B, n_out_dev, x, d.ld_x0,
                                              mu, d.ld_x0, d.history, d.ld_hist, H, nb(out_r), d.ld_out,
                                              nb(outmu_r), d.ld_out, concat ? out_r : nullptr, d.ld_out,
                                              concat ? outmu_r : nullptr, d.ld_out, 1, d_nb, d.ld_dout, d.dx,
                                              d.ld_dx, side));
        }
        if (zero_next && !zero_early) {
            // out[s] held pass k-1's rows: zero it for pass k+1 once their D2H copy has finished
            if (k >= 1) SGCN_CUDA(cudaStreamWaitEvent(side, st->ev_d2h[(k - 1) % R], 0));
            STEP_TRY(sgcn_copy_rows_pad_pair(nullptr, 0, 0, nullptr, B, H, nb(d.out[s]), d.ld_out, nullptr, 0, 0,
                                             nullptr, cvd ? B : 0, H, cvd ? nb(d.out_mu[s]) : nullptr, d.ld_out,
                                             side));
        }
        SGCN_CUDA(cudaEventRecord(st->ev_fwd[k % R], side));

        // ---- chain: write-back after every forward read of history (gcn/models.py:186-194) ----
        SGCN_CUDA(cudaStreamWaitEvent(chain, st->ev_fwd[k % R], 0));
        if (!cv) {
            STEP_TRY(sgcn_sampler_mark_consumed(smp, chain));
        } else if (multi) {
            STEP_TRY(sgcn_wb_wait_apply(d.history, d.ld_hist, H, d.recv_even, d.recv_odd, d.slot_bytes, d.world,
                                        d.wb_bound, d.owner, d.flags, d.epoch, d.timeout_flag, st->pipe + 1, chain));
        } else {
            STEP_TRY(sgcn_history_update(d.history, d.ld_hist, v.field, d.x0_rows, n_in_dev, new_hist, d.ld_x0, H,
                                         st->pipe + 1, chain));
        }
        SGCN_CUDA(cudaEventRecord(st->ev_rest[k % R], chain));

        // ---- copy stream: the pass's aggregated rows to pinned host memory (both aggregate kernels done) ----
        if (out_host) {
            SGCN_CUDA(cudaStreamWaitEvent(copy, st->ev_full[k % R], 0));
            SGCN_CUDA(cudaStreamWaitEvent(copy, st->ev_fwd[k % R], 0));
            SGCN_CUDA(cudaMemcpy2DAsync(out_host + (int64_t)k * B * width, sizeof(float) * (size_t)width, out_r,
                                        sizeof(float) * (size_t)d.ld_out, sizeof(float) * (size_t)width, (size_t)B,
                                        cudaMemcpyDeviceToHost, copy));
            SGCN_CUDA(cudaEventRecord(st->ev_d2h[k % R], copy));
        }
    }
    SGCN_CUDA(cudaEventRecord(st->ev_side_end, side));
    SGCN_CUDA(cudaEventRecord(st->ev_samp_end, samp));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_rest[(n - 1) % R], 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_side_end, 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_samp_end, 0));
    if (out_host) SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_d2h[(n - 1) % R], 0));
    return SGCN_OK;
}

// ---- trains of batches + gather ahead + the write-back fused into the full-neighbour mean ------------------
// What the device timelines of the schedule above showed (profiles/r02_timeline_ahead_sampler_bound.txt): with
// the gather hoisted one pass ahead, the per-batch sampler became the critical path -- one CTA, 17-19 us per
// batch under load, strictly serial from batch to batch (engine state), i.e. ~21 us per pass however far ahead
// it runs.  Here a whole train of batches is sampled by ONE launch (sgcn_sampler_expand_train: one CTA per
// batch, serial only in the prefix sum of the draw counts), one train ahead of the passes that consume it:
//   samp  : [rest of train c-2] (H2D ids of train c) -> mt_stream -> expand_train(c)
//   pre   : [train of pass k+1; x0 / dx / out copies free] gather(k+1) + dX init(k+1) + zero out(k+1)
//   chain : [train, pre k] full_mean(k) + write-back(k) in its tail -PDL-> full_mean(k+1) ...   (fuse_write_back)
//   side  : [pre k, chain k-1] sampled fwd+bwd(k)  (signals the tail of full_mean(k) on the device)
// fuse_write_back (single GPU, CV / CVD): the write-back is carried by the last thread blocks of the pass's own
// full-neighbour mean (sgcn_full_history_mean_wb), after every read of the table by that pass.  Without it
// (multi-GPU exchange) the write-back is a launch of its own on the chain:
//   chain : full_mean(k) -> [sampled k] write-back(k) -> full_mean(k+1)
int sgcn_step_run_trains(sgcn_step* st, const int32_t* ids, int32_t ids_on_host, int32_t n, float* out_host,
                         int32_t first_train, void* stream) {
    SGCN_REQUIRE(st && n >= 0 && (n == 0 || ids) && first_train >= 0, "step_run_trains: bad argument");
    if (n == 0) return SGCN_OK;
    if (st->train <= 0) {
        set_error("step_run_trains: the sampler cannot sample trains of batches (importance sampling, batch or "
                  "degree beyond the fused sampler's bounds); use sgcn_step_run");
        return SGCN_ESTATE;
    }
    const sgcn_step_desc& d = st->d;
    SGCN_REQUIRE(d.x0_alt[0] && d.x0_alt[1] && d.dx_alt, "step_run_trains: desc.x0_alt / dx_alt are required");
    sgcn_sampler* smp = st->sampler;
    const int B = d.batch, H = d.hidden, R = sgcn_step::kRing2, T = st->train;
    const bool cv = d.mode != 0, cvd = d.mode == 2, concat = d.concat != 0, multi = d.world > 1 && cv;
    // Two device-side waits are for work the host submits LATER: the fused write-back waits for the pass's sampled
    // aggregate, and a train's sampler waits for the consumer marks of the previous train's passes when they share
    // a node.  With CUDA's lazy module loading the FIRST launch of a kernel may have to wait for the device to go
    // idle -- behind a thread block that is waiting for that very kernel (seen as a 4 s bounded spin and a
    // mis-ordered row permutation in a cold process).  So the first run of a step object uses the unfused chain
    // and samples every train only after the previous train's passes have finished (stream-ordered, nothing to
    // wait for on the device), and thereby loads every kernel.
    const bool cold = !st->warmed;           // first run of this step object: kernels may not be loaded yet
    const bool fuse = cv && !multi && d.fuse_write_back != 0 && !cold;
    st->warmed = true;
    const bool ring = multi && d.ring > 0;
    const bool sharded = d.shard_rows > 0 && d.world > 1;
    if (sharded) SGCN_REQUIRE(!cv || ring, "step_run_trains: sharded tables need the ring form of the exchange");
    // sharded tables: history rows / feature rows of other ranks are read over NVLink through the shard maps
    struct ShardScope {
        bool on;
        ShardScope(bool enable, const sgcn_step_desc& d) : on(enable) {
            if (on) {
                sgcn_shard_set(0, d.world, d.shard_rows, d.hist_shards);
                sgcn_shard_set(1, d.world, d.shard_rows, d.feat_shards);
            }
        }
        ~ShardScope() {
            if (on) {
                sgcn_shard_set(0, 0, 0, nullptr);
                sgcn_shard_set(1, 0, 0, nullptr);
            }
        }
    } shard_scope(sharded, d);
    // ... and the handshakes of the sharded write-back rely on plain stream order: no programmatic launches
    PdlOff sharded_plain(sharded);
    const int width = H * (concat ? 2 : 1);
    cudaStream_t user = (cudaStream_t)stream, chain = st->chain, side = st->side, samp = st->samp, pre = st->pre,
                 copy = st->copy;
    float* x0b[3] = {d.x0, d.x0_alt[0], d.x0_alt[1]};
    float* dxb[2] = {d.dx, d.dx_alt};
    auto nb = [&](float* base) { return base + (concat ? H : 0); };

    // train c covers passes [tb(c), tb(c + 1)); its batches live in sampler sets (c & 1) * T + position
    const int T0 = first_train > 0 ? std::min<int>(first_train, T) : T;
    auto tb = [&](int c) { return c <= 0 ? 0 : std::min(n, T0 + (c - 1) * T); };
    const int n_trains = n <= T0 ? 1 : 1 + (n - T0 + T - 1) / T;
    auto train_of = [&](int k) { return k < T0 ? 0 : 1 + (k - T0) / T; };
    auto set_of = [&](int k) { const int c = train_of(k); return (c & 1) * T + (k - tb(c)); };
    auto ids_of = [&](int c) -> const int32_t* {         // where the kernels find the ids of train c
        return ids_on_host ? st->ids_stage[c & 1] : ids + (int64_t)tb(c) * B;
    };

    SGCN_CUDA(cudaEventRecord(st->ev_begin, user));
    for (cudaStream_t s : {chain, side, samp, pre}) SGCN_CUDA(cudaStreamWaitEvent(s, st->ev_begin, 0));
    if (out_host) SGCN_CUDA(cudaStreamWaitEvent(copy, st->ev_begin, 0));
    if (fuse) {       // counters of the fused write-back: zero before the first sampled aggregate / mean of the run
        STEP_TRY(sgcn_wb_counters_reset(st->flags, chain));
        SGCN_CUDA(cudaEventRecord(st->ev_zero0, chain));
        SGCN_CUDA(cudaStreamWaitEvent(side, st->ev_zero0, 0));
    }

    auto issue_train = [&](int c) -> int {
        const int len = tb(c + 1) - tb(c);
        // its buffer sets (and its id staging) were those of train c-2: every pass of that train has finished
        if (c >= 2) SGCN_CUDA(cudaStreamWaitEvent(samp, st->t_rest[(tb(c - 1) - 1) % R], 0));
        // cold run: ... and so has every pass of train c-1 (issued by now, see the pass loop)
        if (cold && c >= 1) SGCN_CUDA(cudaStreamWaitEvent(samp, st->t_rest[(tb(c) - 1) % R], 0));
        if (ids_on_host)
            SGCN_CUDA(cudaMemcpyAsync(st->ids_stage[c & 1], ids + (int64_t)tb(c) * B, sizeof(int32_t) * (size_t)len * B,
                                      cudaMemcpyHostToDevice, samp));
        const bool prev = c >= 1 && cv;      // only the full-neighbour mean reads adjacency rows in place
        STEP_TRY(sgcn_sampler_expand_train(smp, ids_of(c), len, (c & 1) * T, prev ? ids_of(c - 1) : nullptr,
                                           prev ? (tb(c) - tb(c - 1)) * B : 0, samp));
        SGCN_CUDA(cudaEventRecord(st->t_train[c % 4], samp));
        return SGCN_OK;
    };
    // gather + dX init + output zeroing of pass k, into copies k % 3 (x0) and k & 1 (dx, out)
    auto ahead = [&](int k) -> int {
        const sgcn_step::Lv& v = st->tlv[(size_t)set_of(k)];
        const int r = k & 1;
        SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_train[train_of(k) % 4], 0));
        if (k >= 2) {     // out / dx copy r: pass k-2 wrote them; its rows must also have left for the host
            SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_full[(k - 2) % R], 0));
            SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_fwd[(k - 2) % R], 0));
            if (out_host) SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_d2h[(k - 2) % R], 0));
        }
        // x0 copy k % 3: read by sampled(k-3) and write-back(k-3)
        if (k >= 3) SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_rest[(k - 3) % R], 0));
        if (k >= 1 && st->gather_after_sampled) SGCN_CUDA(cudaStreamWaitEvent(pre, st->t_fwd[(k - 1) % R], 0));
        STEP_TRY(sgcn_gather_pad_pair(d.features, d.ld_feat, v.field, d.x0_rows, v.meta + 1, d.feat_dim, x0b[k % 3], d.ld_x0,
                                      concat ? d.d_out : nullptr, d.ld_dout, concat ? B : 0,
                                      concat ? v.meta + 0 : nullptr, d.x0_rows, H, dxb[r], d.ld_dx,
                                      nullptr, 0, 0, nullptr, cv ? B : 0, H, cv ? nb(d.out[r]) : nullptr, d.ld_out, pre));
        if (cvd) STEP_TRY(sgcn_copy_rows_pad(nullptr, 0, 0, nullptr, B, H, nb(d.out_mu[r]), d.ld_out, pre));
        SGCN_CUDA(cudaEventRecord(st->t_pre[k % R], pre));
        // multi-GPU, ring form: the rows this pass will write back exist now -- publish them to every rank a
        // whole pass before anybody applies them (their flags are long up when the claim pass looks)
        if (ring)
            STEP_TRY(sgcn_wb_push_ring(v.field, v.meta + 1, d.wb_bound, x0b[k % 3] + (cvd ? H : 0), d.ld_x0, H, d.ring_dst,
                                       d.world, d.ring, d.ring_stride, d.ring_peer_flags, d.rank, d.push_epoch,
                                       d.block_counter, pre));
        return SGCN_OK;
    };

    STEP_TRY(sgcn_sampler_set_stream_async(smp, samp));
    for (int slot = 0; slot < 3; ++slot) {           // forget the batches of the three-set drivers' earlier runs
        STEP_TRY(sgcn_sampler_set_slot(smp, slot));
        STEP_TRY(sgcn_sampler_start_batch_device(smp, 0, nullptr));
    }
    STEP_TRY(issue_train(0));
    if (n_trains > 1 && !cold) STEP_TRY(issue_train(1));
    STEP_TRY(ahead(0));

    for (int k = 0; k < n; ++k) {
        const int r = k & 1, c = train_of(k);
        if (cold && k >= 1 && k == tb(c)) {       // cold run: train c is sampled here, behind the passes of train c-1
            STEP_TRY(issue_train(c));
            STEP_TRY(ahead(k));
        }
        const sgcn_step::Lv& v = st->tlv[(size_t)set_of(k)];
        const int32_t* n_out_dev = v.meta + 0;
        const int32_t* n_in_dev = v.meta + 1;
        float* out_r = d.out[r];
        float* outmu_r = d.out_mu[r];
        const float* x = x0b[k % 3];
        const float* mu = x + H;
        const float* new_hist = cvd ? mu : x;
        const float* d_nb = d.d_out + (concat ? H : 0);

        // ---- chain: the full-neighbour history mean of pass k ----
        SGCN_CUDA(cudaStreamWaitEvent(chain, st->t_train[c % 4], 0));
        SGCN_CUDA(cudaStreamWaitEvent(chain, st->t_pre[k % R], 0));
        if (cv) {
            if (fuse) {
                STEP_TRY(sgcn_full_history_mean_wb(v.field, v.rowptr_f, B, n_out_dev, st->adj_p, st->adj_i, st->adj_w,
                                                   d.history, d.ld_hist, H, cvd ? nb(outmu_r) : nb(out_r), d.ld_out,
                                                   cvd ? nb(out_r) : nullptr, d.ld_out, v.field, n_in_dev, d.x0_rows,
                                                   new_hist, d.ld_x0, st->flags, st->pipe + 1, chain));
                SGCN_CUDA(cudaEventRecord(st->t_rest[k % R], chain));
            } else {
                STEP_TRY(sgcn_full_history_mean(v.field, v.rowptr_f, B, n_out_dev, st->adj_p, st->adj_i, st->adj_w,
                                                d.history, d.ld_hist, H, cvd ? nb(outmu_r) : nb(out_r), d.ld_out,
                                                cvd ? nb(out_r) : nullptr, d.ld_out, nullptr, chain));
            }
        }
        SGCN_CUDA(cudaEventRecord(st->t_full[k % R], chain));
        // ---- pre: everything of pass k+1 that does not read the history ----
        if (k + 1 < n && !st->gather_after_sampled && !(cold && k + 1 == tb(c + 1))) STEP_TRY(ahead(k + 1));
        // ---- side: the sampled aggregate + backward of pass k (history as of write-back k-1) ----
        // plain launches: a programmatic launch here would sit resident in griddepcontrol.wait for most of a
        // full-neighbour mean and keep the NEXT mean's thread blocks off those SMs (profiles/r02_timeline_*)
        {
        PdlOff side_plain;
        SGCN_CUDA(cudaStreamWaitEvent(side, st->t_pre[k % R], 0));
        if (k >= 1) SGCN_CUDA(cudaStreamWaitEvent(side, st->t_rest[(k - 1) % R], 0));     // history as of write-back k-1
        if (fuse) STEP_TRY(sgcn_sampled_done_attach(st->flags));
        const bool split = ring && !sharded && st->split_apply;
        if (split) {
            // multi-GPU, ring form, opt-in (SGCN_WB_SPLIT=1): the claims of this pass's exchange epoch (who refreshes
            // which row: the highest rank wins) need the peers' rows and the previous epoch's copy (awaited just
            // above), not this pass's reads of the table, so they can leave the chain.  Measured on 2 GPUs: 39.8 us
            // per pass against 29.3 with claim -> copy on the chain -- the claim has to wait for the peers' pushes of
            // THIS pass, which land ~10 us into the mean, and every event hop between streams costs 4-8 us of launch
            // latency under load (profiles/r02_timeline_2gpu_split.txt).
            STEP_TRY(sgcn_wb_claim_ring(d.ring_recv, d.slot_bytes, d.world, d.wb_bound, d.owner, d.ring_flags, d.ring,
                                        d.ring_stride, d.apply_epoch, d.apply_stash, d.timeout_flag, side));
            SGCN_CUDA(cudaEventRecord(st->t_claim[k % R], side));
        }
        if (d.mode == 0) {
            STEP_TRY(sgcn_spmm_csr(v.rowptr_s, v.edg_t, v.edg_w, nullptr, B, n_out_dev, x, d.ld_x0, H, nb(out_r),
                                   d.ld_out, 0, side));
            if (concat) STEP_TRY(sgcn_copy_rows_pad(x, d.ld_x0, B, n_out_dev, B, H, out_r, d.ld_out, side));
            STEP_TRY(sgcn_spmm_csr_bwd(v.rowptr_s, v.edg_t, v.edg_w, nullptr, B, n_out_dev, d_nb, d.ld_dout, H, dxb[r],
                                       d.ld_dx, side));
        } else if (!cvd) {
            if (multi && !ring)      // two-area form: the push rides on the sampled launch (sgcn_wb_push_attach)
                STEP_TRY(sgcn_wb_push_attach(v.field, n_in_dev, d.wb_bound, new_hist, d.ld_x0, H, d.dst_even, d.dst_odd,
                                             d.world, d.peer_flags, d.rank, d.epoch, d.block_counter));
            STEP_TRY(sgcn_cv_sampled_fwd_bwd(v.rowptr_s, v.edg_t, v.edg_w, v.tgt, B, n_out_dev, x, d.ld_x0, d.history,
                                             d.ld_hist, H, nb(out_r), d.ld_out, concat ? out_r : nullptr, d.ld_out, 1,
                                             d_nb, d.ld_dout, dxb[r], d.ld_dx, side));
        } else {
            if (multi && !ring)
                STEP_TRY(sgcn_wb_push_attach(v.field, n_in_dev, d.wb_bound, new_hist, d.ld_x0, H, d.dst_even, d.dst_odd,
                                             d.world, d.peer_flags, d.rank, d.epoch, d.block_counter));
            STEP_TRY(sgcn_cvd_sampled_fwd_bwd(v.rowptr_s, v.edg_t, v.edg_w, v.tgt, v.scales, B, n_out_dev, x, d.ld_x0,
                                              mu, d.ld_x0, d.history, d.ld_hist, H, nb(out_r), d.ld_out,
                                              nb(outmu_r), d.ld_out, concat ? out_r : nullptr, d.ld_out,
                                              concat ? outmu_r : nullptr, d.ld_out, 1, d_nb, d.ld_dout, dxb[r],
                                              d.ld_dx, side));
        }
        SGCN_CUDA(cudaEventRecord(st->t_fwd[k % R], side));
        }
        if (k + 1 < n && st->gather_after_sampled && !(cold && k + 1 == tb(c + 1))) STEP_TRY(ahead(k + 1));
        // ---- write-back after every forward read of history (gcn/models.py:186-194) ----
        if (!fuse) {                     // on the chain (programmatic launch: it is the chain's next link)
            SGCN_CUDA(cudaStreamWaitEvent(chain, st->t_fwd[k % R], 0));
            if (!cv) {
                STEP_TRY(sgcn_sampler_mark_consumed(smp, chain));
            } else if (ring && sharded) {
                STEP_TRY(sgcn_wb_wait_apply_sharded(d.history, d.ld_hist, H, d.ring_recv, d.slot_bytes, d.world, d.rank,
                                                    d.shard_rows, d.wb_bound, d.owner, d.ring_flags, d.ring,
                                                    d.ring_stride, d.apply_epoch, d.apply_stash, d.reads_flags,
                                                    d.reads_peer_flags, d.applied_flags, d.applied_peer_flags,
                                                    d.shard_counter, d.timeout_flag, st->pipe + 1, chain));
            } else if (ring && st->split_apply) {
                // ... and the chain keeps the copy alone: ids, claims and rows in registers before the mean has
                // finished (programmatic launch), the stores behind it
                SGCN_CUDA(cudaStreamWaitEvent(chain, st->t_claim[k % R], 0));
                STEP_TRY(sgcn_wb_copy_ring(d.history, d.ld_hist, H, d.ring_recv, d.slot_bytes, d.world, d.wb_bound, d.owner,
                                           d.ring, d.ring_stride, d.apply_epoch, d.apply_stash, st->pipe + 1, chain));
            } else if (ring) {
                STEP_TRY(sgcn_wb_wait_apply_ring(d.history, d.ld_hist, H, d.ring_recv, d.slot_bytes, d.world, d.wb_bound,
                                                 d.owner, d.ring_flags, d.ring, d.ring_stride, d.apply_epoch,
                                                 d.apply_stash, d.timeout_flag, st->pipe + 1, chain));
            } else if (multi) {
                STEP_TRY(sgcn_wb_wait_apply(d.history, d.ld_hist, H, d.recv_even, d.recv_odd, d.slot_bytes, d.world,
                                            d.wb_bound, d.owner, d.flags, d.epoch, d.timeout_flag, st->pipe + 1, chain));
            } else {
                STEP_TRY(sgcn_history_update(d.history, d.ld_hist, v.field, d.x0_rows, n_in_dev, new_hist, d.ld_x0, H,
                                             st->pipe + 1, chain));
            }
            SGCN_CUDA(cudaEventRecord(st->t_rest[k % R], chain));
        }
        // ---- copy: the pass's aggregated rows to pinned host memory (both aggregate kernels done) ----
        if (out_host) {
            SGCN_CUDA(cudaStreamWaitEvent(copy, st->t_full[k % R], 0));
            SGCN_CUDA(cudaStreamWaitEvent(copy, st->t_fwd[k % R], 0));
            SGCN_CUDA(cudaMemcpy2DAsync(out_host + (int64_t)k * B * width, sizeof(float) * (size_t)width, out_r,
                                        sizeof(float) * (size_t)d.ld_out, sizeof(float) * (size_t)width, (size_t)B,
                                        cudaMemcpyDeviceToHost, copy));
            SGCN_CUDA(cudaEventRecord(st->t_d2h[k % R], copy));
        }
        // ---- samp: the train after next, once the last pass of this train has been issued ----
        if (k + 1 == tb(c + 1) && c + 2 < n_trains && !cold) STEP_TRY(issue_train(c + 2));
    }
    SGCN_CUDA(cudaEventRecord(st->ev_side_end, side));
    SGCN_CUDA(cudaEventRecord(st->ev_samp_end, samp));
    SGCN_CUDA(cudaEventRecord(st->ev_pre_end, pre));
    SGCN_CUDA(cudaEventRecord(st->ev_zero0, chain));
    if (out_host) SGCN_CUDA(cudaStreamWaitEvent(user, st->t_d2h[(n - 1) % R], 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->t_rest[(n - 1) % R], 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_zero0, 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_side_end, 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_samp_end, 0));
    SGCN_CUDA(cudaStreamWaitEvent(user, st->ev_pre_end, 0));
    return SGCN_OK;
}

}  // extern "C"
