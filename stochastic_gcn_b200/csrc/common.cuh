// Shared helpers for libsgcn_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "sgcn_b200.h"

namespace sgcn {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

void set_error(const std::string& msg);
extern std::atomic<int64_t> g_launches;
// optional device timeline: trace[2*id] = earliest block start, trace[2*id+1] = latest block end
// (globaltimer ns) of kernel class `id`; NULL = tracing off (sgcn_trace_set)
extern unsigned long long* g_trace;
enum { TR_SAMPLER = 0, TR_FULL = 1, TR_GATHER = 2, TR_SAMPLED = 3, TR_BWD = 4, TR_UPDATE = 5, TR_PAD = 6,
       TR_EXCHANGE = 7, TR_CLASSES = 8,
       // event-log only (no min / max slots): the three kernels of the multi-GPU write-back exchange
       TR_WB_PUSH = 8, TR_WB_CLAIM = 9, TR_WB_COPY = 10,
       // single stamps of the fused write-back: the LAST block of a full-neighbour mean has finished its
       // positions / the tail has stored the rows; the last block of a sampled aggregate has finished
       TR_FULL_BODY_END = 11, TR_FULL_TAIL_END = 12, TR_SAMPLED_END = 13 };

inline int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    set_error(buf);
    return SGCN_ECUDA;
}

#define SGCN_CUDA(call)                                                           \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return ::sgcn::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Counts the launch and checks the launch status (cheap: no sync).
#define SGCN_LAUNCHED()                                                           \
    do {                                                                          \
        ::sgcn::g_launches.fetch_add(1, std::memory_order_relaxed);               \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) return ::sgcn::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define SGCN_REQUIRE(cond, msg)                                                   \
    do {                                                                          \
        if (!(cond)) { ::sgcn::set_error(std::string("invalid argument: ") + (msg)); return SGCN_EINVAL; } \
    } while (0)

// (PDL) when set, the kernels of the step's main chain (full-neighbour mean -> history write-back ->
// next full-neighbour mean) are launched with programmatic stream serialization: each starts while its
// stream predecessor is still running, does the part of its work that does not depend on it, and
// orders the rest with griddepcontrol.wait (sgcn_tune_set SGCN_TUNE_PDL)
extern int g_pdl;
// per host thread: > 0 while a driver wants plain stream-ordered launches from the PDL-capable launch sites
// (a kernel launched programmatically becomes resident as soon as its predecessors let it and then sits in
// griddepcontrol.wait holding registers / shared memory -- right for the one kernel that continues the
// critical chain, wrong for side-branch kernels whose inputs are a whole full-neighbour mean away)
extern thread_local int t_pdl_off;
struct PdlOff {
    bool on;
    explicit PdlOff(bool enable = true) : on(enable) { if (on) ++t_pdl_off; }
    ~PdlOff() { if (on) --t_pdl_off; }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t dyn, cudaStream_t st,
                              Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (g_pdl && t_pdl_off == 0) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// An SM re-partitions its L1 / shared memory only when idle: a kernel whose preferred carve-out differs
// from that of the kernel occupying the SMs (the full-neighbour mean holds every SM for most of a
// step) may have to wait for an SM to drain.  For the sampler / full-mean pair that cost 13 us per step
// (DESIGN section 3); the other kernels that run beside the mean ask for the SAME 132 KB carve-out (58 %: two
// full-mean CTAs with their override tables, 2 x 27 KB, plus one 47 KB train-sampler CTA must fit together)
// as a precaution -- an A/B on the side-branch kernels (SGCN_NO_MATCH_CARVEOUT=1) showed no measurable
// difference in round 1, their late starts are dependent-launch latency under load, not re-partitioning.
constexpr int kStepCarveout = 58;
template <typename K>
inline void match_step_carveout(K kernel, bool* done) {
    if (!*done) {
        if (!getenv("SGCN_NO_MATCH_CARVEOUT"))       // A/B switch for the measurement in profiles/
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kStepCarveout);
        *done = true;
    }
}
#define SGCN_MATCH_CARVEOUT(kernel)                          \
    do {                                                     \
        static bool done__ = false;                          \
        ::sgcn::match_step_carveout(kernel, &done__);        \
    } while (0)

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Row-sharded tables (multi-GPU): row i of a [N, ld] table lives on rank i / rows_per_shard at
// base[i / rows_per_shard] + (i % rows_per_shard) * ld, where base[r] is rank r's shard mapped into this
// process (cudaIpc over NVLink; base[own rank] is local HBM).  world <= 1: the table is one local array.
struct ShardMap {
    int world, rows;
    const float* base[16];
};
// per host thread: the maps the next launches of the history-reading / feature-gathering kernels use
// (sgcn_shard_set); world = 0 = off
extern thread_local ShardMap t_hist_map, t_feat_map;
__device__ __forceinline__ const float* shard_row(const ShardMap& m, const float* table, int64_t row, int64_t ld) {
    if (m.world <= 1) return table + row * ld;
    const int o = (int)row / m.rows;
    return m.base[o] + (int64_t)((int)row - o * m.rows) * ld;
}

// ---- device-side helpers -------------------------------------------------------------------

// 128-bit read-only global load that does not allocate in L1 (rows are read once per kernel).
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// History rows: tables that kernels of the same step WRITE (write-back fused into the tail of the full-neighbour
// mean, programmatic dependent launches) -- coherent loads, never the read-only (.nc) path; no L1 allocation
// (every row is read once per kernel); L2 eviction priority from a policy word (l2_policy).
__device__ __forceinline__ float4 ld_hist4(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return r;
}
__device__ __forceinline__ float ld_hist1(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_hist4(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
// read-once streams (feature rows of the gather): same, through the read-only path
__device__ __forceinline__ float4 ldg_stream4_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
// L2 eviction policy word: kind 0 = normal, 1 = evict_last for `pct` % of the accesses (the rest unchanged),
// 2 = evict_first (sgcn_tune_set SGCN_TUNE_HIST_L2 / SGCN_TUNE_STREAM_L2)
__device__ __forceinline__ uint64_t l2_policy(int kind, int pct) {
    uint64_t pol;
    const float frac = (float)pct * 0.01f;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol) : "f"(frac));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// host-side tunables behind them: {kind, percent}
extern int g_hist_l2[2], g_stream_l2[2];
// (PDL) the write-back kernels (history_update, wb_copy) let their stream successor -- the next full-neighbour mean --
// launch at their END, not at entry: launched early it cannot fit beside the running mean anyway, and a launch that
// waits for room holds up every later launch (the next pass's gather / push) for the whole mean
// (sgcn_tune_set SGCN_TUNE_WB_TRIGGER: 1 = late (default), 0 = at entry)
extern int g_wb_late_trigger;

// 128-bit vector reduction (sm_90+): one RED for 4 floats.
__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

// acquire / release accesses at GPU scope for the device-side counters and flags kernels hand each other
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// bounded wait for a device-side counter: false + *err = 1 after ~2 s, and at once when another wait has
// already given up
__device__ __forceinline__ bool spin_until_ge(const int32_t* ctr, int want, int32_t* err) {
    long long spins = 0;
    while ((int)ld_acquire_u32((const unsigned*)ctr) < want) {
        if (*(volatile int32_t*)err != 0) return false;
        if (++spins > 8000000LL) {
            atomicExch(err, 1);
            return false;
        }
        __nanosleep(50);
    }
    return true;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// RAII-style stamps: construct at kernel entry; the destructor stamps the exit.
// trace[0..15]: per-class min block start / max block end.  trace[16]: event-log cursor;
// trace[17 + 2i], trace[18 + 2i]: (class << 1 | is_end, time) of block 0 of every launch (i < 1024).
constexpr int kTraceLogCap = 1024;
struct TraceScope {
    unsigned long long* t;
    int id;
    __device__ __forceinline__ void log(int is_end, unsigned long long now) {
        if (blockIdx.x == 0 && blockIdx.y == 0) {
            const unsigned long long i = atomicAdd(t + 16, 1ull);
            if (i < (unsigned long long)kTraceLogCap) {
                t[17 + 2 * i] = (unsigned long long)((id << 1) | is_end);
                t[18 + 2 * i] = now;
            }
        }
    }
    __device__ __forceinline__ TraceScope(unsigned long long* trace, int id_) : t(trace), id(id_) {
        if (t && threadIdx.x == 0) {
            const unsigned long long now = global_ns();
            if (id < TR_CLASSES) atomicMin(t + 2 * id, now);
            log(0, now);
        }
    }
    __device__ __forceinline__ ~TraceScope() {
        if (t && threadIdx.x == 0) {
            const unsigned long long now = global_ns();
            if (id < TR_CLASSES) atomicMax(t + 2 * id + 1, now);
            log(1, now);
        }
    }
};

// one stamp in the event log (class id, begin == end), whatever block calls it
__device__ __forceinline__ void trace_stamp(unsigned long long* t, int id) {
    if (!t) return;
    const unsigned long long now = global_ns();
    const unsigned long long i = atomicAdd(t + 16, 2ull);
    if (i + 1 < (unsigned long long)kTraceLogCap) {
        t[17 + 2 * i] = (unsigned long long)(id << 1);       t[18 + 2 * i] = now;
        t[19 + 2 * i] = (unsigned long long)((id << 1) | 1); t[20 + 2 * i] = now;
    }
}

__device__ __forceinline__ int dev_count(const int32_t* n_dev, int n_host) {
    return n_dev ? min(*n_dev, n_host) : n_host;
}

}  // namespace sgcn
