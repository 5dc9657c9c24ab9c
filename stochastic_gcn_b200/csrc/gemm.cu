// First dense layer of the pre-processed models with the feature-row gather fused into its A-operand load
// (gcn/layers.py:100-138 `Dense`: act(MyLayerNorm(x @ W)), x = rows of the PP feature matrix
// picked by the input field, gcn/train.py:190 + gcn/history.cpp:50-88):
//
//     out[i, :] = act(LN(features[idx[i], :K] @ W))        W: [K, 128]
//
// on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 issued by one thread per thread block, operands in
// shared memory (128-byte swizzled, K-major), the 128 x 128 fp32 accumulator in tensor memory (TMEM), read back
// with tcgen05.ld for the epilogue (row moments, normalisation, activation: one thread = one row).
//
// fp32 parity (1e-4, north_star) rules out a single TF32 pass (10-bit mantissa: ~1e-3).  Both operands are split
// x = hi + lo with hi = tf32(x), lo = tf32(x - hi) and three products accumulate into the same TMEM tile:
//     hi.hi + lo.hi + hi.lo        (the dropped lo.lo term is ~2^-22 relative)
// which reproduces the fp32 product to ~1e-6 (tests/test_gemm_gpu.py, against float64).
//
// Data movement.  The gather needs a transformation on the way (the split), so the A tile goes global ->
// registers (coalesced 128-byte row segments, three K-chunks prefetched ahead) -> hi / lo -> shared memory in the
// swizzled layout the matrix descriptor names.  W is split and laid out ONCE (sgcn_gemm_pack_w) as the exact
// shared-memory image of every K-chunk, so a stage's B operand (hi and lo tile, 32 KB) is one bulk copy by the
// TMA engine (cp.async.bulk, completion counted on an mbarrier).
//
// Roles (no block-wide barrier inside the K loop): eight producer warps stage A (3 stages; each warp arrives on
// the stage's "A full" mbarrier after its stores and the proxy fence); one thread of a ninth warp keeps the bulk
// copies of W three chunks ahead (4 stages), waits for "A full" / "B full", issues the 12 MMAs of the chunk and
// hands both stages back with tcgen05.commit ("A empty" / "B empty" arrive when those MMAs have retired).
// The first form of this kernel (all threads staging, one __syncthreads and the W copy's full latency per
// chunk) took 59 us under ncu for the headline shape; profiles/r02_gather_gemm_ncu.json.
#include "common.cuh"

namespace sgcn {

constexpr int kGemmM = 128;            // rows per thread block = TMEM lanes
constexpr int kGemmN = 128;            // output width = accumulator columns (fp32)
constexpr int kGemmKC = 32;            // K per stage: 32 tf32 = one 128-byte swizzle row
constexpr int kGemmStagesA = 3;        // A stages (hi | lo tile), filled by the producer warps
constexpr int kGemmStagesB = 4;        // B stages (hi | lo tile), filled by the TMA engine three chunks ahead
constexpr int kGemmProducers = 256;    // 8 producer warps stage the A rows ...
constexpr int kGemmThreads = kGemmProducers + 32;        // ... one more warp issues the bulk copies and the MMAs
constexpr int kGemmPrefetch = 3;       // K-chunks of A rows held in registers ahead of the one being staged
constexpr int kTileBytes = kGemmM * 128;                 // 16 KB: 128 rows x 128 bytes
constexpr int kPairBytes = 2 * kTileBytes;               // a hi | lo pair
constexpr int kGemmSmem = (kGemmStagesA + kGemmStagesB) * kPairBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ +
                          kGemmM * 4;

// byte offset of element (row r, k in [0, 32)) inside a 128-row K-major tile with the 128-byte swizzle:
// 8-row groups of 1024 bytes, 16-byte chunk index XOR row-in-group (Swizzle<3,4,3>)
__host__ __device__ inline int tile_offset(int r, int k) {
    const int chunk = k >> 2, within = k & 3;
    return (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4) + within * 4;
}

__device__ __forceinline__ float tf32_round(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier / TMA / tcgen05 wrappers (PTX ISA 8.6+, sm_100a) --------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// bounded: a barrier that never completes (a bug, a bad pointer) traps instead of hanging the GPU
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T : M = 128, N = 128, K = 8 (tf32), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc) : "memory");
}
// the mbarrier gets one arrival once every MMA issued so far by this thread has retired
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major tile, 128-byte swizzle.
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused under a swizzle: 1) |
//   [32,46) stride byte offset >> 4 (8-row group pitch: 1024 B) | [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_byte_addr) {
    return (uint64_t)((smem_byte_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = fp32 [4,6) = 1, A = B = tf32 [7,10) = [10,13) = 2,
// both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kGemmN >> 3) << 17) |
                                ((uint32_t)(kGemmM >> 4) << 24);

struct GemmArgs {
    const float* src; int64_t ld_src; const int32_t* idx; int n; const int32_t* n_dev; int K;
    const float* w_packed;             // sgcn_gemm_pack_w: per K-chunk {hi tile, lo tile}, 32 KB each chunk
    float* out; int64_t ld_out;        // act(LN(.)) or the raw product (epilogue = 0)
    float* pre; int64_t ld_pre;        // optional: the raw product x @ W (the backward's input)
    float* stats;                      // optional: {mean, rstd} per row (as sgcn_ln_act_fwd)
    int epilogue;                      // 0: none, 1: layer norm + relu, 2: layer norm
    float eps;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gather_gemm_tf32x3_kernel(const GemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment: the swizzle pattern is a function of the address bits
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;                                             // [3] x {hi tile, lo tile}
    uint8_t* smem_b = smem + kGemmStagesA * kPairBytes;                 // [4] x {hi tile, lo tile}
    uint64_t* bars = (uint64_t*)(smem + (kGemmStagesA + kGemmStagesB) * kPairBytes);
    uint64_t* a_full = bars;                    // [3]  8 arrivals: the producer warps
    uint64_t* a_empty = bars + 3;               // [3]  tcgen05.commit
    uint64_t* b_full = bars + 6;                // [4]  bulk copy (transaction bytes)
    uint64_t* b_empty = bars + 10;              // [4]  tcgen05.commit
    uint64_t* done = bars + 14;                 //      the accumulator is complete
    uint32_t* s_tmem = (uint32_t*)(bars + 16);
    int32_t* s_row = (int32_t*)(s_tmem + 4);                             // source row of each tile row, -1 = none
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = dev_count(a.n_dev, a.n);
    const int m0 = blockIdx.x * kGemmM;
    if (m0 >= n) return;
    const int nk = (a.K + kGemmKC - 1) / kGemmKC;

    if (tid == 0) {
        for (int i = 0; i < 3; ++i) { mbar_init(smem_addr(a_full + i), kGemmProducers / 32); mbar_init(smem_addr(a_empty + i), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(smem_addr(b_full + i), 1); mbar_init(smem_addr(b_empty + i), 1); }
        mbar_init(smem_addr(done), 1);
        fence_barrier_init();
    }
    if (warp == kGemmProducers / 32) tmem_alloc(smem_addr(s_tmem), kGemmN);   // 128 columns x 128 lanes x fp32
    for (int r = tid; r < kGemmM; r += kGemmThreads)
        s_row[r] = m0 + r < n ? (a.idx ? a.idx[m0 + r] : m0 + r) : -1;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < kGemmProducers / 32) {
        // ===== producers: A rows global -> registers -> tf32 hi / lo -> swizzled shared memory =====
        // the tile is 128 rows x 8 chunks of 16 bytes; thread t covers rows t/8 + 32 i (i < 4), chunk t%8 --
        // 8 consecutive threads read one row's contiguous 128 bytes
        const int chunk = tid & 7;
        int64_t row_off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int src_row = s_row[(tid >> 3) + 32 * i];
            row_off[i] = src_row >= 0 ? (int64_t)src_row * a.ld_src : -1;
        }
        auto load_chunk = [&](int kc, float4 (&dst)[4]) {
            const int k = kc * kGemmKC + chunk * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = (kc < nk && row_off[i] >= 0 && k < a.K) ? ldg_stream4(a.src + row_off[i] + k)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // one K-chunk: `cur` holds its rows (loaded three chunks ago), `far` receives the rows of chunk kc + 3
        auto step = [&](int kc, float4 (&cur)[4], float4 (&far)[4]) {
            const int s = kc % kGemmStagesA, use = kc / kGemmStagesA;
            uint8_t* stage = smem_a + s * kPairBytes;
            if (kc >= kGemmStagesA) mbar_wait(smem_addr(a_empty + s), (use - 1) & 1);   // its last readers have retired
            load_chunk(kc + kGemmPrefetch, far);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = (tid >> 3) + 32 * i;
                float4 hi, lo;
                hi.x = tf32_round(cur[i].x); lo.x = tf32_round(cur[i].x - hi.x);
                hi.y = tf32_round(cur[i].y); lo.y = tf32_round(cur[i].y - hi.y);
                hi.z = tf32_round(cur[i].z); lo.z = tf32_round(cur[i].z - hi.z);
                hi.w = tf32_round(cur[i].w); lo.w = tf32_round(cur[i].w - hi.w);
                const int off = tile_offset(r, chunk * 4);
                *(float4*)(stage + off) = hi;
                *(float4*)(stage + kTileBytes + off) = lo;
            }
            fence_proxy_async_smem();                     // generic-proxy stores -> visible to the tensor core's reads
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_addr(a_full + s));
        };
        static_assert(kGemmPrefetch == 3, "the register rotation below is written for three chunks of prefetch");
        float4 r0[4], r1[4], r2[4], r3[4];                // rotating register sets (static indices: no local memory)
        load_chunk(0, r0);
        load_chunk(1, r1);
        load_chunk(2, r2);
        for (int kc = 0; kc < nk; kc += 4) {              // chunk kc lives in set kc % 4, chunk kc + 3 goes to (kc + 3) % 4
            step(kc, r0, r3);
            if (kc + 1 < nk) step(kc + 1, r1, r0);
            if (kc + 2 < nk) step(kc + 2, r2, r1);
            if (kc + 3 < nk) step(kc + 3, r3, r2);
        }
    } else if (lane == 0) {
        // ===== one thread: W tiles by bulk copy (three chunks ahead), MMA issue, stage hand-back =====
        auto copy_b = [&](int kc) {
            const int sb = kc % kGemmStagesB;
            mbar_expect_tx(smem_addr(b_full + sb), kPairBytes);
            bulk_g2s(smem_addr(smem_b + sb * kPairBytes), a.w_packed + (size_t)kc * (kPairBytes / 4), kPairBytes,
                     smem_addr(b_full + sb));
        };
        for (int kc = 0; kc < kGemmStagesB - 1 && kc < nk; ++kc) copy_b(kc);
        for (int kc = 0; kc < nk; ++kc) {
            const int sa = kc % kGemmStagesA, sb = kc % kGemmStagesB;
            mbar_wait(smem_addr(a_full + sa), (kc / kGemmStagesA) & 1);
            mbar_wait(smem_addr(b_full + sb), (kc / kGemmStagesB) & 1);
            tc_fence_after();
            const uint32_t a_hi = smem_addr(smem_a + sa * kPairBytes), a_lo = a_hi + kTileBytes;
            const uint32_t b_hi = smem_addr(smem_b + sb * kPairBytes), b_lo = b_hi + kTileBytes;
#pragma unroll
            for (int ks = 0; ks < kGemmKC / 8; ++ks) {    // one MMA = 8 tf32 of K = 32 bytes along the swizzle row
                const uint32_t o = ks * 32;
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_hi + o), kIdescTf32, (kc | ks) != 0);
                umma_tf32(tmem, umma_desc(a_lo + o), umma_desc(b_hi + o), kIdescTf32, 1);
                umma_tf32(tmem, umma_desc(a_hi + o), umma_desc(b_lo + o), kIdescTf32, 1);
            }
            umma_commit(smem_addr(a_empty + sa));         // both stages free once these MMAs have retired
            umma_commit(smem_addr(b_empty + sb));
            if (kc == nk - 1) umma_commit(smem_addr(done));   // ... and the accumulator complete
            // behind the issue: the W tiles of chunk kc + 3 go where chunk kc - 1 was read from
            if (kc + kGemmStagesB - 1 < nk) {
                if (kc >= 1) mbar_wait(smem_addr(b_empty + (kc - 1) % kGemmStagesB), ((kc - 1) / kGemmStagesB) & 1);
                copy_b(kc + kGemmStagesB - 1);
            }
        }
    }
    if (warp == kGemmProducers / 32) __syncwarp();      // its other lanes park here instead of polling beside the issuer

    // ---- epilogue: TMEM -> registers, one thread = one row (warps 0..3 own TMEM lanes 32 w .. 32 w + 31) ----
    mbar_wait(smem_addr(done), 0);
    tc_fence_after();
    if (warp < 4) {
        const int r = tid;                                 // tile row = TMEM lane
        const int row = m0 + r;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        float mean = 0.f, rstd = 1.f;
        if (a.epilogue != 0) {
            float sum = 0.f;
            for (int c = 0; c < kGemmN; c += 16) {
                float v[16];
                tmem_ld16(lane_addr + c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) sum += v[i];
            }
            mean = sum / (float)kGemmN;
            float q = 0.f;
            for (int c = 0; c < kGemmN; c += 16) {
                float v[16];
                tmem_ld16(lane_addr + c, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) q += (v[i] - mean) * (v[i] - mean);
            }
            rstd = rsqrtf(q / (float)kGemmN + a.eps);
            if (a.stats && row < n) *(float2*)(a.stats + 2 * (int64_t)row) = make_float2(mean, rstd);
        }
        for (int c = 0; c < kGemmN; c += 16) {
            float v[16];
            tmem_ld16(lane_addr + c, v);                   // (all 32 lanes execute the load: .sync.aligned)
            if (row < n) {
                if (a.pre) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *(float4*)(a.pre + (int64_t)row * a.ld_pre + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                if (a.epilogue != 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] = (v[i] - mean) * rstd;
                        if (a.epilogue == 1) v[i] = fmaxf(v[i], 0.f);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *(float4*)(a.out + (int64_t)row * a.ld_out + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kGemmProducers / 32) tmem_dealloc(tmem, kGemmN);
}

// W [K, 128] row-major (x @ W) -> per K-chunk of 32 the two B tiles (n-major rows of 32 k: B[n][k] = W[k][n]),
// split into tf32 hi / lo and laid out exactly as the kernel's shared-memory stage expects them
__global__ void gemm_pack_w_kernel(const float* __restrict__ w, int64_t ld_w, int K, float* __restrict__ packed) {
    const int kc = blockIdx.x;
    uint8_t* base = (uint8_t*)packed + (size_t)kc * kPairBytes;
    for (int e = threadIdx.x; e < kGemmN * kGemmKC; e += blockDim.x) {
        const int kk = e / kGemmN, nn = e % kGemmN;        // consecutive threads: consecutive n (coalesced reads of W)
        const int k = kc * kGemmKC + kk;
        const float x = k < K ? w[(int64_t)k * ld_w + nn] : 0.f;
        const float hi = tf32_round(x), lo = tf32_round(x - hi);
        const int off = tile_offset(nn, kk);
        *(float*)(base + off) = hi;
        *(float*)(base + kTileBytes + off) = lo;
    }
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

int64_t sgcn_gemm_packed_floats(int32_t K, int32_t N) {
    if (K <= 0 || N != kGemmN) return -1;
    return (int64_t)((K + kGemmKC - 1) / kGemmKC) * (kPairBytes / 4);
}

int sgcn_gemm_pack_w(const float* w, int64_t ld_w, int32_t K, int32_t N, float* packed, void* stream) {
    SGCN_REQUIRE(w && packed && K > 0, "gemm_pack_w: bad argument");
    SGCN_REQUIRE(N == kGemmN && ld_w >= N, "gemm_pack_w: the fused layer is built for 128 output columns");
    SGCN_REQUIRE((((uintptr_t)packed) & 15) == 0, "gemm_pack_w: packed must be 16-byte aligned");
    gemm_pack_w_kernel<<<(K + kGemmKC - 1) / kGemmKC, 256, 0, (cudaStream_t)stream>>>(w, ld_w, K, packed);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_gather_gemm_tf32x3(const float* src, int64_t ld_src, const int32_t* idx, int32_t n, const int32_t* n_dev,
                            int32_t K, const float* w_packed, int32_t N, float* out, int64_t ld_out, float* pre,
                            int64_t ld_pre, float* stats, int32_t epilogue, float eps, void* stream) {
    SGCN_REQUIRE(n >= 0 && K > 0, "gather_gemm: bad size");
    if (n == 0) return SGCN_OK;
    SGCN_REQUIRE(src && w_packed && out, "gather_gemm: null pointer");
    SGCN_REQUIRE(N == kGemmN, "gather_gemm: the fused layer is built for 128 output columns");
    SGCN_REQUIRE(epilogue >= 0 && epilogue <= 2, "gather_gemm: epilogue is 0 (none), 1 (layer norm + relu) or 2 (layer norm)");
    SGCN_REQUIRE(K % 4 == 0 && ld_src % 4 == 0 && ld_src >= K && (((uintptr_t)src) & 15) == 0,
                 "gather_gemm: source rows must be 16-byte aligned multiples of 4 floats");
    SGCN_REQUIRE(ld_out >= N && ld_out % 4 == 0 && (((uintptr_t)out) & 15) == 0 &&
                     (!pre || (ld_pre >= N && ld_pre % 4 == 0 && (((uintptr_t)pre) & 15) == 0)),
                 "gather_gemm: output rows must be 16-byte aligned");
    SGCN_REQUIRE((((uintptr_t)w_packed) & 15) == 0 && (!stats || (((uintptr_t)stats) & 7) == 0),
                 "gather_gemm: packed weights / stats misaligned");
    static bool attr_set = false;
    if (!attr_set) {
        SGCN_CUDA(cudaFuncSetAttribute(gather_gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
        attr_set = true;
    }
    GemmArgs a{src, ld_src, idx, n, n_dev, K, w_packed, out, ld_out, pre, ld_pre, stats, epilogue, eps};
    gather_gemm_tf32x3_kernel<<<(n + kGemmM - 1) / kGemmM, kGemmThreads, kGemmSmem, (cudaStream_t)stream>>>(a);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

}  // extern "C"
