// Multi-GPU history write-back exchange (row-range sharded training).
//
// Every rank keeps a full replica of the history table in its own HBM (119 MB at Reddit shape,
// 1 GB at 2M nodes: nothing next to 180 GB) so that the dominant full-neighbour gather never
// leaves the GPU; what crosses NVLink per step is only each rank's write-back -- the <= B(1+d)
// rows it refreshed -- packed as {count, node ids, rows}.  Two transports:
//   * NCCL:  pack -> ncclAllGather (torch.distributed) -> apply
//   * peer:  pack_push writes the payload straight into every peer's receive slot through
//            NVLink-mapped pointers (cudaIpc) and raises a per-rank flag; apply spins on the
//            flags of the current step and then merges -- no collective launch, capturable in
//            the same CUDA graph as the rest of the step.
// apply is deterministic: when several ranks refreshed the same node in one step the highest
// (rank, position) wins as a WHOLE row (claim pass with atomicMax, copy pass by the winner).
#include <cstring>

#include "common.cuh"

namespace sgcn {

constexpr int kWbHeaderInts = 4;   // {count, step, 0, 0}: keeps ids 16-byte aligned

__host__ __device__ inline int64_t wb_ids_offset() { return kWbHeaderInts * 4; }
__host__ __device__ inline int64_t wb_rows_offset(int n_bound) {
    return wb_ids_offset() + (((int64_t)n_bound * 4 + 15) & ~int64_t(15));
}
__host__ __device__ inline int64_t wb_payload_bytes(int n_bound, int D) {
    return (wb_rows_offset(n_bound) + (int64_t)n_bound * D * 4 + 255) & ~int64_t(255);
}

constexpr int kMaxPeers = 16;
struct PeerPtrs { char* p[kMaxPeers]; };

// pack {count, ids, rows} into up to `n_dst` destinations (own buffer and/or peers' receive slots)
__global__ void __launch_bounds__(256)
wb_pack_kernel(const int32_t* __restrict__ field, const int32_t* __restrict__ n_dev, int n_bound,
               const float* __restrict__ rows, int64_t ld_rows, int D, PeerPtrs dst_even, PeerPtrs dst_odd,
               int n_dst, int step, int32_t* epoch, PeerPtrs flags, int my_rank,
               int32_t* block_counter) {
    // peer transport: this push belongs to epoch *epoch + 1 and lands in the slot set of its parity
    if (epoch) step = *epoch + 1;
    const PeerPtrs& dst = (epoch && (step & 1)) ? dst_odd : dst_even;
    const int n = min(*n_dev, n_bound);
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t ids_off = wb_ids_offset(), rows_off = wb_rows_offset(n_bound);
    if (t0 == 0)
        for (int k = 0; k < n_dst; ++k) {
            int32_t* h = (int32_t*)dst.p[k];
            h[0] = n; h[1] = step; h[2] = 0; h[3] = 0;
        }
    for (int64_t i = t0; i < n; i += stride) {
        const int32_t id = field[i];
        for (int k = 0; k < n_dst; ++k) ((int32_t*)(dst.p[k] + ids_off))[i] = id;
    }
    if ((D & 3) == 0 && (ld_rows & 3) == 0 && (((uintptr_t)rows) & 15) == 0) {
        const int d4 = D >> 2;
        const int64_t total = (int64_t)n * d4;
        for (int64_t t = t0; t < total; t += stride) {
            const int64_t r = t / d4;
            const int c = (int)(t - r * d4) * 4;
            const float4 v = ldg_stream4(rows + r * ld_rows + c);
            for (int k = 0; k < n_dst; ++k) *(float4*)(dst.p[k] + rows_off + (r * D + c) * 4) = v;
        }
    } else {
        const int64_t total = (int64_t)n * D;
        for (int64_t t = t0; t < total; t += stride) {
            const int64_t r = t / D;
            const int c = (int)(t - r * D);
            const float v = rows[r * ld_rows + c];
            for (int k = 0; k < n_dst; ++k) *(float*)(dst.p[k] + rows_off + (r * D + c) * 4) = v;
        }
    }
    if (block_counter) {
        // fused signal: the last block to finish advances the epoch and publishes it to every rank
        __shared__ int s_last;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(block_counter, 1) == (int)gridDim.x - 1;
        __syncthreads();
        if (s_last && threadIdx.x == 0) {
            *block_counter = 0;
            *epoch = step;
            __threadfence_system();
            for (int k = 0; k < n_dst; ++k) {
                volatile int32_t* f = (volatile int32_t*)flags.p[k];
                f[my_rank] = step;
            }
            __threadfence_system();
        }
    }
}

// after every CTA of the pack has finished: advance the local epoch and publish it as
// flag[my_rank] in every peer (and in this rank's own flag array)
__global__ void wb_signal_kernel(PeerPtrs flags, int n_peers, int my_rank, int32_t* epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int e = *epoch + 1;
        *epoch = e;
        __threadfence_system();
        for (int k = 0; k < n_peers; ++k) {
            volatile int32_t* f = (volatile int32_t*)flags.p[k];
            f[my_rank] = e;
        }
        __threadfence_system();
    }
}

// spin until flag[r] >= step for every rank (bounded: sets *timeout_flag instead of hanging)
__global__ void wb_wait_kernel(const int32_t* flags, int world, const int32_t* epoch,
                               int32_t* __restrict__ timeout_flag, long long max_spins) {
    if (blockIdx.x != 0 || threadIdx.x >= world) return;
    const int step = *epoch;
    volatile const int32_t* f = (volatile const int32_t*)flags;
    long long spins = 0;
    while (f[threadIdx.x] < step) {
        if (++spins > max_spins) {
            atomicExch(timeout_flag, 1 + threadIdx.x);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

// claim: owner[node] = max over (rank, position) codes of the ranks that refreshed `node`
__global__ void __launch_bounds__(256)
wb_claim_kernel(const char* g_even, const char* g_odd,
                const int32_t* __restrict__ epoch, int64_t slot_bytes, int world, int n_bound,
                int32_t* __restrict__ owner, const int32_t* flags, int32_t* timeout_flag,
                long long max_spins, int32_t* done_counter) {
    // optional: count this launch as "everything stream-ordered before it has finished" (see
    // sgcn_history_update: the pipelined step's consumer mark)
    if (done_counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(done_counter, 1);
    if (flags) {
        // peer transport: every block first waits (bounded) until all ranks have published this epoch
        if (threadIdx.x < world) {
            const int step = *epoch;
            volatile const int32_t* f = (volatile const int32_t*)flags;
            long long spins = 0;
            while (f[threadIdx.x] < step) {
                if (++spins > max_spins) {
                    atomicExch(timeout_flag, 1 + threadIdx.x);
                    break;
                }
                __nanosleep(100);
            }
            __threadfence_system();
        }
        __syncthreads();
    }
    const int r = blockIdx.y;
    const char* gathered = (epoch && (*epoch & 1)) ? g_odd : g_even;
    const char* slot = gathered + (int64_t)r * slot_bytes;
    // payloads were written by other GPUs while this kernel may already have been resident: read
    // them through L2 (ld.global.cg / volatile), never through a possibly stale L1 / read-only path
    const int n = min(__ldcg((const int32_t*)slot), n_bound);
    const int32_t* ids = (const int32_t*)(slot + wb_ids_offset());
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        atomicMax(owner + __ldcg(ids + j), r * n_bound + j);
}

// copy: the winning (rank, position) of each node writes its whole row, then releases the claim
template <bool VEC>
__global__ void __launch_bounds__(256)
wb_copy_kernel(const char* g_even, const char* g_odd,
               const int32_t* __restrict__ epoch, int64_t slot_bytes, int world, int n_bound,
               int32_t* __restrict__ owner, float* __restrict__ hist, int64_t ld_h, int D) {
    const int r = blockIdx.y;
    const char* gathered = (epoch && (*epoch & 1)) ? g_odd : g_even;
    const char* slot = gathered + (int64_t)r * slot_bytes;
    const int n = min(__ldcg((const int32_t*)slot), n_bound);
    const int32_t* ids = (const int32_t*)(slot + wb_ids_offset());
    const float* rows = (const float*)(slot + wb_rows_offset(n_bound));
    constexpr int W = VEC ? 4 : 1;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += warps) {
        const int node = __ldcg(ids + j);
        const bool mine = owner[node] == r * n_bound + j;         // uniform across the warp
        if (mine) {
            for (int c = lane * W; c < D; c += 32 * W) {
                if (VEC) *(float4*)(hist + (int64_t)node * ld_h + c) = __ldcg((const float4*)(rows + (int64_t)j * D + c));
                else hist[(int64_t)node * ld_h + c] = __ldcg(rows + (int64_t)j * D + c);
            }
        }
        __syncwarp();
        if (mine && lane == 0) owner[node] = -1;                  // losers only ever compare for equality
    }
}

}  // namespace sgcn

using namespace sgcn;

extern "C" {

int64_t sgcn_wb_payload_bytes(int32_t n_bound, int32_t D) {
    if (n_bound < 0 || D < 0) return -1;
    return wb_payload_bytes(n_bound, D);
}

static int fill_ptrs(PeerPtrs& p, void* const* src, int n, const char* what) {
    for (int k = 0; k < n; ++k) {
        if (!src[k] || (((uintptr_t)src[k]) & 15) != 0) {
            set_error(std::string("invalid argument: ") + what + ": pointers must be non-null and 16-byte aligned");
            return SGCN_EINVAL;
        }
        p.p[k] = (char*)src[k];
    }
    return SGCN_OK;
}

static int pack_blocks(int n_bound, int D) {
    const int64_t work = std::max<int64_t>((int64_t)n_bound * std::max(D / 4, 1), 1);
    return (int)std::min<int64_t>((work + 255) / 256, kNumSMs * 2);
}

int sgcn_wb_pack(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                 int64_t ld_rows, int32_t D, void* const* dst, int32_t n_dst, int32_t step,
                 void* stream) {
    SGCN_REQUIRE(field && n_dev && rows && dst, "wb_pack: null pointer");
    SGCN_REQUIRE(n_bound >= 0 && D > 0 && ld_rows >= D, "wb_pack: bad size");
    SGCN_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers, "wb_pack: 1..16 destinations");
    PeerPtrs p{};
    int rc = fill_ptrs(p, dst, n_dst, "wb_pack");
    if (rc != SGCN_OK) return rc;
    wb_pack_kernel<<<pack_blocks(n_bound, D), 256, 0, (cudaStream_t)stream>>>(
        field, n_dev, n_bound, rows, ld_rows, D, p, p, n_dst, step, nullptr, p, 0, nullptr);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_wb_push(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                 int64_t ld_rows, int32_t D, void* const* dst_even, void* const* dst_odd, int32_t n_dst,
                 void* const* peer_flags, int32_t my_rank, int32_t* epoch, int32_t* block_counter,
                 void* stream) {
    SGCN_REQUIRE(field && n_dev && rows && dst_even && dst_odd && peer_flags && epoch, "wb_push: null pointer");
    SGCN_REQUIRE(n_bound >= 0 && D > 0 && ld_rows >= D, "wb_push: bad size");
    SGCN_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers && my_rank >= 0 && my_rank < kMaxPeers, "wb_push: 1..16 ranks");
    PeerPtrs pe{}, po{}, pf{};
    int rc = fill_ptrs(pe, dst_even, n_dst, "wb_push");
    if (rc == SGCN_OK) rc = fill_ptrs(po, dst_odd, n_dst, "wb_push");
    if (rc == SGCN_OK) rc = fill_ptrs(pf, peer_flags, n_dst, "wb_push");
    if (rc != SGCN_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    wb_pack_kernel<<<pack_blocks(n_bound, D), 256, 0, st>>>(field, n_dev, n_bound, rows, ld_rows, D, pe, po,
                                                           n_dst, 0, epoch, pf, my_rank, block_counter);
    SGCN_LAUNCHED();
    if (!block_counter) {      // no scratch counter given: publish from a second, stream-ordered launch
        wb_signal_kernel<<<1, 32, 0, st>>>(pf, n_dst, my_rank, epoch);
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

static int launch_apply(float* hist, int64_t ld_h, int32_t D, const void* g_even, const void* g_odd,
                        const int32_t* epoch, int64_t slot_bytes, int32_t world, int32_t n_bound,
                        int32_t* owner, cudaStream_t st, const int32_t* flags = nullptr,
                        int32_t* timeout_flag = nullptr, int32_t* done_counter = nullptr) {
    dim3 g1(std::min(div_up(std::max(n_bound, 1), 256), 64), world);
    // ~2 s at 100 ns per spin: a peer that never arrives raises the flag instead of hanging the GPU
    wb_claim_kernel<<<g1, 256, 0, st>>>((const char*)g_even, (const char*)g_odd, epoch, slot_bytes, world,
                                        n_bound, owner, flags, timeout_flag, 20000000LL, done_counter);
    SGCN_LAUNCHED();
    dim3 g2(std::min(div_up(std::max(n_bound, 1), 8), kNumSMs), world);
    const bool vec = D % 4 == 0 && ld_h % 4 == 0 && (((uintptr_t)hist) & 15) == 0 && slot_bytes % 16 == 0 &&
                     (((uintptr_t)g_even) & 15) == 0 && (((uintptr_t)g_odd) & 15) == 0;
    if (vec)
        wb_copy_kernel<true><<<g2, 256, 0, st>>>((const char*)g_even, (const char*)g_odd, epoch, slot_bytes,
                                                 world, n_bound, owner, hist, ld_h, D);
    else
        wb_copy_kernel<false><<<g2, 256, 0, st>>>((const char*)g_even, (const char*)g_odd, epoch, slot_bytes,
                                                  world, n_bound, owner, hist, ld_h, D);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_wb_apply(float* hist, int64_t ld_h, int32_t D, const void* gathered, int64_t slot_bytes,
                  int32_t world, int32_t n_bound, int32_t* owner, void* stream) {
    SGCN_REQUIRE(hist && gathered && owner, "wb_apply: null pointer");
    SGCN_REQUIRE(world >= 1 && n_bound >= 0 && D > 0 && ld_h >= D, "wb_apply: bad size");
    SGCN_REQUIRE(slot_bytes >= wb_payload_bytes(n_bound, D), "wb_apply: slot smaller than a payload");
    SGCN_REQUIRE((int64_t)world * std::max(n_bound, 1) < 0x7fffffff, "wb_apply: world * n_bound overflows");
    if (n_bound == 0) return SGCN_OK;
    return launch_apply(hist, ld_h, D, gathered, gathered, nullptr, slot_bytes, world, n_bound, owner,
                        (cudaStream_t)stream);
}

int sgcn_wb_wait_apply(float* hist, int64_t ld_h, int32_t D, const void* recv_even, const void* recv_odd,
                       int64_t slot_bytes, int32_t world, int32_t n_bound, int32_t* owner,
                       const int32_t* flags, const int32_t* epoch, int32_t* timeout_flag,
                       int32_t* done_counter, void* stream) {
    SGCN_REQUIRE(hist && recv_even && recv_odd && owner && flags && epoch && timeout_flag,
                 "wb_wait_apply: null pointer");
    SGCN_REQUIRE(world >= 1 && world <= 32 && n_bound >= 0 && D > 0 && ld_h >= D, "wb_wait_apply: bad size");
    SGCN_REQUIRE(slot_bytes >= wb_payload_bytes(n_bound, D), "wb_wait_apply: slot smaller than a payload");
    SGCN_REQUIRE((int64_t)world * std::max(n_bound, 1) < 0x7fffffff, "wb_wait_apply: world * n_bound overflows");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_bound == 0) {
        wb_wait_kernel<<<1, 32, 0, st>>>(flags, world, epoch, timeout_flag, 20000000LL);
        SGCN_LAUNCHED();
        return SGCN_OK;
    }
    // the wait is fused into the claim pass (every block polls the flags before touching a payload)
    return launch_apply(hist, ld_h, D, recv_even, recv_odd, epoch, slot_bytes, world, n_bound, owner, st,
                        flags, timeout_flag, done_counter);
}

// ---- peer memory plumbing (cudaIpc): plain cudaMalloc'd buffers that other ranks can map -----
int sgcn_ipc_alloc(void** ptr, int64_t bytes, int32_t zero) {
    SGCN_REQUIRE(ptr && bytes > 0, "ipc_alloc: bad argument");
    SGCN_CUDA(cudaMalloc(ptr, (size_t)bytes));
    if (zero) SGCN_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    return SGCN_OK;
}

int sgcn_ipc_free(void* ptr) {
    if (ptr) SGCN_CUDA(cudaFree(ptr));
    return SGCN_OK;
}

int sgcn_ipc_export(void* ptr, void* handle64) {
    SGCN_REQUIRE(ptr && handle64, "ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    SGCN_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, ptr));
    return SGCN_OK;
}

int sgcn_ipc_open(const void* handle64, void** ptr) {
    SGCN_REQUIRE(handle64 && ptr, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    SGCN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SGCN_OK;
}

int sgcn_ipc_close(void* ptr) {
    if (ptr) SGCN_CUDA(cudaIpcCloseMemHandle(ptr));
    return SGCN_OK;
}

}  // extern "C"
