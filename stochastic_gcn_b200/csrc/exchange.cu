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
#include "exchange.cuh"

namespace sgcn {

// pack {count, ids, rows} into up to `n_dst` destinations (own buffer and/or peers' receive slots)
__global__ void __launch_bounds__(256)
wb_pack_kernel(const WbPushArgs a, unsigned long long* trace) {
    TraceScope ts(trace, TR_WB_PUSH);
    wb_pack_body(a, blockIdx.x, gridDim.x);
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
                long long max_spins, int32_t* done_counter, unsigned long long* trace,
                int ring, int64_t ring_stride, int32_t* cur_stash,
                int shard_rank, int shard_rows, PeerPtrs reads_peers, const int32_t* reads_flags) {
    TraceScope ts(trace, TR_WB_CLAIM);
    // ring protocol: `epoch` counts the epochs APPLIED so far (the pushes run ahead on a counter of their own);
    // this launch applies epoch *epoch + 1.  Under programmatic launches the only ordering between the kernels
    // of the chain is "a dependent launches after EVERY block of its predecessor has executed
    // launch_dependents", so each kernel reads the counter it needs BEFORE that instruction and publishes what
    // its successor needs BEFORE it too: this kernel stashes the epoch for its copy kernel, the copy kernel
    // advances *epoch for the next claim (which can only launch two kernels later).
    const int cur = epoch ? (ring > 0 ? *(volatile const int32_t*)epoch + 1 : *epoch) : 0;
    if (ring > 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
            *(volatile int32_t*)cur_stash = cur;
            __threadfence();
        }
        __syncthreads();
    }
    // (PDL) this kernel is launched while its stream predecessor -- the full-neighbour mean -- is still
    // running: waiting for the peers' payloads and the claim pass touch nothing the mean reads, so they
    // overlap it; the dependent copy kernel may become resident right away
    asm volatile("griddepcontrol.launch_dependents;");
    // without flags (NCCL transport) the payloads come from the stream predecessor itself: order first
    if (!flags) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (flags) {
        // peer transport: every block first waits (bounded) until all ranks have published this epoch
        if (threadIdx.x < world) {
            const int step = cur;
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
    if (shard_rows > 0) {
        // Sharded history (plain launch: everything stream-ordered before it has finished, i.e. this rank
        // has read all it needs of the table as of the previous epoch): tell every rank, and do not touch
        // a row before every rank has said the same -- their full-neighbour means read MY shard over NVLink.
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < world) {
            volatile int32_t* f = (volatile int32_t*)reads_peers.p[threadIdx.x];
            f[shard_rank] = cur;
            __threadfence_system();
        }
        if (threadIdx.x < world) {
            volatile const int32_t* f = (volatile const int32_t*)reads_flags;
            long long spins = 0;
            while (f[threadIdx.x] < cur) {
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
    const char* gathered = ring > 0 ? g_even + (int64_t)(cur % ring) * ring_stride : ((cur & 1) ? g_odd : g_even);
    const char* slot = gathered + (int64_t)r * slot_bytes;
    // payloads were written by other GPUs while this kernel may already have been resident: read
    // them through L2 (ld.global.cg / volatile), never through a possibly stale L1 / read-only path
    const int n = min(__ldcg((const int32_t*)slot), n_bound);
    const int32_t* ids = (const int32_t*)(slot + wb_ids_offset());
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        int node = __ldcg(ids + j);
        if (shard_rows > 0) {                       // only the rows this rank owns, by local row index
            if (node / shard_rows != shard_rank) continue;
            node -= shard_rank * shard_rows;
        }
        atomicMax(owner + node, r * n_bound + j);
    }
    // (PDL) completion of this grid must imply completion of the predecessor (the copy kernel orders
    // its history writes behind THIS grid only): wait for it here, after the independent work
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // optional: count this launch as "everything stream-ordered before it has finished" (see
    // sgcn_history_update: the pipelined step's consumer mark)
    if (done_counter && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(done_counter, 1);
}

// copy: the winning (rank, position) of each node writes its whole row, then releases the claim
template <bool VEC>
__global__ void __launch_bounds__(256)
wb_copy_kernel(const char* g_even, const char* g_odd,
               const int32_t* __restrict__ epoch, int64_t slot_bytes, int world, int n_bound,
               int32_t* __restrict__ owner, float* __restrict__ hist, int64_t ld_h, int D,
               unsigned long long* trace, int ring, int64_t ring_stride, int32_t* epoch_out,
               const int32_t* cur_stash, int shard_rank, int shard_rows, PeerPtrs applied_peers,
               int32_t* apply_counter, const int late_trigger) {
    TraceScope ts(trace, TR_WB_COPY);
    // ring protocol: the epoch being applied was stashed by the claim kernel; advance the applied-epoch
    // counter for the NEXT claim before any dependent can launch (see wb_claim_kernel)
    const int cur = ring > 0 ? *(volatile const int32_t*)cur_stash : (epoch ? *epoch : 0);
    if (ring > 0) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
            *(volatile int32_t*)epoch_out = cur;
            __threadfence();
        }
        __syncthreads();
    }
    // (PDL) the stream successor is the next full-neighbour mean: at entry, or (late_trigger, see
    // g_wb_late_trigger) once this block's rows are stored
    if (!late_trigger) asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");      // (PDL) claims final, history no longer read
    const int r = blockIdx.y;
    const char* gathered = ring > 0 ? g_even + (int64_t)(cur % ring) * ring_stride : ((cur & 1) ? g_odd : g_even);
    const char* slot = gathered + (int64_t)r * slot_bytes;
    const int n = min(__ldcg((const int32_t*)slot), n_bound);
    const int32_t* ids = (const int32_t*)(slot + wb_ids_offset());
    const float* rows = (const float*)(slot + wb_rows_offset(n_bound));
    constexpr int W = VEC ? 4 : 1;
    constexpr int R = 4;                                          // rows a warp moves per round trip
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int j0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * R; j0 < n; j0 += warps * R) {
        int node[R];
        bool mine[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            node[q] = j0 + q < n ? __ldcg(ids + j0 + q) : -1;
            if (shard_rows > 0 && node[q] >= 0)      // sharded history: local row index, or not mine at all
                node[q] = node[q] / shard_rows == shard_rank ? node[q] - shard_rank * shard_rows : -1;
        }
#pragma unroll
        for (int q = 0; q < R; ++q) mine[q] = node[q] >= 0 && owner[node[q]] == r * n_bound + j0 + q;   // warp-uniform
        for (int c = lane * W; c < D; c += 32 * W) {
            if (VEC) {
                float4 v[R];
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (mine[q]) v[q] = __ldcg((const float4*)(rows + (int64_t)(j0 + q) * D + c));
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (mine[q]) *(float4*)(hist + (int64_t)node[q] * ld_h + c) = v[q];
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q)
                    if (mine[q]) hist[(int64_t)node[q] * ld_h + c] = __ldcg(rows + (int64_t)(j0 + q) * D + c);
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R; ++q)
            if (mine[q] && lane == 0) owner[node[q]] = -1;        // losers only ever compare for equality
    }
    if (late_trigger) asm volatile("griddepcontrol.launch_dependents;");
    if (shard_rows > 0) {      // sharded history: the last block to finish tells every rank "my shard holds epoch cur"
        __shared__ int s_last;
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(apply_counter, 1) == (int)(gridDim.x * gridDim.y) - 1;
        __syncthreads();
        if (s_last && threadIdx.x < world) {
            if (threadIdx.x == 0) *apply_counter = 0;
            volatile int32_t* f = (volatile int32_t*)applied_peers.p[threadIdx.x];
            f[shard_rank] = cur;
            __threadfence_system();
        }
    }
}

// copy, flat form (ring protocol, replicated tables; sgcn_wb_copy_ring): the claims of this epoch were taken by a
// launch that has FINISHED (sgcn_wb_claim_ring on another stream, ordered by an event), so everything up to the
// stores is independent of the stream predecessor -- the full-neighbour mean still reading the table -- and runs
// before griddepcontrol.wait: ids, claim look-ups and the winners' row vectors land in registers (one thread =
// up to 8 independent 16-byte vectors of a flattened (rank, row, vector) space), behind the wait only the stores
// are left.  Rows never straddle warps (D / 4 divides 32, checked by the launcher): every
// lane that compares a claim has done so before the winner's first lane releases it.
constexpr int kCopyFlatU = 8;
__global__ void __launch_bounds__(256)
wb_copy_flat_kernel(const char* recv_base, int64_t slot_bytes, int world, int n_bound, int32_t* __restrict__ owner,
                    float* __restrict__ hist, int64_t ld_h, int D, unsigned long long* trace, int ring,
                    int64_t ring_stride, int32_t* epoch_out, const int32_t* cur_stash, int32_t* done_counter,
                    const int late_trigger) {
    TraceScope ts(trace, TR_WB_COPY);
    __shared__ int s_n[kMaxPeers];
    const int cur = *(volatile const int32_t*)cur_stash;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *(volatile int32_t*)epoch_out = cur;       // applied-epoch counter for the NEXT claim (launched after this grid)
        __threadfence();
    }
    if (!late_trigger) asm volatile("griddepcontrol.launch_dependents;");     // (PDL) see wb_copy_kernel
    const char* gathered = recv_base + (int64_t)(cur % ring) * ring_stride;
    if (threadIdx.x < world)
        s_n[threadIdx.x] = min(__ldcg((const int32_t*)(gathered + (int64_t)threadIdx.x * slot_bytes)), n_bound);
    __syncthreads();
    const int c4 = D >> 2;
    const int64_t per_rank = (int64_t)n_bound * c4;
    const int64_t total = per_rank * world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool waited = false;
    // (trip count uniform within a warp: the __syncwarp below is executed by all 32 lanes or by none)
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base - (threadIdx.x & 31) < total || !waited;
         base += stride * kCopyFlatU) {
        float4 v[kCopyFlatU];
        int64_t off[kCopyFlatU];
        int32_t rel[kCopyFlatU];                   // node whose claim this thread releases (first vector of a row)
#pragma unroll
        for (int u = 0; u < kCopyFlatU; ++u) {
            const int64_t e = base + (int64_t)u * stride;
            off[u] = -1;
            rel[u] = -1;
            if (e < total) {
                const int r = (int)(e / per_rank);
                const int64_t rem = e - (int64_t)r * per_rank;
                const int j = (int)(rem / c4);
                const int c = (int)(rem - (int64_t)j * c4);
                if (j < s_n[r]) {
                    const char* slot = gathered + (int64_t)r * slot_bytes;
                    const int node = __ldcg((const int32_t*)(slot + wb_ids_offset()) + j);
                    if (__ldcg(owner + node) == r * n_bound + j) {
                        v[u] = __ldcg((const float4*)((const float*)(slot + wb_rows_offset(n_bound)) + (int64_t)j * D) + c);
                        off[u] = (int64_t)node * ld_h + c * 4;
                        if (c == 0) rel[u] = node;
                    }
                }
            }
        }
        if (!waited) {
            asm volatile("griddepcontrol.wait;" ::: "memory");  // (PDL) the table is no longer read
            if (done_counter && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(done_counter, 1);
            waited = true;
        }
        __syncwarp();                              // every lane of the row has compared the claim
#pragma unroll
        for (int u = 0; u < kCopyFlatU; ++u) {
            if (off[u] >= 0) *(float4*)(hist + off[u]) = v[u];
            if (rel[u] >= 0) owner[rel[u]] = -1;
        }
    }
    if (late_trigger) asm volatile("griddepcontrol.launch_dependents;");
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

// Thread blocks of a pack / push launch.  Few enough to be resident all at once beside a full-neighbour mean (a
// launch with blocks still waiting for room holds up every later launch), enough to keep the peer stores flowing:
// 64 for up to 4 destinations (2 GPUs: 28.8 us per pass against 29.5 with 192 and 35.7 with 24; 4 GPUs: 39.7
// against 41.8), 192 beyond (the measured 8-GPU form).  SGCN_PUSH_BLOCKS overrides.
static int pack_blocks(int n_bound, int D, int n_dst = 1) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("SGCN_PUSH_BLOCKS");
        forced = e ? std::max(1, std::min(atoi(e), kNumSMs * 2)) : 0;
    }
    const int cap = forced > 0 ? forced : (n_dst > 4 ? 192 : 64);
    const int64_t work = std::max<int64_t>((int64_t)n_bound * std::max(D / 4, 1), 1);
    return (int)std::min<int64_t>((work + 255) / 256, cap);
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
    WbPushArgs a{field, n_dev, n_bound, rows, ld_rows, D, p, p, n_dst, step, nullptr, p, 0, nullptr, 0, 0};
    wb_pack_kernel<<<pack_blocks(n_bound, D), 256, 0, (cudaStream_t)stream>>>(a, g_trace);
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
    WbPushArgs a{field, n_dev, n_bound, rows, ld_rows, D, pe, po, n_dst, 0, epoch, pf, my_rank, block_counter, 0, 0};
    wb_pack_kernel<<<pack_blocks(n_bound, D), 256, 0, st>>>(a, g_trace);
    SGCN_LAUNCHED();
    if (!block_counter) {      // no scratch counter given: publish from a second, stream-ordered launch
        wb_signal_kernel<<<1, 32, 0, st>>>(pf, n_dst, my_rank, epoch);
        SGCN_LAUNCHED();
    }
    return SGCN_OK;
}

int sgcn_wb_push_ring(const int32_t* field, const int32_t* n_dev, int32_t n_bound, const float* rows,
                      int64_t ld_rows, int32_t D, void* const* dst_base, int32_t n_dst, int32_t ring,
                      int64_t ring_stride, void* const* peer_flags, int32_t my_rank, int32_t* push_epoch,
                      int32_t* block_counter, void* stream) {
    SGCN_REQUIRE(field && n_dev && rows && dst_base && peer_flags && push_epoch && block_counter,
                 "wb_push_ring: null pointer");
    SGCN_REQUIRE(n_bound >= 0 && D > 0 && ld_rows >= D, "wb_push_ring: bad size");
    SGCN_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers && my_rank >= 0 && my_rank < kMaxPeers, "wb_push_ring: 1..16 ranks");
    SGCN_REQUIRE(ring >= 2 && ring <= 64 && ring_stride > 0 && ring_stride % 16 == 0, "wb_push_ring: bad ring");
    PeerPtrs pb{}, pf{};
    int rc = fill_ptrs(pb, dst_base, n_dst, "wb_push_ring");
    if (rc == SGCN_OK) rc = fill_ptrs(pf, peer_flags, n_dst, "wb_push_ring");
    if (rc != SGCN_OK) return rc;
    WbPushArgs a{field, n_dev, n_bound, rows, ld_rows, D, pb, pb, n_dst, 0, push_epoch, pf, my_rank, block_counter,
                 ring, ring_stride};
    wb_pack_kernel<<<pack_blocks(n_bound, D, n_dst), 256, 0, (cudaStream_t)stream>>>(a, g_trace);
    SGCN_LAUNCHED();
    return SGCN_OK;
}

static int launch_apply(float* hist, int64_t ld_h, int32_t D, const void* g_even, const void* g_odd,
                        const int32_t* epoch, int64_t slot_bytes, int32_t world, int32_t n_bound,
                        int32_t* owner, cudaStream_t st, const int32_t* flags = nullptr,
                        int32_t* timeout_flag = nullptr, int32_t* done_counter = nullptr, int ring = 0,
                        int64_t ring_stride = 0, int32_t* epoch_out = nullptr, int32_t* apply_counter = nullptr,
                        int shard_rank = 0, int shard_rows = 0, const PeerPtrs* reads_peers = nullptr,
                        const int32_t* reads_flags = nullptr, const PeerPtrs* applied_peers = nullptr,
                        int32_t* shard_counter = nullptr) {
    const PeerPtrs none{};
    dim3 g1(std::min(div_up(std::max(n_bound, 1), 256), 64), world);
    // ~2 s at 100 ns per spin: a peer that never arrives raises the flag instead of hanging the GPU
    SGCN_CUDA(launch_pdl(wb_claim_kernel, g1, dim3(256), 0, st, (const char*)g_even, (const char*)g_odd, epoch,
                         slot_bytes, world, n_bound, owner, flags, timeout_flag, 20000000LL, done_counter, g_trace,
                         ring, ring_stride, apply_counter, shard_rank, shard_rows, reads_peers ? *reads_peers : none,
                         reads_flags));
    SGCN_LAUNCHED();
    // A/B (SGCN_WB_COPY_PDL=0): the copy pass launched plainly instead of programmatically.  Launched
    // programmatically it is resident for the whole full-neighbour mean (256 threads x 64 registers per block take
    // every register the mean leaves free on up to 96 SMs) -- yet the plain form measured SLOWER on 2 GPUs (37.6 vs
    // 28.6 us per pass): its launch latency lands on the chain, and the gather / push of the next pass did not
    // start any earlier (profiles/r02_timeline_2gpu_copy_plain.txt).
    static int copy_pdl = -1;
    if (copy_pdl < 0) {
        const char* e = getenv("SGCN_WB_COPY_PDL");
        copy_pdl = !(e && e[0] == '0');
    }
    PdlOff copy_plain(ring > 0 && shard_rows == 0 && !copy_pdl);
    // few CTAs: with PDL they sit resident beside the full-neighbour mean, and the next batch's sampler
    // CTA still has to find an SM with registers to spare
    dim3 g2(std::max(1, std::min(div_up(std::max(n_bound, 1), 32), 96 / std::max(world, 1))), world);
    const bool vec = D % 4 == 0 && ld_h % 4 == 0 && (((uintptr_t)hist) & 15) == 0 && slot_bytes % 16 == 0 &&
                     (((uintptr_t)g_even) & 15) == 0 && (((uintptr_t)g_odd) & 15) == 0;
    if (vec)
        SGCN_CUDA(launch_pdl(wb_copy_kernel<true>, g2, dim3(256), 0, st, (const char*)g_even, (const char*)g_odd,
                             epoch, slot_bytes, world, n_bound, owner, hist, ld_h, D, g_trace, ring, ring_stride,
                             epoch_out, apply_counter, shard_rank, shard_rows, applied_peers ? *applied_peers : none,
                             shard_counter, g_wb_late_trigger));
    else
        SGCN_CUDA(launch_pdl(wb_copy_kernel<false>, g2, dim3(256), 0, st, (const char*)g_even, (const char*)g_odd,
                             epoch, slot_bytes, world, n_bound, owner, hist, ld_h, D, g_trace, ring, ring_stride,
                             epoch_out, apply_counter, shard_rank, shard_rows, applied_peers ? *applied_peers : none,
                             shard_counter, g_wb_late_trigger));
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

int sgcn_wb_wait_apply_ring(float* hist, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes,
                            int32_t world, int32_t n_bound, int32_t* owner, const int32_t* flags, int32_t ring,
                            int64_t ring_stride, int32_t* apply_epoch, int32_t* apply_counter,
                            int32_t* timeout_flag, int32_t* done_counter, void* stream) {
    SGCN_REQUIRE(hist && recv_base && owner && flags && apply_epoch && apply_counter && timeout_flag,
                 "wb_wait_apply_ring: null pointer");
    SGCN_REQUIRE(world >= 1 && world <= 32 && n_bound > 0 && D > 0 && ld_h >= D, "wb_wait_apply_ring: bad size");
    SGCN_REQUIRE(slot_bytes >= wb_payload_bytes(n_bound, D), "wb_wait_apply_ring: slot smaller than a payload");
    SGCN_REQUIRE(ring >= 2 && ring <= 64 && ring_stride >= (int64_t)world * slot_bytes && ring_stride % 16 == 0,
                 "wb_wait_apply_ring: bad ring");
    SGCN_REQUIRE((int64_t)world * n_bound < 0x7fffffff, "wb_wait_apply_ring: world * n_bound overflows");
    return launch_apply(hist, ld_h, D, recv_base, recv_base, apply_epoch, slot_bytes, world, n_bound, owner,
                        (cudaStream_t)stream, flags, timeout_flag, done_counter, ring, ring_stride, apply_epoch,
                        apply_counter);
}

int sgcn_wb_claim_ring(const void* recv_base, int64_t slot_bytes, int32_t world, int32_t n_bound, int32_t* owner,
                       const int32_t* flags, int32_t ring, int64_t ring_stride, const int32_t* apply_epoch,
                       int32_t* apply_stash, int32_t* timeout_flag, void* stream) {
    SGCN_REQUIRE(recv_base && owner && flags && apply_epoch && apply_stash && timeout_flag, "wb_claim_ring: null pointer");
    SGCN_REQUIRE(world >= 1 && world <= kMaxPeers && n_bound > 0, "wb_claim_ring: bad size");
    SGCN_REQUIRE(ring >= 2 && ring <= 64 && ring_stride >= (int64_t)world * slot_bytes && ring_stride % 16 == 0,
                 "wb_claim_ring: bad ring");
    SGCN_REQUIRE((int64_t)world * n_bound < 0x7fffffff, "wb_claim_ring: world * n_bound overflows");
    const PeerPtrs none{};
    dim3 g1(std::min(div_up(n_bound, 256), 64), world);
    PdlOff plain;           // ordered by events: after the previous epoch's copy, before this epoch's
    SGCN_CUDA(launch_pdl(wb_claim_kernel, g1, dim3(256), 0, (cudaStream_t)stream, (const char*)recv_base,
                         (const char*)recv_base, apply_epoch, slot_bytes, world, n_bound, owner, flags, timeout_flag,
                         20000000LL, (int32_t*)nullptr, g_trace, ring, ring_stride, apply_stash, 0, 0, none,
                         (const int32_t*)nullptr));
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_wb_copy_ring(float* hist, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes, int32_t world,
                      int32_t n_bound, int32_t* owner, int32_t ring, int64_t ring_stride, int32_t* apply_epoch,
                      const int32_t* apply_stash, int32_t* done_counter, void* stream) {
    SGCN_REQUIRE(hist && recv_base && owner && apply_epoch && apply_stash, "wb_copy_ring: null pointer");
    SGCN_REQUIRE(world >= 1 && world <= kMaxPeers && n_bound > 0 && D > 0 && ld_h >= D, "wb_copy_ring: bad size");
    SGCN_REQUIRE(slot_bytes >= wb_payload_bytes(n_bound, D), "wb_copy_ring: slot smaller than a payload");
    SGCN_REQUIRE(ring >= 2 && ring <= 64 && ring_stride >= (int64_t)world * slot_bytes && ring_stride % 16 == 0,
                 "wb_copy_ring: bad ring");
    const int c4 = D / 4;
    SGCN_REQUIRE(D % 4 == 0 && ld_h % 4 == 0 && (((uintptr_t)hist) & 15) == 0 && slot_bytes % 16 == 0 &&
                     (((uintptr_t)recv_base) & 15) == 0 && c4 <= 32 && 32 % c4 == 0,
                 "wb_copy_ring: rows must be 16-byte aligned and 1, 2, 4, 8, 16 or 32 vectors of 16 bytes wide; use "
                 "sgcn_wb_wait_apply_ring");
    const int64_t total = (int64_t)world * n_bound * c4;
    // one round of 8 vectors per thread when it fits in 96 thread blocks (they sit resident beside the mean)
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(96, (total + 256 * kCopyFlatU - 1) / (256 * kCopyFlatU)));
    SGCN_CUDA(launch_pdl(wb_copy_flat_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (const char*)recv_base,
                         slot_bytes, world, n_bound, owner, hist, ld_h, D, g_trace, ring, ring_stride, apply_epoch,
                         apply_stash, done_counter, g_wb_late_trigger));
    SGCN_LAUNCHED();
    return SGCN_OK;
}

int sgcn_wb_wait_apply_sharded(float* hist_shard, int64_t ld_h, int32_t D, const void* recv_base, int64_t slot_bytes,
                               int32_t world, int32_t rank, int32_t rows_per_shard, int32_t n_bound, int32_t* owner,
                               const int32_t* flags, int32_t ring, int64_t ring_stride, int32_t* apply_epoch,
                               int32_t* apply_stash, const int32_t* reads_flags, void* const* reads_peer_flags,
                               const int32_t* applied_flags, void* const* applied_peer_flags, int32_t* shard_counter,
                               int32_t* timeout_flag, int32_t* done_counter, void* stream) {
    SGCN_REQUIRE(hist_shard && recv_base && owner && flags && apply_epoch && apply_stash && reads_flags &&
                     reads_peer_flags && applied_flags && applied_peer_flags && shard_counter && timeout_flag,
                 "wb_wait_apply_sharded: null pointer");
    SGCN_REQUIRE(world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world && rows_per_shard > 0 && n_bound > 0 &&
                     D > 0 && ld_h >= D, "wb_wait_apply_sharded: bad size");
    SGCN_REQUIRE(slot_bytes >= wb_payload_bytes(n_bound, D), "wb_wait_apply_sharded: slot smaller than a payload");
    SGCN_REQUIRE(ring >= 2 && ring <= 64 && ring_stride >= (int64_t)world * slot_bytes && ring_stride % 16 == 0,
                 "wb_wait_apply_sharded: bad ring");
    PeerPtrs pr{}, pa{};
    int rc = fill_ptrs(pr, reads_peer_flags, world, "wb_wait_apply_sharded");
    if (rc == SGCN_OK) rc = fill_ptrs(pa, applied_peer_flags, world, "wb_wait_apply_sharded");
    if (rc != SGCN_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PdlOff plain;      // the handshakes below rely on "everything stream-ordered before this launch has finished"
    rc = launch_apply(hist_shard, ld_h, D, recv_base, recv_base, apply_epoch, slot_bytes, world, n_bound, owner, st,
                      flags, timeout_flag, done_counter, ring, ring_stride, apply_epoch, apply_stash, rank,
                      rows_per_shard, &pr, reads_flags, &pa, shard_counter);
    if (rc != SGCN_OK) return rc;
    // nobody reads a shard before its owner has applied this epoch: wait for every rank's "applied" flag
    wb_wait_kernel<<<1, 32, 0, st>>>(applied_flags, world, apply_epoch, timeout_flag, 20000000LL);
    SGCN_LAUNCHED();
    return SGCN_OK;
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
