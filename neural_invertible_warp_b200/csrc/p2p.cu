// Gradient all-reduce over NVLink peer memory (one node, up to 8 GPUs) -- SURVEY.md section 8 row e.
// The reference trains on one GPU (options.py:103); data parallelism over rays is this repository's addition, and the
// only exchange step of a training step is the sum of the flat gradient bucket (0.53 M floats of NeRF weights,
// 0.17 M of pose / warp parameters).  At these sizes an NCCL all-reduce is pure launch + protocol latency
// (~50 us at 8 GPUs, exposed at the end of the step), so the sum is done by two small kernels of our own:
//
//   publish   copies this rank's segment into its exchange buffer (peer-mapped with CUDA IPC), then the last block
//             to finish raises this rank's flag in EVERY peer's flag block (st.release.sys over NVLink)
//   reduce    waits until all ranks' flags in the local flag block carry the current sequence number, then every
//             thread sums the `world` exchange buffers element-wise IN RANK ORDER (bit-identical result on all
//             ranks) with 128-bit loads straight from peer memory, and writes the sum over the local segment
//
// One-shot (2 ranks): every rank reads world * n floats.  From 4 ranks up that is 17 MB per rank for the NeRF segment
// (42 - 59 us in the 8-GPU timeline), so the sum becomes two-shot: each rank reduces ITS 1/world slice into a result
// area of its block (reduce-scatter), raises a second flag, and everybody copies the other slices from their owners
// (all-gather): 2 * n floats per rank for one more launch.  The exchange buffer is double-buffered on the sequence number: a rank
// overwrites half (s & 1) only after it has seen every peer's flag s-1, i.e. after every peer has left reduce s-2.
// The sequence number lives in device memory, so the pair is CUDA-graph capturable.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

struct P2PState {               // one per rank, in that rank's own memory (peer-mapped as part of the flag block)
    unsigned int flag[8];       // flag[r]  = last sequence number rank r has published
    unsigned int seq;           // sequence number of this rank's last publish
    unsigned int blocks_done;   // publish: blocks that have finished copying
    unsigned int error;         // a peer's flag did not arrive within the time-out
    unsigned int blocks_done2;  // two-shot: blocks that have finished their part of the slice
    unsigned int pad[4];
    unsigned int flag2[8];      // two-shot: flag2[r] = last sequence number whose reduced slice rank r has written
    unsigned int pad2[8];
};
static_assert(sizeof(P2PState) == 128, "flag block layout");

struct Peers { float* buf[8]; P2PState* st[8]; };

__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {      // peer memory: no L1 allocation (the line changes every step)
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
// the last block to get here (all blocks' earlier writes released at GPU scope through `counter`) raises this rank's flag
// `which` in every peer: ONE system-scope fence, then plain stores (a release per store costs a fence per peer)
__device__ __forceinline__ void block_done_raise(unsigned int* counter, const Peers& peers, int rank, int world, unsigned int seq, bool second) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            for (int r = 0; r < world; ++r) st_relaxed_sys(second ? &peers.st[r]->flag2[rank] : &peers.st[r]->flag[rank], seq);
            if (!second) peers.st[rank]->seq = seq;
        }
    }
}
// threads 0..world-1 wait until every rank's flag in the local block has reached seq
__device__ __forceinline__ void wait_flags(P2PState* me, const unsigned int* flags, unsigned int seq, int world) {
    if (threadIdx.x < world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(&flags[threadIdx.x]) - seq) < 0) {
            __nanosleep(40);
            if (clock64() - t0 > (1ll << 32)) { me->error = 1u + threadIdx.x; break; }      // ~2 s: a peer is gone; do not hang the GPU
        }
    }
    __syncthreads();
}

__global__ void p2p_publish_kernel(const float* __restrict__ data, int64_t n4, int64_t half_floats, Peers peers, int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq + 1;                     // (me->seq is only written by the last block, in block_done_raise)
    float4* dst = reinterpret_cast<float4*>(peers.buf[rank] + (seq & 1) * half_floats);
    const float4* src = reinterpret_cast<const float4*>(data);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    block_done_raise(&me->blocks_done, peers, rank, world, seq, false);
}

// one-shot: every rank sums all `world` buffers over the whole vector
__global__ void p2p_reduce_kernel(float* __restrict__ data, int64_t n4, int64_t half_floats, Peers peers, int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq;                         // written by the publish kernel before this one in stream order
    wait_flags(me, me->flag, seq, world);
    const int64_t off = (int64_t)(seq & 1) * half_floats;
    float4* out = reinterpret_cast<float4*>(data);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = ld_peer4(peers.buf[0] + off + 4 * i);
        for (int r = 1; r < world; ++r) {                     // fixed order: the same bits on every rank
            const float4 v = ld_peer4(peers.buf[r] + off + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}

// two-shot, step 1 (reduce-scatter): this rank sums ITS slice of the vector over all ranks' buffers (rank order), keeps it in
// its result area (behind the two exchange halves) and in `data`, and raises flag2 in every peer
__global__ void p2p_reduce_slice_kernel(float* __restrict__ data, int64_t n4, int64_t slice4, int64_t half_floats, Peers peers,
                                        int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq;
    wait_flags(me, me->flag, seq, world);
    const int64_t off = (int64_t)(seq & 1) * half_floats;
    const int64_t lo = rank * slice4, hi = lo + slice4 < n4 ? lo + slice4 : n4;
    float4* out = reinterpret_cast<float4*>(data);
    float4* res = reinterpret_cast<float4*>(peers.buf[rank] + 2 * half_floats + off);
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = ld_peer4(peers.buf[0] + off + 4 * i);
        for (int r = 1; r < world; ++r) {
            const float4 v = ld_peer4(peers.buf[r] + off + 4 * i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
        res[i] = acc;
    }
    block_done_raise(&me->blocks_done2, peers, rank, world, seq, true);
}
// two-shot, step 2 (all-gather): the other ranks' reduced slices, straight from their result areas
__global__ void p2p_gather_kernel(float* __restrict__ data, int64_t n4, int64_t slice4, int64_t half_floats, Peers peers,
                                  int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq;
    wait_flags(me, me->flag2, seq, world);
    const int64_t off = (int64_t)(seq & 1) * half_floats;
    float4* out = reinterpret_cast<float4*>(data);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / slice4);
        if (r != rank) out[i] = ld_peer4(peers.buf[r] + 2 * half_floats + off + 4 * i);
    }
}

}  // namespace

// ---- CUDA IPC plumbing: each rank allocates its exchange buffer + flag block, hands the 64-byte handle to its peers ----
extern "C" int niw_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64) {
    NIW_CHECK_ARG(bytes > 0 && dev_ptr && handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    NIW_CUDA(cudaMalloc(dev_ptr, bytes));
    NIW_CUDA(cudaMemset(*dev_ptr, 0, bytes));
    NIW_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), *dev_ptr));
    return 0;
}
extern "C" int niw_p2p_open(const void* handle64, void** dev_ptr) {
    NIW_CHECK_ARG(handle64 && dev_ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    NIW_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int niw_p2p_close(void* dev_ptr) { NIW_CHECK_ARG(dev_ptr); NIW_CUDA(cudaIpcCloseMemHandle(dev_ptr)); return 0; }
extern "C" int niw_p2p_free(void* dev_ptr) { NIW_CHECK_ARG(dev_ptr); NIW_CUDA(cudaFree(dev_ptr)); return 0; }

// In-place sum of `n` floats (n % 4 == 0, 16-byte aligned) over `world` ranks.  blocks[r]: rank r's allocation as mapped
// in THIS process (own pointer for r == rank): a 64-byte flag block followed, at byte 256, by 2 * half_floats floats of
// exchange buffer and 2 * half_floats floats of result area (two-shot); n <= half_floats.  Every rank must make the same
// sequence of calls on its block set.
extern "C" int niw_allreduce_p2p(float* data, int64_t n, void* const* blocks, int rank, int world, int64_t half_floats,
                                 void* stream) {
    NIW_CHECK_ARG(data && blocks && n > 0 && world >= 1 && world <= 8 && rank >= 0 && rank < world && n <= half_floats);
    if ((n & 3) || (half_floats & 3) || !niw_aligned16(data)) return NIW_E_UNSUPP;
    Peers p{};
    for (int r = 0; r < world; ++r) {
        if (!blocks[r]) return NIW_E_BADARG;
        p.st[r] = reinterpret_cast<P2PState*>(blocks[r]);
        p.buf[r] = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(blocks[r]) + 256);
    }
    const int64_t n4 = n / 4;
    const unsigned grid = (unsigned)((n4 + 255) / 256 < 2 * niw_num_sms() ? (n4 + 255) / 256 : 2 * niw_num_sms());
    cudaStream_t st = niw_stream(stream);
    niw::note_launch(), p2p_publish_kernel<<<grid, 256, 0, st>>>(data, n4, half_floats, p, rank, world);
    // NIW_P2P_TWO_SHOT: 0 never, 1 always, default: from 4 ranks up (one-shot reads world * n floats per rank, two-shot 2 * n
    // for one more launch and flag round)
    static const int two_shot_env = getenv("NIW_P2P_TWO_SHOT") ? atoi(getenv("NIW_P2P_TWO_SHOT")) : -1;
    const bool two_shot = two_shot_env >= 0 ? two_shot_env != 0 : world >= 4;
    if (two_shot) {
        const int64_t slice4 = (n4 + world - 1) / world;
        const unsigned g1 = (unsigned)((slice4 + 255) / 256 < (int64_t)niw_num_sms() ? (slice4 + 255) / 256 : niw_num_sms());
        niw::note_launch(), p2p_reduce_slice_kernel<<<g1 < 1 ? 1 : g1, 256, 0, st>>>(data, n4, slice4, half_floats, p, rank, world);
        niw::note_launch(), p2p_gather_kernel<<<grid, 256, 0, st>>>(data, n4, slice4, half_floats, p, rank, world);
    } else {
        niw::note_launch(), p2p_reduce_kernel<<<grid, 256, 0, st>>>(data, n4, half_floats, p, rank, world);
    }
    NIW_LAUNCH_CHECK();
    return 0;
}

// the time-out flag of rank `rank`'s block (0 = every wait so far was satisfied); synchronises the device
extern "C" int niw_p2p_error(const void* block, unsigned int* error) {
    NIW_CHECK_ARG(block && error);
    NIW_CUDA(cudaMemcpy(error, reinterpret_cast<const uint8_t*>(block) + 40, 4, cudaMemcpyDeviceToHost));
    return 0;
}
