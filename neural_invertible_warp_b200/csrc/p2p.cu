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
// One-shot: every rank reads world * n floats (17 MB at 8 GPUs for the NeRF segment, NVSwitch gives each GPU its
// full NVLink bandwidth to all peers at once).  The exchange buffer is double-buffered on the sequence number: a rank
// overwrites half (s & 1) only after it has seen every peer's flag s-1, i.e. after every peer has left reduce s-2.
// The sequence number lives in device memory, so the pair is CUDA-graph capturable.
#include "common.cuh"
#include <string.h>

namespace {

struct P2PState {               // one per rank, in that rank's own memory (peer-mapped as part of the flag block)
    unsigned int flag[8];       // flag[r] = last sequence number rank r has published
    unsigned int seq;           // sequence number of this rank's last publish
    unsigned int blocks_done;   // publish: blocks that have finished copying
    unsigned int error;         // reduce: a peer's flag did not arrive within the time-out
    unsigned int pad[5];
};
static_assert(sizeof(P2PState) == 64, "flag block layout");

struct Peers { float* buf[8]; P2PState* st[8]; };

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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

__global__ void p2p_publish_kernel(const float* __restrict__ data, int64_t n4, int64_t half_floats, Peers peers, int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq + 1;                     // (me->seq is only written by the last block, below)
    float4* dst = reinterpret_cast<float4*>(peers.buf[rank] + (seq & 1) * half_floats);
    const float4* src = reinterpret_cast<const float4*>(data);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                      // the block's copies (ordered before this by the barrier): released at GPU
        const unsigned int ticket = atomicAdd(&me->blocks_done, 1u);   // scope through the counter; the last block carries them to system scope
        if (ticket == gridDim.x - 1) {
            me->blocks_done = 0;
            __threadfence_system();                           // every block's copy (observed through the counter) before the flags
            for (int r = 0; r < world; ++r) st_release_sys(&peers.st[r]->flag[rank], seq);
            me->seq = seq;
        }
    }
}

__global__ void p2p_reduce_kernel(float* __restrict__ data, int64_t n4, int64_t half_floats, Peers peers, int rank, int world) {
    P2PState* me = peers.st[rank];
    const unsigned int seq = me->seq;                         // written by the publish kernel before this one in stream order
    if (threadIdx.x < world) {
        long long t0 = clock64();
        while ((int)(ld_acquire_sys(&me->flag[threadIdx.x]) - seq) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > (1ll << 32)) { me->error = 1u + threadIdx.x; break; }      // ~2 s: a peer is gone; do not hang the GPU
        }
    }
    __syncthreads();
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
// exchange buffer; n <= half_floats.  Every rank must make the same sequence of calls on its block set.
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
    niw::note_launch(), p2p_reduce_kernel<<<grid, 256, 0, st>>>(data, n4, half_floats, p, rank, world);
    NIW_LAUNCH_CHECK();
    return 0;
}

// the time-out flag of rank `rank`'s block (0 = every wait so far was satisfied); synchronises the device
extern "C" int niw_p2p_error(const void* block, unsigned int* error) {
    NIW_CHECK_ARG(block && error);
    NIW_CUDA(cudaMemcpy(error, reinterpret_cast<const uint8_t*>(block) + 40, 4, cudaMemcpyDeviceToHost));
    return 0;
}
