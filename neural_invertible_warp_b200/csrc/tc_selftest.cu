// tcgen05 self-test: D[128,N] = A[128,K] . B[N,K]^T with BF16 operands and FP32 accumulation in
// TMEM, staged through the exact shared-memory layouts / descriptors the MLP kernels use.
//   variant 0: A and B K-major   (forward / dX GEMMs)        LBO = K-direction, SBO = M/N-direction
//   variant 1: as 0 with LBO and SBO exchanged                (descriptor-semantics probe)
//   variant 2: A and B MN-major  (dW GEMM: K = samples)       LBO = K-direction, SBO = MN-direction
//   variant 3: as 2 with LBO and SBO exchanged
// It exists so that descriptor semantics are verified on hardware by tests/test_gpu_tc.py before
// the fused kernels rely on them.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace niw {
namespace {

__global__ void __launch_bounds__(128)
tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int N, int K, int variant,
                   float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* Bs = As + 128 * K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool mn_major = variant >= 2;
    for (int i = tid; i < 128 * K; i += 128) {
        int m = i / K, k = i % K;
        size_t off = mn_major ? ((size_t)(m >> 3) * K + k) * 8 + (m & 7) : ((size_t)(k >> 3) * 128 + m) * 8 + (k & 7);
        As[off] = __float2bfloat16(A[i]);
    }
    for (int i = tid; i < N * K; i += 128) {
        int n = i / K, k = i % K;
        size_t off = mn_major ? ((size_t)(n >> 3) * K + k) * 8 + (n & 7) : ((size_t)(k >> 3) * N + n) * 8 + (k & 7);
        Bs[off] = __float2bfloat16(Bm[i]);
    }
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_base_s, (uint32_t)(N < 32 ? 32 : N));
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t a0 = ptx::smem_addr(As), b0 = ptx::smem_addr(Bs);
        const uint32_t idesc = ptx::idesc_bf16(128, N, mn_major, mn_major);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint32_t a_addr, b_addr, a_k, a_mn, b_k, b_mn;   // byte strides along K / along M-or-N
            if (!mn_major) {
                a_addr = a0 + ks * 2 * (128 * 16); a_k = 128 * 16; a_mn = 128;
                b_addr = b0 + ks * 2 * (N * 16);   b_k = N * 16;   b_mn = 128;
            } else {
                a_addr = a0 + ks * 16 * 16; a_k = 128; a_mn = K * 16;
                b_addr = b0 + ks * 16 * 16; b_k = 128; b_mn = K * 16;
            }
            const bool swap = (variant & 1) != 0;
            uint64_t ad = swap ? ptx::smem_desc(a_addr, a_mn, a_k) : ptx::smem_desc(a_addr, a_k, a_mn);
            uint64_t bd = swap ? ptx::smem_desc(b_addr, b_mn, b_k) : ptx::smem_desc(b_addr, b_k, b_mn);
            ptx::mma_bf16(tmem_base, ad, bd, idesc, ks > 0);
        }
        ptx::mma_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, (uint32_t)(N < 32 ? 32 : N));
}

// variant 4: CTA pair (cta_group::2), K-major: D[256,N] = A[256,K] . B[N,K]^T.  Each CTA stages its 128 rows of A
// and its N/2 rows of B; the peer tells the leader that its operands are in place with a remote mbarrier arrive;
// the leader issues the M = 256 MMAs and commits to the `done` barrier of both CTAs.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
tc_selftest_pair_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int N, int K, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t ready, done;
    __shared__ uint32_t tmem_base_s;
    __nv_bfloat16* As = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* Bs = As + 128 * K;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int NH = N / 2;
    for (int i = tid; i < 128 * K; i += 128) {
        int m = i / K, k = i % K;
        As[((size_t)(k >> 3) * 128 + m) * 8 + (k & 7)] = __float2bfloat16(A[(size_t)(rank * 128 + m) * K + k]);
    }
    for (int i = tid; i < NH * K; i += 128) {
        int n = i / K, k = i % K;
        Bs[((size_t)(k >> 3) * NH + n) * 8 + (k & 7)] = __float2bfloat16(Bm[(size_t)(rank * NH + n) * K + k]);
    }
    if (tid == 0) { ptx::mbar_init(&ready, 2); ptx::mbar_init(&done, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc2(&tmem_base_s, (uint32_t)(N < 32 ? 32 : N));
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // barriers of both CTAs initialised, operands of both CTAs written
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) ptx::mbar_arrive_cluster(ptx::mapa(&ready, 0));
    if (tid == 0 && rank == 0) {
        ptx::mbar_wait_cluster(&ready, 0);
        ptx::tc_fence_after();
        const uint32_t a0 = ptx::smem_addr(As), b0 = ptx::smem_addr(Bs);
        const uint32_t idesc = ptx::idesc_bf16(256, N, 0, 0);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t ad = ptx::smem_desc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
            uint64_t bd = ptx::smem_desc(b0 + ks * 2 * (NH * 16), NH * 16, 128);
            ptx::mma2_bf16(tmem_base, ad, bd, idesc, ks > 0);
        }
        ptx::mma2_commit(&done);
    }
    ptx::mbar_wait(&done, 0);
    ptx::tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + row) * N + c0 + j] = __uint_as_float(v[j]);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();          // both CTAs are done with TMEM and with each other's shared memory
    if (warp == 0) ptx::tmem_dealloc2(tmem_base, (uint32_t)(N < 32 ? 32 : N));
}

// ---- probe: do tcgen05.mma (accumulating into TMEM) and tcgen05.ld (draining TMEM) overlap on one SM? -------------
// what & 1: one thread issues `iters` x 16 MMAs (M = 128, N = 256, K = 16: one 256-deep layer of the MLP per iteration)
//           into TMEM columns [0, 256);
// what & 2: warps 4-7 read TMEM columns [256, 512) `iters` times (128 lanes x 256 columns x 4 B = 128 KB per iteration:
//           one layer's accumulator, as the MLP epilogue does).
// out[0] = cycles of the MMA stream, out[1] = cycles of the slowest reading warp (clock64 on the SM).
__global__ void __launch_bounds__(256)
tc_probe_kernel(int what, int iters, long long* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (128 + 256) * 16 * 2 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); out[0] = 0; out[1] = 0; }
    if (warp == 0) ptx::tmem_alloc(&tmem_base_s, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if ((what & 1) && tid == 32) {
        const uint32_t a0 = ptx::smem_addr(smem), b0 = a0 + 128 * 16 * 2;
        const uint32_t idesc = ptx::idesc_bf16(128, 256, 0, 0);
        const uint64_t ad = ptx::smem_desc(a0, 128 * 16, 128), bd = ptx::smem_desc(b0, 256 * 16, 128);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int ks = 0; ks < 16; ++ks) ptx::mma_bf16(tmem_base, ad, bd, idesc, (it | ks) != 0);
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        out[0] = clock64() - t0;
    }
    if ((what & 2) && warp >= 4) {
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        uint32_t v[32], acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll 1
            for (int c = 0; c < 256; c += 32) {
                ptx::tmem_ld32(tacc + c, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v[j];
            }
        }
        const long long dt = clock64() - t0;
        if (acc == 0x12345678u) out[1] = -1;            // keep the loads alive
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(out + 1), (unsigned long long)dt);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace
}  // namespace niw

extern "C" int niw_tc_selftest(const float* A, const float* Bm, int N, int K, int variant, float* D, void* stream) {
    NIW_CHECK_ARG(A && Bm && D);
    if (!(N == 32 || N == 64 || N == 128 || N == 256) || K % 16 != 0 || K <= 0 || K > 256 || variant < 0 || variant > 4)
        return NIW_E_UNSUPP;
    if (variant == 4) {   // CTA pair: A has 256 rows, D is [256, N]
        size_t smem2 = (size_t)(128 + N / 2) * K * 2;
        NIW_CUDA(cudaFuncSetAttribute(niw::tc_selftest_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        niw::note_launch(), niw::tc_selftest_pair_kernel<<<2, 128, smem2, niw_stream(stream)>>>(A, Bm, N, K, D);
        NIW_LAUNCH_CHECK();
        return 0;
    }
    size_t smem = (size_t)(128 + N) * K * 2;
    NIW_CUDA(cudaFuncSetAttribute(niw::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    niw::note_launch(), niw::tc_selftest_kernel<<<1, 128, smem, niw_stream(stream)>>>(A, Bm, N, K, variant, D);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_tc_probe(int what, int iters, long long* out, void* stream) {
    NIW_CHECK_ARG(out && iters > 0 && what >= 1 && what <= 3);
    niw::note_launch(), niw::tc_probe_kernel<<<1, 256, (128 + 256) * 16 * 2, niw_stream(stream)>>>(what, iters, out);
    NIW_LAUNCH_CHECK();
    return 0;
}
