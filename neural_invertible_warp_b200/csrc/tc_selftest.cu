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

}  // namespace
}  // namespace niw

extern "C" int niw_tc_selftest(const float* A, const float* Bm, int N, int K, int variant, float* D, void* stream) {
    NIW_CHECK_ARG(A && Bm && D);
    if (!(N == 32 || N == 64 || N == 128 || N == 256) || K % 16 != 0 || K <= 0 || K > 256 || variant < 0 || variant > 3)
        return NIW_E_UNSUPP;
    size_t smem = (size_t)(128 + N) * K * 2;
    NIW_CUDA(cudaFuncSetAttribute(niw::tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    niw::note_launch(), niw::tc_selftest_kernel<<<1, 128, smem, niw_stream(stream)>>>(A, Bm, N, K, variant, D);
    NIW_LAUNCH_CHECK();
    return 0;
}
