// Depth sampling.
//   a4  Graph.sample_depth                      reference model/nerf.py:334-344
//   a5  Graph.sample_depth_from_pdf + cat+sort  reference model/nerf.py:346-365, :313-315
// Both are HBM-bound streaming kernels.  Arithmetic uses explicit round-to-nearest intrinsics so
// that no multiply-add is contracted: the depths and, for a5, the bin indices are bit-exact with
// the reference's fp32 CPU evaluation.
#include "common.cuh"
#include <math_constants.h>

namespace {

// ((u+k)/N)*scale + dmin ; inverse: 1/(d+1e-8).  One thread per 4 consecutive samples.
__global__ void stratified_kernel(const float* __restrict__ u, int64_t total, int N, float scale, float dmin,
                                  int inverse, float* __restrict__ depth) {
    int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= total) return;
    float uu[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    bool vec = (i0 + 3 < total);
    if (u) {
        if (vec) {
            float4 v = *reinterpret_cast<const float4*>(u + i0);
            uu[0] = v.x; uu[1] = v.y; uu[2] = v.z; uu[3] = v.w;
        } else {
            for (int j = 0; j < 4 && i0 + j < total; ++j) uu[j] = u[i0 + j];
        }
    }
    float out[4];
    const float fN = (float)N;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float k = (float)((i0 + j) % N);
        float d = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(uu[j], k), fN), scale), dmin);
        if (inverse) d = __fdiv_rn(1.0f, __fadd_rn(d, 1e-8f));
        out[j] = d;
    }
    if (vec) {
        *reinterpret_cast<float4*>(depth + i0) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
        for (int j = 0; j < 4 && i0 + j < total; ++j) depth[i0 + j] = out[j];
    }
}

// One warp per ray.  Shared memory per warp: cdf[N+1] | sort buffer[npow2].
template <int WARPS>
__global__ void pdf_merge_kernel(const float* __restrict__ pdf, const float* __restrict__ depth_coarse,
                                 const float* __restrict__ unif, const float* __restrict__ bins, int64_t R, int N,
                                 int Nf, int npow2, float* __restrict__ fine, int64_t* __restrict__ idx_out,
                                 float* __restrict__ merged) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = (N + 1) + npow2;
    float* cdf = smem + (size_t)warp * per_warp;
    float* buf = cdf + (N + 1);
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        const float* p = pdf + r * N;
        for (int i = lane; i < N; i += 32) cdf[i + 1] = p[i];
        __syncwarp();
        if (lane == 0) {
            // torch.cumsum on CPU: sequential fp64 accumulation, each output rounded to fp32
            double acc = 0.0;
            cdf[0] = 0.f;
            for (int i = 1; i <= N; ++i) { acc += (double)cdf[i]; cdf[i] = (float)acc; }
        }
        __syncwarp();
        for (int j = lane; j < Nf; j += 32) {
            float uj = unif[j];
            // searchsorted(right=True): first index with cdf[index] > u, in [0, N+1]
            int lo = 0, hi = N + 1;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (cdf[mid] <= uj) lo = mid + 1; else hi = mid;
            }
            int il = lo - 1 < 0 ? 0 : lo - 1;
            int ih = lo > N ? N : lo;
            float cl = cdf[il], ch = cdf[ih], dl = bins[il], dh = bins[ih];
            float t = __fdiv_rn(__fsub_rn(uj, cl), __fadd_rn(__fsub_rn(ch, cl), 1e-8f));
            float d = __fadd_rn(dl, __fmul_rn(t, __fsub_rn(dh, dl)));
            if (fine) fine[r * Nf + j] = d;
            if (idx_out) idx_out[r * Nf + j] = (int64_t)lo;
            buf[N + j] = d;
        }
        if (merged) {
            const float* dc = depth_coarse + r * N;
            for (int i = lane; i < N; i += 32) buf[i] = dc[i];
            for (int i = N + Nf + lane; i < npow2; i += 32) buf[i] = CUDART_INF_F;
            __syncwarp();
            // bitonic sort of npow2 keys by one warp (values only; equals torch.sort(...).values)
            for (int k = 2; k <= npow2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = lane; i < npow2; i += 32) {
                        int ixj = i ^ j;
                        if (ixj > i) {
                            float a = buf[i], b = buf[ixj];
                            bool up = ((i & k) == 0);
                            if ((a > b) == up) { buf[i] = b; buf[ixj] = a; }
                        }
                    }
                    __syncwarp();
                }
            }
            float* m = merged + r * (N + Nf);
            for (int i = lane; i < N + Nf; i += 32) m[i] = buf[i];
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------
// Random pixel subset: the first k entries of a random permutation of [0, n) -- what the reference draws with
// torch.randperm(H*W)[:rand_rays//B] (model/nerf.py:268), which sorts H*W random keys (5 radix-sort passes,
// ~70 us at 480x640) to keep 64 of them.  Here entry i is pi(i) for a keyed bijection pi of [0, 2^b) (4-round
// Feistel network, b = bit length of n-1 rounded up to even) with cycle walking back into [0, n): a prefix of a
// permutation by construction, O(k) work, no global state besides a call counter (so CUDA-graph replays draw new
// subsets).  Different random stream than torch's generator: used where the draw itself is the only requirement.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // lowbias32
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__global__ void sample_pixels_kernel(int64_t n, int k, uint64_t seed, unsigned long long* __restrict__ counter,
                                     int64_t* __restrict__ out) {
    const unsigned long long call = *counter;
    int bits = 2;
    while ((1ull << bits) < (unsigned long long)n) bits += 2;
    const int hb = bits / 2;
    const uint32_t hmask = (1u << hb) - 1u;
    uint32_t key[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) key[r] = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u * (uint32_t)(4 * call + r + 1)) ^ (uint32_t)(call >> 32));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
        uint64_t x = (uint64_t)i;
        do {                                             // cycle walking: expected < 4 trips (2^b < 4n)
            uint32_t L = (uint32_t)(x >> hb) & hmask, R = (uint32_t)x & hmask;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const uint32_t t = L ^ (mix32(R ^ key[r]) & hmask); L = R; R = t; }
            x = ((uint64_t)L << hb) | R;
        } while (x >= (uint64_t)n);
        out[i] = (int64_t)x;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = call + 1;      // single block (k <= a few thousand)
}

}  // namespace

extern "C" int niw_sample_stratified(const float* u, int64_t n_rays, int N, float scale, float dmin, int inverse,
                                     float* depth, void* stream) {
    NIW_CHECK_ARG(depth && n_rays > 0 && N > 0);
    int64_t total = n_rays * N;
    niw::note_launch(), stratified_kernel<<<niw_blocks((total + 3) / 4, 256), 256, 0, niw_stream(stream)>>>(u, total, N, scale, dmin,
                                                                                       inverse, depth);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_sample_pdf_merge(const float* pdf, const float* depth_coarse, const float* unif, const float* bins,
                                    int64_t R, int N, int Nf, float* fine, int64_t* idx, float* merged,
                                    void* stream) {
    NIW_CHECK_ARG(pdf && unif && bins && R > 0 && N > 0 && Nf > 0 && (!merged || depth_coarse));
    int npow2 = 1;
    while (npow2 < N + Nf) npow2 <<= 1;
    if (npow2 > 4096) return NIW_E_UNSUPP;
    constexpr int WARPS = 4;
    size_t smem = sizeof(float) * WARPS * ((size_t)(N + 1) + npow2);
    if (smem > 200 * 1024) return NIW_E_UNSUPP;
    if (smem > 48 * 1024)
        NIW_CUDA(cudaFuncSetAttribute(pdf_merge_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (R + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)niw_num_sms() * 16;
    if (blocks > cap) blocks = cap;
    niw::note_launch(), pdf_merge_kernel<WARPS><<<(unsigned)blocks, WARPS * 32, smem, niw_stream(stream)>>>(
        pdf, depth_coarse, unif, bins, R, N, Nf, npow2, fine, idx, merged);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_sample_pixels(int64_t n, int k, uint64_t seed, uint64_t* counter, int64_t* out, void* stream) {
    NIW_CHECK_ARG(counter && out && n > 0 && k > 0 && k <= n && n <= (int64_t(1) << 40));
    // one block: the counter is read by every thread before the last one advances it
    niw::note_launch(), sample_pixels_kernel<<<1, 1024, 0, niw_stream(stream)>>>(n, k, seed, reinterpret_cast<unsigned long long*>(counter), out);
    NIW_LAUNCH_CHECK();
    return 0;
}
