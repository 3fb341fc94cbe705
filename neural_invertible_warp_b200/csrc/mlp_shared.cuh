// Declarations shared by the fp32 (mlp_fp32.cu) and tcgen05 (mlp_tc.cu) MLP paths and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nerf_layout.cuh"

namespace niw {

struct Bands3 { float w[L3]; };   // coarse-to-fine band weights, point encoding  (barf.py:260-264)
struct BandsV { float w[LV]; };   // same, view encoding

// BARF coarse-to-fine schedule: `progress` is a DEVICE scalar (the reference keeps it as a Parameter,
// model/barf.py:254), so the band weights are evaluated on the device and no host read is needed.
// progress == nullptr: no annealing (all weights 1).
struct C2F { const float* progress; float start, end; };
constexpr int NBANDS = 16;        // [w3(10), wv(4), pad(2)] as stored in the workspaces

// w_k = (1 - cos(pi * clamp(alpha - k, 0, 1))) / 2,  alpha = (progress - start) / (end - start) * L
__device__ __forceinline__ float band_weight(const C2F& c, int k, int L) {
    if (c.progress == nullptr) return 1.f;
    float alpha = (c.progress[0] - c.start) / (c.end - c.start) * (float)L;
    float t = fminf(fmaxf(alpha - (float)k, 0.f), 1.f);
    return (1.f - cosf(t * PI_F)) * 0.5f;
}
__device__ __forceinline__ void store_bands(const C2F& c, int i, float* out) {   // i in [0, NBANDS)
    out[i] = i < L3 ? band_weight(c, i, L3) : (i < L3 + LV ? band_weight(c, i - L3, LV) : 0.f);
}
__device__ __forceinline__ void load_bands(const float* __restrict__ bands, Bands3& b3, BandsV& bv) {
#pragma unroll
    for (int k = 0; k < L3; ++k) b3.w[k] = bands[k];
#pragma unroll
    for (int k = 0; k < LV; ++k) bv.w[k] = bands[L3 + k];
}

inline int64_t fp32_eval_chunk_rays(int N) { int64_t r = (int64_t(1) << 20) / N; return r < 1 ? 1 : r; }

size_t fp32_workspace_bytes(int64_t R, int N, int training);
int fp32_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             const C2F& c2f, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma,
             cudaStream_t st);
int fp32_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
             float* dP, float* d_center, float* d_ray, cudaStream_t st);

size_t tc_workspace_bytes(int64_t R, int N, int training);
int tc_pack(const float* P, const C2F& c2f, int training, int64_t R, int N, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
           const C2F& c2f, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma,
           cudaStream_t st);
int tc_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
           void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
           float* dP, float* d_center, float* d_ray, cudaStream_t st);

// the two halves of tc_bwd, for callers that run the HBM-bound weight-gradient pass on another stream
int tc_bwd_dx(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
              void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
              float* dP, float* d_center, float* d_ray, cudaStream_t st);
int tc_bwd_dw(int64_t R, int N, void* ws, size_t ws_bytes, float* dP, int max_ctas, cudaStream_t st);

// split-precision (hi + lo BF16 operands, 3 MMAs per product) tcgen05 forward, mlp_tc_x3.cu; its backward is tc_bwd
size_t tc_x3_workspace_bytes(int64_t R, int N, int training);
int tc_x3_pack(const float* P, const C2F& c2f, int training, int64_t R, int N, void* ws, size_t ws_bytes, cudaStream_t st);
int tc_x3_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
              const C2F& c2f, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma, cudaStream_t st);

// kernels defined in mlp_fp32.cu that the tcgen05 path reuses
__global__ void encode_bwd_kernel(const float* __restrict__ center, const float* __restrict__ ray,
                                  const float* __restrict__ depth, int64_t R, int N, const float* __restrict__ bands,
                                  const float* __restrict__ d_enc, int ld_enc, const float* __restrict__ d_venc_s,
                                  int ld_venc, float* __restrict__ d_center, float* __restrict__ d_ray);

}  // namespace niw
