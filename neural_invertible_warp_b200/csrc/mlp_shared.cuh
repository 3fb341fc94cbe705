// Declarations shared by the fp32 (mlp_fp32.cu) and tcgen05 (mlp_tc.cu) MLP paths and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nerf_layout.cuh"

namespace niw {

struct Bands3 { float w[L3]; };   // coarse-to-fine band weights, point encoding  (barf.py:260-264)
struct BandsV { float w[LV]; };   // same, view encoding

inline int64_t fp32_eval_chunk_rays(int N) { int64_t r = (int64_t(1) << 20) / N; return r < 1 ? 1 : r; }

size_t fp32_workspace_bytes(int64_t R, int N, int training);
int fp32_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             const Bands3& b3, const BandsV& bv, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma,
             cudaStream_t st);
int fp32_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             const Bands3& b3, const BandsV& bv, void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
             float* dP, float* d_center, float* d_ray, cudaStream_t st);

size_t tc_workspace_bytes(int64_t R, int N, int training);
int tc_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
           const Bands3& b3, const BandsV& bv, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma,
           cudaStream_t st);
int tc_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
           const Bands3& b3, const BandsV& bv, void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
           float* dP, float* d_center, float* d_ray, cudaStream_t st);

// kernels defined in mlp_fp32.cu that the tcgen05 path reuses
__global__ void encode_bwd_kernel(const float* __restrict__ center, const float* __restrict__ ray,
                                  const float* __restrict__ depth, int64_t R, int N, Bands3 bw3, BandsV bwv,
                                  const float* __restrict__ d_enc, int ld_enc, const float* __restrict__ d_venc_s,
                                  int ld_venc, float* __restrict__ d_center, float* __restrict__ d_ray);

}  // namespace niw
