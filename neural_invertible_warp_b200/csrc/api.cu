// C-ABI glue: dispatch of the NeRF MLP entry points by precision, the loss head, error strings.
#include "common.cuh"
#include "mlp_shared.cuh"
#include <string.h>

using namespace niw;

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
namespace niw { void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); } }

extern "C" int niw_abi_version(void) { return NIW_ABI_VERSION; }
extern "C" unsigned long long niw_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* niw_error_string(int code) {
    if (code == 0) return "success";
    if (code == NIW_E_BADARG) return "niw: bad argument (null pointer or non-positive size)";
    if (code == NIW_E_UNSUPP) return "niw: unsupported shape or option";
    if (code == NIW_E_WORKSPACE) return "niw: workspace too small";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "niw: unknown error";
}

extern "C" size_t niw_nerf_workspace_bytes(int64_t R, int N, int precision, int training) {
    if (R <= 0 || N <= 0) return 0;
    training &= 1;
    if (precision == NIW_PREC_BF16) return tc_workspace_bytes(R, N, training);
    if (precision == NIW_PREC_BF16X3) return tc_x3_workspace_bytes(R, N, training);
    return precision == NIW_PREC_FP32 ? fp32_workspace_bytes(R, N, training) : 0;
}

// `progress` is the DEVICE scalar of the BARF schedule (model/barf.py:254); the band weights of
// model/barf.py:260-264 are evaluated on the device and kept in the workspace for the backward pass.
extern "C" int niw_nerf_fwd(const float* params, const float* center, const float* ray, const float* depth, int64_t R,
                            int N, const float* progress, float c2f_start, float c2f_end, int precision, int training,
                            void* workspace, size_t workspace_bytes, float* rgb, float* sigma, void* stream) {
    NIW_CHECK_ARG(params && center && ray && depth && workspace && rgb && sigma && R > 0 && N > 0);
    if (progress && !(c2f_end != c2f_start)) return NIW_E_BADARG;
    C2F c2f{progress, c2f_start, c2f_end};
    if (precision == NIW_PREC_FP32)
        return fp32_fwd(params, center, ray, depth, R, N, c2f, training & 1, workspace, workspace_bytes, rgb, sigma,
                        niw_stream(stream));
    if (precision == NIW_PREC_BF16)
        return tc_fwd(params, center, ray, depth, R, N, c2f, training, workspace, workspace_bytes, rgb, sigma,
                      niw_stream(stream));
    if (precision == NIW_PREC_BF16X3)
        return tc_x3_fwd(params, center, ray, depth, R, N, c2f, training, workspace, workspace_bytes, rgb, sigma,
                         niw_stream(stream));
    return NIW_E_UNSUPP;
}

extern "C" int niw_nerf_pack(const float* params, const float* progress, float c2f_start, float c2f_end, int precision,
                             int training, int64_t R, int N, void* workspace, size_t workspace_bytes, void* stream) {
    NIW_CHECK_ARG(params && workspace && R > 0 && N > 0);
    if (progress && !(c2f_end != c2f_start)) return NIW_E_BADARG;
    C2F c2f{progress, c2f_start, c2f_end};
    if (precision == NIW_PREC_BF16X3)
        return tc_x3_pack(params, c2f, training & 1, R, N, workspace, workspace_bytes, niw_stream(stream));
    if (precision != NIW_PREC_BF16) return NIW_E_UNSUPP;      // the FP32 path reads the parameters in place
    return tc_pack(params, c2f, training & 1, R, N, workspace, workspace_bytes, niw_stream(stream));
}

extern "C" int niw_nerf_bwd(const float* params, const float* center, const float* ray, const float* depth, int64_t R,
                            int N, int precision, void* workspace,
                            size_t workspace_bytes, const float* d_rgb, const float* d_sigma, float* d_params,
                            float* d_center, float* d_ray, void* stream) {
    NIW_CHECK_ARG(params && center && ray && depth && workspace && d_rgb && d_sigma && d_params &&
                  d_center && d_ray && R > 0 && N > 0);
    if (precision == NIW_PREC_FP32)
        return fp32_bwd(params, center, ray, depth, R, N, workspace, workspace_bytes, d_rgb, d_sigma, d_params,
                        d_center, d_ray, niw_stream(stream));
    if (precision == NIW_PREC_BF16 || precision == NIW_PREC_BF16X3)     // the split-precision forward saves BF16 tile records too
        return tc_bwd(params, center, ray, depth, R, N, workspace, workspace_bytes, d_rgb, d_sigma, d_params,
                      d_center, d_ray, niw_stream(stream));
    return NIW_E_UNSUPP;
}

// niw_nerf_bwd in two calls (tensor-core precisions only): the activation-gradient chain, then the weight-gradient pass,
// which may be enqueued on a different stream once the first has been (the caller orders the streams)
extern "C" int niw_nerf_bwd_dx(const float* params, const float* center, const float* ray, const float* depth, int64_t R,
                               int N, int precision, void* workspace, size_t workspace_bytes, const float* d_rgb,
                               const float* d_sigma, float* d_params, float* d_center, float* d_ray, void* stream) {
    NIW_CHECK_ARG(params && center && ray && depth && workspace && d_rgb && d_sigma && d_params &&
                  d_center && d_ray && R > 0 && N > 0);
    if (precision != NIW_PREC_BF16 && precision != NIW_PREC_BF16X3) return NIW_E_UNSUPP;
    return tc_bwd_dx(params, center, ray, depth, R, N, workspace, workspace_bytes, d_rgb, d_sigma, d_params, d_center,
                     d_ray, niw_stream(stream));
}

extern "C" int niw_nerf_bwd_dw(int64_t R, int N, int precision, void* workspace, size_t workspace_bytes, float* d_params,
                               int max_ctas, void* stream) {
    NIW_CHECK_ARG(workspace && d_params && R > 0 && N > 0 && max_ctas >= 0);
    if (precision != NIW_PREC_BF16 && precision != NIW_PREC_BF16X3) return NIW_E_UNSUPP;
    return tc_bwd_dw(R, N, workspace, workspace_bytes, d_params, max_ctas, niw_stream(stream));
}

// ---- loss head: pixel gather + squared error  (model/nerf.py:276-288, model/base.py:209-211) ----
namespace {
__global__ void mse_gather_kernel(const float* __restrict__ image, const float* __restrict__ rgb,
                                  const int64_t* __restrict__ ray_idx, int64_t idx_start, int B, int P, int HW,
                                  float scale, float* __restrict__ loss, float* __restrict__ d_rgb, int direct) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    if (t < (int64_t)B * P) {
        int b = (int)(t / P), p = (int)(t % P);
        int64_t pix = ray_idx ? ray_idx[p] : idx_start + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float diff = rgb[t * 3 + c] - image[((int64_t)b * 3 + c) * HW + pix];
            acc += diff * diff;
            if (d_rgb) d_rgb[t * 3 + c] = 2.f * scale * diff;
        }
    }
    acc = warp_sum(acc);
    __shared__ float red[32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w];
        if (direct) *loss = v * scale;          // the only block: no accumulation, `loss` need not be zeroed
        else atomicAdd(loss, v * scale);
    }
}
}  // namespace

extern "C" int niw_mse_gather_needs_zero(int B, int P) { return (int64_t)B * P > 1024 ? 1 : 0; }

extern "C" int niw_mse_gather(const float* image, const float* rgb, const int64_t* ray_idx, int64_t idx_start, int B,
                              int P, int H, int W, float scale, float* loss, float* d_rgb, void* stream) {
    NIW_CHECK_ARG(image && rgb && loss && B > 0 && P > 0 && H > 0 && W > 0);
    int64_t n = (int64_t)B * P;
    // up to 1 024 rays: ONE block that writes the loss (no zero-fill launch in front, no atomics); more: `loss` accumulates
    // over the blocks and the caller zeroes it first (niw_mse_gather_needs_zero)
    if (n <= 1024)
        niw::note_launch(), mse_gather_kernel<<<1, (unsigned)((n + 31) / 32 * 32), 0, niw_stream(stream)>>>(image, rgb, ray_idx, idx_start, B, P,
                                                                                       H * W, scale, loss, d_rgb, 1);
    else
        niw::note_launch(), mse_gather_kernel<<<niw_blocks(n, 256), 256, 0, niw_stream(stream)>>>(image, rgb, ray_idx, idx_start, B, P, H * W,
                                                                             scale, loss, d_rgb, 0);
    NIW_LAUNCH_CHECK();
    return 0;
}
