// Evaluation metrics on the device: the consumers of the eval render (C4).
//   PSNR   -10 log10(MSE(rgb_map, image))                                  reference model/nerf.py:179
//   SSIM   11x11 Gaussian window (sigma 1.5), zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over the map
//          reference external/pohsun_ssim/pytorch_ssim/__init__.py:7-37 (five grouped conv2d + elementwise)
//   depth  masked mean |gt - pred| and RMSE, with and without a scale on the prediction
//          reference core/metrics.py:64-111
// The rendered image is read in the renderer's own layout ([B, H*W, 3], pixel-major) -- the reference permutes it to
// [B,3,H,W] first -- and the window is applied separably in shared memory: one pass over both images, 24 B/pixel read.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int WIN = 11, HALO = WIN / 2, TS = 16, IN = TS + 2 * HALO;   // 16x16 outputs from a 26x26 input tile

struct Window { float w[WIN]; };

__global__ void __launch_bounds__(TS * TS)
image_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ image, int H, int W, Window win,
                     float* __restrict__ out) {
    __shared__ float sx[IN][IN + 1], sy[IN][IN + 1];
    __shared__ float hs[5][IN][TS];
    __shared__ float red[2][TS * TS / 32];
    const int b = blockIdx.z / 3, c = blockIdx.z % 3;
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS;
    const int tid = threadIdx.y * TS + threadIdx.x;
    const int64_t HW = (int64_t)H * W;
    for (int i = tid; i < IN * IN; i += TS * TS) {
        const int ly = i / IN, lx = i % IN, gy = y0 + ly - HALO, gx = x0 + lx - HALO;
        const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;                    // conv2d(padding = 5): zeros outside
        sx[ly][lx] = in ? pred[((int64_t)b * HW + (int64_t)gy * W + gx) * 3 + c] : 0.f;
        sy[ly][lx] = in ? image[((int64_t)(b * 3 + c) * H + gy) * W + gx] : 0.f;
    }
    __syncthreads();
    for (int i = tid; i < IN * TS; i += TS * TS) {                                   // horizontal pass
        const int ly = i / TS, lx = i % TS;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
            const float x = sx[ly][lx + k], y = sy[ly][lx + k], w = win.w[k];
            a0 += w * x; a1 += w * y; a2 += w * (x * x); a3 += w * (y * y); a4 += w * (x * y);
        }
        hs[0][ly][lx] = a0; hs[1][ly][lx] = a1; hs[2][ly][lx] = a2; hs[3][ly][lx] = a3; hs[4][ly][lx] = a4;
    }
    __syncthreads();
    float ssim = 0.f, se = 0.f;
    {
        const int lx = threadIdx.x, ly = threadIdx.y;
        if (y0 + ly < H && x0 + lx < W) {
            float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
            for (int k = 0; k < WIN; ++k) {                                          // vertical pass
                const float w = win.w[k];
                m1 += w * hs[0][ly + k][lx]; m2 += w * hs[1][ly + k][lx];
                s11 += w * hs[2][ly + k][lx]; s22 += w * hs[3][ly + k][lx]; s12 += w * hs[4][ly + k][lx];
            }
            const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            const float v1 = s11 - m11, v2 = s22 - m22, v12 = s12 - m12;
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            ssim = ((2.f * m12 + C1) * (2.f * v12 + C2)) / ((m11 + m22 + C1) * (v1 + v2 + C2));
            const float d = sx[ly + HALO][lx + HALO] - sy[ly + HALO][lx + HALO];
            se = d * d;
        }
    }
    ssim = warp_sum(ssim); se = warp_sum(se);
    if ((tid & 31) == 0) { red[0][tid >> 5] = se; red[1][tid >> 5] = ssim; }
    __syncthreads();
    if (tid < 2) {
        float v = 0.f;
        for (int i = 0; i < TS * TS / 32; ++i) v += red[tid][i];
        atomicAdd(out + b * 2 + tid, v);
    }
}

// out[0] = #valid, out[1] = sum |gt - p|, out[2] = sum (gt - p)^2, out[3] / out[4] = the same for scale * p
__global__ void depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                     const uint8_t* __restrict__ valid, int64_t n, float scale, float* __restrict__ out) {
    float a[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        const float g = gt[i], p = pred[i], d = g - p, ds = g - p * scale;
        a[0] += 1.f; a[1] += fabsf(d); a[2] += d * d; a[3] += fabsf(ds); a[4] += ds * ds;
    }
    __shared__ float red[5][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float v = warp_sum(a[k]);
        if (lane == 0) red[k][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, v);
    }
}

}  // namespace

extern "C" int niw_image_metrics(const float* pred_rgb, const float* image, int B, int H, int W, float* out, void* stream) {
    NIW_CHECK_ARG(pred_rgb && image && out && B > 0 && H > 0 && W > 0);
    cudaStream_t st = niw_stream(stream);
    NIW_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 2 * B, st));
    // the reference's window: gaussian(11, 1.5) evaluated in Python doubles, stored and normalised in fp32
    Window win;
    float g[WIN], sum = 0.f;
    for (int i = 0; i < WIN; ++i) { g[i] = (float)exp(-(double)((i - HALO) * (i - HALO)) / (2.0 * 1.5 * 1.5)); sum += g[i]; }
    for (int i = 0; i < WIN; ++i) win.w[i] = g[i] / sum;
    dim3 grid((W + TS - 1) / TS, (H + TS - 1) / TS, 3 * B), block(TS, TS);
    niw::note_launch(), image_metrics_kernel<<<grid, block, 0, st>>>(pred_rgb, image, H, W, win, out);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_depth_metrics(const float* pred, const float* gt, const uint8_t* valid, int64_t n, float scale,
                                 float* out, void* stream) {
    NIW_CHECK_ARG(pred && gt && out && n > 0);
    cudaStream_t st = niw_stream(stream);
    NIW_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 5, st));
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)niw_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    niw::note_launch(), depth_metrics_kernel<<<(unsigned)blocks, 256, 0, st>>>(pred, gt, valid, n, scale, out);
    NIW_LAUNCH_CHECK();
    return 0;
}
