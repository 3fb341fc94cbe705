// Ray generation.
//   a1  camera.get_center_and_ray + sub-selection   reference camera.py:419-443, model/nerf.py:298-300
//   a2  camera.get_unwarped_center_and_ray          reference camera.py:359-390
// The reference materialises the full B x HW grid and then indexes it; here only the requested
// pixels are ever generated (one thread per ray), and the pose gradient is reduced per image.
#include "common.cuh"

namespace {

// world-frame quantities of one pose: Rinv = R^T, tinv = -R^T t   (camera.py:89-95)
struct PoseInv { float R[9]; float t[3]; float tinv[3]; };

__device__ __forceinline__ PoseInv load_pose(const float* __restrict__ p) {
    PoseInv q;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) q.R[i * 3 + j] = p[i * 4 + j];
        q.t[i] = p[i * 4 + 3];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
        q.tinv[j] = -(q.R[0 * 3 + j] * q.t[0] + q.R[1 * 3 + j] * q.t[1] + q.R[2 * 3 + j] * q.t[2]);
    return q;
}

__device__ __forceinline__ void cam_to_world(const PoseInv& q, const float g[3], float out[3]) {
    // X_hom @ pose_inv^T  (camera.py:343-346)
#pragma unroll
    for (int j = 0; j < 3; ++j)
        out[j] = q.R[0 * 3 + j] * g[0] + q.R[1 * 3 + j] * g[1] + q.R[2 * 3 + j] * g[2] + q.tinv[j];
}

// grid (chunks of pixels, B): the image's K^-1 and inverted pose are computed once per block; a warp produces 128
// consecutive rays per iteration (4 per lane), stages the 384 floats of each output in shared memory and writes them
// back as 128-bit stores that are contiguous across the warp (24 B/ray written, nothing read but ray_idx)
constexpr int RG_WARPS = 8;
__global__ void __launch_bounds__(RG_WARPS * 32)
raygen_pose_fwd_kernel(const float* __restrict__ pose, const float* __restrict__ intr,
                       const int64_t* __restrict__ ray_idx, int64_t idx_start, int P, int W, int vec,
                       float* __restrict__ center, float* __restrict__ ray) {
    __shared__ Mat3 s_Ki;
    __shared__ PoseInv s_q;
    __shared__ __align__(16) float s_stage[RG_WARPS][384];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_Ki = inverse3x3(intr + b * 9); s_q = load_pose(pose + b * 12); }
    __syncthreads();
    const Mat3 Ki = s_Ki;
    const PoseInv q = s_q;
    float* stage = s_stage[warp];
    for (int w0 = (blockIdx.x * RG_WARPS + warp) * 128; w0 < P; w0 += gridDim.x * RG_WARPS * 128) {
        const int p0 = w0 + lane * 4;
        float orr[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = p0 + j < P ? p0 + j : P - 1;
            const int64_t pix = ray_idx ? ray_idx[p] : idx_start + p;
            float g[3], gw[3];
            pixel_to_cam(Ki, pix, W, g);
            cam_to_world(q, g, gw);
#pragma unroll
            for (int c = 0; c < 3; ++c) orr[j * 3 + c] = gw[c] - q.tinv[c];   // grid_3D - center_3D  (camera.py:442)
        }
        const int64_t t0 = (int64_t)b * P + w0;                  // first ray of the warp
        if (vec && w0 + 128 <= P) {
            __syncwarp();
            vstore<12>(stage + lane * 12, orr);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int f = (i * 32 + lane) * 4;               // float offset inside the warp's 384-float block
                *reinterpret_cast<float4*>(ray + t0 * 3 + f) = *reinterpret_cast<const float4*>(stage + f);
                // the centre repeats with period 3 floats: float f holds component f % 3
                const int m = f % 3;
                *reinterpret_cast<float4*>(center + t0 * 3 + f) =
                    make_float4(q.tinv[m], q.tinv[(m + 1) % 3], q.tinv[(m + 2) % 3], q.tinv[m]);
            }
        } else {
            for (int j = 0; j < 4 && p0 + j < P; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) { center[(t0 + lane * 4 + j) * 3 + c] = q.tinv[c]; ray[(t0 + lane * 4 + j) * 3 + c] = orr[j * 3 + c]; }
        }
    }
}

// d_pose[b] = sum over rays.  ray_j = sum_i R[i][j] g_i, center_j = -sum_i R[i][j] t_i, so
//   dR[i][j] = sum g_i d_ray_j - t_i * sum d_center_j ,  dt_i = -sum_j R[i][j] sum d_center_j.
__global__ void raygen_pose_bwd_kernel(const float* __restrict__ pose, const float* __restrict__ intr,
                                       const int64_t* __restrict__ ray_idx, int64_t idx_start, int B, int P, int W,
                                       const float* __restrict__ d_center, const float* __restrict__ d_ray,
                                       float* __restrict__ d_pose) {
    int b = blockIdx.y;
    Mat3 Ki = inverse3x3(intr + b * 9);
    float acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0.f;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        int64_t t = (int64_t)b * P + p;
        int64_t pix = ray_idx ? ray_idx[p] : idx_start + p;
        float g[3];
        pixel_to_cam(Ki, pix, W, g);
        if (d_ray) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[i * 3 + j] += g[i] * d_ray[t * 3 + j];
        }
        if (d_center) {
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[9 + j] += d_center[t * 3 + j];
        }
    }
    __shared__ float red[12][8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        float v = warp_sum(acc[i]);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        red[threadIdx.x][0] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const float* pb = pose + b * 12;
        int i = threadIdx.x / 4, j = threadIdx.x % 4;
        float out;
        if (j < 3) out = red[i * 3 + j][0] - pb[i * 4 + 3] * red[9 + j][0];
        else out = -(pb[i * 4 + 0] * red[9][0] + pb[i * 4 + 1] * red[10][0] + pb[i * 4 + 2] * red[11][0]);
        atomicAdd(d_pose + b * 12 + threadIdx.x, out);
    }
}

// pts [B, P + n_center, 3] = [P grid rows ; n_center centre rows], n_center = P (the reference's list) or 1 (all centre
// rows of an image are the same point; the warp then evaluates it once)
__global__ void raygen_unwarped_kernel(const float* __restrict__ intr, const float* __restrict__ pose_init,
                                       const int64_t* __restrict__ ray_idx, int64_t idx_start, int B, int P, int n_center,
                                       int W, float* __restrict__ pts) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)B * P) return;
    int b = (int)(t / P), p = (int)(t % P);
    int64_t pix = ray_idx ? ray_idx[p] : idx_start + p;
    Mat3 Ki = inverse3x3(intr + b * 9);
    float g[3], c[3] = {0.f, 0.f, 0.f};
    pixel_to_cam(Ki, pix, W, g);
    if (pose_init) {
        PoseInv q = load_pose(pose_init + b * 12);
        float gw[3];
        cam_to_world(q, g, gw);
#pragma unroll
        for (int j = 0; j < 3; ++j) { g[j] = gw[j]; c[j] = q.tinv[j]; }
    }
    const int64_t rows = (int64_t)P + n_center;
    float* grid_row = pts + ((int64_t)b * rows + p) * 3;
#pragma unroll
    for (int j = 0; j < 3; ++j) grid_row[j] = g[j];
    if (p < n_center) {
        float* cen_row = pts + ((int64_t)b * rows + P + p) * 3;
#pragma unroll
        for (int j = 0; j < 3; ++j) cen_row[j] = c[j];
    }
}

// ---- rays from the warped point list (model/barf_inn_llff.py:352-356, pose_models/inn.py:75-77) ----------------
// warped [B, P + n_center, 3] -> ray = grid - centre, centre (expanded to P rows), both [B,P,3]
__global__ void rays_from_warp_fwd_kernel(const float* __restrict__ warped, int B, int P, int n_center,
                                          float* __restrict__ ray, float* __restrict__ center) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread per output float
    if (i >= (int64_t)B * P * 3) return;
    const int c = (int)(i % 3);
    const int64_t t = i / 3;
    const int b = (int)(t / P), p = (int)(t % P);
    const int64_t rows = (int64_t)P + n_center;
    const float g = warped[((int64_t)b * rows + p) * 3 + c];
    const float ce = warped[((int64_t)b * rows + P + (n_center == 1 ? 0 : p)) * 3 + c];
    ray[i] = g - ce;
    center[i] = ce;
}

// d_warped = [d_ray ; d_centre - d_ray] (n_center = P) or [d_ray ; sum_p (d_centre - d_ray)] (n_center = 1); one block
// per image, either upstream gradient may be NULL
__global__ void __launch_bounds__(256)
rays_from_warp_bwd_kernel(const float* __restrict__ d_ray, const float* __restrict__ d_center, int P, int n_center,
                          float* __restrict__ d_warped) {
    const int b = blockIdx.x;
    const int64_t rows = (int64_t)P + n_center;
    float acc[3] = {0.f, 0.f, 0.f};
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t i = ((int64_t)b * P + p) * 3 + c;
            const float dr = d_ray ? d_ray[i] : 0.f, lower = (d_center ? d_center[i] : 0.f) - dr;
            d_warped[((int64_t)b * rows + p) * 3 + c] = dr;
            if (n_center == 1) acc[c] += lower; else d_warped[((int64_t)b * rows + P + p) * 3 + c] = lower;
        }
    }
    if (n_center == 1) {
        __shared__ float red[3][8];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = warp_sum(acc[c]);
            if (lane == 0) red[c][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            float v = 0.f;
            for (int w = 0; w < 8; ++w) v += red[threadIdx.x][w];
            d_warped[((int64_t)b * rows + P) * 3 + threadIdx.x] = v;
        }
    }
}

}  // namespace

extern "C" int niw_raygen_pose_fwd(const float* pose, const float* intr, const int64_t* ray_idx, int64_t idx_start,
                                   int B, int P, int H, int W, float* center, float* ray, void* stream) {
    NIW_CHECK_ARG(pose && intr && center && ray && B > 0 && P > 0 && H > 0 && W > 0);
    const int threads = RG_WARPS * 32;
    // a block amortises its K^-1 / pose prologue over many pixels: a few waves of blocks in total
    unsigned bx = niw_blocks(P, RG_WARPS * 128);
    const unsigned want = (unsigned)((niw_num_sms() * 32 + B - 1) / B);
    if (bx > want) bx = want;
    dim3 grid(bx, B);
    const int vec = ((int64_t)P * 3 % 4 == 0) && niw_aligned16(center) && niw_aligned16(ray);
    niw::note_launch(), raygen_pose_fwd_kernel<<<grid, threads, 0, niw_stream(stream)>>>(pose, intr, ray_idx, idx_start, P, W, vec, center, ray);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_raygen_pose_bwd(const float* pose, const float* intr, const int64_t* ray_idx, int64_t idx_start,
                                   int B, int P, int H, int W, const float* d_center, const float* d_ray,
                                   float* d_pose, void* stream) {
    NIW_CHECK_ARG(pose && intr && d_pose && (d_center || d_ray) && B > 0 && P > 0 && H > 0 && W > 0);
    NIW_CUDA(cudaMemsetAsync(d_pose, 0, sizeof(float) * 12 * B, niw_stream(stream)));
    int bx = (P + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid(bx, B);
    niw::note_launch(), raygen_pose_bwd_kernel<<<grid, 256, 0, niw_stream(stream)>>>(pose, intr, ray_idx, idx_start, B, P, W, d_center,
                                                               d_ray, d_pose);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_raygen_unwarped(const float* intr, const float* pose_init, const int64_t* ray_idx,
                                   int64_t idx_start, int B, int P, int n_center, int H, int W, float* pts, void* stream) {
    NIW_CHECK_ARG(intr && pts && B > 0 && P > 0 && H > 0 && W > 0 && (n_center == P || n_center == 1));
    int64_t n = (int64_t)B * P;
    niw::note_launch(), raygen_unwarped_kernel<<<niw_blocks(n, 256), 256, 0, niw_stream(stream)>>>(intr, pose_init, ray_idx, idx_start, B,
                                                                              P, n_center, W, pts);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_rays_from_warp_fwd(const float* warped, int B, int P, int n_center, float* ray, float* center, void* stream) {
    NIW_CHECK_ARG(warped && ray && center && B > 0 && P > 0 && (n_center == P || n_center == 1));
    niw::note_launch(), rays_from_warp_fwd_kernel<<<niw_blocks((int64_t)B * P * 3, 256), 256, 0, niw_stream(stream)>>>(warped, B, P, n_center, ray, center);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_rays_from_warp_bwd(const float* d_ray, const float* d_center, int B, int P, int n_center, float* d_warped,
                                      void* stream) {
    NIW_CHECK_ARG((d_ray || d_center) && d_warped && B > 0 && P > 0 && (n_center == P || n_center == 1));
    niw::note_launch(), rays_from_warp_bwd_kernel<<<B, 256, 0, niw_stream(stream)>>>(d_ray, d_center, P, n_center, d_warped);
    NIW_LAUNCH_CHECK();
    return 0;
}
