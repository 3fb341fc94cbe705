// Adam update over a flat parameter segment.
//   optim.step() / sched.step()   reference model/nerf.py:33-46,87,92  (Adam + ExponentialLR on graph.nerf)
//   optim_pose.step()             reference model/barf_inn_llff.py:84-120 (Adam on warp_mlp + warp_latent)
// The reference's two torch.optim.Adam instances walk ~60 small tensors; here every optimiser group is one
// contiguous fp32 segment (parameters, gradients, both moments), updated by one streaming launch:
// 16 B read + 12 B written per parameter, HBM-bound.  The step counter lives on the device (CUDA-graph
// replayable); the ExponentialLR decay lr_t = lr * gamma^(t-1) is evaluated from it in double precision.
#include "common.cuh"
#include <cmath>

namespace {

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, double lr, double log_gamma, float beta1, float beta2, float eps, float weight_decay,
                 float warmup_iters, int64_t warmup_n, float max_iter, float* __restrict__ progress0,
                 float* __restrict__ progress1, float* __restrict__ step, unsigned int* __restrict__ ticket) {
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const bool vec = i0 + 3 < n && (i0 + 3 < warmup_n || i0 >= warmup_n);     // a straddling float4 is handled per element
    // the element loads go out first: the scalar chain below (step counter -> double exp -> two powf) would otherwise sit
    // in front of them, two dependent memory round trips in a kernel that is all latency
    float4 pp, gg, mm, vv;
    if (vec) {
        pp = *reinterpret_cast<float4*>(p + i0); gg = *reinterpret_cast<const float4*>(g + i0);
        mm = *reinterpret_cast<float4*>(m + i0); vv = *reinterpret_cast<float4*>(v + i0);
    }
    const float t = *reinterpret_cast<volatile float*>(step) + 1.f;   // every thread reads it before any block can advance it
    // ExponentialLR: torch multiplies the (double) learning rate by gamma once per step; gamma^(t-1) is evaluated in
    // double here as well -- an fp32 gamma drifts by ~0.5 % over the 200 000 iterations of the target schedules
    const float lr_full = (float)(log_gamma == 0.0 ? lr : lr * exp(log_gamma * (double)(t - 1.f)));
    // pose-LR warm-up (model/barf.py:48-51): lr *= min(1, it / warmup) with it = iterations completed before this one,
    // applied to the first warmup_n elements only (the reference scales optim_pose.param_groups[0] = warp_mlp, not the
    // latent codes: model/barf_inn_llff.py:108-111)
    const float lr_warm = warmup_iters > 0.f ? lr_full * fminf(1.f, (t - 1.f) / warmup_iters) : lr_full;
    // torch.optim.Adam (single-tensor form): step_size = lr / (1 - b1^t), denom = sqrt(v) / sqrt(1 - b2^t) + eps
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float step_size = (i0 + 3 < warmup_n ? lr_warm : lr_full) / bc1;
    if (vec) {
        float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float gr = G[j] + weight_decay * P[j];
            M[j] = beta1 * M[j] + (1.f - beta1) * gr;
            V[j] = beta2 * V[j] + (1.f - beta2) * gr * gr;
            P[j] -= step_size * (M[j] / (sqrtf(V[j]) / bc2_sqrt + eps));
        }
        *reinterpret_cast<float4*>(p + i0) = pp;
        *reinterpret_cast<float4*>(m + i0) = mm;
        *reinterpret_cast<float4*>(v + i0) = vv;
    } else {
        for (int64_t i = i0; i < n && i < i0 + 4; ++i) {
            float gr = g[i] + weight_decay * p[i];
            float mi = beta1 * m[i] + (1.f - beta1) * gr;
            float vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
            m[i] = mi; v[i] = vi;
            p[i] -= (i < warmup_n ? lr_warm : lr_full) / bc1 * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        }
    }
    // the last block to finish advances the step counter (all blocks have read it by then)
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        step[0] = t; *ticket = 0u;
        // the BARF schedule scalar (model/barf.py:57-59: progress.data.fill_(it / max_iter) after the pose step), kept on
        // the device so that a captured training loop anneals without a host write
        const float prog = (float)((double)t / (double)max_iter);
        if (progress0) progress0[0] = prog;
        if (progress1) progress1[0] = prog;
    }
}

}  // namespace

extern "C" int niw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                             double lr_gamma, float beta1, float beta2, float eps, float weight_decay, float warmup_iters,
                             int64_t warmup_n, float max_iter, float* progress0, float* progress1, float* state,
                             void* stream) {
    NIW_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state && n > 0 && warmup_iters >= 0.f && lr_gamma > 0.0 &&
                  warmup_n >= 0 && (!(progress0 || progress1) || max_iter > 0.f));
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return NIW_E_BADARG;                            // segments must be 16-byte aligned
    niw::note_launch(), adam_flat_kernel<<<niw_blocks((n + 3) / 4, 256), 256, 0, niw_stream(stream)>>>(
        params, grads, exp_avg, exp_avg_sq, n, lr, lr_gamma == 1.0 ? 0.0 : log(lr_gamma), beta1, beta2, eps, weight_decay,
        warmup_iters, warmup_n, max_iter, progress0, progress1, state, reinterpret_cast<unsigned int*>(state + 1));
    NIW_LAUNCH_CHECK();
    return 0;
}
