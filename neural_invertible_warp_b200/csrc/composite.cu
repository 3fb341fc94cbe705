// Volume compositing.  reference model/nerf.py:458-474.
// One warp per ray.  Fast kernels (N = 32*CH, CH in {2,4,6,8}): every lane owns CH CONSECUTIVE samples, moved
// as 128/64-bit vectors (all of a ray's sigma / depth / rgb loads are issued before any arithmetic: ~3 KB in
// flight per warp), a per-lane serial prefix plus one warp-shuffle scan of the lane totals gives the optical
// depth, so transmittance never leaves registers.  The backward pass is the mirrored suffix scan and reuses
// the forward's saved transmittance (weights are recomputed as T (1 - exp(-sigma delta)) when the caller did
// not keep them).  Generic kernels (any N): samples visited 32 at a time with a running carry.
// HBM-bound: 24 B/sample forward (one of prob / trans written), 40 B/sample backward.
#include "common.cuh"

namespace {

constexpr int WARPS = 8;

__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ float warp_suffix_incl_scan(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(WARPS * 32)
composite_fwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ sigma,
                     const float* __restrict__ depth_s, int64_t R, int N, float bg, float* __restrict__ rgb,
                     float* __restrict__ depth, float* __restrict__ opacity, float* __restrict__ prob,
                     float* __restrict__ trans) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        float rx = ray[r * 3], ry = ray[r * 3 + 1], rz = ray[r * 3 + 2];
        float len = sqrtf(rx * rx + ry * ry + rz * rz);
        const float* sg = sigma + r * N;
        const float* dp = depth_s + r * N;
        const float* cs = rgb_s + r * N * 3;
        float carry = 0.f, a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_o = 0.f;
        for (int base = 0; base < N; base += 32) {
            int i = base + lane;
            bool ok = i < N;
            float d = ok ? dp[i] : 0.f;
            float dn = (i + 1 < N) ? dp[i + 1] : 0.f;
            float intv = (i + 1 < N) ? (dn - d) : 1e10f;           // last interval = 1e10 (nerf.py:461)
            float sd = ok ? sg[i] * (intv * len) : 0.f;           // sigma * (intv * ray_length)
            float incl = warp_incl_scan(sd, lane);
            // exclusive prefix by shifting (incl - sd would cancel against the 1e10 last interval)
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.f;
            float T = expf(-(excl + carry));
            float w = T * (1.f - expf(-sd));
            carry += __shfl_sync(0xffffffffu, incl, 31);
            if (ok) {
                if (prob) prob[r * N + i] = w;
                if (trans) trans[r * N + i] = T;
                a_r += w * cs[i * 3]; a_g += w * cs[i * 3 + 1]; a_b += w * cs[i * 3 + 2];
                a_d += w * d; a_o += w;
            }
        }
        a_r = warp_sum(a_r); a_g = warp_sum(a_g); a_b = warp_sum(a_b); a_d = warp_sum(a_d); a_o = warp_sum(a_o);
        if (lane == 0) {
            if (bg >= 0.f) { float k = bg * (1.f - a_o); a_r += k; a_g += k; a_b += k; }
            rgb[r * 3] = a_r; rgb[r * 3 + 1] = a_g; rgb[r * 3 + 2] = a_b;
            depth[r] = a_d; opacity[r] = a_o;
        }
    }
}

// dL/d(sd_i) = T_{i+1} v_i - sum_{j>i} w_j v_j,  v_i = g_rgb.c_i + g_depth d_i + g_op,  T_{i+1} = T_i - w_i
__global__ void __launch_bounds__(WARPS * 32)
composite_bwd_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ sigma,
                     const float* __restrict__ depth_s, const float* __restrict__ prob,
                     const float* __restrict__ trans, int64_t R, int N, float bg, const float* __restrict__ d_rgb,
                     const float* __restrict__ d_depth, const float* __restrict__ d_opacity,
                     float* __restrict__ d_rgb_s, float* __restrict__ d_sigma, float* __restrict__ d_ray) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        float rx = ray[r * 3], ry = ray[r * 3 + 1], rz = ray[r * 3 + 2];
        float len = sqrtf(rx * rx + ry * ry + rz * rz);
        float gr = d_rgb ? d_rgb[r * 3] : 0.f, gg = d_rgb ? d_rgb[r * 3 + 1] : 0.f, gb = d_rgb ? d_rgb[r * 3 + 2] : 0.f;
        float gd = d_depth ? d_depth[r] : 0.f;
        float go = d_opacity ? d_opacity[r] : 0.f;
        if (bg >= 0.f) go -= bg * (gr + gg + gb);
        const float* sg = sigma + r * N;
        const float* dp = depth_s + r * N;
        const float* cs = rgb_s + r * N * 3;
        float carry = 0.f, dlen = 0.f;
        int nchunks = (N + 31) / 32;
        for (int c = nchunks - 1; c >= 0; --c) {
            int i = c * 32 + lane;
            bool ok = i < N;
            float d = ok ? dp[i] : 0.f;
            float dn = (i + 1 < N) ? dp[i + 1] : 0.f;
            float intv = (i + 1 < N) ? (dn - d) : 1e10f;
            float T = ok ? trans[r * N + i] : 0.f;
            float w = !ok ? 0.f : (prob ? prob[r * N + i] : T * (1.f - expf(-(sg[i] * (intv * len)))));
            float c0 = ok ? cs[i * 3] : 0.f, c1 = ok ? cs[i * 3 + 1] : 0.f, c2 = ok ? cs[i * 3 + 2] : 0.f;
            float v = gr * c0 + gg * c1 + gb * c2 + gd * d + go;
            float wv = w * v;
            float sfx = warp_suffix_incl_scan(wv, lane);          // sum_{j>=i} within chunk
            float nxt = __shfl_down_sync(0xffffffffu, sfx, 1);
            if (lane == 31) nxt = 0.f;
            float after = nxt + carry;                           // sum_{j>i} over the whole ray
            carry += __shfl_sync(0xffffffffu, sfx, 0);
            float dsd = (T - w) * v - after;
            if (ok) {
                float s = sg[i];
                d_sigma[r * N + i] = dsd * (intv * len);
                dlen += dsd * s * intv;
                d_rgb_s[(r * N + i) * 3] = w * gr;
                d_rgb_s[(r * N + i) * 3 + 1] = w * gg;
                d_rgb_s[(r * N + i) * 3 + 2] = w * gb;
            }
        }
        dlen = warp_sum(dlen);
        if (lane == 0 && d_ray) {
            float k = len > 0.f ? dlen / len : 0.f;               // d||ray||/dray = ray/||ray||
            d_ray[r * 3] = k * rx; d_ray[r * 3 + 1] = k * ry; d_ray[r * 3 + 2] = k * rz;
        }
    }
}

// ---- fast kernels: N = 32 * CH ---------------------------------------------------------------------------
// Loss head in the compositor's epilogue (reference model/nerf.py:276-288, model/base.py:209-211):
// the warp that has composited ray r = b P + p also gathers the ground-truth pixel image[b, :, ray_idx[p]], writes
// d_unit[r] = 2 scale (rgb - gt) (the loss gradient for a unit upstream gradient) and adds its squared error to the block's
// sum; the blocks leave partial sums and the LAST block to finish (ticket) adds them in index order and writes the loss --
// deterministic, nothing to zero per call (the ticket resets itself; `ticket` is zero before the first call).
struct MseEpilogue {
    const float* image;         // [B,3,HW] or nullptr: no loss head
    const int64_t* ray_idx;     // [P] or nullptr (pixel = idx_start + p)
    int64_t idx_start;
    int P, HW;
    float scale;                // 1 / (3 R)
    float* d_unit;              // [R,3]
    float* partial;             // [gridDim.x]
    unsigned int* ticket;
    float* loss;
};

// lanes 0..2 of the warp fetch the ground-truth channel early (index load -> pixel load: two dependent round trips that
// overlap the ray's sample loads)
__device__ __forceinline__ float mse_fetch(const MseEpilogue& e, int64_t r, int lane) {
    if (!e.image || lane >= 3) return 0.f;
    const int b = (int)(r / e.P), p = (int)(r % e.P);
    const int64_t pix = e.ray_idx ? e.ray_idx[p] : e.idx_start + p;
    return e.image[((int64_t)b * 3 + lane) * e.HW + pix];
}
// every lane passes its channel of the composited colour (lane c < 3 holds channel c); returns the ray's squared error in lane 0
__device__ __forceinline__ float mse_ray(const MseEpilogue& e, int64_t r, int lane, float mine, float gt) {
    float sq = 0.f;
    if (lane < 3) {
        const float diff = mine - gt;
        e.d_unit[r * 3 + lane] = 2.f * e.scale * diff;
        sq = diff * diff;
    }
    sq += __shfl_down_sync(0xffffffffu, sq, 1) + __shfl_down_sync(0xffffffffu, sq, 2);
    return sq;                  // lane 0: all three channels
}
// end of kernel: block sum of the warps' errors, partial[], ticket, ordered final sum by the last block
__device__ __forceinline__ void mse_finish(const MseEpilogue& e, float warp_err, int lane, int warp) {
    __shared__ float red[WARPS];
    __shared__ bool last;
    if (lane == 0) red[warp] = warp_err;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += red[w];
        e.partial[blockIdx.x] = v;
        __threadfence();
        last = atomicAdd(e.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    float v = 0.f;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) v += __ldcg(e.partial + i);
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += red[w];
        *e.loss = t * e.scale;
        *e.ticket = 0u;
    }
}

template <int CH>
__global__ void __launch_bounds__(WARPS * 32)
composite_fwd_vec_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ sigma,
                         const float* __restrict__ depth_s, int64_t R, float bg, float* __restrict__ rgb,
                         float* __restrict__ depth, float* __restrict__ opacity, float* __restrict__ prob,
                         float* __restrict__ trans, MseEpilogue mse) {
    constexpr int N = 32 * CH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float warp_err = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        const float gt = mse_fetch(mse, r, lane);
        float sg[CH], d[CH], c[3 * CH];
        vload<CH>(sigma + r * N + lane * CH, sg);
        vload<CH>(depth_s + r * N + lane * CH, d);
        vload<3 * CH>(rgb_s + (r * N + lane * CH) * 3, c);
        const float rx = ray[r * 3], ry = ray[r * 3 + 1], rz = ray[r * 3 + 2];
        const float len = sqrtf(rx * rx + ry * ry + rz * rz);
        const float dnext = __shfl_down_sync(0xffffffffu, d[0], 1);
        float sd[CH], ps[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float intv = k + 1 < CH ? d[k + 1] - d[k] : (lane < 31 ? dnext - d[k] : 1e10f);   // last = 1e10 (nerf.py:461)
            sd[k] = sg[k] * (intv * len);
            ps[k] = k ? ps[k - 1] + sd[k] : sd[k];
        }
        float incl = warp_incl_scan(ps[CH - 1], lane);
        float before = __shfl_up_sync(0xffffffffu, incl, 1);      // optical depth in front of this lane's first sample
        if (lane == 0) before = 0.f;
        float a_r = 0.f, a_g = 0.f, a_b = 0.f, a_d = 0.f, a_o = 0.f, w[CH], T[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            T[k] = expf(-(k ? before + ps[k - 1] : before));
            w[k] = T[k] * (1.f - expf(-sd[k]));
            a_r += w[k] * c[3 * k]; a_g += w[k] * c[3 * k + 1]; a_b += w[k] * c[3 * k + 2];
            a_d += w[k] * d[k]; a_o += w[k];
        }
        if (prob) vstore<CH>(prob + r * N + lane * CH, w);
        if (trans) vstore<CH>(trans + r * N + lane * CH, T);
        a_r = warp_sum(a_r); a_g = warp_sum(a_g); a_b = warp_sum(a_b); a_d = warp_sum(a_d); a_o = warp_sum(a_o);
        if (bg >= 0.f) { float k = bg * (1.f - a_o); a_r += k; a_g += k; a_b += k; }     // (all lanes hold the sums)
        if (lane == 0) {
            rgb[r * 3] = a_r; rgb[r * 3 + 1] = a_g; rgb[r * 3 + 2] = a_b;
            depth[r] = a_d; opacity[r] = a_o;
        }
        if (mse.image) warp_err += mse_ray(mse, r, lane, lane == 0 ? a_r : (lane == 1 ? a_g : a_b), gt);
    }
    if (mse.image) mse_finish(mse, warp_err, lane, warp);
}

template <int CH>
__global__ void __launch_bounds__(WARPS * 32)
composite_bwd_vec_kernel(const float* __restrict__ ray, const float* __restrict__ rgb_s, const float* __restrict__ sigma,
                         const float* __restrict__ depth_s, const float* __restrict__ prob,
                         const float* __restrict__ trans, int64_t R, float bg, const float* __restrict__ d_rgb,
                         const float* __restrict__ d_depth, const float* __restrict__ d_opacity,
                         float* __restrict__ d_rgb_s, float* __restrict__ d_sigma, float* __restrict__ d_ray,
                         const float* __restrict__ d_unit, const float* __restrict__ d_loss) {
    constexpr int N = 32 * CH;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // loss head of the forward's epilogue: d_rgb += d_loss * d_unit (d_loss: the upstream gradient of the scalar loss)
    const float gl = d_unit ? d_loss[0] : 0.f;
    for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
        float sg[CH], d[CH], c[3 * CH], T[CH], w[CH];
        vload<CH>(sigma + r * N + lane * CH, sg);
        vload<CH>(depth_s + r * N + lane * CH, d);
        vload<3 * CH>(rgb_s + (r * N + lane * CH) * 3, c);
        vload<CH>(trans + r * N + lane * CH, T);
        if (prob) vload<CH>(prob + r * N + lane * CH, w);
        const float rx = ray[r * 3], ry = ray[r * 3 + 1], rz = ray[r * 3 + 2];
        const float len = sqrtf(rx * rx + ry * ry + rz * rz);
        float gr = d_rgb ? d_rgb[r * 3] : 0.f, gg = d_rgb ? d_rgb[r * 3 + 1] : 0.f, gb = d_rgb ? d_rgb[r * 3 + 2] : 0.f;
        if (d_unit) { gr += gl * d_unit[r * 3]; gg += gl * d_unit[r * 3 + 1]; gb += gl * d_unit[r * 3 + 2]; }
        const float gd = d_depth ? d_depth[r] : 0.f;
        float go = d_opacity ? d_opacity[r] : 0.f;
        if (bg >= 0.f) go -= bg * (gr + gg + gb);
        const float dnext = __shfl_down_sync(0xffffffffu, d[0], 1);
        float intv[CH], v[CH], sfx[CH];                           // sfx[k] = sum_{j >= k, same lane} w_j v_j
#pragma unroll
        for (int k = CH - 1; k >= 0; --k) {
            intv[k] = k + 1 < CH ? d[k + 1] - d[k] : (lane < 31 ? dnext - d[k] : 1e10f);
            if (!prob) w[k] = T[k] * (1.f - expf(-(sg[k] * (intv[k] * len))));
            v[k] = gr * c[3 * k] + gg * c[3 * k + 1] + gb * c[3 * k + 2] + gd * d[k] + go;
            sfx[k] = k + 1 < CH ? sfx[k + 1] + w[k] * v[k] : w[k] * v[k];
        }
        const float incl = warp_suffix_incl_scan(sfx[0], lane);
        float later = __shfl_down_sync(0xffffffffu, incl, 1);     // sum over the lanes behind this one
        if (lane == 31) later = 0.f;
        float ds[CH], dc[3 * CH], dlen = 0.f;
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const float after = (k + 1 < CH ? sfx[k + 1] : 0.f) + later;
            const float dsd = (T[k] - w[k]) * v[k] - after;
            ds[k] = dsd * (intv[k] * len);
            dlen += dsd * sg[k] * intv[k];
            dc[3 * k] = w[k] * gr; dc[3 * k + 1] = w[k] * gg; dc[3 * k + 2] = w[k] * gb;
        }
        vstore<CH>(d_sigma + r * N + lane * CH, ds);
        vstore<3 * CH>(d_rgb_s + (r * N + lane * CH) * 3, dc);
        dlen = warp_sum(dlen);
        if (lane == 0 && d_ray) {
            const float k = len > 0.f ? dlen / len : 0.f;         // d||ray||/dray = ray/||ray||
            d_ray[r * 3] = k * rx; d_ray[r * 3 + 1] = k * ry; d_ray[r * 3 + 2] = k * rz;
        }
    }
}

inline unsigned grid_for(int64_t R) {
    int64_t blocks = (R + WARPS - 1) / WARPS;
    int64_t cap = (int64_t)niw_num_sms() * 8;   // 8 x 256 threads = 64 warps / SM resident
    return (unsigned)(blocks < cap ? blocks : cap);
}

template <typename K>
inline unsigned grid_vec(K kernel, int64_t R) {
    return niw_resident_grid(kernel, WARPS * 32, 0, (R + WARPS - 1) / WARPS);
}

}  // namespace

// scratch of the loss-head variant: [0] the ticket (unsigned, zero before the first call), [1 ..] per-block partial sums
constexpr int MSE_SCRATCH_FLOATS = 4096;

static int composite_fwd_launch(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s, int64_t R, int N,
                                float bg, float* rgb, float* depth, float* opacity, float* prob, float* trans,
                                const MseEpilogue& mse, cudaStream_t st) {
    const bool al = niw_aligned16(rgb_s) && niw_aligned16(sigma) && niw_aligned16(depth_s) && niw_aligned16(prob) && niw_aligned16(trans);
    const bool vec = al && (N == 64 || N == 128 || N == 192 || N == 256);
    if (mse.image && !vec) return NIW_E_UNSUPP;         // the loss head rides on the vector kernels only
    niw::note_launch();
#define NIW_FWD(CH) do { unsigned g = grid_vec(composite_fwd_vec_kernel<CH>, R); if (g > MSE_SCRATCH_FLOATS - 1) g = MSE_SCRATCH_FLOATS - 1; \
        composite_fwd_vec_kernel<CH><<<g, WARPS * 32, 0, st>>>(ray, rgb_s, sigma, depth_s, R, bg, rgb, depth, opacity, prob, trans, mse); } while (0)
    if (al && N == 64) NIW_FWD(2);
    else if (al && N == 128) NIW_FWD(4);
    else if (al && N == 192) NIW_FWD(6);
    else if (al && N == 256) NIW_FWD(8);
    else composite_fwd_kernel<<<grid_for(R), WARPS * 32, 0, st>>>(ray, rgb_s, sigma, depth_s, R, N, bg, rgb, depth, opacity, prob, trans);
#undef NIW_FWD
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_composite_fwd(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                                 int64_t R, int N, float bg, float* rgb, float* depth, float* opacity, float* prob,
                                 float* trans, void* stream) {
    NIW_CHECK_ARG(ray && rgb_s && sigma && depth_s && rgb && depth && opacity && R > 0 && N > 0);
    return composite_fwd_launch(ray, rgb_s, sigma, depth_s, R, N, bg, rgb, depth, opacity, prob, trans, MseEpilogue{}, niw_stream(stream));
}

extern "C" int niw_composite_mse_scratch_floats(void) { return MSE_SCRATCH_FLOATS; }

extern "C" int niw_composite_fwd_mse(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                                     int64_t R, int N, float bg, float* rgb, float* depth, float* opacity, float* prob,
                                     float* trans, const float* image, const int64_t* ray_idx, int64_t idx_start, int B, int P,
                                     int H, int W, float* d_unit, float* scratch, float* loss, void* stream) {
    NIW_CHECK_ARG(ray && rgb_s && sigma && depth_s && rgb && depth && opacity && R > 0 && N > 0);
    NIW_CHECK_ARG(image && d_unit && scratch && loss && B > 0 && P > 0 && H > 0 && W > 0 && (int64_t)B * P == R);
    MseEpilogue mse;
    mse.image = image; mse.ray_idx = ray_idx; mse.idx_start = idx_start; mse.P = P; mse.HW = H * W;
    mse.scale = 1.0f / (float)(R * 3); mse.d_unit = d_unit; mse.partial = scratch + 1;
    mse.ticket = reinterpret_cast<unsigned int*>(scratch); mse.loss = loss;
    return composite_fwd_launch(ray, rgb_s, sigma, depth_s, R, N, bg, rgb, depth, opacity, prob, trans, mse, niw_stream(stream));
}

static int composite_bwd_launch(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                                const float* prob, const float* trans, int64_t R, int N, float bg,
                                const float* d_rgb, const float* d_depth, const float* d_opacity, float* d_rgb_s,
                                float* d_sigma, float* d_ray, const float* d_unit, const float* d_loss, cudaStream_t st) {
    const bool al = niw_aligned16(rgb_s) && niw_aligned16(sigma) && niw_aligned16(depth_s) && niw_aligned16(prob) &&
                    niw_aligned16(trans) && niw_aligned16(d_rgb_s) && niw_aligned16(d_sigma);
    const bool vec = al && (N == 64 || N == 128 || N == 192 || N == 256);
    if (d_unit && !vec) return NIW_E_UNSUPP;
    niw::note_launch();
#define NIW_BWD(CH) composite_bwd_vec_kernel<CH><<<grid_vec(composite_bwd_vec_kernel<CH>, R), WARPS * 32, 0, st>>>(ray, rgb_s, sigma, depth_s, prob, trans, R, bg, d_rgb, d_depth, d_opacity, d_rgb_s, d_sigma, d_ray, d_unit, d_loss)
    if (al && N == 64) NIW_BWD(2);
    else if (al && N == 128) NIW_BWD(4);
    else if (al && N == 192) NIW_BWD(6);
    else if (al && N == 256) NIW_BWD(8);
    else composite_bwd_kernel<<<grid_for(R), WARPS * 32, 0, st>>>(ray, rgb_s, sigma, depth_s, prob, trans, R, N, bg, d_rgb, d_depth, d_opacity, d_rgb_s, d_sigma, d_ray);
#undef NIW_BWD
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_composite_bwd(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                                 const float* prob, const float* trans, int64_t R, int N, float bg,
                                 const float* d_rgb, const float* d_depth, const float* d_opacity, float* d_rgb_s,
                                 float* d_sigma, float* d_ray, void* stream) {
    NIW_CHECK_ARG(ray && rgb_s && sigma && depth_s && trans && d_rgb_s && d_sigma && R > 0 && N > 0);   // prob may be NULL
    return composite_bwd_launch(ray, rgb_s, sigma, depth_s, prob, trans, R, N, bg, d_rgb, d_depth, d_opacity, d_rgb_s, d_sigma,
                                d_ray, nullptr, nullptr, niw_stream(stream));
}

extern "C" int niw_composite_bwd_mse(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                                     const float* prob, const float* trans, int64_t R, int N, float bg,
                                     const float* d_rgb, const float* d_depth, const float* d_opacity, const float* d_unit,
                                     const float* d_loss, float* d_rgb_s, float* d_sigma, float* d_ray, void* stream) {
    NIW_CHECK_ARG(ray && rgb_s && sigma && depth_s && trans && d_rgb_s && d_sigma && R > 0 && N > 0 && d_unit && d_loss);
    return composite_bwd_launch(ray, rgb_s, sigma, depth_s, prob, trans, R, N, bg, d_rgb, d_depth, d_opacity, d_rgb_s, d_sigma,
                                d_ray, d_unit, d_loss, niw_stream(stream));
}
