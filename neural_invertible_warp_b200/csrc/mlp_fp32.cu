// FP32 CUDA-core implementation of NeRF.forward_samples and its backward
// (reference model/nerf.py:416-456, model/barf.py:256-268,
// camera.py:517-521).  This is the parity / high-precision path (NIW_PREC_FP32): every GEMM
// accumulates in fp32 from fp32 operands, so rendered outputs stay within 1e-3 of the fp32
// reference.  The throughput path is the tcgen05 kernel in mlp_tc.cu; both share the encode /
// gradient-reduction kernels in this file.
#include "common.cuh"
#include "nerf_layout.cuh"
#include "mlp_shared.cuh"

namespace niw {

// ------------------------------------------------------------------------------------------
// positional encoding
// ------------------------------------------------------------------------------------------

// one thread per sample: x = c + d*v, enc = [x, per coord: w_k sin(2^k pi x) (k<10), w_k cos(...)]
__global__ void encode_points_kernel(const float* __restrict__ center, const float* __restrict__ ray,
                                     const float* __restrict__ depth, int64_t S, int N, const float* __restrict__ bands,
                                     float* __restrict__ enc) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    Bands3 bw; BandsV bwv_unused;
    load_bands(bands, bw, bwv_unused);
    int64_t r = s / N;
    float d = depth[s];
    float* e = enc + s * ENC3_PAD;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = __fadd_rn(center[r * 3 + c], __fmul_rn(ray[r * 3 + c], d));   // camera.py:520
        e[c] = x;
#pragma unroll
        for (int k = 0; k < L3; ++k) {
            float sn, cs;
            sincosf(x * ((float)(1 << k) * PI_F), &sn, &cs);
            e[3 + c * 2 * L3 + k] = bw.w[k] * sn;
            e[3 + c * 2 * L3 + L3 + k] = bw.w[k] * cs;
        }
    }
    e[ENC3] = 0.f;
}

// one thread per ray: view = normalize(ray) (torch F.normalize, eps 1e-12), encoded with L=4
__global__ void bands_kernel(C2F c2f, float* __restrict__ bands) {
    if (threadIdx.x < NBANDS) store_bands(c2f, threadIdx.x, bands);
}

__global__ void encode_view_kernel(const float* __restrict__ ray, int64_t R, const float* __restrict__ bands,
                                   float* __restrict__ venc) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    Bands3 bw3_unused; BandsV bw;
    load_bands(bands, bw3_unused, bw);
    float v[3] = {ray[r * 3], ray[r * 3 + 1], ray[r * 3 + 2]};
    float inv = 1.f / fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), 1e-12f);
    float* e = venc + r * ENCV_PAD;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = v[c] * inv;
        e[c] = x;
#pragma unroll
        for (int k = 0; k < LV; ++k) {
            float sn, cs;
            sincosf(x * ((float)(1 << k) * PI_F), &sn, &cs);
            e[3 + c * 2 * LV + k] = bw.w[k] * sn;
            e[3 + c * 2 * LV + LV + k] = bw.w[k] * cs;
        }
    }
#pragma unroll
    for (int i = ENCV; i < ENCV_PAD; ++i) e[i] = 0.f;
}

// One warp per ray.  Consumes d_enc [S,64] (gradient wrt the encoded sample position) and
// d_venc_s [S,32] (gradient wrt the per-sample copy of the view encoding) and produces
// d_center [R,3], d_ray [R,3] (the position route x = c + d v and the view route normalize(v)).
__global__ void encode_bwd_kernel(const float* __restrict__ center, const float* __restrict__ ray,
                                  const float* __restrict__ depth, int64_t R, int N, const float* __restrict__ bands,
                                  const float* __restrict__ d_enc, int ld_enc, const float* __restrict__ d_venc_s,
                                  int ld_venc, float* __restrict__ d_center, float* __restrict__ d_ray) {
    const int lane = threadIdx.x & 31;
    Bands3 bw3; BandsV bwv;
    load_bands(bands, bw3, bwv);
    int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    float c3[3] = {center[r * 3], center[r * 3 + 1], center[r * 3 + 2]};
    float v3[3] = {ray[r * 3], ray[r * 3 + 1], ray[r * 3 + 2]};
    float dc[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
    float dve[ENCV];
#pragma unroll
    for (int i = 0; i < ENCV; ++i) dve[i] = 0.f;
    for (int i = lane; i < N; i += 32) {
        int64_t s = r * N + i;
        float d = depth[s];
        const float* g = d_enc + s * ld_enc;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float x = __fadd_rn(c3[c], __fmul_rn(v3[c], d));
            float acc = g[c];
#pragma unroll
            for (int k = 0; k < L3; ++k) {
                float f = (float)(1 << k) * PI_F, sn, cs;
                sincosf(x * f, &sn, &cs);
                acc += bw3.w[k] * f * (cs * g[3 + c * 2 * L3 + k] - sn * g[3 + c * 2 * L3 + L3 + k]);
            }
            dc[c] += acc;
            dv[c] += acc * d;
        }
        if (d_venc_s) {
            const float* gv = d_venc_s + s * ld_venc;
#pragma unroll
            for (int j = 0; j < ENCV; ++j) dve[j] += gv[j];
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) { dc[c] = warp_sum(dc[c]); dv[c] = warp_sum(dv[c]); }
    if (d_venc_s) {
#pragma unroll
        for (int j = 0; j < ENCV; ++j) dve[j] = warp_sum(dve[j]);
    }
    if (lane == 0) {
        if (d_venc_s) {
            float nrm = fmaxf(sqrtf(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]), 1e-12f);
            float inv = 1.f / nrm;
            float u[3] = {v3[0] * inv, v3[1] * inv, v3[2] * inv};
            float du[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float acc = dve[c];
#pragma unroll
                for (int k = 0; k < LV; ++k) {
                    float f = (float)(1 << k) * PI_F, sn, cs;
                    sincosf(u[c] * f, &sn, &cs);
                    acc += bwv.w[k] * f * (cs * dve[3 + c * 2 * LV + k] - sn * dve[3 + c * 2 * LV + LV + k]);
                }
                du[c] = acc;
            }
            float dot = u[0] * du[0] + u[1] * du[1] + u[2] * du[2];
#pragma unroll
            for (int c = 0; c < 3; ++c) dv[c] += (du[c] - u[c] * dot) * inv;   // d normalize(v)
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { d_center[r * 3 + c] = dc[c]; d_ray[r * 3 + c] = dv[c]; }
    }
}

// ------------------------------------------------------------------------------------------
// fp32 SIMT GEMMs (64x64x16 tiles, 256 threads, 4x4 micro-tiles)
// ------------------------------------------------------------------------------------------

constexpr int BM = 64, BN = 64, BK = 16;
enum { EPI_RELU = 0, EPI_LAYER7 = 1, EPI_SIGMOID = 2 };

__device__ __forceinline__ float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// C = epi([A1 | A2] . W^T + bias);  A2 row = m / a2_div (per-ray operand broadcast over samples)
template <int EPI>
__global__ void __launch_bounds__(256)
linear_fwd_kernel(const float* __restrict__ A1, int lda1, int K1, const float* __restrict__ A2, int lda2, int K2,
                  int a2_div, const float* __restrict__ W, int ldw, const float* __restrict__ bias, int64_t M,
                  int Nout, float* __restrict__ C, int ldc, float* __restrict__ aux_pre, float* __restrict__ aux_act) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int K = K1 + K2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;           // 0..1023
            int row = idx >> 4, kk = idx & 15;  // row-major tile read: 16 consecutive k per row
            int k = k0 + kk;
            int64_t m = m0 + row;
            float a = 0.f;
            if (m < M && k < K) a = k < K1 ? A1[m * lda1 + k] : A2[(m / a2_div) * lda2 + (k - K1)];
            As[kk][row] = a;
            int n = n0 + row;
            Ws[kk][row] = (n < Nout && k < K) ? W[(int64_t)n * ldw + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= Nout) continue;
            float v = acc[i][j] + bias[n];
            if (EPI == EPI_RELU) {
                C[m * ldc + n] = fmaxf(v, 0.f);
            } else if (EPI == EPI_LAYER7) {
                if (n == 0) { aux_pre[m] = v; aux_act[m] = softplus1(v); }      // nerf.py:427-431
                else C[m * ldc + (n - 1)] = fmaxf(v, 0.f);
            } else {
                C[m * ldc + n] = 1.f / (1.f + expf(-v));
            }
        }
    }
}

// dX[m,k] (+)= sum_n G[m,n] W[n,k] (+ g1[m]*w1[k]); optional ReLU mask from the saved activation
__global__ void __launch_bounds__(256)
linear_bwd_dx_kernel(const float* __restrict__ G, int ldg, int Nout, const float* __restrict__ W, int ldw, int K,
                     const float* __restrict__ g1, const float* __restrict__ w1, const float* __restrict__ mask,
                     int ldm, int64_t M, float* __restrict__ dX, int ldx, int accumulate) {
    __shared__ float Gs[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int c0 = blockIdx.y * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int n0 = 0; n0 < Nout; n0 += BK) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;
            int row = idx >> 4, kk = idx & 15;
            int64_t m = m0 + row;
            int n = n0 + kk;
            Gs[kk][row] = (m < M && n < Nout) ? G[m * ldg + n] : 0.f;
            int wr = idx >> 6, wc = idx & 63;   // W tile [16 n][64 k], coalesced along k
            int nn = n0 + wr, k = c0 + wc;
            Ws[wr][wc] = (nn < Nout && k < K) ? W[(int64_t)nn * ldw + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Gs[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float gm = g1 ? g1[m] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = c0 + tx * 4 + j;
            if (k >= K) continue;
            float v = acc[i][j];
            if (g1) v += gm * w1[k];
            if (mask && !(mask[m * ldm + k] > 0.f)) v = 0.f;
            if (accumulate) dX[m * ldx + k] += v; else dX[m * ldx + k] = v;
        }
    }
}

// dW[n,k] += sum_m G[m,n] X[m / x_div, k]   (split over m by blockIdx.z, atomics into dW)
__global__ void __launch_bounds__(256)
linear_bwd_dw_kernel(const float* __restrict__ G, int ldg, int Nout, const float* __restrict__ X, int ldx, int x_div,
                     int K, int64_t M, int64_t m_per_block, float* __restrict__ dW, int ldw) {
    __shared__ float Gs[BK][BM + 4];   // [m][n]
    __shared__ float Xs[BK][BN + 4];   // [m][k]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
    const int64_t mb = (int64_t)blockIdx.z * m_per_block;
    int64_t me = mb + m_per_block;
    if (me > M) me = M;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int64_t ms = mb; ms < me; ms += BK) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            int idx = tid + it * 256;
            int mr = idx >> 6, cc = idx & 63;
            int64_t m = ms + mr;
            int n = n0 + cc, k = k0 + cc;
            bool okm = m < me;
            Gs[mr][cc] = (okm && n < Nout) ? G[m * ldg + n] : 0.f;
            Xs[mr][cc] = (okm && k < K) ? X[(m / x_div) * ldx + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Gs[kk][ty * 4 + i]; w[i] = Xs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= Nout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(dW + (int64_t)n * ldw + k, acc[i][j]);
        }
    }
}

// db[n] += sum_m G[m,n]
__global__ void colsum_kernel(const float* __restrict__ G, int ldg, int Nout, int64_t M, int64_t m_per_block,
                              float* __restrict__ db) {
    __shared__ float red[8][33];
    int n = blockIdx.x * 32 + (threadIdx.x & 31);
    int lane_m = threadIdx.x >> 5;   // 8 row lanes
    int64_t mb = (int64_t)blockIdx.y * m_per_block, me = mb + m_per_block;
    if (me > M) me = M;
    float acc = 0.f;
    if (n < Nout)
        for (int64_t m = mb + lane_m; m < me; m += 8) acc += G[m * ldg + n];
    red[lane_m][threadIdx.x & 31] = acc;
    __syncthreads();
    if (lane_m == 0 && n < Nout) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x & 31];
        atomicAdd(db + n, v);
    }
}

__global__ void rgb_sigmoid_bwd_kernel(const float* __restrict__ d_rgb, const float* __restrict__ rgb, int64_t n,
                                       float* __restrict__ g) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float y = rgb[i]; g[i] = d_rgb[i] * y * (1.f - y); }
}

__global__ void softplus_bwd_kernel(const float* __restrict__ d_sigma, const float* __restrict__ pre, int64_t n,
                                    float* __restrict__ g) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float x = pre[i]; g[i] = d_sigma[i] * (x > 20.f ? 1.f : 1.f / (1.f + expf(-x))); }
}

// ------------------------------------------------------------------------------------------
// host-side orchestration
// ------------------------------------------------------------------------------------------

namespace {

struct Fp32Workspace {
    float *bands, *enc, *venc, *h[NFEAT], *hr, *sig_pre, *rgb_keep, *gA, *gB, *g_enc, *g_venc, *g_hr, *g3, *gs;
};

size_t carve(Fp32Workspace* w, float* base, int64_t S, int64_t R, bool training) {
    size_t off = 0;
    auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 63) & ~size_t(63); return p; };
    Fp32Workspace t;
    t.bands = take(NBANDS);
    t.enc = take(S * ENC3_PAD);
    t.venc = take(R * ENCV_PAD);
    for (int l = 0; l < NFEAT; ++l) t.h[l] = nullptr;
    float* ping[2] = {nullptr, nullptr};
    if (training) { for (int l = 0; l < NFEAT; ++l) t.h[l] = take(S * WIDTH); }
    else { ping[0] = take(S * WIDTH); ping[1] = take(S * WIDTH); for (int l = 0; l < NFEAT; ++l) t.h[l] = ping[l & 1]; }
    t.hr = take(S * RGBW);
    t.sig_pre = take(S);
    t.rgb_keep = training ? take(S * 3) : nullptr;
    if (training) {
        t.gA = take(S * WIDTH); t.gB = take(S * WIDTH); t.g_enc = take(S * ENC3_PAD);
        t.g_venc = take(S * ENCV_PAD); t.g_hr = take(S * RGBW); t.g3 = take(S * 3); t.gs = take(S);
    } else {
        t.gA = t.gB = t.g_enc = t.g_venc = t.g_hr = t.g3 = t.gs = nullptr;
    }
    if (w) *w = t;
    return off * sizeof(float);
}

inline dim3 grid_mn(int64_t M, int N) { return dim3((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN)); }

void launch_dw(const float* G, int ldg, int Nout, const float* X, int ldx, int x_div, int K, int64_t M, float* dW,
               int ldw, cudaStream_t st) {
    int tiles = ((Nout + BM - 1) / BM) * ((K + BN - 1) / BN);
    int64_t want = (int64_t)niw_num_sms() * 4 / tiles;
    if (want < 1) want = 1;
    int64_t m_per = (M + want - 1) / want;
    m_per = ((m_per + BK - 1) / BK) * BK;
    if (m_per < 256) m_per = 256;
    unsigned z = (unsigned)((M + m_per - 1) / m_per);
    dim3 grid((Nout + BM - 1) / BM, (K + BN - 1) / BN, z);
    niw::note_launch(), linear_bwd_dw_kernel<<<grid, 256, 0, st>>>(G, ldg, Nout, X, ldx, x_div, K, M, m_per, dW, ldw);
}

void launch_colsum(const float* G, int ldg, int Nout, int64_t M, float* db, cudaStream_t st) {
    int64_t m_per = 4096;
    dim3 grid((Nout + 31) / 32, (unsigned)((M + m_per - 1) / m_per));
    niw::note_launch(), colsum_kernel<<<grid, 256, 0, st>>>(G, ldg, Nout, M, m_per, db);
}

}  // namespace

size_t fp32_workspace_bytes(int64_t R, int N, int training) {
    int64_t S = R * N;
    if (!training) {
        int64_t rays = fp32_eval_chunk_rays(N);
        if (R > rays) { R = rays; S = R * N; }
    }
    return carve(nullptr, nullptr, S, R, training != 0);
}

static int fp32_fwd_chunk(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
                          const C2F& c2f, bool training, float* wsbase, float* rgb, float* sigma,
                          cudaStream_t st) {
    const int64_t S = R * N;
    Fp32Workspace w;
    carve(&w, wsbase, S, R, training);
    niw::note_launch(), bands_kernel<<<1, 32, 0, st>>>(c2f, w.bands);
    niw::note_launch(), encode_points_kernel<<<niw_blocks(S, 128), 128, 0, st>>>(center, ray, depth, S, N, w.bands, w.enc);
    niw::note_launch(), encode_view_kernel<<<niw_blocks(R, 128), 128, 0, st>>>(ray, R, w.bands, w.venc);
    for (int l = 0; l < NFEAT; ++l) {
        const float* Wl = P + feat_w_off(l);
        const float* bl = P + feat_b_off(l);
        const float* A1 = l == 0 ? w.enc : w.h[l - 1];
        int lda1 = l == 0 ? ENC3_PAD : WIDTH, K1 = l == 0 ? ENC3 : WIDTH;
        const float* A2 = l == SKIP ? w.enc : nullptr;
        int K2 = l == SKIP ? ENC3 : 0;
        if (l < NFEAT - 1)
            niw::note_launch(), linear_fwd_kernel<EPI_RELU><<<grid_mn(S, WIDTH), 256, 0, st>>>(A1, lda1, K1, A2, ENC3_PAD, K2, 1, Wl,
                                                                         feat_in(l), bl, S, WIDTH, w.h[l], WIDTH,
                                                                         nullptr, nullptr);
        else
            niw::note_launch(), linear_fwd_kernel<EPI_LAYER7><<<grid_mn(S, WIDTH + 1), 256, 0, st>>>(A1, lda1, K1, nullptr, 0, 0, 1, Wl,
                                                                               feat_in(l), bl, S, WIDTH + 1, w.h[l],
                                                                               WIDTH, w.sig_pre, sigma);
    }
    niw::note_launch(), linear_fwd_kernel<EPI_RELU><<<grid_mn(S, RGBW), 256, 0, st>>>(w.h[NFEAT - 1], WIDTH, WIDTH, w.venc, ENCV_PAD, ENCV, N,
                                                                P + RGB0_W, WIDTH + ENCV, P + RGB0_B, S, RGBW, w.hr,
                                                                RGBW, nullptr, nullptr);
    niw::note_launch(), linear_fwd_kernel<EPI_SIGMOID><<<grid_mn(S, 3), 256, 0, st>>>(w.hr, RGBW, RGBW, nullptr, 0, 0, 1, P + RGB1_W, RGBW,
                                                               P + RGB1_B, S, 3, rgb, 3, nullptr, nullptr);
    if (training) cudaMemcpyAsync(w.rgb_keep, rgb, sizeof(float) * S * 3, cudaMemcpyDeviceToDevice, st);
    return (int)cudaPeekAtLastError();
}

int fp32_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             const C2F& c2f, int training, void* ws, size_t ws_bytes, float* rgb, float* sigma,
             cudaStream_t st) {
    if (ws_bytes < fp32_workspace_bytes(R, N, training)) return NIW_E_WORKSPACE;
    if (training) return fp32_fwd_chunk(P, center, ray, depth, R, N, c2f, true, (float*)ws, rgb, sigma, st);
    const int64_t chunk = fp32_eval_chunk_rays(N);
    for (int64_t r0 = 0; r0 < R; r0 += chunk) {
        int64_t rc = R - r0 < chunk ? R - r0 : chunk;
        int e = fp32_fwd_chunk(P, center + r0 * 3, ray + r0 * 3, depth + r0 * N, rc, N, c2f, false, (float*)ws,
                               rgb + r0 * N * 3, sigma + r0 * N, st);
        if (e) return e;
    }
    return 0;
}

int fp32_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
             void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma,
             float* dP, float* d_center, float* d_ray, cudaStream_t st) {
    if (ws_bytes < fp32_workspace_bytes(R, N, 1)) return NIW_E_WORKSPACE;
    const int64_t S = R * N;
    Fp32Workspace w;
    carve(&w, (float*)ws, S, R, true);
    // RGB head
    niw::note_launch(), rgb_sigmoid_bwd_kernel<<<niw_blocks(S * 3, 256), 256, 0, st>>>(d_rgb, w.rgb_keep, S * 3, w.g3);
    launch_dw(w.g3, 3, 3, w.hr, RGBW, 1, RGBW, S, dP + RGB1_W, RGBW, st);
    launch_colsum(w.g3, 3, 3, S, dP + RGB1_B, st);
    niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, RGBW), 256, 0, st>>>(w.g3, 3, 3, P + RGB1_W, RGBW, RGBW, nullptr, nullptr, w.hr,
                                                         RGBW, S, w.g_hr, RGBW, 0);
    launch_dw(w.g_hr, RGBW, RGBW, w.h[NFEAT - 1], WIDTH, 1, WIDTH, S, dP + RGB0_W, WIDTH + ENCV, st);
    launch_dw(w.g_hr, RGBW, RGBW, w.venc, ENCV_PAD, N, ENCV, S, dP + RGB0_W + WIDTH, WIDTH + ENCV, st);
    launch_colsum(w.g_hr, RGBW, RGBW, S, dP + RGB0_B, st);
    float* g = w.gA;       // gradient wrt the current layer's post-activation output (already masked)
    float* gn = w.gB;
    niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, WIDTH), 256, 0, st>>>(w.g_hr, RGBW, RGBW, P + RGB0_W, WIDTH + ENCV, WIDTH, nullptr,
                                                          nullptr, w.h[NFEAT - 1], WIDTH, S, g, WIDTH, 0);
    cudaMemsetAsync(w.g_venc, 0, sizeof(float) * S * ENCV_PAD, st);
    niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, ENCV), 256, 0, st>>>(w.g_hr, RGBW, RGBW, P + RGB0_W + WIDTH, WIDTH + ENCV, ENCV,
                                                         nullptr, nullptr, nullptr, 0, S, w.g_venc, ENCV_PAD, 0);
    niw::note_launch(), softplus_bwd_kernel<<<niw_blocks(S, 256), 256, 0, st>>>(d_sigma, w.sig_pre, S, w.gs);
    for (int l = NFEAT - 1; l >= 0; --l) {
        const float* Wl = P + feat_w_off(l);
        float* dWl = dP + feat_w_off(l);
        float* dbl = dP + feat_b_off(l);
        const int ldw = feat_in(l);
        const float* X = l == 0 ? w.enc : w.h[l - 1];
        const int ldx = l == 0 ? ENC3_PAD : WIDTH, Kx = l == 0 ? ENC3 : WIDTH;
        if (l == NFEAT - 1) {
            // rows 1..256 = features, row 0 = density  (nerf.py:427-432)
            launch_dw(g, WIDTH, WIDTH, X, ldx, 1, Kx, S, dWl + ldw, ldw, st);
            launch_dw(w.gs, 1, 1, X, ldx, 1, Kx, S, dWl, ldw, st);
            launch_colsum(g, WIDTH, WIDTH, S, dbl + 1, st);
            launch_colsum(w.gs, 1, 1, S, dbl, st);
            niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, WIDTH), 256, 0, st>>>(g, WIDTH, WIDTH, Wl + ldw, ldw, WIDTH, w.gs, Wl, X,
                                                                  WIDTH, S, gn, WIDTH, 0);
        } else {
            launch_dw(g, WIDTH, WIDTH, X, ldx, 1, Kx, S, dWl, ldw, st);
            if (l == SKIP) launch_dw(g, WIDTH, WIDTH, w.enc, ENC3_PAD, 1, ENC3, S, dWl + WIDTH, ldw, st);
            launch_colsum(g, WIDTH, WIDTH, S, dbl, st);
            if (l == SKIP)   // gradient into the re-injected encoding (overwrites; layer 0 adds later)
                niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, ENC3), 256, 0, st>>>(g, WIDTH, WIDTH, Wl + WIDTH, ldw, ENC3, nullptr,
                                                                     nullptr, nullptr, 0, S, w.g_enc, ENC3_PAD, 0);
            if (l > 0)
                niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, WIDTH), 256, 0, st>>>(g, WIDTH, WIDTH, Wl, ldw, WIDTH, nullptr, nullptr,
                                                                      X, WIDTH, S, gn, WIDTH, 0);
            else
                niw::note_launch(), linear_bwd_dx_kernel<<<grid_mn(S, ENC3), 256, 0, st>>>(g, WIDTH, WIDTH, Wl, ldw, ENC3, nullptr, nullptr,
                                                                     nullptr, 0, S, w.g_enc, ENC3_PAD, 1);
        }
        float* t = g; g = gn; gn = t;
    }
    niw::note_launch(), encode_bwd_kernel<<<niw_blocks(R * 32, 128), 128, 0, st>>>(center, ray, depth, R, N, w.bands, w.g_enc, ENC3_PAD,
                                                              w.g_venc, ENCV_PAD, d_center, d_ray);
    return (int)cudaPeekAtLastError();
}

}  // namespace niw
