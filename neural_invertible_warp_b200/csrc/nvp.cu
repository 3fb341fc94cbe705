// NVP invertible coupling-layer warp.
//   DeformNetwork.forward    reference model/nvp/nvp_ndr.py:365-468
//   Embedder.embed (anneal)  reference model/nvp/embedder.py:41-50   (point-index quirk kept)
//   euler2rot_2dinv          reference model/nvp/nvp_ndr.py:166-174
// The reference runs ~900 tiny ATen launches per call.  Here one kernel evaluates all three
// coupling blocks for one point per thread; the per-image latent contribution to each first
// layer (W[:,emb:] . code_b + bias) is folded into `code_bias` by the host (B x 128, PyTorch),
// so the kernel only sees the 26- / 13-wide embedded coordinates.  Backward recomputes the
// forward per block, and reduces weight gradients over the block's 128 points in shared memory
// (thread j owns hidden unit j) before one atomicAdd per weight per CTA.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int HID = NIW_NVP_HIDDEN;        // 128
constexpr int NF = NIW_NVP_FREQS;          // 6
constexpr int EA = 2 * (1 + 2 * NF);       // 26 embedded inputs of part a
constexpr int EB = 1 * (1 + 2 * NF);       // 13 embedded inputs of part b
constexpr int OFF_W1A = 0;
constexpr int OFF_W2A = OFF_W1A + HID * EA;
constexpr int OFF_B2A = OFF_W2A + HID;
constexpr int OFF_W1B = OFF_B2A + 1;
constexpr int OFF_W2B = OFF_W1B + HID * EB;
constexpr int OFF_B2B = OFF_W2B + 3 * HID;
constexpr int BLOCK_FLOATS = OFF_B2B + 3;
static_assert(BLOCK_FLOATS == NIW_NVP_BLOCK_FLOATS, "wpack layout");
constexpr int PTS = 128;                   // points (= threads) per CTA
constexpr float BETA = 100.f;

struct Bands { float w[NF]; };

__device__ __forceinline__ float softplus100(float x) {
    float bx = BETA * x;
    return bx > 20.f ? x : log1pf(expf(bx)) / BETA;
}
__device__ __forceinline__ float softplus100_grad(float x) {
    float bx = BETA * x;
    return bx > 20.f ? 1.f : 1.f / (1.f + expf(-bx));
}

// the reference anneals `output[:, a:b]` of a [B,P,1,C] tensor, i.e. along the point axis
template <int D>
__device__ __forceinline__ float quirk_scale(const Bands& bw, int n) {
    if (n >= D && n < D * (2 * NF + 1)) return bw.w[(n - D) / (2 * D)];
    return 1.f;
}

template <int D>
__device__ __forceinline__ void embed(const float* x, float scale, float* e) {
#pragma unroll
    for (int c = 0; c < D; ++c) e[c] = scale * x[c];
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        const float f = (float)(1 << k) * 3.14159274101257324f;  // fp32(pi) * 2^k, as the reference's fp32 freq tensor
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float s, co;
            sincosf(x[c] * f, &s, &co);
            e[D + k * 2 * D + c] = scale * s;
            e[D + k * 2 * D + D + c] = scale * co;
        }
    }
}

// d(embedding)/dx contracted with de
template <int D>
__device__ __forceinline__ void embed_bwd(const float* x, float scale, const float* de, float* dx) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
        float acc = de[c];
#pragma unroll
        for (int k = 0; k < NF; ++k) {
            const float f = (float)(1 << k) * 3.14159274101257324f;
            float s, co;
            sincosf(x[c] * f, &s, &co);
            acc += f * (co * de[D + k * 2 * D + c] - s * de[D + k * 2 * D + D + c]);
        }
        dx[c] += scale * acc;
    }
}

__device__ __forceinline__ void axes(int blk, int& foc, int& o0, int& o1) {
    // form 0 (blocks 0..2): focus z, y, x; the other two in ascending order (nvp_ndr.py:389-399)
    int m = blk % 3;
    foc = m == 0 ? 2 : (m == 1 ? 1 : 0);
    o0 = m == 2 ? 1 : 0;
    o1 = m == 0 ? 1 : 2;
}

template <int NIN>
__device__ __forceinline__ float hidden_pre(const float* __restrict__ W1, const float* __restrict__ bias, int j,
                                            const float* e) {
    float pre = bias[j];
#pragma unroll
    for (int i = 0; i < NIN; ++i) pre += W1[j * NIN + i] * e[i];
    return pre;
}

__device__ __forceinline__ void load_block_weights(float* sw, const float* __restrict__ wpack, int blk) {
    for (int i = threadIdx.x; i < BLOCK_FLOATS; i += blockDim.x) sw[i] = wpack[blk * BLOCK_FLOATS + i];
}

// forward of one coupling block for one point; returns intermediate values needed by backward
struct BlockFwd { float xo[2]; float xf_in; float xf; float y[2]; float c, s; };

__device__ __forceinline__ BlockFwd block_forward(const float* sw, const float* biasA, const float* biasB,
                                                  float sa, float sb, float x[3], int blk) {
    int foc, o0, o1;
    axes(blk, foc, o0, o1);
    BlockFwd r;
    r.xo[0] = x[o0]; r.xo[1] = x[o1]; r.xf_in = x[foc];
    float e[EA];
    embed<2>(r.xo, sa, e);
    float delta = sw[OFF_B2A];
    for (int j = 0; j < HID; ++j)
        delta += sw[OFF_W2A + j] * softplus100(hidden_pre<EA>(sw + OFF_W1A, biasA, j, e));
    r.xf = r.xf_in - delta;
    float e2[EB];
    embed<1>(&r.xf, sb, e2);
    float o[3] = {sw[OFF_B2B], sw[OFF_B2B + 1], sw[OFF_B2B + 2]};
    for (int j = 0; j < HID; ++j) {
        float h = softplus100(hidden_pre<EB>(sw + OFF_W1B, biasB, j, e2));
        o[0] += sw[OFF_W2B + j] * h; o[1] += sw[OFF_W2B + HID + j] * h; o[2] += sw[OFF_W2B + 2 * HID + j] * h;
    }
    sincosf(o[0], &r.s, &r.c);
    r.y[0] = r.xo[0] - o[1]; r.y[1] = r.xo[1] - o[2];
    x[foc] = r.xf;
    x[o0] = r.c * r.y[0] + r.s * r.y[1];      // [[c, s], [-s, c]] (euler2rot_2dinv as assembled)
    x[o1] = -r.s * r.y[0] + r.c * r.y[1];
    return r;
}

__global__ void __launch_bounds__(PTS)
nvp_fwd_kernel(const float* __restrict__ wpack, const float* __restrict__ code_bias, const float* __restrict__ pts,
               Bands bw, int B, int Pt, float* __restrict__ out) {
    __shared__ float sw[BLOCK_FLOATS];
    int64_t t = (int64_t)blockIdx.x * PTS + threadIdx.x;
    bool valid = t < (int64_t)B * Pt;
    int b = valid ? (int)(t / Pt) : 0, n = valid ? (int)(t % Pt) : 0;
    float x[3] = {0.f, 0.f, 0.f};
    if (valid) { x[0] = pts[t * 3]; x[1] = pts[t * 3 + 1]; x[2] = pts[t * 3 + 2]; }
    float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
    for (int blk = 0; blk < NIW_NVP_BLOCKS; ++blk) {
        __syncthreads();
        load_block_weights(sw, wpack, blk);
        __syncthreads();
        const float* biasA = code_bias + ((size_t)(blk * 2 + 0) * B + b) * HID;
        const float* biasB = code_bias + ((size_t)(blk * 2 + 1) * B + b) * HID;
        block_forward(sw, biasA, biasB, sa, sb, x, blk);
    }
    if (valid) { out[t * 3] = x[0]; out[t * 3 + 1] = x[1]; out[t * 3 + 2] = x[2]; }
}

// sum_p A[p][j] * V[p][i] for i < ncols, thread j; results atomically added to dst[j*ldj + i*ldi]
__device__ __forceinline__ void reduce_outer(const float* A, const float* V, int vstride, int ncols, float* dst,
                                             int ldj, int ldi) {
    int j = threadIdx.x;
    for (int i = 0; i < ncols; ++i) {
        float acc = 0.f;
        for (int p = 0; p < PTS; ++p) acc += A[p * (HID + 1) + j] * V[p * vstride + i];
        atomicAdd(dst + (size_t)j * ldj + (size_t)i * ldi, acc);
    }
}

// per-image column sums of A (bias gradient): thread j walks the CTA's points in order
__device__ __forceinline__ void reduce_bias(const float* A, int64_t t0, int Pt, int64_t total, float* dbias_part,
                                            int B) {
    int j = threadIdx.x;
    float acc = 0.f;
    int cur = -1;
    for (int p = 0; p < PTS; ++p) {
        int64_t t = t0 + p;
        if (t >= total) break;
        int b = (int)(t / Pt);
        if (b != cur) {
            if (cur >= 0) atomicAdd(dbias_part + (size_t)cur * HID + j, acc);
            cur = b; acc = 0.f;
        }
        acc += A[p * (HID + 1) + j];
    }
    if (cur >= 0) atomicAdd(dbias_part + (size_t)cur * HID + j, acc);
}

__global__ void __launch_bounds__(PTS)
nvp_bwd_kernel(const float* __restrict__ wpack, const float* __restrict__ code_bias, const float* __restrict__ pts,
               Bands bw, int B, int Pt, const float* __restrict__ d_out, float* __restrict__ d_wpack,
               float* __restrict__ d_code_bias) {
    extern __shared__ float smem[];
    float* sw = smem;                               // BLOCK_FLOATS (padded to 5512)
    float* A = smem + 5512;                         // [PTS][HID+1]
    float* V = A + PTS * (HID + 1);                 // [PTS][EA]  (embedding / dout staging)
    const int64_t total = (int64_t)B * Pt;
    const int64_t t0 = (int64_t)blockIdx.x * PTS;
    const int64_t t = t0 + threadIdx.x;
    const bool valid = t < total;
    const int b = valid ? (int)(t / Pt) : 0, n = valid ? (int)(t % Pt) : 0;
    const float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
    float xin[NIW_NVP_BLOCKS][3];
    float x[3] = {0.f, 0.f, 0.f};
    if (valid) { x[0] = pts[t * 3]; x[1] = pts[t * 3 + 1]; x[2] = pts[t * 3 + 2]; }
    // pass 0: forward, remembering each block's input
    for (int blk = 0; blk < NIW_NVP_BLOCKS; ++blk) {
        __syncthreads();
        load_block_weights(sw, wpack, blk);
        __syncthreads();
        xin[blk][0] = x[0]; xin[blk][1] = x[1]; xin[blk][2] = x[2];
        if (blk + 1 < NIW_NVP_BLOCKS) {
            const float* biasA = code_bias + ((size_t)(blk * 2 + 0) * B + b) * HID;
            const float* biasB = code_bias + ((size_t)(blk * 2 + 1) * B + b) * HID;
            block_forward(sw, biasA, biasB, sa, sb, x, blk);
        }
    }
    float dx[3] = {0.f, 0.f, 0.f};
    if (valid) { dx[0] = d_out[t * 3]; dx[1] = d_out[t * 3 + 1]; dx[2] = d_out[t * 3 + 2]; }
    for (int blk = NIW_NVP_BLOCKS - 1; blk >= 0; --blk) {
        if (blk != NIW_NVP_BLOCKS - 1) {
            __syncthreads();
            load_block_weights(sw, wpack, blk);
            __syncthreads();
        }
        const float* biasA = code_bias + ((size_t)(blk * 2 + 0) * B + b) * HID;
        const float* biasB = code_bias + ((size_t)(blk * 2 + 1) * B + b) * HID;
        float* dW = d_wpack + (size_t)blk * BLOCK_FLOATS;
        float* dbA = d_code_bias + (size_t)(blk * 2 + 0) * B * HID;
        float* dbB = d_code_bias + (size_t)(blk * 2 + 1) * B * HID;
        int foc, o0, o1;
        axes(blk, foc, o0, o1);
        float xx[3] = {xin[blk][0], xin[blk][1], xin[blk][2]};
        BlockFwd f = block_forward(sw, biasA, biasB, sa, sb, xx, blk);
        // ---------------- part b backward ----------------
        float g0 = dx[o0], g1 = dx[o1];
        float dy0 = f.c * g0 - f.s * g1, dy1 = f.s * g0 + f.c * g1;
        float dth = g0 * (-f.s * f.y[0] + f.c * f.y[1]) + g1 * (-f.c * f.y[0] - f.s * f.y[1]);
        float dout[3] = {dth, -dy0, -dy1};
        if (!valid) { dout[0] = dout[1] = dout[2] = 0.f; }
        float dxo[2] = {dy0, dy1};
        float dxf = dx[foc];
        float e2[EB];
        embed<1>(&f.xf, sb, e2);
        for (int j = 0; j < HID; ++j)
            A[threadIdx.x * (HID + 1) + j] = valid ? softplus100(hidden_pre<EB>(sw + OFF_W1B, biasB, j, e2)) : 0.f;
        V[threadIdx.x * EA + 0] = dout[0]; V[threadIdx.x * EA + 1] = dout[1]; V[threadIdx.x * EA + 2] = dout[2];
        __syncthreads();
        reduce_outer(A, V, EA, 3, dW + OFF_W2B, 1, HID);               // dW2b[m][j]
        if (threadIdx.x < 3) {
            float acc = 0.f;
            for (int p = 0; p < PTS; ++p) acc += V[p * EA + threadIdx.x];
            atomicAdd(dW + OFF_B2B + threadIdx.x, acc);
        }
        __syncthreads();
        float de2[EB];
#pragma unroll
        for (int i = 0; i < EB; ++i) de2[i] = 0.f;
        for (int j = 0; j < HID; ++j) {
            float pre = hidden_pre<EB>(sw + OFF_W1B, biasB, j, e2);
            float dh = sw[OFF_W2B + j] * dout[0] + sw[OFF_W2B + HID + j] * dout[1] + sw[OFF_W2B + 2 * HID + j] * dout[2];
            float dpre = dh * softplus100_grad(pre);
            A[threadIdx.x * (HID + 1) + j] = dpre;
#pragma unroll
            for (int i = 0; i < EB; ++i) de2[i] += sw[OFF_W1B + j * EB + i] * dpre;
        }
#pragma unroll
        for (int i = 0; i < EB; ++i) V[threadIdx.x * EA + i] = valid ? e2[i] : 0.f;
        __syncthreads();
        reduce_outer(A, V, EA, EB, dW + OFF_W1B, EB, 1);               // dW1b[j][i]
        reduce_bias(A, t0, Pt, total, dbB, B);
        __syncthreads();
        embed_bwd<1>(&f.xf, sb, de2, &dxf);
        // ---------------- part a backward ----------------
        float ddelta = valid ? -dxf : 0.f;
        float e[EA];
        embed<2>(f.xo, sa, e);
        for (int j = 0; j < HID; ++j)
            A[threadIdx.x * (HID + 1) + j] = valid ? softplus100(hidden_pre<EA>(sw + OFF_W1A, biasA, j, e)) : 0.f;
        V[threadIdx.x * EA + 0] = ddelta;
        __syncthreads();
        reduce_outer(A, V, EA, 1, dW + OFF_W2A, 1, HID);                // dW2a[j]
        if (threadIdx.x == 0) {
            float acc = 0.f;
            for (int p = 0; p < PTS; ++p) acc += V[p * EA];
            atomicAdd(dW + OFF_B2A, acc);
        }
        __syncthreads();
        float de[EA];
#pragma unroll
        for (int i = 0; i < EA; ++i) de[i] = 0.f;
        for (int j = 0; j < HID; ++j) {
            float pre = hidden_pre<EA>(sw + OFF_W1A, biasA, j, e);
            float dpre = sw[OFF_W2A + j] * ddelta * softplus100_grad(pre);
            A[threadIdx.x * (HID + 1) + j] = dpre;
#pragma unroll
            for (int i = 0; i < EA; ++i) de[i] += sw[OFF_W1A + j * EA + i] * dpre;
        }
#pragma unroll
        for (int i = 0; i < EA; ++i) V[threadIdx.x * EA + i] = valid ? e[i] : 0.f;
        __syncthreads();
        reduce_outer(A, V, EA, EA, dW + OFF_W1A, EA, 1);               // dW1a[j][i]
        reduce_bias(A, t0, Pt, total, dbA, B);
        __syncthreads();
        embed_bwd<2>(f.xo, sa, de, dxo);
        dx[foc] = dxf; dx[o0] = dxo[0]; dx[o1] = dxo[1];
    }
}

Bands make_bands(float alpha_ratio) {
    Bands bw;
    for (int i = 0; i < NF; ++i) {
        double a = (double)alpha_ratio * NF - i;
        a = a < 0.0 ? 0.0 : (a > 1.0 ? 1.0 : a);
        bw.w[i] = (float)((1.0 - cos(M_PI * a)) * 0.5);             // embedder.py:47-49
    }
    return bw;
}

constexpr size_t BWD_SMEM = sizeof(float) * (5512 + PTS * (HID + 1) + PTS * EA);

}  // namespace

extern "C" int niw_nvp_warp_fwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                                int B, int Pt, float* out, void* stream) {
    NIW_CHECK_ARG(wpack && code_bias && pts && out && B > 0 && Pt > 0);
    int64_t total = (int64_t)B * Pt;
    niw::note_launch(), nvp_fwd_kernel<<<niw_blocks(total, PTS), PTS, 0, niw_stream(stream)>>>(wpack, code_bias, pts,
                                                                          make_bands(alpha_ratio), B, Pt, out);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_nvp_warp_bwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                                int B, int Pt, const float* d_out, float* d_wpack, float* d_code_bias, void* stream) {
    NIW_CHECK_ARG(wpack && code_bias && pts && d_out && d_wpack && d_code_bias && B > 0 && Pt > 0);
    cudaStream_t st = niw_stream(stream);
    NIW_CUDA(cudaMemsetAsync(d_wpack, 0, sizeof(float) * NIW_NVP_BLOCKS * BLOCK_FLOATS, st));
    NIW_CUDA(cudaMemsetAsync(d_code_bias, 0, sizeof(float) * NIW_NVP_BLOCKS * 2 * (size_t)B * HID, st));
    NIW_CUDA(cudaFuncSetAttribute(nvp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    int64_t total = (int64_t)B * Pt;
    niw::note_launch(), nvp_bwd_kernel<<<niw_blocks(total, PTS), PTS, BWD_SMEM, st>>>(wpack, code_bias, pts, make_bands(alpha_ratio), B, Pt,
                                                                 d_out, d_wpack, d_code_bias);
    NIW_LAUNCH_CHECK();
    return 0;
}
