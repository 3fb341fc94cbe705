// NVP invertible coupling-layer warp.
//   DeformNetwork.forward    reference model/nvp/nvp_ndr.py:365-468
//   Embedder.embed (anneal)  reference model/nvp/embedder.py:41-50   (point-index quirk kept)
//   euler2rot_2dinv          reference model/nvp/nvp_ndr.py:166-174
//   weight_norm first layers reference model/nvp/nvp_ndr.py:291-292,335-336
// The reference issues ~900 tiny ATen launches per call (plus as many again in backward).  Here:
//
//   nvp_pack_fwd_kernel   (3 x 16 CTAs) weight-norm resolution w = g v/||v||, code projection
//                                   code_b = W_c code + b_c + code, and the per-image first-layer
//                                   biases b_0 + w[:, emb:] code_b  ->  wpack, code_bias
//   nvp_fwd_kernel        one WARP per point, lanes = hidden units (4 each): all three coupling
//                         blocks; the weights of all blocks arrive in shared memory by bulk async
//                         copies while the warp loads its point and its image's biases
//   nvp_rays_fwd_kernel   the same with the un-warped grid point computed from the pixel index in
//                         front and ray = warped grid - warped centre behind (one launch for the
//                         train-mode ray generation; opt-in, see functional.fused_warped_rays)
//   nvp_bwd_kernel        one warp per point and round, forward recomputed; the weight gradients
//                         (outer products dpre (x) e) are deferred to a per-round pass in which
//                         every thread adds the round's records into register accumulators for its
//                         fixed slice of the weight image; one global atomic per weight per CTA
//   nvp_pack_bwd_kernel   (3 x 16 CTAs) back through the biases, the code projector and weight-norm,
//                                   accumulating straight into the parameters' .grad buffers
#include "common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

namespace {

constexpr int HID = NIW_NVP_HIDDEN;        // 128
constexpr int NF = NIW_NVP_FREQS;          // 6
constexpr int NB = NIW_NVP_BLOCKS;         // 3
constexpr int EA = 2 * (1 + 2 * NF);       // 26 embedded inputs of part a
constexpr int EB = 1 * (1 + 2 * NF);       // 13 embedded inputs of part b
constexpr int DF = 128;                    // latent code width
// One block of the packed weights (wpack, d_wpack; include/niw_b200.h) IS the shared-memory image the warp kernels
// work on, with odd row strides (bank-conflict free for lane-strided rows):
//     W1a [128][27] (26 used), W2a [128], b2a, W1b [128][13], W2b [3][128], b2b [3], pad
constexpr int SA = EA + 1;                 // 27
constexpr int S_W1A = 0;
constexpr int S_W2A = S_W1A + HID * SA;
constexpr int S_B2A = S_W2A + HID;
constexpr int S_W1B = S_B2A + 1;
constexpr int S_W2B = S_W1B + HID * EB;
constexpr int S_B2B = S_W2B + 3 * HID;
constexpr int S_BLOCK = ((S_B2B + 3 + 3) / 4) * 4;   // 5636 floats
constexpr int BLOCK_FLOATS = S_BLOCK;
static_assert(BLOCK_FLOATS == NIW_NVP_BLOCK_FLOATS, "wpack layout");
constexpr int OFF_W1A = S_W1A, OFF_W2A = S_W2A, OFF_B2A = S_B2A, OFF_W1B = S_W1B, OFF_W2B = S_W2B, OFF_B2B = S_B2B;
constexpr float BETA = 100.f;
constexpr float PI_F = 3.14159274101257324f;   // fp32(pi), as the reference's fp32 freq tensor
constexpr int U = HID / 32;                // hidden units per lane

struct Bands { float w[NF]; };
// position of local point n in the per-image point list the reference would have built (the annealing quirk below is
// keyed on it): n < split ? n + offset : n + offset + jump.  Identity = {0, Pt, 0}.  A ray shard passes the index of its
// first ray in the global per-image list (offset) and the rows it does not hold (jump); the de-duplicated centre row
// of [grid rows ; one centre] lands on a centre index of the full list.
struct IndexMap { int offset, split, jump; };
__device__ __forceinline__ int list_index(const IndexMap& m, int n) { return n < m.split ? n + m.offset : n + m.offset + m.jump; }

// softplus(beta = 100, threshold 20) and its derivative from one ex2, one lg2 and one reciprocal of the special-function
// unit (the libm log1pf(expf()) pair was 22 % of the backward kernel's instructions): with y = fl(1 + e),
// log1p(e) = log(y) - ((y - 1) - e) / y compensates the rounding of 1 + e to first order, so small e keeps its relative
// accuracy; MUFU.LG2 is within 2^-21.4 absolute on [0.5, 2] and 3 ulp elsewhere, i.e. h is within 5e-9 absolute + 4e-7
// relative of the reference's value (tests: 5e-6 on the warped points).  Forward and backward use the same function.
__device__ __forceinline__ void softplus100_both(float x, float& h, float& g) {
    const float bx = BETA * x;
    const float e = __expf(fminf(bx, 20.f));
    const float y = 1.f + e;
    const float ry = __fdividef(1.f, y);
    const float l = __logf(y) - ((y - 1.f) - e) * ry;
    const bool lin = bx > 20.f;
    h = lin ? x : l * (1.f / BETA);
    g = lin ? 1.f : e * ry;
}
__device__ __forceinline__ float softplus100(float x) {
    float h, g;
    softplus100_both(x, h, g);
    return h;
}

// the reference anneals `output[:, a:b]` of a [B,P,1,C] tensor, i.e. along the point axis
template <int D>
__device__ __forceinline__ float quirk_scale(const Bands& bw, int n) {
    if (n >= D && n < D * (2 * NF + 1)) return bw.w[(n - D) / (2 * D)];
    return 1.f;
}

__device__ __forceinline__ void axes(int blk, int& foc, int& o0, int& o1) {
    // form 0 (blocks 0..2): focus z, y, x; the other two in ascending order (nvp_ndr.py:389-399)
    int m = blk % 3;
    foc = m == 0 ? 2 : (m == 1 ? 1 : 0);
    o0 = m == 2 ? 1 : 0;
    o1 = m == 0 ? 1 : 2;
}

// register-array access with a runtime index, compiled to selects (no local-memory spill)
__device__ __forceinline__ float sel3(const float x[3], int i) { return i == 0 ? x[0] : (i == 1 ? x[1] : x[2]); }
__device__ __forceinline__ void put3(float x[3], int i, float v) { if (i == 0) x[0] = v; else if (i == 1) x[1] = v; else x[2] = v; }

__device__ __forceinline__ void load_weights_smem(float* sw, const float* __restrict__ wpack, int nblocks) {
    const uint4* src = reinterpret_cast<const uint4*>(wpack);
    uint4* dst = reinterpret_cast<uint4*>(sw);
    for (int i = threadIdx.x; i < nblocks * (S_BLOCK / 4); i += blockDim.x) dst[i] = src[i];
}

// the forward kernels' form: all blocks' weights with one thread's bulk async copies (TMA engine, one mbarrier) while the
// other threads compute their points' inputs (wpack is 16-byte aligned: the launchers check).  Every thread calls
// weights_arrive() before its first read.
__device__ __forceinline__ void weights_issue(float* sw, const float* __restrict__ wpack, int nblocks, uint64_t* bar) {
    if (threadIdx.x == 0) {
        niw::ptx::mbar_init(bar, 1);
        niw::ptx::fence_mbar_init();
        niw::ptx::mbar_arrive_expect_tx(bar, (uint32_t)(nblocks * S_BLOCK * sizeof(float)));
        for (int b = 0; b < nblocks; ++b)
            niw::ptx::bulk_g2s(sw + (size_t)b * S_BLOCK, wpack + (size_t)b * S_BLOCK, (uint32_t)(S_BLOCK * sizeof(float)), bar);
    }
}
__device__ __forceinline__ void weights_arrive(uint64_t* bar) {
    __syncthreads();                                  // the barrier's initialisation
    niw::ptx::mbar_wait(bar, 0);
}

// embedding of D coordinates into e[D*(1+2NF)], computed cooperatively: lane i < D*NF evaluates
// one (frequency, coordinate) sin/cos pair; every lane then reads the whole vector from shared memory
template <int D>
__device__ __forceinline__ void embed_coop(const float* x, float scale, float* e, int lane) {
    if (lane < D) e[lane] = scale * ((D == 1 || lane == 0) ? x[0] : x[D - 1]);
    if (lane < D * NF) {
        const int k = lane / D, c = lane % D;
        const float xc = (D == 1 || c == 0) ? x[0] : x[D - 1];
        float s, co;
        sincosf(xc * ((float)(1 << k) * PI_F), &s, &co);
        e[D + k * 2 * D + c] = scale * s;
        e[D + k * 2 * D + D + c] = scale * co;
    }
    __syncwarp();
}

struct BlockFwd { float xo[2]; float xf; float y[2]; float c, s; };

// forward of one coupling block for the warp's point; x is replicated in every lane.
// es: per-warp scratch in shared memory (EA floats)
// biasA / biasB: the lane's U per-image first-layer biases (code_bias rows), loaded by the caller ahead of the chain
__device__ __forceinline__ BlockFwd block_forward(const float* sw, const float biasA[U], const float biasB[U], float sa,
                                                  float sb, float x[3], int blk, float* es, int lane) {
    int foc, o0, o1;
    axes(blk, foc, o0, o1);
    BlockFwd r;
    r.xo[0] = sel3(x, o0); r.xo[1] = sel3(x, o1);
    __syncwarp();
    embed_coop<2>(r.xo, sa, es, lane);
    float part = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = lane + 32 * u;
        float pre = biasA[u];
        const float* w = sw + S_W1A + j * SA;
#pragma unroll
        for (int i = 0; i < EA; ++i) pre += w[i] * es[i];
        part += sw[S_W2A + j] * softplus100(pre);
    }
    const float delta = sw[S_B2A] + warp_sum(part);
    r.xf = sel3(x, foc) - delta;
    __syncwarp();
    embed_coop<1>(&r.xf, sb, es, lane);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = lane + 32 * u;
        float pre = biasB[u];
        const float* w = sw + S_W1B + j * EB;
#pragma unroll
        for (int i = 0; i < EB; ++i) pre += w[i] * es[i];
        const float h = softplus100(pre);
        p0 += sw[S_W2B + j] * h; p1 += sw[S_W2B + HID + j] * h; p2 += sw[S_W2B + 2 * HID + j] * h;
    }
    const float o0v = sw[S_B2B] + warp_sum(p0), o1v = sw[S_B2B + 1] + warp_sum(p1), o2v = sw[S_B2B + 2] + warp_sum(p2);
    sincosf(o0v, &r.s, &r.c);
    r.y[0] = r.xo[0] - o1v; r.y[1] = r.xo[1] - o2v;
    put3(x, foc, r.xf);
    put3(x, o0, r.c * r.y[0] + r.s * r.y[1]);      // [[c, s], [-s, c]] (euler2rot_2dinv as assembled)
    put3(x, o1, -r.s * r.y[0] + r.c * r.y[1]);
    return r;
}

// the lane's first-layer biases of image b: all blocks and both parts at once, so that ONE global-memory latency is paid
// ahead of the chain (loaded inside the chain, the six loads were 43 % of the forward kernel's stall samples)
struct LaneBias { float a[NB][U], b[NB][U]; };
__device__ __forceinline__ void load_lane_bias(LaneBias& lb, const float* __restrict__ code_bias, int B, int b, int lane) {
#pragma unroll
    for (int blk = 0; blk < NB; ++blk)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            lb.a[blk][u] = code_bias[((size_t)(blk * 2 + 0) * B + b) * HID + lane + 32 * u];
            lb.b[blk][u] = code_bias[((size_t)(blk * 2 + 1) * B + b) * HID + lane + 32 * u];
        }
}

constexpr int FWD_WARPS = 8;
constexpr size_t FWD_SMEM = sizeof(float) * ((size_t)NB * S_BLOCK + FWD_WARPS * 32) + 8;    // + the weights' mbarrier
static_assert((FWD_SMEM - 8) % 8 == 0, "mbarrier alignment");

__global__ void __launch_bounds__(FWD_WARPS * 32)
nvp_fwd_kernel(const float* __restrict__ wpack, const float* __restrict__ code_bias, const float* __restrict__ pts,
               Bands bw, IndexMap im, int B, int Pt, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* es = smem + (size_t)NB * S_BLOCK + warp * 32;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)NB * S_BLOCK + FWD_WARPS * 32);
    weights_issue(sw, wpack, NB, bar);
    const int64_t total = (int64_t)B * Pt;
    int64_t t = (int64_t)blockIdx.x * FWD_WARPS + warp;
    LaneBias lb;
    float x[3] = {0.f, 0.f, 0.f};
    if (t < total) {                                                      // in flight while the weights arrive
        load_lane_bias(lb, code_bias, B, (int)(t / Pt), lane);
        x[0] = pts[t * 3]; x[1] = pts[t * 3 + 1]; x[2] = pts[t * 3 + 2];
    }
    weights_arrive(bar);
    while (t < total) {
        const int n = list_index(im, (int)(t % Pt));
        const float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) block_forward(sw + (size_t)blk * S_BLOCK, lb.a[blk], lb.b[blk], sa, sb, x, blk, es, lane);
        if (lane < 3) out[t * 3 + lane] = sel3(x, lane);
        t += (int64_t)gridDim.x * FWD_WARPS;
        if (t < total) {
            load_lane_bias(lb, code_bias, B, (int)(t / Pt), lane);
            x[0] = pts[t * 3]; x[1] = pts[t * 3 + 1]; x[2] = pts[t * 3 + 2];
        }
    }
}

// ------------------------------------------------------------------------------------------
// Warped ray generation in ONE kernel (north_star (1); reference model/barf_inn_llff.py:325-364, camera.py:359-390):
// pixel index -> un-warped grid point (K^-1, initial pose) -> three coupling blocks -> ray = warped grid - warped centre.
// Grid (chunks of 8 pixels, images).  Warps 0-7 of a CTA warp one grid point each; warp 8 warps the image's camera centre
// (the same point for every ray of the image, evaluated once per CTA instead of once per ray: it runs next to the grid
// points, so it adds no latency), and after one barrier every grid warp subtracts it.  Also written: the un-warped list
// [grid ; centre] (the backward kernel's input, var.grid_cam / var.center_cam) and the warped list.
// ------------------------------------------------------------------------------------------
struct PoseInvW { float R[9]; float tinv[3]; };
constexpr size_t RAYS_SMEM = sizeof(float) * ((size_t)NB * S_BLOCK + (FWD_WARPS + 1) * 32 + 4) + 8;   // + the weights' mbarrier
static_assert((RAYS_SMEM - 8) % 8 == 0, "mbarrier alignment");

__global__ void __launch_bounds__((FWD_WARPS + 1) * 32)
nvp_rays_fwd_kernel(const float* __restrict__ wpack, const float* __restrict__ code_bias, const float* __restrict__ intr,
                    const float* __restrict__ pose_init, const int64_t* __restrict__ ray_idx, int64_t idx_start, Bands bw,
                    IndexMap im, int B, int P, int W, float* __restrict__ pts, float* __restrict__ warped,
                    float* __restrict__ ray, float* __restrict__ center) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* es = smem + (size_t)NB * S_BLOCK + warp * 32;
    float* s_c = smem + (size_t)NB * S_BLOCK + (FWD_WARPS + 1) * 32;
    const int b = blockIdx.y;
    LaneBias lb;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)NB * S_BLOCK + (FWD_WARPS + 1) * 32 + 4);
    weights_issue(sw, wpack, NB, bar);
    load_lane_bias(lb, code_bias, B, b, lane);                            // in flight while the weights arrive
    const bool is_center = warp == FWD_WARPS;
    const int p = is_center ? P : blockIdx.x * FWD_WARPS + warp;          // local row of the [grid ; centre] list
    const bool active = is_center || p < P;
    float x[3] = {0.f, 0.f, 0.f};
    if (active) {
        // world-frame quantities of the initial pose: Rinv = R^T, tinv = -R^T t   (camera.py:89-95, :343-346)
        float R[9], tinv[3] = {0.f, 0.f, 0.f};
        if (pose_init) {
            const float* q = pose_init + b * 12;
            float t3[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j) R[i * 3 + j] = q[i * 4 + j];
                t3[i] = q[i * 4 + 3];
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) tinv[j] = -(R[0 * 3 + j] * t3[0] + R[1 * 3 + j] * t3[1] + R[2 * 3 + j] * t3[2]);
        }
        if (is_center) {
            x[0] = tinv[0]; x[1] = tinv[1]; x[2] = tinv[2];
        } else {
            const Mat3 Ki = inverse3x3(intr + b * 9);
            const int64_t pix = ray_idx ? ray_idx[p] : idx_start + p;
            float g[3];
            pixel_to_cam(Ki, pix, W, g);
            if (pose_init) {
#pragma unroll
                for (int j = 0; j < 3; ++j) x[j] = R[0 * 3 + j] * g[0] + R[1 * 3 + j] * g[1] + R[2 * 3 + j] * g[2] + tinv[j];
            } else {
                x[0] = g[0]; x[1] = g[1]; x[2] = g[2];
            }
        }
        if (lane < 3 && (!is_center || blockIdx.x == 0)) pts[((int64_t)b * (P + 1) + p) * 3 + lane] = sel3(x, lane);
    }
    weights_arrive(bar);                                            // weights in shared memory
    if (active) {
        const int n = list_index(im, p);
        const float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) block_forward(sw + (size_t)blk * S_BLOCK, lb.a[blk], lb.b[blk], sa, sb, x, blk, es, lane);
        if (is_center && lane < 3) s_c[lane] = sel3(x, lane);
        if (lane < 3 && warped && (!is_center || blockIdx.x == 0)) warped[((int64_t)b * (P + 1) + p) * 3 + lane] = sel3(x, lane);
    }
    __syncthreads();                                                      // the image's warped centre
    if (active && !is_center && lane < 3) {
        const float c = s_c[lane];
        ray[((int64_t)b * P + p) * 3 + lane] = sel3(x, lane) - c;
        center[((int64_t)b * P + p) * 3 + lane] = c;
    }
}

// ------------------------------------------------------------------------------------------
// backward: block-outer order, one warp per point and round.  A warp's pass over a point is a long dependent chain
// (~1 500 instructions per block), so the kernel lives on warps per SM.  The weight gradients are outer products
// dpre (x) e summed over points; they are NOT accumulated by the warp that owns the point (a private accumulator image
// per warp is 22 KB of shared memory and capped the CTA at 8 warps): phase 1 leaves dpre, the activations and the
// embedding of its point in a small record, and after one barrier per round phase 2 adds the round's records into
// accumulators that every thread keeps in REGISTERS for its fixed slice of the weight image (thread = hidden unit x
// 13 embedding columns), flushed with one global atomic per weight per CTA at the end of the block pass.
// ------------------------------------------------------------------------------------------
#ifndef NIW_NVP_BWD_WARPS
#define NIW_NVP_BWD_WARPS 16
#endif
constexpr int BWD_WARPS = NIW_NVP_BWD_WARPS;
constexpr int BWD_THREADS = BWD_WARPS * 32;
static_assert(BWD_THREADS >= 384 && BWD_THREADS % 128 == 0, "phase 2 maps 3 x 128 threads onto the first-layer gradients");
constexpr int MAX_ROUNDS = 32;             // points per warp; per-point state kept in shared memory: xin[3][3] + dx[3]
constexpr int PT_STATE = 12;
constexpr int ES_FLOATS = 2 * (EA + EB) + 2;   // scaled + raw embeddings of both parts
// record of one point and block (phase 1 -> phase 2), floats
constexpr int R_DA = 0;                    // dpre of part a        [128]
constexpr int R_DB = HID;                  // dpre of part b        [128]
constexpr int R_VA = 2 * HID;              // ddelta * h_a          [128]   (gradient of W2a)
constexpr int R_HB = 3 * HID;              // h_b                   [128]   (x dout[m] = gradient of W2b)
constexpr int R_EA = 4 * HID;              // scaled embedding a, two halves of 13 padded to 16
constexpr int R_EB = R_EA + 32;            // scaled embedding b, 13 padded to 16
constexpr int R_MISC = R_EB + 16;          // dout[3], ddelta
constexpr int R_IMG = R_MISC + 4;          // image index (bits)
constexpr int REC = R_IMG + 4;             // 568 floats, a multiple of 4
constexpr size_t BWD_SMEM = sizeof(float) * ((size_t)S_BLOCK + 2 * (size_t)BWD_WARPS * REC + BWD_WARPS * ES_FLOATS +
                                             (size_t)MAX_ROUNDS * BWD_WARPS * PT_STATE);
static_assert(BWD_SMEM <= 227 * 1024, "shared memory budget (NVP backward)");
static_assert(S_BLOCK % 4 == 0 && REC % 4 == 0 && (BWD_WARPS * ES_FLOATS) % 4 == 0, "16-byte alignment of the records");

// scaled embedding e (what the MLP sees) and the raw sin/cos (needed by the derivative)
template <int D>
__device__ __forceinline__ void embed_coop2(const float* x, float scale, float* e, float* raw, int lane) {
    if (lane < D) { const float xv = (D == 1 || lane == 0) ? x[0] : x[D - 1]; e[lane] = scale * xv; raw[lane] = xv; }
    if (lane < D * NF) {
        const int k = lane / D, c = lane % D;
        const float xc = (D == 1 || c == 0) ? x[0] : x[D - 1];
        float s, co;
        sincosf(xc * ((float)(1 << k) * PI_F), &s, &co);
        e[D + k * 2 * D + c] = scale * s; raw[D + k * 2 * D + c] = s;
        e[D + k * 2 * D + D + c] = scale * co; raw[D + k * 2 * D + D + c] = co;
    }
    __syncwarp();
}
// Sums over the lanes of several per-lane values, transposed: after log2 halving steps lane l holds the total of v[l]
// (31 shuffles for 32 values, where one butterfly per value costs 160).  The 16-value form leaves v[l & 15] in lanes
// l and l ^ 16.
__device__ __forceinline__ float warp_sum_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = hi ? v[i] : v[i + o], keep = hi ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}
__device__ __forceinline__ float warp_sum_scatter16(float (&v)[16], int lane) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const bool hi = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = hi ? v[i] : v[i + o], keep = hi ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 16);
}
// element q of the embedding's backward: d(e . de)/dx_c = de[c] + sum_k f_k (cos_kc de[sin_kc] - sin_kc de[cos_kc]); the lane
// holding de[q] contributes its term to coordinate q % D (raw: the un-scaled sin / cos values, embed_coop2)
template <int D>
__device__ __forceinline__ float embed_bwd_term(const float* raw, float de_q, int q) {
    if (q >= D * (1 + 2 * NF)) return 0.f;
    if (q < D) return de_q;
    const int r = q - D, k = r / (2 * D), m = r % (2 * D);
    const float f = (float)(1 << k) * PI_F;
    return m < D ? f * raw[q + D] * de_q : -f * raw[q - D] * de_q;
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
nvp_bwd_kernel(const float* __restrict__ wpack, const float* __restrict__ code_bias, const float* __restrict__ pts,
               Bands bw, IndexMap im, int B, int Pt, int rounds, const float* __restrict__ d_out, float* __restrict__ d_wpack,
               float* __restrict__ d_code_bias) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sw = smem;                                            // one block's weights (padded image)
    float* recs = smem + S_BLOCK;                                // [2][BWD_WARPS][REC]
    float* es = recs + 2 * (size_t)BWD_WARPS * REC + warp * ES_FLOATS;
    float* eA = es, *rawA = es + EA, *eB = es + 2 * EA, *rawB = es + 2 * EA + EB;
    float* state = recs + 2 * (size_t)BWD_WARPS * REC + BWD_WARPS * ES_FLOATS;       // [rounds * BWD_WARPS][PT_STATE]
    const int64_t total = (int64_t)B * Pt;
    const int64_t cta0 = (int64_t)blockIdx.x * rounds * BWD_WARPS;   // the CTA's points: cta0 + round * BWD_WARPS + warp
    const int n_cta = (int)(total - cta0 < (int64_t)rounds * BWD_WARPS ? total - cta0 : (int64_t)rounds * BWD_WARPS);

    // per-point state in shared memory: stt[3*blk .. 3*blk+2] = input of block blk, stt[9..11] = running gradient
    // ---- pass 0: forward through blocks 0 .. NB-2 remembering each block's input; dx <- d_out ----
    for (int blk = 0; blk + 1 < NB; ++blk) {
        __syncthreads();
        load_weights_smem(sw, wpack + (size_t)blk * BLOCK_FLOATS, 1);
        __syncthreads();
        for (int i = warp; i < n_cta; i += BWD_WARPS) {
            const int64_t t = cta0 + i;
            const int b = (int)(t / Pt), n = list_index(im, (int)(t % Pt));
            float* stt = state + i * PT_STATE;
            float x[3];
            if (blk == 0) {
                x[0] = pts[t * 3]; x[1] = pts[t * 3 + 1]; x[2] = pts[t * 3 + 2];
                if (lane < 3) { stt[lane] = sel3(x, lane); stt[9 + lane] = d_out[t * 3 + lane]; }
            } else {
                x[0] = stt[blk * 3]; x[1] = stt[blk * 3 + 1]; x[2] = stt[blk * 3 + 2];
            }
            const float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
            float biasA[U], biasB[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                biasA[u] = code_bias[((size_t)(blk * 2 + 0) * B + b) * HID + lane + 32 * u];
                biasB[u] = code_bias[((size_t)(blk * 2 + 1) * B + b) * HID + lane + 32 * u];
            }
            block_forward(sw, biasA, biasB, sa, sb, x, blk, eA, lane);
            if (lane < 3) stt[(blk + 1) * 3 + lane] = sel3(x, lane);
            __syncwarp();
        }
    }

    // phase-2 roles: threads 0..383 own 13 columns of one hidden unit's first-layer gradient row (part 0, 1: the two halves
    // of W1a; part 2: W1b); the last 128 threads also own the unit's second-layer gradients and the per-image bias sums
    const int j = tid & (HID - 1), part = tid >> 7;
    const bool heavy = tid < 3 * HID, light = tid >= BWD_THREADS - HID;
    const int doff = part < 2 ? R_DA : R_DB, eoff = part < 2 ? R_EA + part * 16 : R_EB;

    // ---- passes NB-1 .. 0: backward of one block for all points of the CTA ----
    for (int blk = NB - 1; blk >= 0; --blk) {
        __syncthreads();
        load_weights_smem(sw, wpack + (size_t)blk * BLOCK_FLOATS, 1);
        __syncthreads();
        int foc, o0, o1;
        axes(blk, foc, o0, o1);
        float a[EB];                                             // heavy: 13 columns
        float a2[4] = {0.f, 0.f, 0.f, 0.f}, ab[2] = {0.f, 0.f}, amisc = 0.f;   // light: W2a, W2b[3]; bias sums; b2b / b2a
#pragma unroll
        for (int q = 0; q < EB; ++q) a[q] = 0.f;
        int cur_img = -1;
        float* dbA = d_code_bias + (size_t)(blk * 2 + 0) * B * HID;
        float* dbB = d_code_bias + (size_t)(blk * 2 + 1) * B * HID;
        for (int r = 0; r * BWD_WARPS < n_cta; ++r) {
            float* rbuf = recs + (size_t)(r & 1) * BWD_WARPS * REC;
            const int i = r * BWD_WARPS + warp;
            if (i < n_cta) {
                // ================= phase 1: this warp's point =================
                float* rc = rbuf + (size_t)warp * REC;
                const int64_t t = cta0 + i;
                const int b = (int)(t / Pt), n = list_index(im, (int)(t % Pt));
                float* stt = state + i * PT_STATE;
                const float x[3] = {stt[blk * 3], stt[blk * 3 + 1], stt[blk * 3 + 2]};
                const float dx[3] = {stt[9], stt[10], stt[11]};
                const float sa = quirk_scale<2>(bw, n), sb = quirk_scale<1>(bw, n);
                float biasA[U], biasB[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    biasA[u] = code_bias[((size_t)(blk * 2 + 0) * B + b) * HID + lane + 32 * u];
                    biasB[u] = code_bias[((size_t)(blk * 2 + 1) * B + b) * HID + lane + 32 * u];
                }
                // ---------------- forward, keeping activations ----------------
                const float xo[2] = {sel3(x, o0), sel3(x, o1)};
                embed_coop2<2>(xo, sa, eA, rawA, lane);
                float hA[U], gA[U], part_sum = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int jj = lane + 32 * u;
                    float pre = biasA[u];
                    const float* w = sw + S_W1A + jj * SA;
#pragma unroll
                    for (int q = 0; q < EA; ++q) pre += w[q] * eA[q];
                    softplus100_both(pre, hA[u], gA[u]);
                    part_sum += sw[S_W2A + jj] * hA[u];
                }
                const float xf = sel3(x, foc) - (sw[S_B2A] + warp_sum(part_sum));
                embed_coop2<1>(&xf, sb, eB, rawB, lane);
                float hB[U], gB[U], q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int jj = lane + 32 * u;
                    float pre = biasB[u];
                    const float* w = sw + S_W1B + jj * EB;
#pragma unroll
                    for (int q = 0; q < EB; ++q) pre += w[q] * eB[q];
                    softplus100_both(pre, hB[u], gB[u]);
                    q0 += sw[S_W2B + jj] * hB[u]; q1 += sw[S_W2B + HID + jj] * hB[u]; q2 += sw[S_W2B + 2 * HID + jj] * hB[u];
                }
                const float th = sw[S_B2B] + warp_sum(q0), t1 = sw[S_B2B + 1] + warp_sum(q1), t2 = sw[S_B2B + 2] + warp_sum(q2);
                float sn, cs;
                sincosf(th, &sn, &cs);
                const float y0 = xo[0] - t1, y1 = xo[1] - t2;
                // ---------------- part b backward ----------------
                const float g0 = sel3(dx, o0), g1 = sel3(dx, o1);
                const float dy0 = cs * g0 - sn * g1, dy1 = sn * g0 + cs * g1;
                const float dth = g0 * (-sn * y0 + cs * y1) + g1 * (-cs * y0 - sn * y1);
                const float dout[3] = {dth, -dy0, -dy1};
                float dxo[2] = {dy0, dy1};
                float dxf = sel3(dx, foc);
                float de2[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) de2[q] = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int jj = lane + 32 * u;
                    const float* w = sw + S_W1B + jj * EB;
                    const float dh = sw[S_W2B + jj] * dout[0] + sw[S_W2B + HID + jj] * dout[1] + sw[S_W2B + 2 * HID + jj] * dout[2];
                    const float dpre = dh * gB[u];
                    rc[R_DB + jj] = dpre;
                    rc[R_HB + jj] = hB[u];
#pragma unroll
                    for (int q = 0; q < EB; ++q) de2[q] += w[q] * dpre;
                }
                {
                    float tb = embed_bwd_term<1>(rawB, warp_sum_scatter16(de2, lane), lane & 15);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) tb += __shfl_xor_sync(0xffffffffu, tb, o);
                    dxf += sb * tb;
                }
                // ---------------- part a backward ----------------
                const float ddelta = -dxf;
                float de[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) de[q] = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int jj = lane + 32 * u;
                    const float* w = sw + S_W1A + jj * SA;
                    const float dpre = sw[S_W2A + jj] * ddelta * gA[u];
                    rc[R_DA + jj] = dpre;
                    rc[R_VA + jj] = ddelta * hA[u];
#pragma unroll
                    for (int q = 0; q < EA; ++q) de[q] += w[q] * dpre;
                }
                {
                    // coordinate of element q is q % 2: the butterfly over lane bits 4..1 sums each parity class
                    float ta = embed_bwd_term<2>(rawA, warp_sum_scatter32(de, lane), lane);
#pragma unroll
                    for (int o = 16; o > 1; o >>= 1) ta += __shfl_xor_sync(0xffffffffu, ta, o);
                    dxo[0] += sa * __shfl_sync(0xffffffffu, ta, 0);
                    dxo[1] += sa * __shfl_sync(0xffffffffu, ta, 1);
                }
                if (lane < EA) rc[R_EA + (lane / EB) * 16 + lane % EB] = eA[lane];
                if (lane < EB) rc[R_EB + lane] = eB[lane];
                if (lane < 3) rc[R_MISC + lane] = sel3(dout, lane);
                if (lane == 3) rc[R_MISC + 3] = ddelta;
                if (lane == 4) rc[R_IMG] = __int_as_float(b);
                __syncwarp();
                if (lane == 0) { stt[9 + foc] = dxf; stt[9 + o0] = dxo[0]; stt[9 + o1] = dxo[1]; }
            }
            __syncthreads();
            // ================= phase 2: the round's records into the register accumulators =================
            const int nv = n_cta - r * BWD_WARPS < BWD_WARPS ? n_cta - r * BWD_WARPS : BWD_WARPS;
            if (heavy) {
#pragma unroll 4
                for (int p = 0; p < nv; ++p) {
                    const float* rp = rbuf + (size_t)p * REC;
                    const float d = rp[doff + j];
                    const float4* e4 = reinterpret_cast<const float4*>(rp + eoff);
                    const float4 e0 = e4[0], e1 = e4[1], e2 = e4[2];
                    const float e12 = rp[eoff + 12];
                    a[0] += d * e0.x; a[1] += d * e0.y; a[2] += d * e0.z; a[3] += d * e0.w;
                    a[4] += d * e1.x; a[5] += d * e1.y; a[6] += d * e1.z; a[7] += d * e1.w;
                    a[8] += d * e2.x; a[9] += d * e2.y; a[10] += d * e2.z; a[11] += d * e2.w;
                    a[12] += d * e12;
                }
            }
            if (light) {
                for (int p = 0; p < nv; ++p) {
                    const float* rp = rbuf + (size_t)p * REC;
                    const int img = __float_as_int(rp[R_IMG]);
                    if (img != cur_img) {
                        if (cur_img >= 0) {
                            atomicAdd(dbA + (size_t)cur_img * HID + j, ab[0]);
                            atomicAdd(dbB + (size_t)cur_img * HID + j, ab[1]);
                            ab[0] = 0.f; ab[1] = 0.f;
                        }
                        cur_img = img;
                    }
                    ab[0] += rp[R_DA + j]; ab[1] += rp[R_DB + j];
                    a2[0] += rp[R_VA + j];
                    const float hb = rp[R_HB + j];
                    a2[1] += rp[R_MISC] * hb; a2[2] += rp[R_MISC + 1] * hb; a2[3] += rp[R_MISC + 2] * hb;
                    if (j < 4) amisc += rp[R_MISC + j];
                }
            }
            // (no second barrier: the next round writes the other record buffer)
        }
        // ---- one global atomic per weight per CTA ----
        float* dW = d_wpack + (size_t)blk * BLOCK_FLOATS;
        if (heavy && n_cta > 0) {
            float* row = part < 2 ? dW + OFF_W1A + j * SA + part * EB : dW + OFF_W1B + j * EB;
#pragma unroll
            for (int q = 0; q < EB; ++q) if (a[q] != 0.f) atomicAdd(row + q, a[q]);
        }
        if (light && cur_img >= 0) {
            atomicAdd(dbA + (size_t)cur_img * HID + j, ab[0]);
            atomicAdd(dbB + (size_t)cur_img * HID + j, ab[1]);
            atomicAdd(dW + OFF_W2A + j, a2[0]);
#pragma unroll
            for (int m = 0; m < 3; ++m) atomicAdd(dW + OFF_W2B + m * HID + j, a2[1 + m]);
            if (j < 3) atomicAdd(dW + OFF_B2B + j, amisc);
            if (j == 3) atomicAdd(dW + OFF_B2A, amisc);
        }
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

// ------------------------------------------------------------------------------------------
// pack: parameters -> effective weights + per-image biases, and its backward
// ------------------------------------------------------------------------------------------
// pointer table, per block b (12 entries): v_a, g_a, b0_a, W1_a, b1_a, v_b, g_b, b0_b, W1_b, b1_b, W_c, b_c
struct PtrTable { const float* p[NIW_NVP_PARAM_PTRS]; };
struct GradTable { float* p[NIW_NVP_PARAM_PTRS]; };
constexpr int MAX_IMG = 96;                // images per call (dynamic shared memory: ~2 KB per image)
constexpr int LDC = DF + 1;
constexpr int PACK_THREADS = 1024;

// squared row norm of v[j][0..ld) by one warp (coalesced)
__device__ __forceinline__ float row_norm2(const float* __restrict__ row, int ld, int lane) {
    // ld <= EA + DF = 154 < 160: all five loads of a lane are issued before the first is used (the rolled loop paid one
    // global-memory latency per 32 columns, 40 in sequence for a warp's 8 rows)
    float t[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { const int c = lane + 32 * i; t[i] = c < ld ? row[c] : 0.f; }
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) a += t[i] * t[i];
    return warp_sum(a);
}
static_assert(EA + DF <= 160 && EB + DF <= 160, "row_norm2 covers 160 columns");

// Grid (coupling block, split): the work of one block is latency-bound in a single CTA (about 270 k warp
// instructions), so it is spread over `gridDim.y` CTAs.  Every reduction is done by a warp with the lanes along
// the contiguous (reduction) axis, so all global reads are coalesced.
//   forward:  split y owns images [y*ipc, (y+1)*ipc): their code projection and per-image biases; every CTA
//             recomputes the 256 weight-norm row scales it needs (8 rows per warp); split 0 writes wpack
__global__ void __launch_bounds__(PACK_THREADS)
nvp_pack_fwd_kernel(PtrTable T, const float* __restrict__ code, int B, int ipc, float* __restrict__ wpack,
                    float* __restrict__ code_bias, float* __restrict__ cb_out) {
    extern __shared__ float sm[];
    float* s_cb = sm;                          // [ipc][LDC]
    float* s_scale = s_cb + (size_t)ipc * LDC; // [2][HID]
    const int blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = PACK_THREADS / 32;
    const int img0 = blockIdx.y * ipc, nimg = min(ipc, B - img0);
    const bool writer = blockIdx.y == 0;
    const float* const* P = T.p + blk * 12;
    const float *Wc = P[10], *bc = P[11];
    float* wp = wpack + (size_t)blk * BLOCK_FLOATS;
    if (nimg <= 0 && !writer) return;          // fewer images than splits: nothing to do here
    // (1) row scales g/||v|| and the effective embedded-coordinate columns.  A warp's rows are taken four at a time with
    // all their loads issued before the first norm (one global-memory round trip per four rows, not per row)
    constexpr int RB = 4;
    static_assert((2 * HID) % (RB * (PACK_THREADS / 32)) == 0, "rows per warp");
    for (int r0 = warp; r0 < 2 * HID; r0 += RB * nwarp) {
        float t[RB][5], gsc[RB];
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int r = r0 + a * nwarp, part = r >> 7, j = r & 127;
            const int ld = (part == 0 ? EA : EB) + DF;
            const float* v = P[part * 5 + 0] + (size_t)j * ld;
#pragma unroll
            for (int i = 0; i < 5; ++i) { const int c = lane + 32 * i; t[a][i] = c < ld ? v[c] : 0.f; }
            gsc[a] = P[part * 5 + 1][j];
        }
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int r = r0 + a * nwarp, part = r >> 7, j = r & 127;
            const int emb = part == 0 ? EA : EB;
            float n2 = 0.f;
#pragma unroll
            for (int i = 0; i < 5; ++i) n2 += t[a][i] * t[a][i];
            const float scale = gsc[a] / sqrtf(warp_sum(n2));
            if (lane == 0) s_scale[r] = scale;
            if (writer && lane < emb) wp[(part == 0 ? OFF_W1A + j * SA : OFF_W1B + j * EB) + lane] = t[a][0] * scale;
        }
    }
    if (writer) {   // output layers: copies
        for (int i = tid; i < HID; i += PACK_THREADS) wp[OFF_W2A + i] = P[3][i];
        for (int i = tid; i < 3 * HID; i += PACK_THREADS) wp[OFF_W2B + i] = P[8][i];
        if (tid == 0) wp[OFF_B2A] = P[4][0];
        if (tid < 3) wp[OFF_B2B + tid] = P[9][tid];
    }
    // (2) code_b[img][k] = code + b_c + W_c code      (nvp_ndr.py:382): warp per output, lanes over m
    for (int o0 = warp; o0 < nimg * DF; o0 += RB * nwarp) {
        float wv[RB][DF / 32], cv[RB][DF / 32];
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int o = o0 + a * nwarp;
            const bool on = o < nimg * DF;
            const int li = on ? o / DF : 0, k = on ? o % DF : 0;
            const float* c = code + (size_t)(img0 + li) * DF;
            const float* w = Wc + (size_t)k * DF;
#pragma unroll
            for (int q = 0; q < DF / 32; ++q) { wv[a][q] = w[lane + 32 * q]; cv[a][q] = c[lane + 32 * q]; }
        }
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int o = o0 + a * nwarp;
            if (o < nimg * DF) {
                const int li = o / DF, k = o % DF;
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < DF / 32; ++q) acc += wv[a][q] * cv[a][q];
                acc = warp_sum(acc);
                if (lane == 0) {
                    acc += code[(size_t)(img0 + li) * DF + k] + bc[k];
                    s_cb[li * LDC + k] = acc;
                    cb_out[((size_t)blk * B + img0 + li) * DF + k] = acc;
                }
            }
        }
    }
    __syncthreads();
    // (3) per-image first-layer biases: warp per (part, j), lanes over the latent axis, loop over images (rows four at a time
    // as above)
    for (int r0 = warp; r0 < 2 * HID; r0 += RB * nwarp) {
        float vl[RB][DF / 32], b0[RB];
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int r = r0 + a * nwarp, part = r >> 7, j = r & 127;
            const int emb = part == 0 ? EA : EB, ld = emb + DF;
            const float* v = P[part * 5 + 0] + (size_t)j * ld + emb;
#pragma unroll
            for (int q = 0; q < DF / 32; ++q) vl[a][q] = v[lane + 32 * q];
            b0[a] = P[part * 5 + 2][j];
        }
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            const int r = r0 + a * nwarp, part = r >> 7, j = r & 127;
            const float scale = s_scale[r];
            for (int li = 0; li < nimg; ++li) {
                float acc = 0.f;
#pragma unroll
                for (int q = 0; q < DF / 32; ++q) acc += vl[a][q] * s_cb[li * LDC + lane + 32 * q];
                acc = warp_sum(acc);
                if (lane == 0) code_bias[((size_t)(blk * 2 + part) * B + img0 + li) * HID + j] = b0[a] + scale * acc;
            }
        }
    }
}

//   backward: split y owns first-layer rows [y*rpc, (y+1)*rpc) of the 256 (part, j) rows (weight-norm backward,
//             single writer per gradient element) and latent columns k in [y*kpc, (y+1)*kpc) of d code_b, from
//             which it derives its rows of the code-projector gradient and its partial sum of d code
constexpr int PACK_SPLIT = 16;
constexpr int RPC = 2 * HID / PACK_SPLIT;      // 16 rows per split
constexpr int KPC = DF / PACK_SPLIT;           // 8 latent columns per split
constexpr int JQ = 8;                          // partitions of the (part, j) reduction of d code_b
static_assert(PACK_SPLIT * RPC == 2 * HID && PACK_SPLIT * KPC == DF, "pack split");

__global__ void __launch_bounds__(PACK_THREADS)
nvp_pack_bwd_kernel(PtrTable T, GradTable G, const float* __restrict__ code, const float* __restrict__ cb, int B,
                    const float* __restrict__ d_wpack, const float* __restrict__ d_code_bias, float* __restrict__ d_code) {
    extern __shared__ float sm[];
    float* s_cb = sm;                          // [B][LDC]      code_b
    float* s_dcb = s_cb + (size_t)B * LDC;     // [2][B][LDC]   d(code_bias) of both parts
    float* s_dc = s_dcb + 2 * (size_t)B * LDC; // [B][KPC]      d(code_b), this split's columns
    float* s_scale = s_dc + (size_t)B * KPC;   // [2][HID]  g / ||v||
    float* s_nrm2 = s_scale + 2 * HID;         // [2][HID]  ||v||^2
    const int blk = blockIdx.x, y = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = PACK_THREADS / 32;
    const float* const* P = T.p + blk * 12;
    float* const* Gp = G.p + blk * 12;
    const float* dwp = d_wpack + (size_t)blk * BLOCK_FLOATS;
    for (int idx = tid; idx < B * DF; idx += PACK_THREADS) s_cb[(idx / DF) * LDC + idx % DF] = cb[(size_t)blk * B * DF + idx];
    for (int idx = tid; idx < 2 * B * HID; idx += PACK_THREADS) {
        const int p = idx / (B * HID), r = idx % (B * HID);
        s_dcb[((size_t)p * B + r / HID) * LDC + r % HID] = d_code_bias[((size_t)(blk * 2 + p) * B) * HID + r];
    }
    for (int idx = tid; idx < B * KPC; idx += PACK_THREADS) s_dc[idx] = 0.f;
    if (y == 0) {   // output layers: the gradient is the warp kernel's own (every element has exactly one writer here)
        for (int i = tid; i < HID; i += PACK_THREADS) Gp[3][i] += dwp[OFF_W2A + i];
        for (int i = tid; i < 3 * HID; i += PACK_THREADS) Gp[8][i] += dwp[OFF_W2B + i];
        if (tid == 0) Gp[4][0] += dwp[OFF_B2A];
        if (tid < 3) Gp[9][tid] += dwp[OFF_B2B + tid];
    }
    // row norms of all 256 rows (the d code_b reduction below runs over every row)
    for (int r = warp; r < 2 * HID; r += nwarp) {
        const int part = r >> 7, j = r & 127;
        const int ld = (part == 0 ? EA : EB) + DF;
        const float nrm2 = row_norm2(P[part * 5 + 0] + (size_t)j * ld, ld, lane);
        if (lane == 0) { s_scale[r] = P[part * 5 + 1][j] / sqrtf(nrm2); s_nrm2[r] = nrm2; }
    }
    __syncthreads();
    // first layers, this split's rows: warp per (part, j); lanes along the columns of row j
    for (int rr = warp; rr < RPC; rr += nwarp) {
        const int r = y * RPC + rr, part = r >> 7, j = r & 127;
        const int emb = part == 0 ? EA : EB, ld = emb + DF;
        const float* v = P[part * 5 + 0] + (size_t)j * ld;
        const float scale = s_scale[r], nrm2 = s_nrm2[r], nrm = sqrtf(nrm2);
        const float* my_dcb = s_dcb + (size_t)part * B * LDC + j;
        // d b0[j] = sum_img dcb
        float db0 = 0.f;
        for (int img = lane; img < B; img += 32) db0 += my_dcb[img * LDC];
        db0 = warp_sum(db0);
        if (lane == 0) Gp[part * 5 + 2][j] += db0;
        // gradient of the effective row: embedded columns from the warp kernel, latent columns
        // d w0[j][emb+k] = sum_img dcb[img][j] cb[img][k]
        const float* dW1 = dwp + (part == 0 ? OFF_W1A + j * SA : OFF_W1B + j * EB);
        float dwe = lane < emb ? dW1[lane] : 0.f;
        float dwl[DF / 32];
#pragma unroll
        for (int q = 0; q < DF / 32; ++q) dwl[q] = 0.f;
        for (int img = 0; img < B; ++img) {
            const float d = my_dcb[img * LDC];
#pragma unroll
            for (int q = 0; q < DF / 32; ++q) dwl[q] += d * s_cb[img * LDC + lane + 32 * q];
        }
        float dot = lane < emb ? dwe * v[lane] : 0.f;
#pragma unroll
        for (int q = 0; q < DF / 32; ++q) dot += dwl[q] * v[emb + lane + 32 * q];
        dot = warp_sum(dot);
        // weight-norm backward: w = g v/||v||
        if (lane == 0) Gp[part * 5 + 1][j] += dot / nrm;
        float* dv = Gp[part * 5 + 0] + (size_t)j * ld;
        const float coef = dot / nrm2;
        if (lane < emb) dv[lane] += scale * (dwe - v[lane] * coef);
#pragma unroll
        for (int q = 0; q < DF / 32; ++q) {
            const int c = emb + lane + 32 * q;
            dv[c] += scale * (dwl[q] - v[c] * coef);
        }
    }
    // d code_b[img][k] = sum_part sum_j dcb[part][img][j] w0[j][emb+k] for this split's k: thread (k, img, jq) sums a
    // 1/JQ share of the 256 (part, j) rows, shares are combined with shared-memory atomics
    for (int idx = tid; idx < KPC * B * JQ; idx += PACK_THREADS) {
        const int kk = idx % KPC, img = (idx / KPC) % B, jq = idx / (KPC * B);
        const int k = y * KPC + kk;
        float acc = 0.f;
        for (int r = jq * (2 * HID / JQ); r < (jq + 1) * (2 * HID / JQ); ++r) {
            const int p = r >> 7, jj = r & 127, e2 = p == 0 ? EA : EB;
            acc += s_dcb[((size_t)p * B + img) * LDC + jj] * P[p * 5 + 0][(size_t)jj * (e2 + DF) + e2 + k] * s_scale[r];
        }
        atomicAdd(&s_dc[img * KPC + kk], acc);
    }
    __syncthreads();
    // code projector: code_b = W_c code + b_c + code; this split's rows k of W_c / b_c
    const float* Wc = P[10];
    for (int idx = tid; idx < KPC * DF; idx += PACK_THREADS) {
        const int kk = idx / DF, m = idx % DF;
        float acc = 0.f;
        for (int img = 0; img < B; ++img) acc += s_dc[img * KPC + kk] * code[(size_t)img * DF + m];
        Gp[10][(size_t)(y * KPC + kk) * DF + m] += acc;
    }
    for (int kk = tid; kk < KPC; kk += PACK_THREADS) {
        float acc = 0.f;
        for (int img = 0; img < B; ++img) acc += s_dc[img * KPC + kk];
        Gp[11][y * KPC + kk] += acc;
    }
    // d code[img][m] += (identity term for this split's columns) + sum_{k in split} d code_b[img][k] W_c[k][m]
    for (int idx = tid; idx < B * DF; idx += PACK_THREADS) {
        const int img = idx / DF, m = idx % DF;
        float acc = (m / KPC == y) ? s_dc[img * KPC + m % KPC] : 0.f;
#pragma unroll
        for (int kk = 0; kk < KPC; ++kk) acc += s_dc[img * KPC + kk] * Wc[(size_t)(y * KPC + kk) * DF + m];
        atomicAdd(d_code + idx, acc);        // three blocks x PACK_SPLIT splits contribute
    }
}

}  // namespace

extern "C" int niw_nvp_pack_fwd(const float* const* params, const float* code, int B, float* wpack, float* code_bias,
                                float* cb, void* stream) {
    NIW_CHECK_ARG(params && code && wpack && code_bias && cb && B > 0);
    PtrTable T;
    for (int i = 0; i < NIW_NVP_PARAM_PTRS; ++i) { NIW_CHECK_ARG(params[i]); T.p[i] = params[i]; }
    if (B > MAX_IMG) return NIW_E_UNSUPP;
    const int nsplit = B < 32 ? B : 32, ipc = (B + nsplit - 1) / nsplit;      // images per CTA
    const size_t smem = sizeof(float) * ((size_t)ipc * LDC + 2 * HID);
    niw::note_launch(), nvp_pack_fwd_kernel<<<dim3(NB, (B + ipc - 1) / ipc), PACK_THREADS, smem, niw_stream(stream)>>>(
        T, code, B, ipc, wpack, code_bias, cb);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_nvp_pack_bwd(const float* const* params, float* const* grads, const float* code, const float* cb,
                                const float* d_wpack, const float* d_code_bias, int B, float* d_code, void* stream) {
    NIW_CHECK_ARG(params && grads && code && cb && d_wpack && d_code_bias && d_code && B > 0);
    if (B > MAX_IMG) return NIW_E_UNSUPP;
    PtrTable T; GradTable G;
    for (int i = 0; i < NIW_NVP_PARAM_PTRS; ++i) { NIW_CHECK_ARG(params[i] && grads[i]); T.p[i] = params[i]; G.p[i] = grads[i]; }
    cudaStream_t st = niw_stream(stream);
    NIW_CUDA(cudaMemsetAsync(d_code, 0, sizeof(float) * (size_t)B * DF, st));
    const size_t smem = sizeof(float) * (3 * (size_t)B * LDC + (size_t)B * KPC + 4 * HID);
    if (smem > 48 * 1024)
        NIW_CUDA(cudaFuncSetAttribute(nvp_pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    niw::note_launch(), nvp_pack_bwd_kernel<<<dim3(NB, PACK_SPLIT), PACK_THREADS, smem, st>>>(T, G, code, cb, B, d_wpack, d_code_bias, d_code);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_nvp_warp_fwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                                int B, int Pt, int idx_offset, int idx_split, int idx_jump, float* out, void* stream) {
    NIW_CHECK_ARG(wpack && code_bias && pts && out && B > 0 && Pt > 0 && idx_offset >= 0 && idx_split >= 0 && idx_jump >= 0);
    NIW_CHECK_ARG(((uintptr_t)wpack & 15) == 0);          // the weight images move as 128-bit vectors / bulk copies
    const IndexMap im{idx_offset, idx_split, idx_jump};
    const int64_t total = (int64_t)B * Pt;
    NIW_CUDA(cudaFuncSetAttribute(nvp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
    int64_t blocks = (total + FWD_WARPS - 1) / FWD_WARPS;
    const int64_t cap = (int64_t)niw_num_sms() * 2;
    if (blocks > cap) blocks = cap;
    niw::note_launch(), nvp_fwd_kernel<<<(unsigned)blocks, FWD_WARPS * 32, FWD_SMEM, niw_stream(stream)>>>(
        wpack, code_bias, pts, make_bands(alpha_ratio), im, B, Pt, out);
    NIW_LAUNCH_CHECK();
    return 0;
}

// pixel indices -> warped rays in one launch (shared centre row: the caller has checked that no centre row of the image's
// list is an annealed one).  pts / warped: [B, P+1, 3] = [grid rows ; the centre row]; ray, center: [B, P, 3].
extern "C" int niw_nvp_rays_fwd(const float* wpack, const float* code_bias, const float* intr, const float* pose_init,
                                const int64_t* ray_idx, int64_t idx_start, float alpha_ratio, int B, int P, int H, int W,
                                int idx_offset, int idx_split, int idx_jump, float* pts, float* warped, float* ray,
                                float* center, void* stream) {
    NIW_CHECK_ARG(wpack && code_bias && intr && pts && ray && center && B > 0 && P > 0 && H > 0 && W > 0 && idx_offset >= 0 &&
                  idx_split >= 0 && idx_jump >= 0 && ((uintptr_t)wpack & 15) == 0);
    const IndexMap im{idx_offset, idx_split, idx_jump};
    NIW_CUDA(cudaFuncSetAttribute(nvp_rays_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RAYS_SMEM));
    const dim3 grid((unsigned)((P + FWD_WARPS - 1) / FWD_WARPS), (unsigned)B);
    niw::note_launch(), nvp_rays_fwd_kernel<<<grid, (FWD_WARPS + 1) * 32, RAYS_SMEM, niw_stream(stream)>>>(
        wpack, code_bias, intr, pose_init, ray_idx, idx_start, make_bands(alpha_ratio), im, B, P, W, pts, warped, ray, center);
    NIW_LAUNCH_CHECK();
    return 0;
}

extern "C" int niw_nvp_warp_bwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                                int B, int Pt, int idx_offset, int idx_split, int idx_jump, const float* d_out,
                                float* d_wpack, float* d_code_bias, int max_ctas, void* stream) {
    NIW_CHECK_ARG(wpack && code_bias && pts && d_out && d_wpack && d_code_bias && B > 0 && Pt > 0 && idx_offset >= 0 &&
                  idx_split >= 0 && idx_jump >= 0 && max_ctas >= 0 && ((uintptr_t)wpack & 15) == 0);
    const IndexMap im{idx_offset, idx_split, idx_jump};
    cudaStream_t st = niw_stream(stream);
    NIW_CUDA(cudaMemsetAsync(d_wpack, 0, sizeof(float) * NB * BLOCK_FLOATS, st));
    NIW_CUDA(cudaMemsetAsync(d_code_bias, 0, sizeof(float) * NB * 2 * (size_t)B * HID, st));
    NIW_CUDA(cudaFuncSetAttribute(nvp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM));
    const int64_t total = (int64_t)B * Pt;
    // rounds (points per warp): the pass is a latency chain per warp (blocks in sequence, points in sequence), so as few as
    // one wave of CTAs allows
    // (max_ctas > 0: the caller keeps the other SMs for a kernel running concurrently on another stream)
    const int sms = max_ctas > 0 && max_ctas < niw_num_sms() ? max_ctas : niw_num_sms();
    const int64_t warps_max = (int64_t)sms * BWD_WARPS;
    int64_t rounds = (total + warps_max - 1) / warps_max;
    if (rounds < 1) rounds = 1;
    if (rounds > MAX_ROUNDS) rounds = MAX_ROUNDS;
    const int64_t blocks = (total + rounds * BWD_WARPS - 1) / (rounds * BWD_WARPS);
    niw::note_launch(), nvp_bwd_kernel<<<(unsigned)blocks, BWD_THREADS, BWD_SMEM, st>>>(
        wpack, code_bias, pts, make_bands(alpha_ratio), im, B, Pt, (int)rounds, d_out, d_wpack, d_code_bias);
    NIW_LAUNCH_CHECK();
    return 0;
}
