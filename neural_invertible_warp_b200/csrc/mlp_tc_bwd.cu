// Backward of the fused positional-encoding + NeRF MLP on tcgen05 (sm_100a).
// SURVEY.md section 8 rows a6+a7+a8 (autograd of reference model/nerf.py:416-456).
//
// Two kernels:
//
//  tc_dx_kernel   activation-gradient chain, one pass over the tiles, same warp roles and
//                 shared-memory operand images as the forward kernel: the TRANSPOSED weights are
//                 streamed (bulk async copies) through a ring, G_l = dL/dz_l of a 128-sample tile
//                 is the K-major A operand, the product lands in TMEM, and the epilogue threads
//                 (one per sample row) apply the saved ReLU bit masks, re-pack to BF16 as the next
//                 A operand and store the G image for the weight-gradient pass.  The first step
//                 (sigmoid/softplus derivatives and the 3->128 rgb layer) and the last one (the
//                 derivative of the positional encoding and the reduction of d_center / d_ray over
//                 a ray's samples) run on the CUDA cores in the same threads.
//
//  tc_dw_kernel   weight gradients dW_l = G_l^T . X_l as tensor-core GEMMs with K = samples: the
//                 saved X_l (forward) and G_l (dX pass) tile images are read as MN-major operands
//                 exactly as they lie in HBM (no transposition), each CTA owns one
//                 (layer, 128-output-row half, sample slice) and keeps its 128 x 256 fp32
//                 accumulator in TMEM across all its tiles; bias gradients and the two CUDA-core
//                 head layers ride along as N = 16 products against a small image holding
//                 [g_rgb_pre(3), g_sigma_pre, 1].  Slices are summed by tc_dw_reduce_kernel.
#include "tc_layout.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace niw {
namespace tc {

// ==========================================================================================
// dX pass
// ==========================================================================================

constexpr int BX_NSTAGE = 10;
#ifndef NIW_BX_GROUP
#define NIW_BX_GROUP 2
#endif
constexpr int BX_GROUP = NIW_BX_GROUP;           // weight chunks (2 MMAs each) per elected region of an issuer
static_assert(step_chunks(0) % BX_GROUP == 0 && step_chunks(2) % BX_GROUP == 0, "the issuer takes the weight chunks in groups");
constexpr int BX_ACT = 0;                                   // 2 x 64 KB G tiles
constexpr int BX_RING = BX_ACT + 2 * ACT_BYTES;
constexpr int BX_CONST = BX_RING + BX_NSTAGE * HSTAGE_BYTES; // W7 row 0 [256] + Wrgb1 [3][128]
constexpr int BX_CONST_FLOATS = WIDTH + 3 * RGBW + NBANDS;   // ... + band weights
constexpr int BX_BAR = BX_CONST + BX_CONST_FLOATS * 4;
constexpr int BX_TOTAL = BX_BAR + 512;
static_assert(BX_TOTAL <= 227 * 1024, "shared memory budget (dX pass)");

// reduce v over the 32 rows of a warp when they all belong to ray r (uniform), else per-thread atomics
__device__ __forceinline__ void ray_atomic_add3(float* dst, int64_t r, const float v[3], bool valid, bool uniform) {
    if (uniform) {
        float a = warp_sum(valid ? v[0] : 0.f), b = warp_sum(valid ? v[1] : 0.f), c = warp_sum(valid ? v[2] : 0.f);
        if ((threadIdx.x & 31) == 0 && r >= 0) { atomicAdd(dst + r * 3, a); atomicAdd(dst + r * 3 + 1, b); atomicAdd(dst + r * 3 + 2, c); }
    } else if (valid) {
        atomicAdd(dst + r * 3, v[0]); atomicAdd(dst + r * 3 + 1, v[1]); atomicAdd(dst + r * 3 + 2, v[2]);
    }
}

// CTA pairs (cta_group::2), same organisation as the forward kernel (mlp_tc.cu): slot s of CTA r holds tile 2*hq + r of
// the slot's tile pair hq, every MMA is M = 256 over the two tiles of a slot, each CTA stages half of every transposed-weight
// chunk, the two slots run one step apart, one issuer thread per slot in the leader CTA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
tc_dx_kernel(const uint8_t* __restrict__ bstream, const float* __restrict__ consts_g, const float* __restrict__ center,
             const float* __restrict__ ray, const float* __restrict__ depth, int64_t S, int N,
             const float* __restrict__ d_rgb, const float* __restrict__ d_sigma, const float* __restrict__ sig_pre,
             const float* __restrict__ rgb_keep, uint8_t* __restrict__ save, float* __restrict__ scratch,
             float* __restrict__ dP, float* __restrict__ d_center, float* __restrict__ d_ray) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BX_BAR);
    uint64_t* w_full = bars;                    // [2][BX_NSTAGE] (alternating trips round the ring, see mlp_tc.cu)
    uint64_t* w_empty = bars + 2 * BX_NSTAGE;   // [BX_NSTAGE]
    uint64_t* a_ready = bars + 3 * BX_NSTAGE;   // [2]  (leader)
    uint64_t* acc_full = a_ready + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    float* cst = reinterpret_cast<float*>(smem + BX_CONST);   // [0,256): W7 row 0; [256, 640): Wrgb1

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int64_t ntiles = (S + TILE - 1) / TILE;
    // Work unit = one SLOT of a CTA pair = one tile pair (tiles 2*hq + rank).  In round k slot s of pair p holds tile pair
    // (2k + s) * npairs + p, so the last round of a pair may use slot 0 only: 1 024 tiles over 74 pairs are 6.92 tile pairs per
    // CTA pair -- three rounds of two slots and one of a single slot (which runs its layers back to back, ~25 % faster than a
    // shared round) instead of four full rounds for half of the pairs and three for the rest.
    const int64_t nhq = (ntiles + 1) / 2;
    const int64_t pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    auto hq_of = [&](int64_t k, int s) { return (2 * k + s) * npairs + pair0; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * BX_NSTAGE; ++i) ptx::mbar_init(&w_full[i], rank == 0 ? 2 : 1);
        for (int i = 0; i < BX_NSTAGE; ++i) ptx::mbar_init(&w_empty[i], 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&a_ready[i], 2 * TILE / 32); ptx::mbar_init(&acc_full[i], 1); }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    if (warp == 3) {
        for (int i = lane; i < WIDTH + 3 * RGBW; i += 32) cst[i] = consts_g[C_W7R0 + i];
        if (lane < NBANDS) cst[WIDTH + 3 * RGBW + lane] = consts_g[C_BANDS + lane];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    Bands3 bw3; BandsV bwv;
    load_bands(cst + WIDTH + 3 * RGBW, bw3, bwv);

    if (warp == 0) {
        // ================= transposed-weight producer (this CTA's half of every chunk, once per slot) =================
        if (lane == 0) {
            uint32_t st = 0, cyc = 0;
            for (int64_t k = 0; hq_of(k, 0) < nhq; ++k) {
                const int nslots = hq_of(k, 1) < nhq ? 2 : 1;
                const uint8_t* ssrc = bstream;
                for (int s = 0; s < NSTEP; ++s) {
                    const uint32_t bytes = (uint32_t)(step_n(s) / 2) * CHUNK_K * 2;
                    for (int sl = 0; sl < nslots; ++sl) {
                        const uint8_t* src = ssrc + rank * bytes;
                        for (int c = 0; c < step_chunks(s); ++c) {
                            uint64_t* full = &w_full[(cyc & 1) * BX_NSTAGE + st];
                            ptx::mbar_wait(&w_empty[st], (cyc & 1) ^ 1);
                            ptx::mbar_arrive_expect_tx(full, bytes);
                            ptx::bulk_g2s(smem + BX_RING + st * HSTAGE_BYTES, src, bytes, full);
                            src += 2 * bytes;
                            if (++st == BX_NSTAGE) { st = 0; ++cyc; }
                        }
                    }
                    ssrc += (int64_t)step_chunks(s) * 2 * bytes;
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================= peer CTA: tell the leader when this CTA's half of a chunk has landed =================
        if (lane == 0) {
            uint32_t st = 0, cyc = 0;
            const uint32_t full0 = ptx::mapa(&w_full[0], 0);
            for (int64_t k = 0; hq_of(k, 0) < nhq; ++k)
                for (int s = 0; s < NSTEP; ++s)
                    for (int c = 0; c < (hq_of(k, 1) < nhq ? 2 : 1) * step_chunks(s); ++c) {
                        const uint32_t fb = (cyc & 1) * BX_NSTAGE + st;
                        ptx::mbar_wait(&w_full[fb], (cyc >> 1) & 1);
                        ptx::mbar_arrive_cluster(full0 + fb * 8);
                        if (++st == BX_NSTAGE) { st = 0; ++cyc; }
                    }
        }
    } else if ((warp == 1 || warp == 2) && rank == 0) {
        // ================= leader CTA: MMA issuers, one warp per slot (converged, one elected lane issues) =================
        {
            const int sl = warp - 1;
            uint32_t st = 0, cyc = 0, ready_ph = 0;
            auto skip = [&](int n) { st += n; while (st >= BX_NSTAGE) { st -= BX_NSTAGE; ++cyc; } };   // the other slot's chunks
            static_assert(2 * BX_NSTAGE >= WIDTH / CHUNK_K + 1, "ring too short for two skipping issuers (see mlp_tc.cu)");
            const uint32_t act_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + BX_ACT + sl * ACT_BYTES), KROW);
            const uint32_t ring_a = ptx::smem_addr(smem + BX_RING) >> 4;
            const uint32_t desc_hi = ptx::smem_desc_hi(128);
            const uint32_t tacc = tmem_base + sl * WIDTH;
            for (int64_t k = 0; hq_of(k, sl) < nhq; ++k) {
                const bool both = hq_of(k, 1) < nhq;                 // the other slot works in this round too
                for (int s = 0; s < NSTEP; ++s) {
                    const int hrows = step_n(s) / 2, nch = step_chunks(s);
                    const uint32_t idesc = ptx::idesc_bf16(2 * TILE, 2 * hrows, 0, 0);
                    const uint32_t b_lbo = (uint32_t)hrows << 16, b_kstep = (uint32_t)hrows * 2;
                    if (sl == 1) skip(nch);
                    ptx::mbar_wait_fast(&a_ready[sl], ready_ph);
                    ready_ph ^= 1;
                    ptx::tc_fence_after();
                    // BX_GROUP weight chunks (nch is 4 or 8) per elected region, as in the forward kernel
                    for (int c = 0; c < nch; c += BX_GROUP) {
                        uint32_t stg[BX_GROUP];
#pragma unroll
                        for (int i = 0; i < BX_GROUP; ++i) {
                            stg[i] = st;
                            ptx::mbar_wait(&w_full[(cyc & 1) * BX_NSTAGE + st], (cyc >> 1) & 1);
                            if (++st == BX_NSTAGE) { st = 0; ++cyc; }
                        }
                        ptx::tc_fence_after();
                        const uint32_t a_lo = act_lo + (uint32_t)c * (CHUNK_K / 8) * (KROW >> 4);
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int i = 0; i < BX_GROUP; ++i) {
                                const uint32_t b_lo = (ring_a + stg[i] * (HSTAGE_BYTES >> 4)) | b_lbo;
                                ptx::mma2_bf16_w(tacc, a_lo + (4 * i) * (KROW >> 4), desc_hi, b_lo, desc_hi, idesc, (c + i) != 0);
                                ptx::mma2_bf16_w(tacc, a_lo + (4 * i + 2) * (KROW >> 4), desc_hi, b_lo + b_kstep, desc_hi, idesc, 1u);
                                ptx::mma2_commit(&w_empty[stg[i]]);
                            }
                            if (c + BX_GROUP >= nch) ptx::mma2_commit(&acc_full[sl]);
                        }
                        __syncwarp();
                    }
                    if (sl == 0 && both) skip(nch);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups (one thread per sample row) =================
        const int slot = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) | lane;
        uint8_t* act = smem + BX_ACT + slot * ACT_BYTES;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + slot * WIDTH;
        float* scr = scratch + ((size_t)blockIdx.x * 2 + slot) * ENC3_PAD * TILE;
        const bool uniform = (N % 32) == 0;       // the 32 rows of a warp then share one ray
        const uint32_t ready_bar = ptx::mapa(&a_ready[slot], 0);     // the leader's barrier
        uint32_t full_uses = 0;
        for (int64_t k = 0; hq_of(k, slot) < nhq; ++k) {
            const int64_t tile = hq_of(k, slot) * 2 + rank;
            const int64_t g = tile * TILE + row;
            const bool valid = tile < ntiles && g < S;
            const int64_t r = valid ? g / N : -1;
            uint8_t* rec = tile < ntiles ? save + tile * SAVE_TILE_BYTES : nullptr;
            const uint32_t* mask = rec ? reinterpret_cast<const uint32_t*>(rec + SV_MASK) : nullptr;
            // ---- step "-1": derivatives of sigmoid / softplus, the 128->3 layer, G8 -> A tile ----
            float g3[3] = {0.f, 0.f, 0.f}, gs = 0.f;
            if (valid) {
#pragma unroll
                for (int c = 0; c < 3; ++c) { float y = rgb_keep[g * 3 + c]; g3[c] = d_rgb[g * 3 + c] * y * (1.f - y); }
                gs = d_sigma[g] * sigmoid_f(sig_pre[g]);             // softplus'(x) = sigmoid(x)
            }
            {
                // bias gradients of the two CUDA-core heads
                float b0 = warp_sum(g3[0]), b1 = warp_sum(g3[1]), b2 = warp_sum(g3[2]), b3 = warp_sum(gs);
                if (lane == 0) {
                    atomicAdd(dP + RGB1_B, b0); atomicAdd(dP + RGB1_B + 1, b1); atomicAdd(dP + RGB1_B + 2, b2);
                    atomicAdd(dP + feat_b_off(7), b3);
                }
            }
#pragma unroll 1
            for (int cc = 0; cc < RGBW / 32; ++cc) {
                const uint32_t bits = mask ? mask[(8 * MASK_WORDS + cc) * TILE + row] : 0u;
                const float4* w0 = reinterpret_cast<const float4*>(cst + WIDTH + cc * 32);
                uint32_t pk[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 x0 = w0[q], x1 = w0[RGBW / 4 + q], x2 = w0[2 * RGBW / 4 + q];
                    const float a = g3[0] * x0.x + g3[1] * x1.x + g3[2] * x2.x, b = g3[0] * x0.y + g3[1] * x1.y + g3[2] * x2.y;
                    const float c = g3[0] * x0.z + g3[1] * x1.z + g3[2] * x2.z, d = g3[0] * x0.w + g3[1] * x1.w + g3[2] * x2.w;
                    pk[2 * q] = ptx::pack_bf16(a, b) & ptx::relu_mask_expand(bits << q, 2 * q);
                    pk[2 * q + 1] = ptx::pack_bf16(c, d) & ptx::relu_mask_expand(bits << q, 2 * q + 1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                    *reinterpret_cast<uint4*>(act + (cc * 4 + q) * KROW + row * 16) = o;
                    if (rec) *reinterpret_cast<uint4*>(rec + SV_G8 + hbm_img_off(RGBW, row, cc * 4 + q)) = o;
                }
            }
            if (rec)   // small image: columns [g_rgb_pre(3), g_sigma_pre, 1, 0, 0, 0]
                *reinterpret_cast<uint4*>(rec + SV_SMALL + row * 16) =
                    make_uint4(ptx::pack_bf16(g3[0], g3[1]), ptx::pack_bf16(g3[2], gs), ptx::pack_bf16(valid ? 1.f : 0.f, 0.f), 0u);
            ptx::fence_proxy_async();
            ptx::warp_arrive_cluster(ready_bar);

            for (int s = 0; s < NSTEP; ++s, ++full_uses) {
                const int lo = step_out_layer(s);
                uint32_t mw[MASK_WORDS];
                if (lo >= 0) {   // the ReLU flags do not depend on the products: fetch them while the MMAs run
#pragma unroll
                    for (int cc = 0; cc < MASK_WORDS; ++cc) mw[cc] = mask ? mask[(lo * MASK_WORDS + cc) * TILE + row] : 0u;
                }
                ptx::mbar_wait_fast(&acc_full[slot], full_uses & 1);
                ptx::tc_fence_after();
                if (lo >= 0) {
                    // ---- hidden layers: (+ density rank-1 term) -> ReLU mask -> BF16 -> next A tile + G image ----
                    uint8_t* save_img = rec ? rec + SV_G + (int64_t)lo * ACT_BYTES : nullptr;
                    uint32_t v[2][32];
                    ptx::tmem_ld32(tacc, v[0]);
#pragma unroll
                    for (int cc = 0; cc < WIDTH / 32; ++cc) {
                        // TMEM loads run one chunk ahead of the arithmetic
                        ptx::tmem_ld_wait();
                        if (cc + 1 < WIDTH / 32) ptx::tmem_ld32(tacc + (cc + 1) * 32, v[(cc + 1) & 1]);
                        const uint32_t (&vc)[32] = v[cc & 1];
                        uint32_t pk[16];
                        if (s == 2) {   // dL/dh6 += g_sigma_pre * W7[0, :]   (density head, nerf.py:427)
                            const float4* w = reinterpret_cast<const float4*>(cst + cc * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 w4 = w[q];
                                pk[2 * q] = ptx::pack_bf16(__uint_as_float(vc[4 * q]) + gs * w4.x, __uint_as_float(vc[4 * q + 1]) + gs * w4.y);
                                pk[2 * q + 1] = ptx::pack_bf16(__uint_as_float(vc[4 * q + 2]) + gs * w4.z, __uint_as_float(vc[4 * q + 3]) + gs * w4.w);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16(__uint_as_float(vc[2 * j]), __uint_as_float(vc[2 * j + 1]));
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] &= ptx::relu_mask_expand(mw[cc] << (j >> 1), j);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 o = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                            *reinterpret_cast<uint4*>(act + (cc * 4 + q) * KROW + row * 16) = o;
                            if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(WIDTH, row, cc * 4 + q)) = o;
                        }
                    }
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::warp_arrive_cluster(ready_bar);
                } else if (s == 0) {
                    // ---- view branch: d(encoded view) -> d(unit view) -> d ray through normalize ----
                    uint32_t v[32];
                    ptx::tmem_ld32(tacc, v);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before();
                    ptx::warp_arrive_cluster(ready_bar);            // accumulator drained; A tile unchanged
                    float dr[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        float v3[3] = {ray[r * 3], ray[r * 3 + 1], ray[r * 3 + 2]};
                        float inv = 1.f / fmaxf(sqrtf(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]), 1e-12f);
                        float u[3] = {v3[0] * inv, v3[1] * inv, v3[2] * inv}, du[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float acc = __uint_as_float(v[c]);
#pragma unroll
                            for (int k = 0; k < LV; ++k) {
                                float f = (float)(1 << k) * PI_F, sn, cs;
                                sincos_reduced(u[c] * f, sn, cs);
                                acc += bwv.w[k] * f * (cs * __uint_as_float(v[3 + c * 2 * LV + k]) -
                                                       sn * __uint_as_float(v[3 + c * 2 * LV + LV + k]));
                            }
                            du[c] = acc;
                        }
                        float dot = u[0] * du[0] + u[1] * du[1] + u[2] * du[2];
#pragma unroll
                        for (int c = 0; c < 3; ++c) dr[c] = (du[c] - u[c] * dot) * inv;
                    }
                    ray_atomic_add3(d_ray, r, dr, valid, uniform);
                } else if (s == 5) {
                    // ---- skip connection: park G4 . W4[:, 256:319] (fp32) until step 10 ----
#pragma unroll
                    for (int cc = 0; cc < ENC3_PAD / 32; ++cc) {
                        uint32_t v[32];
                        ptx::tmem_ld32(tacc + cc * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) scr[(cc * 32 + j) * TILE + row] = __uint_as_float(v[j]);
                    }
                    ptx::tc_fence_before();
                    ptx::warp_arrive_cluster(ready_bar);
                } else {
                    // ---- s == 10: d(encoded position) -> d x -> d center, d ray ----
                    float ge[ENC3_PAD];
#pragma unroll
                    for (int cc = 0; cc < ENC3_PAD / 32; ++cc) {
                        uint32_t v[32];
                        ptx::tmem_ld32(tacc + cc * 32, v);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) ge[cc * 32 + j] = __uint_as_float(v[j]) + scr[(cc * 32 + j) * TILE + row];
                    }
                    ptx::tc_fence_before();   // TMEM reads done before the next pair overwrites the accumulator
                    float dc[3] = {0.f, 0.f, 0.f}, dv[3] = {0.f, 0.f, 0.f};
                    if (valid) {
                        const float d = depth[g];
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float x = __fadd_rn(center[r * 3 + c], __fmul_rn(ray[r * 3 + c], d));
                            float acc = ge[c];
#pragma unroll
                            for (int k = 0; k < L3; ++k) {
                                float f = (float)(1 << k) * PI_F, sn, cs;
                                sincos_reduced(x * f, sn, cs);
                                acc += bw3.w[k] * f * (cs * ge[3 + c * 2 * L3 + k] - sn * ge[3 + c * 2 * L3 + L3 + k]);
                            }
                            dc[c] = acc; dv[c] = acc * d;
                        }
                    }
                    ray_atomic_add3(d_center, r, dc, valid, uniform);
                    ray_atomic_add3(d_ray, r, dv, valid, uniform);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // neither CTA leaves while its peer may still touch its shared memory / TMEM
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

// ==========================================================================================
// dW pass
// ==========================================================================================
// One CTA per (unit, sample slice).  A unit is one weight block dW = G^T . X with all its output rows:
// G (up to 256 features = two M = 128 blocks) and X (N <= 256 features) are streamed as 64-sample half
// images (contiguous in the tile record, MN-major operands with K = samples) through a 3-stage ring, and the
// two 128 x N fp32 accumulators fill TMEM for the whole slice, so every tile costs 128 KB of HBM reads for
// 16.8 MFLOP.  The thin products (bias gradients = G^T . 1, density row = h6^T . g_sigma, rgb1 weights =
// hr^T . g_rgb) have no TMEM columns left: warps 4-7 compute them from the same staged half images with
// warp-level mma.sync (m16n8k16, A = small image^T via ldmatrix.trans), accumulating in registers.

constexpr int BW_NSTAGE = 3;
constexpr int BW_A = 0;                                  // 32 KB: 64 samples x up to 256 features (G [+ hr])
constexpr int BW_B = BW_A + ACT_BYTES / 2;               // 32 KB: 64 samples x up to 256 features (X)
constexpr int BW_B2 = BW_B + ACT_BYTES / 2;              // 4 KB:  64 samples x 32 columns (encoded view)
constexpr int BW_S = BW_B2 + VENC_BYTES / 2;             // 1 KB:  64 samples x 8 columns (small image)
constexpr int BW_STAGE = BW_S + SMALL_BYTES / 2;         // 70656
constexpr int BW_BAR = BW_NSTAGE * BW_STAGE;
constexpr int BW_TOTAL = BW_BAR + 64;
static_assert(BW_TOTAL <= 227 * 1024, "shared memory budget (dW pass)");
static_assert(4 * 32 * 33 * 4 <= BW_STAGE, "epilogue transpose buffers fit in stage 0");

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr) : "memory");
}
// D(16x8, only rows 0-7 used) += A(16x16, rows 8-15 zero) . B(16x8)
// (d[2], d[3] are the unused rows 8-15: they stay zero, but every accumulator needs its own pair -- shared
// dummies would chain all MMAs through one register dependency)
__device__ __forceinline__ void mma_16816_top(float (&d)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(256, 1)
tc_dw_kernel(const uint8_t* __restrict__ save, int64_t ntiles, DwPlan plan, float* __restrict__ partial, int direct,
             unsigned long long* __restrict__ dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    unsigned long long t_start = 0;
    if (dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + BW_BAR);   // [BW_NSTAGE]
    uint64_t* empty = full + BW_NSTAGE;                            // [BW_NSTAGE]
    uint64_t* done = empty + BW_NSTAGE;                            // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // which unit / slice is this CTA?
    int ui = 0;
    for (int i = 0; i < plan.n_units; ++i) if ((int)blockIdx.x >= plan.u[i].first_cta) ui = i;
    const DwUnit& U = plan.u[ui];
    const int slice = (int)blockIdx.x - U.first_cta;
    // slice s owns tiles s, s + n_slices, ... and walks them from the highest down: the dX pass has just written the
    // G images in ascending tile order, so every CTA starts on the records most likely to be still in L2
    const int64_t my_tiles = slice < ntiles ? (ntiles - slice + U.n_slices - 1) / U.n_slices : 0;
    const int64_t nstages = 2 * my_tiles;      // 64-sample halves

    if (threadIdx.x == 0) {
        for (int i = 0; i < BW_NSTAGE; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 5); }
        ptx::mbar_init(done, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc(tmem_slot, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t stage_bytes = (uint32_t)(U.a_half + U.a2_half + U.b_half + U.b2_half + SMALL_BYTES / 2);
            for (int64_t it = 0; it < nstages; ++it) {
                const uint32_t st = (uint32_t)(it % BW_NSTAGE), ph = (uint32_t)(it / BW_NSTAGE) & 1;
                const uint8_t* rec = save + (slice + (my_tiles - 1 - (it >> 1)) * U.n_slices) * SAVE_TILE_BYTES;
                const int half = (int)(it & 1);
                uint8_t* dst = smem + st * BW_STAGE;
                ptx::mbar_wait(&empty[st], ph ^ 1);
                ptx::mbar_arrive_expect_tx(&full[st], stage_bytes);
                ptx::bulk_g2s(dst + BW_A, rec + U.a_off + half * U.a_half, (uint32_t)U.a_half, &full[st]);
                if (U.a2_half) ptx::bulk_g2s(dst + BW_A + HR_BYTES / 2, rec + U.a2_off + half * U.a2_half, (uint32_t)U.a2_half, &full[st]);
                if (U.b_half) ptx::bulk_g2s(dst + BW_B, rec + U.b_off + half * U.b_half, (uint32_t)U.b_half, &full[st]);
                if (U.b2_half) ptx::bulk_g2s(dst + BW_B2, rec + U.b2_off + half * U.b2_half, (uint32_t)U.b2_half, &full[st]);
                ptx::bulk_g2s(dst + BW_S, rec + SV_SMALL + half * (SMALL_BYTES / 2), SMALL_BYTES / 2, &full[st]);
            }
        }
    } else if (warp == 1) {
        // MMA issuer: the warp stays converged (loop state and descriptor words uniform), one elected lane issues a whole
        // stage -- a single-thread issuer computes every descriptor in per-thread registers and pays register-to-uniform
        // moves per MMA, which made the ISSUE (not the memory system) the limit of a dW CTA (~40 GB/s)
        const uint32_t idesc = ptx::idesc_bf16(TILE, U.n_main > 0 ? U.n_main : 16, 1, 1);
        const uint32_t idesc2 = ptx::idesc_bf16(TILE, U.n2 > 0 ? U.n2 : 16, 1, 1);
        const uint32_t desc_hi = ptx::smem_desc_hi(HROW);          // MN-major operands, K = samples: LBO = 128 (8 samples), SBO = HROW
        const int m_halves = U.m_halves;
        const bool second = U.n2 > 0;
        for (int64_t it = 0; it < nstages; ++it) {
            const uint32_t st = (uint32_t)(it % BW_NSTAGE), ph = (uint32_t)(it / BW_NSTAGE) & 1;
            ptx::mbar_wait(&full[st], ph);
            ptx::tc_fence_after();
            const uint32_t base = ptx::smem_addr(smem + st * BW_STAGE);
            const uint32_t a_lo = ptx::smem_desc_lo(base + BW_A, 128), b_lo = ptx::smem_desc_lo(base + BW_B, 128);
            const uint32_t b2_lo = ptx::smem_desc_lo(base + BW_B2, 128);
            const uint32_t first = it != 0;
            if (ptx::elect_one()) {
#pragma unroll
                for (int ks = 0; ks < HALF / 16; ++ks) {
                    // 16 samples = 256 B along a feature group
                    const uint32_t k = (uint32_t)(ks * 256) >> 4, acc = first | (uint32_t)(ks != 0);
                    ptx::mma_bf16_w(tmem_base, a_lo + k, desc_hi, b_lo + k, desc_hi, idesc, acc);
                    if (m_halves > 1) ptx::mma_bf16_w(tmem_base + WIDTH, a_lo + ((16 * HROW) >> 4) + k, desc_hi, b_lo + k, desc_hi, idesc, acc);
                    if (second) ptx::mma_bf16_w(tmem_base + WIDTH, a_lo + k, desc_hi, b2_lo + k, desc_hi, idesc2, acc);
                }
                ptx::mma_commit(&empty[st]);
            }
            __syncwarp();
        }
        if (ptx::elect_one()) ptx::mma_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int wq = warp & 3;
        // ---- thin products on the warp-level tensor path while the tiles stream ----
        // this warp owns feature groups [wq*g, (wq+1)*g) of each side image, g = groups / 4 (8 or 4)
        const int g0 = U.side[0].kind ? U.side[0].groups / 4 : 0, g1 = U.side[1].kind ? U.side[1].groups / 4 : 0;
        float acc0[8][4], acc1[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc0[i][q] = 0.f; acc1[i][q] = 0.f; }
        for (int64_t it = 0; it < nstages; ++it) {
            const uint32_t st = (uint32_t)(it % BW_NSTAGE), ph = (uint32_t)(it / BW_NSTAGE) & 1;
            ptx::mbar_wait(&full[st], ph);
            if (g0 | g1) {
                const uint32_t base = ptx::smem_addr(smem + st * BW_STAGE);
                uint32_t sa[2][4];
                ldsm_x4_trans(sa[0], base + BW_S + lane * 16);
                ldsm_x4_trans(sa[1], base + BW_S + 512 + lane * 16);
                // four feature groups at a time, k outermost: consecutive MMAs hit different accumulators
                // (a chain of dependent mma.sync costs ~33 cycles per instruction)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int gq = j == 0 ? g0 : g1;
                    const uint32_t img = base + U.side[j].smem_off + lane * 16;
#pragma unroll
                    for (int i0 = 0; i0 < 8; i0 += 4) {
                        if (i0 < gq) {
                            uint32_t xb[4][2][4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                ldsm_x4_trans(xb[i][0], img + (wq * gq + i0 + i) * HROW);
                                ldsm_x4_trans(xb[i][1], img + (wq * gq + i0 + i) * HROW + 512);
                            }
#pragma unroll
                            for (int k = 0; k < 4; ++k)
#pragma unroll
                                for (int i = 0; i < 4; ++i)
                                    mma_16816_top(j == 0 ? acc0[i0 + i] : acc1[i0 + i], sa[k >> 1][(k & 1) * 2], sa[k >> 1][(k & 1) * 2 + 1],
                                                  xb[i][k >> 1][(k & 1) * 2], xb[i][k >> 1][(k & 1) * 2 + 1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[st]);
        }
        // ---- epilogue: TMEM -> (per-warp transpose in shared memory) -> coalesced partial-sum rows ----
        ptx::mbar_wait(done, 0);
        ptx::tc_fence_after();
        asm volatile("bar.sync 1, 128;" ::: "memory");   // all four warps are done with the stage buffers
        // direct: `partial` is the gradient vector itself and every slice adds its product with reductions at the L2
        // (RED.ADD.F32, warp-coalesced) -- no partial rows to clear before and no reduction pass after this kernel
        float* out = direct ? partial : partial + (size_t)slice * NPARAMS;
        auto put = [&](int64_t i, float v) { if (direct) atomicAdd(out + i, v); else out[i] = v; };
        if (my_tiles > 0) {
            float* tr = reinterpret_cast<float*>(smem) + wq * (32 * 33);   // stage buffers are idle now
            for (int h = 0; h < U.m_halves + (U.n2 > 0 ? 1 : 0); ++h) {
                // h < m_halves: main product of M block h;  h == m_halves (= 1): second X image of M block 0
                const bool second = h >= U.m_halves;
                const int n = second ? U.n2 : U.n_main, ncols = second ? U.ncols2 : U.ncols, col0 = second ? U.col0_2 : U.col0;
                const int frow = second ? 0 : h * 128;
                const uint32_t tacc = tmem_base + ((uint32_t)(wq * 32) << 16) + h * WIDTH;
                for (int c0 = 0; c0 < n; c0 += 32) {
                    uint32_t v[32];
                    ptx::tmem_ld32(tacc + c0, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const int col = c0 + lane;
                    if (col < ncols) {
                        for (int rr = 0; rr < 32; ++rr)
                            put(U.w_base + (int64_t)(frow + wq * 32 + rr) * U.ld + col0 + col, tr[rr * 33 + lane]);
                    }
                    __syncwarp();
                }
            }
            // thin products: this lane holds row m = lane / 4 of small^T . X for features 8*group + 2*(lane % 4) + {0, 1}
            const int m = lane >> 2, n0 = (lane & 3) * 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const DwSide& sd = U.side[j];
                const int gq = j == 0 ? g0 : g1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < gq) {
                        const int f = (wq * gq + i) * 8 + n0;
                        const float x0 = j == 0 ? acc0[i][0] : acc1[i][0], x1 = j == 0 ? acc0[i][1] : acc1[i][1];
                        if ((sd.kind == 1 && m == 4) || (sd.kind == 3 && m == 3)) { put(sd.base + f, x0); put(sd.base + f + 1, x1); }
                        if (sd.kind == 2 && m < 3) { put(sd.base + m * RGBW + f, x0); put(sd.base + m * RGBW + f + 1, x1); }
                    }
                }
            }
        }
        ptx::tc_fence_before();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
    if (dbg && threadIdx.x == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        dbg[blockIdx.x * 2] = t_start; dbg[blockIdx.x * 2 + 1] = t_end;
    }
}

// dP[i] += sum over slices of partial[s][i]
__global__ void tc_dw_reduce_kernel(const float* __restrict__ partial, int n_slices, float* __restrict__ dP) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPARAMS) return;
    float acc = 0.f;
    for (int s = 0; s < n_slices; ++s) acc += partial[(size_t)s * NPARAMS + i];
    dP[i] += acc;
}

// ---- host: work plan of the dW pass ----
static DwPlan make_plan(int64_t ntiles, int n_sms) {
    DwPlan p;
    int n = 0;
    auto add = [&](int64_t a_off, int a_cols, int64_t b_off, int b_cols, int ld, int col0, int ncols, int64_t w_base) -> DwUnit& {
        DwUnit& u = p.u[n++];
        u = DwUnit{};
        u.a_off = (int32_t)a_off; u.a_half = a_cols / 8 * HROW; u.m_halves = a_cols / 128;
        u.b_off = (int32_t)b_off; u.b_half = b_cols / 8 * HROW; u.n_main = b_cols;
        u.ld = ld; u.col0 = col0; u.ncols = ncols; u.w_base = w_base;
        u.first_cta = 0; u.n_slices = 1;
        return u;
    };
    auto side = [](DwUnit& u, int j, int smem_off, int cols, int kind, int64_t base) {
        u.side[j].smem_off = smem_off; u.side[j].groups = cols / 8; u.side[j].kind = kind; u.side[j].base = base;
    };
    for (int l = 0; l < NFEAT; ++l) {
        const int ro = layer_rowoff(l);
        const int64_t a_off = SV_G + (int64_t)l * ACT_BYTES;
        const int64_t w_base = feat_w_off(l) + (int64_t)ro * feat_in(l), b_base = feat_b_off(l) + ro;
        if (l == 0) {
            side(add(a_off, WIDTH, SV_ENC, ENC3_PAD, feat_in(l), 0, ENC3, w_base), 0, BW_A, WIDTH, 1, b_base);
        } else {
            DwUnit& u = add(a_off, WIDTH, SV_H + (int64_t)(l - 1) * ACT_BYTES, WIDTH, feat_in(l), 0, WIDTH, w_base);
            side(u, 0, BW_A, WIDTH, 1, b_base);
            if (l == NFEAT - 1) side(u, 1, BW_B, WIDTH, 3, feat_w_off(l));                       // + density row from h6
            if (l == SKIP) add(a_off, WIDTH, SV_ENC, ENC3_PAD, feat_in(l), WIDTH, ENC3, w_base);
        }
    }
    {   // rgb0 (h7 and encoded-view columns) + rgb1 weights from the hr image staged behind G8
        DwUnit& u = add(SV_G8, RGBW, SV_H + 7 * (int64_t)ACT_BYTES, WIDTH, WIDTH + ENCV, 0, WIDTH, RGB0_W);
        u.b2_off = (int32_t)SV_VENC; u.b2_half = ENCV_PAD / 8 * HROW; u.n2 = ENCV_PAD; u.col0_2 = WIDTH; u.ncols2 = ENCV;
        u.a2_off = (int32_t)SV_HR; u.a2_half = HR_BYTES / 2;
        side(u, 0, BW_A, RGBW, 1, RGB0_B);
        side(u, 1, BW_A + HR_BYTES / 2, RGBW, 2, RGB1_W);
    }
    p.n_units = n;
    // slices proportional to bytes streamed per tile, capped by the tile count and PARTIAL_SLICES
    auto bytes = [&](int i) { return 2.0 * (p.u[i].a_half + p.u[i].a2_half + p.u[i].b_half + p.u[i].b2_half) + SMALL_BYTES; };
    double total = 0;
    for (int i = 0; i < n; ++i) total += bytes(i);
    int budget = n_sms > n ? n_sms : n;
    int used = 0;
    for (int i = 0; i < n; ++i) {
        int s = (int)(bytes(i) / total * budget);
        if (s < 1) s = 1;
        if (s > PARTIAL_SLICES) s = PARTIAL_SLICES;
        if (s > ntiles) s = (int)ntiles;
        p.u[i].n_slices = s; used += s;
    }
    for (int pass = 0; pass < 4 && used < budget; ++pass)      // hand out the remainder, largest units first
        for (int i = 0; i < n && used < budget; ++i)
            if (p.u[i].b_half == ACT_BYTES / 2 && p.u[i].n_slices < PARTIAL_SLICES && p.u[i].n_slices < ntiles) { ++p.u[i].n_slices; ++used; }
    int cta = 0, mx = 1;
    for (int i = 0; i < n; ++i) { p.u[i].first_cta = cta; cta += p.u[i].n_slices; if (p.u[i].n_slices > mx) mx = p.u[i].n_slices; }
    p.n_ctas = cta; p.max_slices = mx;
    return p;
}

}  // namespace tc

// activation-gradient chain: G images into the tile records, d_center / d_ray, the CUDA-core heads' bias gradients
int tc_bwd_dx(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
              void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma, float* dP,
              float* d_center, float* d_ray, cudaStream_t st) {
    using namespace tc;
    (void)P;
    const int64_t S = R * (int64_t)N;
    Workspace w = carve(ws, S, true);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    const int64_t ntiles = (S + TILE - 1) / TILE;
    if (d_ray == d_center + R * 3) {          // adjacent buffers (functional.py allocates them so): one memset node
        NIW_CUDA(cudaMemsetAsync(d_center, 0, sizeof(float) * R * 6, st));
    } else {
        NIW_CUDA(cudaMemsetAsync(d_center, 0, sizeof(float) * R * 3, st));
        NIW_CUDA(cudaMemsetAsync(d_ray, 0, sizeof(float) * R * 3, st));
    }
    NIW_CUDA(cudaFuncSetAttribute(tc_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BX_TOTAL));
    const int64_t nhq = (ntiles + 1) / 2;         // tile pairs: one per slot of a CTA pair and round
    int64_t pairs = niw_num_sms() / 2;
    if (pairs > nhq) pairs = nhq;
    const int grid = (int)(2 * (pairs < 1 ? 1 : pairs));
    niw::note_launch(), tc_dx_kernel<<<grid, 384, BX_TOTAL, st>>>(w.bstream, w.consts, center, ray, depth, S, N, d_rgb, d_sigma,
                                                                 w.sig_pre, w.rgb_keep, w.save, w.scratch, dP, d_center, d_ray);
    NIW_LAUNCH_CHECK();
    return 0;
}

// weight gradients from the tile records (after tc_bwd_dx on the same workspace): dP += G^T X.  `max_ctas` (0 = one per
// SM): the pass is HBM-bound, so fewer CTAs carry it at nearly the same speed and leave the other SMs to kernels running
// concurrently on another stream (the pose / warp backward of the training step)
int tc_bwd_dw(int64_t R, int N, void* ws, size_t ws_bytes, float* dP, int max_ctas, cudaStream_t st) {
    using namespace tc;
    const int64_t S = R * (int64_t)N;
    Workspace w = carve(ws, S, true);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    const int64_t ntiles = (S + TILE - 1) / TILE;
    // NIW_DW_CTAS overrides the CTA budget (tuning knob)
    static const int env_ctas = getenv("NIW_DW_CTAS") ? atoi(getenv("NIW_DW_CTAS")) : 0;
    int budget = env_ctas > 0 ? env_ctas : max_ctas;
    if (budget <= 0 || budget > niw_num_sms()) budget = niw_num_sms();
    DwPlan plan = make_plan(ntiles, budget);
    // NIW_DW_PARTIALS=1: the first form of the cross-slice sum (each slice stores a row of partial sums, cleared before and
    // reduced after the kernel: 30 MB written, 30 MB cleared and 30 MB read again, 16 us of the C2 step on its critical path)
    static const bool direct = !(getenv("NIW_DW_PARTIALS") && atoi(getenv("NIW_DW_PARTIALS")) != 0);
    if (!direct) NIW_CUDA(cudaMemsetAsync(w.partial, 0, sizeof(float) * (size_t)plan.max_slices * NPARAMS, st));
    NIW_CUDA(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_TOTAL));
    // NIW_DW_DEBUG=1: per-CTA start / end times of the dW pass on stderr (synchronises; diagnostics only)
    static const bool dw_debug = getenv("NIW_DW_DEBUG") != nullptr;
    unsigned long long* dbg = nullptr;
    if (dw_debug) NIW_CUDA(cudaMalloc(&dbg, sizeof(unsigned long long) * 2 * plan.n_ctas));
    niw::note_launch(), tc_dw_kernel<<<plan.n_ctas, 256, BW_TOTAL, st>>>(w.save, ntiles, plan, direct ? dP : w.partial, direct ? 1 : 0, dbg);
    NIW_LAUNCH_CHECK();
    if (dbg) {
        std::vector<unsigned long long> h(2 * plan.n_ctas);
        NIW_CUDA(cudaStreamSynchronize(st));
        NIW_CUDA(cudaMemcpy(h.data(), dbg, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        unsigned long long t0 = h[0];
        for (int i = 0; i < plan.n_ctas; ++i) if (h[2 * i] < t0) t0 = h[2 * i];
        for (int u = 0; u < plan.n_units; ++u)
            for (int sl = 0; sl < plan.u[u].n_slices; ++sl) {
                const int c = plan.u[u].first_cta + sl;
                fprintf(stderr, "dw unit %2d slice %2d/%2d start %7.1f us  dur %7.1f us\n", u, sl, plan.u[u].n_slices,
                        (h[2 * c] - t0) * 1e-3, (h[2 * c + 1] - h[2 * c]) * 1e-3);
            }
    }
    if (!direct) {
        niw::note_launch(), tc_dw_reduce_kernel<<<niw_blocks(NPARAMS, 256), 256, 0, st>>>(w.partial, plan.max_slices, dP);
        NIW_LAUNCH_CHECK();
    }
    return 0;
}

int tc_bwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N,
           void* ws, size_t ws_bytes, const float* d_rgb, const float* d_sigma, float* dP,
           float* d_center, float* d_ray, cudaStream_t st) {
    int e = tc_bwd_dx(P, center, ray, depth, R, N, ws, ws_bytes, d_rgb, d_sigma, dP, d_center, d_ray, st);
    if (e) return e;
    return tc_bwd_dw(R, N, ws, ws_bytes, dP, 0, st);
}

}  // namespace niw
