// Streaming ("layer-outer") activation-gradient chain of the fused NeRF MLP backward (tcgen05, sm_100a).
// SURVEY.md section 8 rows a6+a7+a8 (autograd of reference model/nerf.py:416-456).
//
// tc_dx_kernel (mlp_tc_bwd.cu) walks a tile through all eleven steps of the chain with the gradient tile resident in shared
// memory; every step of a tile waits for the previous one, two such chains fit one SM, and the tensor pipe idles ~60 % of the
// time.  Here the loops are interchanged for steps 2..10 (the 256-wide layers; the head steps -1, 0, 1 stay with tc_dx_kernel
// in its "head" mode): for each step the CTA pair's half of the transposed weights is RESIDENT in shared memory and the pair's
// tiles stream through it -- the gradient image G_in of a tile comes back from the tile record with one bulk copy (written a
// few microseconds earlier by this very CTA: L2 hits), D = G_in . W lands in TMEM, the epilogue warps apply the ReLU masks and
// store G_out to the record.  Consecutive MMAs belong to different tiles, so nothing of a tile's chain is on the critical
// path of the tensor pipe.
//
// Warp roles (384 threads, CTA pairs, tcgen05 cta_group::2, M = 256 = the pair's two 128-sample tiles):
//   warp 0      gradient-image producer (bulk async copies into two 64 KB buffers), gated on the epilogue that wrote the image
//   warp 1      leader: MMA issuer;  peer: relays "my image has landed"
//   warp 2      TMEM allocator;  peer: relays "my half of the weights has landed"
//   warp 3      constants, then weight producer (once per step)
//   warps 4-7   epilogue of accumulator 0,  warps 8-11 epilogue of accumulator 1 (one thread per sample row)
#include "tc_layout.cuh"
#include <cstdio>
#include <cstdlib>

namespace niw {
namespace tc {

constexpr int FB_NPH = 9;                                  // steps 2 .. 10 of the chain (tc_layout.cuh step table)
__host__ __device__ constexpr int fb_step(int ph) { return ph + 2; }
__host__ __device__ constexpr int fb_in_layer(int s) { return s <= 4 ? 9 - s : (s <= 6 ? 4 : 10 - s); }   // G image the step multiplies

constexpr int FB_GX = 0;                                   // 2 x 64 KB: G_in images (K-major A operand, [32 groups][128 rows][8])
constexpr int FB_W = FB_GX + 2 * ACT_BYTES;                // this CTA's half of the step's transposed weights
constexpr int FB_W_BYTES = (WIDTH / 2) * WIDTH * 2;        // 65536
constexpr int FB_CONST = FB_W + FB_W_BYTES;                // W7 row 0 [256] + band weights
constexpr int FB_CONST_FLOATS = WIDTH + NBANDS;
constexpr int FB_BAR = FB_CONST + FB_CONST_FLOATS * 4;
constexpr int FB_TOTAL = FB_BAR + 256;
static_assert(FB_TOTAL <= 227 * 1024, "shared memory budget (streaming dX)");

__device__ __forceinline__ void fb_fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t fb_ld_volatile_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_addr(p)) : "memory");
    return v;
}
// reduce v over the 32 rows of a warp when they all belong to ray r (uniform), else per-thread atomics
__device__ __forceinline__ void fb_ray_atomic_add3(float* dst, int64_t r, const float v[3], bool valid, bool uniform) {
    if (uniform) {
        float a = warp_sum(valid ? v[0] : 0.f), b = warp_sum(valid ? v[1] : 0.f), c = warp_sum(valid ? v[2] : 0.f);
        if ((threadIdx.x & 31) == 0 && r >= 0) { atomicAdd(dst + r * 3, a); atomicAdd(dst + r * 3 + 1, b); atomicAdd(dst + r * 3 + 2, c); }
    } else if (valid) {
        atomicAdd(dst + r * 3, v[0]); atomicAdd(dst + r * 3 + 1, v[1]); atomicAdd(dst + r * 3 + 2, v[2]);
    }
}

// Work items of a CTA pair: (phase ph, tile pair i) in phase-major order, seq = ph * T + i; item seq uses accumulator /
// epilogue warpgroup / image buffer seq & 1.  G images live UNSPLIT in the records here ([32 groups][128 rows][8], one bulk copy).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
tc_dx_stream_kernel(const uint8_t* __restrict__ bstream, const float* __restrict__ consts_g, const float* __restrict__ center,
                    const float* __restrict__ ray, const float* __restrict__ depth, int64_t S, int N,
                    const float* __restrict__ d_sigma, const float* __restrict__ sig_pre, uint8_t* __restrict__ save,
                    float* __restrict__ park, float* __restrict__ d_center, float* __restrict__ d_ray) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FB_BAR);
    uint64_t* gx_full = bars;                     // [2] this CTA's image has landed (leader: and the peer's)
    uint64_t* gx_empty = gx_full + 2;             // [2] the MMAs reading the buffer have completed (both CTAs)
    uint64_t* w_full = gx_empty + 2;              // [1] the step's weights have landed (leader: in both CTAs)
    uint64_t* w_empty = w_full + 1;               // [1] the step's last MMA has completed (both CTAs)
    uint64_t* acc_full = w_empty + 1;             // [2] a tile pair has been accumulated (both CTAs)
    uint64_t* acc_empty = acc_full + 2;           // [2] (leader) both CTAs have drained the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint32_t* done = tmem_slot + 2;               // [2] epilogue warps x items published, per accumulator
    float* cst = reinterpret_cast<float*>(smem + FB_CONST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int64_t npairs_t = ((S + TILE - 1) / TILE + 1) / 2;
    const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
    const int64_t T = pair0 < npairs_t ? (npairs_t - pair0 + pair_step - 1) / pair_step : 0;   // tile pairs of this CTA pair
    auto tile_of = [&](int64_t i) { return 2 * (pair0 + i * pair_step) + rank; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&gx_full[i], rank == 0 ? 2 : 1); ptx::mbar_init(&gx_empty[i], 1);
            ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 2 * TILE / 32);
        }
        ptx::mbar_init(w_full, rank == 0 ? 2 : 1);
        ptx::mbar_init(w_empty, 1);
        done[0] = 0; done[1] = 0;
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    if (warp == 3) {
        for (int i = lane; i < WIDTH; i += 32) cst[i] = consts_g[C_W7R0 + i];
        if (lane < NBANDS) cst[WIDTH + lane] = consts_g[C_BANDS + lane];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= gradient-image producer =================
        if (lane == 0) {
            for (int ph = 0; ph < FB_NPH; ++ph) {
                const int64_t img = SV_G + (int64_t)fb_in_layer(fb_step(ph)) * ACT_BYTES;
                for (int64_t i = 0; i < T; ++i) {
                    const int64_t seq = (int64_t)ph * T + i;
                    const uint32_t b = (uint32_t)(seq & 1), use = (uint32_t)(seq >> 1);
                    if (ph > 0) {      // the image was written by the previous step's epilogue of this tile (item seq - T)
                        const int64_t dep = seq - T;
                        const uint32_t need = 4u * (uint32_t)((dep >> 1) + 1);
                        while (fb_ld_volatile_shared(done + (dep & 1)) < need) __nanosleep(20);
                        __threadfence();
                        fb_fence_proxy_async_global();
                    }
                    ptx::mbar_wait(&gx_empty[b], (use & 1) ^ 1);
                    ptx::mbar_arrive_expect_tx(&gx_full[b], ACT_BYTES);
                    ptx::bulk_g2s(smem + FB_GX + b * ACT_BYTES, save + tile_of(i) * SAVE_TILE_BYTES + img, ACT_BYTES, &gx_full[b]);
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================= peer: relay "image landed" to the leader =================
        if (lane == 0) {
            const uint32_t full0 = ptx::mapa(&gx_full[0], 0);
            for (int64_t seq = 0; seq < (int64_t)FB_NPH * T; ++seq) {
                ptx::mbar_wait(&gx_full[seq & 1], (uint32_t)(seq >> 1) & 1);
                ptx::mbar_arrive_cluster(full0 + (uint32_t)(seq & 1) * 8);
            }
        }
    } else if (warp == 2 && rank != 0) {
        // ================= peer: relay "weights landed" to the leader =================
        if (lane == 0 && T > 0) {
            const uint32_t wfull0 = ptx::mapa(w_full, 0);
            for (int ph = 0; ph < FB_NPH; ++ph) {
                ptx::mbar_wait(w_full, ph & 1);
                ptx::mbar_arrive_cluster(wfull0);
            }
        }
    } else if (warp == 3) {
        // ================= weight producer: this CTA's N half of every K = 32 chunk of the step, resident for the step =================
        if (lane == 0 && T > 0) {
            for (int ph = 0; ph < FB_NPH; ++ph) {
                const int s = fb_step(ph);
                const uint32_t cb = (uint32_t)(step_n(s) / 2) * CHUNK_K * 2;
                const uint8_t* src = bstream + bstream_off(s) + rank * cb;
                ptx::mbar_wait(w_empty, (ph & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(w_full, (uint32_t)step_chunks(s) * cb);
                for (int c = 0; c < step_chunks(s); ++c) ptx::bulk_g2s(smem + FB_W + c * cb, src + (int64_t)c * 2 * cb, cb, w_full);
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= leader: MMA issuer (warp converged, one elected region per tile pair) =================
        uint32_t acc_uses[2] = {0u, 0u};
        const uint32_t desc_hi = ptx::smem_desc_hi(128);
        const uint32_t w_a = ptx::smem_addr(smem + FB_W) >> 4;
        for (int ph = 0; ph < FB_NPH; ++ph) {
            const int s = fb_step(ph), hrows = step_n(s) / 2, nk = step_k(s) / 16;
            const uint32_t idesc = ptx::idesc_bf16(2 * TILE, 2 * hrows, 0, 0);
            const uint32_t b_lbo = (uint32_t)hrows << 16, b_kstep = (uint32_t)hrows * 2;
            if (T > 0) ptx::mbar_wait(w_full, ph & 1);
            ptx::tc_fence_after();
            for (int64_t i = 0; i < T; ++i) {
                const int64_t seq = (int64_t)ph * T + i;
                const uint32_t a = (uint32_t)(seq & 1);
                ptx::mbar_wait_fast(&acc_empty[a], (acc_uses[a] & 1) ^ 1);
                ptx::mbar_wait(&gx_full[a], acc_uses[a] & 1);
                ++acc_uses[a];
                ptx::tc_fence_after();
                const uint32_t tacc = tmem_base + a * WIDTH;
                const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + FB_GX + a * ACT_BYTES), KROW);
                const uint32_t b_lo = w_a | b_lbo;
                if (ptx::elect_one()) {
#pragma unroll 4
                    for (int kk = 0; kk < nk; ++kk)
                        ptx::mma2_bf16_w(tacc, a_lo + kk * 2 * (KROW >> 4), desc_hi, b_lo + kk * b_kstep, desc_hi, idesc, kk != 0);
                    ptx::mma2_commit(&gx_empty[a]);
                    ptx::mma2_commit(&acc_full[a]);
                    if (i == T - 1) ptx::mma2_commit(w_empty);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups (one thread per sample row) =================
        const int wg = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) | lane;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + wg * WIDTH;
        const uint32_t empty_bar = ptx::mapa(&acc_empty[wg], 0);     // the leader's barrier
        const bool uniform = (N % 32) == 0;       // the 32 rows of a warp then share one ray
        Bands3 bw3; BandsV bwv;
        load_bands(cst + WIDTH, bw3, bwv);
        uint32_t full_uses = 0;
        for (int ph = 0; ph < FB_NPH; ++ph) {
            const int s = fb_step(ph), lo = step_out_layer(s);
            for (int64_t i = 0; i < T; ++i) {
                const int64_t seq = (int64_t)ph * T + i;
                if ((seq & 1) != wg) continue;
                const int64_t tile = tile_of(i), g = tile * TILE + row;
                const bool valid = g < S;
                const int64_t r = valid ? g / N : -1;
                uint8_t* rec = save + tile * SAVE_TILE_BYTES;
                const uint32_t* mask = reinterpret_cast<const uint32_t*>(rec + SV_MASK);
                float* scr = park + tile * (int64_t)(ENC3_PAD * TILE);
                uint32_t mw[MASK_WORDS];
                float gs = 0.f;
                if (lo >= 0) {   // the ReLU flags do not depend on the products: fetch them while the MMAs run
#pragma unroll
                    for (int cc = 0; cc < MASK_WORDS; ++cc) mw[cc] = mask[(lo * MASK_WORDS + cc) * TILE + row];
                    if (s == 2 && valid) gs = d_sigma[g] * sigmoid_f(sig_pre[g]);      // softplus'(x) = sigmoid(x)
                }
                ptx::mbar_wait_fast(&acc_full[wg], full_uses & 1);
                ++full_uses;
                ptx::tc_fence_after();
                if (lo >= 0) {
                    // ---- hidden layers: (+ density rank-1 term) -> ReLU mask -> BF16 -> G image of layer lo ----
                    uint8_t* save_img = rec + SV_G + (int64_t)lo * ACT_BYTES;
                    uint32_t v[2][32];
                    ptx::tmem_ld32(tacc, v[0]);
#pragma unroll
                    for (int cc = 0; cc < WIDTH / 32; ++cc) {
                        ptx::tmem_ld_wait();
                        if (cc + 1 < WIDTH / 32) {
                            ptx::tmem_ld32(tacc + (cc + 1) * 32, v[(cc + 1) & 1]);
                        } else {
                            ptx::tc_fence_before();                 // accumulator drained: the next tile pair may be multiplied
                            ptx::warp_arrive_cluster(empty_bar);
                        }
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
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(save_img + (cc * 4 + q) * KROW + row * 16) =
                                make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
                    }
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
                    ptx::warp_arrive_cluster(empty_bar);
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
                    ptx::tc_fence_before();
                    ptx::warp_arrive_cluster(empty_bar);
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
                    fb_ray_atomic_add3(d_center, r, dc, valid, uniform);
                    fb_ray_atomic_add3(d_ray, r, dv, valid, uniform);
                }
                // this warp's record stores of the item are complete: visible to the producer's bulk copies, counted in
                __threadfence();
                fb_fence_proxy_async_global();
                __syncwarp();
                if (lane == 0) atomicAdd(&done[wg], 1u);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // neither CTA leaves while its peer may still touch its shared memory / TMEM
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

}  // namespace tc

// steps 2 .. 10 of the chain in streaming form (after tc_dx_kernel in head mode has written G7 UNSPLIT into the records)
int tc_dx_stream(const tc::Workspace& w, const float* center, const float* ray, const float* depth, int64_t S, int N,
                 const float* d_sigma, float* d_center, float* d_ray, cudaStream_t st) {
    using namespace tc;
    NIW_CUDA(cudaFuncSetAttribute(tc_dx_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_TOTAL));
    const int64_t npairs = ((S + TILE - 1) / TILE + 1) / 2;
    int64_t pairs = niw_num_sms() / 2;
    if (pairs > npairs) pairs = npairs;
    const int grid = (int)(2 * (pairs < 1 ? 1 : pairs));
    niw::note_launch(), tc_dx_stream_kernel<<<grid, 384, FB_TOTAL, st>>>(w.bstream, w.consts, center, ray, depth, S, N, d_sigma,
                                                                        w.sig_pre, w.save, w.park, d_center, d_ray);
    NIW_LAUNCH_CHECK();
    return 0;
}

}  // namespace niw
