// High-precision forward of the fused positional-encoding + 8x256 NeRF MLP on the tcgen05 tensor cores: every GEMM
// operand is carried as a BF16 pair (hi = bf16(x), lo = bf16(x - hi)) and every product is issued as three BF16 MMAs
//     x . w  ~=  hi_x . hi_w  +  lo_x . hi_w  +  hi_x . lo_w          (FP32 accumulate in TMEM; the dropped lo.lo term is 2^-16 relative)
// so that the network is evaluated to ~fp32 operand accuracy -- the 1e-3 max-abs contract on rendered rgb / depth / opacity
// that plain BF16 operands miss (1.1-1.6e-3 rgb, 5e-3-1.2e-2 depth) -- while staying on the tensor pipe.
// SURVEY.md section 8 rows a6+a7+a8 (row n1 of VERDICT r1); reference model/nerf.py:416-456, model/barf.py:256-268 (fp32 there).
//
// Same CTA-pair organisation as mlp_tc.cu (clusters of 2, cta_group::2, M = 256 over the two CTAs' 128-sample tiles, N = 256,
// K = 16; operand images [K/8][rows][8 bf16], no swizzle), with the differences the doubled operands force:
//   * ONE tile pair in flight per CTA pair (hi + lo activation images are 2 x 64 KB + 2 x 16 KB of shared memory), one issuer;
//   * the weight ring carries [hi half chunk ; lo half chunk] per stage (16 KB, 3 stages), 6 MMAs per K = 32 chunk;
//   * the epilogue splits ReLU(acc) into hi / lo images for the next layer; the CUDA-core heads (density row, 128 -> 3 RGB
//     layer) read the FP32 accumulator values directly.
// In training mode the forward also writes the tile records of mlp_tc.cu (hi images, ReLU masks): the BF16 backward
// (mlp_tc_bwd.cu) runs unchanged on them -- north_star's gradient contract is "1e-2 relative for BF16 operands".
#include "tc_layout.cuh"

namespace niw {
namespace tc {

constexpr int X3_NSTAGE = 3;
constexpr int X3_STAGE = 2 * HSTAGE_BYTES;                 // 16384: hi + lo halves of one weight chunk
constexpr int X3_ACT_HI = 0;
constexpr int X3_ACT_LO = X3_ACT_HI + ACT_BYTES;
constexpr int X3_ENC_HI = X3_ACT_LO + ACT_BYTES;
constexpr int X3_ENC_LO = X3_ENC_HI + ENC_BYTES;
constexpr int X3_RING = X3_ENC_LO + ENC_BYTES;
constexpr int X3_ONES = X3_RING + X3_NSTAGE * X3_STAGE;
constexpr int X3_CONST = X3_ONES + ONES_BYTES;
constexpr int X3_BAR = X3_CONST + ((C_FLOATS * 4 + 15) / 16) * 16;
constexpr int X3_TOTAL = X3_BAR + 128;
static_assert(X3_TOTAL <= 227 * 1024, "shared memory budget (split-precision forward)");

// ---- weight stream: per layer [chunk][2 CTA halves][hi | lo][4 k-groups][rows/2][8 bf16], then the K = 16 bias chunk ----
__global__ void pack_weights_x3_kernel(const float* __restrict__ P, uint8_t* __restrict__ stream) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= STREAM_X3_BYTES / 16) return;
    const int64_t byte = gid * 16;
    int l = 0;
#pragma unroll
    for (int i = 1; i < NLAYER; ++i) if (byte >= stream_x3_off(i)) l = i;
    const int rows = layer_rows(l), hrows = rows / 2;
    int64_t rel = byte - stream_x3_off(l);
    const int64_t chunk_bytes = (int64_t)rows * CHUNK_K * 4;               // hi + lo
    const int64_t main_bytes = (int64_t)layer_chunks(l) * chunk_bytes;
    uint32_t out[4] = {0u, 0u, 0u, 0u};
    if (rel < main_bytes) {
        const int chunk = (int)(rel / chunk_bytes), in_chunk = (int)(rel % chunk_bytes);
        const int half_bytes = hrows * CHUNK_K * 4, part_bytes = hrows * CHUNK_K * 2;
        const int half = in_chunk / half_bytes, in_half = in_chunk % half_bytes;
        const int part = in_half / part_bytes, in_part = in_half % part_bytes;
        const int kc = in_part / (hrows * 16);
        const int n = half * hrows + (in_part % (hrows * 16)) / 16;
        const int k0 = chunk * CHUNK_K + kc * 8;
        const int in_dim = layer_in(l);
        const float* Wl = P + layer_woff(l) + (int64_t)(n + layer_rowoff(l)) * in_dim;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ka = k0 + 2 * j, kb = ka + 1;
            const float a = ka < in_dim ? Wl[ka] : 0.f, b = kb < in_dim ? Wl[kb] : 0.f;
            const float ah = __bfloat162float(__float2bfloat16(a)), bh = __bfloat162float(__float2bfloat16(b));
            out[j] = part == 0 ? ptx::pack_bf16(ah, bh) : ptx::pack_bf16(a - ah, b - bh);
        }
    } else {
        rel -= main_bytes;
        const int half = (int)(rel / (hrows * BIAS_K * 2)), in_half = (int)(rel % (hrows * BIAS_K * 2));
        const int kg = in_half / (hrows * 16);
        const int n = half * hrows + (in_half % (hrows * 16)) / 16;
        if (kg == 0) {
            const float b = P[layer_boff(l) + n + layer_rowoff(l)];
            const float hi = __bfloat162float(__float2bfloat16(b));
            out[0] = ptx::pack_bf16(hi, b - hi);
        }
    }
    *reinterpret_cast<uint4*>(stream + byte) = make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ float x3_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float x3_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// 8 fp32 values -> one 16-byte group of the hi image and one of the lo image
__device__ __forceinline__ void split8(const float* e, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = ptx::pack_bf16(e[2 * j], e[2 * j + 1]);
        l[j] = ptx::pack_bf16(e[2 * j] - x3_lo(h[j]), e[2 * j + 1] - x3_hi(h[j]));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void write_enc_row_x3(uint8_t* enc_hi, uint8_t* enc_lo, uint8_t* save_img, int row, const float x[3],
                                                 const Bands3& bw, bool valid) {
    float e[ENC3_PAD];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        e[c] = valid ? x[c] : 0.f;
#pragma unroll
        for (int k = 0; k < L3; ++k) {
            float sn, cs;
            sincos_reduced(x[c] * ((float)(1 << k) * PI_F), sn, cs);
            e[3 + c * 2 * L3 + k] = valid ? bw.w[k] * sn : 0.f;
            e[3 + c * 2 * L3 + L3 + k] = valid ? bw.w[k] * cs : 0.f;
        }
    }
    e[ENC3] = 0.f;
#pragma unroll
    for (int ch = 0; ch < ENC3_PAD / 8; ++ch) {
        uint4 hi, lo;
        split8(e + ch * 8, hi, lo);
        *reinterpret_cast<uint4*>(enc_hi + ch * KROW + row * 16) = hi;
        *reinterpret_cast<uint4*>(enc_lo + ch * KROW + row * 16) = lo;
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(ENC3_PAD, row, ch)) = hi;
    }
}

__device__ __forceinline__ void write_venc_row_x3(uint8_t* enc_hi, uint8_t* enc_lo, uint8_t* save_img, int row, const float v3[3],
                                                  const BandsV& bw, bool valid) {
    float e[ENCV_PAD];
    const float inv = 1.f / fmaxf(sqrtf(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]), 1e-12f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float x = v3[c] * inv;
        e[c] = valid ? x : 0.f;
#pragma unroll
        for (int k = 0; k < LV; ++k) {
            float sn, cs;
            sincos_reduced(x * ((float)(1 << k) * PI_F), sn, cs);
            e[3 + c * 2 * LV + k] = valid ? bw.w[k] * sn : 0.f;
            e[3 + c * 2 * LV + LV + k] = valid ? bw.w[k] * cs : 0.f;
        }
    }
#pragma unroll
    for (int i = ENCV; i < ENCV_PAD; ++i) e[i] = 0.f;
#pragma unroll
    for (int ch = 0; ch < ENCV_PAD / 8; ++ch) {
        uint4 hi, lo;
        split8(e + ch * 8, hi, lo);
        *reinterpret_cast<uint4*>(enc_hi + ch * KROW + row * 16) = hi;
        *reinterpret_cast<uint4*>(enc_lo + ch * KROW + row * 16) = lo;
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(ENCV_PAD, row, ch)) = hi;
    }
}

// One 32-column chunk of a layer epilogue: x = ReLU(acc) (bias already added by the tensor cores) -> hi / lo images.
__device__ __forceinline__ void epilogue_chunk_x3(const uint32_t (&v)[32], int cc, int row, int C, uint8_t* act_hi, uint8_t* act_lo,
                                                  uint8_t* save_img, uint32_t* flags, float (&x)[32]) {
    uint32_t bits = 0, ph[16], pl[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float a = fmaxf(__uint_as_float(v[2 * j]), 0.f), b = fmaxf(__uint_as_float(v[2 * j + 1]), 0.f);
        x[2 * j] = a; x[2 * j + 1] = b;
        ph[j] = ptx::pack_bf16(a, b);
        pl[j] = ptx::pack_bf16(a - x3_lo(ph[j]), b - x3_hi(ph[j]));
        bits |= ptx::gt0_mask_bf16x2(ph[j]) & ptx::relu_mask_const(j);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 oh = make_uint4(ph[q * 4], ph[q * 4 + 1], ph[q * 4 + 2], ph[q * 4 + 3]);
        if (act_hi) {
            *reinterpret_cast<uint4*>(act_hi + (cc * 4 + q) * KROW + row * 16) = oh;
            *reinterpret_cast<uint4*>(act_lo + (cc * 4 + q) * KROW + row * 16) = make_uint4(pl[q * 4], pl[q * 4 + 1], pl[q * 4 + 2], pl[q * 4 + 3]);
        }
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(C, row, cc * 4 + q)) = oh;
    }
    if (flags) flags[cc * TILE + row] = bits;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
tc_fwd_x3_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ consts_g, const float* __restrict__ center,
                 const float* __restrict__ ray, const float* __restrict__ depth, int64_t S, int N,
                 float* __restrict__ rgb_out, float* __restrict__ sigma_out, float* __restrict__ sig_pre,
                 float* __restrict__ rgb_keep, uint8_t* __restrict__ save) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + X3_BAR);
    uint64_t* w_full = bars;                    // [X3_NSTAGE] this CTA's half of the chunk has landed (leader: and the peer's)
    uint64_t* w_empty = bars + X3_NSTAGE;       // [X3_NSTAGE] the MMAs reading the stage have completed (both CTAs)
    uint64_t* a_ready = bars + 2 * X3_NSTAGE;   // [1] (leader) both A tiles are written, accumulators drained
    uint64_t* acc_full = a_ready + 1;           // [1] the layer has been accumulated (both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    float* cst = reinterpret_cast<float*>(smem + X3_CONST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int64_t ntiles = (S + TILE - 1) / TILE;
    const int64_t nunits = (ntiles + 1) / 2;
    const int64_t unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < X3_NSTAGE; ++i) { ptx::mbar_init(&w_full[i], rank == 0 ? 2 : 1); ptx::mbar_init(&w_empty[i], 1); }
        ptx::mbar_init(a_ready, 2 * TILE / 32);
        ptx::mbar_init(acc_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 256);
    if (warp == 3) {
        for (int i = lane; i < C_FLOATS; i += 32) cst[i] = consts_g[i];
        uint4* ones = reinterpret_cast<uint4*>(smem + X3_ONES);
        for (int i = lane; i < ONES_BYTES / 16; i += 32) ones[i] = i < TILE ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    Bands3 bw3; BandsV bwv;
    load_bands(cst + C_BANDS, bw3, bwv);

    if (warp == 0) {
        // ================= weight producer: this CTA's [hi ; lo] half of every chunk =================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int64_t unit = unit0; unit < nunits; unit += unit_step) {
                const uint8_t* src = wstream;
                for (int l = 0; l < NLAYER; ++l) {
                    const int nch = layer_chunks(l), hrows = layer_rows(l) / 2;
                    for (int c = 0; c <= nch; ++c) {                     // chunk nch is the K = 16 bias chunk
                        const uint32_t bytes = c < nch ? (uint32_t)hrows * CHUNK_K * 4 : (uint32_t)hrows * BIAS_K * 2;
                        ptx::mbar_wait(&w_empty[st], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(&w_full[st], bytes);
                        ptx::bulk_g2s(smem + X3_RING + st * X3_STAGE, src + rank * bytes, bytes, &w_full[st]);
                        src += 2 * bytes;
                        if (++st == X3_NSTAGE) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================= peer CTA: tell the leader when this CTA's half of a chunk has landed =================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            const uint32_t full0 = ptx::mapa(&w_full[0], 0);
            for (int64_t unit = unit0; unit < nunits; unit += unit_step)
                for (int l = 0; l < NLAYER; ++l)
                    for (int c = 0; c <= layer_chunks(l); ++c) {
                        ptx::mbar_wait(&w_full[st], ph);
                        ptx::mbar_arrive_cluster(full0 + st * 8);
                        if (++st == X3_NSTAGE) { st = 0; ph ^= 1; }
                    }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= leader CTA: MMA issuer (warp converged, one elected lane issues) =================
        uint32_t st = 0, ph = 0, ready_ph = 0;
        const uint32_t act_hi = ptx::smem_desc_lo(ptx::smem_addr(smem + X3_ACT_HI), KROW);
        const uint32_t act_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + X3_ACT_LO), KROW);
        const uint32_t enc_hi = ptx::smem_desc_lo(ptx::smem_addr(smem + X3_ENC_HI), KROW);
        const uint32_t enc_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + X3_ENC_LO), KROW);
        const uint32_t ones_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + X3_ONES), KROW);
        const uint32_t ring_a = ptx::smem_addr(smem + X3_RING) >> 4;
        const uint32_t desc_hi = ptx::smem_desc_hi(128);
        const uint32_t tacc = tmem_base;
        const uint32_t k2 = 2 * (KROW >> 4);                              // second K = 16 step of a chunk inside an A image
        for (int64_t unit = unit0; unit < nunits; unit += unit_step) {
            for (int l = 0; l < NLAYER; ++l) {
                const int hrows = layer_rows(l) / 2, nch = layer_chunks(l);
                const uint32_t idesc = ptx::idesc_bf16(2 * TILE, 2 * hrows, 0, 0);
                const uint32_t b_lbo = (uint32_t)hrows << 16;
                const uint32_t b_kstep = (uint32_t)hrows * 2;
                const uint32_t b_lopart = (uint32_t)(hrows * CHUNK_K * 2) >> 4;     // lo half chunk behind the hi one
                ptx::mbar_wait_fast(a_ready, ready_ph);
                ready_ph ^= 1;
                ptx::tc_fence_after();
                for (int c = 0; c < nch; ++c) {
                    ptx::mbar_wait(&w_full[st], ph);
                    ptx::tc_fence_after();
                    const bool from_enc = (l == 0) || (c >= 8);
                    const uint32_t off = (uint32_t)((l == 0 || c < 8) ? c : c - 8) * (CHUNK_K / 8) * (KROW >> 4);
                    const uint32_t ah = (from_enc ? enc_hi : act_hi) + off, al = (from_enc ? enc_lo : act_lo) + off;
                    const uint32_t bh = (ring_a + st * (X3_STAGE >> 4)) | b_lbo, bl = bh + b_lopart;
                    if (ptx::elect_one()) {
                        ptx::mma2_bf16_w(tacc, ah, desc_hi, bh, desc_hi, idesc, c != 0);
                        ptx::mma2_bf16_w(tacc, ah + k2, desc_hi, bh + b_kstep, desc_hi, idesc, 1u);
                        ptx::mma2_bf16_w(tacc, al, desc_hi, bh, desc_hi, idesc, 1u);
                        ptx::mma2_bf16_w(tacc, al + k2, desc_hi, bh + b_kstep, desc_hi, idesc, 1u);
                        ptx::mma2_bf16_w(tacc, ah, desc_hi, bl, desc_hi, idesc, 1u);
                        ptx::mma2_bf16_w(tacc, ah + k2, desc_hi, bl + b_kstep, desc_hi, idesc, 1u);
                        ptx::mma2_commit(&w_empty[st]);
                    }
                    __syncwarp();
                    if (++st == X3_NSTAGE) { st = 0; ph ^= 1; }
                }
                {   // bias: D += ones[256 x 16] . [bf16(b), b - bf16(b), 0 ...]^T
                    ptx::mbar_wait(&w_full[st], ph);
                    ptx::tc_fence_after();
                    const uint32_t bh = (ring_a + st * (X3_STAGE >> 4)) | b_lbo;
                    if (ptx::elect_one()) {
                        ptx::mma2_bf16_w(tacc, ones_lo, desc_hi, bh, desc_hi, idesc, 1u);
                        ptx::mma2_commit(&w_empty[st]);
                        ptx::mma2_commit(acc_full);
                    }
                    __syncwarp();
                    if (++st == X3_NSTAGE) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroup (one thread per sample row) =================
        const int row = ((warp & 3) << 5) | lane;
        uint8_t* a_hi = smem + X3_ACT_HI;
        uint8_t* a_lo = smem + X3_ACT_LO;
        uint8_t* e_hi = smem + X3_ENC_HI;
        uint8_t* e_lo = smem + X3_ENC_LO;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t ready_bar = ptx::mapa(a_ready, 0);
        uint32_t full_uses = 0;
        for (int64_t unit = unit0; unit < nunits; unit += unit_step) {
            const int64_t tile = unit * 2 + rank;
            const int64_t g = tile * TILE + row;
            const bool valid = tile < ntiles && g < S;
            uint8_t* save_tile = (save && tile < ntiles) ? save + tile * SAVE_TILE_BYTES : nullptr;
            uint32_t* mask_tile = save_tile ? reinterpret_cast<uint32_t*>(save_tile + SV_MASK) : nullptr;
            float v3[3] = {0.f, 0.f, 1.f};
            {
                float x[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    const int64_t r = g / N;
                    const float d = depth[g];
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        v3[c] = ray[r * 3 + c];
                        x[c] = __fadd_rn(center[r * 3 + c], __fmul_rn(v3[c], d));
                    }
                }
                write_enc_row_x3(e_hi, e_lo, save_tile ? save_tile + SV_ENC : nullptr, row, x, bw3, valid);
            }
            ptx::fence_proxy_async();
            ptx::warp_arrive_cluster(ready_bar);
            for (int l = 0; l < NLAYER; ++l, ++full_uses) {
                ptx::mbar_wait_fast(acc_full, full_uses & 1);
                ptx::tc_fence_after();
                uint32_t v[32];
                float x[32];
                if (l < 8) {
                    uint8_t* save_img = save_tile ? save_tile + SV_H + (int64_t)l * ACT_BYTES : nullptr;
                    uint32_t* flags = mask_tile ? mask_tile + l * MASK_WORDS * TILE : nullptr;
                    float sig_acc = 0.f;
#pragma unroll 1
                    for (int cc = 0; cc < WIDTH / 32; ++cc) {
                        ptx::tmem_ld32(tacc + cc * 32, v);
                        ptx::tmem_ld_wait();
                        epilogue_chunk_x3(v, cc, row, WIDTH, a_hi, a_lo, save_img, flags, x);
                        if (l == 6) {   // density head: row 0 of layer 7 applied to h6 (nerf.py:427), fp32 activations
                            const float4* w = reinterpret_cast<const float4*>(cst + C_W7R0 + cc * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 w4 = w[q];
                                sig_acc += w4.x * x[4 * q] + w4.y * x[4 * q + 1] + w4.z * x[4 * q + 2] + w4.w * x[4 * q + 3];
                            }
                        }
                    }
                    if (l == 7)   // A columns 256..287 of rgb0
                        write_venc_row_x3(e_hi, e_lo, save_tile ? save_tile + SV_VENC : nullptr, row, v3, bwv, valid);
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::warp_arrive_cluster(ready_bar);
                    if (l == 6 && valid) {
                        const float pre = sig_acc + cst[C_MISC];
                        sigma_out[g] = softplus_f(pre);
                        if (sig_pre) sig_pre[g] = pre;
                    }
                } else {
                    // rgb0 epilogue: hr = relu(.), rgb = sigmoid(W_rgb1 hr + b)   (nerf.py:442-446), fp32 throughout
                    float o0 = cst[C_MISC + 1], o1 = cst[C_MISC + 2], o2 = cst[C_MISC + 3];
                    uint8_t* save_img = save_tile ? save_tile + SV_HR : nullptr;
                    uint32_t* flags = mask_tile ? mask_tile + 8 * MASK_WORDS * TILE : nullptr;
#pragma unroll 1
                    for (int cc = 0; cc < RGBW / 32; ++cc) {
                        ptx::tmem_ld32(tacc + cc * 32, v);
                        ptx::tmem_ld_wait();
                        epilogue_chunk_x3(v, cc, row, RGBW, nullptr, nullptr, save_img, flags, x);
                        const float4* w0 = reinterpret_cast<const float4*>(cst + C_WRGB1 + cc * 32);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float a = x[4 * q], b = x[4 * q + 1], c = x[4 * q + 2], d = x[4 * q + 3];
                            const float4 x0 = w0[q], x1 = w0[RGBW / 4 + q], x2 = w0[2 * RGBW / 4 + q];
                            o0 += x0.x * a + x0.y * b + x0.z * c + x0.w * d;
                            o1 += x1.x * a + x1.y * b + x1.z * c + x1.w * d;
                            o2 += x2.x * a + x2.y * b + x2.z * c + x2.w * d;
                        }
                    }
                    ptx::tc_fence_before();   // TMEM reads done before the next unit overwrites the accumulator
                    if (valid) {
                        const float r0 = sigmoid_f(o0), r1 = sigmoid_f(o1), r2 = sigmoid_f(o2);
                        rgb_out[g * 3] = r0; rgb_out[g * 3 + 1] = r1; rgb_out[g * 3 + 2] = r2;
                        if (rgb_keep) { rgb_keep[g * 3] = r0; rgb_keep[g * 3 + 1] = r1; rgb_keep[g * 3 + 2] = r2; }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 256);
}

}  // namespace tc

size_t tc_x3_workspace_bytes(int64_t R, int N, int training) { return tc::carve(nullptr, R * (int64_t)N, training != 0, true).bytes; }

// fp32 parameters -> hi / lo BF16 weight stream + fp32 constants (+ the BF16 streams of the backward pass when training)
int tc_x3_pack(const float* P, const C2F& c2f, int training, int64_t R, int N, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tc;
    Workspace w = carve(ws, R * (int64_t)N, training != 0, true);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    // constants, band weights (and, when a backward will follow, the BF16 forward / transposed streams it reads)
    int e = tc_pack(P, c2f, training, R, N, ws, ws_bytes, st);
    if (e) return e;
    niw::note_launch(), pack_weights_x3_kernel<<<niw_blocks(STREAM_X3_BYTES / 16, 256), 256, 0, st>>>(P, w.wstream_x3);
    NIW_LAUNCH_CHECK();
    return 0;
}

int tc_x3_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N, const C2F& c2f,
              int training, void* ws, size_t ws_bytes, float* rgb, float* sigma, cudaStream_t st) {
    using namespace tc;
    const int64_t S = R * (int64_t)N;
    const bool prepacked = (training & NIW_NERF_PREPACKED) != 0;
    training &= 1;
    Workspace w = carve(ws, S, training != 0, true);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    if (!prepacked) {
        int e = tc_x3_pack(P, c2f, training, R, N, ws, ws_bytes, st);
        if (e) return e;
    }
    NIW_CUDA(cudaFuncSetAttribute(tc_fwd_x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X3_TOTAL));
    const int64_t nunits = ((S + TILE - 1) / TILE + 1) / 2;       // one tile pair per CTA pair and round
    int64_t pairs = niw_num_sms() / 2;
    if (pairs > nunits) pairs = nunits;
    const int grid = (int)(2 * (pairs < 1 ? 1 : pairs));
    niw::note_launch(), tc_fwd_x3_kernel<<<grid, 256, X3_TOTAL, st>>>(w.wstream_x3, w.consts, center, ray, depth, S, N, rgb, sigma,
                                                                     training ? w.sig_pre : nullptr, training ? w.rgb_keep : nullptr,
                                                                     training ? w.save : nullptr);
    NIW_LAUNCH_CHECK();
    return 0;
}

}  // namespace niw
