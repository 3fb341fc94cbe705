// Layout shared by the tcgen05 MLP kernels (mlp_tc.cu forward, mlp_tc_bwd.cu backward):
// tile geometry, shared-memory operand images, the per-tile record saved for the backward pass,
// weight-stream step tables and the workspace carve-up.
//
// Operand image in shared memory (forward / dX kernels, every activation / gradient tile):
//     [C/8 groups][128 rows][8 bf16]      (C = feature columns, rows = samples of the tile)
// i.e. element (row r, column c) lives at (c/8)*2048 + r*16 + (c%8)*2 bytes: a K-major A operand with
// K = features, LBO = 2048, SBO = 128.
// The copy saved in HBM for the weight-gradient pass is split into two 64-sample halves,
//     [2 halves][C/8 groups][64 rows][8 bf16]     (hbm_img_off below)
// so that a half image is one contiguous block (one bulk copy) and, in shared memory, an MN-major operand
// with K = samples: LBO = 128 (8 samples), SBO = 1024 (8 features).  A warp of epilogue threads (32 rows,
// one 16-byte group each) writes 512 contiguous bytes in either layout.
// (no swizzle; both conventions verified on hardware by tests/test_gpu_tc.py::test_tc_selftest_variants).
#pragma once
#include "common.cuh"
#include "mlp_shared.cuh"
#include "tc_ptx.cuh"

namespace niw {
namespace tc {

constexpr int TILE = 128;                         // samples per tile (UMMA M)
constexpr int CHUNK_K = 32;                       // K extent of one streamed weight chunk
constexpr int KROW = TILE * 16;                   // 2048: bytes of one 8-column group of an image
constexpr int ACT_BYTES = TILE * WIDTH * 2;       // 65536: 256-column image
constexpr int HR_BYTES = TILE * RGBW * 2;         // 32768: 128-column image
constexpr int ENC_BYTES = TILE * ENC3_PAD * 2;    // 16384: 64-column image
constexpr int VENC_BYTES = TILE * ENCV_PAD * 2;   // 8192:  32-column image
constexpr int SMALL_BYTES = TILE * 8 * 2;         // 2048:  8-column image (g_rgb_pre[3], g_sigma_pre, 1, 0, 0, 0)
constexpr int HALF = 64;                          // samples per half image (K extent of one dW stage)
constexpr int HROW = HALF * 16;                   // 1024: bytes of one 8-column group of a half image
constexpr int STAGE_BYTES = WIDTH * CHUNK_K * 2;  // 16384: one weight chunk (256 rows x 32 k)
constexpr int HSTAGE_BYTES = STAGE_BYTES / 2;     // 8192:  the half of a weight chunk one CTA of a pair stages

// ---- fp32 constants kept in shared memory (CUDA-core head weights, band weights) ------------------
constexpr int NLAYER = 9;                         // forward: 8 feature layers + rgb0
constexpr int C_W7R0 = 0;                         // [256]  density row of layer 7
constexpr int C_WRGB1 = C_W7R0 + WIDTH;           // [3][128]
constexpr int C_MISC = C_WRGB1 + 3 * RGBW;        // b7[0], brgb1[0..2]
constexpr int C_BANDS = C_MISC + 4;               // [NBANDS] coarse-to-fine band weights (evaluated on the device)
constexpr int C_FLOATS = C_BANDS + NBANDS;
// The biases of the 9 GEMM layers ride on the tensor cores: every layer's weight stream ends with a K = 16
// chunk whose k = 0 / 1 columns hold bf16(b) and bf16(b - bf16(b)), multiplied by a constant A image with
// ones in those two columns (ONES_BYTES in shared memory), so the accumulator leaves TMEM with the bias added
// (to ~16 mantissa bits) and the epilogue needs no add.
constexpr int BIAS_K = 16;
constexpr int ONES_BYTES = TILE * BIAS_K * 2;     // 4096: [2 k-groups][128 rows][8 bf16]

// ---- per-tile record saved by the forward pass (training) and extended by the dX pass ---------
constexpr int64_t SV_H = 0;                                  // h0..h7 images
constexpr int64_t SV_HR = SV_H + 8 * (int64_t)ACT_BYTES;     // hr image
constexpr int64_t SV_ENC = SV_HR + HR_BYTES;                 // encoded position image
constexpr int64_t SV_VENC = SV_ENC + ENC_BYTES;              // encoded view direction image
constexpr int MASK_WORDS = 8;                                // 256 ReLU bits per row and layer
constexpr int64_t SV_MASK = SV_VENC + VENC_BYTES;            // [9][8 words][128 rows] uint32 (bit = activation > 0)
constexpr int64_t MASK_BYTES = 9 * MASK_WORDS * TILE * 4;
constexpr int64_t SV_G = SV_MASK + MASK_BYTES;               // G0..G7: gradient wrt the pre-activation of layer l
constexpr int64_t SV_G8 = SV_G + 8 * (int64_t)ACT_BYTES;     // gradient wrt the pre-activation of rgb0
constexpr int64_t SV_SMALL = SV_G8 + HR_BYTES;               // 8-column image, see SMALL_BYTES
constexpr int64_t SAVE_TILE_BYTES = SV_SMALL + SMALL_BYTES;
static_assert(SAVE_TILE_BYTES % 256 == 0, "tile record alignment");
// byte offset of the 16-byte group (row, column group cg) inside a saved image of C columns
__host__ __device__ constexpr int hbm_img_off(int C, int row, int cg) {
    return (row >> 6) * (C / 8) * HROW + cg * HROW + (row & (HALF - 1)) * 16;
}

// ---- forward weight stream ----------------------------------------------------------------------
__host__ __device__ constexpr int layer_chunks(int l) { return l == 0 ? 2 : (l == 4 ? 10 : (l == 8 ? 9 : 8)); }
__host__ __device__ constexpr int layer_rows(int l) { return l == 8 ? RGBW : WIDTH; }
__host__ __device__ constexpr int layer_in(int l) { return l == 8 ? WIDTH + ENCV : feat_in(l); }
__host__ __device__ constexpr int64_t layer_woff(int l) { return l == 8 ? RGB0_W : feat_w_off(l); }
__host__ __device__ constexpr int64_t layer_boff(int l) { return l == 8 ? RGB0_B : feat_b_off(l); }
__host__ __device__ constexpr int layer_rowoff(int l) { return l == 7 ? 1 : 0; }   // layer 7: row 0 is the density head
__host__ __device__ constexpr int64_t layer_stream_bytes(int l) {
    return (int64_t)layer_rows(l) * (layer_chunks(l) * CHUNK_K + BIAS_K) * 2;
}
__host__ __device__ constexpr int64_t stream_off(int l) {
    int64_t o = 0;
    for (int i = 0; i < l; ++i) o += layer_stream_bytes(i);
    return o;
}
constexpr int64_t STREAM_BYTES = stream_off(NLAYER);
static_assert(STREAM_BYTES == 1126400, "weight stream size");
// split-precision forward (mlp_tc_x3.cu): every main chunk carries a hi and a lo BF16 image of the weights
__host__ __device__ constexpr int64_t layer_stream_x3_bytes(int l) {
    return (int64_t)layer_rows(l) * (2 * layer_chunks(l) * CHUNK_K + BIAS_K) * 2;
}
__host__ __device__ constexpr int64_t stream_x3_off(int l) {
    int64_t o = 0;
    for (int i = 0; i < l; ++i) o += layer_stream_x3_bytes(i);
    return o;
}
constexpr int64_t STREAM_X3_BYTES = stream_x3_off(NLAYER);

// ---- dX-pass weight stream (transposed weights): D[samples, N = inputs] = G[samples, K = outputs] . B^T ----
// step:        0 view   1 rgb0   2 L7    3 L6    4 L5    5 L4enc  6 L4    7 L3    8 L2    9 L1    10 L0
constexpr int NSTEP = 11;
__host__ __device__ constexpr int step_n(int s) { return s == 0 ? ENCV_PAD : ((s == 5 || s == 10) ? ENC3_PAD : WIDTH); }
__host__ __device__ constexpr int step_k(int s) { return s <= 1 ? RGBW : WIDTH; }
__host__ __device__ constexpr int step_chunks(int s) { return step_k(s) / CHUNK_K; }
__host__ __device__ constexpr int step_layer(int s) {       // forward layer whose weights the step reads (8 = rgb0)
    return s <= 1 ? 8 : (s <= 4 ? 9 - s : (s == 5 ? 4 : 10 - s));
}
__host__ __device__ constexpr int step_col0(int s) { return (s == 0 || s == 5) ? WIDTH : 0; }       // first input column
__host__ __device__ constexpr int step_nvalid(int s) { return s == 0 ? ENCV : ((s == 5 || s == 10) ? ENC3 : WIDTH); }
__host__ __device__ constexpr int64_t bstream_off(int s) {
    int64_t o = 0;
    for (int i = 0; i < s; ++i) o += (int64_t)step_chunks(i) * step_n(i) * CHUNK_K * 2;
    return o;
}
constexpr int64_t BSTREAM_BYTES = bstream_off(NSTEP);
// which saved mask gates the output of a step, and which G image it produces (-1: none)
__host__ __device__ constexpr int step_out_layer(int s) {
    return s == 1 ? 7 : (s >= 2 && s <= 4 ? 8 - s : (s >= 6 && s <= 9 ? 9 - s : -1));
}

// ---- weight-gradient pass: work units ----------------------------------------------------------
struct DwSide { int32_t smem_off, groups, kind, pad_; int64_t base; };   // thin product on one staged image
struct DwUnit {
    int32_t a_off, a_half;       // G image: byte offset inside the tile record, bytes of one 64-sample half
    int32_t m_halves;            // 128-feature M blocks of the G image (1 or 2)
    int32_t a2_off, a2_half;     // second image staged behind a 128-feature G image (thin products only), 0 = none
    int32_t b_off, b_half;       // X image; b_half = 0 -> no main product
    int32_t n_main;              // N = columns of the X image
    int32_t b2_off, b2_half, n2; // second X image of the first M block (TMEM columns 256..), 0 = none
    int32_t ld, col0, ncols;     // output: dP[w_base + feature*ld + col0 + j], j < ncols
    int32_t col0_2, ncols2;      // same for the second X image
    int32_t first_cta, n_slices; // CTAs [first_cta, first_cta + n_slices) split the tiles of this unit
    int64_t w_base;
    // thin products (kind 0 none; 1 bias: row 4 -> base[f]; 2 rgb1 weights: rows 0-2 -> base[m*128 + f];
    // 3 density row: row 3 -> base[f]) of small^T . image, image at smem_off inside the stage
    DwSide side[2];
};
constexpr int MAX_UNITS = 24;
struct DwPlan { DwUnit u[MAX_UNITS]; int n_units; int n_ctas; int max_slices; };

// ---- workspace -----------------------------------------------------------------------------------
struct Workspace {
    uint8_t* wstream;     // forward weight stream (bf16)
    uint8_t* bstream;     // dX-pass (transposed) weight stream
    float* consts;        // C_FLOATS
    float* sig_pre;       // [S]
    float* rgb_keep;      // [S,3]
    uint8_t* save;        // [tiles] x SAVE_TILE_BYTES
    float* scratch;       // [SMs*2 slots][64][128] fp32: skip-connection gradient parked between steps 5 and 10
    float* partial;       // [max_slices][NPARAMS] fp32 weight-gradient partial sums
    uint8_t* wstream_x3;  // split-precision forward weight stream (hi / lo bf16), carved last: the other offsets do not move
    size_t bytes;
};
constexpr int PARTIAL_SLICES = 16;

inline Workspace carve(void* base, int64_t S, bool training, bool x3 = false) {
    Workspace w;
    size_t off = 0;
    auto take = [&](size_t n) { uint8_t* p = base ? (uint8_t*)base + off : nullptr; off += (n + 255) & ~size_t(255); return p; };
    int64_t tiles = (S + TILE - 1) / TILE;
    w.wstream = take(STREAM_BYTES);
    w.bstream = take(BSTREAM_BYTES);
    w.consts = (float*)take(C_FLOATS * 4);
    w.sig_pre = (float*)take(training ? S * 4 : 0);
    w.rgb_keep = (float*)take(training ? S * 12 : 0);
    w.save = take(training ? tiles * SAVE_TILE_BYTES : 0);
    w.scratch = (float*)take(training ? (size_t)niw_num_sms() * 2 * ENC3_PAD * TILE * 4 : 0);
    w.partial = (float*)take(training ? (size_t)PARTIAL_SLICES * NPARAMS * 4 : 0);
    w.wstream_x3 = take(x3 ? STREAM_X3_BYTES : 0);
    w.bytes = off;
    return w;
}

// ---- device helpers ------------------------------------------------------------------------------

// sin / cos of an fp32 argument that may be huge (inverse-depth samples reach |x| ~ 1e8): exact
// range reduction of the *rounded fp32 argument* in fp64 (so the value matches the reference's
// sin(fp32(x*freq)) rather than the mathematically exact sin(2^k pi x)), then MUFU on |r| <= pi/2.
__device__ __forceinline__ void sincos_reduced(float arg, float& s, float& c) {
    double t = (double)arg * 0.31830988618379067154;
    long long n = __double2ll_rn(t);
    float fr = (float)(t - (double)n) * PI_F;
    s = __sinf(fr); c = __cosf(fr);
    if (n & 1) { s = -s; c = -c; }
}
__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(__expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + __expf(-x)); }

}  // namespace tc
}  // namespace niw
