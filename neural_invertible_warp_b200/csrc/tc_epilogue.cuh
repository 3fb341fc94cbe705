// Per-row epilogue helpers of the tcgen05 forward kernel (mlp_tc.cu):
// positional encoding straight into an operand image, and the ReLU + BF16 re-pack of one 32-column accumulator chunk.
#pragma once
#include "tc_layout.cuh"

namespace niw {
namespace tc {

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------

// write the 64-wide encoded position of one row into an A-tile image ([8 chunks][128 rows][8 bf16])
__device__ __forceinline__ void write_enc_row(uint8_t* enc_tile, uint8_t* save_img, int row, const float x[3],
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
        uint4 v = make_uint4(ptx::pack_bf16(e[ch * 8], e[ch * 8 + 1]), ptx::pack_bf16(e[ch * 8 + 2], e[ch * 8 + 3]),
                             ptx::pack_bf16(e[ch * 8 + 4], e[ch * 8 + 5]), ptx::pack_bf16(e[ch * 8 + 6], e[ch * 8 + 7]));
        if (enc_tile) *reinterpret_cast<uint4*>(enc_tile + ch * KROW + row * 16) = v;
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(ENC3_PAD, row, ch)) = v;
    }
}

// write the 32-wide encoded view direction (27 + zero pad) of one row into chunks 0..3 of an enc tile
__device__ __forceinline__ void write_venc_row(uint8_t* enc_tile, uint8_t* save_img, int row, const float v3[3],
                                               const BandsV& bw, bool valid) {
    float e[ENCV_PAD];
    float inv = 1.f / fmaxf(sqrtf(v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2]), 1e-12f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = v3[c] * inv;
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
        uint4 v = make_uint4(ptx::pack_bf16(e[ch * 8], e[ch * 8 + 1]), ptx::pack_bf16(e[ch * 8 + 2], e[ch * 8 + 3]),
                             ptx::pack_bf16(e[ch * 8 + 4], e[ch * 8 + 5]), ptx::pack_bf16(e[ch * 8 + 6], e[ch * 8 + 7]));
        if (enc_tile) *reinterpret_cast<uint4*>(enc_tile + ch * KROW + row * 16) = v;
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(ENCV_PAD, row, ch)) = v;
    }
}

// One 32-column chunk of a layer epilogue.  The accumulator already holds W.x + b (the bias rides on the tensor
// cores), so a pair of columns costs one F2FP.RELU (ReLU + BF16 pack) and two logic ops for the ReLU flags.
//   act      next layer's A image in shared memory (nullptr: not needed, rgb0)
//   save_img this layer's image in the tile record (nullptr: inference); C = image columns
//   flags    this layer's ReLU-flag words of the tile record (nullptr: inference)
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], int cc, int row, int C, uint8_t* act,
                                               uint8_t* save_img, uint32_t* flags, uint32_t (&pk)[16]) {
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        pk[j] = ptx::pack_relu_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        bits |= ptx::gt0_mask_bf16x2(pk[j]) & ptx::relu_mask_const(j);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 o = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        if (act) *reinterpret_cast<uint4*>(act + (cc * 4 + q) * KROW + row * 16) = o;
        if (save_img) *reinterpret_cast<uint4*>(save_img + hbm_img_off(C, row, cc * 4 + q)) = o;
    }
    if (flags) flags[cc * TILE + row] = bits;
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

}  // namespace tc
}  // namespace niw
