// Architecture constants of the reference NeRF MLP (options/nerf_inn_llff.yaml:3-11,
// model/nerf.py:373-402) and offsets into the flat fp32 parameter vector (state_dict order).
#pragma once
#include <stdint.h>

namespace niw {

constexpr int L3 = 10, LV = 4;
constexpr int ENC3 = 3 + 6 * L3;   // 63
constexpr int ENCV = 3 + 6 * LV;   // 27
constexpr int ENC3_PAD = 64, ENCV_PAD = 32;
constexpr int WIDTH = 256, RGBW = 128, NFEAT = 8, SKIP = 4;

// in / out width of mlp_feat layer l
__host__ __device__ constexpr int feat_in(int l) { return l == 0 ? ENC3 : (l == SKIP ? WIDTH + ENC3 : WIDTH); }
__host__ __device__ constexpr int feat_out(int l) { return l == NFEAT - 1 ? WIDTH + 1 : WIDTH; }

__host__ __device__ constexpr int64_t feat_w_off(int l) {
    int64_t o = 0;
    for (int i = 0; i < l; ++i) o += (int64_t)feat_out(i) * feat_in(i) + feat_out(i);
    return o;
}
__host__ __device__ constexpr int64_t feat_b_off(int l) { return feat_w_off(l) + (int64_t)feat_out(l) * feat_in(l); }
constexpr int64_t RGB0_W = feat_w_off(NFEAT);
constexpr int64_t RGB0_B = RGB0_W + (int64_t)RGBW * (WIDTH + ENCV);
constexpr int64_t RGB1_W = RGB0_B + RGBW;
constexpr int64_t RGB1_B = RGB1_W + 3 * RGBW;
constexpr int64_t NPARAMS = RGB1_B + 3;
static_assert(NPARAMS == 530052, "527872 weights + 2180 biases (the reference count 530053 includes barf progress)");

constexpr float PI_F = 3.14159274101257324f;  // fp32(pi): the reference builds freq = 2^k * pi in fp32

}  // namespace niw
