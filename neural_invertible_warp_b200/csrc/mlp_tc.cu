// Fused positional-encoding + 8x256 NeRF MLP on the 5th-generation tensor cores (tcgen05, sm_100a).
// SURVEY.md section 8 rows a6+a7+a8; reference model/nerf.py:416-456, model/barf.py:256-268.
//
// Forward (tc_fwd_kernel): persistent, CTA PAIRS (clusters of 2 on neighbouring SMs, tcgen05 cta_group::2),
// 384 threads per CTA.  A pair works on up to four 128-sample tiles at a time: slot s of CTA r holds tile 2*hq + r of the
// tile pair hq assigned to the slot (round k: hq = (2k + s) * pairs + pair; the last round may use slot 0 only),
// and every MMA is an M = 256 instruction over the two tiles of one slot (rows 0-127 in the leader, 128-255 in
// its peer), so each CTA only stages HALF of every weight chunk.  That halving pays for streaming the weights
// once per SLOT instead of once per pair of slots, which is what lets the two slots run a layer apart: while the
// tensor cores work on slot 1, the epilogue warps turn slot 0's accumulator into the next layer's A operand.
//   warp 0      producer: streams this CTA's half of the pre-packed BF16 weight chunks from L2 into a
//               7-stage shared-memory ring with 1-D bulk async copies (TMA engine, UBLKCP)
//   warp 1      leader CTA: MMA issuer -- one thread issues tcgen05.mma.cta_group::2 (M=256, N=256|128, K=16)
//               with the activation tiles as A (shared memory, K-major, no swizzle) and the weight halves as B,
//               and commits to the barriers of BOTH CTAs.  peer CTA: relays "my half of the chunk has landed"
//               to the leader
//   warp 2      TMEM allocator (2 x 256 fp32 columns = all 512 columns, in both CTAs)
//   warp 3      loads head weights into shared memory, builds the constant A image of the bias products
//   warps 4-7   epilogue of tile slot 0,  warps 8-11 epilogue of tile slot 1: one thread per
//               sample row; TMEM -> registers -> ReLU -> BF16 -> next layer's A tile in
//               shared memory (and to HBM for the backward pass).  The density head (row 0 of
//               layer 7) and the 128->3 RGB layer are folded into the epilogues as FP32 dot
//               products, the positional encoding is computed straight into the A tile.  The biases are added
//               by the tensor cores (a K = 16 chunk against a constant "ones" A image).
//
// Shared-memory operand layout (both operands, all layers): [K/8 chunks][rows][8 bf16], i.e. the
// canonical no-swizzle K-major UMMA layout with 128 B core matrices, SBO = 128 B (8-row groups are
// contiguous) and LBO = rows*16 B.  Epilogue thread r writes 16 B at chunk*rows*16 + r*16: a warp
// writes 512 contiguous bytes (bank-conflict free).
#include "tc_layout.cuh"
#include "tc_epilogue.cuh"
#ifndef NIW_NSTAGE
#define NIW_NSTAGE 7
#endif

namespace niw {

namespace tc {

constexpr int NSTAGE = NIW_NSTAGE;
#ifndef NIW_ISSUE_GROUP
#define NIW_ISSUE_GROUP 2
#endif
constexpr int ISSUE_GROUP = NIW_ISSUE_GROUP;     // ring entries (2 MMAs each) per elected region of an issuer
static_assert(ISSUE_GROUP >= 1 && ISSUE_GROUP <= NSTAGE - 2, "an issuer waits for a whole group of ring stages");

// shared memory map of the forward kernel
constexpr int SM_ACT = 0;
constexpr int SM_ENC = SM_ACT + 2 * ACT_BYTES;
constexpr int SM_RING = SM_ENC + 2 * ENC_BYTES;
constexpr int SM_ONES = SM_RING + NSTAGE * HSTAGE_BYTES;
constexpr int SM_CONST = SM_ONES + ONES_BYTES;
constexpr int SM_BAR = SM_CONST + ((C_FLOATS * 4 + 15) / 16) * 16;
constexpr int SM_TOTAL = SM_BAR + 256;
static_assert(SM_TOTAL <= 227 * 1024, "shared memory budget");

// ------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> BF16 chunk stream in MMA consumption order + fp32 constants
// ------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ P, C2F c2f, uint8_t* __restrict__ stream,
                                    float* __restrict__ consts) {
    // one thread per 16-byte group (8 consecutive k of one row)
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ngroups = STREAM_BYTES / 16;
    if (gid < ngroups) {
        int64_t byte = gid * 16;
        int l = 0;
#pragma unroll
        for (int i = 1; i < NLAYER; ++i) if (byte >= stream_off(i)) l = i;
        const int rows = layer_rows(l);
        int64_t rel = byte - stream_off(l);
        const int64_t main_bytes = (int64_t)layer_chunks(l) * rows * CHUNK_K * 2;
        uint32_t out[4] = {0u, 0u, 0u, 0u};
        const int hrows = rows / 2;                          // rows staged by one CTA of the pair
        if (rel < main_bytes) {
            // chunk = [2 CTA halves][4 k-groups][rows/2][8 bf16]
            int chunk = (int)(rel / ((int64_t)rows * CHUNK_K * 2));
            int in_chunk = (int)(rel % ((int64_t)rows * CHUNK_K * 2));
            int half = in_chunk / (hrows * CHUNK_K * 2);
            int in_half = in_chunk % (hrows * CHUNK_K * 2);
            int kc = in_half / (hrows * 16);                 // which 8-wide k group inside the chunk
            int n = half * hrows + (in_half % (hrows * 16)) / 16;
            int k0 = chunk * CHUNK_K + kc * 8;
            const int in_dim = layer_in(l);
            const float* Wl = P + layer_woff(l) + (int64_t)(n + layer_rowoff(l)) * in_dim;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int ka = k0 + 2 * j, kb = ka + 1;
                out[j] = ptx::pack_bf16(ka < in_dim ? Wl[ka] : 0.f, kb < in_dim ? Wl[kb] : 0.f);
            }
        } else {
            // bias chunk [2 CTA halves][2 k-groups][rows/2][8 bf16]: k = 0 -> bf16(b), k = 1 -> bf16(b - bf16(b)), rest 0
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
    if (gid < C_FLOATS) {
        int i = (int)gid;
        float v;
        if (i >= C_BANDS) {
            store_bands(c2f, i - C_BANDS, consts + C_BANDS);
            return;
        }
        if (i < C_WRGB1) {
            v = P[feat_w_off(7) + (i - C_W7R0)];
        } else if (i < C_MISC) {
            v = P[RGB1_W + (i - C_WRGB1)];
        } else {
            int j = i - C_MISC;
            v = j == 0 ? P[feat_b_off(7)] : P[RGB1_B + (j - 1)];
        }
        consts[i] = v;
    }
}

// transposed stream of the dX pass: step s, chunk c holds B[n][k] = W_layer[k + rowoff][col0 + n]
// for k in [32c, 32c+32) as [2 CTA halves][4 k-groups][N/2 rows][8 bf16]
__global__ void pack_weights_bwd_kernel(const float* __restrict__ P, uint8_t* __restrict__ stream) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= BSTREAM_BYTES / 16) return;
    const int64_t byte = gid * 16;
    int s = 0;
#pragma unroll
    for (int i = 1; i < NSTEP; ++i) if (byte >= bstream_off(i)) s = i;
    const int n_rows = step_n(s), hrows = n_rows / 2;
    const int64_t rel = byte - bstream_off(s);
    const int chunk = (int)(rel / ((int64_t)n_rows * CHUNK_K * 2));
    const int in_chunk = (int)(rel % ((int64_t)n_rows * CHUNK_K * 2));
    const int half = in_chunk / (hrows * CHUNK_K * 2), in_half = in_chunk % (hrows * CHUNK_K * 2);
    const int kg = in_half / (hrows * 16);
    const int n = half * hrows + (in_half % (hrows * 16)) / 16;
    const int k0 = chunk * CHUNK_K + kg * 8;
    const int l = step_layer(s);
    const int in_dim = layer_in(l);
    const float* Wl = P + layer_woff(l) + (int64_t)layer_rowoff(l) * in_dim + step_col0(s) + n;
    const bool ok = n < step_nvalid(s);
    uint32_t out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = ok ? Wl[(int64_t)(k0 + 2 * j) * in_dim] : 0.f;
        float b = ok ? Wl[(int64_t)(k0 + 2 * j + 1) * in_dim] : 0.f;
        out[j] = ptx::pack_bf16(a, b);
    }
    *reinterpret_cast<uint4*>(stream + byte) = make_uint4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------------------------------
// fused forward kernel
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
tc_fwd_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ consts_g, const float* __restrict__ center,
              const float* __restrict__ ray, const float* __restrict__ depth, int64_t S, int N,
              float* __restrict__ rgb_out, float* __restrict__ sigma_out, float* __restrict__ sig_pre,
              float* __restrict__ rgb_keep, uint8_t* __restrict__ save) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
    // w_full has TWO barriers per stage, used on alternating trips round the ring: the two issuer threads skip each
    // other's chunks without observing their phases, and with one barrier per stage the parity test of a phase
    // aliases as soon as a thread is a full trip away from the barrier's current phase (seen as a deadlock)
    uint64_t* w_full = bars;                 // [2][NSTAGE] this CTA's half of the chunk has landed (leader: and the peer's)
    uint64_t* w_empty = bars + 2 * NSTAGE;   // [NSTAGE] the MMAs reading the stage have completed (both CTAs)
    uint64_t* a_ready = bars + 3 * NSTAGE;   // [2]      (leader) both A tiles of the slot are written, accumulators drained
    uint64_t* acc_full = a_ready + 2;        // [2]      the slot's layer has been accumulated (both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 2);
    float* cst = reinterpret_cast<float*>(smem + SM_CONST);

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
        // leader: a stage is full when its own copy has landed (expect_tx arrive) and the peer has reported its half
        for (int i = 0; i < 2 * NSTAGE; ++i) ptx::mbar_init(&w_full[i], rank == 0 ? 2 : 1);
        for (int i = 0; i < NSTAGE; ++i) ptx::mbar_init(&w_empty[i], 1);
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&a_ready[i], 2 * TILE / 32); ptx::mbar_init(&acc_full[i], 1); }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    if (warp == 3) {
        for (int i = lane; i < C_FLOATS; i += 32) cst[i] = consts_g[i];
        // constant A image of the bias products: columns 0 and 1 are 1.0, the other 14 are 0
        uint4* ones = reinterpret_cast<uint4*>(smem + SM_ONES);
        for (int i = lane; i < ONES_BYTES / 16; i += 32) ones[i] = i < TILE ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // the peer's barriers are initialised before anyone signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    Bands3 bw3; BandsV bwv;
    load_bands(cst + C_BANDS, bw3, bwv);

    if (warp == 0) {
        // ================= weight producer (this CTA's half of every chunk, once per slot) =================
        if (lane == 0) {
            uint32_t st = 0, cyc = 0;                  // ring stage, trips round the ring
            for (int64_t k = 0; hq_of(k, 0) < nhq; ++k) {
                const int nslots = hq_of(k, 1) < nhq ? 2 : 1;
                const uint8_t* lsrc = wstream;
                for (int l = 0; l < NLAYER; ++l) {
                    const int nch = layer_chunks(l), hrows = layer_rows(l) / 2;
                    for (int s = 0; s < nslots; ++s) {
                        const uint8_t* src = lsrc;
                        for (int c = 0; c <= nch; ++c) {                // chunk nch is the K = 16 bias chunk
                            const uint32_t bytes = (uint32_t)hrows * (c < nch ? CHUNK_K : BIAS_K) * 2;
                            uint64_t* full = &w_full[(cyc & 1) * NSTAGE + st];
                            ptx::mbar_wait(&w_empty[st], (cyc & 1) ^ 1);
                            ptx::mbar_arrive_expect_tx(full, bytes);
                            ptx::bulk_g2s(smem + SM_RING + st * HSTAGE_BYTES, src + rank * bytes, bytes, full);
                            src += 2 * bytes;
                            if (++st == NSTAGE) { st = 0; ++cyc; }
                        }
                    }
                    lsrc += layer_stream_bytes(l);
                }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================= peer CTA: tell the leader when this CTA's half of a chunk has landed =================
        if (lane == 0) {
            uint32_t st = 0, cyc = 0;
            const uint32_t full0 = ptx::mapa(&w_full[0], 0);
            for (int64_t k = 0; hq_of(k, 0) < nhq; ++k)
                for (int l = 0; l < NLAYER; ++l)
                    for (int c = 0; c < (hq_of(k, 1) < nhq ? 2 : 1) * (layer_chunks(l) + 1); ++c) {
                        const uint32_t fb = (cyc & 1) * NSTAGE + st;
                        ptx::mbar_wait(&w_full[fb], (cyc >> 1) & 1);
                        ptx::mbar_arrive_cluster(full0 + fb * 8);
                        if (++st == NSTAGE) { st = 0; ++cyc; }
                    }
        }
    } else if ((warp == 1 || warp == 2) && rank == 0) {
        // ================= leader CTA: MMA issuers, one warp per slot (M = 256 over the pair) =================
        // The chunk stream alternates [slot 0: layer l][slot 1: layer l]; each issuer walks its own segments.  One
        // thread cannot issue fast enough for both slots.  The warp stays converged (uniform loop state), one elected
        // lane issues the tcgen05 instructions.
        {
            const int s = warp - 1;
            uint32_t st = 0, cyc = 0, ready_ph = 0;
            auto skip = [&](int n) { st += n; while (st >= NSTAGE) { st -= NSTAGE; ++cyc; } };   // the other slot's chunks
            auto wait_full = [&]() { ptx::mbar_wait(&w_full[(cyc & 1) * NSTAGE + st], (cyc >> 1) & 1); };
            // sound as long as the previous phase of a barrier (chunk X - 2*NSTAGE) has completed when an issuer starts
            // to wait for chunk X: its previous own chunk X' was waited for and X - X' <= longest segment + 1
            static_assert(2 * NSTAGE >= (10 + 1) + 1, "ring too short for two skipping issuers");
            const uint32_t act_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + SM_ACT + s * ACT_BYTES), KROW);
            const uint32_t enc_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + SM_ENC + s * ENC_BYTES), KROW);
            const uint32_t ones_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + SM_ONES), KROW);
            const uint32_t ring_a = ptx::smem_addr(smem + SM_RING) >> 4;
            const uint32_t desc_hi = ptx::smem_desc_hi(128);
            const uint32_t tacc = tmem_base + s * WIDTH;
            for (int64_t k = 0; hq_of(k, s) < nhq; ++k) {
                const bool both = hq_of(k, 1) < nhq;                 // the other slot works in this round too
                for (int l = 0; l < NLAYER; ++l) {
                    const int hrows = layer_rows(l) / 2, nch = layer_chunks(l);
                    const uint32_t idesc = ptx::idesc_bf16(2 * TILE, 2 * hrows, 0, 0);
                    const uint32_t b_lbo = (uint32_t)hrows << 16;            // LBO field: hrows * 16 B >> 4
                    const uint32_t b_kstep = (uint32_t)hrows * 2;            // one K = 16 step: 2 k-groups x hrows x 16 B >> 4
                    if (s == 1) skip(nch + 1);
                    ptx::mbar_wait_fast(&a_ready[s], ready_ph);
                    ready_ph ^= 1;
                    ptx::tc_fence_after();
                    // ring entries 0 .. nch-1 are the K = 32 weight chunks, entry nch the K = 16 bias chunk
                    // (D += ones[256 x 16] . [bf16(b), b - bf16(b), 0 ...]^T); ISSUE_GROUP entries per elected region: every
                    // elect / reconverge round trip costs issue time (profiles/r2_stream_experiment.md)
                    for (int e = 0; e <= nch; e += ISSUE_GROUP) {
                        const int n_in = nch + 1 - e < ISSUE_GROUP ? nch + 1 - e : ISSUE_GROUP;
                        uint32_t stg[ISSUE_GROUP], a_lo[ISSUE_GROUP], b_lo[ISSUE_GROUP];
#pragma unroll
                        for (int i = 0; i < ISSUE_GROUP; ++i) {
                            if (i < n_in) {
                                stg[i] = st;
                                ptx::mbar_wait(&w_full[(cyc & 1) * NSTAGE + st], (cyc >> 1) & 1);
                                if (++st == NSTAGE) { st = 0; ++cyc; }
                            } else {
                                stg[i] = 0;
                            }
                            const int c = e + i;
                            // which A tile region does this chunk multiply?
                            const bool from_enc = (l == 0) || (c >= 8);
                            a_lo[i] = c >= nch ? ones_lo
                                               : (from_enc ? enc_lo : act_lo) + (uint32_t)((l == 0 || c < 8) ? c : c - 8) * (CHUNK_K / 8) * (KROW >> 4);
                            b_lo[i] = (ring_a + stg[i] * (HSTAGE_BYTES >> 4)) | b_lbo;
                        }
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
#pragma unroll
                            for (int i = 0; i < ISSUE_GROUP; ++i) {
                                const int c = e + i;
                                if (i < n_in) {
                                    ptx::mma2_bf16_w(tacc, a_lo[i], desc_hi, b_lo[i], desc_hi, idesc, c != 0);
                                    if (c < nch) ptx::mma2_bf16_w(tacc, a_lo[i] + 2 * (KROW >> 4), desc_hi, b_lo[i] + b_kstep, desc_hi, idesc, 1u);
                                    ptx::mma2_commit(&w_empty[stg[i]]);
                                    if (c == nch) ptx::mma2_commit(&acc_full[s]);
                                }
                            }
                        }
                        __syncwarp();
                    }
                    if (s == 0 && both) skip(nch + 1);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups (one thread per sample row) =================
        const int slot = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) | lane;
        uint8_t* act = smem + SM_ACT + slot * ACT_BYTES;
        uint8_t* enc = smem + SM_ENC + slot * ENC_BYTES;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + slot * WIDTH;
        const uint32_t ready_bar = ptx::mapa(&a_ready[slot], 0);     // the leader's barrier
        uint32_t full_uses = 0;
        for (int64_t k = 0; hq_of(k, slot) < nhq; ++k) {
            const int64_t tile = hq_of(k, slot) * 2 + rank;
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
                write_enc_row(enc, save_tile ? save_tile + SV_ENC : nullptr, row, x, bw3, valid);
            }
            ptx::fence_proxy_async();
            ptx::warp_arrive_cluster(ready_bar);
            for (int l = 0; l < NLAYER; ++l, ++full_uses) {
                ptx::mbar_wait_fast(&acc_full[slot], full_uses & 1);
                ptx::tc_fence_after();
                uint32_t va[32], vb[32], pk[16];
                if (l < 8) {
                    uint8_t* save_img = save_tile ? save_tile + SV_H + (int64_t)l * ACT_BYTES : nullptr;
                    uint32_t* flags = mask_tile ? mask_tile + l * MASK_WORDS * TILE : nullptr;
                    float sig_acc = 0.f;
                    ptx::tmem_ld32(tacc, va);
#pragma unroll 1
                    for (int c2 = 0; c2 < WIDTH / 64; ++c2) {
                        // TMEM loads run one chunk ahead of the arithmetic
                        ptx::tmem_ld_wait();
                        ptx::tmem_ld32(tacc + (2 * c2 + 1) * 32, vb);
                        epilogue_chunk(va, 2 * c2, row, WIDTH, act, save_img, flags, pk);
                        if (l == 6) {   // density head: row 0 of layer 7 applied to h6 (nerf.py:427)
                            const float4* w = reinterpret_cast<const float4*>(cst + C_W7R0 + (2 * c2) * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 w4 = w[q];
                                sig_acc += w4.x * bf16_lo(pk[2 * q]) + w4.y * bf16_hi(pk[2 * q]) +
                                           w4.z * bf16_lo(pk[2 * q + 1]) + w4.w * bf16_hi(pk[2 * q + 1]);
                            }
                        }
                        ptx::tmem_ld_wait();
                        if (c2 + 1 < WIDTH / 64) ptx::tmem_ld32(tacc + (2 * c2 + 2) * 32, va);
                        epilogue_chunk(vb, 2 * c2 + 1, row, WIDTH, act, save_img, flags, pk);
                        if (l == 6) {
                            const float4* w = reinterpret_cast<const float4*>(cst + C_W7R0 + (2 * c2 + 1) * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 w4 = w[q];
                                sig_acc += w4.x * bf16_lo(pk[2 * q]) + w4.y * bf16_hi(pk[2 * q]) +
                                           w4.z * bf16_lo(pk[2 * q + 1]) + w4.w * bf16_hi(pk[2 * q + 1]);
                            }
                        }
                    }
                    if (l == 7)   // A columns 256..287 of rgb0
                        write_venc_row(enc, save_tile ? save_tile + SV_VENC : nullptr, row, v3, bwv, valid);
                    ptx::tc_fence_before();
                    ptx::fence_proxy_async();
                    ptx::warp_arrive_cluster(ready_bar);
                    if (l == 6 && valid) {
                        float pre = sig_acc + cst[C_MISC];
                        sigma_out[g] = softplus_f(pre);
                        if (sig_pre) sig_pre[g] = pre;
                    }
                } else {
                    // rgb0 epilogue: hr = relu(.), rgb = sigmoid(W_rgb1 hr + b)   (nerf.py:442-446)
                    float o0 = cst[C_MISC + 1], o1 = cst[C_MISC + 2], o2 = cst[C_MISC + 3];
                    uint8_t* save_img = save_tile ? save_tile + SV_HR : nullptr;
                    uint32_t* flags = mask_tile ? mask_tile + 8 * MASK_WORDS * TILE : nullptr;
#pragma unroll 1
                    for (int cc = 0; cc < RGBW / 32; ++cc) {
                        ptx::tmem_ld32(tacc + cc * 32, va);
                        ptx::tmem_ld_wait();
                        epilogue_chunk(va, cc, row, RGBW, nullptr, save_img, flags, pk);
                        const float4* w0 = reinterpret_cast<const float4*>(cst + C_WRGB1 + cc * 32);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float a = bf16_lo(pk[2 * q]), b = bf16_hi(pk[2 * q]), c = bf16_lo(pk[2 * q + 1]), d = bf16_hi(pk[2 * q + 1]);
                            const float4 x0 = w0[q], x1 = w0[RGBW / 4 + q], x2 = w0[2 * RGBW / 4 + q];
                            o0 += x0.x * a + x0.y * b + x0.z * c + x0.w * d;
                            o1 += x1.x * a + x1.y * b + x1.z * c + x1.w * d;
                            o2 += x2.x * a + x2.y * b + x2.z * c + x2.w * d;
                        }
                    }
                    ptx::tc_fence_before();   // TMEM reads done before the next pass overwrites the accumulator
                    if (valid) {
                        float r0 = sigmoid_f(o0), r1 = sigmoid_f(o1), r2 = sigmoid_f(o2);
                        rgb_out[g * 3] = r0; rgb_out[g * 3 + 1] = r1; rgb_out[g * 3 + 2] = r2;
                        if (rgb_keep) { rgb_keep[g * 3] = r0; rgb_keep[g * 3 + 1] = r1; rgb_keep[g * 3 + 2] = r2; }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();            // neither CTA leaves while its peer may still touch its shared memory / TMEM
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

}  // namespace tc

// ------------------------------------------------------------------------------------------
// host entry points
// ------------------------------------------------------------------------------------------

size_t tc_workspace_bytes(int64_t R, int N, int training) { return tc::carve(nullptr, R * (int64_t)N, training != 0).bytes; }

// fp32 parameters -> the BF16 weight streams + fp32 constants of the workspace.  Independent of the rays: callers may run
// it early, on another stream, while the rays are still being produced (niw_nerf_pack).
int tc_pack(const float* P, const C2F& c2f, int training, int64_t R, int N, void* ws, size_t ws_bytes, cudaStream_t st) {
    using namespace tc;
    Workspace w = carve(ws, R * (int64_t)N, training != 0);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    niw::note_launch(), pack_weights_kernel<<<niw_blocks(STREAM_BYTES / 16, 256), 256, 0, st>>>(P, c2f, w.wstream, w.consts);
    if (training)
        niw::note_launch(), pack_weights_bwd_kernel<<<niw_blocks(BSTREAM_BYTES / 16, 256), 256, 0, st>>>(P, w.bstream);
    NIW_LAUNCH_CHECK();
    return 0;
}

int tc_fwd(const float* P, const float* center, const float* ray, const float* depth, int64_t R, int N, const C2F& c2f,
           int training, void* ws, size_t ws_bytes, float* rgb, float* sigma, cudaStream_t st) {
    using namespace tc;
    const int64_t S = R * (int64_t)N;
    const bool prepacked = (training & NIW_NERF_PREPACKED) != 0;     // tc_pack already ran on this workspace
    training &= 1;
    Workspace w = carve(ws, S, training != 0);
    if (ws_bytes < w.bytes) return NIW_E_WORKSPACE;
    if (!prepacked) {
        int e = tc_pack(P, c2f, training, R, N, ws, ws_bytes, st);
        if (e) return e;
    }
    NIW_CUDA(cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    const int64_t nhq = ((S + TILE - 1) / TILE + 1) / 2;        // tile pairs: one per slot of a CTA pair and round
    int64_t pairs = niw_num_sms() / 2;
    if (pairs > nhq) pairs = nhq;
    const int grid = (int)(2 * (pairs < 1 ? 1 : pairs));
    niw::note_launch(), tc_fwd_kernel<<<grid, 384, SM_TOTAL, st>>>(w.wstream, w.consts, center, ray, depth, S, N, rgb, sigma,
                                             training ? w.sig_pre : nullptr, training ? w.rgb_keep : nullptr,
                                             training ? w.save : nullptr);
    NIW_LAUNCH_CHECK();
    return 0;
}

}  // namespace niw
