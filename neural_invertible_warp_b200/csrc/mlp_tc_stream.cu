// Streaming ("layer-outer") form of the fused NeRF MLP kernels for the TRAINING step (tcgen05, sm_100a).
// SURVEY.md section 8 rows a6+a7+a8; reference model/nerf.py:416-456, model/barf.py:256-268.
//
// The slot-form kernels (mlp_tc.cu / mlp_tc_bwd.cu) walk a tile through all layers: the activations never leave
// shared memory, but every layer of a tile waits for the previous one (MMA -> accumulator drain -> re-pack -> MMA),
// and TMEM holds only two 256-column accumulators, so at most two such chains run per SM and the tensor pipe idles
// about half the time.  In training mode every activation image is written to the tile record
// anyway (the weight-gradient pass reads it), so here the loops are interchanged:
//
//     for layer l:   (this CTA pair's half of W_l stays RESIDENT in shared memory, no weight stream)
//         for every tile pair u owned by this CTA pair:
//             A = h_{l-1}(u) comes back from the tile record with bulk async copies (L2 hits: it was written a few
//                 microseconds ago by this very CTA), D = A . W_l^T lands in one of two TMEM accumulators,
//                 the epilogue warps turn it into h_l(u) and store it straight to the record
//
// Consecutive MMAs belong to different tiles and do not depend on each other: tile u+1 is multiplied while tile u
// drains, and nothing of the chain MMA -> epilogue -> shared-memory store -> proxy fence -> cluster arrive ->
// MMA is left on the critical path.  L2 traffic is the same as in the slot form (64 KB of operand per tile and
// layer and CTA: there the weights, here the activations).
//
// Warp roles (384 threads, CTA pairs, tcgen05 cta_group::2, M = 256 = the pair's two 128-sample tiles):
//   warp 0      operand producer: K = 128 slices (32 KB) of this CTA's tile images -> 4-unit ring (bulk async copies);
//               a slice is requested only after the epilogue that wrote it has published it (done[] counters)
//   warp 1      leader: MMA issuer (one elected lane);  peer: relays "my slice has landed" to the leader
//   warp 2      TMEM allocator;  peer: relays "my half of the weights has landed"
//   warp 3      constants, the "ones" image of the bias product, then weight producer (once per layer)
//   warps 4-7   epilogue of accumulator 0,  warps 8-11 epilogue of accumulator 1 (one thread per sample row)
#include "tc_layout.cuh"
#include "tc_epilogue.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace niw {
namespace tc {

#ifndef NIW_SF_UNITK
#define NIW_SF_UNITK 128
#endif
constexpr int SF_UNITK = NIW_SF_UNITK;                   // K extent of one ring unit
constexpr int SF_UNIT = TILE * SF_UNITK * 2;             // 32768: K = 128 columns of a 128-row image (16 groups x 2 KB)
constexpr int SF_NUNIT = 128 * 1024 / SF_UNIT;
constexpr int SF_HP = WIDTH / SF_UNITK;                  // ring units of a 256-column image
constexpr int SF_REG = 8192;                             // weight buffer regions, recycled one by one at a layer change
constexpr int SF_NREG = 11;                              // layer 4 (K = 320): 10 chunks x 8 KB + 4 KB of bias chunk
constexpr int SF_RING = 0;
constexpr int SF_W = SF_RING + SF_NUNIT * SF_UNIT;       // this CTA's N half of the current layer's weights (+ bias chunk)
constexpr int SF_ONES = SF_W + SF_NREG * SF_REG;
constexpr int SF_CONST = SF_ONES + ONES_BYTES;
constexpr int SF_BAR = SF_CONST + ((C_FLOATS * 4 + 15) / 16) * 16;
constexpr int SF_TOTAL = SF_BAR + 384;
static_assert(SF_TOTAL <= 227 * 1024, "shared memory budget (streaming forward)");

// operand pieces of forward layer l: which record image, how many bytes, how many K = 16 steps
__host__ __device__ constexpr int sf_pieces(int l) { return l == 0 ? 1 : ((l == SKIP || l == 8) ? SF_HP + 1 : SF_HP); }
__host__ __device__ constexpr int64_t sf_piece_off(int l, int p) {
    return l == 0 ? SV_ENC : (p < SF_HP ? SV_H + (int64_t)(l - 1) * ACT_BYTES + (int64_t)p * SF_UNIT : (l == SKIP ? SV_ENC : SV_VENC));
}
__host__ __device__ constexpr int sf_piece_bytes(int l, int p) {
    return l == 0 ? ENC_BYTES : (p < SF_HP ? SF_UNIT : (l == SKIP ? ENC_BYTES : VENC_BYTES));
}
__host__ __device__ constexpr int sf_piece_ksteps(int l, int p) { return sf_piece_bytes(l, p) / (2 * KROW); }
// weight bytes of layer l held by one CTA, the regions they fill, and how many earlier (band, layer) phases used region r
__host__ __device__ constexpr int sf_w_bytes(int l) { return (layer_chunks(l) * CHUNK_K + BIAS_K) * (layer_rows(l) / 2) * 2; }
__host__ __device__ constexpr int sf_nreg(int l) { return (sf_w_bytes(l) + SF_REG - 1) / SF_REG; }
__host__ __device__ constexpr uint32_t sf_reg_uses(int r, int64_t band, int l) {
    int per_band = 0, before = 0;
    for (int i = 0; i < NLAYER; ++i) {
        if (sf_nreg(i) > r) { ++per_band; if (i < l) ++before; }
    }
    return (uint32_t)(band * per_band + before);
}

// bulk copy global -> shared with an L2 evict-first hint: an activation image is read back exactly once by the forward pass
// (the weight-gradient pass streams it from HBM much later), so the line should be the first to go once it has been read
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(ptx::smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(ptx::smem_addr(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_shared(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_addr(p)) : "memory");
    return v;
}
// the calling warp's record stores of one tile are complete: make them visible to the bulk-copy (async) proxy of this
// CTA's producer and count the warp in
__device__ __forceinline__ void publish_tile(uint32_t* done_ctr) {
    __threadfence();
    fence_proxy_async_global();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicAdd(done_ctr, 1u);
}
__device__ __forceinline__ void wait_published(const uint32_t* done, int64_t seq) {
    const uint32_t need = 4u * (uint32_t)((seq >> 1) + 1);
    const uint32_t* ctr = done + (seq & 1);
    while (ld_volatile_shared(ctr) < need) __nanosleep(20);
    __threadfence();
    fence_proxy_async_global();
}

// DBG: CTA 0 records clock64() at the hand-offs of every work item (seq): dbg[seq * 16 + k], k = 0 producer dependency
// met, 1 slices requested, 2 accumulator free, 3 operands landed + MMAs issued, 4 committed, 5 accumulator full (epilogue),
// 6 drained, 7 stored, 8 published, 9 first weight region landed (first tile of a phase)
#define SF_STAMP(seq, k) do { if (DBG && blockIdx.x == 0 && dbg) dbg[(seq) * 16 + (k)] = clock64(); } while (0)

// Work items of a CTA pair: its T tile pairs are cut into bands of `band` tile pairs; inside a band the layers are the outer
// loop, so an image written in layer l is read back in layer l+1 after at most `band` tile pairs of every CTA pair
// (148 x band x 128 KB of L2 footprint per layer: stays L2-resident).  Items are numbered in processing order
// (the T encoding items first); item seq uses accumulator / epilogue warpgroup seq & 1.
template <bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
tc_fwd_stream_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ consts_g, const float* __restrict__ center,
                     const float* __restrict__ ray, const float* __restrict__ depth, int64_t S, int N, int band, int evict_first,
                     float* __restrict__ rgb_out, float* __restrict__ sigma_out, float* __restrict__ sig_pre,
                     float* __restrict__ rgb_keep, uint8_t* __restrict__ save, long long* __restrict__ dbg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SF_BAR);
    uint64_t* a_full = bars;                      // [SF_NUNIT] this CTA's slice has landed (leader: and the peer's)
    uint64_t* a_empty = a_full + SF_NUNIT;        // [SF_NUNIT] the MMAs reading the unit have completed (both CTAs)
    uint64_t* w_full = a_empty + SF_NUNIT;        // [SF_NREG]  the weights of a region have landed (leader: in both CTAs)
    uint64_t* w_empty = w_full + SF_NREG;         // [SF_NREG]  the phase's last MMA on the region has completed (both CTAs)
    uint64_t* acc_full = w_empty + SF_NREG;       // [2]  a tile pair has been accumulated (both CTAs)
    uint64_t* acc_empty = acc_full + 2;           // [2]  (leader) both CTAs have drained the accumulator
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint32_t* done = tmem_slot + 2;               // [2]  epilogue warps x items published, per accumulator
    float* cst = reinterpret_cast<float*>(smem + SF_CONST);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int64_t npairs = ((S + TILE - 1) / TILE + 1) / 2;
    const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
    const int64_t T = pair0 < npairs ? (npairs - pair0 + pair_step - 1) / pair_step : 0;   // tile pairs of this CTA pair
    const int64_t nbands = (T + band - 1) / band;
    auto tile_of = [&](int64_t i) { return 2 * (pair0 + i * pair_step) + rank; };
    auto band_size = [&](int64_t b) { const int64_t n = T - b * band; return n < band ? n : (int64_t)band; };

    if (threadIdx.x == 0) {
        for (int i = 0; i < SF_NUNIT; ++i) { ptx::mbar_init(&a_full[i], rank == 0 ? 2 : 1); ptx::mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < SF_NREG; ++i) { ptx::mbar_init(&w_full[i], rank == 0 ? 2 : 1); ptx::mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], 2 * TILE / 32); }
        done[0] = 0; done[1] = 0;
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc2(tmem_slot, 512);
    if (warp == 3) {
        for (int i = lane; i < C_FLOATS; i += 32) cst[i] = consts_g[i];
        uint4* ones = reinterpret_cast<uint4*>(smem + SF_ONES);
        for (int i = lane; i < ONES_BYTES / 16; i += 32) ones[i] = i < TILE ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= operand producer =================
        if (lane == 0) {
            uint32_t slot = 0, cyc = 0;
            int64_t seq = T;
            const uint64_t pol = l2_policy_evict_first();
            for (int64_t b = 0; b < nbands; ++b) {
                const int64_t nb = band_size(b);
                for (int l = 0; l < NLAYER; ++l)
                    for (int64_t ii = 0; ii < nb; ++ii, ++seq) {
                        const int64_t i = b * band + ii;
                        wait_published(done, l == 0 ? i : seq - nb);       // the item that wrote these images
                        SF_STAMP(seq, 0);
                        const uint8_t* rec = save + tile_of(i) * SAVE_TILE_BYTES;
                        for (int p = 0; p < sf_pieces(l); ++p) {
                            const uint32_t bytes = (uint32_t)sf_piece_bytes(l, p);
                            ptx::mbar_wait(&a_empty[slot], (cyc & 1) ^ 1);
                            ptx::mbar_arrive_expect_tx(&a_full[slot], bytes);
                            if (evict_first & 1) bulk_g2s_hint(smem + SF_RING + slot * SF_UNIT, rec + sf_piece_off(l, p), bytes, &a_full[slot], pol);
                            else ptx::bulk_g2s(smem + SF_RING + slot * SF_UNIT, rec + sf_piece_off(l, p), bytes, &a_full[slot]);
                            if (++slot == SF_NUNIT) { slot = 0; ++cyc; }
                        }
                        SF_STAMP(seq, 1);
                    }
            }
        }
    } else if (warp == 1 && rank != 0) {
        // ================= peer: relay "slice landed" to the leader =================
        if (lane == 0) {
            uint32_t slot = 0, cyc = 0;
            const uint32_t full0 = ptx::mapa(&a_full[0], 0);
            for (int64_t b = 0; b < nbands; ++b)
                for (int l = 0; l < NLAYER; ++l)
                    for (int64_t n = band_size(b) * sf_pieces(l); n > 0; --n) {
                        ptx::mbar_wait(&a_full[slot], cyc & 1);
                        ptx::mbar_arrive_cluster(full0 + slot * 8);
                        if (++slot == SF_NUNIT) { slot = 0; ++cyc; }
                    }
        }
    } else if (warp == 2 && rank != 0) {
        // ================= peer: relay "weight region landed" to the leader =================
        if (lane == 0) {
            const uint32_t wfull0 = ptx::mapa(&w_full[0], 0);
            for (int64_t b = 0; b < nbands; ++b)
                for (int l = 0; l < NLAYER; ++l) {
                    ptx::mbar_wait(&w_full[0], (uint32_t)(b * NLAYER + l) & 1);
                    ptx::mbar_arrive_cluster(wfull0);
                }
        }
    } else if (warp == 3) {
        // ================= weight producer: this CTA's half of every chunk of the layer, region by region =================
        // a region is refilled as soon as the previous phase's last tile pair is through with it, so the change of layer
        // overlaps that tile pair's MMAs instead of draining the pipeline
        if (lane == 0) {
            for (int64_t b = 0; b < nbands; ++b) {
                const uint8_t* lsrc = wstream;
                for (int l = 0; l < NLAYER; ++l) {
                    const int nch = layer_chunks(l), hrows = layer_rows(l) / 2;
                    const uint32_t cb = (uint32_t)hrows * CHUNK_K * 2, bb = (uint32_t)hrows * BIAS_K * 2, total = (uint32_t)nch * cb + bb;
                    int c = 0;
                    uint32_t off = 0;
                    for (int r = 0; r < sf_nreg(l); ++r) {
                        const uint32_t lim = (uint32_t)(r + 1) * SF_REG;
                        ptx::mbar_wait(&w_empty[r], (sf_reg_uses(r, b, l) & 1) ^ 1);
                        if (r == 0) ptx::mbar_arrive_expect_tx(&w_full[0], total);      // one "landed" barrier per phase
                        while (c <= nch && off < lim) {
                            const uint32_t sz = c < nch ? cb : bb;
                            ptx::bulk_g2s(smem + SF_W + off, lsrc + (int64_t)c * 2 * cb + rank * sz, sz, &w_full[0]);
                            off += sz; ++c;
                        }
                    }
                    lsrc += layer_stream_bytes(l);
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ================= leader: MMA issuer (warp converged, one elected lane issues) =================
        uint32_t slot = 0, cyc = 0, acc_uses[2] = {0u, 0u};
        const uint32_t ring_a = ptx::smem_addr(smem + SF_RING) >> 4;
        const uint32_t w_a = ptx::smem_addr(smem + SF_W) >> 4;
        const uint32_t ones_lo = ptx::smem_desc_lo(ptx::smem_addr(smem + SF_ONES), KROW);
        const uint32_t desc_hi = ptx::smem_desc_hi(128);
        const uint32_t a_lbo = (uint32_t)(KROW >> 4) << 16;
        int64_t seq = T;
        for (int64_t b = 0; b < nbands; ++b) {
            const int64_t nb = band_size(b);
            for (int l = 0; l < NLAYER; ++l) {
                const int hrows = layer_rows(l) / 2;
                const uint32_t idesc = ptx::idesc_bf16(2 * TILE, 2 * hrows, 0, 0);
                const uint32_t b_lbo = (uint32_t)hrows << 16, b_kstep = (uint32_t)hrows * 2, kbytes = (uint32_t)hrows * 32;
                for (int64_t ii = 0; ii < nb; ++ii, ++seq) {
                    const bool first = ii == 0, last = ii == nb - 1;
                    const uint32_t a = (uint32_t)(seq & 1);
                    ptx::mbar_wait_fast(&acc_empty[a], (acc_uses[a] & 1) ^ 1);
                    ++acc_uses[a];
                    ptx::tc_fence_after();
                    if (lane == 0) SF_STAMP(seq, 2);
                    const uint32_t tacc = tmem_base + a * WIDTH;
                    // Everything stays warp-uniform (loop state and descriptors live in uniform registers); only the tcgen05
                    // instructions sit in an elected region, ONE per operand piece: every elect / reconverge round trip costs
                    // ~200 clk of issue time, and a region that computes its own operands pays R2UR moves per MMA
                    uint32_t kstep = 0, w_freed = 0;      // regions [0, w_freed) are released
                    const int np = sf_pieces(l);
                    for (int p = 0; p < np; ++p) {
                        const bool lastp = p == np - 1;      // the bias product rides behind the last piece:
                        const int ks = sf_piece_ksteps(l, p); // D += ones[256 x 16] . [bf16(b), b - bf16(b), 0 ...]^T
                        ptx::mbar_wait(&a_full[slot], cyc & 1);
                        if (first && p == 0) {      // the layer's weights: have they landed (in both CTAs)?
                            ptx::mbar_wait(&w_full[0], (uint32_t)(b * NLAYER + l) & 1);
                            if (lane == 0) SF_STAMP(seq, 9);
                        }
                        ptx::tc_fence_after();
                        const uint32_t a_lo = (ring_a + slot * (SF_UNIT >> 4)) | a_lbo;
                        // last tile pair of the phase: it is through with these regions, the next layer's weights may land in them
                        const uint32_t upto = !last ? w_freed : (lastp ? (uint32_t)sf_nreg(l) : ((kstep + ks) * kbytes) / SF_REG);
                        if (ptx::elect_one()) {
                            for (int kk = 0; kk < ks; ++kk)
                                ptx::mma2_bf16_w(tacc, a_lo + kk * 2 * (KROW >> 4), desc_hi, (w_a + (kstep + kk) * b_kstep) | b_lbo, desc_hi,
                                                 idesc, (kstep + kk) != 0);
                            ptx::mma2_commit(&a_empty[slot]);
                            if (lastp) {
                                ptx::mma2_bf16_w(tacc, ones_lo, desc_hi, (w_a + (kstep + ks) * b_kstep) | b_lbo, desc_hi, idesc, 1u);
                                ptx::mma2_commit(&acc_full[a]);
                            }
                            for (uint32_t r = w_freed; r < upto; ++r) ptx::mma2_commit(&w_empty[r]);
                        }
                        __syncwarp();
                        w_freed = upto;
                        kstep += ks;
                        if (++slot == SF_NUNIT) { slot = 0; ++cyc; }
                    }
                    if (lane == 0) SF_STAMP(seq, 3);
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups (one thread per sample row) =================
        const int wg = (warp - 4) >> 2;
        const int row = ((warp & 3) << 5) | lane;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + wg * WIDTH;
        const uint32_t empty_bar = ptx::mapa(&acc_empty[wg], 0);     // the leader's barrier
        // images that the next layer reads back are stored with L2 evict-last priority (and read with evict-first)
        const uint64_t spol = (evict_first & 2) ? l2_policy_evict_last() : 0ull;
        Bands3 bw3; BandsV bwv;
        load_bands(cst + C_BANDS, bw3, bwv);
        // ---- encoding items: positional encodings of the points and the view directions into the tile records ----
        for (int64_t i = wg; i < T; i += 2) {
            const int64_t tile = tile_of(i), g = tile * TILE + row;
            const bool valid = g < S;
            uint8_t* rec = save + tile * SAVE_TILE_BYTES;
            float v3[3] = {0.f, 0.f, 1.f}, x[3] = {0.f, 0.f, 0.f};
            if (valid) {
                const int64_t r = g / N;
                const float d = depth[g];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v3[c] = ray[r * 3 + c];
                    x[c] = __fadd_rn(center[r * 3 + c], __fmul_rn(v3[c], d));
                }
            }
            write_enc_row(nullptr, rec + SV_ENC, row, x, bw3, valid);
            write_venc_row(nullptr, rec + SV_VENC, row, v3, bwv, valid);
            publish_tile(&done[wg]);
        }
        uint32_t full_uses = 0;
        int64_t seq = T;
        for (int64_t b = 0; b < nbands; ++b) {
            const int64_t nb = band_size(b);
            for (int l = 0; l < NLAYER; ++l) {
                for (int64_t ii = 0; ii < nb; ++ii, ++seq) {
                    if ((seq & 1) != wg) continue;
                    const int64_t tile = tile_of(b * band + ii), g = tile * TILE + row;
                    const bool valid = g < S;
                    uint8_t* rec = save + tile * SAVE_TILE_BYTES;
                    uint32_t* mask_tile = reinterpret_cast<uint32_t*>(rec + SV_MASK);
                    ptx::mbar_wait_fast(&acc_full[wg], full_uses & 1);
                    ++full_uses;
                    ptx::tc_fence_after();
                    if ((warp & 3) == 0 && lane == 0) SF_STAMP(seq, 5);
                    uint32_t va[32], vb[32], pk[16];
                    if (l < 8) {
                        uint8_t* save_img = rec + SV_H + (int64_t)l * ACT_BYTES;
                        uint32_t* flags = mask_tile + l * MASK_WORDS * TILE;
                        float sig_acc = 0.f;
                        ptx::tmem_ld32(tacc, va);
#pragma unroll 1
                        for (int c2 = 0; c2 < WIDTH / 64; ++c2) {
                            ptx::tmem_ld_wait();
                            ptx::tmem_ld32(tacc + (2 * c2 + 1) * 32, vb);
                            epilogue_chunk(va, 2 * c2, row, WIDTH, nullptr, save_img, flags, pk, spol);
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
                            if (c2 + 1 < WIDTH / 64) {
                                ptx::tmem_ld32(tacc + (2 * c2 + 2) * 32, va);
                            } else {
                                ptx::tc_fence_before();                     // accumulator drained: the next tile pair may be multiplied
                                ptx::warp_arrive_cluster(empty_bar);
                                if ((warp & 3) == 0 && lane == 0) SF_STAMP(seq, 6);
                            }
                            epilogue_chunk(vb, 2 * c2 + 1, row, WIDTH, nullptr, save_img, flags, pk, spol);
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
                        if (l == 6 && valid) {
                            float pre = sig_acc + cst[C_MISC];
                            sigma_out[g] = softplus_f(pre);
                            if (sig_pre) sig_pre[g] = pre;
                        }
                    } else {
                        // rgb0 epilogue: hr = relu(.), rgb = sigmoid(W_rgb1 hr + b)   (nerf.py:442-446)
                        float o0 = cst[C_MISC + 1], o1 = cst[C_MISC + 2], o2 = cst[C_MISC + 3];
                        uint8_t* save_img = rec + SV_HR;
                        uint32_t* flags = mask_tile + 8 * MASK_WORDS * TILE;
#pragma unroll 1
                        for (int cc = 0; cc < RGBW / 32; ++cc) {
                            ptx::tmem_ld32(tacc + cc * 32, va);
                            ptx::tmem_ld_wait();
                            if (cc == RGBW / 32 - 1) {
                                ptx::tc_fence_before();
                                ptx::warp_arrive_cluster(empty_bar);
                            }
                            epilogue_chunk(va, cc, row, RGBW, nullptr, save_img, flags, pk);
                            const float4* w0 = reinterpret_cast<const float4*>(cst + C_WRGB1 + cc * 32);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float a = bf16_lo(pk[2 * q]), b2 = bf16_hi(pk[2 * q]), c = bf16_lo(pk[2 * q + 1]), d = bf16_hi(pk[2 * q + 1]);
                                const float4 x0 = w0[q], x1 = w0[RGBW / 4 + q], x2 = w0[2 * RGBW / 4 + q];
                                o0 += x0.x * a + x0.y * b2 + x0.z * c + x0.w * d;
                                o1 += x1.x * a + x1.y * b2 + x1.z * c + x1.w * d;
                                o2 += x2.x * a + x2.y * b2 + x2.z * c + x2.w * d;
                            }
                        }
                        if (valid) {
                            float r0 = sigmoid_f(o0), r1 = sigmoid_f(o1), r2 = sigmoid_f(o2);
                            rgb_out[g * 3] = r0; rgb_out[g * 3 + 1] = r1; rgb_out[g * 3 + 2] = r2;
                            if (rgb_keep) { rgb_keep[g * 3] = r0; rgb_keep[g * 3 + 1] = r1; rgb_keep[g * 3 + 2] = r2; }
                        }
                    }
                    if ((warp & 3) == 0 && lane == 0) SF_STAMP(seq, 7);
                    publish_tile(&done[wg]);       // the next layer reads this image back (every item counts, rgb0 too)
                    if ((warp & 3) == 0 && lane == 0) SF_STAMP(seq, 8);
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

// training-mode forward in streaming form (same outputs, same tile records as tc_fwd_kernel; the workspace holds the
// packed weights already)
int tc_fwd_stream(const tc::Workspace& w, const float* center, const float* ray, const float* depth, int64_t S, int N,
                  float* rgb, float* sigma, cudaStream_t st) {
    using namespace tc;
    const int64_t npairs = ((S + TILE - 1) / TILE + 1) / 2;
    int64_t pairs = niw_num_sms() / 2;
    if (pairs > npairs) pairs = npairs;
    const int grid = (int)(2 * (pairs < 1 ? 1 : pairs));
    static const int band = getenv("NIW_STREAM_BAND") && atoi(getenv("NIW_STREAM_BAND")) > 0 ? atoi(getenv("NIW_STREAM_BAND")) : 8;
    static const int evict_first = getenv("NIW_STREAM_EVICT") ? atoi(getenv("NIW_STREAM_EVICT")) : 1;
    // NIW_STREAM_DEBUG=1: hand-off time stamps of CTA 0 on stderr (synchronises; diagnostics only)
    static const bool debug = getenv("NIW_STREAM_DEBUG") != nullptr;
    if (debug) {
        const int64_t T = (npairs + pairs - 1) / pairs, nseq = (NLAYER + 1) * T;
        long long* dbg = nullptr;
        NIW_CUDA(cudaMalloc(&dbg, sizeof(long long) * 16 * nseq));
        NIW_CUDA(cudaMemsetAsync(dbg, 0, sizeof(long long) * 16 * nseq, st));
        NIW_CUDA(cudaFuncSetAttribute(tc_fwd_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_TOTAL));
        tc_fwd_stream_kernel<true><<<grid, 384, SF_TOTAL, st>>>(w.wstream, w.consts, center, ray, depth, S, N, band, evict_first, rgb, sigma,
                                                                 w.sig_pre, w.rgb_keep, w.save, dbg);
        NIW_LAUNCH_CHECK();
        std::vector<long long> h(16 * nseq);
        NIW_CUDA(cudaStreamSynchronize(st));
        NIW_CUDA(cudaMemcpy(h.data(), dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        cudaFree(dbg);
        long long t0 = 0;
        for (int64_t q = T; q < nseq; ++q) if (h[q * 16 + 0] && (!t0 || h[q * 16 + 0] < t0)) t0 = h[q * 16 + 0];
        fprintf(stderr, "stream fwd, CTA 0, T = %lld tile pairs; clocks since the first request\n"
                        " seq | dep_ok  request | acc_free   issued        - | acc_full  drained   stored  published | w_landed\n", (long long)T);
        for (int64_t q = T; q < nseq; ++q) {
            fprintf(stderr, "%4lld |", (long long)q);
            for (int k = 0; k < 10; ++k) {
                if (k == 2 || k == 5 || k == 9) fprintf(stderr, " |");
                if (h[q * 16 + k]) fprintf(stderr, " %8lld", h[q * 16 + k] - t0); else fprintf(stderr, "        -");
            }
            fprintf(stderr, "\n");
        }
        return 0;
    }
    NIW_CUDA(cudaFuncSetAttribute(tc_fwd_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_TOTAL));
    niw::note_launch(), tc_fwd_stream_kernel<false><<<grid, 384, SF_TOTAL, st>>>(w.wstream, w.consts, center, ray, depth, S, N, band, evict_first, rgb, sigma,
                                                                                w.sig_pre, w.rgb_keep, w.save, nullptr);
    NIW_LAUNCH_CHECK();
    return 0;
}

}  // namespace niw
