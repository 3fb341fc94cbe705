/*
 * niw_b200.h -- C ABI of the B200-native NeRF ray-render path for the invertible-warp BARF
 * codebase (sfchng/neural_invertible_warp).
 *
 * The reference has NO native boundary (it is eager PyTorch); each entry point below replaces a
 * chain of ATen ops inside one reference Python function, cited as reference file:line.  The
 * Python mirror of the reference's Graph/NeRF API (the .py files of neural_invertible_warp_b200) binds these
 * symbols with ctypes and calls them from torch.autograd.Function.forward/backward; see
 * INTEGRATION.md for the binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous row-major fp32 array unless noted;
 *   - the caller owns and allocates every buffer, including workspaces (sizes are queried);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - return value 0 = success; >0 = a cudaError_t; <0 = NIW_E_* argument error.  Nothing throws.
 *     niw_error_string() turns a code into text.
 *   - no hidden state apart from idempotent function attributes (max dynamic shared memory).
 *   - B = images, P = rays per image, R = B*P rays, N = samples per ray, S = R*N samples.
 */
#ifndef NIW_B200_H_
#define NIW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIW_ABI_VERSION 6

#define NIW_E_BADARG   (-1)  /* null pointer / non-positive size */
#define NIW_E_UNSUPP   (-2)  /* shape or option outside what the kernels implement */
#define NIW_E_WORKSPACE (-3) /* workspace too small */

/* precision of the MLP GEMMs */
#define NIW_PREC_FP32 0 /* CUDA-core fp32 (parity / high-precision path)          */
#define NIW_PREC_BF16 1 /* tcgen05 BF16 operands, FP32 accumulate in TMEM (fast)  */
#define NIW_PREC_BF16X3 2 /* tcgen05, every operand a hi + lo BF16 pair and every product 3 MMAs (hi.hi + lo.hi + hi.lo), FP32
                           * accumulate: ~fp32 operand accuracy on the tensor cores -- the path that meets the 1e-3 max-abs
                           * contract on rendered rgb / depth / opacity.  Forward only; niw_nerf_bwd runs the BF16 backward
                           * on the tile records this forward saves (gradient contract: 1e-2 relative, BF16 operands). */

int niw_abi_version(void);
const char* niw_error_string(int code);
/* number of kernels this library has launched since it was loaded (monotonic; all streams) */
unsigned long long niw_launch_count(void);

/* ---- (a1) camera.get_center_and_ray + [:, ray_idx]   camera.py:419-443, model/nerf.py:298-300
 * pose [B,3,4] world->camera, intr [B,3,3].  Pixel p -> ((p%W)+.5, (p/W)+.5, 1).  If ray_idx is
 * NULL the P pixels idx_start .. idx_start+P-1 are used (render_by_slices, model/nerf.py:326-327).
 * Backward produces d_pose [B,3,4] (overwritten) from d_center / d_ray (either may be NULL). */
int niw_raygen_pose_fwd(const float* pose, const float* intr, const int64_t* ray_idx, int64_t idx_start,
                        int B, int P, int H, int W, float* center, float* ray, void* stream);
int niw_raygen_pose_bwd(const float* pose, const float* intr, const int64_t* ray_idx, int64_t idx_start,
                        int B, int P, int H, int W, const float* d_center, const float* d_ray,
                        float* d_pose, void* stream);

/* ---- (a2) camera.get_unwarped_center_and_ray   camera.py:359-390
 * Writes pts [B, P + n_center, 3] = [P grid rows ; n_center centre rows] in the camera frame, or in the world frame
 * of pose_init [B,3,4] when it is non-NULL.  n_center = P is the layout barf_inn_llff.py:348 concatenates; n_center = 1
 * writes the (identical) centre row once. */
int niw_raygen_unwarped(const float* intr, const float* pose_init, const int64_t* ray_idx, int64_t idx_start,
                        int B, int P, int n_center, int H, int W, float* pts, void* stream);

/* ---- rays from the warped point list   model/barf_inn_llff.py:352-356, model/pose_models/inn.py:75-77
 * warped [B, P + n_center, 3] -> ray = grid rows - centre rows, center (expanded), both [B,P,3].  Backward:
 * d_warped = [d_ray ; d_center - d_ray] (n_center = P) or [d_ray ; sum over the image's rays] (n_center = 1); d_ray or
 * d_center may be NULL. */
int niw_rays_from_warp_fwd(const float* warped, int B, int P, int n_center, float* ray, float* center, void* stream);
int niw_rays_from_warp_bwd(const float* d_ray, const float* d_center, int B, int P, int n_center, float* d_warped,
                           void* stream);

/* ---- (a3) DeformNetwork.forward   model/nvp/nvp_ndr.py:365-468, model/nvp/embedder.py:41-50
 * 3 coupling blocks, hidden 128, 6 frequency bands, Softplus(beta=100), including the reference's
 * point-index annealing quirk.  `wpack` holds the EFFECTIVE (weight-norm resolved) weights,
 * NIW_NVP_BLOCK_FLOATS floats per block in the order
 *     W1a[128][27] (26 used, row stride 27), W2a[128], b2a[1], W1b[128][13], W2b[3][128], b2b[3], pad to a multiple of 4
 * (W1* = the columns of lin{b}_{a,b}_0 that multiply the embedded coordinates) -- the shared-memory image the warp
 * kernels use (odd row strides: bank-conflict free), so staging a block is one straight 16-byte-vector copy.  `code_bias`
 * [3][2][B][128] holds, per block / part / image, lin_0.bias + W_0[:, emb:] . code_b where
 * code_b = lin{b}_c(code)+code -- a B-sized product done by the host in PyTorch.
 * pts/out are [B,Pt,3].  Backward overwrites d_wpack (same layout) and d_code_bias. */
#define NIW_NVP_HIDDEN 128
#define NIW_NVP_FREQS 6
#define NIW_NVP_BLOCKS 3
#define NIW_NVP_BLOCK_FLOATS 5636   /* ((128*27 + 128 + 1 + 128*13 + 3*128 + 3) + 3) / 4 * 4 */
/* Parameter packing of the NVP network (the B-sized part of DeformNetwork.forward): resolves the
 * weight-norm re-parametrisation w = g v/||v||_row (torch.nn.utils.weight_norm, nvp_ndr.py:291-292,335-336),
 * the code projector code_b = lin{b}_c(code) + code (nvp_ndr.py:382) and folds the latent columns of the
 * first layers into per-image biases.  `params` / `grads` are HOST arrays of NIW_NVP_PARAM_PTRS device
 * pointers, 12 per block b in the order
 *     lin{b}_a_0.weight_v [128,154], .weight_g [128], .bias [128], lin{b}_a_1.weight [1,128], .bias [1],
 *     lin{b}_b_0.weight_v [128,141], .weight_g [128], .bias [128], lin{b}_b_1.weight [3,128], .bias [3],
 *     lin{b}_c.weight [128,128], .bias [128]
 * code [B,128] -> wpack [3*NIW_NVP_BLOCK_FLOATS], code_bias [3,2,B,128], cb [3,B,128] (code_b, kept for backward).
 * Backward ADDS into the `grads` buffers (the parameters' .grad tensors) and overwrites d_code [B,128]. */
#define NIW_NVP_PARAM_PTRS 36
int niw_nvp_pack_fwd(const float* const* params, const float* code, int B, float* wpack, float* code_bias,
                     float* cb, void* stream);
int niw_nvp_pack_bwd(const float* const* params, float* const* grads, const float* code, const float* cb,
                     const float* d_wpack, const float* d_code_bias, int B, float* d_code, void* stream);
/* idx_offset / idx_split / idx_jump: position of local point n of an image in the point list the reference would have
 * built, n < idx_split ? n + idx_offset : n + idx_offset + idx_jump (the embedder's annealing quirk, embedder.py:46-49,
 * is keyed on it).  Identity: (0, Pt, 0).  A ray shard passes its first ray's index in the global per-image list. */
/* (max_ctas of the backward: 0 = one CTA per SM; > 0 caps the grid, each warp then walks more points -- for callers that
 * run it next to another kernel, e.g. niw_nerf_bwd_dw on a second stream; wpack must be 16-byte aligned: the weight images
 * move as bulk copies / 128-bit vectors, NIW_E_BADARG otherwise) */
int niw_nvp_warp_fwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                     int B, int Pt, int idx_offset, int idx_split, int idx_jump, float* out, void* stream);
int niw_nvp_warp_bwd(const float* wpack, const float* code_bias, const float* pts, float alpha_ratio,
                     int B, int Pt, int idx_offset, int idx_split, int idx_jump, const float* d_out, float* d_wpack,
                     float* d_code_bias, int max_ctas, void* stream);
/* Warped ray generation in ONE launch (model/barf_inn_llff.py:325-364 + camera.py:359-390): pixel index -> un-warped grid
 * point (K^-1, pose_init [B,3,4] or NULL) -> coupling blocks -> ray = warped grid - warped camera centre, the centre warped
 * once per image (valid when no centre row of the image's point list is an annealed row: P_global >= 26).  Outputs: pts and
 * warped [B, P+1, 3] = [grid rows ; the centre row] (pts is what niw_nvp_warp_bwd takes; warped may be NULL), ray and
 * center [B, P, 3].  Backward: niw_rays_from_warp_bwd (n_center = 1) -> niw_nvp_warp_bwd -> niw_nvp_pack_bwd. */
int niw_nvp_rays_fwd(const float* wpack, const float* code_bias, const float* intr, const float* pose_init,
                     const int64_t* ray_idx, int64_t idx_start, float alpha_ratio, int B, int P, int H, int W,
                     int idx_offset, int idx_split, int idx_jump, float* pts, float* warped, float* ray, float* center,
                     void* stream);

/* ---- random pixel subset   model/nerf.py:268 (`torch.randperm(H*W)[:rand_rays//B]`)
 * out[i] = pi(i), i < k, for a keyed random bijection pi of [0,n): the first k entries of a random permutation
 * without sorting n keys.  `counter` (device, caller-zeroed) is advanced by the kernel, so replays of a captured
 * graph draw fresh subsets.  Own random stream (not torch's generator): parity tests feed the reference's draws. */
int niw_sample_pixels(int64_t n, int k, uint64_t seed, uint64_t* counter, int64_t* out, void* stream);

/* ---- (a4) Graph.sample_depth   model/nerf.py:334-344
 * depth[r,k] = ((u+k)/N)*scale + dmin, optionally 1/(.+1e-8); evaluated with the reference's
 * rounding sequence (no FMA contraction).  u [n_rays*N] or NULL (=0.5, un-stratified). */
int niw_sample_stratified(const float* u, int64_t n_rays, int N, float scale, float dmin, int inverse,
                          float* depth, void* stream);
/* same with the depth range [min, max] read from DEVICE memory (the DTU graphs carry it as a tensor, data/dtu.py:110-111):
 * no host read of the range, so the step can be captured in a CUDA graph */
int niw_sample_stratified_dev(const float* u, int64_t n_rays, int N, const float* range_dev, int inverse, float* depth,
                              void* stream);
/* The same depths with the uniforms drawn INSIDE the kernel (Philox4x32-10, 24-bit uniforms in [0, 1) as torch.rand; counter =
 * (sample group, call number), key = seed): what `torch.rand` + niw_sample_stratified give in distribution, without the torch
 * RNG op in a captured step (its graph-safe generator state costs two fill launches in front of every replay).  `rng`: two
 * zero-initialised 64-bit device words owned by the caller; the kernel advances the call number, so every launch / replay
 * draws afresh.  range_dev (device [min, max]) or NULL -> scale = max - min, dmin. */
int niw_sample_stratified_rng(int64_t n_rays, int N, float scale, float dmin, const float* range_dev, int inverse,
                              unsigned long long seed, unsigned long long* rng, float* depth, void* stream);

/* ---- (a5) Graph.sample_depth_from_pdf + cat + sort   model/nerf.py:346-365, :313-315
 * pdf [R,N]; unif [Nf] = 0.5*(grid[:-1]+grid[1:]); bins [N+1] = linspace(dmin,dmax,N+1) (both made
 * by the host with torch so they carry the reference's fp32 rounding).  The CDF is accumulated
 * sequentially in fp64 and rounded per element (what torch.cumsum does on CPU), searchsorted is
 * right=True.  Outputs (each may be NULL): fine [R,Nf], idx [R,Nf] int64, merged [R,N+Nf] =
 * ascending sort of depth_coarse [R,N] ++ fine. */
int niw_sample_pdf_merge(const float* pdf, const float* depth_coarse, const float* unif, const float* bins,
                         int64_t R, int N, int Nf, float* fine, int64_t* idx, float* merged, void* stream);

/* ---- (a9) NeRF.composite   model/nerf.py:458-474
 * ray [R,3], rgb_s [R,N,3], sigma [R,N], depth_s [R,N] -> rgb [R,3], depth [R], opacity [R],
 * prob [R,N], trans [R,N] (transmittance, saved for backward).  bg < 0 disables setbg_opaque.
 * Backward: d_rgb_s [R,N,3], d_sigma [R,N], d_ray [R,3] (through ||ray||), all overwritten. */
int niw_composite_fwd(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                      int64_t R, int N, float bg, float* rgb, float* depth, float* opacity,
                      float* prob, float* trans, void* stream);
int niw_composite_bwd(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                      const float* prob, const float* trans, int64_t R, int N, float bg,
                      const float* d_rgb, const float* d_depth, const float* d_opacity,
                      float* d_rgb_s, float* d_sigma, float* d_ray, void* stream);
/* (a9 + f1) the compositor with the loss head of model/nerf.py:276-288 (model/base.py:209-211 MSE_loss) in its epilogue:
 * the warp that composited ray r = b P + p gathers image[b, :, pixel] (image [B,3,H,W]; pixel = ray_idx[p], or idx_start + p
 * when ray_idx is NULL), writes d_unit [R,3] = 2 (rgb - gt) / (3 R) and the kernel leaves loss = mean((rgb - gt)^2) (summed
 * in block order by the last block to finish: deterministic).  scratch: niw_composite_mse_scratch_floats() floats, zeroed
 * once by the caller before the first call and owned by one stream of calls.  N in {64,128,192,256} (NIW_E_UNSUPP else).
 * Backward: d_rgb of niw_composite_bwd becomes d_rgb (may be NULL) + d_loss[0] * d_unit, d_loss a device scalar. */
int niw_composite_mse_scratch_floats(void);
int niw_composite_fwd_mse(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                          int64_t R, int N, float bg, float* rgb, float* depth, float* opacity, float* prob, float* trans,
                          const float* image, const int64_t* ray_idx, int64_t idx_start, int B, int P, int H, int W,
                          float* d_unit, float* scratch, float* loss, void* stream);
int niw_composite_bwd_mse(const float* ray, const float* rgb_s, const float* sigma, const float* depth_s,
                          const float* prob, const float* trans, int64_t R, int N, float bg,
                          const float* d_rgb, const float* d_depth, const float* d_opacity, const float* d_unit,
                          const float* d_loss, float* d_rgb_s, float* d_sigma, float* d_ray, void* stream);

/* ---- (a6+a7+a8) NeRF.forward_samples   model/nerf.py:416-456, model/barf.py:256-268, camera.py:517-521
 * x = center + depth*ray, view = normalize(ray), BARF-weighted positional encoding (L=10 / 4),
 * 8x256 MLP with skip at layer 4, softplus density from pre-ReLU column 0 of layer 7, ReLU on the
 * other 256, RGB head (256+27)->128->3, sigmoid.
 * `params` is the flat fp32 parameter vector in reference state_dict order
 *     mlp_feat.0.weight [256,63], .bias, ..., mlp_feat.7.weight [257,256], .bias,
 *     mlp_rgb.0.weight [128,283], .bias, mlp_rgb.1.weight [3,128], .bias      (NIW_NERF_PARAMS floats; the BARF `progress` scalar is not part of it)
 * `progress` is the DEVICE scalar of the BARF coarse-to-fine schedule (model/barf.py:254) and
 * [c2f_start, c2f_end] is opt.barf_c2f: band k of an L-band encoding is weighted by
 * (1 - cos(pi clamp((progress-start)/(end-start) L - k, 0, 1)))/2, evaluated on the device (no host read).
 * progress == NULL: no annealing (plain NeRF, or barf_c2f unset).
 * center/ray [R,3], depth [R,N] -> rgb [R,N,3], sigma [R,N].  `workspace` keeps what backward
 * needs (niw_nerf_workspace_bytes; `training`=0 allows a smaller, forward-only workspace).
 * Backward ADDS into d_params (caller zeroes it once per step) and overwrites d_center, d_ray. */
#define NIW_NERF_PARAMS 530052
size_t niw_nerf_workspace_bytes(int64_t R, int N, int precision, int training);
int niw_nerf_fwd(const float* params, const float* center, const float* ray, const float* depth,
                 int64_t R, int N, const float* progress, float c2f_start, float c2f_end, int precision, int training,
                 void* workspace, size_t workspace_bytes, float* rgb, float* sigma, void* stream);
/* The parameter-only part of niw_nerf_fwd for NIW_PREC_BF16 / NIW_PREC_BF16X3 (fp32 parameters -> BF16 weight streams and constants in
 * `workspace`): it does not depend on the rays, so a caller may run it early on another stream; pass
 * `training | NIW_NERF_PREPACKED` to the following niw_nerf_fwd on the same workspace to skip it there. */
#define NIW_NERF_PREPACKED 2
int niw_nerf_pack(const float* params, const float* progress, float c2f_start, float c2f_end, int precision,
                  int training, int64_t R, int N, void* workspace, size_t workspace_bytes, void* stream);
int niw_nerf_bwd(const float* params, const float* center, const float* ray, const float* depth,
                 int64_t R, int N, int precision,
                 void* workspace, size_t workspace_bytes, const float* d_rgb, const float* d_sigma,
                 float* d_params, float* d_center, float* d_ray, void* stream);
/* niw_nerf_bwd as two calls (NIW_PREC_BF16 / NIW_PREC_BF16X3): _dx runs the activation-gradient chain (d_center, d_ray, the
 * gradient images of the tile records), _dw the weight-gradient GEMMs over those records (d_params += ...).  _dw is
 * HBM-bound and independent of everything downstream of d_center / d_ray, so a training step may enqueue it on a second
 * stream -- ordered after _dx by the caller -- while the pose / warp backward runs on the first.  max_ctas > 0 caps its
 * grid (one CTA per SM) so that the concurrent kernels find free SMs; 0 = all SMs. */
int niw_nerf_bwd_dx(const float* params, const float* center, const float* ray, const float* depth, int64_t R, int N,
                    int precision, void* workspace, size_t workspace_bytes, const float* d_rgb, const float* d_sigma,
                    float* d_params, float* d_center, float* d_ray, void* stream);
int niw_nerf_bwd_dw(int64_t R, int N, int precision, void* workspace, size_t workspace_bytes, float* d_params,
                    int max_ctas, void* stream);


/* ---- loss head: image gather + MSE   model/nerf.py:276-288, model/base.py:209-211
 * image [B,3,H,W], rgb [B,P,3]; loss += scale * sum((rgb - image[:, :, pix])^2) (scale = 1/(B*P*3)
 * for the mean); d_rgb = 2*scale*(rgb - target) is written when non-NULL.  loss is a device scalar
 * that the caller zeroes. */
int niw_mse_gather(const float* image, const float* rgb, const int64_t* ray_idx, int64_t idx_start,
                   int B, int P, int H, int W, float scale, float* loss, float* d_rgb, void* stream);
/* 1: `loss` accumulates over several blocks and must be zeroed by the caller; 0 (B*P <= 1024): one block writes it. */
int niw_mse_gather_needs_zero(int B, int P);

/* ---- evaluation metrics (row f3)   model/nerf.py:176-183, external/pohsun_ssim/pytorch_ssim/__init__.py:7-37
 * pred_rgb [B,H*W,3] (the renderer's layout), image [B,3,H,W] -> out [B,2] = per image (sum of squared errors,
 * sum of the SSIM map over the 3*H*W entries; 11x11 Gaussian window sigma 1.5, zero padding).  out is zeroed here.
 * PSNR = -10 log10(out[b][0] / (3 H W)), SSIM = out[b][1] / (3 H W). */
int niw_image_metrics(const float* pred_rgb, const float* image, int B, int H, int W, float* out, void* stream);

/* ---- depth error (row f3)   core/metrics.py:64-111
 * pred, gt [n], valid [n] bytes or NULL -> out[5] = (#valid, sum|gt-p|, sum(gt-p)^2, sum|gt-s p|, sum(gt-s p)^2), zeroed here. */
int niw_depth_metrics(const float* pred, const float* gt, const uint8_t* valid, int64_t n, float scale, float* out,
                      void* stream);

/* ---- rigid registration of the global-alignment loss (row f1)   model/nerf_inn_llff.py:563-572, model/pose_models/inn.py:96-102
 * (roma.rigid_points_registration, roma==1.4.1).  x, y [B,M,3] -> the least-squares proper rotation R [B,3,3] and
 * translation t [B,3] with y ~ R x + t (Kabsch optimum via Horn's quaternion form, fp64 Jacobi; no host sync). */
int niw_kabsch(const float* x, const float* y, int B, int M, float* R, float* t, void* stream);
/* The same fit in two steps for point lists sharded over data-parallel ranks (the reference fits each image's WHOLE list,
 * model/nerf_inn_llff.py:566-572): _stats writes stats [B,16] doubles = (rows, sum x [3], sum y [3], sum x_a y_c [9]) of this
 * rank's rows; the caller sums the stats over the ranks (one all-reduce of 16 doubles per image); _solve fits from the sums.
 * With one rank, _stats followed by _solve equals niw_kabsch up to fp64 rounding of the un-centred sums. */
int niw_kabsch_stats(const float* x, const float* y, int B, int M, double* stats, void* stream);
int niw_kabsch_solve(const double* stats, int B, float* R, float* t, void* stream);

/* ---- optimiser step (row f2): torch.optim.Adam + ExponentialLR over ONE flat fp32 segment
 *      model/nerf.py:33-46,87,92 (optim / sched), model/barf_inn_llff.py:84-120 (optim_pose)
 * params/grads/exp_avg/exp_avg_sq [n] (16-byte aligned).  `state` is a DEVICE array of 2 floats owned by the
 * caller and zero-initialised: state[0] = number of steps taken so far (advanced by the kernel, so the call can be
 * replayed from a CUDA graph), state[1] = scratch.  Update (torch's Adam, amsgrad off, maximize off):
 *   t = state[0]+1; lr_t = lr * lr_gamma^(t-1) (evaluated in double, as torch's ExponentialLR accumulates it);
 *   g += weight_decay*p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *   p -= lr_t/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * warmup_iters > 0: lr_t *= min(1, (t-1)/warmup_iters) for the first warmup_n elements of the segment (the pose-LR warm-up
 * of model/barf.py:48-51; the reference applies it to optim_pose.param_groups[0] only, i.e. to warp_mlp and not to the
 * latent codes that follow it in the segment: model/barf_inn_llff.py:108-111).  progress0/1 (device scalars, may be NULL):
 * set to t / max_iter after the update -- `nerf.progress.data.fill_(it/max_iter)`, model/barf.py:57-59. */
int niw_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                  double lr_gamma, float beta1, float beta2, float eps, float weight_decay, float warmup_iters,
                  int64_t warmup_n, float max_iter, float* progress0, float* progress1, float* state, void* stream);

/* ---- tcgen05 self-test: D[128,N] = A[128,K] . B[N,K]^T with BF16 operands staged exactly as the
 * MLP kernel stages them.  variant selects the operand form under test (see csrc/tc_selftest.cu); variant 4 is the
 * CTA-pair form (tcgen05 cta_group::2): A [256,K], D [256,N]. */
int niw_tc_selftest(const float* A, const float* Bm, int N, int K, int variant, float* D, void* stream);

/* ---- probe (diagnostics, scripts/probe_tmem.py): cycles of a tcgen05.mma stream (what & 1: iters x 16 MMAs of
 * M = 128, N = 256, K = 16 into TMEM columns 0-255) and of four warps draining TMEM columns 256-511 (what & 2: iters x
 * 128 KB), alone or concurrently, on one SM.  out[0] = MMA cycles, out[1] = slowest reader's cycles (device int64[2]). */
int niw_tc_probe(int what, int iters, long long* out, void* stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (no counterpart in the reference, which
 * asserts a single GPU at options.py:103) ------------------------------------------------------------------------------
 * The only exchange step of a training step is the sum of the flat gradient bucket (0.53 M + 0.17 M floats): at that size
 * a library all-reduce is launch + protocol latency, so the sum is two kernels of our own over CUDA-IPC-mapped buffers
 * (csrc/p2p.cu): publish (copy into this rank's exchange buffer, raise its flag in every peer) and reduce (wait for all
 * flags, sum the `world` buffers in rank order -- bit-identical on all ranks -- straight from peer memory).
 * niw_p2p_alloc: a device block of `bytes` (zeroed) and its 64-byte IPC handle; niw_p2p_open maps a peer's handle into
 * this process.  Block layout: flag block, at byte 256 the exchange buffer of 2 * half_floats floats (double buffered on
 * the sequence number kept in the block, so the call is CUDA-graph capturable), then 2 * half_floats floats of result area:
 * bytes >= 256 + 16 * half_floats.  From 4 ranks up the sum is two-shot (reduce-scatter into the result areas, all-gather).
 * niw_allreduce_p2p: in-place sum of n floats (n % 4 == 0, n <= half_floats) over `world` <= 8 ranks of one node;
 * blocks[r] = rank r's block as mapped here.  All ranks make the same sequence of calls.  niw_p2p_error reads the
 * time-out flag of a block (a wait for a peer that never arrived gives up after ~2 s instead of hanging the GPU). */
int niw_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64);
int niw_p2p_open(const void* handle64, void** dev_ptr);
int niw_p2p_close(void* dev_ptr);
int niw_p2p_free(void* dev_ptr);
int niw_allreduce_p2p(float* data, int64_t n, void* const* blocks, int rank, int world, int64_t half_floats, void* stream);
int niw_p2p_error(const void* block, unsigned int* error);

#ifdef __cplusplus
}
#endif
#endif /* NIW_B200_H_ */
