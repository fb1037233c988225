/* pea_b200.h -- C ABI of the B200-native Pixel-Embedded-Affinity hot path.
 *
 * The reference (weih527/Pixel-Embedded-Affinity) has no FFI: its boundary for this path is a set
 * of Python callables.  Each entry point below is the device-side replacement of one (group of)
 * reference callable(s), cited as file:line under the reference tree.  The Python drop-in layer
 * (the .py files in pixel-embedded-affinity_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only.  No torch / C++ types cross this boundary.
 *   - Every data pointer is a DEVICE pointer owned by the caller; `offsets`, `loss_norm`,
 *     `loss_coef` are small HOST arrays that are copied into kernel parameters during the call.
 *   - The library allocates nothing, keeps no global mutable state, and enqueues all work on the
 *     `stream` argument (a cudaStream_t passed as void*).  No call synchronises the device.
 *     Calls are re-entrant and CUDA-graph capturable.
 *   - Return value: 0 on success, a negative PEA_E_* code for a rejected argument, or a positive
 *     cudaError_t if a launch failed.  Nothing throws, nothing exits.
 *   - Volumes are channel-planar, contiguous: embedding (B, C, D, H, W) float32, 2-D callers pass
 *     D = 1.  Per-offset tensors (target / weight / mask / affs) are (B, K, D, H, W).
 *   - Built for sm_100a only.  There is no CPU path.
 */
#ifndef PEA_B200_H_
#define PEA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PEA_ABI_VERSION 7
#define PEA_MAX_OFFSETS 64

/* error codes (negative); positive return values are cudaError_t */
#define PEA_OK 0
#define PEA_E_NULL (-1)         /* a required pointer is NULL */
#define PEA_E_SHAPE (-2)        /* non-positive or overflowing extent */
#define PEA_E_CHANNELS (-3)     /* embedding channel count not one of 4, 8, 16, 32 */
#define PEA_E_OFFSETS (-4)      /* K outside [1, PEA_MAX_OFFSETS] */
#define PEA_E_WORKSPACE (-5)    /* workspace smaller than pea_workspace_bytes() */
#define PEA_E_DTYPE (-6)        /* unknown mask / label dtype code */
#define PEA_E_ALIGN (-7)        /* pointer not aligned for its element type */
#define PEA_E_MODE (-8)         /* unknown border / variant code */

/* border rule for the neighbour p + offset */
#define PEA_BORDER_WRAP 0 /* circular, torch.roll semantics: scripts_cvppp/loss/loss_embedding_mse.py:8 */
#define PEA_BORDER_ZERO 1 /* valid region only, affinity 0 elsewhere: scripts_ac3ac4/loss/loss_embedding_mse.py:147-151,186-190 */

/* dtype codes */
#define PEA_MASK_NONE 0
#define PEA_MASK_U8 1
#define PEA_MASK_F32 2
#define PEA_LABEL_U8 0
#define PEA_LABEL_U16 1
#define PEA_LABEL_U32 2
#define PEA_LABEL_U64 3
#define PEA_LABEL_F32 4
#define PEA_LABEL_S8 5
#define PEA_LABEL_S16 6
#define PEA_LABEL_S32 7
#define PEA_LABEL_S64 8

/* kernel-selection hint for pea_affs_loss_fwd_bwd / pea_affs_fwd (results are identical) */
#define PEA_IMPL_AUTO 0
#define PEA_IMPL_GENERIC 1 /* table-driven gather kernel, any offsets */
#define PEA_IMPL_TILED 2   /* shared-memory / register-tiled kernels for axis-aligned offset sets */
/* Epilogue flag, or-ed into `impl` of pea_affs_fwd: store max(a, 0) -- the caller-side F.relu(pred) of
 * scripts_ac3ac4/inference.py:165, scripts_cvppp/main.py:312 folded into the kernel's store. */
#define PEA_EPI_RELU 0x100
/* Flag or-ed into `impl` of pea_affs_loss_fwd_bwd: grad_emb += d loss / d emb instead of grad_emb = ... -- the two losses of
 * a training step that share an embedding (scripts_ac3ac4/main.py:223-224: embedding_loss_norm5, then
 * ema_embedding_loss_norm5 on the same `embedding`) write ONE gradient buffer, instead of two buffers that autograd adds
 * in a third pass.  Taken by the tiled two-view 3-D kernel only (12-offset set, partner != NULL, C = 16); PEA_E_MODE else. */
#define PEA_EPI_GRAD_ACCUMULATE 0x200

int pea_abi_version(void);
const char* pea_strerror(int code);

/* TEST-ONLY: dispatch knobs of the parity tests (process-wide, 0 = production behaviour; the only mutable state in the
 * library, written by nothing but this call).  PEA_KNOB_NO_CANON != 0: the canonical 2-D tables take the table-driven
 * axis kernel; PEA_KNOB_MAX_CTAS = n > 0: persistent 2-D kernels launch at most n CTAs (every CTA walks many tiles);
 * PEA_KNOB_TUNE_TY2D = 8 / 12 / 16: pins the tile height of the generic 2-D tiled kernel; PEA_KNOB_GT_TWO_PASS != 0:
 * pea_gt_targets takes its two-launch route even where the one-launch (cooperative) route is available. */
#define PEA_KNOB_NO_CANON 0
#define PEA_KNOB_MAX_CTAS 1
#define PEA_KNOB_TUNE_TY2D 2
#define PEA_KNOB_GT_TWO_PASS 3
#define PEA_KNOB_K3_ROT_WAVES 4   /* z-shard forward kernel: 0 = default policy; n > 0: only CTAs of the first n waves visit
                                   * the halo-dependent slices last; -1: none do (timing experiments, same results) */
#define PEA_KNOB_COUNT 5
int pea_debug_configure(int knob, int value);

/* Bytes of device workspace the affinity entry points need for a problem of this size.
 * `cross` != 0 when a partner (second-view) embedding is passed. */
size_t pea_workspace_bytes(int B, int C, int D, int H, int W, int K, int cross);

/* embedding -> affinities, forward only.
 *   replaces  embedding2affs                 scripts_cvppp/loss/loss_embedding_mse.py:58-66   (border = WRAP)
 *             inf_embedding_loss_norm1/5     scripts_ac3ac4/loss/loss_embedding_mse.py:54-67,212-234 (border = ZERO)
 * affs[b,k,p] = <e^[b,:,p], u^[b,:,p+o_k]>,  e^ = emb / max(|emb|_2, 1e-12), u^ likewise from `partner`
 * (or e^ itself when partner == NULL).  offsets: host int32[K*3] as (dz, dy, dx).
 * z-sharded volumes: the first `z_skip` slices of emb are halo received from the previous shard --
 * they are read as neighbours but produce no output, so affs has D - z_skip slices.  `emb_plane` / `affs_plane`
 * are the element strides between consecutive channel planes (0 = contiguous: D*H*W and (D - z_skip)*H*W), which
 * lets emb / affs be z-range views of deeper buffers without a copy.
 * Alternatively (z_skip == 0) `halo` points at the previous shard's last `halo_slices` (<= 4) slices wherever they
 * live -- typically the PEER GPU's memory mapped over NVLink (symmetric memory) -- laid out (B, C, halo_slices, H, W)
 * with channel-plane stride `halo_plane`: the kernel reads them in place, so no exchange step or staging copy
 * exists at all.  halo == NULL: no external halo.  Only the tiled 3-D forward kernel takes an external halo. */
int pea_affs_fwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                 const int32_t* offsets, int K, int border, int impl,
                 int z_skip, int64_t emb_plane, int64_t affs_plane,
                 const float* halo, int halo_slices, int64_t halo_plane,
                 float* affs, void* workspace, size_t ws_bytes, void* stream);

/* z-sharded volumes, STAGED halo (ABI 7; SURVEY.md 8e, BASELINE config 4 -- the reference has no counterpart: it
 * blends overlapping patches on the host, scripts_ac3ac4/data/provider_valid.py:306-348; the per-voxel arithmetic is
 * inf_embedding_loss_norm5, scripts_ac3ac4/loss/loss_embedding_mse.py:212-234).
 * pea_affs_fwd with `halo` in peer memory pulls 4 slices over NVLink from inside the compute kernel.  The staged
 * flavour lets a COPY ENGINE pull them instead, while the kernel runs:
 *   pea_zshard_stage_halo   enqueues, on `stream` (a side stream), one 3-D copy per band of `band_rows` image rows of
 *                           tail (B, C, halo_slices, H, W; channel-plane stride tail_plane, a multiple of H*W) into
 *                           `stage` (contiguous), rows [row0, H) only, each followed by a stream-ordered 32-bit write
 *                           of `epoch` to flags[j] (j = band index; device memory, ceil((H - row0) / band_rows) words);
 *   pea_zshard_affs_fwd     the 12-offset AC3/AC4 forward kernel on `stream` (the main stream): tiles whose first row
 *                           is below row0 read `halo` (peer memory) directly, as pea_affs_fwd does -- their CTAs start
 *                           before a band can have arrived; every other tile waits until its band's flag has reached
 *                           `epoch` (signed 32-bit distance, so epochs may wrap) and reads `stage`.
 * row0 and band_rows must be multiples of 8 (the kernel's tile height).  The caller advances `epoch` every step and
 * orders the side stream behind the producer's writes of the tail (and the next step's copies behind this kernel).
 * Results are bit-identical to pea_affs_fwd with the same halo. */
int pea_zshard_stage_halo(const float* tail, int64_t tail_plane, float* stage, int B, int C, int halo_slices, int H, int W,
                          int row0, int band_rows, uint32_t* flags, uint32_t epoch, void* stream);
int pea_zshard_affs_fwd(const float* emb, int B, int C, int D, int H, int W, int64_t emb_plane, int64_t affs_plane,
                        const float* halo, int halo_slices, int64_t halo_plane,
                        const float* stage, const uint32_t* flags, uint32_t epoch, int row0, int band_rows,
                        int relu, float* affs, void* stream);

/* Fused embedding -> affinities -> weighted-MSE loss -> d loss / d embedding, one pass.
 *   replaces  embedding_loss / ema_embedding_loss      scripts_cvppp/loss/loss_embedding_mse.py:18-47,79-95
 *             embedding_loss_norm1/5, ema_* variants   scripts_ac3ac4/loss/loss_embedding_mse.py:7-51,169-194,263-289
 *             WeightedMSE                              scripts_cvppp/loss/loss.py:106-124
 *             and the autograd replay of all of the above.
 * Per offset k:  S_k = sum_p w (a m - t m)^2  over the pairs the border rule keeps;
 *                loss_k = S_k / loss_norm[k];   loss = sum_k loss_coef[k] * loss_k  (float32, in offset order).
 * loss_out (device, 1 + K floats) receives [loss, loss_0 .. loss_{K-1}].
 * grad_emb (B,C,D,H,W) receives d loss / d emb for an upstream gradient of 1; grad_partner likewise for
 * the partner embedding; either may be NULL (not computed).  affs may be NULL (not stored).
 * mask may be NULL with mask_dtype = PEA_MASK_NONE. */
int pea_affs_loss_fwd_bwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                          const int32_t* offsets, int K, int border, int impl,
                          const float* target, const float* weight, const void* mask, int mask_dtype,
                          const float* loss_norm, const float* loss_coef,
                          float* affs, float* loss_out, float* grad_emb, float* grad_partner,
                          void* workspace, size_t ws_bytes, void* stream);

/* Backward of pea_affs_fwd for an arbitrary upstream gradient (criterion other than WeightedMSE, or a
 * caller that back-propagates through the returned 2-D affs, scripts_cvppp/loss/loss_embedding_mse.py:46).
 * grad_affs is (B,K,D,H,W); grad_emb / grad_partner are overwritten (either may be NULL). */
int pea_affs_bwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                 const int32_t* offsets, int K, int border,
                 const float* grad_affs, float* grad_emb, float* grad_partner,
                 void* workspace, size_t ws_bytes, void* stream);

/* Several SMALL fused losses in ONE launch: the deep-supervision calls of a training step,
 *   scripts_ac3ac4/main.py:227-230   embedding_loss_norm1(emd_i, down_i[:, :3], down_i[:, 3:], criterion)   x 4 scales
 *   scripts_cvppp/main.py:284-287    embedding_loss(emd_i, ..., offsets[:2 i])                              x 4 scales
 * Every problem has the semantics of pea_affs_loss_fwd_bwd (same fields, same loss_out layout); `partner` != NULL makes
 * it the two-view loss with the second view constant (no grad_partner here).  Up to PEA_MULTI_MAX_PROBLEMS problems of
 * up to PEA_MULTI_MAX_OFFSETS offsets (|component| <= 127) and the same channel count C.  The kernel finalizes every
 * problem's loss itself (last CTA to arrive, fixed order: reproducible): `ticket` is REQUIRED -- one device uint32 that is
 * ZERO on entry and is left zero; one call in flight per ticket word (the Python layer keeps one per stream).  workspace: at least
 * pea_affs_loss_multi_workspace(problems, n) bytes. */
#define PEA_MULTI_MAX_PROBLEMS 8
#define PEA_MULTI_MAX_OFFSETS 12
typedef struct pea_loss_problem {
    const float* emb;        /* (B, C, D, H, W) */
    const float* partner;    /* second view or NULL */
    const float* target;     /* (B, K, D, H, W) */
    const float* weight;
    const void* mask;        /* NULL with mask_dtype = PEA_MASK_NONE */
    float* affs;             /* (B, K, D, H, W) or NULL */
    float* grad_emb;         /* (B, C, D, H, W) or NULL */
    float* loss_out;         /* 1 + K floats */
    const int32_t* offsets;  /* host, K * 3 as (dz, dy, dx) */
    const float* loss_norm;  /* host, K */
    const float* loss_coef;  /* host, K */
    int32_t B, D, H, W, K, border, mask_dtype;
} pea_loss_problem;
size_t pea_affs_loss_multi_workspace(const pea_loss_problem* problems, int n);
int pea_affs_loss_multi(const pea_loss_problem* problems, int n, int C, void* workspace, size_t ws_bytes,
                        uint32_t* ticket, void* stream);
/* bufs[i][0 .. lens[i]) *= *scales[i] for i < n <= PEA_MULTI_MAX_PROBLEMS in one launch (the upstream gradients of the
 * losses of one pea_affs_loss_multi call); a buffer whose scalar is exactly 1 is not touched.  Host arrays of device
 * pointers. */
int pea_scale_multi(float* const* bufs, const int64_t* lens, const float* const* scales, int n, void* stream);

/* Full-volume inference epilogue: Gaussian-weighted accumulation of overlapping patches,
 *   replaces  Provider_valid.add_vol / get_results     scripts_ac3ac4/data/provider_valid.py:306-348
 *             `1 - output_affs` for the Mutex Watershed scripts_ac3ac4/inference.py:196
 * pea_blend_add:    acc[k, z0+z, y0+y, x0+x] += patch[k,z,y,x] * weight[z,y,x]   (product rounded, then added: NumPy's
 *                   two roundings),  wmap[z0+z, y0+y, x0+x] += weight[z,y,x].   patch (K,d,h,w), weight (d,h,w),
 *                   acc (K,D,H,W), wmap (D,H,W).  Patches are added one call at a time (stream order), like the loop.
 * pea_blend_finish: out[k] = acc[k] / wmap cropped by (pad_z, pad_y, pad_x) on both sides of every axis
 *                   (valid_padding, :330-340), or 1 - that when one_minus != 0.  out (K, D-2pz, H-2py, W-2px). */
int pea_blend_add(const float* patch, const float* weight, int K, int d, int h, int w, float* acc, float* wmap,
                  int D, int H, int W, int z0, int y0, int x0, void* stream);
int pea_blend_finish(const float* acc, const float* wmap, int K, int D, int H, int W, int pad_z, int pad_y, int pad_x,
                     int one_minus, float* out, void* stream);

/* Undo the per-sample flip / transpose augmentation of the second view before the two-view loss, one pass, rules read on
 * the device (no host synchronisation)  
 *   replaces  convert_consistency_flip   scripts_cvppp/data/data_consistency.py:34-45 (+ :19-31)
 *                                        scripts_ac3ac4/utils/consistency_aug.py:217-228 (+ :58-77)
 * in / out (B, C, D, H, W) float32, out != in; rules: DEVICE int32 [B][4] as (z reflection, x reflection, y reflection,
 * xy transpose); 2-D callers pass D = 1 and rz = 0.  H must equal W for samples whose transpose bit is set (the
 * reference stacks the samples back into one tensor, so it needs that too).
 * out[b,c,z,y,x] = in[b,c,z',y',x'], or in[b,c,z',x',y'] under the transpose, with z' = D-1-z, y' = H-1-y, x' = W-1-x
 * where the corresponding reflection bit is set. */
int pea_consistency_flip(const float* in, const int32_t* rules, int B, int C, int D, int H, int W, float* out, void* stream);

/* buf[i] *= *scale (a device scalar), in place; exits without touching memory when *scale == 1.
 * Used by the autograd backward to apply the upstream gradient of the scalar loss to the gradient the
 * fused forward already produced (the common `loss.backward()` case costs one empty launch). */
int pea_scale_inplace(float* buf, int64_t n, const float* scale, void* stream);

/* labels -> (affinity float32, validity mask uint8), 2-D, per offset (dy, dx).
 *   replaces  gen_affs_ours   scripts_cvppp/utils/affinity_ours.py:17-39
 * labels (B,H,W) of `label_dtype`; offsets host int32[K*2]; affs/mask (B,K,H,W). mask may be NULL. */
int pea_gt_affs2d(const void* labels, int label_dtype, int B, int H, int W,
                  const int32_t* offsets, int K, int ignore, int padding,
                  float* affs, uint8_t* mask, void* stream);

/* segmentation -> background-aware affinities by neighbourhood (dz, dy, dx); 2-D callers pass D = 1.
 *   replaces  seg_to_aff      scripts_cvppp/data/data_affinity.py:53-102
 * replicate_pad applies the `pad == 'replicate'` border rule (:94-100): 0 = off, 3 = three edges of a 3-D
 * volume (K must be 3), 2 = two edges of a 2-D image passed as D = 1 (K must be 2).
 * seg (B,D,H,W); affs (B,K,D,H,W).  Signed label codes make the `seg > 0` test signed, as NumPy's is. */
int pea_gt_affs3d(const void* seg, int label_dtype, int B, int D, int H, int W,
                  const int32_t* nhood, int K, int replicate_pad,
                  float* affs, void* stream);

/* Class-balance weights, one independent problem per channel.
 *   replaces  weight_binary_ratio   scripts_cvppp/data/data_segmentation.py:205-228   (mask=None)
 * label (n_channels, n_per_channel) float32 -> weight, same shape.  stats: device scratch of
 * pea_wbr_stats_bytes(n_channels) bytes. */
size_t pea_wbr_stats_bytes(int n_channels);
int pea_weight_binary_ratio(const float* label, int n_channels, int64_t n_per_channel, double alpha,
                            float* weight, void* stats, size_t stats_bytes, void* stream);

/* Whole-batch ground truth (the training-loop form of the three functions above,
 * scripts_cvppp/data/data_provider.py:204-225, scripts_ac3ac4/data/data_provider_labeled_deep.py:246-285 with
 * separate_weight): labels (B,D,H,W) -> target (B,K,D,H,W), [mask u8, semantics 0 only], weight (B,K,D,H,W), one
 * weight problem per (sample, channel).  semantics 0 = gen_affs_ours (ignore / padding apply), 1 = seg_to_aff
 * without replicate pad.  offsets: host int32[K*3] (dz,dy,dx).  counts: device scratch of pea_gt_counts_bytes(B,K)
 * (the per-(sample, channel) counts plus one completion counter per sample; ABI 6: 4 B more per sample than ABI 5).
 * One cooperative launch when the whole batch can be resident (counting phase with the results parked in shared memory,
 * then the stores); two launches (count + write, then recompute + write weights) otherwise.  Results are identical. */
size_t pea_gt_counts_bytes(int B, int K);
int pea_gt_targets(const void* labels, int label_dtype, int B, int D, int H, int W,
                   const int32_t* offsets, int K, int semantics, int ignore, int padding, double alpha,
                   float* target, uint8_t* mask, float* weight, void* counts, size_t counts_bytes, void* stream);

/* The border fix-up that follows the 3-D affinities in the reference's training / inference loops
 * (scripts_ac3ac4/main.py:233-236, inference.py:161-163), in place:
 *   affs[:, 0, :shift] = affs[:, 0, shift:2 shift];  affs[:, 1, :, :shift] = affs[:, 1, :, shift:2 shift];
 *   affs[:, 2, :, :, :shift] = affs[:, 2, :, :, shift:2 shift]
 * affs (B,K>=3,D,H,W) with `plane` elements between channel planes (0 = D*H*W); extents smaller than 2 shift on an
 * axis leave that channel untouched.  One launch over the three thin slabs. */
int pea_replicate_first(float* affs, int B, int K, int D, int H, int W, int64_t plane, int shift, void* stream);

/* Label-driven fused 2-D loss (the step before the path -- the reference's Dataset.__getitem__,
 * scripts_cvppp/data/data_provider.py:204-225 -- folded into the path).  Two calls:
 *   pea_gt_pair_weights       labels -> wtab (device, B*K*2 floats): the two values the class-balance weight map of each
 *                             (sample, channel) takes where the target is 1 / 0 (weight_binary_ratio of the
 *                             gen_affs_ours / seg_to_aff target, counted without materialising the target);
 *   pea_affs_loss_labels_fwd_bwd   embedding + int32 labels + wtab -> loss, affinities, gradient, with
 *                             target = gen_affs_ours(labels, offsets, ignore, padding)[0], mask = [1] (pairs across the
 *                             image edge dropped) and weight derived inside the kernel: results equal
 *                             pea_gt_targets followed by pea_affs_loss_fwd_bwd with a uint8 mask.
 * Implemented (C = 16): semantics 0 (gen_affs_ours), D = 1, either border, with the canonical 2-D tables
 * multi_offset(shifts <= 12, neighbor=4) (unrolled kernel) and with every other axis-aligned table whose shifts are
 * <= 28, e.g. CVPPP's multi_offset([1, 3, 5, 9, 27], 4) (table-driven kernel, W % 4 == 0); semantics 1 (seg_to_aff without replicate pad) with the 12-offset AC3/AC4 table,
 * PEA_BORDER_ZERO.  Anything else returns PEA_E_MODE and the caller takes the two-step route.
 * counts: device scratch of pea_gt_counts_bytes(B, K). */
int pea_gt_pair_weights(const void* labels, int label_dtype, int B, int D, int H, int W,
                        const int32_t* offsets, int K, int semantics, int ignore, int padding, double alpha,
                        float* wtab, void* counts, size_t counts_bytes, void* stream);
int pea_affs_loss_labels_fwd_bwd(const float* emb, const int32_t* labels, int B, int C, int D, int H, int W,
                                 const int32_t* offsets, int K, int border, int semantics, int ignore, const float* wtab,
                                 const float* loss_norm, const float* loss_coef,
                                 float* affs, float* loss_out, float* grad_emb,
                                 void* workspace, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PEA_B200_H_ */
