/* sphk.h -- C ABI of the B200 (sm_100a) spherical-box IoU kernels.
 *
 * Drop-in boundary for the IoU hot path of ManuelVeras/sph-retina.  Each entry point names
 * the reference interface (file:line, relative to the reference root) whose arithmetic it
 * replaces; the Python layer in sph_retina_b200/sphdet keeps the reference's signatures
 * and calls these through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in
 *     _host; nothing is allocated or freed by the library, the caller owns all buffers;
 *   - boxes are float32 row-major [n, D], D = 4 (theta, phi, alpha, beta) or 5 (+gamma),
 *     degrees, exactly the reference layout (sphdet/iou/sph_iou_calculator.py:22-31);
 *   - inputs are never written (tests/test_all_ious.py:322-331);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; 0 = legacy default) and
 *     the call returns without synchronising;
 *   - return value: SPHK_OK or a negative error code; sphk_last_error_string() gives the
 *     reason for the calling thread.  No C++ exception crosses the boundary.
 */
#ifndef SPHK_H_
#define SPHK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHK_ABI_VERSION 8

/* OR-ed into `kind` of the NMS entry points: suppress iff IoU > threshold (mmcv nms) instead of !(IoU <= threshold) (SphNMS) */
#define SPHK_NMS_RULE_GT 0x100

enum sphk_status {
    SPHK_OK = 0,
    SPHK_ERR_INVALID_ARGUMENT = -1,
    SPHK_ERR_CUDA = -2,
    SPHK_ERR_UNSUPPORTED = -3
};

/* IoU flavour: reference backend strings of sph_overlaps (sphdet/iou/sph_iou_calculator.py:58-113) */
enum sphk_kind {
    SPHK_KIND_SPH2POB_EFFICIENT = 0, /* 'sph2pob_efficient_iou'  sphdet/iou/sph_iou_api.py:97-98   */
    SPHK_KIND_SPH2POB_STANDARD = 1,  /* 'sph2pob_standard_iou'   sphdet/iou/sph_iou_api.py:94-95   */
    SPHK_KIND_SPH = 2,               /* 'sph_iou'                sphdet/iou/sph_iou_api.py:130-151 */
    SPHK_KIND_FOV = 3,               /* 'fov_iou'                sphdet/iou/sph_iou_api.py:156-177 */
    SPHK_KIND_NAIVE = 4,             /* 'naive_iou'              sphdet/iou/sph_iou_api.py:181-198 (planar IoU of the
                                        sph2pix boxes; BFoV or RBFoV, mode 'iou' only)                            */
    SPHK_KIND_UNBIASED = 5,          /* 'unbiased_iou'           sphdet/iou/sph_iou_api.py:103-125 (exact spherical IoU,
                                        unbiased_iou_bfov.py / unbiased_iou_rbfov.py; the default backend of
                                        SphOverlaps2D; evaluated in double; mode 'iou' only)                      */
    SPHK_KIND_SPH2POB_LEGACY = 6     /* 'sph2pob_legacy_iou'     sphdet/iou/sph_iou_api.py:91-92 (sph2pob_legacy.py:8-31,
                                        the hand-crafted first transform; BFoV only; rbb_angle ignored)           */
};
enum sphk_mode { SPHK_MODE_IOU = 0, SPHK_MODE_IOF = 1 };                     /* sph_iou_api.py:49     */
enum sphk_edge { SPHK_EDGE_ARC = 0, SPHK_EDGE_CHORD = 1, SPHK_EDGE_TANGENT = 2 }; /* sph2pob_efficient.py:100-108 */
enum sphk_angle { SPHK_ANGLE_EQUATOR = 0, SPHK_ANGLE_PROJECT = 1 };                /* sph2pob_efficient.py:81-97    */

int sphk_abi_version(void);
const char* sphk_last_error_string(void);
/* number of SMs / compute capability of the current device, for grid sizing and sanity checks */
int sphk_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Aligned IoU: out[p] = IoU(b1[p], b2[p]), p < P.
 * Replaces _sph2pob_iou_auxiliary(..., is_aligned=True) (sphdet/iou/sph_iou_api.py:48-86: both
 * jitters :222-260, transform sph2pob_efficient.py:9-73 | sph2pob_standard.py:8-80, rotated IoU
 * mmcv.ops.box_iou_rotated at :79, clamp :86) and sph_iou / fov_iou (:130-177 with
 * approximate_ious.py:3-55; D must be 4, mode must be IOU for those two kinds).
 * angle = SPHK_ANGLE_PROJECT (the reference's ablation option rbb_angle='project') runs a plain
 * double-precision kernel: correct, not tuned. */
int sphk_iou_aligned(int kind, const float* b1, const float* b2, int64_t P, int D, int mode, int edge, int angle,
                     float* out, void* stream);

/* Pairwise IoU of rows[R,D] x cols[C,D]; pair (i,j) = (rows[i] as bboxes1, cols[j] as bboxes2),
 * the reference's expansion order (sphdet/iou/sph_iou_api.py:59-61).  Any output may be NULL:
 *   out        [R, ld]  the matrix (ld >= C, in elements)                    (:85 view(rows, cols))
 *   row_max/row_arg [R] max / argmax over the columns  -- MaxIoUAssigner's overlaps.max(dim=1)
 *   col_max/col_arg [C] max / argmax over the rows     -- overlaps.max(dim=0)
 *                       (mmdet/core/bbox/assigners/max_iou_assigner.py:173-176)
 * Ties resolve to the lowest index.  row_base / col_base are added to the reported indices so a
 * shard of a larger matrix reports global indices.  `workspace` (16-byte aligned device memory of
 * sphk_iou_pairwise_workspace_bytes(R, C) bytes) holds the per-box precompute of the Sph2Pob kinds and
 * the packed max/argmax keys; it may be NULL only for the sph/fov kinds without max outputs.
 * With angle = SPHK_ANGLE_PROJECT only the matrix output is supported. */
int64_t sphk_iou_pairwise_workspace_bytes(int64_t R, int64_t C);
int sphk_iou_pairwise(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode,
                      int edge, int angle, float* out, int64_t ld, float* row_max, int32_t* row_arg, float* col_max,
                      int32_t* col_arg, int32_t row_base, int32_t col_base, void* workspace, void* stream);

/* The fused max/argmax of sphk_iou_pairwise as PACKED keys (no matrix, no unpacking), for reductions across
 * shards: key = float32 bits << 32 | (0xFFFFFFFF - index), 0 = "no positive overlap".  The integer maximum of
 * such keys over shards is (max value, lowest index) -- the single-device tie rule.
 *   row_keys [R]: max over the columns (index = col_base + j);  col_keys [C]: max over the rows (row_base + i)
 *   keep: bit 0 / bit 1 = do NOT reset the row / column keys first but raise the ones already there (a sweep processed in
 *         row chunks accumulates the per-column maxima of all chunks in one key array: keep = 2 from the second chunk on) */
int sphk_iou_pairwise_keys(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                           uint64_t* row_keys, uint64_t* col_keys, int32_t row_base, int32_t col_base, int keep, void* workspace,
                           void* stream);

/* Row-sharded N x M across GPUs (BASELINE configs[4]; sph_retina_b200/sharded.py): what follows the all-gather of the
 * shards' packed keys.  The long operand (anchors) is split contiguously and balanced over `world` shards (the first
 * n_long % world shards own one row more), the short operand (ground truths) is replicated.  Each shard contributes a
 * block of cap + n_short keys (cap >= ceil(n_long / world)): the sphk_iou_pairwise_keys of its rows (entries past its own
 * slice are padding), then its keys of the short operand with GLOBAL row indices.
 *   gathered [world, cap + n_short]  the all-gathered blocks (for world = 1: the single block)
 *   long_max / long_arg [n_long]     per long-operand box: max over the short set / its index  (int64, as torch.max)
 *   short_max / short_arg [n_short]  per short-operand box: max over ALL shards / global index, ties -> lowest index
 * Contract reproduced: overlaps.max(dim=0) / .max(dim=1) of mmdet/core/bbox/assigners/max_iou_assigner.py:173-176 on
 * the unsharded matrix.  One launch; a key of 0 ("no positive overlap") reads as (0.0, index 0). */
int sphk_unpack_gathered_keys(const uint64_t* gathered, int32_t world, int64_t n_long, int64_t n_short, int64_t cap,
                              float* long_max, int64_t* long_arg, float* short_max, int64_t* short_arg, void* stream);

/* The sharded sweep WITHOUT a collective: the exchange is fused into the compute launch (stores into the peers' buffers
 * over NVLink while the launch runs) and into the unpack launch (flag handshake + reads).  Every rank owns a symmetric
 * buffer of identical layout (uint64 elements), mapped into all processes of the node
 * (torch.distributed._symmetric_memory):
 *     [ parity 0: world slots | parity 1: world slots | flags : >= world ]
 *     slot s = the block of rank s = [ `parts` arrays of cap keys of its rows | its n_short keys of the short operand ]
 * peer_bufs is a DEVICE array of the `world` base pointers (the handle's buffer_ptrs_dev).
 *
 * sphk_iou_pairwise_keys_push: the fused max / argmax sweep of sphk_iou_pairwise_keys for rows = the rank's shard of the
 * long operand.  The column keys go to col_keys (the short-operand part of the rank's own slot) as before; the row keys
 * are NOT merged over the column tiles: every CTA stores the 32 keys of its (row tile, column tile) as partial array
 * t = column tile (0 <= t < sphk_key_push_parts(C) = ceil(C / 256)) at element push_offset + t * part_stride + i of the
 * buffer of EVERY rank, its own included (plain 8-byte stores while the launch runs; part_stride >= R, normally cap).
 * The maximum over the partial arrays is the row key.  Needs no zero-fill: every partial entry of rows [0, R) is written.
 * Replaces: nothing in the reference (it is single-device on this path); the contract reproduced is the
 * single-device overlaps.max(dim=0/1) of mmdet/core/bbox/assigners/max_iou_assigner.py:173-176. */
int32_t sphk_key_push_parts(int64_t C);
int sphk_iou_pairwise_keys_push(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                                uint64_t* col_keys, int32_t row_base, int32_t col_base, uint64_t* const* peer_bufs, int32_t world,
                                int64_t push_offset, int64_t part_stride, void* workspace, void* stream);

/* The exchange step after the rank's compute kernel of step s (same stream): raises flags[rank] = s in every peer's
 * buffer, waits until all flags of the local buffer have reached s, then writes the same outputs as
 * sphk_unpack_gathered_keys.  block_offset = the step's parity (s & 1) times world * (long_parts * cap + n_short).
 * long_pushed != 0: the long operand's keys are the maximum over the long_parts partial arrays of the LOCAL slots (filled
 * by every rank's sphk_iou_pairwise_keys_push); long_pushed == 0 (long_parts = 1): they are read from slot s of rank s's
 * buffer by peer loads (filled by sphk_iou_pairwise_keys).  The short operand's keys (final only when the owner's kernel
 * has ended) always come from their owners' slots.  `step` counts from 1 and must advance by one per call on every rank;
 * the flags must be zero before the first step.  A peer that does not arrive within ~5 s makes the kernel trap (the job
 * fails instead of hanging). */
int sphk_unpack_peer_keys(const uint64_t* const* peer_bufs, int32_t rank, int32_t world, uint64_t step, int64_t block_offset,
                          int64_t flag_offset, int64_t n_long, int64_t n_short, int64_t cap, int32_t long_parts, int32_t long_pushed,
                          float* long_max, int64_t* long_arg, float* short_max, int64_t* short_arg, void* stream);

/* Second pass of MaxIoUAssigner's low-quality matching with gt_max_assign_all=True
 * (mmdet/core/bbox/assigners/max_iou_assigner.py:201-205: for each GT i in ascending order,
 * `assigned[overlaps[i, :] == gt_max_overlaps[i]] = i + 1`) without the K x N matrix:
 *   row_target [R]  the row maxima of a previous sphk_iou_pairwise call on the SAME operands (same kernel,
 *                   bit-identical values); a negative entry disables that row (gt_max < min_pos_iou)
 *   col_tie    [C]  out: the largest row_base + i + 1 over the rows i with IoU(i, j) == row_target[i] > 0,
 *                   0 if none.  Rows whose target is exactly 0 must be handled by the caller. */
int sphk_iou_pairwise_ties(int kind, const float* rows, int64_t R, const float* cols, int64_t C, int D, int mode, int edge,
                           const float* row_target, int32_t* col_tie, int32_t row_base, void* workspace, void* stream);

/* MaxIoUAssigner.assign for a batch of images that share one anchor list, without the K x N matrices
 * (mmdet/core/bbox/assigners/max_iou_assigner.py:67-220; per-image loop mmdet/models/dense_heads/anchor_head.py:368-377):
 *   gts             [sumK, D]   ground truths of all images, concatenated (bboxes1 role, as :113 calls the calculator)
 *   gt_offsets_host [batch+1]   HOST array, image b owns gts[gt_offsets[b] .. gt_offsets[b+1])
 *   boxes           [N, D]      anchors / proposals (bboxes2 role), shared by the images
 *   neg_iou_lo/hi               the negative range [lo, hi): (0, neg_iou_thr) for a float threshold (:178-184)
 *   gt_labels       [sumK] int64 or NULL;  labels [batch, N] int64 or NULL
 *   gt_inds [batch, N] int64: -1 ignore, 0 background, i+1 = assigned to the image's GT i;  max_overlaps [batch, N]
 * Launches: records, pass 1 (fused max/argmax), targets, pass 2 (ties, only for gt_max_assign_all), epilogue. */
int64_t sphk_max_iou_assign_workspace_bytes(int64_t sumK, int64_t N, int32_t batch);
int sphk_max_iou_assign(int kind, const float* gts, const int32_t* gt_offsets_host, int32_t batch, const float* boxes,
                        int64_t N, int D, float pos_iou_thr, float neg_iou_lo, float neg_iou_hi, float min_pos_iou,
                        int gt_max_assign_all, int match_low_quality, const int64_t* gt_labels, int64_t* gt_inds,
                        float* max_overlaps, int64_t* labels, void* workspace, void* stream);

/* Training targets of the anchor head for a batch of images that share the anchors, PseudoSampler case
 * (mmdet/models/dense_heads/anchor_head.py:254-285 _get_targets_single; mmdet/core/bbox/samplers/pseudo_sampler.py),
 * computed from the assignment of sphk_max_iou_assign without leaving the device:
 *   gt_inds      [batch, N]     0 = negative, -1 = ignored, k + 1 = positive for GT k of its image
 *   anchors [N, D], gts [sumK, D] (all images), gt_labels [sumK] or NULL (then positives get label 0),
 *   gt_offsets   [batch + 1]    DEVICE array: image b owns gts[gt_offsets[b] : gt_offsets[b + 1]]
 *   num_classes  background label; pos_weight: train_cfg.pos_weight (<= 0: positives weigh 1)
 *   reg_decoded_bbox != 0: bbox_targets = the assigned GT box; else bbox_coder.encode(anchor, GT) with the HOST arrays
 *                means / stds (D floats each, NULL = 0 / 1)
 *   labels [batch, N] int64, label_weights [batch, N], bbox_targets / bbox_weights [batch, N, D]
 *   counts [batch, 2] int32: positives, negatives per image (zeroed here) */
int sphk_anchor_targets(const int64_t* gt_inds, int32_t batch, int64_t N, int D, const float* anchors, const float* gts,
                        const int64_t* gt_labels, const int32_t* gt_offsets, int64_t num_classes, float pos_weight,
                        int reg_decoded_bbox, const float* means, const float* stds, int64_t* labels, float* label_weights,
                        float* bbox_targets, float* bbox_weights, int32_t* counts, void* stream);

/* Sph2Pob loss, fused forward + backward (Sph2PobIoULoss, mode='iou'):
 * replaces Sph2PobTransfrom.new_forward (sphdet/losses/sph2pob_transform.py:24-35: jitter_1,
 * sph2pob_standard, jitter_2) followed by diff_iou_rotated_2d(...).clamp(0,1)
 * (sphdet/losses/sph2pob_iou_loss.py:122) and its autograd backward.
 *   iou        [n]      clamped IoU of (pred[i], target[i])
 *   grad_iou   [n]      upstream d(total)/d(iou[i]); NULL = 1 for every row (the outputs are then
 *                       d(iou[i])/d(box) itself)
 *   grad_pred  [n, D]   d(total)/d(pred)   (NULL to skip)
 *   grad_target[n, D]   d(total)/d(target) (NULL to skip); both NULL = forward only  */
int sphk_loss_fwd_bwd(const float* pred, const float* target, int64_t n, int D, float* iou, const float* grad_iou,
                      float* grad_pred, float* grad_target, void* stream);

/* Reduced Sph2Pob IoU loss in one launch: Sph2PobIoULoss(mode='iou') with reduction 'mean' / 'sum'
 * (sphdet/losses/sph2pob_iou_loss.py:25-58,138-140 + mmdet/models/losses/utils.py weight_reduce_loss):
 *   partial [sphk_loss_reduce_partials(n)]  per-block sums of weight[i] * (1 - iou[i]); the loss is
 *                                           scale * sum(partial)  (scale = loss_weight / n | / (avg_factor+eps) | 1)
 *   weight  [n] or NULL (= 1)
 *   grad_pred / grad_target [n, D] or NULL: d(loss)/d(box) = -weight[i] * scale * d(iou[i])/d(box) */
int64_t sphk_loss_reduce_partials(int64_t n);
int sphk_loss_reduce(const float* pred, const float* target, const float* weight, int64_t n, int D, float scale,
                     float* partial, float* grad_pred, float* grad_target, void* stream);

/* The same launch, finished on the device: *total = scale * sum_i weight[i] * (1 - iou[i]) (one float), so that the
 * loss of a training step is ONE kernel and no reduction op follows.  The block that finishes last adds the per-block
 * sums in index order (deterministic).  `scratch`: sphk_loss_total_scratch_bytes(n) bytes, 16-byte aligned, whose FIRST
 * 16 bytes (the ticket counter) must be zero when the call starts; the kernel leaves them zero, so a buffer that was
 * zeroed once can serve every later call of the same stream, whatever its n. */
int64_t sphk_loss_total_scratch_bytes(int64_t n);
int sphk_loss_reduce_total(const float* pred, const float* target, const float* weight, int64_t n, int D, float scale,
                           float* total, void* scratch, float* grad_pred, float* grad_target, void* stream);

/* The same two stages exposed separately so that the GIoU/DIoU/CIoU epilogues
 * (sphdet/losses/sph2pob_iou_loss.py:142-194) can stay as autograd code on the OBBs:
 *   sphk_obb_fwd : (pred,target)[n,D] -> obb1, obb2 [n,5] = (x, y, w, h, angle rad) after
 *                  jitter_1 + transform(kind) + jitter_2
 *   sphk_obb_bwd : grads of obb1/obb2 -> grads of pred/target (either may be NULL)
 *   sphk_riou_fwd_bwd : rotated IoU of OBB pairs (diff_iou_rotated_2d, sphdet/iou/diff_iou_rotated.py:325-343)
 *                  and, when grad_iou != NULL, its gradient w.r.t. both OBBs. */
int sphk_obb_fwd(int kind, const float* b1, const float* b2, int64_t n, int D, int edge, float* obb1, float* obb2,
                 void* stream);
int sphk_obb_bwd(int kind, const float* b1, const float* b2, int64_t n, int D, int edge, const float* grad_obb1,
                 const float* grad_obb2, float* grad_b1, float* grad_b2, void* stream);
int sphk_riou_fwd_bwd(const float* obb1, const float* obb2, int64_t n, float* iou, const float* grad_iou,
                      float* grad_obb1, float* grad_obb2, void* stream);

/* The other regression losses on the Sph2Pob OBBs, forward and backward in ONE launch:
 *   Sph2PobGDLoss (sphdet/losses/sph2pob_gd_loss.py:7-26; mmrotate 0.3.2 GDLoss: gwd, kld, jd, kld_symmax, kld_symmin),
 *   Sph2PobKFLoss (sphdet/losses/sph2pob_kf_loss.py:8-26; mmrotate 0.3.2 KFLoss, decoded boxes swapped as at :26),
 *   Sph2PobL1Loss (sphdet/losses/sph2pob_l1_loss.py:9-94; mmdet L1Loss on bbox2delta of the OBBs),
 * each behind the Sph2PobTransfrom decorator (sphdet/losses/sph2pob_transform.py:11-37): jitter_1 -> transform ->
 * jitter_2 per row.  L = 5 columns per row for SPHK_LOSS_L1, else 1.
 *   fun      GD: 0 none, 1 log1p, 2 sqrt;  KF: 0 none, 1 ln, 2 exp
 *   flags    GD: bit0 = normalize (gwd) / sqrt (kld family);  L1: bit0 encode, bit1 swap, bit2 angle_modifier='modulus'
 *   tau, alpha  GDLoss;  beta, eps  KFLoss (mmrotate defaults 1/9, 1e-6)
 *   transform   SPHK_KIND_SPH2POB_STANDARD (the decorator's default) or _EFFICIENT
 *   upstream [n] (up_cols = 1), [n, L] (up_cols = L) or NULL (= 1): loss weights or d(total)/d(loss); times `scale`
 *   loss     [n, L] unweighted elementwise loss, or NULL
 *   partial  [sphk_loss_reduce_partials(n)] per-block sums of upstream * loss (without scale), or NULL
 *   grad_pred / grad_target [n, D]: scale * sum_j upstream[i, j] * d(loss[i, j])/d(box), or NULL
 * Rows whose upstream is entirely zero get a zero gradient and add nothing to `partial`. */
#define SPHK_LOSS_GWD 0
#define SPHK_LOSS_KLD 1
#define SPHK_LOSS_JD 2
#define SPHK_LOSS_KLD_SYMMAX 3
#define SPHK_LOSS_KLD_SYMMIN 4
#define SPHK_LOSS_KFIOU 5
#define SPHK_LOSS_L1 6
int sphk_obb_loss(int loss_kind, int fun, int flags, float tau, float alpha, float beta, float eps, int transform,
                  const float* pred, const float* target, int64_t n, int D, const float* upstream, int up_cols, float scale,
                  float* loss, float* partial, float* grad_pred, float* grad_target, void* stream);

/* The reduced loss of sphk_obb_loss finished on the device: *total = scale * sum_i sum_j upstream[i, j] * loss[i, j] as one
 * float written by the launch itself (the block that draws the last ticket adds the per-block sums in index order, as
 * sphk_loss_reduce_total does), next to the gradients of that total -- the forward of a training step is ONE launch, no
 * reduction op on the host side.  scratch: sphk_loss_total_scratch_bytes(n) bytes, 16-byte aligned, its first 16 bytes
 * zero before the first call (every call hands the ticket counter back at zero). */
int sphk_obb_loss_total(int loss_kind, int fun, int flags, float tau, float alpha, float beta, float eps, int transform,
                        const float* pred, const float* target, int64_t n, int D, const float* upstream, int up_cols, float scale,
                        float* total, void* scratch, float* grad_pred, float* grad_target, void* stream);

/* The spherical delta box coders: DeltaXYWHSphBBoxCoder (D = 4) and DeltaXYWHASphBBoxCoder (D = 5)
 * (sphdet/bbox/coder/delta_xywh_sph_bbox_coder.py:45-115,117-262; delta_xywha_rsph_bbox_coder.py:45-115,117-268).
 * means / stds: HOST arrays of D floats (target_means / target_stds); wh_ratio_clip, clip_border, add_ctr_clamp,
 * ctr_clamp as in the coder's constructor / decode().
 *   sphk_coder_decode     : out[n, D] = delta2bbox(rois[n, D], deltas[n, D])
 *   sphk_coder_decode_bwd : grad_deltas[n, D] = d(total)/d(deltas) from grad_out = d(total)/d(out) (zero where a clamp
 *                           of the decode is active, as torch.clamp's backward does)
 *   sphk_coder_encode     : out[n, D] = bbox2delta(proposals[n, D], gt[n, D]) */
int sphk_coder_decode(const float* rois, const float* deltas, int64_t n, int D, const float* means, const float* stds,
                      float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp, float* out, void* stream);
int sphk_coder_decode_bwd(const float* rois, const float* deltas, const float* grad_out, int64_t n, int D, const float* means,
                          const float* stds, float wh_ratio_clip, int clip_border, int add_ctr_clamp, float ctr_clamp,
                          float* grad_deltas, void* stream);
int sphk_coder_encode(const float* proposals, const float* gt, int64_t n, int D, const float* means, const float* stds,
                      float* out, void* stream);

/* The regression branch of the head's loss with reg_decoded_bbox=True, in one launch
 * (sphdet/models/heads/sph_retina_head.py:252-265: bbox_coder.decode(anchors, bbox_pred) then
 * Sph2PobIoULoss(mode='iou')(pred, target, weight, avg_factor)).  It is called on ALL anchors of a
 * batch with zero weights for the negatives: rows whose weight is 0 are skipped (loss 0 * finite, gradient 0 * finite),
 * the others are compacted and evaluated 32 at a time: decode, jitter_1, sph2pob_standard, jitter_2, rotated IoU, and
 * the whole backward down to the deltas.
 *   anchors, deltas, target [n, D]
 *   weight   NULL (= 1), [n] (weight_cols = 1) or [n, weight_cols] (row mean, sph2pob_iou_loss.py:43-48;
 *            weight_cols = D as the head passes bbox_weights)
 *   partial  [sphk_decode_loss_partials(n)] per-CTA sums of weight[i] * (1 - iou[i]); loss = scale * sum(partial)
 *   grad_deltas [n, D] or NULL: d(loss)/d(deltas) (every row is written) */
int64_t sphk_decode_loss_partials(int64_t n);
int sphk_decode_loss_reduce(const float* anchors, const float* deltas, const float* target, const float* weight,
                            int weight_cols, int64_t n, int D, const float* means, const float* stds, float wh_ratio_clip,
                            int clip_border, int add_ctr_clamp, float ctr_clamp, float scale, float* partial,
                            float* grad_deltas, void* stream);

/* Box format conversions either side of the IoU path (sphdet/bbox/box_formator.py), one launch, row-wise:
 *   fmt  0 xyxy2xywh (:17-23)      1 xywh2xyxy (:25-31)       2 obb2hbb_wywh (:33-50, [n,5] -> [n,4])   3 obb2hbb_xyxy (:52-55)
 *        4 bfov2rbfov (:57-61)     5 geo2sph (:64-68)         6 sph2geo (:70-74)                        7 sph2pix (:77-84)
 *        8 pix2sph (:86-93)        9 sph2tan (:99-107)       10 tan2sph (:109-117)
 *       11 / 12 Sph2PlanarBoxTransform('sph2pix' / 'sph2tan') (:166-182): [n,4] -> xyxy, [n,5] -> (x, y, w, h, -gamma rad)
 *       13 / 14 Planar2SphBoxTransform(pix / tan) (:185-200): xyxy -> bfov [n,4] or rbfov [n,5] (gamma = 0)
 *   in [n, d_in], out [n, d_out]; img_h, img_w: the equirectangular image size of the pix / tan formats ((512, 1024) in the
 *   reference).  The pure-arithmetic formats are bit-identical to the reference's fp32 torch expressions. */
int sphk_box_format(int fmt, const float* in, int64_t n, int d_in, int d_out, float img_h, float img_w, float* out, void* stream);

/* Batched greedy spherical NMS (SphNMS / sph_batched_nms / sph_nms_op,
 * sphdet/bbox/nms/sph_nms.py:22-74).  kind: the IoU SphNMS was built with (sph_nms.py:8-16) -- SPHK_KIND_SPH2POB_EFFICIENT
 * (its default), SPHK_KIND_NAIVE (the reference's indoor360 configs: test_cfg.iou_calculator = 'naive_iou') or
 * SPHK_KIND_UNBIASED (its pandora configs: 'unbiased_iou').  SPHK_KIND_NAIVE | SPHK_NMS_RULE_GT is mmcv's `nms` on the
 * sph2pix boxes, i.e. PlanarNMS (sphdet/bbox/nms/planar_nms.py:7-18): same IoU, but suppression iff IoU > threshold, so
 * that a NaN IoU (two zero-area boxes) KEEPS the box as mmcv's `inter > thr * union` does, where SphNMS's
 * `ious <= thr` (:70) drops it.
 *   boxes       [M, D]
 *   order       [M]    int32 indices into boxes, grouped by segment (one segment = one
 *                      (image, class) group), score-descending inside a segment (:65)
 *   seg_offsets [S+1]  int32, segment s covers order[seg_offsets[s] .. seg_offsets[s+1])
 *   max_seg_len        an upper bound of the segment lengths (sizes the per-CTA shared memory;
 *                      a longer segment is refused: its keep bytes are set to 0xFF)
 *   typical_seg_len    0, or the usual segment length when max_seg_len is only a loose bound (sizes the CTA)
 *   keep        [M]    uint8, keep[q] = 1 iff the box order[q] survives (same positions as `order`)
 * A box is suppressed iff IoU(pivot as bboxes1, box as bboxes2) > iou_threshold (:70-73). */
int sphk_nms_batched(const float* boxes, const int32_t* order, const int32_t* seg_offsets, int32_t S,
                     int32_t max_seg_len, int32_t typical_seg_len, int D, int kind, float iou_threshold, uint8_t* keep,
                     void* stream);

/* The same NMS for a test-time batch laid out as `num_images` equal blocks of `per_image` candidates (what the head's
 * post-processing produces: nms_pre candidates per level and image, sph_retina_head.py:169-212 then :35-101), with no
 * host-side sort and no synchronisation: per image one CTA sorts (label, score descending) in shared memory and emits the
 * (image, class) segments, `k_nms` runs on them, and a second per-image CTA orders the survivors of all classes by
 * descending score.
 *   boxes [M, D], scores [M], labels [M] int64 in [0, num_classes), M = num_images * per_image (<= 16384 per image)
 *   valid [M] uint8 or NULL: 0 leaves a candidate out (padding, scores under score_thr)
 *   out_idx   [num_images, max_out] int32: indices into boxes of the kept candidates of each image, score-descending
 *             (equal scores: lower index first), -1 beyond the count -- max_out plays nms_cfg.max_num / max_per_img
 *   out_count [num_images] int32; -1 for an image that holds a wanted candidate whose label is outside
 *             [0, num_classes): its result must not be used (the caller re-runs it through sphk_nms_batched)
 *   workspace sphk_nms_images_workspace_bytes(num_images, per_image, num_classes) bytes, 16-byte aligned */
int64_t sphk_nms_images_workspace_bytes(int32_t num_images, int32_t per_image, int32_t num_classes);
int sphk_nms_images(const float* boxes, const float* scores, const int64_t* labels, const uint8_t* valid, int32_t num_images,
                    int32_t per_image, int32_t num_classes, int D, int kind, float iou_threshold, int32_t max_out,
                    int32_t* out_idx, int32_t* out_count, void* workspace, void* stream);

/* Measurement helpers (bench.py; no counterpart in the reference).
 * sphk_probe_fp32: FMA-chain microbenchmark that yields the FP32 CUDA-core peak the Sph2Pob kernels
 *   are bounded by (SURVEY.md 8d asks for a measured denominator): launches `blocks` x 256 threads,
 *   each running 8 independent chains of `iters` FMAs, i.e. 2*8*iters*256*blocks flop; sink[blocks*256].
 * sphk_set_dense: 1 disables the parity-safe "disjoint pair" early-outs of the Sph2Pob kernels so the
 *   dense throughput can be reported next to the real one; returns the previous setting.
 * sphk_prefilter_count: *live_count (device) = number of pairs of rows[R] x cols[C] that survive the prefilter of
 *   the N x M kernels (circumscribed-circle test, then the box-frame or the separating-axis test, whichever the kernel
 *   instance of a call of that shape runs: csrc/sphk_fast.cuh), i.e. that the expensive
 *   transform + clipping code is run for; early-out rate = 1 - live / (R * C).  Workspace as sphk_iou_pairwise. */
int sphk_probe_fp32(int32_t blocks, int32_t iters, float* sink, void* stream);
int sphk_set_dense(int on);
int sphk_prefilter_count(const float* rows, int64_t R, const float* cols, int64_t C, int D, int edge, uint64_t* live_count,
                         void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPHK_H_ */
