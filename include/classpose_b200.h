/*
 * classpose_b200 -- C ABI of the B200-native Classpose post-network path.
 *
 * One shared library (libclasspose_b200.so), plain pointers and sizes, no torch types.
 * Every *_device entry point takes DEVICE pointers, a caller-owned device workspace
 * (size from cpb_workspace_bytes) and the caller's CUDA stream (cudaStream_t passed as
 * void*; NULL = legacy default stream).  Calls are asynchronous on that stream, never
 * synchronise the host, keep no global state and may be issued concurrently from
 * several host threads as long as each call has its own workspace (the reference runs
 * >= 2 inference threads per process: predict_wsi.py:728-797).
 * The *_host entry points take HOST pointers and perform the H2D / D2H copies
 * themselves (chunked, double-buffered on internal streams); they return when the
 * outputs are in host memory.
 *
 * Return value: 0 on success, a negative CPB_E_* code on argument errors, or a
 * positive cudaError_t if a launch/copy failed.  There is no CPU fallback.
 *
 * Each entry point cites the reference interface it stands in for.  The arithmetic of
 * rows (2)-(5) lives in cellpose==4.0.8 (not vendored by the reference); citations for
 * those are the reference's call sites plus SURVEY.md Appendix A.
 *
 * Label images are int32 [B,H,W], 0 = background, instance ids 1..n per tile (n in
 * counts[b]).  Flow fields are float32 [B,2,H,W] (dY, dX) at network scale (5x unit
 * vectors), cell probabilities float32 [B,H,W], class logits float32 [B,C,H,W].
 */
#ifndef CLASSPOSE_B200_H
#define CLASSPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPB_ABI_VERSION 2

#define CPB_E_ARG       (-1) /* bad shape / null pointer */
#define CPB_E_WORKSPACE (-2) /* workspace too small */
#define CPB_E_RANGE     (-3) /* B*H*W does not fit the 31-bit pixel index (split the batch), or one tile exceeds
                                 2^24 padded pixels (about 4090 x 4090) in follow_flows */
#define CPB_E_CAPACITY  (-4) /* a tile exhausted an internal pool (hole fill of labels spanning more than ~512 x 512
                                 pixels draws its bitmaps from a bounded pool).  *_device calls are asynchronous and
                                 report this as counts[b] = -1 for the tile; *_host calls return this code. */

/* Parameters of dynamics.resize_and_compute_masks as called at
 * /root/reference/src/classpose/models.py:149-159; defaults from models.py:490-498,751-752. */
typedef struct cpb_params {
    int32_t niter;              /* 200 */
    float   cellprob_threshold; /* 0.0 */
    double  flow_threshold;     /* 0.4; <= 0 disables the flow-error check */
    int32_t min_size;           /* 15; <= 0 skips the size filters (holes are still filled) */
    double  max_size_fraction;  /* 0.4 */
    int32_t remove_border;      /* 0; 1 = also drop instances touching the tile border
                                   (metrics/pq.py:65-92), applied last */
    int32_t fill_holes;         /* 1 = resize_and_compute_masks contract; 0 = dynamics.compute_masks
                                   contract with min_size=-1 (no hole fill / size filter) */
} cpb_params;

int  cpb_abi_version(void);
/* label-table capacity (max label value + 1 a tile can produce) used by the fused path */
int  cpb_label_capacity(int H, int W);
/* bytes of device workspace needed by any *_device call on a [B,H,W] batch with C classes
 * and label values <= lcap-1 (pass lcap = 0 for the fused-path default) */
size_t cpb_workspace_bytes(int B, int H, int W, int C, int lcap);

/* ---- fused path: dP, cellprob (, logits) -> masks (, per-cell class) --------------------
 * Replaces: classpose.models.compute_masks 2-D branch (models.py:97-188, one plane per
 * tile) + compute_class_masks (models.py:191-230) as called from ClassposeModel.eval
 * (models.py:750-770).
 * logits may be NULL (then cell_class / class_masks are not written).
 * masks [B,H,W] int32, counts [B] int32 (instances per tile),
 * cell_class [B,lcap] int32 (entry l = class of instance l, entry 0 = 0; only 0..counts[b]
 * are meaningful), class_masks [B,H,W] uint8 or NULL. */
int cpb_compute_masks_device(const float* dP, const float* cellprob, const float* logits,
                             int B, int H, int W, int C, const cpb_params* prm,
                             int32_t* masks, int32_t* counts, int32_t* cell_class,
                             uint8_t* class_masks, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Measurement aid: the same fused path with CUDA events around each stage; synchronises the stream
 * and writes cpb_num_stages() device times in milliseconds to stage_ms (HOST pointer). */
int cpb_num_stages(void);
const char* cpb_stage_name(int i);
int cpb_compute_masks_profiled_device(const float* dP, const float* cellprob, const float* logits,
                                      int B, int H, int W, int C, const cpb_params* prm,
                                      int32_t* masks, int32_t* counts, int32_t* cell_class,
                                      uint8_t* class_masks, void* workspace, size_t workspace_bytes,
                                      void* stream, float* stage_ms);
/* follow_flows variant: 2 = trajectory pool with many merge points (default), 1 = two merge points per chunk,
 * 0 = plain kernel, -1 = CPB_FOLLOW_MERGE from the environment.  All give bit-identical results; the switch
 * exists for A/B measurements and tests. */
void cpb_debug_set_follow_merge(int mode);
/* A/B switches of the fused path (value 1 = on, 0 = off, -1 = environment variable of the same name, default on
 * unless stated).  Every setting gives the same results (CPB_BLEND_EFT: to one ulp; CPB_FILL_EXACT: see below); they
 * exist for measurements and tests. */
#define CPB_SWITCH_FOLLOW_MERGE 0   /* CPB_FOLLOW_MERGE: values 0 / 1 / 2 as cpb_debug_set_follow_merge */
#define CPB_SWITCH_DIFFUSE_QUEUE 1  /* CPB_DIFFUSE_QUEUE: diffusion warps pull label pairs from a queue */
#define CPB_SWITCH_QC_FUSED 2       /* CPB_QC_FUSED: flow error of isolated labels inside the diffusion warp */
#define CPB_SWITCH_VOTE_FUSED 3     /* CPB_VOTE_FUSED: class vote folded into the final label pass */
#define CPB_SWITCH_QC_SCREEN 4      /* CPB_QC_SCREEN: float32 screen with a proven error bound in front of the float64
                                       flow check; labels it cannot decide take the float64 path (same removal set) */
#define CPB_SWITCH_BLEND_EFT 5      /* CPB_BLEND_EFT: taper blend in float32 with error-free transformations (0: float64
                                       arithmetic per element, numpy's literal sequence; results agree to one ulp on
                                       ~1e-6 of the elements) */
#define CPB_SWITCH_FOLLOW_SMALL 6   /* CPB_FOLLOW_SMALL: 256-entry chunks in the trajectory pool when the batch is a handful of
                                       tiles (latency form; 0: 1024-entry chunks always) */
#define CPB_SWITCH_SEED_CANDS 7     /* CPB_SEED_CANDS: the trajectory kernel lists the bins that pass 10 end points while it counts
                                       them (0: a separate pass streams the whole histogram to find them) */
#define CPB_SWITCH_FILL_EXACT 8     /* CPB_FILL_EXACT (default 0, unlike the others): replay upstream's label-by-label hole fill on
                                       tiles where a label lies partly inside another label's hole; validated on the CPU
                                       simulator only, see INTEGRATION.md section 4 */
void cpb_debug_set_switch(int which, int value);
/* number of kernels this library has launched in this process (statistics for the benchmark) */
long long cpb_debug_launch_count(void);
/* flow-check counters of the last cpb_compute_masks_profiled_device call in this process, 24 ints: [0] float32 screen
 * jobs, [2] labels that took the float64 warp kernel (contact, too large for the screen, or undecided), [4] labels the
 * screen decided, [5] labels the screen left undecided (statistics for the benchmark and the tests) */
void cpb_debug_qc_stats(int32_t* out);

/* Same, HOST buffers in / out (pageable or pinned).  tiles_per_chunk <= 0 picks a default.
 * device = CUDA device ordinal.  Equivalent to cpb_compute_masks_host_ex with default options. */
int cpb_compute_masks_host(const float* dP, const float* cellprob, const float* logits,
                           int B, int H, int W, int C, const cpb_params* prm,
                           int32_t* masks, int32_t* counts, int32_t* cell_class,
                           uint8_t* class_masks, int tiles_per_chunk, int device);

/* The host path is bound by the bytes that cross PCIe; the options say how to spend them.  This is the call the
 * e2e benchmark times. */
#define CPB_HOST_LOGITS_AUTO   0  /* mapped when the logits buffer is pinned / registered host memory, else upload */
#define CPB_HOST_LOGITS_UPLOAD 1  /* copy every logit to the device (4*C bytes per pixel) */
#define CPB_HOST_LOGITS_MAPPED 2  /* the final label pass reads the logits through the mapped host pointer and only
                                     touches the 4-pixel groups that hold a cell; CPB_E_ARG if the buffer is pageable */
typedef struct cpb_host_options {
    int32_t tiles_per_chunk;   /* <= 0: default (128) */
    int32_t device;            /* CUDA device ordinal */
    int32_t logits_mode;       /* CPB_HOST_LOGITS_* */
    int32_t flows_mode;        /* same values, for dP: mapped = the prep kernel reads dP in place and only the 4-pixel
                                  groups that hold foreground (cellprob > threshold) cross the bus; they are kept in a
                                  device copy for the flow check.  cellprob is always uploaded (every pixel is needed) */
    int32_t masks_u16;         /* 1: `masks` is uint16 [B,H,W] (Cellpose's dtype below 65536 labels; ids are
                                  truncated to 16 bits), 0: int32 */
} cpb_host_options;
int cpb_compute_masks_host_ex(const float* dP, const float* cellprob, const float* logits,
                              int B, int H, int W, int C, const cpb_params* prm,
                              void* masks, int32_t* counts, int32_t* cell_class,
                              uint8_t* class_masks, const cpb_host_options* opt);

/* ---- single-tile plans: the per-tile call of the reference's WSI loop -----------------------------------------
 * The loop calls model.eval([tile]) one tile at a time from two inference threads per process
 * (predict_wsi.py:728-797), which reaches compute_masks (models.py:464 -> :97) and then compute_class_masks
 * (models.py:766 -> :191) with host arrays.  A plan owns pinned staging buffers, device buffers and CUDA graphs for
 * "upload -> fused path -> download" and "upload logits -> vote on the labels still on the device -> download", so a
 * call costs one copy into the staging buffer, one graph launch and one synchronise.  One plan per host thread and
 * tile shape; the calls of one plan must not overlap.
 *   cpb_tile_plan_dp / _cellprob   pinned float32 [2,H,W] / [H,W] the caller fills before cpb_tile_plan_run
 *   cpb_tile_plan_masks            pinned int32 [H,W], valid after cpb_tile_plan_run until the next run
 *   cpb_tile_plan_logits(p, C)     pinned float32 [C,H,W] to fill before cpb_tile_plan_vote (sets up the vote for C classes)
 *   cpb_tile_plan_vote             class image uint8 [H,W] and per-instance classes (first min(lcap, 1024) entries),
 *                                  both pinned, for the labels of the LAST cpb_tile_plan_run */
typedef struct cpb_tile_plan cpb_tile_plan;
int  cpb_tile_plan_create(int H, int W, const cpb_params* prm, int device, cpb_tile_plan** out);
void cpb_tile_plan_destroy(cpb_tile_plan* plan);
float* cpb_tile_plan_dp(cpb_tile_plan* plan);
float* cpb_tile_plan_cellprob(cpb_tile_plan* plan);
const int32_t* cpb_tile_plan_masks(cpb_tile_plan* plan);
int  cpb_tile_plan_run(cpb_tile_plan* plan, int32_t* count);
float* cpb_tile_plan_logits(cpb_tile_plan* plan, int C);
int  cpb_tile_plan_vote(cpb_tile_plan* plan, const uint8_t** class_masks, const int32_t** cell_class);

/* ---- stage entry points (each is one row of SURVEY.md section 8a) ------------------------ */

/* (2) dynamics.follow_flows / steps_interp on dP*(cellprob>thr)/5 (SURVEY A.2-A.3;
 * reached from models.py:149).  p_final [B,H,W] int32: (y<<16)|x of the truncated end
 * point of every foreground pixel, -1 for background.  p_float (optional, may be NULL)
 * [B,2,H,W] float32 un-truncated (y,x), undefined on background. */
int cpb_follow_flows_device(const float* dP, const float* cellprob, int B, int H, int W,
                            int niter, float cellprob_threshold, int32_t* p_final,
                            float* p_float, void* workspace, size_t workspace_bytes, void* stream);

/* (3) dynamics.get_masks_torch: end points -> seeds -> labels, big-mask
 * removal, first-appearance renumbering. */
int cpb_get_masks_device(const int32_t* p_final, int B, int H, int W, double max_size_fraction,
                         int32_t* masks, int32_t* counts, void* workspace, size_t workspace_bytes,
                         void* stream);

/* (4a) dynamics.masks_to_flows: float64 heat diffusion from each instance's
 * centre; mu [B,2,H,W] float64 unit vectors (0 on background).  lcap > max label value. */
int cpb_masks_to_flows_device(const int32_t* masks, int B, int H, int W, int lcap, double* mu,
                              void* workspace, size_t workspace_bytes, void* stream);

/* (4) dynamics.remove_bad_flow_masks: zero instances whose mean squared
 * difference between mask-derived flows and dP/5 exceeds `threshold`; no renumbering.
 * flow_err (optional) [B,lcap] float64 receives the per-label error. */
int cpb_remove_bad_flow_masks_device(int32_t* masks, const float* dP, int B, int H, int W,
                                     int lcap, double threshold, double* flow_err,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* (5) utils.fill_holes_and_remove_small_masks (also models.py:172-174). */
int cpb_fill_holes_and_remove_small_masks_device(int32_t* masks, int B, int H, int W, int lcap,
                                                 int min_size, int32_t* counts, void* workspace,
                                                 size_t workspace_bytes, void* stream);

/* (6) compute_class_masks (models.py:191-230): per-pixel arg-max class, per-instance
 * majority (ties -> lowest class, class 0 may win, label 0 -> class 0).
 * cell_class [B,lcap] int32; class_masks [B,H,W] uint8 or NULL. */
int cpb_class_vote_device(const int32_t* masks, const float* logits, int B, int H, int W, int C,
                          int lcap, int32_t* cell_class, uint8_t* class_masks, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Same vote with the per-tile label bound taken from a DEVICE array (counts [B], e.g. the one the fused path wrote)
 * instead of lcap - 1: no host knowledge of the counts is needed, so the call can sit in a CUDA graph behind
 * cpb_compute_masks_device, and the (instance, class) table stays in shared memory for nuclei-scale tiles. */
int cpb_class_vote_counts_device(const int32_t* masks, const float* logits, const int32_t* counts, int B, int H,
                                 int W, int C, int lcap, int32_t* cell_class, uint8_t* class_masks,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* (7) metrics.pq.remove_border_instances (pq.py:65-92).  masks [B,H,W,nch] int32, channel 0
 * holds the instance ids, every channel is zeroed for border instances (nch = 1 for a plain
 * label image).  In place, as the reference mutates its argument. */
int cpb_remove_border_instances_device(int32_t* masks, int B, int H, int W, int nch, int lcap,
                                       void* workspace, size_t workspace_bytes, void* stream);

/* (1) transforms.average_tiles (core.py:215,218-220) fused with
 * unaugment_tiles / unaugment_class_tiles (core.py:207-214; transforms/transforms.py:4-21)
 * and the crop of core.py:226-229.
 * y [B,ntiles,nch,ly,lx] float32 network outputs per sub-tile; y0/x0 [ntiles] window origins;
 * flip [ntiles] bit0 = tile was flipped in Y, bit1 = in X (0 when augment=False);
 * negate_flow != 0: channel 0 changes sign on Y flips and channel 1 on X flips (flow maps);
 * taper_y [ly], taper_x [lx] float64 1-D taper factors (the 2-D weight is their product);
 * output yf [B,nch,Ly-cy0-cy1,Lx-cx0-cx1] float32 = blended map cropped by (cy0,cy1,cx0,cx1).
 * geometry arrays and tapers are DEVICE pointers. */
int cpb_average_tiles_device(const float* y, int B, int ntiles, int nch, int ly, int lx,
                             const int32_t* y0, const int32_t* x0, const int32_t* flip,
                             int negate_flow, const double* taper_y, const double* taper_x,
                             int Ly, int Lx, int cy0, int cy1, int cx0, int cx1, float* yf,
                             void* stream);

/* Same, with a host-side promise that enables the 128-bit path: every window origin x0[j] is a multiple of 4
 * (x0_multiple_of_4 != 0; device arrays are not read back by the library).  max_cover (maximum number of
 * windows over one pixel) is informational.  Pass 0, 0 when unknown. */
int cpb_average_tiles_ex_device(const float* y, int B, int ntiles, int nch, int ly, int lx,
                                const int32_t* y0, const int32_t* x0, const int32_t* flip,
                                int negate_flow, const double* taper_y, const double* taper_x,
                                int Ly, int Lx, int cy0, int cy1, int cx0, int cx1, float* yf,
                                int x0_multiple_of_4, int max_cover, void* stream);

/* ---- next row N3 / north_star (1): the tail of ClassposeModel.eval on the network's sub-tile outputs -------------
 * Replaces core.py:197-231 (un-flip, average_tiles of the flow map and of the class logits, crop) + models.py:750-770
 * (masks, class vote) in one call on device data.  The blend of the flow map is FUSED with the cellprob threshold: the
 * thread that blends (dY, dX, cellprob) of a pixel group also emits the foreground list, the masked / scaled flow field of
 * follow_flows and the zeroed label image, so the blended maps are never re-read by a separate first pass.
 *   y_flows  [B,ntiles,3,ly,lx] float32 (dY, dX, cellprob per sub-tile, as the network emits them)
 *   y_logits [B,ntiles,C,ly,lx] float32 or NULL
 *   y0/x0/flip [ntiles], taper_y [ly], taper_x [lx], (Ly, Lx), crop as in cpb_average_tiles_device; augment != 0 = --tta
 *   out: dP [B,2,H,W], cellprob [B,H,W], logits [B,C,H,W] (blended, cropped; H = Ly-cy0-cy1, W = Lx-cx0-cx1), and the
 *        outputs of cpb_compute_masks_device.
 * Geometry the fused kernel needs: lx % 4 == 0, cx0 % 4 == 0, every x0 % 4 == 0 (host's promise), W % 64 == 0;
 * otherwise CPB_E_ARG (compose cpb_average_tiles_ex_device + cpb_compute_masks_device instead). */
int cpb_eval_tail_device(const float* y_flows, const float* y_logits, int B, int ntiles, int C, int ly, int lx,
                         const int32_t* y0, const int32_t* x0, const int32_t* flip, int augment,
                         const double* taper_y, const double* taper_x, int Ly, int Lx, int cy0, int cy1,
                         int cx0, int cx1, const cpb_params* prm, float* dP, float* cellprob, float* logits,
                         int32_t* masks, int32_t* counts, int32_t* cell_class, uint8_t* class_masks,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- next row N1: PostProcessor features on the device --------------------------------------------
 * Replaces the per-cell host loop of PostProcessor.__call__ (predict_wsi.py:595-656): ndimage.find_objects,
 * cv2.findContours(cell_mask, RETR_EXTERNAL, CHAIN_APPROX_SIMPLE)[0] and the shapely polygon measures.
 * For every label l of every tile (index [b*lcap + l]):
 *   npoints   int32   points of contours[0], identical to cv2's list (0 if the label is absent)
 *   offsets   int64   start of the label's points in `points` (exclusive scan in (tile, label) order)
 *   points    int16   [points_cap][2] (x, y) in tile pixels; total[0] = points needed -- if it exceeds
 *                     points_cap the labels that do not fit are left unwritten (valid = 0), call again
 *   feat      int64   [..][8] pixel area, ymin, ymax, xmin, xmax, A2, Sx, Sy with the exact integer polygon sums:
 *                     area = |A2|/2, centroid = (Sx, Sy) / (3*A2)
 *   perimeter float64 polygon length in tile pixels
 *   valid     int32   ring has >= 4 points, non-zero area and neither touches nor crosses itself
 * Slide coordinates are an affine map on the host: p*scale + tile_origin (area*scale^2, perimeter*scale). */
int cpb_cell_contours_device(const int32_t* masks, int B, int H, int W, int lcap, int32_t* npoints,
                             int64_t* offsets, int64_t* total, int16_t* points, int64_t points_cap,
                             int64_t* feat, double* perimeter, int32_t* valid, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---- next row N2: overlap de-duplication across tiles -------------------------------------------------
 * Replaces deduplicate() (predict_wsi.py:896-965): cells whose centroids are within max_dist (7.5) of each other
 * are grouped (connected components of that relation) and only the largest cell of a group is kept (ties: lowest
 * index).  cx, cy, size: float64 [n] device arrays in slide coordinates; keep int32 [n] (1 = survives);
 * group (optional) int32 [n] = representative index of the cell's group. */
size_t cpb_dedup_workspace_bytes(int64_t n);
int cpb_dedup_cells_device(const double* cx, const double* cy, const double* size, int64_t n, double max_dist,
                           int32_t* keep, int32_t* group, void* workspace, size_t workspace_bytes, void* stream);

/* ---- next row N4: tile preparation in front of the network ----------------------------------------------
 * Replaces transforms.normalize_img with its defaults (models.py:641-666: per channel (x - p1) / (p99 - p1) with
 * numpy's linear percentiles; a constant channel is left as is, a channel with p99 - p1 <= 1e-3 becomes 0) followed
 * by np.pad and transforms.make_tiles with the parity flips (core.py:129-178).
 * img [B,H,W,C] float32 channels-last; (pad_y, pad_x) = leading pads of get_pad_yx; windows y0/x0/flip [ntiles] as in
 * cpb_average_tiles_device (coordinates in the padded image); tiles [B,ntiles,C,ly,lx] float32 out;
 * lowhigh [B,C,2] float32 out = (p_lower, p_upper - p_lower); code [B,C] int32 out (1 normalised, 2 zeroed, 0 left). */
int cpb_prepare_tiles_device(const float* img, int B, int H, int W, int C, double lower, double upper,
                             int pad_y, int pad_x, int ntiles, int ly, int lx, const int32_t* y0,
                             const int32_t* x0, const int32_t* flip, float* tiles, float* lowhigh,
                             int32_t* code, void* stream);

/* (e) global label offsets: exclusive prefix sum of per-tile instance counts.
 * offsets [B] int64 = base + sum(counts[0..b)); total [1] int64 = sum(counts).  `base` is the
 * rank's offset obtained from the cross-GPU all-gather of totals (host side). */
int cpb_label_offsets_device(const int32_t* counts, int B, int64_t base, int64_t* offsets,
                             int64_t* total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLASSPOSE_B200_H */
