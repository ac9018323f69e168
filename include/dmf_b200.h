/*
 * dmf_b200.h — C-ABI of libdmf_b200.so, the B200-native (sm_100a) implementation of the per-pixel
 * MS+PAN scene-classification hot path of salalalala23/Dual-modal-fusion.
 *
 * The reference is pure Python and has no FFI layer; its seams are Python-level:
 * the model plug-in (solver/mainsolver.py:30-34), the Dataset/DataLoader objects
 * (solver/basesolver.py:58,63-105), the data-prep functions (function/function.py:99-169), the IHS
 * transforms (image_convert/IHS.py:6-54) and the metric loop (solver/mainsolver.py:139-141).  Each
 * entry point below names the reference interface it replaces.  The Python host mirror
 * (dual-modal-fusion_b200/{solver,train,function,image_convert,indicators,model}) binds these
 * symbols with ctypes (dual-modal-fusion_b200/dmf/_lib.py); INTEGRATION.md shows the binding a
 * reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative dmf_status; dmf_last_error() returns a
 *    thread-local message for the last failure.  Nothing aborts.
 *  - pointers named *_dev are device pointers (e.g. torch.Tensor.data_ptr()); *_host are host
 *    pointers.  The library never frees caller memory; opaque handles own their device buffers
 *    until the matching *_destroy.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work is
 *    asynchronous on that stream unless stated otherwise.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    DMF_ERR_CUDA.
 */
#ifndef DMF_B200_H
#define DMF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMF_ABI_VERSION 1

typedef enum dmf_status {
    DMF_OK = 0,
    DMF_ERR_ARG = -1,      /* bad argument                         */
    DMF_ERR_CUDA = -2,     /* CUDA runtime / driver error          */
    DMF_ERR_STATE = -3,    /* handle not ready (e.g. weights missing) */
    DMF_ERR_UNSUPPORTED = -4
} dmf_status;

/* raster element types accepted for raw scenes (function/function.py:34-43 returns whatever the
 * TIFF holds; uint16 for these sensors) */
typedef enum dmf_dtype {
    DMF_U8 = 0,
    DMF_U16 = 1,
    DMF_F32 = 2,
    DMF_F64 = 3
} dmf_dtype;

typedef struct dmf_scene dmf_scene;   /* normalised + reflect-padded MS/PAN (and optional MSPAN) on device */
typedef struct dmf_net dmf_net;       /* GMFNet weights packed for the sm_100a kernels + activation workspace */

int dmf_abi_version(void);
const char* dmf_last_error(void);
/* number of kernels this library launched in the calling process since load (bench.py gpu_launches) */
int64_t dmf_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Scene preparation — replaces to_tensor + data_padding (function/function.py:99-124):
 * global min-max normalisation over the whole raster, then BORDER_REFLECT_101 padding of the
 * bottom/right edge by p-1 (MS) / 4p-1 (PAN).  Results are bit-exact with
 * float32(reference float64 value).
 * ------------------------------------------------------------------------------------------ */

/* Generic normalise+pad of one raster: raw[H][W][bands] -> out[H+P-1][W+P-1][bands].
 * out_dtype is DMF_F32 or DMF_F64 (DMF_F64 reproduces data_padding()'s return value exactly). */
int dmf_normalize_pad(const void* raw_dev, int raw_dtype, int H, int W, int bands, int P,
                      void* out_dev, int out_dtype, void* stream);

/* Build a device scene from RAW rasters ms[H][W][4], pan[4H][4W]. `on_device` = 0 if the two
 * pointers are host memory (copied with cudaMemcpyAsync), 1 if already on the device. */
int dmf_scene_create_raw(dmf_scene** out, const void* ms, int ms_dtype, const void* pan, int pan_dtype,
                         int H, int W, int p, int on_device, void* stream);
/* Re-fill an existing scene with new rasters of the same H, W, p (no allocation: same-size scenes in a
 * loop, e.g. tiles of a mosaic). */
int dmf_scene_update_raw(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype,
                         int on_device, void* stream);
/* Row-band scenes (one band of a larger scene per GPU).  to_tensor (function/function.py:120-124) normalises with the min / max of
 * the WHOLE raster: dmf_raster_minmax gives a band's min / max on the device (lohi_out_dev = 2 doubles) for the ranks to all-reduce,
 * and a re-fill that normalises with GIVEN ranges (device {min, max} pairs, e.g. the all-reduced ranges of all bands) instead of
 * the band's own.  A band made of scene rows [s0, s1) with s1 = min(H, r1 + p - 1) reproduces the whole scene's windows for the
 * anchors of rows [r0, r1): interior bands carry the next band's first p-1 rows, the last band's reflect padding is its own. */
int dmf_raster_minmax(const void* raw_dev, int dtype, int64_t n, double* lohi_out_dev, void* stream);
int dmf_scene_update_raw_range(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                               const double* ms_lohi_dev, const double* pan_lohi_dev, void* stream);
/* Build a device scene from ALREADY normalised+padded rasters as data_padding() returns them:
 * ms_pad[H+p-1][W+p-1][4], pan_pad[4H+4p-1][4W+4p-1]; dtype DMF_F32 or DMF_F64 (cast to f32 with
 * round-to-nearest, the cast dataset_dual applies per patch, train/dataset.py:183-184). */
int dmf_scene_create_padded(dmf_scene** out, const void* ms_pad, const void* pan_pad, int dtype,
                            int H, int W, int p, int on_device, void* stream);
/* Attach the third raster of dataset_tri (train/dataset.py:249-268), same shape as pan_pad. */
int dmf_scene_set_mspan(dmf_scene* s, const void* mspan_pad, int dtype, int on_device, void* stream);
/* The same raster computed ON THE DEVICE from the raw rasters of the scene — replaces IHS_tran(to_tensor(ms), to_tensor(pan))
 * (image_convert/IHS.py:40-54 on function/function.py:120-124) + the reflect-101 padding of data_padding (function/function.py:104-110)
 * + the dataset's float32 cast (train/dataset.py:265-268) in one pass: float64 arithmetic in the reference's operation order, so the
 * result is float32(reference float64 value) bit for bit.  ms[H][W][4], pan[4H][4W] as for dmf_scene_create_raw; offsets_dev =
 * int8[4][H][W][2], the (m, n) draws of unpooling() (image_convert/IHS.py:25-28) in band -> row -> col order; ms_lohi_dev /
 * pan_lohi_dev: device {min, max} pairs to normalise with (row-band scenes), or both NULL = the rasters' own ranges. */
int dmf_scene_set_mspan_ihs(dmf_scene* s, const void* ms, int ms_dtype, const void* pan, int pan_dtype, int on_device,
                            const int8_t* offsets_dev, const double* ms_lohi_dev, const double* pan_lohi_dev, void* stream);
/* Attach the label map uint8[H][W] (label.npy, solver/basesolver.py:35-37). */
int dmf_scene_set_labels(dmf_scene* s, const uint8_t* label, int on_device, void* stream);
int dmf_scene_destroy(dmf_scene* s);
/* dims[0..5] = H, W, p, bands, padded MS rows, padded MS cols */
int dmf_scene_dims(const dmf_scene* s, int32_t dims[6]);
/* copy the padded device rasters out: which = 0 MS [Hp][Wp][4], 1 PAN [H4p][W4p], 2 MSPAN; f32. */
int dmf_scene_export(const dmf_scene* s, int which, float* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1 patch gather — replaces dataset_dual.__getitem__ / dataset_tri.__getitem__ + default_collate
 * + .to(device) (train/dataset.py:168-185, 259-279; solver/mainsolver.py:50).
 * flat_idx[i] = row*W + col of the pixel whose patch has its TOP-LEFT corner there.
 * Outputs: ms [N][4][p][p] f32, pan [N][1][4p][4p] f32, mspan like pan (or NULL),
 * target [N] f32 = label at the pixel (NULL to skip; needs dmf_scene_set_labels).
 * ------------------------------------------------------------------------------------------ */
int dmf_gather(const dmf_scene* s, const int64_t* flat_idx_dev, int64_t N,
               float* ms_out_dev, float* pan_out_dev, float* mspan_out_dev, float* target_out_dev,
               void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2 IHS — replaces IHS_tran / pan2ms (image_convert/IHS.py:14-19, 40-54), float64, same operation
 * order.  offsets = int8[4][H][W][2]: the (m, n) draws of unpooling() in band->row->col order.
 * ------------------------------------------------------------------------------------------ */
int dmf_ihs_tran(const double* ms_dev /*[H][W][4]*/, const double* pan_dev /*[4H][4W]*/,
                 const int8_t* offsets_dev, double* mspan_out_dev /*[4H][4W]*/, int H, int W, void* stream);
/* pan[H4][W4] (any dmf_dtype) -> out f64 [H4/4][W4/4][4] */
int dmf_pan2ms(const void* pan_dev, int pan_dtype, int H4, int W4, double* out_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3 network — replaces model.gmfnet.Net (absent from the reference; contract at
 * solver/mainsolver.py:30-38,52,109,169).  Parameters are fed by state_dict name.
 * ------------------------------------------------------------------------------------------ */
int dmf_net_create(dmf_net** out, int p, int num_classes, int max_batch);
int dmf_net_destroy(dmf_net* n);
/* host fp32 tensor of `numel` elements, contiguous in PyTorch's layout; name = state_dict key */
int dmf_net_load_param(dmf_net* n, const char* name, const float* data_host, int64_t numel);
/* fold eval-mode BatchNorm, pack weights for the kernels, upload (synchronous) */
int dmf_net_finalize(dmf_net* n, void* stream);
/* algorithmic FLOPs of one patch forward (2*MAC, no padding credit) */
int64_t dmf_net_flops_per_patch(const dmf_net* n);

/* forward on materialised patches (what Net.forward(ms, pan) receives from a loader):
 * ms [N][4][p][p] f32, pan [N][1][4p][4p] f32 -> logits [N][C] f32 */
int dmf_net_forward_patches(dmf_net* n, const float* ms_dev, const float* pan_dev, int64_t N,
                            float* logits_out_dev, void* stream);
/* fused gather + forward straight from the scene for an index list (flat_idx_dev) or, when it is
 * NULL, for the N consecutive pixels starting at flat index `first`.  Any of the outputs may be
 * NULL: logits [N][C] f32, pred [N] u8, cm int64[C][C] (accumulated: cm[pred][label] += 1, needs
 * labels), pred_map u8[H][W] (written at each pixel's own position). */
int dmf_net_forward_scene(dmf_net* n, const dmf_scene* s, const int64_t* flat_idx_dev, int64_t first,
                          int64_t N, float* logits_out_dev, uint8_t* pred_out_dev, int64_t* cm_dev,
                          uint8_t* pred_map_dev, void* stream);
/* whole row band [row0,row1) of the scene: Solver.color()'s two loader passes + the confusion loop
 * (solver/mainsolver.py:167-185, train/test.py:58-60) in one call. */
int dmf_infer_scene(dmf_net* n, const dmf_scene* s, int row0, int row1, uint8_t* pred_map_dev,
                    int64_t* cm_dev, void* stream);
/* Scene-dense evaluation of the same band (csrc/dense.cu) — replaces the two whole-scene loader passes of Solver.color()
 * (solver/mainsolver.py:167-185) and the full-loader Solver.test() (:104-141): whole-scene inference visits patches at stride 1, so
 * every layer is computed once per scene position and border class instead of once per patch; same results up to
 * fp32 summation order.  logits_out_dev: [(row1-row0)*W][C] f32 or NULL.  dmf_infer_scene uses this path unless
 * dmf_net_set_dense(n, 0, 0) selects the per-patch kernels; band_rows = anchor rows per pass (workspace ~ 11 KB per
 * map position; 0 keeps the current value, default 512). */
int dmf_infer_scene_dense(dmf_net* n, const dmf_scene* s, int row0, int row1, float* logits_out_dev,
                          uint8_t* pred_map_dev, int64_t* cm_dev, void* stream);
int dmf_net_set_dense(dmf_net* n, int enabled, int band_rows);
/* accumulated device time per stage of the dense path (needs dmf_net_set_timing(n,1)): out[0..10] = ms stem maps,
 * ms2 conv+pool, (unused), pan stem maps, pan2 conv+pool, (unused), pan3 conv+pool, (unused), fuse conv + row sums, head, -; out[11] = total */
int dmf_net_get_dense_timing(dmf_net* n, float out_ms[12], int reset);
/* test hook (host only): the per-class window / box table of the fused conv + pool kernel (csrc/dense_tc.cuh), see dense.cu */
int dmf_dense_class_table(int a, int b, int aligned, int16_t* win, int16_t* box_plane, int8_t* box_drow, int8_t* box_dcol,
                          int32_t* n_boxes, int32_t* slot_bytes);
/* test hook: device pointer of a dense-path map ("A","CAT","B1","B2" bf16, "S" fp32); dims = rows, cols of the MS grid */
int dmf_net_dense_buffer(dmf_net* n, const char* name, void** ptr_out, int64_t* bytes_out, int32_t dims[2]);
/* IHS-input models (dataset_tri's third raster, train/dataset.py:259-279; trained with dmf_train_step_scene(..., use_mspan = 1)):
 * scene inference (dmf_net_forward_scene,
 * dmf_infer_scene, dmf_infer_scene_dense) reads the scene's IHS product (dmf_scene_set_mspan) in place of the PAN raster. */
int dmf_net_set_pan_source(dmf_net* n, int use_mspan);
/* device time of each stage of the last forward call, in ms (synchronises): out[0..7] = stem_ms,
 * conv_ms2, stem_pan, conv_pan2, conv_pan3, conv_fuse, head, total; needs dmf_net_set_timing(n,1). */
int dmf_net_set_timing(dmf_net* n, int enabled);
int dmf_net_get_timing(dmf_net* n, float out_ms[8]);

/* test hooks: run ONE layer on caller buffers in the kernels' activation layout
 * [N][C/8][H][W][8] bf16.  layer: 0 ms2, 1 pan2, 2 pan3, 3 fuse.  impl: 0 = tcgen05 path (2,3,4 = same with TMA loads / epilogue / both skipped: timing diagnostics),
 * 1 = CUDA-core direct convolution (debug oracle on device, never used by the product path). */
int dmf_net_debug_layer(dmf_net* n, int layer, int impl, const void* in_dev, void* out_dev, int64_t N,
                        void* stream);
int dmf_net_debug_stem(dmf_net* n, int which /*0 ms, 1 pan*/, const float* patches_dev, void* out_dev,
                       int64_t N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step — replaces the inner loop of Solver.train() (solver/mainsolver.py:49-55): forward in train mode
 * (BatchNorm batch statistics + running-stat update), CrossEntropyLoss(mean) (utils/utils.py:28-29), backward, and
 * torch.optim.Adam (utils/utils.py:12).  bf16 tensor-core arithmetic with fp32 accumulation; the parameters, their
 * gradients and the BatchNorm buffers stay the caller's fp32 device tensors (the nn.Module's own storage), bound
 * once by state_dict name.
 * ------------------------------------------------------------------------------------------ */
typedef struct dmf_train dmf_train;
int dmf_train_create(dmf_train** out, int p, int num_classes, int max_batch);
int dmf_train_destroy(dmf_train* t);
/* name = state_dict key.  param_dev: fp32 device tensor (int64 for "*.num_batches_tracked"); grad_dev: fp32 gradient
 * buffer of the same shape (NULL for BatchNorm buffers).  Gradients are ACCUMULATED by dmf_train_backward. */
int dmf_train_bind(dmf_train* t, const char* name, void* param_dev, float* grad_dev, int64_t numel);
int dmf_train_finalize(dmf_train* t);
/* train-mode forward of a batch of N <= max_batch patches: ms [N][4][p][p], pan [N][1][4p][4p] fp32 -> logits [N][C];
 * activations are kept inside the handle for the backward call. */
int dmf_train_forward(dmf_train* t, const float* ms_dev, const float* pan_dev, int64_t N, float* logits_out_dev, void* stream);
/* dlogits [N][C] fp32 = dLoss/dlogits of the last forward -> parameter gradients (accumulated) */
int dmf_train_backward(dmf_train* t, const float* dlogits_dev, void* stream);
/* CrossEntropyLoss(reduction='mean'): loss (1 float, overwritten; may be NULL) and dLoss/dlogits (may be NULL).
 * target: float32 labels as the loaders deliver them, or int64 (target_is_i64) */
int dmf_softmax_ce(const float* logits_dev, const void* target_dev, int target_is_i64, int64_t N, int C, float* loss_out_dev,
                   float* dlogits_out_dev, void* stream);
/* torch.optim.Adam without weight decay / amsgrad over one flat tensor; step = 1 for the first update */
int dmf_adam_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, int64_t numel, float lr,
                  float beta1, float beta2, float eps, int64_t step, void* stream);
/* forward + loss + backward in one call, from materialised patches ... */
int dmf_train_step_patches(dmf_train* t, const float* ms_dev, const float* pan_dev, const void* target_dev, int target_is_i64, int64_t N,
                           float* loss_out_dev, void* stream);
/* ... or cropped from the scene by flat pixel index (K1 gather in front; targets = scene labels).  use_mspan: feed the
 * IHS product's window (dataset_tri's third tensor, train/dataset.py:259-279) as the PAN input. */
int dmf_train_step_scene(dmf_train* t, const dmf_scene* s, const int64_t* flat_idx_dev, int64_t N, int use_mspan, float* loss_out_dev,
                         void* stream);
/* test hooks */
int dmf_train_buffer(dmf_train* t, const char* name, void** ptr_out, int64_t* bytes_out);
int dmf_train_set_debug(dmf_train* t, int swap_lbo_sbo);
/* run one stage on the internal buffers: op 0 pack weights, 1 forward conv, 2 wgrad, 3 dgrad; layer 0..5 = ms1, ms2, pan1, pan2,
 * pan3, fuse */
int dmf_train_debug_op(dmf_train* t, int op, int layer, int64_t N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4 argmax + confusion — replaces the loop at solver/mainsolver.py:139-141 (train/test.py:58-60):
 * pred = first index of the row maximum; cm[pred][target] += 1 (int64, caller zeroes it).
 * target_dtype: DMF_F32 (the loaders' float labels) or DMF_U8.  pred_out (int64[N]) may be NULL.
 * ------------------------------------------------------------------------------------------ */
int dmf_argmax_confusion(const float* logits_dev, const void* target_dev, int target_dtype, int64_t N,
                         int C, int64_t* pred_out_dev, int64_t* cm_dev, void* stream);

/* The same matrix (solver/mainsolver.py:139-141) from whole-scene maps: cm[pred_map[k]][label_map[k]] += 1 for k in flat_idx (NULL =
 * the first N pixels).
 * Lets Solver.test() take its loader's sample set out of one scene-dense pass instead of running the network per sample. */
int dmf_confusion_at(const uint8_t* pred_map_dev, const uint8_t* label_map_dev, const int64_t* flat_idx_dev, int64_t N,
                     int C, int64_t* cm_dev, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K5 colouring — replaces the scatter + paint loops of Solver.color()
 * (solver/mainsolver.py:171-173, 186-189).
 * ------------------------------------------------------------------------------------------ */
int dmf_scatter_labels(const int64_t* x_dev, const int64_t* y_dev, const int64_t* pred_dev, int64_t N,
                       uint8_t* label_map_dev, int W, void* stream);
int dmf_paint_labels(const uint8_t* label_map_dev, int64_t npix, const uint8_t* palette_host /*[C][3]*/,
                     int C, uint8_t* rgb_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMF_B200_H */
