/*
 * dcrf_b200.h -- C ABI of the B200-native DenseCRF mean-field engine (libdcrf_b200.so).
 *
 * This is the drop-in boundary for the one hot path of lyndonchan/wsss-analysis: the fully-connected
 * DenseCRF that the reference reaches through the (un-vendored) `pydensecrf` Python class API.  Each
 * entry point names the reference interface it replaces.  File:line citations are into
 * /root/reference; "[EXT]" marks pydensecrf behaviour recalled from the public package, which is
 * not present in the reference tree.
 *
 * Conventions
 *   - plain C: pointers, sizes, ints and floats only; no torch / numpy types.
 *   - every function returns 0 on success, a DCRF_E* code otherwise; dcrf_last_error() gives the
 *     thread-local message of the last failure on the calling thread.
 *   - a handle holds a BATCH of B >= 1 independent images that share the label count L.  A single
 *     pydensecrf `DenseCRF2D(w, h, L)` object (03c_hsn/utilities.py:427) is a batch of one.
 *   - pixel index p = y*W + x; per-image matrices are row-major (L, N_b) float32 exactly as the
 *     Python callers hand them over (03c_hsn/utilities.py:431-432, 443); in a batch the per-image
 *     blocks are laid back to back in image order ("concatenated").
 *   - `on_device` = 0: the pointer is host memory (pageable or pinned); 1: device memory on the
 *     handle's device.  All work is enqueued on the handle's stream; calls that return host data
 *     synchronise that stream before returning, calls that only take device pointers do not.
 *   - there is NO CPU fallback: without a CUDA device every call fails with DCRF_ECUDA.
 */
#ifndef DCRF_B200_H
#define DCRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dcrf_handle dcrf_t;

/* error codes */
enum {
    DCRF_OK = 0,
    DCRF_EINVAL = 1, /* bad argument (shape, enum, NULL)   -> Python ValueError */
    DCRF_ECUDA = 2,  /* CUDA runtime failure / no device   -> Python RuntimeError */
    DCRF_ESTATE = 3, /* call sequence error (e.g. inference before setUnaryEnergy) */
    DCRF_ENOMEM = 4
};

/* [EXT] pydensecrf.densecrf enums; only DIAG_KERNEL + NORMALIZE_SYMMETRIC (the defaults) are
 * exercised by the reference (03c_hsn/utilities.py:435,439-440). */
enum { DCRF_CONST_KERNEL = 0, DCRF_DIAG_KERNEL = 1, DCRF_FULL_KERNEL = 2 };
enum {
    DCRF_NO_NORMALIZATION = 0,
    DCRF_NORMALIZE_BEFORE = 1,
    DCRF_NORMALIZE_AFTER = 2,
    DCRF_NORMALIZE_SYMMETRIC = 3
};
/* [EXT] label compatibility: a number -> Potts, a 1-D array -> diagonal, a 2-D array -> matrix */
enum { DCRF_COMPAT_POTTS = 0, DCRF_COMPAT_DIAGONAL = 1, DCRF_COMPAT_MATRIX = 2 };

const char *dcrf_last_error(void);
/* library / build identification, e.g. "dcrf_b200 0.1 sm_100a" */
const char *dcrf_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t dcrf_launch_count(void);
/* bytes this library has copied host -> device and device -> host in the calling process (payload and
 * metadata alike).  Calls that are given device pointers (on_device = 1) move only a few hundred bytes
 * of batch geometry; tests/test_gpu_wsss.py asserts it. */
void dcrf_copy_count(int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- construction --------------------------------------------------------------------------- */

/* `stream` arguments: a cudaStream_t of the caller; NULL = the calling thread's persistent library
 * stream (handles created by one thread then run in order on one stream); DCRF_STREAM_DEDICATED = a
 * stream created for, and destroyed with, this handle (for callers that keep several handles of one
 * thread in flight concurrently). */
#define DCRF_STREAM_DEDICATED ((void *)(intptr_t)-1)

/* Long-lived streams for callers that keep several handles in flight (pipeline.py) without
 * depending on PyTorch for stream objects.  The library keeps one device-memory pool per stream;
 * dcrf_stream_destroy also releases that pool. */
int dcrf_stream_create(int device, void **stream_out);
/* The per-stream pools keep freed device memory cached for the next image (release threshold =
 * max).  dcrf_trim_memory() synchronises the current device and returns all cached, unused memory
 * of every pool to the driver. */
int dcrf_trim_memory(void);
/* free / total device memory, counting the unused memory cached in the library's pools as free (the
 * batching wrappers size their handles with it) */
int dcrf_mem_info(int device, int64_t *free_bytes, int64_t *total_bytes);
int dcrf_stream_destroy(void *stream);

/* Replaces `dcrf.DenseCRF2D(w, h, nlabels)` (03c_hsn/utilities.py:427; width first).
 * device < 0 = current device. */
int dcrf_create(int w, int h, int n_labels, int device, void *stream, dcrf_t **out);

/* [EXT] `DenseCRF(nvar, nlabels)`: a model over n_vars variables without image geometry; only
 * dcrf_add_pairwise_energy() is available on it. */
int dcrf_create_nd(int n_vars, int n_labels, int device, void *stream, dcrf_t **out);

/* Batch of n_images images with per-image sizes.  Replaces the serial per-image loops of the
 * reference: `for iter_input_image in range(num_input_images)` (03c_hsn/utilities.py:424),
 * `for i in range(batch_size)` (03a_sec-dsrg/SEC.py:274, DSRG.py:327). */
int dcrf_create_batch(int n_images, const int *w, const int *h, int n_labels, int device,
                      void *stream, dcrf_t **out);

void dcrf_destroy(dcrf_t *h);

/* Options.  DCRF_OPT_EXACT_ARITHMETIC selects the float arithmetic of the per-iteration kernels:
 *   DCRF_ARITH_FMA (0): fused multiply-add, normalisation folded into the packed entry weights, one
 *     reciprocal per softmax, CUDA expf.  Differs from a sequential CPU evaluation [EXT] by float
 *     rounding only (<= 1e-5 on Q for well-conditioned models).
 *   DCRF_ARITH_REFERENCE (1): the association of the sequential CPU evaluation, operation for
 *     operation -- separately rounded multiply / add in splat and slice, normalisation applied as its
 *     own rounded product, expf as the host libm (glibc) evaluates it, softmax sum in label order,
 *     IEEE division.  Marginals are bit-identical to such an evaluation, with one exception: splat
 *     rows with more than 256 entries (flat image regions) are summed as 256 sequential terms + a
 *     fixed tree over the tail.  About 20 % slower per iteration than DCRF_ARITH_FMA.
 *   DCRF_ARITH_STRICT (2): as 1 without that exception (a lattice vertex shared by thousands of
 *     pixels is then summed by one lane group: slow on flat images).
 *   DCRF_ARITH_AUTO (3, default): DCRF_ARITH_REFERENCE for models with a narrow appearance kernel
 *     (any bilateral colour bandwidth below 8, e.g. the IRN label CRF with srgb = 5 or SEC's ADP-func
 *     setting with srgb = 4: there the mean-field update is expansive at bistable pixels and
 *     rounding differences grow from iteration to iteration), DCRF_ARITH_FMA otherwise.
 * All modes are run-to-run deterministic.  The environment variable DCRF_ARITHMETIC = auto | fma |
 * reference | strict sets the default of handles created afterwards.  dcrf_get_arithmetic returns
 * the mode the handle resolved to (0, 1 or 2). */
enum { DCRF_ARITH_FMA = 0, DCRF_ARITH_REFERENCE = 1, DCRF_ARITH_STRICT = 2, DCRF_ARITH_AUTO = 3 };
/* DCRF_OPT_ASYNC_HOST = 1: calls that read or write caller HOST buffers only enqueue their copies on
 * the handle's stream and return; the caller keeps the buffers alive and untouched until
 * dcrf_synchronize().  Lets one host thread keep two handles (two streams) in flight so that the
 * PCIe copies of one batch overlap the kernels of the other (wsss_analysis_b200/pipeline.py).
 * Host buffers should be page-locked, otherwise the copies are not asynchronous.  In this mode
 * dcrf_set_unary runs its upload and layout change on a separate per-thread stream, so the lattice
 * builds enqueued next overlap the upload; the handle's stream waits for it before the unary is read. */
/* DCRF_OPT_PERSISTENT (experimental): run dcrf_inference / dcrf_map / dcrf_run as ONE cooperative launch
 * with grid barriers between the phases instead of ~16 launches per iteration.  0 = never, 1 = whenever
 * the model allows (Potts terms, > 2 labels), -1 (default) = for handles of at most
 * DCRF_PERSISTENT_MAX_PIXELS pixels (environment; default 0 = never: on B200 the barriers cost as much
 * as the launches they replace -- measured slower).  Same arithmetic, bit-identical marginals. */
enum { DCRF_OPT_EXACT_ARITHMETIC = 1, DCRF_OPT_ASYNC_HOST = 2, DCRF_OPT_PERSISTENT = 4 };
int dcrf_set_option(dcrf_t *h, int option, int value);
int dcrf_get_arithmetic(dcrf_t *h, int *mode_out);
/* block the calling thread until everything enqueued on the handle's stream has finished */
int dcrf_synchronize(dcrf_t *h);

/* ---- model set-up (every setter copies; the caller keeps ownership of its buffers) --------- */

/* Replaces `d.setUnaryEnergy(U)` (03c_hsn/utilities.py:432).  U: concatenated (L, N_b) blocks. */
int dcrf_set_unary(dcrf_t *h, const float *U, int on_device);

/* Unary construction on the GPU, replacing the NumPy glue in front of setUnaryEnergy:
 *  - from class probabilities: `unary_from_softmax(sm, scale, clip)` [EXT pydensecrf.utils] as called
 *    at 03c_hsn/utilities.py:431.  probs: concatenated (L, N_b) blocks, float64 (is_f64) or float32;
 *    U = -log(clip(scale*p + (1-scale)/L, clip, 1)) in double, stored as float32.  scale = 1 and
 *    has_clip = 1, clip = 1e-5 are the defaults the reference uses.
 *  - from a feature map: SEC/DSRG `crf_inference(..., use_log=True)` ([EXT] lib/crf.py; call sites
 *    03a_sec-dsrg/SEC.py:275, model.py:689): feat = concatenated (H_b, W_b, L) float32 blocks,
 *    U = -log softmax_L(feat).  use_log must be non-zero: the wrapper's other branch is exercised by
 *    no call site and its source is not in the reference tree, so it is not guessed (DCRF_EINVAL).
 *  - from hard labels: `unary_from_labels(labels, L, gt_prob, zero_unsure)` [EXT] as used by
 *    crf_inference_label (03b_irn/step/cam_to_ir_label.py:35).  labels: concatenated int32. */
int dcrf_set_unary_from_probs(dcrf_t *h, const void *probs, int is_f64, double scale, double clip,
                              int has_clip, int on_device);
int dcrf_set_unary_from_logits(dcrf_t *h, const float *feat, int use_log, int on_device);
int dcrf_set_unary_from_labels(dcrf_t *h, const int32_t *labels, float gt_prob, int zero_unsure, int on_device);

/* Replaces `d.addPairwiseGaussian(sxy=(sx,sy), compat=...)` (03c_hsn/utilities.py:435).
 * compat: 1 float (Potts), L floats (diagonal) or L*L floats row-major (matrix); HOST memory. */
int dcrf_add_pairwise_gaussian(dcrf_t *h, float sx, float sy, int compat_kind, const float *compat,
                               int kernel_type, int normalization_type);

/* Replaces `d.addPairwiseBilateral(sxy, srgb, rgbim, compat)` (03c_hsn/utilities.py:439-440).
 * rgb: concatenated (H_b, W_b, 3) uint8 images; read only during this call. */
int dcrf_add_pairwise_bilateral(dcrf_t *h, float sx, float sy, float sr, float sg, float sb,
                                const uint8_t *rgb, int on_device, int compat_kind,
                                const float *compat, int kernel_type, int normalization_type);

/* [EXT] `addPairwiseEnergy(features, compat, kernel, normalization)`: features row-major (d, N),
 * 1 <= d <= 7.  Single-image / nd handles only. */
int dcrf_add_pairwise_energy(dcrf_t *h, const float *features, int d, int on_device, int compat_kind,
                             const float *compat, int kernel_type, int normalization_type);

/* ---- inference -------------------------------------------------------------------------------- */

/* Replaces `Q = d.inference(n)` + `np.array(Q)` (03c_hsn/utilities.py:442-443): runs n mean-field
 * iterations from the unary and writes the marginals as concatenated (L, N_b) float32 blocks. */
int dcrf_inference(dcrf_t *h, int n_iter, float *Q_out, int on_device);

/* inference + per-pixel argmax over labels (first maximum wins, like np.argmax):
 * replaces `np.argmax(np.array(Q).reshape(L,H,W), axis=0)` (03c_hsn/utilities.py:443-444,
 * [EXT] crf_inference_label used by 03b_irn/step/cam_to_ir_label.py:35).  labels: sum(N_b) int32. */
int dcrf_map(dcrf_t *h, int n_iter, int32_t *labels_out, int on_device);
/* the same with uint8 labels (n_labels <= 256): what a label consumer downloads -- 1 byte per pixel
 * over PCIe instead of 4 * n_labels bytes of marginals (dcrf_process, crf_inference_label, the
 * evaluation loops of 03a_sec-dsrg/model.py:699 and 03b_irn/step/eval_sem_seg.py only keep argmax) */
int dcrf_map_u8(dcrf_t *h, int n_iter, uint8_t *labels_out, int on_device);

/* The two halves of dcrf_inference / dcrf_map, for callers that overlap the download of one batch
 * with the iterations of the next (wsss_analysis_b200/pipeline.py): dcrf_run = startInference +
 * n_iter steps, blocks until the iterations have finished; dcrf_get_q / dcrf_get_labels then emit the
 * running Q (or its argmax). */
int dcrf_run(dcrf_t *h, int n_iter);
int dcrf_get_labels(dcrf_t *h, int32_t *labels_out, int on_device);
int dcrf_get_labels_u8(dcrf_t *h, uint8_t *labels_out, int on_device);

/* [EXT] startInference / stepInference / klDivergence.  The running Q lives inside the handle. */
int dcrf_start_inference(dcrf_t *h);
int dcrf_step_inference(dcrf_t *h);
int dcrf_get_q(dcrf_t *h, float *Q_out, int on_device);
int dcrf_set_q(dcrf_t *h, const float *Q_in, int on_device);

/* The running Q as concatenated (H_b, W_b, L) float32 blocks -- the layout [EXT] lib/crf.py's
 * crf_inference returns to 03a_sec-dsrg/SEC.py:275 and model.py:689-693 -- without a host transpose.
 * min_prob > 0 additionally applies the epilogue of the `crf` py_func closure (SEC.py:277-278,
 * DSRG.py:330-331): ret[ret < min_prob] = min_prob; ret /= sum over labels (NumPy's float32
 * summation order, bit-identical); take_log != 0 then takes the logarithm (SEC.py:279). */
int dcrf_get_q_hwc(dcrf_t *h, float min_prob, int take_log, float *out, int on_device);
int dcrf_kl_divergence(dcrf_t *h, double *kl_out); /* of the running Q; batch-of-one only */

/* ---- introspection (bit-exact lattice tests) ------------------------- */

int dcrf_num_pairwise(dcrf_t *h, int *n_out);
/* dimension d, total vertex count M over the batch, and (optional) per-image vertex counts */
int dcrf_lattice_info(dcrf_t *h, int kernel, int *d_out, int64_t *M_out, int64_t *M_per_image);
/* Host buffers (any may be NULL), all in the reference numbering of Appendix A.3 (vertex id = rank
 * of the key's first occurrence in pixel-major / remainder-minor scan order), for image `image`:
 *   keys (M_b, d) int16; offsets (N_b, d+1) int32; bary (N_b, d+1) float32;
 *   neighbours (d+1, M_b, 2) int32 with -1 = absent; norm (N_b) float32. */
int dcrf_lattice_export(dcrf_t *h, int kernel, int image, int16_t *keys, int32_t *offsets,
                        float *bary, int32_t *neighbours, float *norm);
/* y[i] = expf(x[i]) as the reference-arithmetic softmax evaluates it (restatement of glibc's
 * double-precision expf for x <= 0); host buffers.  Test hook. */
int dcrf_expf_ref(const float *x, float *y, int64_t n, int device);
/* one application of pairwise kernel `kernel`'s lattice filter (splat, blur, slice; no norm, no
 * compat) to host values (L, N) -> (L, N); batch-of-one only.  Test hook. */
int dcrf_lattice_filter(dcrf_t *h, int kernel, const float *in, float *out, int value_size);

/* ---- measurement hooks (bench.py's roofline object) ------------------------------------------ */

/* Kernel classes timed with CUDA events recorded on the handle's stream around every launch. */
enum {
    DCRF_K_SPLAT = 0, DCRF_K_BLUR = 1, DCRF_K_SLICE = 2,
    /* lattice construction phases (tag = d): point = elevate / round / rank / barycentric; hash = key
     * insertion; number = first-occurrence flags + scan + offsets; neigh = compact table + neighbour
     * look-ups; sort = radix sort of the entries by vertex; csr = CSR rows + packed entry tables;
     * norm = kernel normalisation (filter of all-ones); repl = replication of a position-only lattice */
    DCRF_K_BUILD_POINT = 3, DCRF_K_BUILD_HASH = 4, DCRF_K_BUILD_NUMBER = 5, DCRF_K_BUILD_NEIGH = 6,
    DCRF_K_BUILD_SORT = 7, DCRF_K_BUILD_CSR = 8, DCRF_K_BUILD_NORM = 9, DCRF_K_BUILD_REPL = 10
};
int dcrf_profile_enable(dcrf_t *h, int enable);
/* Synchronises the stream, then sums the recorded launches of `kernel_class` whose tag equals `tag`
 * (tag = lattice dimension d for splat / blur, number of fused pairwise terms for slice; -1 = any).
 * reset != 0 clears the records of every class afterwards. */
int dcrf_profile_read(dcrf_t *h, int kernel_class, int tag, double *total_ms, int64_t *launches,
                      int reset);

/* ---- evaluation reduction (the integer collective behind mIoU) --------------------------------- */

/* Replaces chainercv `calc_semantic_segmentation_confusion` as called at
 * 03b_irn/step/eval_sem_seg.py:41 and the per-class loops of 03a_sec-dsrg/model.py:698-719:
 * conf is (C+1, C) int64 DEVICE memory, row = GT class, column = predicted class, row C collects
 * pixels whose GT is outside [0, C) (ignored).  Accumulates (does not clear).  gt/pred: device int32.
 * Predictions outside [0, C) are counted into *n_bad_pred (device int64, may be NULL). */
int dcrf_confusion_accumulate(const int32_t *gt, const int32_t *pred, int64_t n, int n_classes,
                              int64_t *conf, int64_t *n_bad_pred, int device, void *stream);

/* The one collective of the path: SUM all-reduce of the int64 confusion matrix over the ranks of a
 * communicator (NCCL over NVLink / NVSwitch), in place, on `stream`.  Replaces the single-process
 * accumulation of 03b_irn/step/eval_sem_seg.py:41-50 when the image list is striped over GPUs the way
 * 03b_irn/step/cam_to_ir_label.py:114-117 stripes it over worker processes.  Integer addition is order
 * independent: the result is bit-identical for every rank count.  `comm` is an ncclComm_t -- one the
 * caller already has, or one made by dcrf_nccl_comm_create from a 128-byte ncclUniqueId that rank 0
 * obtained with dcrf_nccl_unique_id and handed to the other ranks by any means (a file, MPI,
 * torch.distributed).  NCCL is bound at run time (the libnccl.so.2 PyTorch ships, or DCRF_NCCL_LIB). */
int dcrf_nccl_unique_id(void *id_out_128_bytes);
int dcrf_nccl_comm_create(int n_ranks, int rank, const void *unique_id_128_bytes, int device, void **comm_out);
int dcrf_nccl_comm_destroy(void *comm);
int dcrf_confusion_allreduce(void *comm, int64_t *conf, int64_t count, int device, void *stream);

/* ---- resizes either side of the path (device pointers) ---------------------------------------- */

/* cv2.resize(labels, (dw, dh), interpolation=cv2.INTER_NEAREST) on an int32 label map:
 * 03b_irn/step/eval_sem_seg.py:36, 03c_hsn/demo.py:181-183.  Bit-exact index arithmetic. */
int dcrf_resize_nearest_i32(const int32_t *src, int sh, int sw, int32_t *dst, int dh, int dw, int device,
                            void *stream);
/* cv2.resize(featmap, (dw, dh)) (INTER_LINEAR) on a float32 (sh, sw, channels) map:
 * 03a_sec-dsrg/model.py:686-687,696. */
int dcrf_resize_bilinear_f32(const float *src, int sh, int sw, int channels, float *dst, int dh, int dw,
                             int device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DCRF_B200_H */
