/*
 * pycnn_b200.h -- C ABI of libpycnn_b200.so, the B200-native replacement for the
 * `pycnn.cuda(True)` backend of 77AXEL/PyCNN.
 *
 * The reference has exactly one C ABI today -- pycnn/lib/optimized.cpp:8-9,33-34
 * (`softmax`, `cross_entropy`, bound with ctypes at pycnn/pycnn.py:18-66) -- and its GPU
 * backend has no source of its own: it is CuPy calls inlined in pycnn/pycnn.py under
 * `if self.use_cuda:` (pycnn/pycnn.py:104-134, 229-263, 474-562).  Each entry point below
 * states which of those reference code paths it replaces.  INTEGRATION.md shows the ctypes
 * stubs a PyCNN maintainer adds; pycnn_b200/_lib.py is that binding, in-tree.
 *
 * Conventions
 *   - plain C types only; every host pointer is BORROWED for the duration of the call
 *     (caller owns it, like the reference's NumPy-owned buffers, pycnn/pycnn.py:39-44);
 *   - the opaque context owns all device memory, streams, events and the NCCL communicator;
 *   - every function returns 0 on success, non-zero on error; pycnn_b200_last_error() returns a
 *     thread-local message.  Nothing aborts or throws across the ABI;
 *   - there is NO CPU fallback: without a CUDA device pycnn_b200_create() fails;
 *   - calls on one context must be serialised by the caller (one context = one training job on
 *     one GPU; multi-GPU = one process per GPU, see pycnn_b200_comm_*).
 *   - host-side parameters cross the ABI as float64 (the reference's dtype, DTYPE=np.float64 in
 *     all four .pyx); device storage and arithmetic are float32 / TF32.
 *
 * Tensor order for `param index`: w1,b1,w2,b2,...,wo,bo (the reference's draw order,
 * pycnn/pycnn.py:303-311), i.e. index 2*l is the weight [out,in] row-major and 2*l+1 the bias
 * [out] of layer l (0-based; l == L is the output layer).
 */
#ifndef PYCNN_B200_H
#define PYCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PYCNN_B200_ABI_VERSION 2

typedef struct pycnn_b200_ctx pycnn_b200_ctx;

/* math modes (set_math_mode) */
#define PYCNN_B200_MATH_FP32_SIMT 0 /* FP32 CUDA-core GEMMs: the 1e-5 verification mode            */
#define PYCNN_B200_MATH_TF32 1      /* tcgen05 TF32 tensor-core GEMMs, FP32 accumulate in TMEM (1e-2) */
/* update rules (set_update_rule) */
#define PYCNN_B200_RULE_CPU 0  /* pycnn/modules/gradient_decent.pyx: clip 5.0, NaN skip, t frozen at 1 */
#define PYCNN_B200_RULE_CUPY 1 /* pycnn/pycnn.py:526-562: no clip, t advances                           */
/* optimizers */
#define PYCNN_B200_OPT_SGD 0
#define PYCNN_B200_OPT_ADAM 1
/* extractor implementations (set_extract_mode) */
#define PYCNN_B200_EXTRACT_SIMT 0 /* FP32 CUDA-core stencil (verification)                    */
#define PYCNN_B200_EXTRACT_TC 1   /* tcgen05 im2col-free 4x4-patch contraction (u8 fast path) */

/* ---- library / errors -------------------------------------------------------------------- */
int pycnn_b200_abi_version(void);
const char* pycnn_b200_last_error(void);
/* number of visible CUDA devices; fails (non-zero) when the CUDA runtime/driver is unusable */
int pycnn_b200_device_count(int* count);

/* ---- context (replaces pycnn/pycnn.py:104-134 _setup_backend(True) and :181-194 cuda()) ---- */
int pycnn_b200_create(int device, pycnn_b200_ctx** out);
int pycnn_b200_destroy(pycnn_b200_ctx* ctx);
int pycnn_b200_synchronize(pycnn_b200_ctx* ctx);
int pycnn_b200_set_math_mode(pycnn_b200_ctx* ctx, int mode);
int pycnn_b200_set_update_rule(pycnn_b200_ctx* ctx, int rule);
int pycnn_b200_set_extract_mode(pycnn_b200_ctx* ctx, int mode);
/* human-readable device description into buf (name, SM count, memory) */
int pycnn_b200_device_info(pycnn_b200_ctx* ctx, char* buf, size_t buflen);

/* ---- feature extractor (replaces _preprocess_image, pycnn/pycnn.py:267-287, the per-image
 *      cupyx convolve2d x3 / maximum / _max_pooling loop; and max_pooling.pyx:10-30) -------- */
/* taps: F x 3 x 3 row-major exactly as the user wrote them (the 180-degree flip of true
 * convolution is applied inside).  Replaces init(filters=), pycnn/pycnn.py:222-223. */
int pycnn_b200_set_filters(pycnn_b200_ctx* ctx, const double* taps, int num_filters);
/* features per image for an S x S input: ((S-2)/2)^2 * F  (pycnn/pycnn.py:330) */
int pycnn_b200_feature_count(pycnn_b200_ctx* ctx, int image_size, int64_t* out);
/* images: n x S x S x 3 uint8 (HWC, host).  features_out: n x D float64 (host), filter-major.
 * One fused conv+ReLU+pool launch for the whole batch. */
int pycnn_b200_extract(pycnn_b200_ctx* ctx, const uint8_t* images, int64_t n, int image_size,
                       double* features_out);
/* same, float32 output (half the D2H bytes) */
int pycnn_b200_extract_f32(pycnn_b200_ctx* ctx, const uint8_t* images, int64_t n, int image_size,
                           float* features_out);
/* 2x2/stride-2 max-pool of one float64 map (drop-in for max_pooling.pyx / pycnn.py:251-265) */
int pycnn_b200_max_pool(pycnn_b200_ctx* ctx, const double* map, int h, int w, double* out);

/* ---- the reference's existing C ABI, on the GPU (pycnn/lib/optimized.cpp:9,34) ------------- */
int pycnn_b200_softmax(pycnn_b200_ctx* ctx, const double* x, double* out, int rows, int cols);
int pycnn_b200_cross_entropy(pycnn_b200_ctx* ctx, const double* preds, const double* labels,
                             double* out, int rows, int cols);

/* ---- network (replaces DatasetLoader._init_layers storage, pycnn/pycnn.py:293-321; values are
 *      still drawn on the host from np.random so both backends start from the same weights) -- */
int pycnn_b200_set_network(pycnn_b200_ctx* ctx, int64_t input_size, const int* layers, int num_layers,
                           int num_classes);
int pycnn_b200_num_params(pycnn_b200_ctx* ctx, int* count);
int pycnn_b200_param_shape(pycnn_b200_ctx* ctx, int index, int64_t* rows, int64_t* cols);
int pycnn_b200_set_param(pycnn_b200_ctx* ctx, int index, const double* values);
int pycnn_b200_get_param(pycnn_b200_ctx* ctx, int index, double* values);
int pycnn_b200_get_grad(pycnn_b200_ctx* ctx, int index, double* values);
/* which: 0 = Adam m, 1 = Adam v (sidecar checkpoints; the reference keeps them in self.m/self.v) */
int pycnn_b200_get_opt_state(pycnn_b200_ctx* ctx, int which, int index, double* values);
int pycnn_b200_set_opt_state(pycnn_b200_ctx* ctx, int which, int index, const double* values);
/* replaces adam() / the sgd default (pycnn/pycnn.py:196-213); resets m, v and t */
int pycnn_b200_set_optimizer(pycnn_b200_ctx* ctx, int kind, double lr, double beta1, double beta2,
                             double eps);
int pycnn_b200_get_step_count(pycnn_b200_ctx* ctx, int64_t* t);
/* resume: restore the Adam step count t of the CuPy update rule (pycnn/pycnn.py:534) after set_optimizer +
 * set_opt_state; the CPU rule never advances t (pycnn/modules/gradient_decent.pyx:58-59) and ignores it */
int pycnn_b200_set_step_count(pycnn_b200_ctx* ctx, int64_t t);

/* ---- device-resident dataset (replaces self.data/self.labels migration, pycnn.py:192-194) -- */
int pycnn_b200_dataset_alloc(pycnn_b200_ctx* ctx, int64_t capacity_rows);
/* extract n images straight into feature rows [row0, row0+n) -- no host round trip */
int pycnn_b200_dataset_extract(pycnn_b200_ctx* ctx, int64_t row0, const uint8_t* images, int64_t n,
                               int image_size);
/* inject precomputed float64 features [n x D] (e.g. produced by the CPU backend) */
int pycnn_b200_dataset_set_features(pycnn_b200_ctx* ctx, int64_t row0, const double* features, int64_t n);
int pycnn_b200_dataset_get_features(pycnn_b200_ctx* ctx, int64_t row0, double* features, int64_t n);
/* class index per row (argmax of the reference's one-hot rows, pycnn/pycnn.py:353-355) */
int pycnn_b200_dataset_set_labels(pycnn_b200_ctx* ctx, int64_t row0, const int32_t* labels, int64_t n);
int pycnn_b200_dataset_rows(pycnn_b200_ctx* ctx, int64_t* rows);

/* ---- training (replaces the batch loop pycnn/pycnn.py:626-642: gather, _forward_pass :474,
 *      _cross_entropy_batch :245, argmax/acc :637-639, _backward_pass :498, _gradient_decent :525) */
/* one step on dataset rows idx[0..bs).  loss_sum/correct may be NULL (no host sync then). */
int pycnn_b200_train_step(pycnn_b200_ctx* ctx, const int64_t* idx, int bs, double* loss_sum,
                          int64_t* correct);
/* one epoch: batches perm[start:start+batch_size] in order (short last batch allowed).  With a
 * communicator of W ranks each rank takes its contiguous 1/W slice of every batch and gradients
 * are all-reduced (sum).  Outputs are the epoch totals over ALL ranks. */
int pycnn_b200_train_epoch(pycnn_b200_ctx* ctx, const int64_t* perm, int64_t n, int batch_size,
                           double* loss_sum, int64_t* correct);
/* streaming step from HOST images: H2D copy, extractor, forward, loss, backward, update.
 * images n x S x S x 3 uint8 (pinned memory recommended), labels n x int32. */
int pycnn_b200_train_step_images(pycnn_b200_ctx* ctx, const uint8_t* images, const int32_t* labels,
                                 int n, int image_size, double* loss_sum, int64_t* correct);
/* streaming epoch from HOST images (the e2e path): batches perm[start:start+batch_size] (perm NULL = rows in
 * order) are brought to the device on a copy stream -- copy engine for contiguous batches, a zero-copy gather
 * kernel over PCIe for permuted rows of pinned/registered memory -- double-buffered against extractor + step on
 * the compute stream; the running (loss, #correct) totals are read back asynchronously after every step.
 * step_stats (nullable) receives 2 doubles per step: cumulative loss sum and cumulative #correct.
 * images / labels hold num_images rows; the epoch trains n rows: perm[0..n) with every entry in
 * [0, num_images) (checked on the host before anything is launched), or rows 0..n-1 (n <= num_images). */
int pycnn_b200_train_epoch_images(pycnn_b200_ctx* ctx, const uint8_t* images, const int32_t* labels,
                                  int64_t num_images, const int64_t* perm, int64_t n, int image_size,
                                  int batch_size, double* step_stats, double* loss_sum, int64_t* correct);
/* `steps` back-to-back steps on dataset rows idx[s*bs .. (s+1)*bs), no host sync inside
 * (device-resident benchmark loop); outputs are totals over the steps. */
int pycnn_b200_train_steps(pycnn_b200_ctx* ctx, const int64_t* idx, int bs, int steps,
                           double* loss_sum, int64_t* correct);

/* ---- inference (replaces _forward_pass for Evaluate._run_inference :873-914 and the GEMV chain
 *      of predict :744-759) ------------------------------------------------------------------ */
/* probabilities [bs x C] float64 for dataset rows idx (idx NULL -> rows row0..row0+bs) */
int pycnn_b200_forward_rows(pycnn_b200_ctx* ctx, const int64_t* idx, int64_t row0, int bs, double* probs);
/* probabilities for host float64 features [bs x D] */
int pycnn_b200_forward_features(pycnn_b200_ctx* ctx, const double* features, int bs, double* probs);
/* batched predict from host uint8 images: class index + confidence per image */
int pycnn_b200_predict_images(pycnn_b200_ctx* ctx, const uint8_t* images, int64_t n, int image_size,
                              int32_t* classes, float* confidences);
/* evaluate dataset rows [row0,row0+n) in chunks of batch_size: total CE loss and #correct,
 * per-row predicted class (may be NULL) */
int pycnn_b200_evaluate_rows(pycnn_b200_ctx* ctx, int64_t row0, int64_t n, int batch_size,
                             double* loss_sum, int64_t* correct, int32_t* predicted);

/* evaluate HOST uint8 images against int32 labels: extractor + forward + CE/argmax per chunk */
int pycnn_b200_evaluate_images(pycnn_b200_ctx* ctx, const uint8_t* images, const int32_t* labels, int64_t n,
                               int image_size, int batch_size, double* loss_sum, int64_t* correct,
                               int32_t* predicted);
/* call-protocol compatibility with the reference's private methods: backward on the activations left
 * on the device by the last pycnn_b200_forward_features(bs) call (_backward_pass, pycnn/pycnn.py:498);
 * gradients are then readable with get_grad and applied with apply_update (_gradient_decent, :525) */
int pycnn_b200_backward_last(pycnn_b200_ctx* ctx, const int32_t* labels, int bs);

/* ---- data parallel: one process per GPU, NCCL over NVLink (no counterpart in the reference) -- */
/* 128-byte ncclUniqueId; rank 0 creates it, the host layer broadcasts it (torch.distributed) */
int pycnn_b200_comm_unique_id(char* id128);
int pycnn_b200_comm_init(pycnn_b200_ctx* ctx, const char* id128, int rank, int world);
int pycnn_b200_comm_destroy(pycnn_b200_ctx* ctx);
/* Peer-memory gradient exchange (pycnn_b200/csrc/p2p.cuh): replaces NCCL all-reduce + norm + update by a fused
 * reduce-scatter -> norms -> sharded clip/Adam -> all-gather over NVLink peer pointers.  Every rank exports
 * the CUDA IPC handles of its gradient, parameter and sync buffers (192 bytes, after set_network), the host
 * layer all-gathers them, and every rank attaches to all `world` blobs (rank order).  world <= 8, one node.
 * set_network / comm_destroy detach.  Adam m, v are then sharded: a rank only holds state for its range until
 * comm_sync_opt_state (a collective: every rank calls it) assembles the shards on every rank -- call it before
 * get_opt_state / a checkpoint; comm_ipc_detach does it itself. */
#define PYCNN_B200_IPC_BLOB_BYTES 192
int pycnn_b200_comm_ipc_export(pycnn_b200_ctx* ctx, char* blob192);
int pycnn_b200_comm_ipc_attach(pycnn_b200_ctx* ctx, const char* blobs, int rank, int world);
int pycnn_b200_comm_ipc_detach(pycnn_b200_ctx* ctx);
int pycnn_b200_comm_sync_opt_state(pycnn_b200_ctx* ctx);
/* NVSwitch multicast (NVLS) gradient exchange, csrc/nvls.cuh -- same place in the step as the peer-memory exchange
 * (replaces pycnn/modules/nn.py's single-device update when the batch is sharded over GPUs): the switch adds every
 * rank's gradient (multimem.ld_reduce) and broadcasts the sum (multimem.st), one launch per step; the update stays
 * replicated, so optimizer state is whole on every rank.  Collective set-up, in this order, the host layer
 * putting a barrier between the steps:
 *   supported   *supported = 1 when the device and driver offer multicast objects and POSIX-fd handles
 *   create      rank 0 only: creates the object sized for the bound network, returns a file descriptor to pass to
 *               the other ranks (SCM_RIGHTS; the caller closes it)
 *   join        every rank: fd = the received descriptor (rank 0: -1), adds this rank's device      [barrier]
 *   bind        every rank: allocates + binds + maps the region; the flat gradient buffer moves into it (pointers
 *               from grad_buffer taken earlier are stale)                                           [barrier]
 *   detach      back to an ordinary gradient buffer and the NCCL all-reduce; set_network / comm_destroy detach too. */
int pycnn_b200_comm_nvls_supported(pycnn_b200_ctx* ctx, int* supported);
int pycnn_b200_comm_nvls_create(pycnn_b200_ctx* ctx, int world, int* fd_out);
int pycnn_b200_comm_nvls_join(pycnn_b200_ctx* ctx, int fd, int rank, int world);
int pycnn_b200_comm_nvls_bind(pycnn_b200_ctx* ctx);
int pycnn_b200_comm_nvls_detach(pycnn_b200_ctx* ctx);
/* device pointer + element count of the flat float32 gradient buffer (for an external
 * all-reduce, e.g. torch.distributed on a wrapped tensor); valid until set_network / comm_nvls_bind / detach */
int pycnn_b200_grad_buffer(pycnn_b200_ctx* ctx, void** device_ptr, int64_t* num_floats);
/* split step for an external all-reduce: backward only / update only */
int pycnn_b200_forward_backward(pycnn_b200_ctx* ctx, const int64_t* idx, int bs);
int pycnn_b200_apply_update(pycnn_b200_ctx* ctx);
int pycnn_b200_read_step_stats(pycnn_b200_ctx* ctx, double* loss_sum, int64_t* correct, int reset);

/* ---- measurement ----------------------------------------------------------------------------- */
/* CUDA events on the context's compute stream */
int pycnn_b200_event_record(pycnn_b200_ctx* ctx, int slot);                       /* slot 0..15 */
int pycnn_b200_event_elapsed_ms(pycnn_b200_ctx* ctx, int slot_a, int slot_b, float* ms);
/* per-kernel-class timing (events around every launch of that class); classes are listed by
 * pycnn_b200_kernel_class_name(i), i in [0, pycnn_b200_kernel_class_count()) */
int pycnn_b200_set_kernel_timing(pycnn_b200_ctx* ctx, int enabled);
int pycnn_b200_kernel_class_count(void);
const char* pycnn_b200_kernel_class_name(int cls);
int pycnn_b200_kernel_stats(pycnn_b200_ctx* ctx, int cls, int64_t* launches, double* total_ms, int reset);
/* total kernel launches issued by this context since creation */
int pycnn_b200_launch_count(pycnn_b200_ctx* ctx, int64_t* launches);
/* write >= L2-size bytes to evict L2 between timed iterations */
/* In-kernel timeline (diagnostics; no reference counterpart): op 0 = enable with `capacity` slots (drops captured graphs
 * so that new captures carry slots), 1 = reset all slots, 2 = disable.  Every traced launch (GEMMs, row gather, norm,
 * update) records {first CTA start, last CTA end} in %globaltimer ns plus (min, max) over its CTAs of up to three
 * phase marks (GEMMs: first operand stage landed, accumulator complete, partial tile parked); read returns the
 * slots handed out so far: 8 words per slot in times[8*i .. 8*i+7] (unused marks: ~0 / 0) and
 * kinds[i] = kernel class + 100 * stream (0 main, 1 side, 2 prefetch). */
int pycnn_b200_trace_control(pycnn_b200_ctx* ctx, int op, int capacity);
int pycnn_b200_trace_read(pycnn_b200_ctx* ctx, uint64_t* times, int32_t* kinds, int capacity, int* count);
int pycnn_b200_flush_l2(pycnn_b200_ctx* ctx);
/* standalone GEMM probe for tests/roofline: C[M,N] = op(A) . op(B) with the context's math mode.
 * a_kmajor: A stored [M,K] row-major (1) or [K,M] row-major (0); b_kmajor: B stored [N,K] (1) or
 * [K,N] (0).  Host float32 in/out. */
int pycnn_b200_gemm_probe(pycnn_b200_ctx* ctx, const float* A, const float* B, float* C, int M, int N,
                          int K, int a_kmajor, int b_kmajor, int iters, float* avg_ms);

#ifdef __cplusplus
}
#endif
#endif /* PYCNN_B200_H */
