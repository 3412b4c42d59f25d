// common.cuh -- context, error plumbing and launch helpers for libpycnn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <nvtx3/nvToolsExt.h>   // header-only; a no-op unless a profiler injects itself

#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/pycnn_b200.h"

namespace pb {

// ---------------------------------------------------------------------------------------------
// errors: thread-local message, int return codes, nothing throws across the ABI
// ---------------------------------------------------------------------------------------------
inline thread_local char g_err[1024] = "";

inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define PB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return pb::fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define PB_CHECK(cond, ...)                  \
  do {                                       \
    if (!(cond)) return pb::fail(__VA_ARGS__); \
  } while (0)

#define PB_TRY(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---------------------------------------------------------------------------------------------
// kernel classes for per-class event timing (pycnn_b200_kernel_stats)
// ---------------------------------------------------------------------------------------------
enum KernelClass : int {
  KC_EXTRACT = 0,   // fused conv3x3 + ReLU + 2x2 max-pool + flatten
  KC_GATHER,        // batch row gather
  KC_GEMM_FWD,      // forward GEMMs of layers 2..L+1 (bias/ReLU epilogue)
  KC_GEMM_DA,       // backward dA GEMMs (ReLU-mask epilogue, fused bias gradient)
  KC_GEMM_DW,       // backward dW GEMMs of layers 2..L+1
  KC_GEMM_FWD1,     // forward GEMM of layer 1: [bs, D] x [D, n1] -- the dominant forward kernel
  KC_GEMM_DW1,      // dW of layer 1: [n1, bs] x [bs, D] -- the dominant backward kernel
  KC_SOFTMAX_CE,    // softmax + CE + argmax + dZ
  KC_COLSUM,        // bias gradients
  KC_SQNORM,        // per-tensor gradient norms
  KC_UPDATE,        // clipped SGD / Adam
  KC_ALLREDUCE,     // NCCL gradient all-reduce (time on the compute stream)
  KC_MISC,          // converts, memsets
  KC_SPLITK,        // ordered reduce of split-K partial tiles with the GEMM's epilogue (bias / ReLU / mask)
  KC_COUNT
};

static const char* const kKernelClassNames[KC_COUNT] = {
    "extract_conv_relu_pool", "gather_rows", "gemm_forward_tail", "gemm_dA", "gemm_dW_tail", "gemm_forward_layer1",
    "gemm_dW_layer1", "softmax_ce_grad",
    "bias_colsum", "grad_finalize", "clip_update", "allreduce", "misc", "splitk_reduce"};

struct TimedLaunch {
  cudaEvent_t a, b;
};

// flat parameter layout entry
struct TensorDesc {
  int64_t off;   // offset in floats into the flat buffers
  int64_t rows;  // out
  int64_t cols;  // in (1 for biases: stored as a single row of `rows` values -> rows=1, cols=out)
  int64_t ld;    // row pitch in floats (multiple of 4)
};

struct Chunk {  // unit of work for the finalize / norm / update kernels: <= 2048 floats of one tensor
  int tensor;
  int first_chunk;   // index of the tensor's first chunk and how many chunks it has (per-tensor norm = ordered sum of
  int n_chunks;      // the chunk norms)
  int pad;
  int64_t off;       // offset in the flat buffers
  int64_t len;
  int64_t rel;       // offset inside the tensor (off - tensor.off)
};

struct GradPartSlot {   // one gradient tensor's region of the ordered-partial-sum arena (elementwise.cuh: GradParts)
  int64_t off = 0;      // floats from d_gpart
  int64_t stride = 0;   // distance between consecutive partials (the tensor's padded length)
  int cap = 0;          // partials the region can hold
  int nparts = 0;       // partials the current step's producer wrote (0: it wrote the flat gradient buffer directly)
};

struct NcclApi;  // comm.cuh

}  // namespace pb

// The opaque context (global namespace: it is the C ABI's handle type).
struct pycnn_b200_ctx {
  int device = 0;
  int num_sms = 148;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;       // compute
  cudaStream_t copy_stream = nullptr;  // H2D staging for streamed batches
  cudaStream_t side_stream = nullptr;  // backward: the shallow dW GEMMs run beside the dA chain (fork/join inside the graph)
  cudaEvent_t ev_fork[8] = {}, ev_join = nullptr;
  cudaStream_t pf_stream = nullptr;    // gather of the NEXT step's rows, overlapped with this step's backward pass
  cudaEvent_t ev_pf_fork = nullptr, ev_pf_join = nullptr;
  cudaEvent_t ev_idx_free = nullptr, ev_idx_ready = nullptr;   // train_epoch: the epoch's index tail goes up on the copy stream
  bool epoch_log = false;                                      // PYCNN_B200_EPOCH_LOG=1: host phases of train_epoch to stderr
  struct Prefetch {                    // armed by graphed_step_rows_pf, consumed by step_on_batch
    bool armed = false, forked = false;
    int next_bs = 0;
    int64_t extra = 0;                 // rows of the current step: the next batch starts at cursor + extra
    float* x = nullptr;
    int32_t* labels = nullptr;
  } pf;
  // in-kernel timeline (PYCNN_B200_TRACE=1, diagnostics): every traced launch owns a slot {first CTA start, last CTA
  // end} in %globaltimer ns; slots are handed out in launch order while a step is issued or captured
  bool trace = false;
  unsigned long long* d_trace = nullptr;
  int trace_cap = 0, trace_next = 0;
  std::vector<int> trace_cls;          // kernel class of each slot, + 100 x stream id (0 main, 1 side, 2 prefetch)
  int force_bn = 0;                    // PYCNN_B200_FORCE_BN: tile width override for deep-K GEMMs (experiments)
  int last_gemm_ctas = 0;              // grid size of the most recent tensor-core GEMM launch
  bool defer_last_dw = true;           // backward: dW of layer 2 runs on the SMs dW of layer 1 leaves idle (PYCNN_B200_DEFER_DW=0)
  bool use_prefetch = true;
  int pf_place = 0;                    // row-gather fork point: 0 auto (1 GPU: behind layer-1 forward; DP: beside the exchange), 1 fwd, 2 dw, 3 xchg (PYCNN_B200_PF_PLACE)
  bool pf_early = false;               // fork the next step's row gather BEFORE the layer-1 forward GEMM instead of behind it (PYCNN_B200_PF_EARLY=1)
  int graph_steps = 16;                // max training steps per captured graph: a graph launch costs ~15 us of idle GPU (PYCNN_B200_GRAPH_STEPS)
  int pf_ctas_per_sm = 2;              // gather CTAs per SM on the prefetch stream (PYCNN_B200_PF_CTAS)
  bool uniform_carveout = true;        // elementwise kernels ask for the GEMMs' L1/shared split (PYCNN_B200_CARVEOUT=0 disables)
  bool fuse_softmax = true;            // training: softmax/CE/dZ in the output-layer GEMM's epilogue (PYCNN_B200_FUSE_SOFTMAX=0: own kernel)
  bool one_wave_tiles = true;          // shallow-K GEMMs: widen the tile rather than run a second wave (PYCNN_B200_ONE_WAVE=0 disables)
  bool nvtx = false;                   // NVTX range per kernel class around every launch / capture (PYCNN_B200_NVTX=1)
  // layer-1 GEMMs fetch their batch rows from the dataset by TMA tile::gather4 -- no gathered batch in HBM (PYCNN_B200_GATHER_FREE=1).
  // Correct (bit-identical) but OFF by default: measured on B200, a gather4 op (4 rows x 128 B) costs ~70 SM cycles of the TMA
  // unit, 32 ops per 16 KB operand tile = 2240 cycles against the tile's 760 MMA cycles: layer-1 forward 22 -> 50 us, dW 22 -> 62 us.
  bool use_gather_free = false;
  struct GatherSrc {                   // set around a step whose batch is "rows idx[cursor + ...] of the dataset" (no d_x)
    bool on = false;
    int64_t advance = 0;               // rows the step's bookkeeping moves the cursor on by (dW of layer 1 looks back by it)
  } gsrc;
  bool use_tail = true;                // forward layers 2.. + softmax + the dA chain as one kernel (tail_fused.cuh; PYCNN_B200_FUSED_TAIL=0 disables)
  bool use_two_sm = true;              // deep 256-wide GEMM tiles run as CTA pairs on the 2-SM tensor-core path (PYCNN_B200_TWO_SM=0 disables)
  int pair_min_kb = 64;                // pairing pays from this many k-blocks per CTA on (PYCNN_B200_PAIR_MIN_KB)
  bool use_pair = true;                // M-adjacent GEMM CTAs share the B tile by TMA multicast (PYCNN_B200_PAIR=0 disables)
  bool use_cluster_splitk = true;      // split-K partials meet in distributed shared memory (PYCNN_B200_CLUSTER_SPLITK=0: global atomics)
  cudaEvent_t events[16] = {};
  cudaEvent_t copy_done = nullptr, compute_done = nullptr;
  int math_mode = PYCNN_B200_MATH_TF32;
  int update_rule = PYCNN_B200_RULE_CPU;
  int extract_mode = PYCNN_B200_EXTRACT_TC;

  // filters
  int F = 0;
  std::vector<double> h_taps;   // as given by the user [F][3][3]
  float* d_taps = nullptr;      // flipped, [F][9] fp32 (SIMT path)
  void* d_tc_filters = nullptr; // packed B operand for the tensor-core extractor
  int tc_ncols = 0;             // accumulator columns used by the TC extractor
  float* d_tc_scale = nullptr;  // per-filter epilogue scale for the TC extractor

  // network
  int64_t D = 0, Dp = 0;
  int L = 0, C = 0;
  std::vector<int64_t> dims;            // [D, n1..nL, C]
  std::vector<int64_t> ldw;             // pitch of act/feature rows per level: pad4(dims[i])
  std::vector<pb::TensorDesc> tensors;  // 2*(L+1): w,b per layer
  int64_t P = 0;                        // flat length in floats
  float *d_params = nullptr, *d_grads = nullptr, *d_m = nullptr, *d_v = nullptr;
  pb::Chunk* d_chunks = nullptr;
  int num_chunks = 0;
  std::vector<pb::Chunk> h_chunks;      // host copy of the chunk table
  int2* d_tensor_chunks = nullptr;      // per tensor: (first chunk, chunk count)
  double* d_tensor_norm = nullptr;      // [tensors] ||g||^2 per tensor (large networks only: tensor_norm_kernel)
  bool big_tensors = false;             // some tensor has more chunks than an update block should add up by itself
  double* d_chunk_norm = nullptr;       // [num_chunks] sum of squares of each chunk of the (reduced) gradient
  float* d_gpart = nullptr;             // ordered-partial-sum arena (sized by ensure_workspace for the batch capacity)
  int64_t gpart_floats = 0;
  pb::GradPartSlot gp[64];              // per tensor

  // optimizer
  int opt = PYCNN_B200_OPT_SGD;
  double lr = 1e-3, beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
  int64_t t = 0;

  // dataset
  int64_t cap_rows = 0, n_rows = 0;
  float* d_feat = nullptr;
  int32_t* d_labels = nullptr;

  // batch workspace
  int64_t bs_cap = 0;
  int64_t* d_idx = nullptr;
  int64_t idx_cap = 0;
  int32_t* d_blabels = nullptr;
  float* d_x = nullptr;               // gathered batch [bs, Dp]
  float* d_x2 = nullptr;              // second batch buffer (row-gather prefetch), allocated on first use
  int32_t* d_blabels2 = nullptr;
  std::vector<float*> d_act;          // [L] hidden activations [bs, ldw[l+1]]
  std::vector<float*> d_dz;           // [L+1] dZ per layer [bs, ldw[l+1]]
  float* d_probs = nullptr;           // [bs, ldw[L+1]]
  int32_t* d_pred = nullptr;          // [bs]
  float* d_conf = nullptr;            // [bs]
  float* d_gemm_ws = nullptr;         // split-K workspace
  size_t gemm_ws_bytes = 0;

  // stats accumulators (elementwise.cuh: stats_add): [0] loss sum in 2^-32 fixed point, [1] #correct, [2] non-finite count
  unsigned long long* d_stats = nullptr;
  unsigned long long* h_stats = nullptr;  // pinned

  // staging
  uint8_t* d_images = nullptr;
  size_t images_cap = 0;
  int32_t* d_img_labels = nullptr;
  size_t img_labels_cap = 0;
  void* d_scratch = nullptr;  // generic device scratch (f64 staging for host<->device converts)
  size_t scratch_cap = 0;
  void* h_pinned = nullptr;   // pinned bounce buffer
  size_t pinned_cap = 0;
  float* d_flush = nullptr;   // L2 flush buffer
  size_t flush_bytes = 0;

  // double-buffered staging for the streaming epoch (copy stream || compute stream)
  uint8_t* d_img2[2] = {nullptr, nullptr};
  int32_t* d_lab2[2] = {nullptr, nullptr};
  size_t img2_cap = 0, lab2_cap = 0;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_compute[2] = {nullptr, nullptr};
  uint8_t* h_bounce[2] = {nullptr, nullptr};   // pinned bounce buffers for pageable host memory
  size_t bounce_cap = 0;
  unsigned long long* h_step_stats = nullptr;  // pinned ring for per-step stats read-backs (kStatsWords words each)
  size_t step_stats_cap = 0;

  // comm
  void* nccl_comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange (p2p.cuh): opened IPC mappings of every rank's buffers
  bool p2p = false;
  void* p2p_sync = nullptr;                 // my SyncBlock (peers write into it)
  void* p2p_peer_ptrs[3][8] = {};           // [grads|params|sync][rank], own entries point at local buffers
  long long* d_p2p_seq = nullptr;           // [0] unused, [1] exchange sequence number (monotonic since attach)
  unsigned int* d_blocks_done = nullptr;
  int p2p_chunk0 = 0, p2p_nchunks = 0;
  // NVSwitch multicast exchange (nvls.cuh): one physical allocation per rank, mapped unicast and multicast
  bool nvls = false;                        // the step uses nvls_allreduce_kernel + the replicated update
  bool nvls_added = false, nvls_bound = false;
  bool nvls_split = true;                   // first layer's gradient exchanged right behind dW_1, the rest beside it (PYCNN_B200_NVLS_SPLIT=0: one exchange)
  bool nvls_fused = false;                  // PYCNN_B200_NVLS_FUSED=1: all-reduce + clip + update in one launch (measured slower: its waiting blocks fill the register files)
  CUmemGenericAllocationHandle nvls_mc = 0, nvls_mem = 0;
  CUdeviceptr nvls_uc = 0, nvls_mcva = 0;
  size_t nvls_size = 0, nvls_gran = 0, nvls_off_gred = 0, nvls_off_norm = 0, nvls_off_flags = 0;

  // timing
  bool timing = false;
  std::vector<pb::TimedLaunch> timed[pb::KC_COUNT];
  std::vector<cudaEvent_t> event_pool;
  int64_t class_launches[pb::KC_COUNT] = {};
  double class_ms[pb::KC_COUNT] = {};
  int64_t launches = 0;

  // TMA descriptor encoder (driver entry point resolved at create time)
  void* encode_tiled = nullptr;

  // CUDA-graph replay of the training step (the step is ~30 short launches: launch-bound when issued eagerly)
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
    int64_t class_launches[pb::KC_COUNT] = {};
  };
  std::map<std::tuple<int, int64_t, int64_t, int64_t, int>, GraphEntry> graphs;
  bool use_graphs = true;
  bool use_pdl = true;           // programmatic dependent launch between the step's kernels (PYCNN_B200_PDL=0 disables)
  int64_t* d_cursor = nullptr;   // [0] offset into d_idx of the current batch, [1] completed steps (device-side t)
};

namespace pb {

// RAII-ish scope that brackets one launch with events when timing is on.
struct LaunchScope {
  pycnn_b200_ctx* c;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  LaunchScope(pycnn_b200_ctx* ctx, int k, int n_launches = 1) : c(ctx), cls(k) {
    c->launches += n_launches;
    c->class_launches[cls] += n_launches;
    if (c->nvtx) nvtxRangePushA(kKernelClassNames[cls]);      // one range per kernel class (PYCNN_B200_NVTX=1)
    if (c->timing) {
      a = take();
      b = take();
      cudaEventRecord(a, c->stream);
    }
  }
  ~LaunchScope() {
    if (c->nvtx) nvtxRangePop();
    if (c->timing) {
      cudaEventRecord(b, c->stream);
      c->timed[cls].push_back({a, b});
    }
  }
  cudaEvent_t take() {
    cudaEvent_t e;
    if (!c->event_pool.empty()) {
      e = c->event_pool.back();
      c->event_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }
};

// Programmatic dependent launch: the next kernel of the step may start launching (and run its prologue) while
// this one drains; every kernel calls pdl_wait() before it touches global memory produced by its predecessor.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A timeline slot is kTraceWords words: [0] min CTA start, [1] max CTA end, then (min, max) pairs of up to three
// kernel-specific phase marks (GEMMs: first operand stage landed, accumulator complete; see gemm_tc.cuh).
constexpr int kTraceWords = 8;
__device__ __forceinline__ void trace_begin(unsigned long long* slot) {
  if (slot && threadIdx.x == 0) atomicMin(slot, global_timer_ns());
}
__device__ __forceinline__ void trace_end(unsigned long long* slot) {
  if (slot && threadIdx.x == 0) atomicMax(slot + 1, global_timer_ns());
}
__device__ __forceinline__ void trace_mark(unsigned long long* slot, int phase) {   // phase 1..3, one thread per CTA
  if (slot) {
    const unsigned long long t = global_timer_ns();
    atomicMin(slot + 2 * phase, t);
    atomicMax(slot + 2 * phase + 1, t);
  }
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch as thread-block clusters of (cluster_x, cluster_y, cluster_z) when that is more than one CTA
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl_cluster3(pycnn_b200_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                                              size_t smem, cudaStream_t stream, int cluster_x, int cluster_y, int cluster_z,
                                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (c->use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster_x * cluster_y * cluster_z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x; attr[na].val.clusterDim.y = (unsigned)cluster_y; attr[na].val.clusterDim.z = (unsigned)cluster_z;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl_cluster(pycnn_b200_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                                             size_t smem, cudaStream_t stream, int cluster_y, int cluster_z, Args&&... args) {
  return launch_pdl_cluster3(c, kernel, grid, block, smem, stream, 1, cluster_y, cluster_z, static_cast<Args&&>(args)...);
}

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(pycnn_b200_ctx* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t stream, Args&&... args) {
  return launch_pdl_cluster(c, kernel, grid, block, smem, stream, 1, 1, static_cast<Args&&>(args)...);
}

// next timeline slot for a launch of class `cls` on c->stream (nullptr when tracing is off or the table is full)
static inline unsigned long long* trace_slot(pycnn_b200_ctx* c, int cls) {
  if (!c->trace || !c->d_trace || c->trace_next >= c->trace_cap) return nullptr;
  const int sid = (c->stream == c->side_stream) ? 1 : (c->stream == c->pf_stream) ? 2 : 0;
  c->trace_cls.push_back(cls + 100 * sid);
  return c->d_trace + kTraceWords * (size_t)(c->trace_next++);
}

static inline int ensure_scratch(pycnn_b200_ctx* c, size_t bytes) {
  if (bytes <= c->scratch_cap) return 0;
  if (c->d_scratch) PB_CUDA(cudaFree(c->d_scratch));
  c->d_scratch = nullptr;
  c->scratch_cap = 0;
  size_t cap = (size_t)round_up((int64_t)bytes, 1 << 20);
  PB_CUDA(cudaMalloc(&c->d_scratch, cap));
  c->scratch_cap = cap;
  return 0;
}

static inline int ensure_pinned(pycnn_b200_ctx* c, size_t bytes) {
  if (bytes <= c->pinned_cap) return 0;
  if (c->h_pinned) PB_CUDA(cudaFreeHost(c->h_pinned));
  c->h_pinned = nullptr;
  c->pinned_cap = 0;
  size_t cap = (size_t)round_up((int64_t)bytes, 1 << 20);
  PB_CUDA(cudaMallocHost(&c->h_pinned, cap));
  c->pinned_cap = cap;
  return 0;
}

}  // namespace pb
