// api.cu -- the C ABI of libpycnn_b200.so (see include/pycnn_b200.h for what each entry replaces).
#include <algorithm>

#include "comm.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "extract_tc.cuh"
#include "gemm_tc.cuh"
#include <chrono>

#include "p2p.cuh"
#include "nvls.cuh"
#include "simt.cuh"
#include "tail_fused.cuh"

using namespace pb;

namespace {

constexpr int64_t kSlackFloats = 128;  // tail slack behind every fp32 buffer (vector tails, partial tiles)

void drop_graphs(pycnn_b200_ctx* c);
int p2p_exchange_update(pycnn_b200_ctx* c);

int alloc_f32(float** p, int64_t n) {
  PB_CUDA(cudaMalloc(p, sizeof(float) * (size_t)(n + kSlackFloats)));
  PB_CUDA(cudaMemset(*p, 0, sizeof(float) * (size_t)(n + kSlackFloats)));
  return 0;
}
void free_dev(void* p) {
  if (p) cudaFree(p);
}

const TensorDesc& Wt(const pycnn_b200_ctx* c, int l) { return c->tensors[2 * l]; }
const TensorDesc& Bt(const pycnn_b200_ctx* c, int l) { return c->tensors[2 * l + 1]; }

int grid_for(int64_t total, int threads, int num_sms) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, threads), (int64_t)num_sms * 16));
}

// ---------------------------------------------------------------------------------------------
// GEMM dispatch: TF32 tensor cores by default, FP32 SIMT in verification mode
// ---------------------------------------------------------------------------------------------
// bias gradient db = column sums of dz [M, N] as ordered partials (one per row chunk) in tensor `slot`'s arena region
int colsum_partials(pycnn_b200_ctx* c, const float* dz, int M, int N, int64_t ld, int slot) {
  GradPartSlot& gs = c->gp[slot];
  const int col_blocks = (int)ceil_div(N, 32);
  const int chunks = std::max(1, std::min({(int)ceil_div(M, 64), (2 * c->num_sms + col_blocks - 1) / col_blocks, gs.cap}));
  const int rows_per_block = (int)ceil_div(M, chunks);
  dim3 cgrid((unsigned)col_blocks, (unsigned)ceil_div(M, rows_per_block));
  PB_CHECK((int)cgrid.y <= gs.cap, "bias-gradient partial arena too small (%d row chunks, cap %d)", (int)cgrid.y, gs.cap);
  gs.nparts = (int)cgrid.y;
  LaunchScope ls(c, KC_COLSUM);
  PB_CUDA(launch_pdl(c, colsum_kernel, cgrid, dim3(32, 8), 0, c->stream, dz, M, N, ld, c->d_gpart + gs.off, gs.stride, rows_per_block));
  return 0;
}

int gemm(pycnn_b200_ctx* c, const GemmCall& g) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (c->math_mode == PYCNN_B200_MATH_TF32) return tc_gemm(c, g);
  SimtGemmArgs a;
  a.A = g.A; a.B = g.B; a.C = g.C; a.ldc = g.ldc;
  a.sam = g.a_kmajor ? g.lda : 1; a.sak = g.a_kmajor ? 1 : g.lda;
  a.sbn = g.b_kmajor ? g.ldb : 1; a.sbk = g.b_kmajor ? 1 : g.ldb;
  a.M = g.M; a.N = g.N; a.K = g.K; a.epi = g.epi;
  a.bias = g.bias; a.mask = g.mask; a.ldmask = g.ldmask;
  dim3 grid((unsigned)ceil_div(g.N, 64), (unsigned)ceil_div(g.M, 64));
  {
    LaunchScope ls(c, g.cls);
    sgemm_simt_kernel<<<grid, 256, 0, c->stream>>>(a);
    PB_CUDA(cudaGetLastError());
  }
  if (g.grad_slot) c->gp[g.grad_slot - 1].nparts = 0;   // unsplit: C is the flat gradient buffer itself
  if (g.colsum_slot) PB_TRY(colsum_partials(c, g.C, g.M, g.N, g.ldc, g.colsum_slot - 1));   // the tensor-core kernel fuses this
  return 0;
}

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
void free_workspace(pycnn_b200_ctx* c) {
  free_dev(c->d_gpart); c->d_gpart = nullptr; c->gpart_floats = 0;
  for (auto& g : c->gp) g = GradPartSlot{};
  free_dev(c->d_blabels); c->d_blabels = nullptr;
  free_dev(c->d_x); c->d_x = nullptr;
  free_dev(c->d_x2); c->d_x2 = nullptr;
  free_dev(c->d_blabels2); c->d_blabels2 = nullptr;
  for (auto p : c->d_act) free_dev(p);
  for (auto p : c->d_dz) free_dev(p);
  c->d_act.clear(); c->d_dz.clear();
  free_dev(c->d_probs); c->d_probs = nullptr;
  free_dev(c->d_pred); c->d_pred = nullptr;
  free_dev(c->d_conf); c->d_conf = nullptr;
  c->bs_cap = 0;
}

int ensure_workspace(pycnn_b200_ctx* c, int64_t bs) {
  PB_CHECK(c->L >= 0 && c->dims.size() >= 2, "set_network must be called first");
  if (bs <= c->bs_cap) return 0;
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  free_workspace(c);
  const int64_t cap = round_up(bs, 128);
  PB_CUDA(cudaMalloc(&c->d_blabels, sizeof(int32_t) * cap));
  PB_CUDA(cudaMemset(c->d_blabels, 0, sizeof(int32_t) * cap));
  PB_TRY(alloc_f32(&c->d_x, cap * c->Dp));
  c->d_act.assign(c->L, nullptr);
  c->d_dz.assign(c->L + 1, nullptr);
  for (int l = 0; l < c->L; ++l) PB_TRY(alloc_f32(&c->d_act[l], cap * c->ldw[l + 1]));
  for (int l = 0; l <= c->L; ++l) PB_TRY(alloc_f32(&c->d_dz[l], cap * c->ldw[l + 1]));
  PB_TRY(alloc_f32(&c->d_probs, cap * c->ldw[c->L + 1]));
  PB_CUDA(cudaMalloc(&c->d_pred, sizeof(int32_t) * cap));
  PB_CUDA(cudaMalloc(&c->d_conf, sizeof(float) * cap));
  // Ordered-partial-sum arena: per weight tensor as many slices as its dW GEMM can be split into at this batch capacity
  // (the split grows with K = batch rows), per bias one partial per row tile / softmax block / column-sum chunk.
  int64_t off = 0;
  for (int l = 0; l <= c->L; ++l) {
    GemmCall g{};
    g.a_kmajor = false; g.b_kmajor = false;
    g.M = (int)c->dims[l + 1]; g.N = (int)c->dims[l]; g.K = (int)cap; g.grad_slot = 2 * l + 1;
    const GemmPlan p = plan_tc_gemm(c, g);
    GradPartSlot& w = c->gp[2 * l];
    w.off = off; w.stride = Wt(c, l).rows * Wt(c, l).ld; w.cap = p.split > 1 ? p.split : 0; w.nparts = 0;
    off = round_up(off + w.stride * w.cap, 64);
    GradPartSlot& b = c->gp[2 * l + 1];
    b.off = off; b.stride = Bt(c, l).ld; b.nparts = 0;
    b.cap = (int)std::max<int64_t>(round_up(ceil_div(cap, tc::G_BM), 2), 2 * c->num_sms);
    off = round_up(off + b.stride * b.cap, 64);
  }
  PB_TRY(alloc_f32(&c->d_gpart, off));
  c->gpart_floats = off;
  c->bs_cap = cap;
  return 0;
}

// second batch buffer for the row-gather prefetch (same capacity as d_x); only the row-training entry points need it
int ensure_prefetch_buffers(pycnn_b200_ctx* c) {
  if (c->d_x2 || c->bs_cap == 0) return 0;
  PB_CUDA(cudaMalloc(&c->d_blabels2, sizeof(int32_t) * c->bs_cap));
  PB_CUDA(cudaMemset(c->d_blabels2, 0, sizeof(int32_t) * c->bs_cap));
  PB_TRY(alloc_f32(&c->d_x2, c->bs_cap * c->Dp));
  return 0;
}

int ensure_idx(pycnn_b200_ctx* c, int64_t n) {
  if (n <= c->idx_cap) return 0;
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  free_dev(c->d_idx);
  c->d_idx = nullptr;
  c->idx_cap = 0;
  PB_CUDA(cudaMalloc(&c->d_idx, sizeof(int64_t) * (size_t)round_up(n, 1024)));
  c->idx_cap = round_up(n, 1024);
  return 0;
}

int ensure_images(pycnn_b200_ctx* c, size_t bytes, size_t n_labels) {
  if (bytes > c->images_cap) {
    PB_CUDA(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    free_dev(c->d_images);
    c->d_images = nullptr;
    c->images_cap = 0;
    size_t cap = (size_t)round_up((int64_t)bytes + 256, 1 << 20);
    PB_CUDA(cudaMalloc(&c->d_images, cap));
    PB_CUDA(cudaMemset(c->d_images, 0, cap));
    c->images_cap = cap;
  }
  if (n_labels > c->img_labels_cap) {
    PB_CUDA(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    free_dev(c->d_img_labels);
    c->d_img_labels = nullptr;
    PB_CUDA(cudaMalloc(&c->d_img_labels, sizeof(int32_t) * (size_t)round_up((int64_t)n_labels, 1024)));
    c->img_labels_cap = (size_t)round_up((int64_t)n_labels, 1024);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// extractor
// ---------------------------------------------------------------------------------------------
int64_t feature_count(const pycnn_b200_ctx* c, int S) {
  const int64_t p = (S - 2) / 2;
  return (S >= 2 ? p * p : 0) * c->F;
}

// d_images: n x S x S x 3 u8 on the device -> out rows [n, ld] fp32 (filter-major), on c->stream
int extract_device(pycnn_b200_ctx* c, const uint8_t* d_images, int64_t n, int S, float* out, int64_t ld) {
  PB_CHECK(c->F > 0, "set_filters must be called before extraction");
  PB_CHECK(S >= 4, "image_size must be >= 4 (got %d)", S);
  if (n == 0) return 0;
  const int PH = (S - 2) / 2, PW = PH;
  if (c->extract_mode == PYCNN_B200_EXTRACT_TC && tc_extract_supported(c, S)) {
    return tc_extract(c, d_images, n, S, out, ld);
  }
  const int tiles = (int)ceil_div(PW, EX_T);
  const size_t smem = sizeof(float) * (EX_IN * (EX_IN + 1) + (size_t)c->F * 9);
  PB_CHECK(smem <= 48 * 1024, "too many filters for the SIMT extractor (%d)", c->F);
  for (int64_t n0 = 0; n0 < n; n0 += 65535) {  // gridDim.y limit
    const int nb = (int)std::min<int64_t>(65535, n - n0);
    dim3 grid(tiles * tiles, nb);
    LaunchScope ls(c, KC_EXTRACT);
    extract_simt_kernel<<<grid, 256, smem, c->stream>>>(d_images + n0 * (int64_t)S * S * 3, S, PH, PW, c->F,
                                                        c->d_taps, out + n0 * ld, ld, tiles);
    PB_CUDA(cudaGetLastError());
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// forward / backward / update on the batch currently in c->d_x (or an external x) and c->d_blabels
// ---------------------------------------------------------------------------------------------

int issue_prefetch(pycnn_b200_ctx* c);
// Where the next step's row gather (HBM-bound, ~30-39 us) is forked.  Single GPU: behind the layer-1 forward GEMM, beside the
// latency-bound tail.  Data parallel: behind grad_finalize, beside the gradient exchange and the update -- that phase is
// NVLink-latency-bound with HBM idle and about as long as the gather, and the GEMMs and finalize then run undisturbed.
bool pf_behind_dw1(const pycnn_b200_ctx* c) { return c->pf_place == 2; }
// 3: behind grad_finalize, i.e. beside the exchange kernel(s) and the update only (finalize runs undisturbed)
bool pf_beside_exchange(const pycnn_b200_ctx* c) { return c->pf_place == 3 || (c->pf_place == 0 && c->world > 1); }

// Gradients are never accumulated with atomics and the flat gradient buffer is never zero-filled: every producer
// either writes a tensor's gradient whole or stores ordered partial sums (GradPartSlot) that finalize_grads adds in
// index order -- training is bit-reproducible.  begin_grads forgets the previous step's partial counts.
void begin_grads(pycnn_b200_ctx* c) {
  for (size_t t = 0; t < c->tensors.size(); ++t) c->gp[t].nparts = 0;
}

// partial sums -> flat gradient buffer (+ per-chunk sums of squares for the clip when want_norm)
int finalize_grads(pycnn_b200_ctx* c, bool want_norm, int chunk0 = 0, int nchunks = -1) {
  if (nchunks < 0) nchunks = c->num_chunks - chunk0;
  if (nchunks <= 0) return 0;
  GradParts gp{};
  gp.base = c->d_gpart;
  bool any = want_norm;
  for (size_t t = 0; t < c->tensors.size(); ++t) {
    gp.off[t] = c->gp[t].off; gp.stride[t] = c->gp[t].stride; gp.nparts[t] = c->gp[t].nparts;
    any = any || c->gp[t].nparts > 0;
  }
  if (!any) return 0;
  LaunchScope ls(c, KC_SQNORM);
  PB_CUDA(launch_pdl(c, grad_finalize_kernel, dim3(nchunks), dim3(256), 0, c->stream, c->d_grads,
                     (const Chunk*)c->d_chunks + chunk0, gp, c->d_chunk_norm + chunk0, want_norm ? 1 : 0, trace_slot(c, KC_SQNORM)));
  return 0;
}

int zero_grads(pycnn_b200_ctx* c) {   // a rank whose shard of the batch is empty contributes zeros
  LaunchScope ls(c, KC_MISC);
  PB_CUDA(cudaMemsetAsync(c->d_grads, 0, sizeof(float) * c->P, c->stream));
  return 0;
}

// The output layer's GEMM can finish the training forward pass itself (softmax, CE, dZ, db_o, tallies, bookkeeping)
bool softmax_fusable(const pycnn_b200_ctx* c) {
  return c->fuse_softmax && c->math_mode == PYCNN_B200_MATH_TF32 && c->C <= 32;
}

// x: [bs, Dp].  Leaves hidden activations in d_act and logits in d_probs (pitch ldw[L+1]).
// train_labels != nullptr (only with softmax_fusable): the last GEMM runs the fused softmax/CE epilogue instead of
// writing logits -- dZ_out lands in d_dz[L], loss/#correct in d_stats, db_o as ordered partials (one per row tile).
int forward_logits(pycnn_b200_ctx* c, const float* x, int bs, const int32_t* train_labels = nullptr, int64_t advance = 0,
                   int last_layer = -1) {
  const float* in = x;
  if (last_layer < 0) last_layer = c->L;
  for (int l = 0; l <= last_layer; ++l) {
    GemmCall g{};
    g.A = in; g.lda = c->ldw[l]; g.a_kmajor = true;
    g.B = c->d_params + Wt(c, l).off; g.ldb = Wt(c, l).ld; g.b_kmajor = true;
    g.M = bs; g.N = (int)c->dims[l + 1]; g.K = (int)c->dims[l];
    if (l == 0 && c->gsrc.on) {            // gather-free: the batch is rows idx[cursor ...) of the dataset
      g.A = c->d_feat; g.gather_idx = c->d_idx; g.gather_cursor = c->d_cursor; g.gather_extra = 0; g.gather_rows = c->n_rows;
    }
    g.bias = c->d_params + Bt(c, l).off;
    g.cls = (l == 0) ? KC_GEMM_FWD1 : KC_GEMM_FWD;
    if (l < c->L) { g.C = c->d_act[l]; g.epi = 2; }
    else          { g.C = c->d_probs;  g.epi = 1; }
    g.ldc = c->ldw[l + 1];
    if (l == c->L && train_labels) {
      g.epi = 4;
      g.labels = train_labels; g.dz = c->d_dz[c->L]; g.lddz = c->ldw[c->L + 1]; g.stats = c->d_stats;
      g.db_slot = 2 * c->L + 2;
      g.bk = StepBookkeeping{c->d_cursor, advance};
    }
    if (l == 0 && c->pf_early) PB_TRY(issue_prefetch(c));   // no-op unless a pipelined training step armed it
    PB_TRY(gemm(c, g));
    if (l == 0 && !pf_behind_dw1(c) && !pf_beside_exchange(c)) PB_TRY(issue_prefetch(c));
    in = g.C;
  }
  return 0;
}

// logits live in c->d_probs.  train: writes dZ, accumulates loss/#correct and db_o, does the step bookkeeping.
int softmax_ce(pycnn_b200_ctx* c, int bs, const int32_t* labels, bool write_probs, bool write_dz, bool write_pred,
               bool stats, bool train = false, int64_t advance = 0) {
  LaunchScope ls(c, KC_SOFTMAX_CE);
  int blocks = (int)ceil_div(bs, 8);
  StepBookkeeping bk{nullptr, 0};
  float* db_part = nullptr;
  size_t smem = 0;
  if (train) {
    bk.cursor = c->d_cursor; bk.advance = advance;
    GradPartSlot& gs = c->gp[2 * c->L + 1];            // db_o: one ordered partial per block
    blocks = std::min(blocks, std::min(gs.cap, 2 * c->num_sms));
    gs.nparts = blocks;
    db_part = c->d_gpart + gs.off;
    smem = sizeof(float) * 8 * (size_t)c->C;
    PB_CHECK(smem <= 48 * 1024, "too many classes for the softmax kernel's bias-gradient staging (%d)", c->C);
  }
  PB_CUDA(launch_pdl(c, softmax_ce_kernel, dim3(blocks), dim3(256), smem, c->stream, (const float*)c->d_probs, bs, c->C,
                     c->ldw[c->L + 1], labels, write_probs ? c->d_probs : (float*)nullptr,
                     write_dz ? c->d_dz[c->L] : (float*)nullptr, write_pred ? c->d_pred : (int32_t*)nullptr,
                     write_pred ? c->d_conf : (float*)nullptr, stats ? c->d_stats : (unsigned long long*)nullptr, db_part, bk));
  return 0;
}

// ---- fused tail (tail_fused.cuh): forward layers 2..L+1, softmax/CE and the whole dA chain in one launch -------------
bool tail_eligible(const pycnn_b200_ctx* c) {
  if (!c->use_tail || c->math_mode != PYCNN_B200_MATH_TF32 || !c->fuse_softmax) return false;
  if (c->L < 1 || 2 * c->L > tail::MAX_STAGES || c->C > 32 || c->dims[1] > tail::MAX_IN) return false;
  for (int l = 2; l <= c->L; ++l)
    if (c->dims[l] > tail::MAX_MID) return false;
  return true;
}

// a1 = relu(z1 + b1) is in d_act[0].  Leaves a2..aL in d_act, dZ_{L+1}..dZ_1 in d_dz, loss/#correct in d_stats and the
// every dZ_l in d_dz, and the bias gradients of every layer as ordered partials (one per 128-row tile).
int tail_forward_backward(pycnn_b200_ctx* c, int bs, const int32_t* labels, int64_t advance) {
  const int L = c->L, tiles = (int)ceil_div(bs, 128);
  auto bias_part = [&](int layer, float** part, int64_t* ld) -> int {      // ordered partial slot of b_layer (0-based layer)
    GradPartSlot& gs = c->gp[2 * layer + 1];
    PB_CHECK(tiles <= gs.cap, "bias-gradient partial arena too small (%d row tiles, cap %d)", tiles, gs.cap);
    gs.nparts = tiles;
    *part = c->d_gpart + gs.off; *ld = gs.stride;
    return 0;
  };
  tail::TailArgs t{};
  t.bs = bs; t.nstages = 2 * L;
  t.a_in = c->d_act[0]; t.ld_in = c->ldw[1]; t.in_width = (int)c->dims[1];
  t.labels = labels; t.stats = c->d_stats;
  t.bk = StepBookkeeping{c->d_cursor, advance};
  tail::TailMaps tm;
  PB_TRY(make_map_kmajor(c, &tm.in, c->d_act[0], bs, c->dims[1], c->ldw[1], 128));
  int s = 0;
  for (int l = 1; l <= L; ++l, ++s) {            // forward: network layer l (0-based) = hidden layers 2..L, then the output layer
    tail::Stage& S = t.st[s];
    const TensorDesc& w = Wt(c, l);
    S.K = (int)c->dims[l]; S.N = (int)c->dims[l + 1]; S.b_mn = 0;
    S.bias = c->d_params + Bt(c, l).off;
    if (l < L) { S.epi = 0; S.mask_slot = l; S.out = c->d_act[l]; S.ld_out = c->ldw[l + 1]; }
    else       { S.epi = 1; S.out = c->d_dz[L]; S.ld_out = c->ldw[L + 1]; PB_TRY(bias_part(L, &S.colsum_part, &S.ld_colsum)); }
    PB_TRY(make_map_kmajor(c, &tm.b[s], c->d_params + w.off, w.rows, w.cols, w.ld, (int)round_up(std::min<int64_t>(w.rows, 128), 16)));
    PB_TRY(make_map_kmajor(c, &tm.o[s], S.out, bs, S.N, S.ld_out, 128));     // the stage's output tile leaves by TMA store
  }
  for (int l = L; l >= 1; --l, ++s) {            // backward: dZ_{l-1} = (dZ_l . W_l) masked by a_{l-1} > 0   (0-based l)
    tail::Stage& S = t.st[s];
    const TensorDesc& w = Wt(c, l);
    S.K = (int)c->dims[l + 1]; S.N = (int)c->dims[l]; S.b_mn = 1; S.epi = 2; S.mask_slot = l - 1;
    S.out = c->d_dz[l - 1]; S.ld_out = c->ldw[l];
    PB_TRY(bias_part(l - 1, &S.colsum_part, &S.ld_colsum));
    PB_TRY(make_map_mnmajor(c, &tm.b[s], c->d_params + w.off, w.rows, w.cols, w.ld));
    PB_TRY(make_map_kmajor(c, &tm.o[s], S.out, bs, S.N, S.ld_out, 128));
  }
  for (; s < tail::MAX_STAGES; ++s) { tm.b[s] = tm.b[0]; tm.o[s] = tm.o[0]; }
  t.trace = trace_slot(c, KC_GEMM_FWD);
  LaunchScope ls(c, KC_GEMM_FWD);
  PB_CUDA(launch_pdl(c, tail::tail_fused_kernel, dim3((unsigned)tiles), dim3(tail::THREADS), (size_t)tail::SMEM_BYTES, c->stream,
                     tm, t));
  return 0;
}

// dW GEMMs after the fused tail: dW of layers 2..L+1 need only what the tail kernel left in memory, so they all run on
// the side stream, capped to the SMs dW of layer 1 (main stream, the critical path) leaves idle.
// join == false: the shallow dW GEMMs are left running on the side stream (the caller finishes their part of the step
// there and joins later); *forked tells whether there is anything to join.
int backward_dw_after_tail(pycnn_b200_ctx* c, const float* x, int bs, bool join = true, bool* forked = nullptr) {
  const bool fork = !c->timing && c->side_stream != nullptr;
  cudaStream_t main_stream = c->stream;
  auto dw_call = [&](int l) {
    GemmCall g{};
    g.A = c->d_dz[l]; g.lda = c->ldw[l + 1]; g.a_kmajor = false;
    g.B = (l == 0) ? x : c->d_act[l - 1]; g.ldb = c->ldw[l]; g.b_kmajor = false;
    g.C = c->d_grads + Wt(c, l).off; g.ldc = Wt(c, l).ld;
    g.M = (int)c->dims[l + 1]; g.N = (int)c->dims[l]; g.K = bs; g.epi = 0; g.cls = (l == 0) ? KC_GEMM_DW1 : KC_GEMM_DW;
    g.grad_slot = 2 * l + 1;
    if (l == 0 && c->gsrc.on) {            // the step's bookkeeping has already moved the cursor past this batch
      g.B = c->d_feat; g.gather_idx = c->d_idx; g.gather_cursor = c->d_cursor; g.gather_extra = -c->gsrc.advance; g.gather_rows = c->n_rows;
    }
    return g;
  };
  if (fork) PB_CUDA(cudaEventRecord(c->ev_fork[0], main_stream));
  PB_TRY(gemm(c, dw_call(0)));
  const int room = c->num_sms - c->last_gemm_ctas;
  if (pf_behind_dw1(c)) PB_TRY(issue_prefetch(c));
  if (fork) {
    PB_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_fork[0], 0));
    c->stream = c->side_stream;
  }
  int rc = 0;
  for (int l = c->L; l >= 1 && !rc; --l) {
    GemmCall g = dw_call(l);
    if (fork && room >= 8) g.max_ctas = room;
    rc = gemm(c, g);
  }
  c->stream = main_stream;
  PB_TRY(rc);
  if (forked) *forked = fork;
  if (fork && join) {
    PB_CUDA(cudaEventRecord(c->ev_join, c->side_stream));
    PB_CUDA(cudaStreamWaitEvent(main_stream, c->ev_join, 0));
  }
  return 0;
}

// The critical path of the backward pass is dZ_L -> dA chain -> dZ_1 -> dW_1; the shallow dW_l (l >= 1) only need
// dZ_l and the stored activations, so they are forked onto a side stream as soon as their dZ exists and joined
// before the gradient exchange / update.  (While per-kernel timing is on everything stays on one stream.)
//
// The last shallow dW (layer 2) becomes ready only ~10 us before dW_1 and its 32 split-K CTAs used to hold SMs that
// dW_1's single wave (136 CTAs at configs[1]) then had to wait for (+7 us on the critical path, scripts/
// step_timeline.py).  It is therefore released together with dW_1 (same dependency: the last dA GEMM) and capped to
// the SMs dW_1 leaves idle, so that it runs entirely in dW_1's shadow.
int backward(pycnn_b200_ctx* c, const float* x, int bs) {
  const bool fork = !c->timing && c->side_stream != nullptr && c->L >= 1 && c->L < 8;
  const bool defer = fork && c->defer_last_dw && c->math_mode == PYCNN_B200_MATH_TF32;
  GemmCall deferred{};
  bool have_deferred = false;
  cudaStream_t main_stream = c->stream;
  for (int l = c->L; l >= 0; --l) {
    const float* in = (l == 0) ? x : c->d_act[l - 1];
    const int out_n = (int)c->dims[l + 1], in_n = (int)c->dims[l];
    {  // dW_l = dZ_l^T . in   (backward_pass.pyx:71,92)
      GemmCall g{};
      g.A = c->d_dz[l]; g.lda = c->ldw[l + 1]; g.a_kmajor = false;
      g.B = in; g.ldb = c->ldw[l]; g.b_kmajor = false;
      g.C = c->d_grads + Wt(c, l).off; g.ldc = Wt(c, l).ld;
      g.M = out_n; g.N = in_n; g.K = bs; g.epi = 0; g.cls = (l == 0) ? KC_GEMM_DW1 : KC_GEMM_DW;
      g.grad_slot = 2 * l + 1;
      if (l == 0 && c->gsrc.on) {
        g.B = c->d_feat; g.gather_idx = c->d_idx; g.gather_cursor = c->d_cursor; g.gather_extra = -c->gsrc.advance; g.gather_rows = c->n_rows;
      }
      if (defer && l == 1) {
        deferred = g; have_deferred = true;      // launched below, behind dW of layer 1
      } else if (fork && l > 0) {
        PB_CUDA(cudaEventRecord(c->ev_fork[l], main_stream));
        PB_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_fork[l], 0));
        c->stream = c->side_stream;
        const int rc = gemm(c, g);
        c->stream = main_stream;
        PB_TRY(rc);
      } else {
        PB_TRY(gemm(c, g));
      }
    }
    if (l > 0) {  // dZ_{l-1} = (dZ_l . W_l) masked by act_{l-1} > 0 (backward_pass.pyx:85-86,97); db_{l-1} = colsum (:93)
      GemmCall g{};
      g.A = c->d_dz[l]; g.lda = c->ldw[l + 1]; g.a_kmajor = true;
      g.B = c->d_params + Wt(c, l).off; g.ldb = Wt(c, l).ld; g.b_kmajor = false;
      g.C = c->d_dz[l - 1]; g.ldc = c->ldw[l];
      g.M = bs; g.N = in_n; g.K = out_n; g.epi = 3;
      g.mask = c->d_act[l - 1]; g.ldmask = c->ldw[l];
      g.colsum_slot = 2 * (l - 1) + 2;
      g.cls = KC_GEMM_DA;
      PB_TRY(gemm(c, g));
      if (have_deferred && l == 1) PB_CUDA(cudaEventRecord(c->ev_fork[1], main_stream));   // = dW_1's dependency
    }
  }
  if (pf_behind_dw1(c)) PB_TRY(issue_prefetch(c));
  if (have_deferred) {
    const int room = c->num_sms - c->last_gemm_ctas;      // last launch on this path: dW of layer 1
    if (room >= 8) deferred.max_ctas = room;
    PB_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_fork[1], 0));
    c->stream = c->side_stream;
    const int rc = gemm(c, deferred);
    c->stream = main_stream;
    PB_TRY(rc);
  }
  if (fork) {   // join: everything the side stream did is ordered before the exchange / update
    PB_CUDA(cudaEventRecord(c->ev_join, c->side_stream));
    PB_CUDA(cudaStreamWaitEvent(main_stream, c->ev_join, 0));
  }
  return 0;
}

int allreduce_grads(pycnn_b200_ctx* c) {
  if (c->world <= 1 || !c->nccl_comm) return 0;
  NcclApi* api = nccl_api();
  PB_CHECK(api, "NCCL is not loaded");
  LaunchScope ls(c, KC_ALLREDUCE);
  PB_NCCL(api, api->AllReduce(c->d_grads, c->d_grads, (size_t)c->P, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
  return 0;
}

// norms_ready: finalize_grads(want_norm) of this step already left the per-chunk sums of squares in d_chunk_norm
int apply_update(pycnn_b200_ctx* c, bool norms_ready = false, const float* grads = nullptr, double* chunk_norm = nullptr) {
  if (!grads) grads = c->d_grads;
  if (!chunk_norm) chunk_norm = c->d_chunk_norm;
  UpdateParams u{};
  u.opt = c->opt; u.rule = c->update_rule;
  u.lr = (float)c->lr; u.beta1 = (float)c->beta1; u.beta2 = (float)c->beta2; u.eps = (float)c->eps;
  u.eps2 = (float)(c->eps * c->eps); u.max_norm = 5.0f;
  double bc1 = 1.0, bc2 = 1.0;
  if (c->opt == PYCNN_B200_OPT_ADAM && c->update_rule == PYCNN_B200_RULE_CPU) {
    // gradient_decent.pyx:58-59 with self.t never advanced by train_model: t == 1 for the whole run
    bc1 = 1.0 - c->beta1;
    bc2 = 1.0 - c->beta2;
    if (std::fabs(bc1) < 1e-8) bc1 = 1e-8;   // gradient_decent.pyx:64-67
    if (std::fabs(bc2) < 1e-8) bc2 = 1e-8;
  }
  u.inv_bc1 = (float)(1.0 / bc1); u.inv_bc2 = (float)(1.0 / bc2);
  u.beta1_d = c->beta1; u.beta2_d = c->beta2;
  u.steps_done = c->d_cursor;   // CuPy rule: bias corrections from the device-side step count (graph-replayable)
  if (c->update_rule == PYCNN_B200_RULE_CPU && !norms_ready) {   // clip norms of a gradient that is complete in d_grads
    LaunchScope ls(c, KC_SQNORM);
    PB_CUDA(launch_pdl(c, grad_sqnorm_kernel, dim3(c->num_chunks), dim3(256), 0, c->stream, grads,
                       (const Chunk*)c->d_chunks, chunk_norm, trace_slot(c, KC_SQNORM)));
  }
  if (c->update_rule == PYCNN_B200_RULE_CPU && c->big_tensors) {     // per-tensor totals once, not once per block
    LaunchScope ls(c, KC_SQNORM);
    PB_CUDA(launch_pdl(c, tensor_norm_kernel, dim3((unsigned)c->tensors.size()), dim3(256), 0, c->stream, (const double*)chunk_norm,
                       (const int2*)c->d_tensor_chunks, c->d_tensor_norm));
    u.tensor_norm = c->d_tensor_norm;
  }
  LaunchScope ls(c, KC_UPDATE);
  PB_CUDA(launch_pdl(c, clip_update_kernel, dim3(c->num_chunks), dim3(256), 0, c->stream, c->d_params,
                     grads, c->d_m, c->d_v, (const Chunk*)c->d_chunks, (const double*)chunk_norm, u,
                     trace_slot(c, KC_UPDATE)));
  return 0;
}

UpdateParams make_update_params(pycnn_b200_ctx* c) {
  UpdateParams u{};
  u.opt = c->opt; u.rule = c->update_rule;
  u.lr = (float)c->lr; u.beta1 = (float)c->beta1; u.beta2 = (float)c->beta2; u.eps = (float)c->eps;
  u.eps2 = (float)(c->eps * c->eps); u.max_norm = 5.0f;
  double bc1 = 1.0, bc2 = 1.0;
  if (c->opt == PYCNN_B200_OPT_ADAM && c->update_rule == PYCNN_B200_RULE_CPU) {
    bc1 = 1.0 - c->beta1;
    bc2 = 1.0 - c->beta2;
    if (std::fabs(bc1) < 1e-8) bc1 = 1e-8;
    if (std::fabs(bc2) < 1e-8) bc2 = 1e-8;
  }
  u.inv_bc1 = (float)(1.0 / bc1); u.inv_bc2 = (float)(1.0 / bc2);
  u.beta1_d = c->beta1; u.beta2_d = c->beta2;
  u.steps_done = c->d_cursor;
  return u;
}

// reduce-scatter -> norms -> sharded update -> all-gather over NVLink peer memory (p2p.cuh)
int p2p_exchange_update(pycnn_b200_ctx* c) {
  p2p::PeerTable t{};
  for (int r = 0; r < c->world; ++r) {
    t.grads[r] = static_cast<float*>(c->p2p_peer_ptrs[0][r]);
    t.params[r] = static_cast<float*>(c->p2p_peer_ptrs[1][r]);
    t.sync[r] = static_cast<p2p::SyncBlock*>(c->p2p_peer_ptrs[2][r]);
  }
  t.rank = c->rank; t.world = c->world; t.chunk0 = c->p2p_chunk0; t.nchunks = c->p2p_nchunks;
  t.lost = c->d_blocks_done + 2;
  const int grid = std::max(1, c->p2p_nchunks);
  {
    LaunchScope ls(c, KC_ALLREDUCE);
    PB_CUDA(launch_pdl(c, p2p::p2p_reduce_kernel, dim3(grid), dim3(256), 0, c->stream, t, (const Chunk*)c->d_chunks,
                       (const long long*)c->d_p2p_seq, c->d_chunk_norm, (const int2*)c->d_tensor_chunks, (int)c->tensors.size(),
                       c->d_blocks_done, trace_slot(c, KC_ALLREDUCE)));
  }
  UpdateParams u = make_update_params(c);
  LaunchScope ls(c, KC_UPDATE);
  PB_CUDA(launch_pdl(c, p2p::p2p_update_kernel, dim3(grid), dim3(256), 0, c->stream, t, (const Chunk*)c->d_chunks,
                     c->d_p2p_seq, c->d_m, c->d_v, u, c->d_blocks_done + 1, trace_slot(c, KC_UPDATE)));
  return 0;
}

// The flat buffer is exchanged in up to two groups of chunks, each with its own counter pair and sequence word:
// group 0 = the first layer's W and b (all but ~1 % of the bytes; complete as soon as dW_1 is), group 1 = every other
// tensor (the shallow dW GEMMs finish later, on the side stream).  group < 0: one exchange over all chunks (set 0).
struct NvlsGroup { int c0, n, set; };
NvlsGroup nvls_group(const pycnn_b200_ctx* c, int group) {
  int split = c->num_chunks;
  for (int i = 0; i < c->num_chunks; ++i)
    if (c->h_chunks[i].tensor >= 2) { split = i; break; }
  if (group < 0) return {0, c->num_chunks, 0};
  return group == 0 ? NvlsGroup{0, split, 0} : NvlsGroup{split, c->num_chunks - split, 1};
}

nvls::Table nvls_table(pycnn_b200_ctx* c, const NvlsGroup& g) {
  nvls::Table t{};
  char* mc = reinterpret_cast<char*>(c->nvls_mcva);
  char* uc = reinterpret_cast<char*>(c->nvls_uc);
  t.mc_grads = reinterpret_cast<float*>(mc);
  t.mc_gred = reinterpret_cast<float*>(mc + c->nvls_off_gred);
  t.mc_norm = reinterpret_cast<double*>(mc + c->nvls_off_norm);
  t.mc_flags = reinterpret_cast<unsigned long long*>(mc + c->nvls_off_flags) + 2 * g.set;
  t.uc_flags = reinterpret_cast<const unsigned long long*>(uc + c->nvls_off_flags) + 2 * g.set;
  t.rank = c->rank; t.world = c->world;
  t.chunk0 = g.c0 + (int)((int64_t)c->rank * g.n / c->world);                    // owned range inside the group
  t.nchunks = g.c0 + (int)((int64_t)(c->rank + 1) * g.n / c->world) - t.chunk0;
  t.lost = c->d_blocks_done + 2;
  return t;
}

// all-reduce of one chunk group through the NVSwitch multicast object (nvls.cuh), on c->stream
int nvls_exchange(pycnn_b200_ctx* c, int group) {
  const NvlsGroup g = nvls_group(c, group);
  if (g.n <= 0) return 0;
  nvls::Table t = nvls_table(c, g);
  LaunchScope ls(c, KC_ALLREDUCE);
  PB_CUDA(launch_pdl(c, nvls::nvls_allreduce_kernel, dim3(std::max(1, t.nchunks)), dim3(256), 0, c->stream, t,
                     (const Chunk*)c->d_chunks, c->d_p2p_seq + g.set, c->d_blocks_done + g.set, trace_slot(c, KC_ALLREDUCE)));
  return 0;
}

// the replicated clip + update on the reduced buffer
int nvls_update(pycnn_b200_ctx* c) {
  char* uc = reinterpret_cast<char*>(c->nvls_uc);
  return apply_update(c, true, reinterpret_cast<const float*>(uc + c->nvls_off_gred),
                      reinterpret_cast<double*>(uc + c->nvls_off_norm));
}

// one exchange over all chunks, then the update (every path but the split one of step_on_batch)
int nvls_exchange_update(pycnn_b200_ctx* c) {
  if (c->nvls_fused && !c->big_tensors && c->num_chunks >= c->world) {     // exchange + clip + update in one launch
    nvls::Table t = nvls_table(c, nvls_group(c, -1));
    char* uc = reinterpret_cast<char*>(c->nvls_uc);
    UpdateParams u = make_update_params(c);
    LaunchScope ls(c, KC_ALLREDUCE);
    PB_CUDA(launch_pdl(c, nvls::nvls_allreduce_update_kernel, dim3(c->num_chunks), dim3(256), 0, c->stream, t,
                       (const Chunk*)c->d_chunks, c->num_chunks, c->d_p2p_seq, c->d_params, c->d_m, c->d_v,
                       reinterpret_cast<const float*>(uc + c->nvls_off_gred),
                       reinterpret_cast<const double*>(uc + c->nvls_off_norm), u, c->d_blocks_done,
                       trace_slot(c, KC_ALLREDUCE)));
    return 0;
  }
  PB_TRY(nvls_exchange(c, -1));
  return nvls_update(c);
}

int gather_batch(pycnn_b200_ctx* c, const int64_t* d_idx, const int64_t* cursor, int64_t extra, int bs, float* x,
                 int32_t* blabels);

// Row-gather prefetch: as soon as the layer-1 forward GEMM (the one L2/HBM-heavy kernel of the forward pass) is
// done, the NEXT step's rows are gathered into the other batch buffer on pf_stream -- beside the latency-bound MLP
// tail, softmax and dA chain -- and joined at the end of the step.  In this mode nothing on the main stream reads
// the row cursor, so the prefetch branch owns it: rows come from idx[cursor + bs ...) and the cursor moves on
// behind the gather.
int issue_prefetch(pycnn_b200_ctx* c) {
  if (!c->pf.armed || c->pf.next_bs == 0) return 0;
  c->pf.armed = false;
  const bool fork = !c->timing && c->pf_stream != nullptr;
  cudaStream_t main_stream = c->stream;
  if (fork) {
    PB_CUDA(cudaEventRecord(c->ev_pf_fork, main_stream));
    PB_CUDA(cudaStreamWaitEvent(c->pf_stream, c->ev_pf_fork, 0));
    c->stream = c->pf_stream;
  }
  int rc = gather_batch(c, c->d_idx, c->d_cursor, c->pf.extra, c->pf.next_bs, c->pf.x, c->pf.labels);
  if (!rc) {
    LaunchScope ls(c, KC_MISC);
    advance_cursor_kernel<<<1, 32, 0, c->stream>>>(c->d_cursor, c->pf.extra);
    if (cudaGetLastError() != cudaSuccess) rc = fail("advance_cursor_kernel launch failed");
  }
  c->stream = main_stream;
  PB_TRY(rc);
  c->pf.forked = fork;
  return 0;
}

int join_prefetch(pycnn_b200_ctx* c) {
  c->pf.armed = false;
  if (!c->pf.forked) return 0;
  c->pf.forked = false;
  PB_CUDA(cudaEventRecord(c->ev_pf_join, c->pf_stream));
  PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_pf_join, 0));
  return 0;
}

// full step on x [bs, Dp] + labels (device).  bs may be 0 on a rank whose slice is empty.
int step_on_batch(pycnn_b200_ctx* c, const float* x, const int32_t* labels, int bs, bool update, int64_t advance = 0) {
  begin_grads(c);
  // Multicast exchange of a network the fused tail handles: the first layer's gradient (all but ~1 % of the bytes) is
  // complete when dW_1 is, so its exchange starts right there on the main stream, while the shallow dW GEMMs, their
  // finalize and their (small) exchange run beside it on the side stream.  The same on every rank, every step.
  const bool split = update && c->nvls && c->nvls_split && !c->nvls_fused && tail_eligible(c) && nvls_group(c, 1).n > 0;
  bool side_open = false;
  if (bs > 0) {
    if (tail_eligible(c)) {
      PB_TRY(forward_logits(c, x, bs, nullptr, 0, 0));          // layer 1 only (+ the row-gather prefetch fork)
      if (c->gsrc.on) PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_pf_join, 0));     // the batch labels (side stream)
      PB_TRY(tail_forward_backward(c, bs, labels, advance));
      PB_TRY(backward_dw_after_tail(c, x, bs, !split, &side_open));
    } else if (softmax_fusable(c)) {
      if (c->gsrc.on) PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_pf_join, 0));
      PB_TRY(forward_logits(c, x, bs, labels, advance));
      PB_TRY(backward(c, x, bs));
    } else {
      if (c->gsrc.on) PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_pf_join, 0));
      PB_TRY(forward_logits(c, x, bs));
      // probs are not written during training; dZ goes to its own buffer
      PB_TRY(softmax_ce(c, bs, labels, false, true, false, true, true, advance));
      PB_TRY(backward(c, x, bs));
    }
  } else {   // empty shard on this rank: contribute zero gradients, still advance the cursor
    PB_TRY(zero_grads(c));
    LaunchScope ls(c, KC_MISC);
    StepBookkeeping bk{c->d_cursor, advance};
    step_bookkeeping_kernel<<<1, 32, 0, c->stream>>>(bk);
    PB_CUDA(cudaGetLastError());
    PB_TRY(issue_prefetch(c));
  }
  if (split) {
    const NvlsGroup g0 = nvls_group(c, 0), g1 = nvls_group(c, 1);
    cudaStream_t main_stream = c->stream;
    if (bs > 0) {
      // plain stream order for this one launch: as a programmatic dependent of dW_1 its ~550 blocks would become
      // resident early, wait, and hold the registers the shallow dW GEMMs on the side stream need (measured: +17 us)
      const bool pdl = c->use_pdl;
      c->use_pdl = false;
      const int rc = finalize_grads(c, false, g0.c0, g0.n);
      c->use_pdl = pdl;
      PB_TRY(rc);
    }
    if (pf_beside_exchange(c)) PB_TRY(issue_prefetch(c));
    PB_TRY(nvls_exchange(c, 0));
    if (side_open) c->stream = c->side_stream;      // behind the shallow dW GEMMs
    int rc = bs > 0 ? finalize_grads(c, false, g1.c0, g1.n) : 0;
    if (!rc) rc = nvls_exchange(c, 1);
    c->stream = main_stream;
    PB_TRY(rc);
    if (side_open) {
      PB_CUDA(cudaEventRecord(c->ev_join, c->side_stream));
      PB_CUDA(cudaStreamWaitEvent(main_stream, c->ev_join, 0));
    }
    return nvls_update(c);
  }
  // single GPU: the clip norms come out of the same pass that adds the partial sums; with an exchange they must be
  // those of the REDUCED gradient and are taken after it
  const bool local_norms = update && c->world <= 1 && c->update_rule == PYCNN_B200_RULE_CPU;
  PB_TRY(finalize_grads(c, local_norms));
  if (pf_beside_exchange(c)) PB_TRY(issue_prefetch(c));
  if (update) {
    if (c->nvls) {
      PB_TRY(nvls_exchange_update(c));
    } else if (c->p2p) {
      PB_TRY(p2p_exchange_update(c));
    } else {
      PB_TRY(allreduce_grads(c));
      PB_TRY(apply_update(c, local_norms));
    }
  }
  return 0;
}

// rows d_idx[*cursor + extra ...] (cursor may be null: offset 0)
int gather_batch(pycnn_b200_ctx* c, const int64_t* d_idx, const int64_t* cursor, int64_t extra, int bs,
                 float* x, int32_t* blabels) {
  if (bs == 0) return 0;
  if (!x) { x = c->d_x; blabels = c->d_blabels; }
  LaunchScope ls(c, KC_GATHER);
  const int64_t total = (int64_t)bs * (c->Dp >> 2);
  // 4 loads in flight per thread; on the prefetch stream only 2 CTAs per SM so that the critical path's GEMM CTAs
  // (one per SM, shared-memory bound) always find thread slots beside it
  const int64_t want = ceil_div(total, 256 * 4);
  const int per_sm = (c->stream == c->pf_stream) ? c->pf_ctas_per_sm : 8;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)per_sm * c->num_sms));
  PB_CUDA(launch_pdl(c, gather_rows_kernel, dim3(blocks), dim3(256), 0, c->stream, c->d_feat,
                     c->d_labels, d_idx, cursor, extra, bs, c->Dp, x, blabels, trace_slot(c, KC_GATHER)));
  return 0;
}

void drop_graphs(pycnn_b200_ctx* c) {
  for (auto& kv : c->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  c->graphs.clear();
}

// Run `body` (a sequence of launches on c->stream, no host synchronisation inside) through a cached CUDA graph.
// Falls back to eager issue when per-kernel timing is on (events inside a capture would not time anything).
template <class Body>
int run_graphed(pycnn_b200_ctx* c, std::tuple<int, int64_t, int64_t, int64_t, int> key, Body body) {
  if (!c->use_graphs || c->timing) return body();
  auto it = c->graphs.find(key);
  if (it == c->graphs.end()) {
    pycnn_b200_ctx::GraphEntry ent;
    const int64_t l0 = c->launches;
    int64_t cl0[KC_COUNT];
    for (int k = 0; k < KC_COUNT; ++k) cl0[k] = c->class_launches[k];
    PB_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = body();
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ent.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    ent.launches = c->launches - l0;
    for (int k = 0; k < KC_COUNT; ++k) ent.class_launches[k] = c->class_launches[k] - cl0[k];
    c->launches = l0;   // counted again below, per replay
    for (int k = 0; k < KC_COUNT; ++k) c->class_launches[k] = cl0[k];
    it = c->graphs.emplace(key, ent).first;
  }
  PB_CUDA(cudaGraphLaunch(it->second.exec, c->stream));
  c->launches += it->second.launches;
  for (int k = 0; k < KC_COUNT; ++k) c->class_launches[k] += it->second.class_launches[k];
  return 0;
}

enum GraphKind { GK_STEP_ROWS = 1, GK_STEP_IMAGES = 2, GK_STEP_ROWS_PF = 3, GK_STEP_ROWS_GF = 8 };

// one training step on dataset rows d_idx[*cursor + lo .. +bs), then cursor += advance
int graphed_step_rows(pycnn_b200_ctx* c, int bs, int64_t lo, int64_t advance) {
  return run_graphed(c, std::make_tuple((int)GK_STEP_ROWS, (int64_t)bs, lo, advance, 0), [&]() -> int {
    PB_TRY(gather_batch(c, c->d_idx, c->d_cursor, lo, bs, nullptr, nullptr));
    PB_TRY(step_on_batch(c, c->d_x, c->d_blabels, bs, true, advance));
    return 0;
  });
}

// Pipelined variant for runs of consecutive steps: this step's rows are already in batch buffer `parity` (gathered
// by the previous step, or by prime_rows for the first one); while it runs, next_bs rows of the following step are
// gathered into the other buffer.  One graph per (bs, next_bs, parity).
float* batch_x(pycnn_b200_ctx* c, int parity) { return parity ? c->d_x2 : c->d_x; }
int32_t* batch_labels(pycnn_b200_ctx* c, int parity) { return parity ? c->d_blabels2 : c->d_blabels; }

int prime_rows(pycnn_b200_ctx* c, int bs) {   // first batch of a run -> buffer 0 (cursor already reset)
  return gather_batch(c, c->d_idx, c->d_cursor, 0, bs, batch_x(c, 0), batch_labels(c, 0));
}

// `count` consecutive steps of bs rows in one graph (count even when > 1 so that the buffer parity comes back to where
// it started): fewer graph launches per epoch, and the PDL chain runs on across the step boundary.
int graphed_step_rows_pf(pycnn_b200_ctx* c, int bs, int next_bs, int parity, int count = 1) {
  return run_graphed(c, std::make_tuple((int)GK_STEP_ROWS_PF, (int64_t)bs, (int64_t)next_bs, (int64_t)count, parity), [&]() -> int {
    for (int j = 0; j < count; ++j) {
      const int p = parity ^ (j & 1);
      c->pf.armed = true; c->pf.forked = false; c->pf.next_bs = (j + 1 < count) ? bs : next_bs; c->pf.extra = bs;
      c->pf.x = batch_x(c, p ^ 1); c->pf.labels = batch_labels(c, p ^ 1);
      const int rc = step_on_batch(c, batch_x(c, p), batch_labels(c, p), bs, true, 0);
      const int rj = join_prefetch(c);
      if (rc || rj) return rc ? rc : rj;
    }
    return 0;
  });
}

// Gather-free steps (the default for batches large enough for the 2-SM layer-1 GEMMs): `count` consecutive steps of bs
// rows in one graph.  No batch buffer, no gather kernel, no prefetch stream: the layer-1 GEMMs fetch rows
// idx[cursor ...) of the dataset themselves (TMA gather4), only the bs labels are gathered -- on the side stream, beside
// the layer-1 forward GEMM.  The step's bookkeeping moves the cursor on by bs.
bool gather_free_ok(pycnn_b200_ctx* c, int bs) {
  if (!c->use_gather_free || c->math_mode != PYCNN_B200_MATH_TF32 || bs <= 0 || !c->side_stream) return false;
  GemmCall f{};
  f.a_kmajor = true; f.b_kmajor = true; f.M = bs; f.N = (int)c->dims[1]; f.K = (int)c->dims[0]; f.epi = 2;
  GemmCall w{};
  w.a_kmajor = false; w.b_kmajor = false; w.M = (int)c->dims[1]; w.N = (int)c->dims[0]; w.K = bs; w.grad_slot = 1;
  return gather_ok(c, f) && gather_ok(c, w);
}

int graphed_step_rows_gf(pycnn_b200_ctx* c, int bs, int count) {
  return run_graphed(c, std::make_tuple((int)GK_STEP_ROWS_GF, (int64_t)bs, (int64_t)count, (int64_t)0, 0), [&]() -> int {
    int rc = 0;
    for (int j = 0; j < count && !rc; ++j) {
      // labels of this batch: side stream, joined before the first kernel that reads them
      PB_CUDA(cudaEventRecord(c->ev_pf_fork, c->stream));
      PB_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_pf_fork, 0));
      {
        LaunchScope ls(c, KC_GATHER);
        gather_labels_kernel<<<(unsigned)std::min<int64_t>(ceil_div(bs, 256), 64), 256, 0, c->side_stream>>>(
            c->d_labels, c->d_idx, c->d_cursor, bs, c->d_blabels);
        PB_CUDA(cudaGetLastError());
      }
      PB_CUDA(cudaEventRecord(c->ev_pf_join, c->side_stream));
      c->gsrc.on = true; c->gsrc.advance = bs;
      rc = step_on_batch(c, nullptr, c->d_blabels, bs, true, bs);
      c->gsrc.on = false;
    }
    return rc;
  });
}

int reset_cursor(pycnn_b200_ctx* c) {
  PB_CUDA(cudaMemsetAsync(c->d_cursor, 0, sizeof(int64_t), c->stream));   // [0] only: [1] keeps counting steps
  return 0;
}

int zero_stats(pycnn_b200_ctx* c) {
  PB_CUDA(cudaMemsetAsync(c->d_stats, 0, kStatsWords * sizeof(unsigned long long), c->stream));
  return 0;
}

// Data-parallel failure detection, polled wherever the host already synchronises: an asynchronous NCCL
// error (a dead peer aborts the communicator) and the peer-memory exchange's lost-peer mark become error returns.
int comm_health(pycnn_b200_ctx* c) {
  if (c->world <= 1) return 0;
  if (c->nccl_comm) {
    NcclApi* api = nccl_api();
    int async_err = 0;
    if (api && api->CommGetAsyncError && api->CommGetAsyncError(c->nccl_comm, &async_err) == 0 && async_err != 0)
      return fail("NCCL reported an asynchronous error on rank %d: %s", c->rank, api->GetErrorString ? api->GetErrorString(async_err) : "?");
  }
  if ((c->p2p || c->nvls) && c->d_blocks_done) {
    unsigned int lost = 0;
    PB_CUDA(cudaMemcpy(&lost, c->d_blocks_done + 2, sizeof(lost), cudaMemcpyDeviceToHost));
    if (lost > (unsigned)c->world)
      return fail("multicast gradient exchange: not every rank arrived within the wait bound (seen by rank %d); the parameters "
                  "of this step are invalid -- destroy the communicator and restart from a checkpoint", c->rank);
    if (lost) return fail("peer-memory gradient exchange: rank %u did not arrive within the wait bound (seen by rank %d); "
                          "the parameters of this step are invalid -- destroy the communicator and restart from a checkpoint", lost - 1, c->rank);
  }
  return 0;
}

// fixed-point loss total -> double (NaN when any contribution was non-finite, as the reference's float sum would be)
double stats_loss(const unsigned long long* w) {
  return w[2] ? std::nan("") : (double)(long long)w[0] / kLossFixedScale;
}

int read_stats(pycnn_b200_ctx* c, double* loss_sum, int64_t* correct, bool allreduce) {
  if (allreduce && c->world > 1 && c->nccl_comm) {
    NcclApi* api = nccl_api();
    PB_NCCL(api, api->AllReduce(c->d_stats, c->d_stats, kStatsWords, kNcclInt64, kNcclSum, c->nccl_comm, c->stream));
  }
  PB_CUDA(cudaMemcpyAsync(c->h_stats, c->d_stats, kStatsWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  PB_TRY(comm_health(c));
  if (loss_sum) *loss_sum = stats_loss(c->h_stats);
  if (correct) *correct = (int64_t)c->h_stats[1];
  return 0;
}

int check_idx_host(const pycnn_b200_ctx* c, const int64_t* idx, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    PB_CHECK(idx[i] >= 0 && idx[i] < c->n_rows, "row index %lld out of range [0,%lld)", (long long)idx[i], (long long)c->n_rows);
  return 0;
}

// host f64 [rows, cols] -> device f32 [rows, ld] via scratch
int upload_f64(pycnn_b200_ctx* c, const double* src, int64_t rows, int64_t cols, float* dst, int64_t ld) {
  if (rows * cols == 0) return 0;
  const size_t bytes = sizeof(double) * (size_t)(rows * cols);
  PB_TRY(ensure_scratch(c, bytes));
  PB_CUDA(cudaMemcpyAsync(c->d_scratch, src, bytes, cudaMemcpyHostToDevice, c->stream));
  f64_to_f32_padded_kernel<<<grid_for(rows * ld, 256, c->num_sms), 256, 0, c->stream>>>(
      static_cast<const double*>(c->d_scratch), rows, cols, dst, ld);
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int download_f64(pycnn_b200_ctx* c, const float* src, int64_t rows, int64_t cols, int64_t ld, double* dst) {
  if (rows * cols == 0) return 0;
  const size_t bytes = sizeof(double) * (size_t)(rows * cols);
  PB_TRY(ensure_scratch(c, bytes));
  f32_padded_to_f64_kernel<<<grid_for(rows * cols, 256, c->num_sms), 256, 0, c->stream>>>(
      src, rows, cols, ld, static_cast<double*>(c->d_scratch));
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaMemcpyAsync(dst, c->d_scratch, bytes, cudaMemcpyDeviceToHost, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// Leave the multicast exchange: the gradient buffer goes back to an ordinary allocation.
int nvls_detach(pycnn_b200_ctx* c) {
  if (!c->nvls_mc && !c->nvls_mem && !c->nvls_uc && !c->nvls_mcva) { c->nvls = false; return 0; }
  nvls::Driver* d = nvls::driver();
  if (c->stream) cudaStreamSynchronize(c->stream);
  const bool grads_inside = c->nvls_uc && c->d_grads == reinterpret_cast<float*>(c->nvls_uc);
  if (d) {
    int dev = 0; cudaGetDevice(&dev);
    CUdevice cudev = 0; d->DeviceGet(&cudev, dev);
    if (c->nvls_mcva) { d->MemUnmap(c->nvls_mcva, c->nvls_size); d->MemAddressFree(c->nvls_mcva, c->nvls_size); }
    if (c->nvls_bound) d->MulticastUnbind(c->nvls_mc, cudev, 0, c->nvls_size);
    if (c->nvls_uc) { d->MemUnmap(c->nvls_uc, c->nvls_size); d->MemAddressFree(c->nvls_uc, c->nvls_size); }
    if (c->nvls_mem) d->MemRelease(c->nvls_mem);
    if (c->nvls_mc) d->MemRelease(c->nvls_mc);
  }
  c->nvls_mcva = c->nvls_uc = 0; c->nvls_mem = c->nvls_mc = 0; c->nvls_bound = c->nvls_added = false; c->nvls = false;
  if (grads_inside) {
    drop_graphs(c);                      // captured launches hold the address inside the region
    c->d_grads = nullptr;
    if (c->P > 0) PB_TRY(alloc_f32(&c->d_grads, c->P));
  }
  free_dev(c->d_p2p_seq); c->d_p2p_seq = nullptr;
  free_dev(c->d_blocks_done); c->d_blocks_done = nullptr;
  return 0;
}

void p2p_detach(pycnn_b200_ctx* c) {
  nvls_detach(c);
  if (!c->p2p_sync && !c->p2p) return;
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int k = 0; k < 3; ++k)
    for (int r = 0; r < 8; ++r) {
      if (c->p2p_peer_ptrs[k][r] && r != c->rank) cudaIpcCloseMemHandle(c->p2p_peer_ptrs[k][r]);
      c->p2p_peer_ptrs[k][r] = nullptr;
    }
  free_dev(c->p2p_sync); c->p2p_sync = nullptr;
  free_dev(c->d_p2p_seq); c->d_p2p_seq = nullptr;
  free_dev(c->d_blocks_done); c->d_blocks_done = nullptr;
  c->p2p = false;
}

void free_network(pycnn_b200_ctx* c) {
  p2p_detach(c);
  free_workspace(c);
  free_dev(c->d_params); free_dev(c->d_grads); free_dev(c->d_m); free_dev(c->d_v);
  free_dev(c->d_chunks); free_dev(c->d_chunk_norm); free_dev(c->d_tensor_chunks); c->d_tensor_chunks = nullptr;
  free_dev(c->d_tensor_norm); c->d_tensor_norm = nullptr;
  c->d_params = c->d_grads = c->d_m = c->d_v = nullptr;
  c->d_chunks = nullptr; c->d_chunk_norm = nullptr;
  c->tensors.clear(); c->dims.clear(); c->ldw.clear();
  c->P = 0;
}

void free_dataset(pycnn_b200_ctx* c) {
  free_dev(c->d_feat); free_dev(c->d_labels);
  c->d_feat = nullptr; c->d_labels = nullptr;
  c->cap_rows = c->n_rows = 0;
}

#define CTX_GUARD(c)                                            \
  PB_CHECK((c) != nullptr, "null context");                     \
  PB_CUDA(cudaSetDevice((c)->device))

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int pycnn_b200_abi_version(void) { return PYCNN_B200_ABI_VERSION; }
const char* pycnn_b200_last_error(void) { return g_err; }

int pycnn_b200_device_count(int* count) {
  PB_CHECK(count, "null argument");
  *count = 0;
  PB_CUDA(cudaGetDeviceCount(count));
  return 0;
}

int pycnn_b200_create(int device, pycnn_b200_ctx** out) {
  PB_CHECK(out, "null argument");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail("no CUDA device available (%s); pycnn_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  PB_CHECK(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  PB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, device));
  PB_CHECK(prop.major == 10, "pycnn_b200 kernels are built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
  auto* c = new pycnn_b200_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->l2_bytes = (size_t)prop.l2CacheSize;
  auto bail = [&](int r) { pycnn_b200_destroy(c); return r; };
#define CREATE_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { fail("%s failed: %s", #expr, cudaGetErrorString(_e)); return bail(1); } } while (0)
  {
    // the critical path (main stream) gets the highest priority: when CTAs of several streams are pending, its
    // CTAs are placed first; the prefetch gather and the shallow dW GEMMs take what is left
    int prio_lo = 0, prio_hi = 0;
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    const char* pe = getenv("PYCNN_B200_PRIORITIES");
    const bool prio = !(pe && pe[0] == '0');
    CREATE_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio ? prio_hi : prio_lo));
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio_lo));
    CREATE_CUDA(cudaStreamCreateWithPriority(&c->pf_stream, cudaStreamNonBlocking, prio_lo));
  }
  CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_pf_fork, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_pf_join, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_idx_free, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_idx_ready, cudaEventDisableTiming));
  for (auto& ev : c->ev_fork) CREATE_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (auto& ev : c->events) CREATE_CUDA(cudaEventCreate(&ev));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
  CREATE_CUDA(cudaEventCreateWithFlags(&c->compute_done, cudaEventDisableTiming));
  CREATE_CUDA(cudaMalloc(&c->d_stats, kStatsWords * sizeof(unsigned long long)));
  CREATE_CUDA(cudaMemset(c->d_stats, 0, kStatsWords * sizeof(unsigned long long)));
  CREATE_CUDA(cudaMallocHost(&c->h_stats, kStatsWords * sizeof(unsigned long long)));
  // shared workspace of split-K GEMMs with an epilogue whose slices cannot meet in distributed shared memory:
  // tiles x slices never exceeds the SM count and a slice of a tile is at most 128 x 256 floats
  c->gemm_ws_bytes = sizeof(float) * (size_t)(c->num_sms + 4) * 128 * 256;
  CREATE_CUDA(cudaMalloc(&c->d_gemm_ws, c->gemm_ws_bytes));
  CREATE_CUDA(cudaMalloc(&c->d_cursor, 2 * sizeof(int64_t)));
  CREATE_CUDA(cudaMemset(c->d_cursor, 0, 2 * sizeof(int64_t)));
#undef CREATE_CUDA
  { const char* g = getenv("PYCNN_B200_GRAPHS"); if (g && g[0] == '0') c->use_graphs = false; }
  { const char* g = getenv("PYCNN_B200_PDL"); if (g && g[0] == '0') c->use_pdl = false; }
  { const char* g = getenv("PYCNN_B200_FORK"); if (g && g[0] == '0') { cudaStreamDestroy(c->side_stream); c->side_stream = nullptr; } }
  { const char* g = getenv("PYCNN_B200_CARVEOUT"); if (g && g[0] == '0') c->uniform_carveout = false; }
  { const char* g = getenv("PYCNN_B200_FORCE_BN"); if (g && (atoi(g) == 64 || atoi(g) == 128 || atoi(g) == 256)) c->force_bn = atoi(g); }
  { const char* g = getenv("PYCNN_B200_DEFER_DW"); if (g && g[0] == '0') c->defer_last_dw = false; }
  { const char* g = getenv("PYCNN_B200_GRAPH_STEPS"); if (g && atoi(g) > 0) c->graph_steps = atoi(g); }
  { const char* g = getenv("PYCNN_B200_EPOCH_LOG"); if (g) c->epoch_log = atoi(g) != 0; }
  { const char* g = getenv("PYCNN_B200_NVLS_SPLIT"); if (g) c->nvls_split = atoi(g) != 0; }
  { const char* g = getenv("PYCNN_B200_NVLS_FUSED"); if (g) c->nvls_fused = atoi(g) != 0; }
  { const char* g = getenv("PYCNN_B200_PF_CTAS"); if (g && atoi(g) > 0) c->pf_ctas_per_sm = atoi(g); }
  { const char* g = getenv("PYCNN_B200_PREFETCH"); if (g && g[0] == '0') c->use_prefetch = false; }
  { const char* g = getenv("PYCNN_B200_PF_EARLY"); if (g) c->pf_early = (g[0] == '1'); }
  { const char* g = getenv("PYCNN_B200_PF_PLACE"); if (g) c->pf_place = !strcmp(g, "fwd") ? 1 : !strcmp(g, "dw") ? 2 : !strcmp(g, "xchg") ? 3 : 0; }
  { const char* g = getenv("PYCNN_B200_FUSE_SOFTMAX"); if (g && g[0] == '0') c->fuse_softmax = false; }
  { const char* g = getenv("PYCNN_B200_ONE_WAVE"); if (g && g[0] == '0') c->one_wave_tiles = false; }
  { const char* g = getenv("PYCNN_B200_PAIR"); if (g && g[0] == '0') c->use_pair = false; }
  { const char* g = getenv("PYCNN_B200_GATHER_FREE"); if (g) c->use_gather_free = (g[0] == '1'); }
  { const char* g = getenv("PYCNN_B200_NVTX"); if (g && g[0] == '1') c->nvtx = true; }
  { const char* g = getenv("PYCNN_B200_FUSED_TAIL"); if (g && g[0] == '0') c->use_tail = false; }
  { cudaError_t e = cudaFuncSetAttribute(tail::tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tail::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tail::tail_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { fail("cudaFuncSetAttribute(tail_fused_kernel) failed: %s", cudaGetErrorString(e)); return bail(1); } }
  { const char* g = getenv("PYCNN_B200_TWO_SM"); if (g && g[0] == '0') c->use_two_sm = false; }
  { const char* g = getenv("PYCNN_B200_PAIR_MIN_KB"); if (g && atoi(g) > 0) c->pair_min_kb = atoi(g); }
  { const char* g = getenv("PYCNN_B200_CLUSTER_SPLITK"); if (g && g[0] == '0') c->use_cluster_splitk = false; }
  if (c->uniform_carveout) {
    // same shared-memory carve-out as the GEMM kernels (see prepare_tc_gemm) for everything that runs beside them
    cudaError_t e = cudaSuccess;
    auto prefer = [&](const void* f) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    };
    prefer((const void*)gather_rows_kernel); prefer((const void*)softmax_ce_kernel); prefer((const void*)grad_sqnorm_kernel);
    prefer((const void*)clip_update_kernel); prefer((const void*)splitk_reduce_kernel); prefer((const void*)grad_finalize_kernel); prefer((const void*)step_bookkeeping_kernel);
    prefer((const void*)colsum_kernel); prefer((const void*)p2p::p2p_reduce_kernel); prefer((const void*)p2p::p2p_update_kernel);
    if (e != cudaSuccess) { fail("cudaFuncSetAttribute(carve-out) failed: %s", cudaGetErrorString(e)); return bail(1); }
  }
  if (tc_gemm_prepare(c) != 0 || tc_extract_prepare(c) != 0) return bail(1);
  if (resolve_encode_tiled(c) != 0) return bail(1);
  *out = c;
  return 0;
}

int pycnn_b200_destroy(pycnn_b200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drop_graphs(c);
  pycnn_b200_comm_destroy(c);
  free_dev(c->d_cursor);
  free_dev(c->d_trace);
  free_network(c);
  free_dataset(c);
  free_dev(c->d_taps); free_dev(c->d_tc_filters); free_dev(c->d_tc_scale);
  free_dev(c->d_idx); free_dev(c->d_images); free_dev(c->d_img_labels);
  for (int s = 0; s < 2; ++s) {
    free_dev(c->d_img2[s]); free_dev(c->d_lab2[s]);
    if (c->ev_copy[s]) cudaEventDestroy(c->ev_copy[s]);
    if (c->ev_compute[s]) cudaEventDestroy(c->ev_compute[s]);
    if (c->h_bounce[s]) cudaFreeHost(c->h_bounce[s]);
  }
  if (c->h_step_stats) cudaFreeHost(c->h_step_stats);
  free_dev(c->d_scratch); free_dev(c->d_flush); free_dev(c->d_gemm_ws); free_dev(c->d_stats);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->h_stats) cudaFreeHost(c->h_stats);
  for (int k = 0; k < KC_COUNT; ++k)
    for (auto& t : c->timed[k]) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : c->event_pool) cudaEventDestroy(e);
  for (auto& ev : c->events) if (ev) cudaEventDestroy(ev);
  if (c->copy_done) cudaEventDestroy(c->copy_done);
  if (c->compute_done) cudaEventDestroy(c->compute_done);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->pf_stream) cudaStreamDestroy(c->pf_stream);
  if (c->ev_pf_fork) cudaEventDestroy(c->ev_pf_fork);
  if (c->ev_pf_join) cudaEventDestroy(c->ev_pf_join);
  if (c->ev_idx_free) cudaEventDestroy(c->ev_idx_free);
  if (c->ev_idx_ready) cudaEventDestroy(c->ev_idx_ready);
  for (auto& ev : c->ev_fork) if (ev) cudaEventDestroy(ev);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int pycnn_b200_synchronize(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int pycnn_b200_set_math_mode(pycnn_b200_ctx* c, int mode) {
  CTX_GUARD(c);
  PB_CHECK(mode == PYCNN_B200_MATH_FP32_SIMT || mode == PYCNN_B200_MATH_TF32, "unknown math mode %d", mode);
  drop_graphs(c);
  c->math_mode = mode;
  return 0;
}
int pycnn_b200_set_update_rule(pycnn_b200_ctx* c, int rule) {
  CTX_GUARD(c);
  PB_CHECK(rule == PYCNN_B200_RULE_CPU || rule == PYCNN_B200_RULE_CUPY, "unknown update rule %d", rule);
  drop_graphs(c);
  c->update_rule = rule;
  return 0;
}
int pycnn_b200_set_extract_mode(pycnn_b200_ctx* c, int mode) {
  CTX_GUARD(c);
  PB_CHECK(mode == PYCNN_B200_EXTRACT_SIMT || mode == PYCNN_B200_EXTRACT_TC, "unknown extract mode %d", mode);
  drop_graphs(c);
  c->extract_mode = mode;
  return 0;
}

int pycnn_b200_device_info(pycnn_b200_ctx* c, char* buf, size_t buflen) {
  CTX_GUARD(c);
  PB_CHECK(buf && buflen > 0, "null buffer");
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, c->device));
  // co-resident GEMM clusters (128x256 forward tile): plain 2/4/8-CTA split-K clusters and paired (x2) ones
  const auto& occ = g_max_clusters[3][0];
  snprintf(buf, buflen, "%s sm_%d%d, %d SMs, %.1f GiB, L2 %.0f MiB; max resident GEMM clusters 1x{2,4,8}=%d/%d/%d 2x{1,2,4}=%d/%d/%d",
           prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.totalGlobalMem / 1073741824.0,
           prop.l2CacheSize / 1048576.0, occ[0][1], occ[0][2], occ[0][3], occ[1][0], occ[1][1], occ[1][2]);
  return 0;
}

// ---- extractor ---------------------------------------------------------------------------------
int pycnn_b200_set_filters(pycnn_b200_ctx* c, const double* taps, int F) {
  CTX_GUARD(c);
  PB_CHECK(taps && F > 0, "set_filters needs at least one 3x3 filter");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  c->h_taps.assign(taps, taps + (size_t)F * 9);
  std::vector<float> flipped((size_t)F * 9);
  for (int f = 0; f < F; ++f)
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)  // true convolution: x[i+a][j+b] * k[2-a][2-b]  
        flipped[(size_t)f * 9 + a * 3 + b] = (float)taps[(size_t)f * 9 + (2 - a) * 3 + (2 - b)];
  free_dev(c->d_taps);
  c->d_taps = nullptr;
  PB_CUDA(cudaMalloc(&c->d_taps, sizeof(float) * flipped.size()));
  PB_CUDA(cudaMemcpy(c->d_taps, flipped.data(), sizeof(float) * flipped.size(), cudaMemcpyHostToDevice));
  c->F = F;
  PB_TRY(tc_extract_set_filters(c));
  return 0;
}

int pycnn_b200_feature_count(pycnn_b200_ctx* c, int image_size, int64_t* out) {
  PB_CHECK(c && out, "null argument");
  *out = feature_count(c, image_size);
  return 0;
}

static int extract_host(pycnn_b200_ctx* c, const uint8_t* images, int64_t n, int S, double* out64, float* out32) {
  CTX_GUARD(c);
  PB_CHECK(n >= 0 && (n == 0 || images), "null images");
  PB_CHECK(c->F > 0, "set_filters must be called before extraction");
  PB_CHECK(S >= 4, "image_size must be >= 4 (got %d)", S);
  if (n == 0) return 0;
  const int64_t D = feature_count(c, S);
  const int64_t ld = round_up(D, 4);
  const size_t img_bytes = (size_t)S * S * 3;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(512ull << 20) / (ld * 4)));
  float* d_out = nullptr;
  PB_CUDA(cudaMalloc(&d_out, sizeof(float) * (size_t)(chunk * ld + kSlackFloats)));
  int rc = 0;
  for (int64_t i0 = 0; i0 < n && rc == 0; i0 += chunk) {
    const int64_t nb = std::min(chunk, n - i0);
    rc = ensure_images(c, nb * img_bytes, 0);
    if (rc) break;
    cudaError_t e = cudaMemcpyAsync(c->d_images, images + i0 * img_bytes, nb * img_bytes, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) { rc = fail("H2D copy failed: %s", cudaGetErrorString(e)); break; }
    rc = extract_device(c, c->d_images, nb, S, d_out, ld);
    if (rc) break;
    if (out64) rc = download_f64(c, d_out, nb, D, ld, out64 + i0 * D);
    else {
      rc = ensure_scratch(c, sizeof(float) * (size_t)(nb * D));
      if (rc) break;
      f32_padded_to_f32_kernel<<<grid_for(nb * D, 256, c->num_sms), 256, 0, c->stream>>>(d_out, nb, D, ld, static_cast<float*>(c->d_scratch));
      e = cudaMemcpyAsync(out32 + i0 * D, c->d_scratch, sizeof(float) * (size_t)(nb * D), cudaMemcpyDeviceToHost, c->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) rc = fail("D2H copy failed: %s", cudaGetErrorString(e));
    }
  }
  cudaFree(d_out);
  return rc;
}

int pycnn_b200_extract(pycnn_b200_ctx* c, const uint8_t* images, int64_t n, int S, double* out) {
  PB_CHECK(out || n == 0, "null output");
  return extract_host(c, images, n, S, out, nullptr);
}
int pycnn_b200_extract_f32(pycnn_b200_ctx* c, const uint8_t* images, int64_t n, int S, float* out) {
  PB_CHECK(out || n == 0, "null output");
  return extract_host(c, images, n, S, nullptr, out);
}

int pycnn_b200_max_pool(pycnn_b200_ctx* c, const double* map, int h, int w, double* out) {
  CTX_GUARD(c);
  PB_CHECK(map && out && h >= 0 && w >= 0, "bad arguments");
  const int nh = h / 2, nw = w / 2;
  if (nh * nw == 0) return 0;
  const size_t in_b = sizeof(double) * (size_t)h * w, out_b = sizeof(double) * (size_t)nh * nw;
  PB_TRY(ensure_scratch(c, in_b + out_b));
  double* d_in = static_cast<double*>(c->d_scratch);
  double* d_out = d_in + (size_t)h * w;
  PB_CUDA(cudaMemcpyAsync(d_in, map, in_b, cudaMemcpyHostToDevice, c->stream));
  { LaunchScope ls(c, KC_MISC);
    max_pool_f64_kernel<<<(unsigned)ceil_div(nh * nw, 256), 256, 0, c->stream>>>(d_in, h, w, d_out); }
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int pycnn_b200_softmax(pycnn_b200_ctx* c, const double* x, double* out, int rows, int cols) {
  CTX_GUARD(c);
  PB_CHECK(x && out && rows >= 0 && cols > 0, "bad arguments");
  if (rows == 0) return 0;
  const size_t b = sizeof(double) * (size_t)rows * cols;
  PB_TRY(ensure_scratch(c, 2 * b));
  double* d_in = static_cast<double*>(c->d_scratch);
  double* d_out = d_in + (size_t)rows * cols;
  PB_CUDA(cudaMemcpyAsync(d_in, x, b, cudaMemcpyHostToDevice, c->stream));
  { LaunchScope ls(c, KC_SOFTMAX_CE);
    softmax_f64_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, c->stream>>>(d_in, d_out, rows, cols); }
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaMemcpyAsync(out, d_out, b, cudaMemcpyDeviceToHost, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int pycnn_b200_cross_entropy(pycnn_b200_ctx* c, const double* preds, const double* labels, double* out, int rows,
                             int cols) {
  CTX_GUARD(c);
  PB_CHECK(preds && labels && out && rows >= 0 && cols > 0, "bad arguments");
  if (rows == 0) return 0;
  const size_t b = sizeof(double) * (size_t)rows * cols;
  PB_TRY(ensure_scratch(c, 2 * b + sizeof(double) * rows));
  double* d_p = static_cast<double*>(c->d_scratch);
  double* d_y = d_p + (size_t)rows * cols;
  double* d_o = d_y + (size_t)rows * cols;
  PB_CUDA(cudaMemcpyAsync(d_p, preds, b, cudaMemcpyHostToDevice, c->stream));
  PB_CUDA(cudaMemcpyAsync(d_y, labels, b, cudaMemcpyHostToDevice, c->stream));
  { LaunchScope ls(c, KC_SOFTMAX_CE);
    cross_entropy_f64_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, c->stream>>>(d_p, d_y, d_o, rows, cols); }
  PB_CUDA(cudaGetLastError());
  PB_CUDA(cudaMemcpyAsync(out, d_o, sizeof(double) * rows, cudaMemcpyDeviceToHost, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- network -----------------------------------------------------------------------------------
int pycnn_b200_set_network(pycnn_b200_ctx* c, int64_t D, const int* layers, int L, int C) {
  CTX_GUARD(c);
  PB_CHECK(D > 0 && L >= 0 && C > 0 && (L == 0 || layers), "bad network shape");
  PB_CHECK(2 * (L + 1) <= GradParts::MAX_T, "pycnn_b200 supports at most %d hidden layers (got %d)", GradParts::MAX_T / 2 - 1, L);
  for (int l = 0; l < L; ++l) PB_CHECK(layers[l] > 0, "layer %d has size %d", l, layers[l]);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  if (c->D != D) free_dataset(c);
  free_network(c);
  c->D = D; c->Dp = round_up(D, 4); c->L = L; c->C = C;
  c->dims.push_back(D);
  for (int l = 0; l < L; ++l) c->dims.push_back(layers[l]);
  c->dims.push_back(C);
  for (auto d : c->dims) c->ldw.push_back(round_up(d, 4));
  int64_t off = 0;
  for (int l = 0; l <= L; ++l) {
    TensorDesc w{off, c->dims[l + 1], c->dims[l], c->ldw[l]};
    off = round_up(off + w.rows * w.ld, 64);
    TensorDesc b{off, 1, c->dims[l + 1], c->ldw[l + 1]};
    off = round_up(off + b.ld, 64);
    c->tensors.push_back(w);
    c->tensors.push_back(b);
  }
  c->P = off;
  PB_TRY(alloc_f32(&c->d_params, c->P));
  PB_TRY(alloc_f32(&c->d_grads, c->P));
  PB_TRY(alloc_f32(&c->d_m, c->P));
  PB_TRY(alloc_f32(&c->d_v, c->P));
  std::vector<Chunk> chunks;
  // one 256-thread block per chunk: 2048 floats for the big tensors (>= 4 waves of blocks at 1.1 M parameters), short
  // chunks for the small ones, which can collect many ordered partials (grad_finalize_kernel)
  for (size_t t = 0; t < c->tensors.size(); ++t) {
    const int64_t len = c->tensors[t].rows * c->tensors[t].ld;
    const int64_t kChunk = len <= 300 * 1024 ? kSmallChunk : 2048;
    const int first = (int)chunks.size(), count = (int)ceil_div(len, kChunk);
    for (int64_t o = 0; o < len; o += kChunk)
      chunks.push_back({(int)t, first, count, 0, c->tensors[t].off + o, std::min(kChunk, len - o), o});
  }
  c->num_chunks = (int)chunks.size();
  PB_CUDA(cudaMalloc(&c->d_chunks, sizeof(Chunk) * chunks.size()));
  PB_CUDA(cudaMemcpy(c->d_chunks, chunks.data(), sizeof(Chunk) * chunks.size(), cudaMemcpyHostToDevice));
  std::vector<int2> tchunks(c->tensors.size());
  for (const Chunk& ch : chunks) tchunks[ch.tensor] = make_int2(ch.first_chunk, ch.n_chunks);
  PB_CUDA(cudaMalloc(&c->d_tensor_chunks, sizeof(int2) * tchunks.size()));
  PB_CUDA(cudaMemcpy(c->d_tensor_chunks, tchunks.data(), sizeof(int2) * tchunks.size(), cudaMemcpyHostToDevice));
  c->h_chunks = chunks;
  c->big_tensors = false;
  for (const int2& tcn : tchunks) c->big_tensors = c->big_tensors || tcn.y > 2048;
  PB_CUDA(cudaMalloc(&c->d_tensor_norm, sizeof(double) * tchunks.size()));
  PB_CUDA(cudaMalloc(&c->d_chunk_norm, sizeof(double) * chunks.size()));
  PB_CUDA(cudaMemset(c->d_chunk_norm, 0, sizeof(double) * chunks.size()));
  c->t = 0;
  return 0;
}

int pycnn_b200_num_params(pycnn_b200_ctx* c, int* count) {
  PB_CHECK(c && count, "null argument");
  *count = (int)c->tensors.size();
  return 0;
}

int pycnn_b200_param_shape(pycnn_b200_ctx* c, int index, int64_t* rows, int64_t* cols) {
  PB_CHECK(c && rows && cols, "null argument");
  PB_CHECK(index >= 0 && index < (int)c->tensors.size(), "param index %d out of range", index);
  *rows = c->tensors[index].rows;
  *cols = c->tensors[index].cols;
  return 0;
}

static int param_io(pycnn_b200_ctx* c, float* base, int index, const double* in, double* out) {
  CTX_GUARD(c);
  PB_CHECK(base, "set_network must be called first");
  PB_CHECK(index >= 0 && index < (int)c->tensors.size(), "param index %d out of range", index);
  const TensorDesc& t = c->tensors[index];
  if (in) return upload_f64(c, in, t.rows, t.cols, base + t.off, t.ld);
  return download_f64(c, base + t.off, t.rows, t.cols, t.ld, out);
}
int pycnn_b200_set_param(pycnn_b200_ctx* c, int index, const double* v) { PB_CHECK(c && v, "null argument"); return param_io(c, c->d_params, index, v, nullptr); }
int pycnn_b200_get_param(pycnn_b200_ctx* c, int index, double* v) { PB_CHECK(c && v, "null argument"); return param_io(c, c->d_params, index, nullptr, v); }
int pycnn_b200_get_grad(pycnn_b200_ctx* c, int index, double* v) { PB_CHECK(c && v, "null argument"); return param_io(c, c->d_grads, index, nullptr, v); }
int pycnn_b200_get_opt_state(pycnn_b200_ctx* c, int which, int index, double* v) {
  PB_CHECK(c && v && (which == 0 || which == 1), "bad arguments");
  return param_io(c, which == 0 ? c->d_m : c->d_v, index, nullptr, v);
}
int pycnn_b200_set_opt_state(pycnn_b200_ctx* c, int which, int index, const double* v) {
  PB_CHECK(c && v && (which == 0 || which == 1), "bad arguments");
  return param_io(c, which == 0 ? c->d_m : c->d_v, index, v, nullptr);
}

int pycnn_b200_set_optimizer(pycnn_b200_ctx* c, int kind, double lr, double beta1, double beta2, double eps) {
  CTX_GUARD(c);
  PB_CHECK(kind == PYCNN_B200_OPT_SGD || kind == PYCNN_B200_OPT_ADAM, "unknown optimizer %d", kind);
  drop_graphs(c);   // lr / betas are baked into the captured update kernel
  c->opt = kind; c->lr = lr; c->beta1 = beta1; c->beta2 = beta2; c->eps = eps; c->t = 0;
  PB_CUDA(cudaMemsetAsync(c->d_cursor, 0, 2 * sizeof(int64_t), c->stream));
  if (c->d_m) {
    PB_CUDA(cudaMemsetAsync(c->d_m, 0, sizeof(float) * c->P, c->stream));
    PB_CUDA(cudaMemsetAsync(c->d_v, 0, sizeof(float) * c->P, c->stream));
  }
  return 0;
}

int pycnn_b200_get_step_count(pycnn_b200_ctx* c, int64_t* t) {
  PB_CHECK(c && t, "null argument");
  *t = c->t;
  return 0;
}

int pycnn_b200_set_step_count(pycnn_b200_ctx* c, int64_t t) {
  CTX_GUARD(c);
  PB_CHECK(t >= 0, "step count must be >= 0 (got %lld)", (long long)t);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  PB_CUDA(cudaMemcpy(c->d_cursor + 1, &t, sizeof(int64_t), cudaMemcpyHostToDevice));
  c->t = t;
  return 0;
}

// ---- dataset -----------------------------------------------------------------------------------
int pycnn_b200_dataset_alloc(pycnn_b200_ctx* c, int64_t rows) {
  CTX_GUARD(c);
  PB_CHECK(c->D > 0, "set_network must be called first");
  PB_CHECK(rows > 0, "dataset needs at least one row");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  free_dataset(c);
  PB_TRY(alloc_f32(&c->d_feat, rows * c->Dp));
  PB_CUDA(cudaMalloc(&c->d_labels, sizeof(int32_t) * rows));
  PB_CUDA(cudaMemset(c->d_labels, 0, sizeof(int32_t) * rows));
  c->cap_rows = rows;
  c->n_rows = 0;
  return 0;
}

static int dataset_range(pycnn_b200_ctx* c, int64_t row0, int64_t n) {
  PB_CHECK(c->d_feat, "dataset_alloc must be called first");
  PB_CHECK(row0 >= 0 && n >= 0 && row0 + n <= c->cap_rows, "rows [%lld,%lld) exceed dataset capacity %lld", (long long)row0, (long long)(row0 + n), (long long)c->cap_rows);
  return 0;
}

int pycnn_b200_dataset_extract(pycnn_b200_ctx* c, int64_t row0, const uint8_t* images, int64_t n, int S) {
  CTX_GUARD(c);
  PB_TRY(dataset_range(c, row0, n));
  PB_CHECK(feature_count(c, S) == c->D, "image_size %d with %d filters gives %lld features, network expects %lld", S, c->F, (long long)feature_count(c, S), (long long)c->D);
  if (n == 0) return 0;
  PB_CHECK(images, "null images");
  const size_t img_bytes = (size_t)S * S * 3;
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(256ull << 20) / (int64_t)img_bytes);
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t nb = std::min(chunk, n - i0);
    PB_TRY(ensure_images(c, nb * img_bytes, 0));
    PB_CUDA(cudaMemcpyAsync(c->d_images, images + i0 * img_bytes, nb * img_bytes, cudaMemcpyHostToDevice, c->stream));
    PB_TRY(extract_device(c, c->d_images, nb, S, c->d_feat + (row0 + i0) * c->Dp, c->Dp));
    PB_CUDA(cudaStreamSynchronize(c->stream));  // d_images is reused by the next chunk
  }
  c->n_rows = std::max(c->n_rows, row0 + n);
  return 0;
}

int pycnn_b200_dataset_set_features(pycnn_b200_ctx* c, int64_t row0, const double* feats, int64_t n) {
  CTX_GUARD(c);
  PB_TRY(dataset_range(c, row0, n));
  if (n == 0) return 0;
  PB_CHECK(feats, "null features");
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(256ull << 20) / (c->D * 8));
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t nb = std::min(chunk, n - i0);
    PB_TRY(upload_f64(c, feats + i0 * c->D, nb, c->D, c->d_feat + (row0 + i0) * c->Dp, c->Dp));
  }
  c->n_rows = std::max(c->n_rows, row0 + n);
  return 0;
}

int pycnn_b200_dataset_get_features(pycnn_b200_ctx* c, int64_t row0, double* feats, int64_t n) {
  CTX_GUARD(c);
  PB_TRY(dataset_range(c, row0, n));
  if (n == 0) return 0;
  PB_CHECK(feats, "null features");
  const int64_t chunk = std::max<int64_t>(1, (int64_t)(256ull << 20) / (c->D * 8));
  for (int64_t i0 = 0; i0 < n; i0 += chunk) {
    const int64_t nb = std::min(chunk, n - i0);
    PB_TRY(download_f64(c, c->d_feat + (row0 + i0) * c->Dp, nb, c->D, c->Dp, feats + i0 * c->D));
  }
  return 0;
}

int pycnn_b200_dataset_set_labels(pycnn_b200_ctx* c, int64_t row0, const int32_t* labels, int64_t n) {
  CTX_GUARD(c);
  PB_TRY(dataset_range(c, row0, n));
  if (n == 0) return 0;
  PB_CHECK(labels, "null labels");
  for (int64_t i = 0; i < n; ++i) PB_CHECK(labels[i] >= 0 && labels[i] < c->C, "label %d at row %lld outside [0,%d)", labels[i], (long long)(row0 + i), c->C);
  PB_CUDA(cudaMemcpyAsync(c->d_labels + row0, labels, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int pycnn_b200_dataset_rows(pycnn_b200_ctx* c, int64_t* rows) {
  PB_CHECK(c && rows, "null argument");
  *rows = c->n_rows;
  return 0;
}

// ---- training ----------------------------------------------------------------------------------
int pycnn_b200_train_step(pycnn_b200_ctx* c, const int64_t* idx, int bs, double* loss_sum, int64_t* correct) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(bs > 0 && idx, "empty batch");
  PB_TRY(check_idx_host(c, idx, bs));
  PB_TRY(ensure_workspace(c, bs));
  PB_TRY(ensure_idx(c, bs));
  PB_CUDA(cudaMemcpyAsync(c->d_idx, idx, sizeof(int64_t) * bs, cudaMemcpyHostToDevice, c->stream));
  PB_TRY(zero_stats(c));
  PB_TRY(reset_cursor(c));
  if (gather_free_ok(c, bs)) PB_TRY(graphed_step_rows_gf(c, bs, 1));
  else                       PB_TRY(graphed_step_rows(c, bs, 0, bs));
  if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += 1;
  if (loss_sum || correct) PB_TRY(read_stats(c, loss_sum, correct, true));
  return 0;
}

int pycnn_b200_train_steps(pycnn_b200_ctx* c, const int64_t* idx, int bs, int steps, double* loss_sum, int64_t* correct) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(bs > 0 && steps > 0 && idx, "empty batch");
  PB_TRY(check_idx_host(c, idx, (int64_t)bs * steps));
  PB_TRY(ensure_workspace(c, bs));
  PB_TRY(ensure_idx(c, (int64_t)bs * steps));
  PB_CUDA(cudaMemcpyAsync(c->d_idx, idx, sizeof(int64_t) * bs * steps, cudaMemcpyHostToDevice, c->stream));
  PB_TRY(zero_stats(c));
  PB_TRY(reset_cursor(c));
  if (gather_free_ok(c, bs)) {
    for (int s = 0; s < steps;) {
      int count = 1;
      while (2 * count <= c->graph_steps && s + 2 * count <= steps) count *= 2;
      PB_TRY(graphed_step_rows_gf(c, bs, count));
      s += count;
    }
  } else if (c->use_prefetch && steps > 1) {
    PB_TRY(ensure_prefetch_buffers(c));
    PB_TRY(prime_rows(c, bs));
    for (int s = 0; s < steps;) {   // greedy 16/8/4/2/1-step graphs, as in train_epoch
      int count = 1;
      if ((s & 1) == 0)
        while (2 * count <= c->graph_steps && s + 2 * count <= steps) count *= 2;
      PB_TRY(graphed_step_rows_pf(c, bs, s + count < steps ? bs : 0, s & 1, count));
      s += count;
    }
  } else {
    for (int s = 0; s < steps; ++s) PB_TRY(graphed_step_rows(c, bs, 0, bs));
  }
  if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += steps;
  if (loss_sum || correct) PB_TRY(read_stats(c, loss_sum, correct, true));
  return 0;
}

int pycnn_b200_train_epoch(pycnn_b200_ctx* c, const int64_t* perm, int64_t n, int batch_size, double* loss_sum,
                           int64_t* correct) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(n > 0 && batch_size > 0 && perm, "empty epoch");
  const int W = c->world, R = c->rank;
  const int64_t per_rank_max = ceil_div(std::min<int64_t>(batch_size, n), W);
  PB_TRY(ensure_workspace(c, per_rank_max));
  // Host side of the epoch's indices: THIS rank's slice of every batch is compacted into pinned staging (memcpy) and
  // validated with a branch-free reduction the compiler vectorises -- ~0.1 ms per 100 k indices, which would leave the
  // GPU idle if done up front; it is done in instalments instead (need_batches below).
  const auto host_t0 = std::chrono::steady_clock::now();
  const int64_t num_batches = ceil_div(n, batch_size);
  PB_TRY(ensure_pinned(c, sizeof(int64_t) * (size_t)(num_batches * per_rank_max)));
  int64_t* h_idx = static_cast<int64_t*>(c->h_pinned);
  int64_t k = 0;
  const uint64_t limit = (uint64_t)c->n_rows;
  auto pack = [&](int64_t batch0, int64_t batch1) -> int {   // batches [batch0, batch1) -> h_idx[k...)
    const int64_t k0 = k;
    uint64_t bad = 0;
    for (int64_t bi = batch0; bi < batch1; ++bi) {             // pycnn/pycnn.py:626-627
      const int64_t start = bi * batch_size;
      const int64_t end = std::min<int64_t>(start + batch_size, n);
      const int64_t len = end - start, per = ceil_div(len, W);
      const int64_t lo = std::min(len, per * R), hi = std::min(len, per * (R + 1));
      const int64_t* src = perm + start + lo;
      const int64_t cnt = hi - lo;
      memcpy(h_idx + k, src, sizeof(int64_t) * (size_t)cnt);
      uint64_t b = 0;
      for (int64_t i = 0; i < cnt; ++i) b |= (uint64_t)((uint64_t)src[i] >= limit);
      bad |= b;
      k += cnt;
    }
    if (bad)                                                   // slow path only to name the offender
      for (int64_t i = k0; i < k; ++i)
        PB_CHECK(h_idx[i] >= 0 && h_idx[i] < c->n_rows, "row index %lld out of range [0,%lld)", (long long)h_idx[i], (long long)c->n_rows);
    return 0;
  };
  PB_TRY(ensure_idx(c, std::max<int64_t>(num_batches * per_rank_max, 1)));
  // indices go up in instalments, each just ahead of the graph that first needs it: batches [0, upto) are made resident.
  // The first instalment travels on the compute stream (in front of the first graph); later ones are packed while the
  // GPU works and travel on the copy stream, which must not overwrite what the previous epoch's last steps still read.
  int64_t packed_batches = 0;
  bool copy_stream_armed = false;
  auto need_batches = [&](int64_t upto) -> int {
    upto = std::min(upto, num_batches);
    if (upto <= packed_batches) return 0;
    const int64_t k0 = k;
    PB_TRY(pack(packed_batches, upto));        // a bad index ends the epoch here, behind the steps already trained --
    packed_batches = upto;                     // like the reference, whose IndexError comes from the batch that holds it
    if (k == k0) return 0;
    if (k0 == 0) {
      PB_CUDA(cudaMemcpyAsync(c->d_idx, h_idx, sizeof(int64_t) * (size_t)k, cudaMemcpyHostToDevice, c->stream));
      PB_CUDA(cudaEventRecord(c->ev_idx_free, c->stream));
      return 0;
    }
    if (!copy_stream_armed) { PB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_idx_free, 0)); copy_stream_armed = true; }
    PB_CUDA(cudaMemcpyAsync(c->d_idx + k0, h_idx + k0, sizeof(int64_t) * (size_t)(k - k0), cudaMemcpyHostToDevice, c->copy_stream));
    PB_CUDA(cudaEventRecord(c->ev_idx_ready, c->copy_stream));
    PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_idx_ready, 0));
    return 0;
  };
  // a short first graph (2 steps) gets the GPU going after packing three batches; the long graphs queue up behind it
  const int first_graph_steps = num_batches > 3 ? 2 : c->graph_steps;
  PB_TRY(need_batches(std::min<int64_t>(first_graph_steps, c->graph_steps) + 1));
  const auto host_t1 = std::chrono::steady_clock::now();
  PB_TRY(zero_stats(c));
  PB_TRY(reset_cursor(c));
  auto local_bs = [&](int64_t start) -> int {
    if (start >= n) return 0;
    const int64_t end = std::min<int64_t>(start + batch_size, n);
    const int64_t len = end - start, per = ceil_div(len, W);
    return (int)(std::min(len, per * (R + 1)) - std::min(len, per * R));
  };
  // gather-free when every batch of this rank is large enough for the 2-SM layer-1 GEMMs (they fetch their rows from the
  // dataset themselves); otherwise the gathered batch buffer, double-buffered behind the previous step when it pays
  bool gf = true;
  for (int64_t start = 0; start < n && gf; start += batch_size)
    if (local_bs(start) > 0) gf = gather_free_ok(c, local_bs(start));
  const bool pipelined = !gf && c->use_prefetch && num_batches > 1;
  if (pipelined) {
    PB_TRY(ensure_prefetch_buffers(c));
    PB_TRY(prime_rows(c, local_bs(0)));
  }
  int parity = 0;
  for (int64_t start = 0; start < n;) {
    const int bs = local_bs(start);
    // one graph per distinct local batch size (x next size x buffer parity when the row gather is pipelined);
    // runs of full batches go graph_steps at a time
    int count = 1;
    if ((pipelined || gf) && parity == 0 && bs > 0) {
      // largest power-of-two run (<= graph_steps) of full batches that starts here: an epoch decomposes greedily
      // into 16, 8, 4, 2, 1-step graphs, so only a handful of graphs are ever instantiated
      const int cap = start == 0 ? std::min(first_graph_steps, c->graph_steps) : c->graph_steps;
      for (int U = 1; 2 * U <= cap; ) {
        U *= 2;
        bool uniform = start + (int64_t)U * batch_size <= n;
        for (int j = count; uniform && j < U; ++j) uniform = local_bs(start + j * (int64_t)batch_size) == bs;
        if (!uniform) break;
        count = U;
      }
    }
    PB_TRY(need_batches(start / batch_size + count + 1));   // this graph's steps + the rows it prefetches for the next one
    if (gf && bs > 0)   PB_TRY(graphed_step_rows_gf(c, bs, count));
    else if (pipelined) PB_TRY(graphed_step_rows_pf(c, bs, local_bs(start + count * (int64_t)batch_size), parity, count));
    else                PB_TRY(graphed_step_rows(c, bs, 0, bs));
    if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += count;
    start += count * (int64_t)batch_size;
    if (!gf && (count & 1)) parity ^= 1;
  }
  const auto host_t2 = std::chrono::steady_clock::now();
  PB_TRY(read_stats(c, loss_sum, correct, true));
  if (c->epoch_log) {
    const auto host_t3 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() / 1e3; };
    fprintf(stderr, "[pycnn_b200 rank %d] train_epoch %lld steps: pack+upload %.1f us, enqueue %.1f us, drain+stats %.1f us\n", c->rank,
            (long long)num_batches, us(host_t0, host_t1), us(host_t1, host_t2), us(host_t2, host_t3));
  }
  return 0;
}

int pycnn_b200_train_step_images(pycnn_b200_ctx* c, const uint8_t* images, const int32_t* labels, int n, int S,
                                 double* loss_sum, int64_t* correct) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_CHECK(n > 0 && images && labels, "empty batch");
  PB_CHECK(feature_count(c, S) == c->D, "image_size %d with %d filters gives %lld features, network expects %lld", S, c->F, (long long)feature_count(c, S), (long long)c->D);
  PB_TRY(ensure_workspace(c, n));
  const size_t img_bytes = (size_t)S * S * 3 * n;
  PB_TRY(ensure_images(c, img_bytes, n));
  PB_CUDA(cudaMemcpyAsync(c->d_images, images, img_bytes, cudaMemcpyHostToDevice, c->stream));
  PB_CUDA(cudaMemcpyAsync(c->d_blabels, labels, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
  PB_TRY(zero_stats(c));
  PB_TRY(run_graphed(c, std::make_tuple((int)GK_STEP_IMAGES, (int64_t)n, (int64_t)S, (int64_t)0, 0), [&]() -> int {
    PB_TRY(extract_device(c, c->d_images, n, S, c->d_x, c->Dp));
    PB_TRY(step_on_batch(c, c->d_x, c->d_blabels, n, true));
    return 0;
  }));
  if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += 1;
  if (loss_sum || correct) PB_TRY(read_stats(c, loss_sum, correct, true));
  return 0;
}

// two staging slots (uint8 images + labels) + their events for copy-stream || compute-stream pipelines
static int ensure_stream_slots(pycnn_b200_ctx* c, int64_t image_bytes, int64_t num_labels) {
  const size_t need_img = (size_t)round_up(image_bytes + 256, 1 << 20), need_lab = (size_t)round_up(std::max<int64_t>(num_labels, 1), 1024);
  if (need_img <= c->img2_cap && need_lab <= c->lab2_cap) return 0;
  PB_CUDA(cudaStreamSynchronize(c->stream));
  PB_CUDA(cudaStreamSynchronize(c->copy_stream));
  drop_graphs(c);
  const size_t img_cap = std::max(need_img, c->img2_cap), lab_cap = std::max(need_lab, c->lab2_cap);
  for (int s = 0; s < 2; ++s) {
    free_dev(c->d_img2[s]); free_dev(c->d_lab2[s]);
    c->d_img2[s] = nullptr; c->d_lab2[s] = nullptr;
    PB_CUDA(cudaMalloc(&c->d_img2[s], img_cap));
    PB_CUDA(cudaMemset(c->d_img2[s], 0, img_cap));
    PB_CUDA(cudaMalloc(&c->d_lab2[s], sizeof(int32_t) * lab_cap));
    if (!c->ev_copy[s]) PB_CUDA(cudaEventCreateWithFlags(&c->ev_copy[s], cudaEventDisableTiming));
    if (!c->ev_compute[s]) PB_CUDA(cudaEventCreateWithFlags(&c->ev_compute[s], cudaEventDisableTiming));
  }
  c->img2_cap = img_cap; c->lab2_cap = lab_cap;
  return 0;
}

static bool host_ptr_is_device_visible(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
}

int pycnn_b200_train_epoch_images(pycnn_b200_ctx* c, const uint8_t* images, const int32_t* labels, int64_t num_images,
                                  const int64_t* perm, int64_t n, int S, int batch_size, double* step_stats,
                                  double* loss_sum, int64_t* correct) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_CHECK(n > 0 && batch_size > 0 && images && labels, "empty epoch");
  PB_CHECK(feature_count(c, S) == c->D, "image_size %d with %d filters gives %lld features, network expects %lld", S, c->F, (long long)feature_count(c, S), (long long)c->D);
  PB_CHECK(num_images > 0 && (perm || n <= num_images), "epoch of %lld rows over %lld images", (long long)n, (long long)num_images);
  if (perm)
    for (int64_t i = 0; i < n; ++i)
      PB_CHECK(perm[i] >= 0 && perm[i] < num_images, "row index %lld at position %lld out of range [0,%lld)", (long long)perm[i], (long long)i, (long long)num_images);
  const int W = c->world, R = c->rank;
  const int64_t img_bytes = (int64_t)S * S * 3;
  const int64_t per_rank_max = ceil_div(std::min<int64_t>(batch_size, n), W);
  const int64_t num_steps = ceil_div(n, batch_size);
  PB_TRY(ensure_workspace(c, per_rank_max));
  PB_TRY(ensure_stream_slots(c, per_rank_max * img_bytes, per_rank_max));
  if ((size_t)(kStatsWords * num_steps) > c->step_stats_cap) {
    if (c->h_step_stats) PB_CUDA(cudaFreeHost(c->h_step_stats));
    c->h_step_stats = nullptr;
    c->step_stats_cap = (size_t)round_up(kStatsWords * num_steps, 1024);
    PB_CUDA(cudaMallocHost(&c->h_step_stats, sizeof(unsigned long long) * c->step_stats_cap));
  }
  const bool zero_copy = perm && host_ptr_is_device_visible(images) && host_ptr_is_device_visible(labels);
  const bool bounce = perm && !zero_copy;
  if (zero_copy) {
    PB_TRY(ensure_idx(c, n));
    PB_CUDA(cudaMemcpyAsync(c->d_idx, perm, sizeof(int64_t) * n, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(cudaStreamSynchronize(c->stream));   // the copy stream reads d_idx
  }
  if (bounce) {
    const size_t need = (size_t)round_up(per_rank_max * (img_bytes + 4), 1 << 20);
    if (need > c->bounce_cap) {
      for (int s = 0; s < 2; ++s) {
        if (c->h_bounce[s]) PB_CUDA(cudaFreeHost(c->h_bounce[s]));
        c->h_bounce[s] = nullptr;
        PB_CUDA(cudaMallocHost(&c->h_bounce[s], need));
      }
      c->bounce_cap = need;
    }
  }
  PB_TRY(zero_stats(c));
  int64_t step = 0;
  for (int64_t start = 0; start < n; start += batch_size, ++step) {
    const int64_t end = std::min<int64_t>(start + batch_size, n);
    const int64_t len = end - start, per = ceil_div(len, W);
    const int64_t lo = std::min(len, per * R), hi = std::min(len, per * (R + 1));
    const int bs = (int)(hi - lo);
    const int slot = (int)(step & 1);
    // ---- copy stream: bring this rank's slice of the batch into staging slot `slot`
    if (step >= 2) PB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_compute[slot], 0));
    if (bs > 0) {
      if (!perm) {
        PB_CUDA(cudaMemcpyAsync(c->d_img2[slot], images + (start + lo) * img_bytes, (size_t)bs * img_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PB_CUDA(cudaMemcpyAsync(c->d_lab2[slot], labels + start + lo, sizeof(int32_t) * bs, cudaMemcpyHostToDevice, c->copy_stream));
      } else if (zero_copy) {
        const int64_t total = (int64_t)bs * (img_bytes >> 4);
        const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, 256 * 4), 2 * c->num_sms));
        gather_host_rows_kernel<<<blocks, 256, 0, c->copy_stream>>>(images, labels, c->d_idx + start + lo, bs, img_bytes,
                                                                    c->d_img2[slot], c->d_lab2[slot]);
        PB_CUDA(cudaGetLastError());
        c->launches += 1;
        c->class_launches[KC_GATHER] += 1;
      } else {
        if (step >= 2) PB_CUDA(cudaEventSynchronize(c->ev_copy[slot]));   // previous H2D out of this bounce buffer is done
        uint8_t* hb = c->h_bounce[slot];
        int32_t* hl = reinterpret_cast<int32_t*>(hb + (size_t)per_rank_max * img_bytes);
        for (int i = 0; i < bs; ++i) {
          const int64_t src = perm[start + lo + i];
          memcpy(hb + (size_t)i * img_bytes, images + src * img_bytes, (size_t)img_bytes);
          hl[i] = labels[src];
        }
        PB_CUDA(cudaMemcpyAsync(c->d_img2[slot], hb, (size_t)bs * img_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        PB_CUDA(cudaMemcpyAsync(c->d_lab2[slot], hl, sizeof(int32_t) * bs, cudaMemcpyHostToDevice, c->copy_stream));
      }
    }
    PB_CUDA(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    // ---- compute stream: extractor + step on that slot (one cached graph per slot and batch size)
    PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[slot], 0));
    PB_TRY(run_graphed(c, std::make_tuple((int)GK_STEP_IMAGES + 1 + slot, (int64_t)bs, (int64_t)S, (int64_t)0, 0), [&]() -> int {
      if (bs > 0) PB_TRY(extract_device(c, c->d_img2[slot], bs, S, c->d_x, c->Dp));
      PB_TRY(step_on_batch(c, c->d_x, c->d_lab2[slot], bs, true));
      return 0;
    }));
    if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += 1;
    PB_CUDA(cudaMemcpyAsync(c->h_step_stats + kStatsWords * step, c->d_stats, kStatsWords * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(cudaEventRecord(c->ev_compute[slot], c->stream));
  }
  PB_TRY(read_stats(c, loss_sum, correct, true));
  if (step_stats)
    for (int64_t i = 0; i < num_steps; ++i) {
      step_stats[2 * i] = stats_loss(c->h_step_stats + kStatsWords * i);
      step_stats[2 * i + 1] = (double)c->h_step_stats[kStatsWords * i + 1];
    }
  return 0;
}

int pycnn_b200_forward_backward(pycnn_b200_ctx* c, const int64_t* idx, int bs) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(bs >= 0 && (bs == 0 || idx), "bad batch");
  if (bs > 0) {
    PB_TRY(check_idx_host(c, idx, bs));
    PB_TRY(ensure_workspace(c, bs));
    PB_TRY(ensure_idx(c, bs));
    PB_CUDA(cudaMemcpyAsync(c->d_idx, idx, sizeof(int64_t) * bs, cudaMemcpyHostToDevice, c->stream));
    PB_TRY(gather_batch(c, c->d_idx, nullptr, 0, bs, nullptr, nullptr));
  }
  PB_TRY(step_on_batch(c, c->d_x, c->d_blabels, bs, false));
  PB_CUDA(cudaStreamSynchronize(c->stream));  // the external all-reduce runs on someone else's stream
  return 0;
}

int pycnn_b200_apply_update(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_TRY(apply_update(c));
  if (c->update_rule == PYCNN_B200_RULE_CUPY && c->opt == PYCNN_B200_OPT_ADAM) c->t += 1;
  return 0;
}

int pycnn_b200_read_step_stats(pycnn_b200_ctx* c, double* loss_sum, int64_t* correct, int reset) {
  CTX_GUARD(c);
  PB_TRY(read_stats(c, loss_sum, correct, false));
  if (reset) PB_TRY(zero_stats(c));
  return 0;
}

// ---- inference ---------------------------------------------------------------------------------
static int probs_to_host(pycnn_b200_ctx* c, int bs, double* probs) {
  PB_TRY(softmax_ce(c, bs, nullptr, true, false, false, false));
  return download_f64(c, c->d_probs, bs, c->C, c->ldw[c->L + 1], probs);
}

int pycnn_b200_forward_rows(pycnn_b200_ctx* c, const int64_t* idx, int64_t row0, int bs, double* probs) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(bs > 0 && probs, "bad arguments");
  PB_TRY(ensure_workspace(c, bs));
  const float* x;
  if (idx) {
    PB_TRY(check_idx_host(c, idx, bs));
    PB_TRY(ensure_idx(c, bs));
    PB_CUDA(cudaMemcpyAsync(c->d_idx, idx, sizeof(int64_t) * bs, cudaMemcpyHostToDevice, c->stream));
    PB_TRY(gather_batch(c, c->d_idx, nullptr, 0, bs, nullptr, nullptr));
    x = c->d_x;
  } else {
    PB_CHECK(row0 >= 0 && row0 + bs <= c->n_rows, "rows [%lld,%lld) outside the dataset", (long long)row0, (long long)(row0 + bs));
    x = c->d_feat + row0 * c->Dp;
  }
  PB_TRY(forward_logits(c, x, bs));
  return probs_to_host(c, bs, probs);
}

int pycnn_b200_forward_features(pycnn_b200_ctx* c, const double* feats, int bs, double* probs) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_CHECK(bs > 0 && feats && probs, "bad arguments");
  PB_TRY(ensure_workspace(c, bs));
  PB_TRY(upload_f64(c, feats, bs, c->D, c->d_x, c->Dp));
  PB_TRY(forward_logits(c, c->d_x, bs));
  return probs_to_host(c, bs, probs);
}

int pycnn_b200_predict_images(pycnn_b200_ctx* c, const uint8_t* images, int64_t n, int S, int32_t* classes,
                              float* confidences) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_CHECK(n >= 0 && (n == 0 || (images && classes)), "bad arguments");
  PB_CHECK(feature_count(c, S) == c->D, "image_size %d with %d filters gives %lld features, network expects %lld", S, c->F, (long long)feature_count(c, S), (long long)c->D);
  const size_t img_bytes = (size_t)S * S * 3;
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(16384, (int64_t)(1ull << 30) / (c->Dp * 4)));
  if (n == 0) return 0;
  PB_TRY(ensure_workspace(c, std::min<int64_t>(chunk, n)));
  PB_TRY(ensure_stream_slots(c, std::min<int64_t>(chunk, n) * (int64_t)img_bytes, 1));
  // double-buffered: the H2D copy of chunk i+1 (copy stream) is enqueued BEFORE chunk i's compute and result
  // read-back (which blocks the host for pageable destinations), so it overlaps both
  const int64_t nchunks = ceil_div(n, chunk);
  auto enqueue_copy = [&](int64_t ci) -> int {
    const int64_t i0 = ci * chunk;
    const int nb = (int)std::min(chunk, n - i0);
    const int slot = (int)(ci & 1);
    if (ci >= 2) PB_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_compute[slot], 0));
    PB_CUDA(cudaMemcpyAsync(c->d_img2[slot], images + i0 * img_bytes, nb * img_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    PB_CUDA(cudaEventRecord(c->ev_copy[slot], c->copy_stream));
    return 0;
  };
  PB_TRY(enqueue_copy(0));
  for (int64_t ci = 0; ci < nchunks; ++ci) {
    const int64_t i0 = ci * chunk;
    const int nb = (int)std::min(chunk, n - i0);
    const int slot = (int)(ci & 1);
    PB_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy[slot], 0));
    PB_TRY(extract_device(c, c->d_img2[slot], nb, S, c->d_x, c->Dp));
    PB_CUDA(cudaEventRecord(c->ev_compute[slot], c->stream));   // the staging slot is free once the extractor has read it
    if (ci + 1 < nchunks) PB_TRY(enqueue_copy(ci + 1));
    PB_TRY(forward_logits(c, c->d_x, nb));
    PB_TRY(softmax_ce(c, nb, nullptr, false, false, true, false));
    PB_CUDA(cudaMemcpyAsync(classes + i0, c->d_pred, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, c->stream));
    if (confidences) PB_CUDA(cudaMemcpyAsync(confidences + i0, c->d_conf, sizeof(float) * nb, cudaMemcpyDeviceToHost, c->stream));
  }
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int pycnn_b200_evaluate_rows(pycnn_b200_ctx* c, int64_t row0, int64_t n, int batch_size, double* loss_sum,
                             int64_t* correct, int32_t* predicted) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_feat, "network and dataset must be set first");
  PB_CHECK(n > 0 && batch_size > 0 && row0 >= 0 && row0 + n <= c->n_rows, "rows [%lld,%lld) outside the dataset", (long long)row0, (long long)(row0 + n));
  PB_TRY(ensure_workspace(c, std::min<int64_t>(batch_size, n)));
  PB_TRY(zero_stats(c));
  for (int64_t s = 0; s < n; s += batch_size) {   // pycnn/pycnn.py:888-898
    const int bs = (int)std::min<int64_t>(batch_size, n - s);
    PB_TRY(forward_logits(c, c->d_feat + (row0 + s) * c->Dp, bs));
    PB_TRY(softmax_ce(c, bs, c->d_labels + row0 + s, false, false, true, true));
    if (predicted) {
      PB_CUDA(cudaMemcpyAsync(predicted + s, c->d_pred, sizeof(int32_t) * bs, cudaMemcpyDeviceToHost, c->stream));
      PB_CUDA(cudaStreamSynchronize(c->stream));
    }
  }
  return read_stats(c, loss_sum, correct, false);
}

int pycnn_b200_evaluate_images(pycnn_b200_ctx* c, const uint8_t* images, const int32_t* labels, int64_t n, int S,
                               int batch_size, double* loss_sum, int64_t* correct, int32_t* predicted) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params, "set_network must be called first");
  PB_CHECK(n > 0 && batch_size > 0 && images && labels, "bad arguments");
  PB_CHECK(feature_count(c, S) == c->D, "image_size %d with %d filters gives %lld features, network expects %lld", S, c->F, (long long)feature_count(c, S), (long long)c->D);
  for (int64_t i = 0; i < n; ++i) PB_CHECK(labels[i] >= 0 && labels[i] < c->C, "label %d at %lld outside [0,%d)", labels[i], (long long)i, c->C);
  const size_t img_bytes = (size_t)S * S * 3;
  const int64_t chunk = std::min<int64_t>(batch_size, n);
  PB_TRY(ensure_workspace(c, chunk));
  PB_TRY(zero_stats(c));
  for (int64_t s = 0; s < n; s += chunk) {
    const int bs = (int)std::min<int64_t>(chunk, n - s);
    PB_TRY(ensure_images(c, bs * img_bytes, bs));
    PB_CUDA(cudaMemcpyAsync(c->d_images, images + s * img_bytes, bs * img_bytes, cudaMemcpyHostToDevice, c->stream));
    PB_CUDA(cudaMemcpyAsync(c->d_img_labels, labels + s, sizeof(int32_t) * bs, cudaMemcpyHostToDevice, c->stream));
    PB_TRY(extract_device(c, c->d_images, bs, S, c->d_x, c->Dp));
    PB_TRY(forward_logits(c, c->d_x, bs));
    PB_TRY(softmax_ce(c, bs, c->d_img_labels, false, false, true, true));
    if (predicted) PB_CUDA(cudaMemcpyAsync(predicted + s, c->d_pred, sizeof(int32_t) * bs, cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(cudaStreamSynchronize(c->stream));  // staging buffers are reused by the next chunk
  }
  return read_stats(c, loss_sum, correct, false);
}

int pycnn_b200_backward_last(pycnn_b200_ctx* c, const int32_t* labels, int bs) {
  CTX_GUARD(c);
  PB_CHECK(c->d_params && c->d_x, "forward_features must be called first");
  PB_CHECK(bs > 0 && bs <= c->bs_cap && labels, "bad arguments");
  for (int i = 0; i < bs; ++i) PB_CHECK(labels[i] >= 0 && labels[i] < c->C, "label %d outside [0,%d)", labels[i], c->C);
  PB_CUDA(cudaMemcpyAsync(c->d_blabels, labels, sizeof(int32_t) * bs, cudaMemcpyHostToDevice, c->stream));
  // d_probs holds probabilities after forward_features; softmax is idempotent on logits only, so redo the
  // forward to regenerate logits from the activations' inputs (cheap at the sizes this compat path sees)
  begin_grads(c);
  PB_TRY(forward_logits(c, c->d_x, bs));
  PB_TRY(zero_stats(c));
  PB_TRY(softmax_ce(c, bs, c->d_blabels, false, true, false, true, true, 0));
  PB_TRY(backward(c, c->d_x, bs));
  PB_TRY(finalize_grads(c, false));
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- data parallel -----------------------------------------------------------------------------
int pycnn_b200_comm_unique_id(char* id128) {
  PB_CHECK(id128, "null argument");
  NcclApi* api = nccl_api();
  PB_CHECK(api, "could not load libnccl.so.2 (set PYCNN_B200_NCCL_LIB)");
  NcclUniqueId id;
  PB_NCCL(api, api->GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}

int pycnn_b200_comm_init(pycnn_b200_ctx* c, const char* id128, int rank, int world) {
  CTX_GUARD(c);
  PB_CHECK(id128 && world >= 1 && rank >= 0 && rank < world, "bad communicator arguments");
  NcclApi* api = nccl_api();
  PB_CHECK(api, "could not load libnccl.so.2 (set PYCNN_B200_NCCL_LIB)");
  PB_TRY(pycnn_b200_comm_destroy(c));
  drop_graphs(c);
  NcclUniqueId id;
  memcpy(id.internal, id128, 128);
  NcclComm comm = nullptr;
  PB_NCCL(api, api->CommInitRank(&comm, world, id, rank));
  c->nccl_comm = comm;
  c->rank = rank;
  c->world = world;
  return 0;
}

int pycnn_b200_comm_destroy(pycnn_b200_ctx* c) {
  PB_CHECK(c, "null context");
  p2p_detach(c);
  if (c->nccl_comm) {
    if (c->stream) cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    NcclApi* api = nccl_api();
    // the stream is drained; ncclCommAbort releases the communicator without waiting for the peers (ncclCommDestroy can
    // block for ever when a peer has already died -- e.g. while an exception unwinds on one rank)
    if (api) { if (api->CommAbort) api->CommAbort(c->nccl_comm); else api->CommDestroy(c->nccl_comm); }
    c->nccl_comm = nullptr;
  }
  c->rank = 0;
  c->world = 1;
  return 0;
}

int pycnn_b200_comm_ipc_export(pycnn_b200_ctx* c, char* blob) {
  CTX_GUARD(c);
  PB_CHECK(blob, "null argument");
  PB_CHECK(c->d_params && c->d_grads, "set_network must be called first");
  PB_CHECK(c->tensors.size() <= (size_t)p2p::MAX_TENSORS, "too many tensors for the peer-memory exchange");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  p2p_detach(c);
  // fresh sync state: zeroed BEFORE the handles travel, so no peer can write into it earlier
  PB_CUDA(cudaMalloc(&c->p2p_sync, sizeof(p2p::SyncBlock)));
  PB_CUDA(cudaMemset(c->p2p_sync, 0, sizeof(p2p::SyncBlock)));
  PB_CUDA(cudaMalloc(&c->d_p2p_seq, 2 * sizeof(long long)));
  PB_CUDA(cudaMalloc(&c->d_blocks_done, 4 * sizeof(unsigned int)));      // [0], [1] block counters, [2] lost-peer mark
  PB_CUDA(cudaMemset(c->d_blocks_done, 0, 4 * sizeof(unsigned int)));
  const long long init[2] = {0, 0};
  PB_CUDA(cudaMemcpy(c->d_p2p_seq, init, sizeof(init), cudaMemcpyHostToDevice));
  cudaIpcMemHandle_t h[3];
  PB_CUDA(cudaIpcGetMemHandle(&h[0], c->d_grads));
  PB_CUDA(cudaIpcGetMemHandle(&h[1], c->d_params));
  PB_CUDA(cudaIpcGetMemHandle(&h[2], c->p2p_sync));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(blob, h, sizeof(h));
  return 0;
}

int pycnn_b200_comm_ipc_attach(pycnn_b200_ctx* c, const char* blobs, int rank, int world) {
  CTX_GUARD(c);
  PB_CHECK(blobs && world >= 1 && world <= p2p::MAX_RANKS && rank >= 0 && rank < world, "bad arguments");
  PB_CHECK(c->p2p_sync, "comm_ipc_export must be called first");
  PB_CHECK(c->world == world && c->rank == rank, "comm_init(rank, world) must match the peer-memory attach");
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->p2p_peer_ptrs[0][r] = c->d_grads; c->p2p_peer_ptrs[1][r] = c->d_params; c->p2p_peer_ptrs[2][r] = c->p2p_sync;
      continue;
    }
    cudaIpcMemHandle_t h[3];
    memcpy(h, blobs + (size_t)r * PYCNN_B200_IPC_BLOB_BYTES, sizeof(h));
    for (int k = 0; k < 3; ++k) {
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h[k], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        p2p_detach(c);
        return fail("cudaIpcOpenMemHandle(rank %d, buffer %d) failed: %s", r, k, cudaGetErrorString(e));
      }
      c->p2p_peer_ptrs[k][r] = ptr;
    }
  }
  const int nc = c->num_chunks;
  c->p2p_chunk0 = (int)((int64_t)rank * nc / world);
  c->p2p_nchunks = (int)((int64_t)(rank + 1) * nc / world) - c->p2p_chunk0;
  c->p2p = true;
  return 0;
}

// Collective (every rank): make Adam's m, v complete on every rank.  In the peer-memory mode a rank only updates the
// state of the chunks it owns; the shards are disjoint, so zero-fill + own range + all-reduce(sum) assembles them.
// The flat gradient buffer is dead between steps and serves as the scratch.
int pycnn_b200_comm_sync_opt_state(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  if (!c->p2p || c->world <= 1 || !c->d_m) return 0;
  NcclApi* api = nccl_api();
  PB_CHECK(api && c->nccl_comm, "the peer-memory mode needs the NCCL communicator to gather optimizer state");
  int64_t lo = 0, hi = 0;
  if (c->p2p_nchunks > 0) {
    lo = c->h_chunks[c->p2p_chunk0].off;
    const Chunk& last = c->h_chunks[c->p2p_chunk0 + c->p2p_nchunks - 1];
    hi = last.off + last.len;
  }
  const int blocks = grid_for(c->P, 256, c->num_sms);
  for (float* state : {c->d_m, c->d_v}) {
    p2p::owned_range_kernel<<<blocks, 256, 0, c->stream>>>(state, c->d_grads, c->P, lo, hi);
    PB_CUDA(cudaGetLastError());
    PB_NCCL(api, api->AllReduce(c->d_grads, state, (size_t)c->P, kNcclFloat32, kNcclSum, c->nccl_comm, c->stream));
  }
  PB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- NVSwitch multicast exchange (nvls.cuh) -----------------------------------------------------------------
// Layout of the per-rank region: [gradient | reduced gradient | chunk norms | counters], every part 4 KB aligned.
static int nvls_layout(pycnn_b200_ctx* c, size_t gran) {
  const size_t pb = (size_t)round_up((int64_t)sizeof(float) * (c->P + kSlackFloats), 4096);
  const size_t nb = (size_t)round_up((int64_t)sizeof(double) * std::max(c->num_chunks, 1), 4096);
  c->nvls_off_gred = pb;
  c->nvls_off_norm = 2 * pb;
  c->nvls_off_flags = 2 * pb + nb;
  c->nvls_size = (size_t)round_up((int64_t)(2 * pb + nb + 4096), (int64_t)gran);
  c->nvls_gran = gran;
  return 0;
}

static int nvls_granularity(pycnn_b200_ctx* c, nvls::Driver* d, int world, int dev, size_t* gran) {
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)world;
  mp.size = 2u << 20;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g_mc = 0, g_mem = 0;
  PB_CU(d, d->MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
  CUmemAllocationProp ap{};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  PB_CU(d, d->MemGetAllocationGranularity(&g_mem, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  *gran = std::max<size_t>(std::max(g_mc, g_mem), 2u << 20);
  (void)c;
  return 0;
}

int pycnn_b200_comm_nvls_supported(pycnn_b200_ctx* c, int* supported) {
  CTX_GUARD(c);
  PB_CHECK(supported, "null argument");
  *supported = 0;
  nvls::Driver* d = nvls::driver();
  if (!d) return 0;
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  CUdevice cudev = 0;
  if (d->DeviceGet(&cudev, dev) != CUDA_SUCCESS) return 0;
  int mc = 0, fd = 0;
  if (d->DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev) != CUDA_SUCCESS) return 0;
  if (d->DeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev) != CUDA_SUCCESS) return 0;
  *supported = (mc && fd) ? 1 : 0;
  return 0;
}

// rank 0: create the multicast object for `world` devices, return a file descriptor the other ranks import
int pycnn_b200_comm_nvls_create(pycnn_b200_ctx* c, int world, int* fd_out) {
  CTX_GUARD(c);
  PB_CHECK(fd_out && world >= 2 && world <= p2p::MAX_RANKS, "bad arguments");
  PB_CHECK(c->d_params && c->d_grads, "set_network must be called first");
  PB_CHECK(c->world == world, "comm_init(rank, world) must come first");
  nvls::Driver* d = nvls::driver();
  PB_CHECK(d, "the driver does not expose the multicast entry points");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  p2p_detach(c);
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  size_t gran = 0;
  PB_TRY(nvls_granularity(c, d, world, dev, &gran));
  PB_TRY(nvls_layout(c, gran));
  CUmulticastObjectProp mp{};
  mp.numDevices = (unsigned)world;
  mp.size = c->nvls_size;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  PB_CU(d, d->MulticastCreate(&c->nvls_mc, &mp));
  int fd = -1;
  CUresult r = d->MemExportToShareableHandle(&fd, c->nvls_mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    nvls_detach(c);
    return fail("cuMemExportToShareableHandle(multicast object) failed: %s", nvls::cu_err(d, r));
  }
  *fd_out = fd;
  return 0;
}

// every rank: import the object (fd < 0 on the rank that created it) and add this rank's device.  The caller puts a
// barrier between join and bind: memory can only be bound once every device has been added.
int pycnn_b200_comm_nvls_join(pycnn_b200_ctx* c, int fd, int rank, int world) {
  CTX_GUARD(c);
  PB_CHECK(world >= 2 && world <= p2p::MAX_RANKS && rank >= 0 && rank < world, "bad arguments");
  PB_CHECK(c->d_params && c->d_grads, "set_network must be called first");
  PB_CHECK(c->world == world && c->rank == rank, "comm_init(rank, world) must match the multicast attach");
  nvls::Driver* d = nvls::driver();
  PB_CHECK(d, "the driver does not expose the multicast entry points");
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  CUdevice cudev = 0;
  PB_CU(d, d->DeviceGet(&cudev, dev));
  if (fd >= 0) {
    PB_CHECK(!c->nvls_mc, "this rank already holds the multicast object");
    PB_CUDA(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    p2p_detach(c);
    size_t gran = 0;
    PB_TRY(nvls_granularity(c, d, world, dev, &gran));
    PB_TRY(nvls_layout(c, gran));
    PB_CU(d, d->MemImportFromShareableHandle(&c->nvls_mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                             CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  }
  PB_CHECK(c->nvls_mc, "comm_nvls_create (rank 0) or an imported descriptor must come first");
  CUresult r = d->MulticastAddDevice(c->nvls_mc, cudev);
  if (r != CUDA_SUCCESS) {
    nvls_detach(c);
    return fail("cuMulticastAddDevice failed: %s", nvls::cu_err(d, r));
  }
  c->nvls_added = true;
  return 0;
}

static int nvls_bind_impl(pycnn_b200_ctx* c, nvls::Driver* d) {
  int dev = 0;
  PB_CUDA(cudaGetDevice(&dev));
  CUmemAllocationProp ap{};
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  PB_CU(d, d->MemCreate(&c->nvls_mem, c->nvls_size, &ap, 0));
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  // unicast view first, zeroed before the memory becomes reachable through the switch
  PB_CU(d, d->MemAddressReserve(&c->nvls_uc, c->nvls_size, c->nvls_gran, 0, 0));
  PB_CU(d, d->MemMap(c->nvls_uc, c->nvls_size, 0, c->nvls_mem, 0));
  PB_CU(d, d->MemSetAccess(c->nvls_uc, c->nvls_size, &acc, 1));
  PB_CUDA(cudaMemset(reinterpret_cast<void*>(c->nvls_uc), 0, c->nvls_size));
  PB_CUDA(cudaDeviceSynchronize());
  PB_CU(d, d->MulticastBindMem(c->nvls_mc, 0, c->nvls_mem, 0, c->nvls_size, 0));
  c->nvls_bound = true;
  PB_CU(d, d->MemAddressReserve(&c->nvls_mcva, c->nvls_size, c->nvls_gran, 0, 0));
  PB_CU(d, d->MemMap(c->nvls_mcva, c->nvls_size, 0, c->nvls_mc, 0));
  PB_CU(d, d->MemSetAccess(c->nvls_mcva, c->nvls_size, &acc, 1));
  PB_CUDA(cudaMalloc(&c->d_p2p_seq, 2 * sizeof(long long)));
  PB_CUDA(cudaMemset(c->d_p2p_seq, 0, 2 * sizeof(long long)));
  PB_CUDA(cudaMalloc(&c->d_blocks_done, 4 * sizeof(unsigned int)));
  PB_CUDA(cudaMemset(c->d_blocks_done, 0, 4 * sizeof(unsigned int)));
  PB_CUDA(cudaDeviceSynchronize());
  return 0;
}

// every rank, after a barrier behind join: allocate the region, bind it, map both views, move the gradient buffer
// into it.  The caller puts another barrier behind bind before anybody steps.
int pycnn_b200_comm_nvls_bind(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  PB_CHECK(c->nvls_mc && c->nvls_added, "comm_nvls_join must come first");
  nvls::Driver* d = nvls::driver();
  PB_CHECK(d, "the driver does not expose the multicast entry points");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  drop_graphs(c);
  if (nvls_bind_impl(c, d) != 0) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep));
    nvls_detach(c);
    memcpy(g_err, keep, sizeof(keep));
    return 1;
  }
  free_dev(c->d_grads);
  c->d_grads = reinterpret_cast<float*>(c->nvls_uc);
  const int nc = c->num_chunks;
  c->p2p_chunk0 = (int)((int64_t)c->rank * nc / c->world);
  c->p2p_nchunks = (int)((int64_t)(c->rank + 1) * nc / c->world) - c->p2p_chunk0;
  c->nvls = true;
  return 0;
}

int pycnn_b200_comm_nvls_detach(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  drop_graphs(c);
  return nvls_detach(c);
}

int pycnn_b200_comm_ipc_detach(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  PB_TRY(pycnn_b200_comm_sync_opt_state(c));   // the replicated update that takes over needs the whole state everywhere
  drop_graphs(c);
  p2p_detach(c);
  return 0;
}

int pycnn_b200_grad_buffer(pycnn_b200_ctx* c, void** ptr, int64_t* n) {
  PB_CHECK(c && ptr && n, "null argument");
  PB_CHECK(c->d_grads, "set_network must be called first");
  *ptr = c->d_grads;
  *n = c->P;
  return 0;
}

// ---- measurement -------------------------------------------------------------------------------
int pycnn_b200_event_record(pycnn_b200_ctx* c, int slot) {
  CTX_GUARD(c);
  PB_CHECK(slot >= 0 && slot < 16, "event slot %d out of range", slot);
  PB_CUDA(cudaEventRecord(c->events[slot], c->stream));
  return 0;
}

int pycnn_b200_event_elapsed_ms(pycnn_b200_ctx* c, int a, int b, float* ms) {
  CTX_GUARD(c);
  PB_CHECK(ms && a >= 0 && a < 16 && b >= 0 && b < 16, "bad arguments");
  PB_CUDA(cudaEventSynchronize(c->events[b]));
  PB_CUDA(cudaEventElapsedTime(ms, c->events[a], c->events[b]));
  return 0;
}

int pycnn_b200_set_kernel_timing(pycnn_b200_ctx* c, int enabled) {
  CTX_GUARD(c);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  c->timing = enabled != 0;   // while on, steps are issued eagerly (not through graphs) so each launch can be timed
  return 0;
}

int pycnn_b200_kernel_class_count(void) { return KC_COUNT; }
const char* pycnn_b200_kernel_class_name(int cls) { return (cls >= 0 && cls < KC_COUNT) ? kKernelClassNames[cls] : ""; }

int pycnn_b200_kernel_stats(pycnn_b200_ctx* c, int cls, int64_t* launches, double* total_ms, int reset) {
  CTX_GUARD(c);
  PB_CHECK(cls >= 0 && cls < KC_COUNT, "kernel class %d out of range", cls);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  for (auto& t : c->timed[cls]) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) c->class_ms[cls] += ms;
    c->event_pool.push_back(t.a);
    c->event_pool.push_back(t.b);
  }
  c->timed[cls].clear();
  if (launches) *launches = c->class_launches[cls];
  if (total_ms) *total_ms = c->class_ms[cls];
  if (reset) { c->class_launches[cls] = 0; c->class_ms[cls] = 0.0; }
  return 0;
}

int pycnn_b200_launch_count(pycnn_b200_ctx* c, int64_t* launches) {
  PB_CHECK(c && launches, "null argument");
  *launches = c->launches;
  return 0;
}

static int trace_fill(pycnn_b200_ctx* c) {
  std::vector<unsigned long long> init(kTraceWords * (size_t)c->trace_cap);
  for (size_t i = 0; i < init.size(); ++i) init[i] = (i & 1) ? 0ull : ~0ull;   // even words take minima, odd words maxima
  PB_CUDA(cudaMemcpy(c->d_trace, init.data(), sizeof(unsigned long long) * init.size(), cudaMemcpyHostToDevice));
  return 0;
}

int pycnn_b200_trace_control(pycnn_b200_ctx* c, int op, int capacity) {
  CTX_GUARD(c);
  PB_CUDA(cudaStreamSynchronize(c->stream));
  if (op == 0) {
    PB_CHECK(capacity > 0 && capacity <= (1 << 20), "trace capacity %d out of range", capacity);
    drop_graphs(c);
    free_dev(c->d_trace); c->d_trace = nullptr;
    PB_CUDA(cudaMalloc(&c->d_trace, sizeof(unsigned long long) * kTraceWords * (size_t)capacity));
    c->trace_cap = capacity; c->trace_next = 0; c->trace_cls.clear(); c->trace = true;
    return trace_fill(c);
  }
  if (op == 1) {
    PB_CHECK(c->d_trace, "trace is not enabled");
    return trace_fill(c);
  }
  PB_CHECK(op == 2, "unknown trace op %d", op);
  drop_graphs(c);   // captured launches hold slot pointers
  c->trace = false;
  free_dev(c->d_trace); c->d_trace = nullptr;
  c->trace_cap = c->trace_next = 0; c->trace_cls.clear();
  return 0;
}

int pycnn_b200_trace_read(pycnn_b200_ctx* c, uint64_t* times, int32_t* kinds, int capacity, int* count) {
  CTX_GUARD(c);
  PB_CHECK(times && kinds && count, "null argument");
  PB_CHECK(c->d_trace, "trace is not enabled");
  PB_CUDA(cudaStreamSynchronize(c->stream));
  const int n = std::min(capacity, c->trace_next);
  if (n > 0) PB_CUDA(cudaMemcpy(times, c->d_trace, sizeof(unsigned long long) * kTraceWords * (size_t)n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; ++i) kinds[i] = c->trace_cls[i];
  *count = n;
  return 0;
}

int pycnn_b200_flush_l2(pycnn_b200_ctx* c) {
  CTX_GUARD(c);
  if (!c->d_flush) {
    c->flush_bytes = std::max<size_t>(2 * c->l2_bytes, 256ull << 20);
    PB_CUDA(cudaMalloc(&c->d_flush, c->flush_bytes));
  }
  PB_CUDA(cudaMemsetAsync(c->d_flush, 0, c->flush_bytes, c->stream));
  return 0;
}

int pycnn_b200_gemm_probe(pycnn_b200_ctx* c, const float* A, const float* B, float* C, int M, int N, int K,
                          int a_kmajor, int b_kmajor, int iters, float* avg_ms) {
  CTX_GUARD(c);
  PB_CHECK(A && B && C && M > 0 && N > 0 && K > 0 && iters > 0, "bad arguments");
  const int64_t a_rows = a_kmajor ? M : K, a_cols = a_kmajor ? K : M;
  const int64_t b_rows = b_kmajor ? N : K, b_cols = b_kmajor ? K : N;
  const int64_t lda = round_up(a_cols, 4), ldb = round_up(b_cols, 4), ldc = round_up(N, 4);
  float *dA = nullptr, *dB = nullptr, *dC = nullptr;
  PB_TRY(alloc_f32(&dA, a_rows * lda));
  PB_TRY(alloc_f32(&dB, b_rows * ldb));
  PB_TRY(alloc_f32(&dC, (int64_t)M * ldc));
  int rc = 0;
  cudaError_t e = cudaMemcpy2D(dA, lda * 4, A, a_cols * 4, a_cols * 4, a_rows, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy2D(dB, ldb * 4, B, b_cols * 4, b_cols * 4, b_rows, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) rc = fail("probe upload failed: %s", cudaGetErrorString(e));
  GemmCall g{};
  g.A = dA; g.lda = lda; g.a_kmajor = a_kmajor != 0;
  g.B = dB; g.ldb = ldb; g.b_kmajor = b_kmajor != 0;
  g.C = dC; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.epi = 0; g.cls = KC_MISC;
  if (!rc) rc = gemm(c, g);  // warm-up
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (!rc) {
    cudaEventRecord(e0, c->stream);
    for (int i = 0; i < iters && !rc; ++i) rc = gemm(c, g);
    cudaEventRecord(e1, c->stream);
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail("gemm probe failed: %s", cudaGetErrorString(e));
  }
  if (!rc) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (avg_ms) *avg_ms = ms / iters;
    e = cudaMemcpy2D(C, (size_t)N * 4, dC, ldc * 4, (size_t)N * 4, M, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail("probe download failed: %s", cudaGetErrorString(e));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dB); cudaFree(dC);
  return rc;
}

}  // extern "C"
