// nvls.cuh -- data-parallel gradient all-reduce through NVSwitch multicast (NVLS), one launch per chunk group and step
// (api.cu exchanges the first layer's tensors right behind dW_1 and the shallow tensors beside it: two groups).
//
// The peer-memory exchange (p2p.cuh) needs three flag rounds per step and moves the gradient W-1 times over each
// GPU's links (every owner reads its range from every peer).  With a multicast object bound on every rank the switch
// does both halves of the all-reduce:
//
//   nvls_allreduce_kernel   rank r owns a contiguous range of the flat buffer's chunks.
//                           arrive  one multimem.red (+1 on every rank's arrive counter) after a system fence; every
//                                   block waits until its LOCAL counter shows world x seq;
//                           reduce  multimem.ld_reduce.add.v4.f32 over the owned range: the switch pulls the 16 bytes
//                                   from every rank's gradient buffer, adds them in fp32 and returns the sum;
//                           publish multimem.st of the sum into every rank's REDUCED buffer, and of the chunk's sum of
//                                   squares (same fixed-shape block reduction as every other path) into every rank's
//                                   chunk-norm table;
//                           done    the last block: system fence, multimem.red on the done counter, and a wait until
//                                   the local one shows world x seq -- the launch retires only when this rank's reduced
//                                   gradient and norm table are complete, and no peer still reads this rank's
//                                   gradient (the next step's finalize may overwrite it).
//
// The update that follows is the ordinary replicated clip + SGD/Adam (elementwise.cuh) on the reduced buffer: optimizer
// state stays whole on every rank, parameters never cross the links.  Per step each GPU sends its gradient once and
// receives one reduced copy -- independent of the rank count.
//
// The gradient buffer, the reduced buffer, the norm table and the two counters live in ONE physical allocation per rank
// (cuMemCreate), mapped twice: a unicast range the local kernels use (finalize writes the gradient there, the update
// reads the reduced gradient there) and the multicast range the kernel below addresses.
#pragma once
#include "p2p.cuh"

namespace pb {
namespace nvls {

struct Table {
  float* mc_grads;                 // multicast views
  float* mc_gred;
  double* mc_norm;
  unsigned long long* mc_flags;    // [0] arrive, [1] done
  const unsigned long long* uc_flags;   // this rank's copy of the counters
  int rank, world;
  int chunk0, nchunks;             // owned chunk range
  unsigned int* lost;              // 0, or a mark that a wait ran into its bound (host reports it)
};

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mc_st(double* p, double v) {
  asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void mc_signal(unsigned long long* p) {
  asm volatile("multimem.red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(1ull) : "memory");
}

// grid = max(1, owned chunks), block 256.  *seq_word = exchanges of this group completed since bind (advanced by the
// last block); t.mc_flags / t.uc_flags point at the group's own {arrive, done} counter pair.
__global__ void __launch_bounds__(256) nvls_allreduce_kernel(Table t, const Chunk* __restrict__ chunks,
                                                             long long* __restrict__ seq_word,
                                                             unsigned int* __restrict__ blocks_done,
                                                             unsigned long long* trace) {
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)*seq_word + 1ull;
  const unsigned long long target = seq * (unsigned long long)t.world;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();            // my gradient (previous kernels of this stream) is visible system-wide
      mc_signal(&t.mc_flags[0]);
    }
    p2p::wait_flag(&t.uc_flags[0], target, t.lost, p2p::MAX_RANKS);
  }
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(trace, 1);          // every rank's gradient is ready
  __shared__ double sh[8];
  if ((int)blockIdx.x < t.nchunks) {
    const int ci = t.chunk0 + blockIdx.x;
    const Chunk ch = chunks[ci];
    const int64_t n4 = ch.len >> 2;
    float acc = 0.f;
    for (int64_t i0 = threadIdx.x; i0 < n4; i0 += 2 * blockDim.x) {
      float4 s[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {          // both switch round trips in flight before the first use
        const int64_t i = i0 + e * (int64_t)blockDim.x;
        if (i < n4) s[e] = mc_ld_reduce(t.mc_grads + ch.off + 4 * i);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int64_t i = i0 + e * (int64_t)blockDim.x;
        if (i < n4) {
          mc_st(t.mc_gred + ch.off + 4 * i, s[e]);
          acc = fmaf(s[e].x, s[e].x, acc); acc = fmaf(s[e].y, s[e].y, acc);
          acc = fmaf(s[e].z, s[e].z, acc); acc = fmaf(s[e].w, s[e].w, acc);
        }
      }
    }
    const double d = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += sh[i];
      mc_st(t.mc_norm + ci, tot);
    }
  }
  if (threadIdx.x == 0) trace_mark(trace, 2);          // owned chunks reduced and published
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();              // cumulative: orders the whole block's multicast stores
    last = (atomicAdd(blocks_done, 1u) == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *blocks_done = 0u;
    mc_signal(&t.mc_flags[1]);
    p2p::wait_flag(&t.uc_flags[1], target, t.lost, p2p::MAX_RANKS);
    *seq_word = (long long)seq;          // every block of this launch has already read the old value
  }
  trace_end(trace);
}

// ---- all-reduce + clip + update in ONE launch -------------------------------------------------------------------
// grid = num_chunks (every chunk of the flat buffer), block 256; block b works on chunk (chunk0 + b) mod num_chunks,
// so the blocks that own a chunk of the reduction come first in launch order and never wait on a later block.
//   every block   requests its chunk's p, m, v right away -- those loads fly during the exchange;
//   owner blocks  arrive barrier -> multimem.ld_reduce -> multimem.st of the sum and of the chunk's ||.||^2 to every
//                 rank -> system fence -> ONE multimem.red on the done counter per owned chunk (no last-block chain);
//   every block   waits until the local done counter shows num_chunks x seq (all sums and norms of the step have landed
//                 here, and nobody reads this rank's gradient any more), forms its tensor's norm from the chunk table,
//                 clips and updates.  Owner blocks use the sum they still hold in registers.
// Small networks only (the host picks it when a block may add its tensor's chunk norms itself, i.e. !big_tensors).
__global__ void __launch_bounds__(256) nvls_allreduce_update_kernel(Table t, const Chunk* __restrict__ chunks, int num_chunks,
                                                                    long long* __restrict__ seq_ptr,
                                                                    float* __restrict__ params, float* __restrict__ m,
                                                                    float* __restrict__ v, const float* __restrict__ uc_gred,
                                                                    const double* __restrict__ uc_norm, UpdateParams u,
                                                                    unsigned int* __restrict__ blocks_done,
                                                                    unsigned long long* trace) {
  __shared__ double s8[8];
  pdl_wait();
  trace_begin(trace);
  const unsigned long long seq = (unsigned long long)seq_ptr[0] + 1ull;
  int ci = t.chunk0 + (int)blockIdx.x;
  if (ci >= num_chunks) ci -= num_chunks;
  const bool own = (int)blockIdx.x < t.nchunks;
  const Chunk ch = chunks[ci];
  const int n4 = (int)(ch.len >> 2);
  float4* p4 = reinterpret_cast<float4*>(params + ch.off);
  float4* m4 = reinterpret_cast<float4*>(m + ch.off);
  float4* v4 = reinterpret_cast<float4*>(v + ch.off);
  float4 p[2], g[2], mm[2], vv[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = threadIdx.x + e * 256;
    mm[e] = vv[e] = g[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n4) {
      p[e] = p4[i];
      if (u.opt == 1) { mm[e] = m4[i]; vv[e] = v4[i]; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();              // my gradient (previous kernels of this stream) is visible system-wide
    mc_signal(&t.mc_flags[0]);
  }
  if (own) {
    if (threadIdx.x == 0) p2p::wait_flag(&t.uc_flags[0], seq * (unsigned long long)t.world, t.lost, p2p::MAX_RANKS);
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(trace, 1);        // every rank's gradient is ready
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = threadIdx.x + e * 256;
      if (i < n4) g[e] = mc_ld_reduce(t.mc_grads + ch.off + 4 * i);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = threadIdx.x + e * 256;
      if (i < n4) {
        mc_st(t.mc_gred + ch.off + 4 * i, g[e]);
        acc = fmaf(g[e].x, g[e].x, acc); acc = fmaf(g[e].y, g[e].y, acc);
        acc = fmaf(g[e].z, g[e].z, acc); acc = fmaf(g[e].w, g[e].w, acc);
      }
    }
    const double d = warp_sum_d((double)acc);
    if ((threadIdx.x & 31) == 0) s8[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += s8[i];
      mc_st(t.mc_norm + ci, tot);
    }
    __syncthreads();                     // the whole block's multicast stores are issued ...
    if (threadIdx.x == 0) {
      __threadfence_system();            // ... and ordered (cumulative) before the signal
      mc_signal(&t.mc_flags[1]);
      trace_mark(trace, 2);              // owned chunk reduced and published
    }
  }
  if (threadIdx.x == 0) p2p::wait_flag(&t.uc_flags[1], seq * (unsigned long long)num_chunks, t.lost, p2p::MAX_RANKS);
  __syncthreads();
  if (threadIdx.x == 0) trace_mark(trace, 3);          // every chunk's sum and norm has landed here
  if (!own) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = threadIdx.x + e * 256;
      if (i < n4) g[e] = __ldcg(reinterpret_cast<const float4*>(uc_gred + ch.off) + i);
    }
  }
  float scale = 1.0f;
  bool skip = false;
  if (u.rule == 1 && u.opt == 1) {
    const double tt = (double)max((long long)u.steps_done[1], 1ll);
    u.inv_bc1 = (float)(1.0 / (1.0 - pow(u.beta1_d, tt)));
    u.inv_bc2 = (float)(1.0 / (1.0 - pow(u.beta2_d, tt)));
  }
  if (u.rule == 0) {
    const double n2 = tensor_norm2(uc_norm, ch, s8);
    if (!isfinite(n2)) skip = true;      // NaN/Inf gradient: this tensor keeps its value (on every rank alike)
    const double nrm = sqrt(n2);
    if (!skip && nrm > (double)u.max_norm) scale = (float)((double)u.max_norm / nrm);
  }
  if (!skip) {
    const float omb1 = 1.0f - u.beta1, omb2 = 1.0f - u.beta2;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = threadIdx.x + e * 256;
      if (i < n4) {
        adam_or_sgd4(p[e], mm[e], vv[e], g[e], scale, u, omb1, omb2);
        p4[i] = p[e];
        if (u.opt == 1) { m4[i] = mm[e]; v4[i] = vv[e]; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(blocks_done, 1u) == gridDim.x - 1) {   // every block has read the old sequence number by now
      *blocks_done = 0u;
      seq_ptr[0] = (long long)seq;
    }
  }
  trace_end(trace);
}

// ---- host side: the driver's virtual-memory and multicast entry points (no link-time dependency on libcuda) -----
struct Driver {
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
};

inline Driver* driver() {
  static Driver d;
  static bool tried = false;
  if (tried) return d.ok ? &d : nullptr;
  tried = true;
  bool all = true;
  auto get = [&](const char* name, void* slot) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
      all = false;
      (void)cudaGetLastError();
      return;
    }
    *reinterpret_cast<void**>(slot) = fn;
  };
  get("cuMulticastCreate", &d.MulticastCreate);
  get("cuMulticastAddDevice", &d.MulticastAddDevice);
  get("cuMulticastBindMem", &d.MulticastBindMem);
  get("cuMulticastUnbind", &d.MulticastUnbind);
  get("cuMulticastGetGranularity", &d.MulticastGetGranularity);
  get("cuMemCreate", &d.MemCreate);
  get("cuMemRelease", &d.MemRelease);
  get("cuMemAddressReserve", &d.MemAddressReserve);
  get("cuMemAddressFree", &d.MemAddressFree);
  get("cuMemMap", &d.MemMap);
  get("cuMemUnmap", &d.MemUnmap);
  get("cuMemSetAccess", &d.MemSetAccess);
  get("cuMemExportToShareableHandle", &d.MemExportToShareableHandle);
  get("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle);
  get("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity);
  get("cuDeviceGet", &d.DeviceGet);
  get("cuDeviceGetAttribute", &d.DeviceGetAttribute);
  get("cuGetErrorString", &d.GetErrorString);
  d.ok = all;
  return d.ok ? &d : nullptr;
}

inline const char* cu_err(Driver* d, CUresult r) {
  const char* s = nullptr;
  if (d && d->GetErrorString) d->GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

#define PB_CU(d, expr)                                                                                      \
  do {                                                                                                      \
    CUresult _r = (expr);                                                                                   \
    if (_r != CUDA_SUCCESS) return pb::fail("%s failed: %s (%s:%d)", #expr, pb::nvls::cu_err(d, _r), __FILE__, __LINE__); \
  } while (0)

}  // namespace nvls
}  // namespace pb
