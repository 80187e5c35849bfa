// Shared host/device helpers of libkrabgpu (sm_100a only).  No torch types, no CPU fallback.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>
#include <vector>

#include "../../include/krabgpu.h"

namespace kg {

// ----------------------------------------------------------------------------- errors
inline std::string& last_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}
#define KG_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return kg::fail(KG_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                      __FILE__, __LINE__);                                                \
  } while (0)
#define KG_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != KG_OK) return _rc; \
  } while (0)

// every kernel launch of the library goes through this counter (bench.py: gpu_launches)
inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

// device-side deferred error word
enum : int { DEV_ERR_OOB = 1, DEV_ERR_SENTINEL = 2 };

// ----------------------------------------------------------------------------- kernel timers
struct Profiler {
  bool enabled = false;
  double ms[KG_K_COUNT] = {0};
  uint64_t launches[KG_K_COUNT] = {0};
  static const int kRing = 64;
  cudaEvent_t ev0[kRing], ev1[kRing];
  int kind[kRing];
  int head = 0, pending = 0;
  bool made = false;
  int ensure() {
    if (made) return KG_OK;
    for (int i = 0; i < kRing; ++i) {
      KG_CUDA(cudaEventCreate(&ev0[i]));
      KG_CUDA(cudaEventCreate(&ev1[i]));
    }
    made = true;
    return KG_OK;
  }
  void drain() {
    while (pending > 0) {
      int i = (head - pending + kRing * 4) % kRing;
      cudaEventSynchronize(ev1[i]);
      float t = 0.f;
      cudaEventElapsedTime(&t, ev0[i], ev1[i]);
      ms[kind[i]] += t;
      --pending;
    }
  }
  void begin(int k, cudaStream_t s) {
    launches[k] += 1;
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    if (!enabled) return;
    if (ensure() != KG_OK) return;
    if (pending == kRing) drain();
    kind[head] = k;
    cudaEventRecord(ev0[head], s);
  }
  void end(cudaStream_t s) {
    if (!enabled || !made) return;
    cudaEventRecord(ev1[head], s);
    head = (head + 1) % kRing;
    ++pending;
  }
  void destroy() {
    if (!made) return;
    for (int i = 0; i < kRing; ++i) {
      cudaEventDestroy(ev0[i]);
      cudaEventDestroy(ev1[i]);
    }
    made = false;
  }
};

// L2 flush: stream-ordered overwrite of a private buffer larger than the 126 MB L2
static __global__ void l2_flush_kernel(uint4* __restrict__ p, uint64_t n16, uint32_t tag) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(tag, tag, tag, tag);
}
struct L2Flusher {
  void* buf = nullptr;
  uint64_t bytes = 0;
  uint32_t tag = 0;
  int run(uint64_t want, cudaStream_t s) {
    if (want == 0) return KG_OK;
    if (bytes < want) {
      if (buf) cudaFree(buf);
      buf = nullptr;
      bytes = 0;
      KG_CUDA(cudaMalloc(&buf, want));
      bytes = want;
    }
    l2_flush_kernel<<<kNumSMsFlush, 256, 0, s>>>((uint4*)buf, want / 16, ++tag);
    return KG_OK;
  }
  void destroy() {
    if (buf) cudaFree(buf);
    buf = nullptr;
    bytes = 0;
  }
  static constexpr int kNumSMsFlush = 148 * 8;
};

// a growable pool of event pairs for per-step timing without host syncs inside the loop
struct EventPool {
  std::vector<cudaEvent_t> ev;
  int get(size_t i, cudaEvent_t* out) {
    while (ev.size() <= i) {
      cudaEvent_t e;
      KG_CUDA(cudaEventCreate(&e));
      ev.push_back(e);
    }
    *out = ev[i];
    return KG_OK;
  }
  void destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

// stopwatch: two events on one stream
struct Stopwatch {
  cudaEvent_t a = nullptr, b = nullptr;
  int start(cudaStream_t s) {
    if (!a) {
      KG_CUDA(cudaEventCreate(&a));
      KG_CUDA(cudaEventCreate(&b));
    }
    KG_CUDA(cudaEventRecord(a, s));
    return KG_OK;
  }
  int stop(cudaStream_t s, double* ms) {
    if (!a) return fail(KG_E_INVALID, "timer_stop without timer_start");
    KG_CUDA(cudaEventRecord(b, s));
    KG_CUDA(cudaEventSynchronize(b));
    float t = 0.f;
    KG_CUDA(cudaEventElapsedTime(&t, a, b));
    if (ms) *ms = t;
    return KG_OK;
  }
  void destroy() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    a = b = nullptr;
  }
};

// ----------------------------------------------------------------------------- device math
// All f32 arithmetic on the path must round exactly like Rust's: IEEE-754 rn, no FMA
// contraction (the TU is compiled with -fmad=false; the explicit intrinsics below make the
// intent local), `%` == fmodf, float->int casts saturate with NaN -> 0 (cvt.rzi does that).
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ int f2i_sat(float v) { return __float2int_rz(v); }  // Rust `as i32`

// Philox4x32-10 (Salmon et al. SC'11).  Stream layout: DESIGN.md §RNG.
struct Philox4 {
  uint32_t v[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{{c0, c1, c2, c3}};
}
// rand 0.9 StandardUniform<f32>: 24 high bits * 2^-24
__host__ __device__ __forceinline__ float u01_f32(uint32_t u) {
  return (float)(u >> 8) * (1.0f / 16777216.0f);
}
enum : uint32_t { DOMAIN_INIT = 0, DOMAIN_STEP = 1, DOMAIN_GRID = 2, DOMAIN_LIFE = 3 };
constexpr uint32_t kIdNone = 0xFFFFFFFFu;  // log entry of an agent that stopped (is_stopped): skipped by the rebuild

constexpr int kNumSMs = 148;  // B200

// ----------------------------------------------------------------------------- dependent launches
// Programmatic dependent launch (sm_90+): a kernel launched with launch_pdl may be scheduled while
// the previous kernel of the stream is still draining, which hides the ~2 us launch latency
// between the three kernels of a step.  Such a kernel must call grid_dep_wait() before it touches
// anything the previous kernel wrote; the wait returns once that kernel has completed and its
// writes are visible (a no-op for ordinary launches).
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the NEXT kernel of the stream (if it was launched with launch_pdl) start running its blocks now.
// Called by a kernel AFTER its own grid_dep_wait(): everything older than this kernel is then complete, so
// the dependent may read it before its own wait — only this kernel's output needs the dependent's wait.
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifndef KG_SCATTER_EARLY
#define KG_SCATTER_EARLY 0  // 1: K3 loads the write log while K2 (the scan) is still running — measured on B200:
                            // 63.4-63.8 vs 62.6-62.8 us per step at 1M agents (the resident scatter blocks slow the scan): off
#endif

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ----------------------------------------------------------------------------- bulk copies (TMA)
// cp.async.bulk: the copy engine moves one contiguous, 16-byte aligned segment global -> shared
// and signals an mbarrier with the byte count (SASS: UBLKCP).  One thread issues a copy; nobody
// spends an instruction per element.  Addresses are shared-window addresses of this CTA.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
// the single arrival of a phase, announcing how many bytes the copies will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// bytes: multiple of 16, > 0; dst and src 16-byte aligned
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

}  // namespace kg
