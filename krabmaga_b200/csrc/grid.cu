// DenseNumberGrid2D on the device: double-buffered dense grid of Option<T>, the Forest-Fire
// class stencil kernel (K5) and the C ABI over them.
//
// Replaces (reference paths relative to the krABMaga crate root):
//   DenseNumberGrid2D::{new,get_value*,set_value_location,remove_value_location,
//   apply_to_all_values,get_location*,get_empty_bags,lazy_update,update}
//                                  src/engine/fields/dense_number_grid_2d.rs:90-561
//
// HBM layout: two buffers of width*height elements of T (1, 2 or 4 bytes), flat index
// x*height + y (y contiguous, as in the reference :351), Option::None stored as the reserved
// value `none`.  lazy_update is a pointer swap plus a *pending* clear of the new write buffer:
// the clear is materialised only if somebody reads or partially writes that buffer before a
// full-grid kernel overwrites it, so a step+swap loop moves exactly 1 read + 1 write per cell.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "jit.cuh"
#include "stencil_device.cuh"

namespace kg {

template <class T>
__global__ void fill_kernel(T* __restrict__ p, uint64_t n, T v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
__global__ void fill16_kernel(uint4* __restrict__ p, uint64_t n16, uint4 v) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = v;
}

// flat index with the reference's i32 arithmetic (dense_number_grid_2d.rs:493); valid iff the
// reference's Vec indexing would not panic
__device__ __forceinline__ bool grid_index(int32_t x, int32_t y, int32_t height, uint64_t ncells,
                                           uint64_t* idx) {
  int32_t i = (int32_t)((uint32_t)x * (uint32_t)height + (uint32_t)y);
  *idx = (uint64_t)(uint32_t)i;
  return i >= 0 && (uint64_t)i < ncells;
}

template <class T>
__global__ void set_values_kernel(T* __restrict__ buf, int32_t height, uint64_t ncells, uint64_t n,
                                  const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                  const T* __restrict__ v, T fill, bool use_fill) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  if (grid_index(x[i], y[i], height, ncells, &idx)) buf[idx] = use_fill ? fill : v[i];
}
__global__ void check_index_kernel(int32_t height, uint64_t ncells, uint64_t n,
                                   const int32_t* __restrict__ x, const int32_t* __restrict__ y,
                                   int* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  if (!grid_index(x[i], y[i], height, ncells, &idx)) atomicOr(err, DEV_ERR_OOB);
}
template <class T>
__global__ void get_values_kernel(const T* __restrict__ buf, int32_t height, uint64_t ncells,
                                  uint64_t n, const int32_t* __restrict__ x,
                                  const int32_t* __restrict__ y, T* __restrict__ out, T none,
                                  bool all_none, int* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t idx;
  if (grid_index(x[i], y[i], height, ncells, &idx))
    out[i] = all_none ? none : buf[idx];
  else {
    out[i] = none;
    atomicOr(err, DEV_ERR_OOB);
  }
}

// apply_to_all_values  dense_number_grid_2d.rs:155-195
template <class T>
__device__ __forceinline__ T apply_op(int op, T v, T c) {
  return op == KG_APPLY_CONST ? c : (T)(v + c);
}
// `Some(v)` is stored as v and `None` as the reserved value `none`, so a closure result equal to
// `none` cannot be represented: it raises DEV_ERR_SENTINEL (-> KG_E_INVALID at the next sync) instead
// of silently turning Some(254 + 1) into None.  ADD otherwise wraps like a Rust release build
// (a debug build would panic on overflow).
template <class T>
__global__ void apply_kernel(T* __restrict__ rd, T* __restrict__ wr, uint64_t n, int op, T c,
                             int option, T none, bool write_all_none, int* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  bool hit = false;
  for (; i < n; i += stride) {
    T r = rd[i];
    if (option == KG_GRID_READ) {
      if (r != none) { T v = apply_op(op, r, c); hit |= v == none; rd[i] = v; }
    } else if (option == KG_GRID_WRITE) {
      if (r != none) { T v = apply_op(op, r, c); hit |= v == none; wr[i] = v; }
      else if (write_all_none) wr[i] = none;
    } else {
      T w = write_all_none ? none : wr[i];
      if (w != none) { T v = apply_op(op, w, c); hit |= v == none; wr[i] = v; }
      else if (r != none) { T v = apply_op(op, r, c); hit |= v == none; wr[i] = v; }
      else if (write_all_none) wr[i] = none;
    }
  }
  if (hit) atomicOr(err, DEV_ERR_SENTINEL);
}

template <class T>
__global__ void find_first_kernel(const T* __restrict__ buf, uint64_t n, T value,
                                  unsigned long long* first) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long best = ~0ull;
  for (; i < n; i += stride)
    if (buf[i] == value) { best = i; break; }
  if (best != ~0ull) atomicMin(first, best);
}
template <class T>
__global__ void count_none_kernel(const T* __restrict__ buf, uint64_t n, T none,
                                  unsigned long long* out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long c = 0;
  for (; i < n; i += stride) c += buf[i] == none;
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// Forest-Fire initial state (row K): tree with probability `density`, column 0 burning
template <class T>
__global__ void init_forest_kernel(T* __restrict__ buf, int32_t width, int32_t height, float density,
                                   uint64_t seed, T none) {
  uint64_t n = (uint64_t)width * (uint64_t)height;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0, DOMAIN_GRID, (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    bool tree = u01_f32(r.v[0]) < density;
    bool col0 = i < (uint64_t)height;
    buf[i] = tree ? (T)(col0 ? FF_BURNING : FF_GREEN) : none;
  }
}

// ------------------------------------------------------------------ K5 generic: one thread per cell
template <class T>
__global__ void forest_fire_generic_kernel(const T* __restrict__ rd, T* __restrict__ wr, int32_t width,
                                           int32_t height, T none, bool write_none) {
  uint64_t n = (uint64_t)width * (uint64_t)height;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t x = (int32_t)(i / (uint64_t)height), y = (int32_t)(i % (uint64_t)height);
  T v = rd[i];
  if (v == none) {
    if (write_none) wr[i] = none;
    return;
  }
  T next = v;
  if (v == (T)FF_GREEN) {
    bool fire = false;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy) {
        int nx = x + dx, ny = y + dy;
        if ((dx | dy) == 0 || nx < 0 || ny < 0 || nx >= width || ny >= height) continue;
        fire |= rd[(uint64_t)nx * height + ny] == (T)FF_BURNING;
      }
    if (fire) next = (T)FF_BURNING;
  } else if (v == (T)FF_BURNING) {
    next = (T)FF_BURNED;
  }
  wr[i] = next;
}

}  // namespace kg

// ====================================================================== handle + C ABI
using namespace kg;

struct kg_grid {
  int device = 0;
  cudaStream_t stream = nullptr;
  int32_t width = 0, height = 0;
  int elem = 1;
  uint32_t none = 0xFF;
  uint64_t ncells = 0;
  void* buf[2] = {nullptr, nullptr};
  int read = 0, write = 1;
  bool write_clear_pending = false;  // the write buffer is logically all-None, not yet filled
  int* d_err = nullptr;
  int* h_err = nullptr;
  void* scratch = nullptr;
  uint64_t scratch_bytes = 0;
  Profiler prof;
  Stopwatch watch;
  EventPool events;
};

namespace {

constexpr int kT = 256;
inline unsigned gblocks(uint64_t n, int t = kT) {
  return (unsigned)std::min<uint64_t>((n + t - 1) / t, (uint64_t)kNumSMs * 32);
}
inline unsigned gblocks_exact(uint64_t n, int t = kT) { return (unsigned)((n + t - 1) / t); }

int guse(kg_grid* g) {
  if (!g) return fail(KG_E_INVALID, "null grid handle");
  KG_CUDA(cudaSetDevice(g->device));
  return KG_OK;
}
int gsync_check(kg_grid* g) {
  KG_CUDA(cudaMemcpyAsync(g->h_err, g->d_err, sizeof(int), cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  if (*g->h_err) {
    const int e = *g->h_err;
    KG_CUDA(cudaMemsetAsync(g->d_err, 0, sizeof(int), g->stream));
    if (e & DEV_ERR_OOB)
      return fail(KG_E_OOB, "grid location outside width*height (reference: index out of bounds panic)");
    return fail(KG_E_INVALID, "apply_to_all_values produced the value reserved for Option::None (0x%x)", g->none);
  }
  return KG_OK;
}
int gscratch(kg_grid* g, uint64_t bytes) {
  if (g->scratch_bytes >= bytes) return KG_OK;
  if (g->scratch) cudaFree(g->scratch);
  g->scratch = nullptr;
  g->scratch_bytes = 0;
  KG_CUDA(cudaMalloc(&g->scratch, bytes + 256));
  g->scratch_bytes = bytes;
  return KG_OK;
}
#define GLAUNCH(g, kind, kernel, grid, block, ...)      \
  do {                                                  \
    (g)->prof.begin(kind, (g)->stream);                 \
    kernel<<<grid, block, 0, (g)->stream>>>(__VA_ARGS__); \
    (g)->prof.end((g)->stream);                         \
  } while (0)

void fill_none(kg_grid* g, void* p) {
  uint64_t bytes = g->ncells * g->elem;
  uint32_t w = g->elem == 1 ? g->none * 0x01010101u : g->elem == 2 ? g->none * 0x00010001u : g->none;
  uint64_t n16 = bytes / 16;
  if (n16)
    GLAUNCH(g, KG_K_MISC, fill16_kernel, gblocks(n16), kT, (uint4*)p, n16, make_uint4(w, w, w, w));
  uint64_t done = n16 * 16;
  if (done < bytes) {
    uint64_t rem = (bytes - done) / g->elem;
    if (g->elem == 1)
      GLAUNCH(g, KG_K_MISC, fill_kernel<uint8_t>, 1, kT, (uint8_t*)p + done, rem, (uint8_t)g->none);
    else if (g->elem == 2)
      GLAUNCH(g, KG_K_MISC, fill_kernel<uint16_t>, 1, kT, (uint16_t*)((uint8_t*)p + done), rem,
              (uint16_t)g->none);
    else
      GLAUNCH(g, KG_K_MISC, fill_kernel<uint32_t>, 1, kT, (uint32_t*)((uint8_t*)p + done), rem,
              (uint32_t)g->none);
  }
}
void materialise_write(kg_grid* g) {
  if (!g->write_clear_pending) return;
  fill_none(g, g->buf[g->write]);
  g->write_clear_pending = false;
}

template <class T>
int set_values_t(kg_grid* g, uint64_t n, const int32_t* x, const int32_t* y, const void* values,
                 bool remove) {
  uint64_t vb = remove ? 0 : n * sizeof(T);
  KG_TRY(gscratch(g, n * 8 + vb + 64));
  int32_t* dx = (int32_t*)g->scratch;
  int32_t* dy = dx + n;
  T* dv = (T*)(dy + n);
  KG_CUDA(cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, g->stream));
  KG_CUDA(cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, g->stream));
  if (!remove) KG_CUDA(cudaMemcpyAsync(dv, values, vb, cudaMemcpyHostToDevice, g->stream));
  GLAUNCH(g, KG_K_MISC, check_index_kernel, gblocks_exact(n), kT, g->height, g->ncells, n, dx, dy,
          g->d_err);
  KG_TRY(gsync_check(g));
  materialise_write(g);
  GLAUNCH(g, KG_K_MISC, set_values_kernel<T>, gblocks_exact(n), kT, (T*)g->buf[g->write], g->height,
          g->ncells, n, dx, dy, dv, (T)g->none, remove);
  return KG_OK;
}

template <class T>
int get_values_t(kg_grid* g, int which, uint64_t n, const int32_t* x, const int32_t* y, void* out) {
  KG_TRY(gscratch(g, n * 8 + n * sizeof(T) + 64));
  int32_t* dx = (int32_t*)g->scratch;
  int32_t* dy = dx + n;
  T* dv = (T*)(dy + n);
  KG_CUDA(cudaMemcpyAsync(dx, x, n * 4, cudaMemcpyHostToDevice, g->stream));
  KG_CUDA(cudaMemcpyAsync(dy, y, n * 4, cudaMemcpyHostToDevice, g->stream));
  bool all_none = which == KG_BUF_WRITE && g->write_clear_pending;
  const T* b = (const T*)g->buf[which == KG_BUF_READ ? g->read : g->write];
  GLAUNCH(g, KG_K_MISC, get_values_kernel<T>, gblocks_exact(n), kT, b, g->height, g->ncells, n, dx, dy,
          dv, (T)g->none, all_none, g->d_err);
  KG_CUDA(cudaMemcpyAsync(out, dv, n * sizeof(T), cudaMemcpyDeviceToHost, g->stream));
  return gsync_check(g);
}

template <class T>
void apply_t(kg_grid* g, int op, uint32_t operand, int option) {
  bool wan = g->write_clear_pending && option != KG_GRID_READ;
  GLAUNCH(g, KG_K_MISC, apply_kernel<T>, gblocks(g->ncells), kT, (T*)g->buf[g->read],
          (T*)g->buf[g->write], g->ncells, op, (T)operand, option, (T)g->none, wan, g->d_err);
  if (wan) g->write_clear_pending = false;
}

int step_stencil(kg_grid* g, int rule) {
  if (rule != KG_RULE_FOREST_FIRE) return fail(KG_E_INVALID, "unknown stencil rule %d", rule);
  bool write_none = g->write_clear_pending;
  const void* rd = g->buf[g->read];
  void* wr = g->buf[g->write];
  if (g->elem == 1 && g->none == 0xFF && g->height % 16 == 0) {
    const int rows = 64;  // rows per block: 2/64 = 3 % halo re-reads (32 and 64 measure equal, 128 slower)
    dim3 grid((unsigned)((g->height + 2047) / 2048), (unsigned)((g->width + rows - 1) / rows));
    if (write_none) {
      // dependent launch: in a run_stencil loop each step's launch overlaps the previous step's tail
      g->prof.begin(KG_K_STENCIL, g->stream);
      cudaError_t le = launch_pdl(forest_fire_u8_kernel<true>, grid, dim3(128), g->stream, (const uint8_t*)rd,
                                  (uint8_t*)wr, g->width, g->height, rows, FFExchange{});
      g->prof.end(g->stream);
      if (le != cudaSuccess) return fail(KG_E_CUDA, "launch of forest_fire_u8_kernel failed: %s", cudaGetErrorString(le));
    }
    else
      GLAUNCH(g, KG_K_STENCIL, forest_fire_u8_kernel<false>, grid, 128, (const uint8_t*)rd,
              (uint8_t*)wr, g->width, g->height, rows, FFExchange{});
  } else if (g->elem == 1) {
    GLAUNCH(g, KG_K_STENCIL, forest_fire_generic_kernel<uint8_t>, gblocks_exact(g->ncells), kT,
            (const uint8_t*)rd, (uint8_t*)wr, g->width, g->height, (uint8_t)g->none, write_none);
  } else if (g->elem == 2) {
    GLAUNCH(g, KG_K_STENCIL, forest_fire_generic_kernel<uint16_t>, gblocks_exact(g->ncells), kT,
            (const uint16_t*)rd, (uint16_t*)wr, g->width, g->height, (uint16_t)g->none, write_none);
  } else {
    GLAUNCH(g, KG_K_STENCIL, forest_fire_generic_kernel<uint32_t>, gblocks_exact(g->ncells), kT,
            (const uint32_t*)rd, (uint32_t*)wr, g->width, g->height, (uint32_t)g->none, write_none);
  }
  g->write_clear_pending = false;  // every cell of the write buffer now holds its value or None
  return KG_OK;
}

// T Forest-Fire steps in one pass over the grid (forest_fire_u8_multi_kernel): read buffer = step t,
// write buffer := step t+T.  Only where K5's fast path applies and the write buffer is due to be
// overwritten completely (the state every step of a run loop is in); the caller swaps ONCE afterwards.
// Returns the number of steps the next pass of a run with `left` steps to go should take.
int fused_steps(const kg_grid* g, int rule, uint64_t left) {
  static const int cap = getenv("KG_FF_FUSE") ? atoi(getenv("KG_FF_FUSE")) : kFFHalo;  // lab / test hook: 0 or 1 = never fuse
  if (rule != KG_RULE_FOREST_FIRE || g->elem != 1 || g->none != 0xFF || g->height % 16 != 0 || g->ncells == 0 ||
      !g->write_clear_pending)
    return 1;
  for (int t = 8; t >= 2; t >>= 1)
    if ((uint64_t)t <= left && t <= cap) return t;
  return 1;
}
// rows per tile of a T-step pass (stencil_device.cuh: ff_multi_rows_per_tile), or the lab hook's
int multi_rows_per_tile(int T, int64_t own, int64_t height) {
  static const int env = getenv("KG_FFT_ROWS") ? atoi(getenv("KG_FFT_ROWS")) : 0;  // lab hook
  if (env > 0) return std::max(env, kFFHalo);
  return ff_multi_rows_per_tile(T, own, height);
}
template <int T>
cudaError_t launch_multi(cudaStream_t st, dim3 grid, const uint8_t* rd, uint8_t* wr, int32_t width, int32_t height,
                         int32_t rows, const FFExchange& ex) {
  return launch_pdl(forest_fire_u8_multi_kernel<T>, grid, dim3(128), st, rd, wr, width, height, rows, ex);
}
int step_stencil_multi(kg_grid* g, int T) {
  const int rows = multi_rows_per_tile(T, g->width, g->height);
  const unsigned spans = (unsigned)((g->height + kFFTSpan - 1) / kFFTSpan);
  dim3 grid((spans + 3) / 4, (unsigned)((g->width + rows - 1) / rows));
  const uint8_t* rd = (const uint8_t*)g->buf[g->read];
  uint8_t* wr = (uint8_t*)g->buf[g->write];
  g->prof.begin(KG_K_STENCIL, g->stream);
  cudaError_t le = T == 8   ? launch_multi<8>(g->stream, grid, rd, wr, g->width, g->height, rows, FFExchange{})
                   : T == 4 ? launch_multi<4>(g->stream, grid, rd, wr, g->width, g->height, rows, FFExchange{})
                            : launch_multi<2>(g->stream, grid, rd, wr, g->width, g->height, rows, FFExchange{});
  g->prof.end(g->stream);
  if (le != cudaSuccess) return fail(KG_E_CUDA, "launch of forest_fire_u8_multi_kernel failed: %s", cudaGetErrorString(le));
  g->write_clear_pending = false;
  return KG_OK;
}

}  // namespace

extern "C" {

int kg_grid_create(int32_t width, int32_t height, int elem_size, uint32_t none, int device,
                   kg_grid** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (elem_size != 1 && elem_size != 2 && elem_size != 4) return fail(KG_E_INVALID, "elem_size must be 1, 2 or 4");
  if (elem_size < 4 && none >= (1u << (8 * elem_size))) return fail(KG_E_INVALID, "none sentinel does not fit the element");
  // the reference sizes the Vec with the signed i32 product (dense_number_grid_2d.rs:116-121)
  int64_t prod = (int64_t)(int32_t)((uint32_t)width * (uint32_t)height);
  if (prod < 0 || (int64_t)width * (int64_t)height != prod)
    return fail(KG_E_INVALID, "width*height overflows i32 (reference: capacity overflow panic)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_grid* g = new kg_grid();
  g->device = device;
  g->width = width < 0 ? -width : width;
  g->height = height < 0 ? -height : height;
  g->elem = elem_size;
  g->none = none;
  g->ncells = (uint64_t)prod;
  uint64_t bytes = g->ncells * elem_size + 256;
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&g->buf[0], bytes) != cudaSuccess || cudaMalloc(&g->buf[1], bytes) != cudaSuccess ||
      cudaMalloc(&g->d_err, sizeof(int)) != cudaSuccess ||
      cudaHostAlloc(&g->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    int rc = fail(KG_E_CUDA, "grid allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    kg_grid_destroy(g);
    return rc;
  }
  cudaMemsetAsync(g->d_err, 0, sizeof(int), g->stream);
  fill_none(g, g->buf[0]);
  fill_none(g, g->buf[1]);
  if (cudaStreamSynchronize(g->stream) != cudaSuccess) {
    int rc = fail(KG_E_CUDA, "grid init failed");
    kg_grid_destroy(g);
    return rc;
  }
  *out = g;
  return KG_OK;
}

int kg_grid_destroy(kg_grid* g) {
  if (!g) return KG_OK;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  g->prof.destroy();
  g->watch.destroy();
  g->events.destroy();
  cudaFree(g->buf[0]);
  cudaFree(g->buf[1]);
  cudaFree(g->d_err);
  cudaFree(g->scratch);
  if (g->h_err) cudaFreeHost(g->h_err);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return KG_OK;
}

int kg_grid_sync(kg_grid* g) {
  KG_TRY(guse(g));
  return gsync_check(g);
}

int kg_grid_set_values(kg_grid* g, uint64_t n, const int32_t* x, const int32_t* y, const void* values) {
  KG_TRY(guse(g));
  if (n == 0) return KG_OK;
  if (!x || !y || !values) return fail(KG_E_INVALID, "null argument");
  for (uint64_t i = 0; i < n; ++i) {  // Some(none) cannot be stored: use kg_grid_remove_values for None
    const uint32_t v = g->elem == 1 ? ((const uint8_t*)values)[i]
                     : g->elem == 2 ? ((const uint16_t*)values)[i] : ((const uint32_t*)values)[i];
    if (v == g->none) return fail(KG_E_INVALID, "value %u is the grid's Option::None sentinel", v);
  }
  if (g->elem == 1) return set_values_t<uint8_t>(g, n, x, y, values, false);
  if (g->elem == 2) return set_values_t<uint16_t>(g, n, x, y, values, false);
  return set_values_t<uint32_t>(g, n, x, y, values, false);
}
int kg_grid_remove_values(kg_grid* g, uint64_t n, const int32_t* x, const int32_t* y) {
  KG_TRY(guse(g));
  if (n == 0) return KG_OK;
  if (!x || !y) return fail(KG_E_INVALID, "null argument");
  if (g->elem == 1) return set_values_t<uint8_t>(g, n, x, y, nullptr, true);
  if (g->elem == 2) return set_values_t<uint16_t>(g, n, x, y, nullptr, true);
  return set_values_t<uint32_t>(g, n, x, y, nullptr, true);
}
int kg_grid_get_values(kg_grid* g, int which, uint64_t n, const int32_t* x, const int32_t* y, void* out) {
  KG_TRY(guse(g));
  if (n == 0) return KG_OK;
  if (!x || !y || !out) return fail(KG_E_INVALID, "null argument");
  if (g->elem == 1) return get_values_t<uint8_t>(g, which, n, x, y, out);
  if (g->elem == 2) return get_values_t<uint16_t>(g, which, n, x, y, out);
  return get_values_t<uint32_t>(g, which, n, x, y, out);
}

int kg_grid_upload(kg_grid* g, int which, const void* cells) {
  KG_TRY(guse(g));
  if (!cells) return fail(KG_E_INVALID, "null argument");
  int b = which == KG_BUF_READ ? g->read : g->write;
  KG_CUDA(cudaMemcpyAsync(g->buf[b], cells, g->ncells * g->elem, cudaMemcpyHostToDevice, g->stream));
  if (which == KG_BUF_WRITE) g->write_clear_pending = false;
  KG_CUDA(cudaStreamSynchronize(g->stream));
  return KG_OK;
}
int kg_grid_download(kg_grid* g, int which, void* cells) {
  KG_TRY(guse(g));
  if (!cells) return fail(KG_E_INVALID, "null argument");
  if (which == KG_BUF_WRITE) materialise_write(g);
  int b = which == KG_BUF_READ ? g->read : g->write;
  KG_CUDA(cudaMemcpyAsync(cells, g->buf[b], g->ncells * g->elem, cudaMemcpyDeviceToHost, g->stream));
  return gsync_check(g);
}

int kg_grid_apply(kg_grid* g, int op, uint32_t operand, int option) {
  KG_TRY(guse(g));
  if (op != KG_APPLY_CONST && op != KG_APPLY_ADD) return fail(KG_E_INVALID, "bad apply op");
  if (option < KG_GRID_READ || option > KG_GRID_READWRITE) return fail(KG_E_INVALID, "bad GridOption");
  if (g->ncells == 0) return KG_OK;
  if (op == KG_APPLY_CONST && operand == g->none)
    return fail(KG_E_INVALID, "apply: the constant %u is the grid's Option::None sentinel", operand);
  if (g->elem == 1) apply_t<uint8_t>(g, op, operand, option);
  else if (g->elem == 2) apply_t<uint16_t>(g, op, operand, option);
  else apply_t<uint32_t>(g, op, operand, option);
  return KG_OK;
}

// ---- run-time compiled closures (jit.cuh)
namespace {
const char* jit_type(int elem) { return elem == 1 ? "unsigned char" : elem == 2 ? "unsigned short" : "unsigned int"; }
// apply_to_all_values :155-195 with the closure's body as a CUDA C expression in `v` (and the cell x, y)
std::string jit_apply_source(int elem, const char* expr) {
  std::string t = jit_type(elem);
  return "typedef " + t + " T;\n"
         "__device__ __forceinline__ T closure(T v, int x, int y) { return (T)(" + std::string(expr) + "); }\n"
         "extern \"C\" __global__ void kg_closure(T* __restrict__ rd, T* __restrict__ wr, unsigned long long n,\n"
         "    int height, int option, T none, int write_all_none, int* err) {\n"
         "  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
         "  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;\n"
         "  bool hit = false;\n"
         "  for (; i < n; i += stride) {\n"
         "    const int x = (int)(i / (unsigned long long)height), y = (int)(i % (unsigned long long)height);\n"
         "    const T r = rd[i];\n"
         "    if (option == 0) {            /* READ :165-171 */\n"
         "      if (r != none) { T q = closure(r, x, y); hit |= q == none; rd[i] = q; }\n"
         "    } else if (option == 1) {     /* WRITE :173-180 */\n"
         "      if (r != none) { T q = closure(r, x, y); hit |= q == none; wr[i] = q; }\n"
         "      else if (write_all_none) wr[i] = none;\n"
         "    } else {                      /* READWRITE :183-194 */\n"
         "      const T w = write_all_none ? none : wr[i];\n"
         "      if (w != none) { T q = closure(w, x, y); hit |= q == none; wr[i] = q; }\n"
         "      else if (r != none) { T q = closure(r, x, y); hit |= q == none; wr[i] = q; }\n"
         "      else if (write_all_none) wr[i] = none;\n"
         "    }\n"
         "  }\n"
         "  if (hit) atomicOr(err, 2);\n"
         "}\n";
}
// a model's per-cell update rule through the field API (get_value of the Moore neighbourhood from the READ
// buffer, set_value_location into the WRITE buffer): the rule's body as a CUDA C expression in `v` (the
// cell's value), `at(dx, dy)` (a neighbour's value; NONE outside the grid or for an empty cell), x, y.
// Like the shipped rule, a cell that is None stays None.
std::string jit_rule_source(int elem, const char* expr) {
  std::string t = jit_type(elem);
  return "typedef " + t + " T;\n"
         "struct Nb { const T* rd; int x, y, w, h; T none;\n"
         "  __device__ __forceinline__ T operator()(int dx, int dy) const {\n"
         "    const int a = x + dx, b = y + dy;\n"
         "    return (a < 0 || b < 0 || a >= w || b >= h) ? none : rd[(unsigned long long)a * h + b]; } };\n"
         "__device__ __forceinline__ T rule(T v, const Nb& at, int x, int y, T NONE) { return (T)(" + std::string(expr) + "); }\n"
         "extern \"C\" __global__ void kg_rule(const T* __restrict__ rd, T* __restrict__ wr, int width, int height,\n"
         "    T none, int write_all_none, int* err) {\n"
         "  const unsigned long long n = (unsigned long long)width * height;\n"
         "  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
         "  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;\n"
         "  bool hit = false;\n"
         "  for (; i < n; i += stride) {\n"
         "    const int x = (int)(i / (unsigned long long)height), y = (int)(i % (unsigned long long)height);\n"
         "    const T v = rd[i];\n"
         "    if (v != none) { Nb at{rd, x, y, width, height, none}; T q = rule(v, at, x, y, none); hit |= q == none; wr[i] = q; }\n"
         "    else if (write_all_none) wr[i] = none;   /* only live cells write (set_value_location per live cell) */\n"
         "  }\n"
         "  if (hit) atomicOr(err, 2);\n"
         "}\n";
}
}  // namespace

int kg_grid_apply_expr(kg_grid* g, const char* expr, int option) {
  KG_TRY(guse(g));
  if (!expr || !*expr) return fail(KG_E_INVALID, "empty closure");
  if (option < KG_GRID_READ || option > KG_GRID_READWRITE) return fail(KG_E_INVALID, "bad GridOption");
  KG_CUDA(cudaFree(nullptr));  // the runtime's primary context is current for the driver calls below
  CUfunction fn;
  KG_TRY(jit::get_kernel(g->device, jit_apply_source(g->elem, expr), "kg_closure", &fn));
  if (g->ncells == 0) return KG_OK;
  void* rd = g->buf[g->read];
  void* wr = g->buf[g->write];
  unsigned long long n = g->ncells;
  int height = g->height, opt = option, wan = g->write_clear_pending && option != KG_GRID_READ ? 1 : 0;
  uint32_t none32 = g->none;
  uint16_t none16 = (uint16_t)g->none;
  uint8_t none8 = (uint8_t)g->none;
  void* none = g->elem == 1 ? (void*)&none8 : g->elem == 2 ? (void*)&none16 : (void*)&none32;
  int* err = g->d_err;
  void* args[] = {&rd, &wr, &n, &height, &opt, none, &wan, &err};
  g->prof.begin(KG_K_MISC, g->stream);
  const int rc = jit::launch(fn, gblocks(g->ncells), kT, g->stream, args);
  g->prof.end(g->stream);
  KG_TRY(rc);
  if (wan) g->write_clear_pending = false;
  return KG_OK;
}

int kg_grid_step_expr(kg_grid* g, const char* expr) {
  KG_TRY(guse(g));
  if (!expr || !*expr) return fail(KG_E_INVALID, "empty rule");
  KG_CUDA(cudaFree(nullptr));
  CUfunction fn;
  KG_TRY(jit::get_kernel(g->device, jit_rule_source(g->elem, expr), "kg_rule", &fn));
  if (g->ncells == 0) return KG_OK;
  const void* rd = g->buf[g->read];
  void* wr = g->buf[g->write];
  int width = g->width, height = g->height;
  uint32_t none32 = g->none;
  uint16_t none16 = (uint16_t)g->none;
  uint8_t none8 = (uint8_t)g->none;
  void* none = g->elem == 1 ? (void*)&none8 : g->elem == 2 ? (void*)&none16 : (void*)&none32;
  int* err = g->d_err;
  int wan = g->write_clear_pending ? 1 : 0;
  void* args[] = {&rd, &wr, &width, &height, none, &wan, &err};
  g->prof.begin(KG_K_STENCIL, g->stream);
  const int rc = jit::launch(fn, gblocks(g->ncells), kT, g->stream, args);
  g->prof.end(g->stream);
  KG_TRY(rc);
  g->write_clear_pending = false;  // the pending clear was materialised by this kernel
  return KG_OK;
}

int kg_grid_get_location(kg_grid* g, int which, uint32_t value, int32_t* x, int32_t* y, int* found) {
  KG_TRY(guse(g));
  if (!x || !y || !found) return fail(KG_E_INVALID, "null argument");
  *found = 0;
  if (g->ncells == 0) return KG_OK;
  if (value == g->none) return KG_OK;  // get_location matches Some(value) only (`elem.is_some() && elem.unwrap() == value`, dense_number_grid_2d.rs:222, :260)
  if (which == KG_BUF_WRITE) materialise_write(g);
  KG_TRY(gscratch(g, 64));
  unsigned long long* d = (unsigned long long*)g->scratch;
  KG_CUDA(cudaMemsetAsync(d, 0xFF, 8, g->stream));
  const void* b = g->buf[which == KG_BUF_READ ? g->read : g->write];
  if (g->elem == 1)
    GLAUNCH(g, KG_K_MISC, find_first_kernel<uint8_t>, gblocks(g->ncells), kT, (const uint8_t*)b, g->ncells, (uint8_t)value, d);
  else if (g->elem == 2)
    GLAUNCH(g, KG_K_MISC, find_first_kernel<uint16_t>, gblocks(g->ncells), kT, (const uint16_t*)b, g->ncells, (uint16_t)value, d);
  else
    GLAUNCH(g, KG_K_MISC, find_first_kernel<uint32_t>, gblocks(g->ncells), kT, (const uint32_t*)b, g->ncells, value, d);
  unsigned long long h = 0;
  KG_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g->stream));
  KG_TRY(gsync_check(g));
  if (h != ~0ull) {
    *found = 1;
    *x = (int32_t)(h / (uint64_t)g->height);
    *y = (int32_t)(h % (uint64_t)g->height);
  }
  return KG_OK;
}

int kg_grid_num_empty(kg_grid* g, uint64_t* out) {
  KG_TRY(guse(g));
  if (!out) return fail(KG_E_INVALID, "null argument");
  *out = 0;
  if (g->ncells == 0) return KG_OK;
  KG_TRY(gscratch(g, 64));
  unsigned long long* d = (unsigned long long*)g->scratch;
  KG_CUDA(cudaMemsetAsync(d, 0, 8, g->stream));
  const void* b = g->buf[g->read];
  if (g->elem == 1)
    GLAUNCH(g, KG_K_MISC, count_none_kernel<uint8_t>, gblocks(g->ncells), kT, (const uint8_t*)b, g->ncells, (uint8_t)g->none, d);
  else if (g->elem == 2)
    GLAUNCH(g, KG_K_MISC, count_none_kernel<uint16_t>, gblocks(g->ncells), kT, (const uint16_t*)b, g->ncells, (uint16_t)g->none, d);
  else
    GLAUNCH(g, KG_K_MISC, count_none_kernel<uint32_t>, gblocks(g->ncells), kT, (const uint32_t*)b, g->ncells, g->none, d);
  unsigned long long h = 0;
  KG_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g->stream));
  KG_TRY(gsync_check(g));
  *out = h;
  return KG_OK;
}

int kg_grid_lazy_update(kg_grid* g) {
  KG_TRY(guse(g));
  // a write buffer nobody touched since the last swap is logically all-None and is about to
  // become readable: fill it now (two swaps in a row, dense_number_grid_2d.rs:541-543)
  materialise_write(g);
  std::swap(g->read, g->write);
  g->write_clear_pending = true;
  return KG_OK;
}
int kg_grid_update(kg_grid* g) {
  KG_TRY(guse(g));
  // copy write -> read, then write := None  (dense_number_grid_2d.rs:553-561)
  if (g->write_clear_pending) {
    fill_none(g, g->buf[g->read]);
  } else {
    KG_CUDA(cudaMemcpyAsync(g->buf[g->read], g->buf[g->write], g->ncells * g->elem,
                            cudaMemcpyDeviceToDevice, g->stream));
    g->write_clear_pending = true;
  }
  return KG_OK;
}

int kg_grid_step_stencil(kg_grid* g, int rule) {
  KG_TRY(guse(g));
  if (g->ncells == 0) return KG_OK;
  return step_stencil(g, rule);
}
int kg_grid_run_stencil(kg_grid* g, int rule, uint64_t nsteps) {
  KG_TRY(guse(g));
  for (uint64_t i = 0; i < nsteps;) {
    // runs of 8 / 4 / 2 steps go through the fused kernel: the read buffer holds step i+T after ONE swap,
    // the write buffer is "all None" either way — the same observable state as T single steps
    const int T = fused_steps(g, rule, nsteps - i);
    if (T > 1) KG_TRY(step_stencil_multi(g, T));
    else if (g->ncells) KG_TRY(step_stencil(g, rule));
    std::swap(g->read, g->write);
    g->write_clear_pending = true;
    i += (uint64_t)T;
  }
  return KG_OK;
}

int kg_grid_init_forest_fire(kg_grid* g, float density, uint64_t seed) {
  KG_TRY(guse(g));
  if (g->ncells) {
    void* wr = g->buf[g->write];
    if (g->elem == 1)
      GLAUNCH(g, KG_K_MISC, init_forest_kernel<uint8_t>, gblocks(g->ncells), kT, (uint8_t*)wr, g->width, g->height, density, seed, (uint8_t)g->none);
    else if (g->elem == 2)
      GLAUNCH(g, KG_K_MISC, init_forest_kernel<uint16_t>, gblocks(g->ncells), kT, (uint16_t*)wr, g->width, g->height, density, seed, (uint16_t)g->none);
    else
      GLAUNCH(g, KG_K_MISC, init_forest_kernel<uint32_t>, gblocks(g->ncells), kT, (uint32_t*)wr, g->width, g->height, density, seed, g->none);
  }
  g->write_clear_pending = false;
  std::swap(g->read, g->write);
  g->write_clear_pending = true;
  return KG_OK;
}

int kg_grid_run_stencil_timed(kg_grid* g, int rule, uint64_t nsteps, double* ms_sum) {
  KG_TRY(guse(g));
  if (!ms_sum) return fail(KG_E_INVALID, "null argument");
  uint64_t passes = 0;  // one event pair per launch: a fused pass advances up to eight steps
  for (uint64_t i = 0; i < nsteps; ++passes) {
    cudaEvent_t a = nullptr, b = nullptr;
    KG_TRY(g->events.get(2 * passes, &a));
    KG_TRY(g->events.get(2 * passes + 1, &b));
    KG_CUDA(cudaEventRecord(a, g->stream));
    const int T = fused_steps(g, rule, nsteps - i);
    if (T > 1) KG_TRY(step_stencil_multi(g, T));
    else if (g->ncells) KG_TRY(step_stencil(g, rule));
    std::swap(g->read, g->write);
    g->write_clear_pending = true;
    KG_CUDA(cudaEventRecord(b, g->stream));
    i += (uint64_t)T;
  }
  KG_TRY(gsync_check(g));
  double sum = 0;
  for (uint64_t i = 0; i < passes; ++i) {
    float t = 0.f;
    KG_CUDA(cudaEventElapsedTime(&t, g->events.ev[2 * i], g->events.ev[2 * i + 1]));
    sum += t;
  }
  *ms_sum = sum;
  return KG_OK;
}

int kg_grid_timer_start(kg_grid* g) {
  KG_TRY(guse(g));
  return g->watch.start(g->stream);
}
int kg_grid_timer_stop(kg_grid* g, double* ms) {
  KG_TRY(guse(g));
  return g->watch.stop(g->stream, ms);
}

int kg_grid_profile(kg_grid* g, int enable) {
  KG_TRY(guse(g));
  g->prof.drain();
  g->prof.enabled = enable != 0;
  return KG_OK;
}
int kg_grid_profile_read(kg_grid* g, double* ms, uint64_t* launches, int reset) {
  KG_TRY(guse(g));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  g->prof.drain();
  for (int k = 0; k < KG_K_COUNT; ++k) {
    if (ms) ms[k] = g->prof.ms[k];
    if (launches) launches[k] = g->prof.launches[k];
    if (reset) {
      g->prof.ms[k] = 0;
      g->prof.launches[k] = 0;
    }
  }
  return KG_OK;
}

}  // extern "C"
