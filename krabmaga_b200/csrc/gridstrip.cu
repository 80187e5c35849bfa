// Multi-GPU DenseNumberGrid2D<u8> for stencil models (Forest Fire, BASELINE config 4): the grid is
// cut into strips of whole x rows (a row = `height` contiguous bytes, index x*height + y as in
// dense_number_grid_2d.rs:351), one kg_gridstrip per GPU.
//
// One kernel per step per GPU (stencil_device.cuh): the blocks that compute a strip's first / last
// row also store it into the line neighbour's inbox with peer stores over NVLink and publish an
// epoch flag behind a system-scope fence; the blocks that need the neighbour's row park on that
// flag before loading it.  No host round trip, no separate pack / wait / unpack launches, no
// collective.  The grid is not toroidal (out-of-range neighbours do not exist), so the halo
// topology is a line.  Inbox slots are double-buffered by step parity: a neighbour can be at most
// one step ahead.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "stencil_device.cuh"

namespace kg {

// Forest-Fire initial state: identical to the single-GPU init because the
// Philox counter is the GLOBAL cell index
__global__ void gs_init_forest_kernel(uint8_t* __restrict__ buf, int32_t x0, int32_t own, int32_t height,
                                      float density, uint64_t seed) {
  uint64_t n = (uint64_t)own * (uint64_t)height;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint64_t gi = i + (uint64_t)x0 * (uint64_t)height;
    Philox4 r = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), 0, DOMAIN_GRID, (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    bool tree = u01_f32(r.v[0]) < density;
    bool col0 = gi < (uint64_t)height;
    buf[i] = tree ? (uint8_t)(col0 ? FF_BURNING : FF_GREEN) : (uint8_t)0xFF;
  }
}

// prepare(): hand the current boundary rows (`bytes` contiguous bytes) to a neighbour (one block)
__global__ void gs_push_row_kernel(const uint8_t* __restrict__ src, uint8_t* dst, int64_t bytes,
                                   unsigned long long* flag, unsigned long long epoch) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  for (int64_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d[i] = s[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    *(volatile unsigned long long*)flag = epoch;
    __threadfence_system();
  }
}

}  // namespace kg

using namespace kg;

struct kg_gridstrip {
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  int32_t width = 0, height = 0;  // global grid
  int32_t x0 = 0, x1 = 0;         // owned rows
  uint8_t* buf[2] = {nullptr, nullptr};
  int read = 0, write = 1;
  size_t slot_bytes = 0;          // 256-byte header (flag) + kFFHalo rows, `height` bytes apart
  void* inbox = nullptr;          // 4 slots: (from_left, from_right) x parity
  void* peer_inbox[2] = {nullptr, nullptr};  // left, right line neighbours
  bool peer_is_ipc[2] = {false, false};
  uint32_t* d_done = nullptr;
  int* d_err = nullptr;
  int* h_err = nullptr;
  unsigned long long steps_done = 0;  // passes done (a pass = one launch = one or two steps): epoch base; parity = steps_done & 1
  bool prepared = false;
  Stopwatch watch;
  EventPool events;
};

namespace {

int gsuse(kg_gridstrip* s) {
  if (!s) return fail(KG_E_INVALID, "null grid strip handle");
  KG_CUDA(cudaSetDevice(s->device));
  return KG_OK;
}
#define GSLAUNCH(s, kernel, grid, block, ...)                                             \
  do {                                                                                    \
    kernel<<<grid, block, 0, (s)->stream>>>(__VA_ARGS__);                                 \
    cudaError_t _le = cudaGetLastError();                                                 \
    if (_le != cudaSuccess)                                                               \
      return fail(KG_E_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_le)); \
    launch_counter().fetch_add(1, std::memory_order_relaxed);                             \
  } while (0)

int gs_sync_check(kg_gridstrip* s) {
  KG_CUDA(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  KG_CUDA(cudaStreamSynchronize(s->stream));
  if (*s->h_err) {
    KG_CUDA(cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream));
    return fail(KG_E_CUDA, "grid strip %d: neighbour flag timed out (peer not stepping?)", s->rank);
  }
  return KG_OK;
}

struct Slot {
  unsigned long long* flag;
  uint8_t* row;
};
inline Slot slot_of(void* base, size_t slot_bytes, int dir, int parity) {
  char* p = (char*)base + (size_t)(dir * 2 + parity) * slot_bytes;
  return Slot{(unsigned long long*)p, (uint8_t*)(p + 256)};
}
inline bool has_left(const kg_gridstrip* s) { return s->rank > 0; }
inline bool has_right(const kg_gridstrip* s) { return s->rank < s->nranks - 1; }

// A pass of T steps needs T rows of every neighbour, so every strip must own at least T; all ranks
// evaluate this the same way (the smallest strip has width / nranks rows).
int gs_fused_steps_of(int32_t width, int nranks, uint64_t left) {
  static const int cap = getenv("KG_FF_FUSE") ? atoi(getenv("KG_FF_FUSE")) : kFFHalo;  // lab / test hook: 0 or 1 = never fuse
  for (int t = 8; t >= 2; t >>= 1)
    if ((uint64_t)t <= left && t <= cap && width / nranks >= t) return t;
  return 1;
}
int gs_fused_steps(const kg_gridstrip* s, uint64_t left) { return gs_fused_steps_of(s->width, s->nranks, left); }
// rows per tile of a pass of T steps over `own` rows (T == 1: K5's tiles), with the guarantee the exchange
// needs: the last tile never holds fewer than kFFHalo rows
int gs_rows_per_tile(int T, int32_t own, int32_t height) {
  static const int rows_env = getenv("KG_FF_ROWS") ? atoi(getenv("KG_FF_ROWS")) : 64;  // lab hooks
  static const int rows_env_t = getenv("KG_FFT_ROWS") ? atoi(getenv("KG_FFT_ROWS")) : 0;
  int rows;
  if (T == 1) rows = std::max(kFFHalo, rows_env);
  else if (rows_env_t > 0) rows = std::max(kFFHalo, rows_env_t);
  else rows = ff_multi_rows_per_tile(T, own, height);
  while (own > rows && own % rows != 0 && own % rows < kFFHalo) ++rows;  // ends at rows == own at the latest
  return rows;
}

template <int T>
cudaError_t gs_launch_multi(cudaStream_t st, dim3 grid, const uint8_t* rd, uint8_t* wr, int32_t width, int32_t height,
                            int32_t rows, const FFExchange& ex) {
  return launch_pdl(forest_fire_u8_multi_kernel<T>, grid, dim3(128), st, rd, wr, width, height, rows, ex);
}

// one pass of T steps: T = 1: K5; 2, 4, 8: forest_fire_u8_multi_kernel
int gs_step(kg_gridstrip* s, int T) {
  if (!s->prepared) return fail(KG_E_INVALID, "grid strip not prepared (call kg_gridstrip_prepare on every rank)");
  const int32_t own = s->x1 - s->x0;
  const unsigned long long t = s->steps_done;
  const int rp = (int)(t & 1), wp = (int)((t + 1) & 1);
  FFExchange ex;
  ex.wait_epoch = t + 1;
  ex.push_epoch = t + 2;
  ex.done = s->d_done;
  ex.err = s->d_err;
  if (has_left(s)) {
    Slot in = slot_of(s->inbox, s->slot_bytes, 0, rp);  // rows -kFFHalo .. -1
    ex.slot_lo = in.row;
    ex.flag_lo = in.flag;
    Slot out = slot_of(s->peer_inbox[0], s->slot_bytes, 1, wp);  // I am the left one's right neighbour
    ex.push_lo = out.row;
    ex.push_flag_lo = out.flag;
  }
  if (has_right(s)) {
    Slot in = slot_of(s->inbox, s->slot_bytes, 1, rp);  // rows own .. own + kFFHalo - 1
    ex.slot_hi = in.row;
    ex.flag_hi = in.flag;
    Slot out = slot_of(s->peer_inbox[1], s->slot_bytes, 0, wp);
    ex.push_hi = out.row;
    ex.push_flag_hi = out.flag;
  }
  // Rows per tile.  K5: 2/64 = 3 % halo re-reads (32 and 64 measure equal, 128 slower); T steps: enough
  // warps to fill the GPU, 2T / rows of redundant work.  All rows pushed to the right neighbour must come
  // from the LAST row tile (its blocks publish the flag): it never holds fewer than kFFHalo rows.
  const int rows = gs_rows_per_tile(T, own, s->height);
  const uint8_t* rd = s->buf[s->read];
  uint8_t* wr = s->buf[s->write];
  cudaError_t le;
  if (T == 1) {
    dim3 grid((unsigned)((s->height + 2047) / 2048), (unsigned)((own + rows - 1) / rows));
    le = launch_pdl(forest_fire_u8_kernel<true>, grid, dim3(128), s->stream, rd, wr, own, s->height, rows, ex);
  } else {
    const unsigned spans = (unsigned)((s->height + kFFTSpan - 1) / kFFTSpan);
    dim3 grid((spans + 3) / 4, (unsigned)((own + rows - 1) / rows));
    le = T == 8   ? gs_launch_multi<8>(s->stream, grid, rd, wr, own, s->height, rows, ex)
         : T == 4 ? gs_launch_multi<4>(s->stream, grid, rd, wr, own, s->height, rows, ex)
                  : gs_launch_multi<2>(s->stream, grid, rd, wr, own, s->height, rows, ex);
  }
  if (le != cudaSuccess) return fail(KG_E_CUDA, "launch of the Forest-Fire kernel (%d steps) failed: %s", T, cudaGetErrorString(le));
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  std::swap(s->read, s->write);
  s->steps_done += 1;
  return KG_OK;
}

// nsteps steps as passes of 8 / 4 / 2 / 1
int gs_run(kg_gridstrip* s, uint64_t nsteps) {
  for (uint64_t i = 0; i < nsteps;) {
    const int T = gs_fused_steps(s, nsteps - i);
    KG_TRY(gs_step(s, T));
    i += (uint64_t)T;
  }
  return KG_OK;
}

}  // namespace

extern "C" {

int kg_gridstrip_create(int32_t width, int32_t height, int rank, int nranks, int device,
                        kg_gridstrip** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (width <= 0 || height <= 0 || height % 16 != 0)
    return fail(KG_E_INVALID, "grid strips need width > 0 and height a positive multiple of 16");
  if (nranks < 1 || rank < 0 || rank >= nranks || nranks > width)
    return fail(KG_E_INVALID, "bad rank %d of %d for %d rows", rank, nranks, width);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_gridstrip* s = new kg_gridstrip();
  s->device = device; s->rank = rank; s->nranks = nranks;
  s->width = width; s->height = height;
  s->x0 = (int32_t)((int64_t)rank * width / nranks);
  s->x1 = (int32_t)((int64_t)(rank + 1) * width / nranks);
  const size_t bytes = (size_t)(s->x1 - s->x0) * (size_t)height;
  s->slot_bytes = 256 + ((size_t)kFFHalo * (size_t)height + 255) / 256 * 256;
  auto bail = [&](int code) { kg_gridstrip_destroy(s); return code; };
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  // load the step path's kernels now: a first launch may otherwise wait for an idle device while a
  // neighbour strip of the same process is parked on a flag (see strip.cu)
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, forest_fire_u8_kernel<true>);
  cudaFuncGetAttributes(&fa, forest_fire_u8_multi_kernel<2>);
  cudaFuncGetAttributes(&fa, forest_fire_u8_multi_kernel<4>);
  cudaFuncGetAttributes(&fa, forest_fire_u8_multi_kernel<8>);
  cudaFuncGetAttributes(&fa, gs_push_row_kernel);
  cudaFuncGetAttributes(&fa, gs_init_forest_kernel);
  if (cudaMalloc(&s->buf[0], bytes + 64) != cudaSuccess || cudaMalloc(&s->buf[1], bytes + 64) != cudaSuccess ||
      cudaMalloc(&s->inbox, 4 * s->slot_bytes) != cudaSuccess ||
      cudaMalloc(&s->d_done, 2 * sizeof(uint32_t)) != cudaSuccess ||
      cudaMalloc(&s->d_err, sizeof(int)) != cudaSuccess ||
      cudaHostAlloc(&s->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess)
    return bail(fail(KG_E_CUDA, "grid strip allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaMemsetAsync(s->buf[0], 0xFF, bytes, s->stream);
  cudaMemsetAsync(s->buf[1], 0xFF, bytes, s->stream);
  cudaMemsetAsync(s->inbox, 0, 4 * s->slot_bytes, s->stream);
  cudaMemsetAsync(s->d_done, 0, 2 * sizeof(uint32_t), s->stream);
  cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream);
  if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(KG_E_CUDA, "grid strip init failed"));
  s->prepared = nranks == 1;
  *out = s;
  return KG_OK;
}

int kg_gridstrip_destroy(kg_gridstrip* s) {
  if (!s) return KG_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  for (int k = 0; k < 2; ++k)
    if (s->peer_inbox[k] && s->peer_is_ipc[k]) cudaIpcCloseMemHandle(s->peer_inbox[k]);
  s->watch.destroy();
  s->events.destroy();
  cudaFree(s->buf[0]);
  cudaFree(s->buf[1]);
  cudaFree(s->inbox);
  cudaFree(s->d_done);
  cudaFree(s->d_err);
  if (s->h_err) cudaFreeHost(s->h_err);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return KG_OK;
}

int kg_gridstrip_pass_plan(int32_t width, int32_t height, int nranks, uint64_t nsteps, int32_t* steps_of_pass,
                           int32_t* rows_per_tile, uint64_t cap, uint64_t* npasses) {
  if (width <= 0 || height <= 0 || nranks < 1 || nranks > width || !npasses)
    return fail(KG_E_INVALID, "bad arguments");
  const int32_t own_min = width / nranks;  // every rank plans with the smallest strip
  uint64_t k = 0;
  for (uint64_t i = 0; i < nsteps; ++k) {
    const int T = gs_fused_steps_of(width, nranks, nsteps - i);
    if (k < cap) {
      if (steps_of_pass) steps_of_pass[k] = T;
      if (rows_per_tile) rows_per_tile[k] = gs_rows_per_tile(T, own_min, height);
    }
    i += (uint64_t)T;
  }
  *npasses = k;
  return KG_OK;
}

int kg_gridstrip_rows(kg_gridstrip* s, int32_t* x0, int32_t* x1) {
  if (!s) return fail(KG_E_INVALID, "null grid strip handle");
  if (x0) *x0 = s->x0;
  if (x1) *x1 = s->x1;
  return KG_OK;
}

int kg_gridstrip_ipc_export(kg_gridstrip* s, void* handle64) {
  KG_TRY(gsuse(s));
  if (!handle64) return fail(KG_E_INVALID, "null handle buffer");
  cudaIpcMemHandle_t h;
  KG_CUDA(cudaIpcGetMemHandle(&h, s->inbox));
  memcpy(handle64, &h, sizeof(h));
  return KG_OK;
}

int kg_gridstrip_connect_ipc(kg_gridstrip* s, const void* left_handle64, const void* right_handle64) {
  KG_TRY(gsuse(s));
  const void* hs[2] = {left_handle64, right_handle64};
  const bool need[2] = {has_left(s), has_right(s)};
  for (int k = 0; k < 2; ++k) {
    if (!need[k]) continue;
    if (!hs[k]) return fail(KG_E_INVALID, "missing IPC handle of the %s neighbour", k ? "right" : "left");
    cudaIpcMemHandle_t h;
    memcpy(&h, hs[k], 64);
    KG_CUDA(cudaIpcOpenMemHandle(&s->peer_inbox[k], h, cudaIpcMemLazyEnablePeerAccess));
    s->peer_is_ipc[k] = true;
  }
  return KG_OK;
}

int kg_gridstrip_connect_local(kg_gridstrip* s, kg_gridstrip* left, kg_gridstrip* right) {
  KG_TRY(gsuse(s));
  kg_gridstrip* nb[2] = {left, right};
  const bool need[2] = {has_left(s), has_right(s)};
  for (int k = 0; k < 2; ++k) {
    if (!need[k]) continue;
    if (!nb[k]) return fail(KG_E_INVALID, "missing %s neighbour", k ? "right" : "left");
    if (nb[k]->device != s->device) {
      int can = 0;
      KG_CUDA(cudaDeviceCanAccessPeer(&can, s->device, nb[k]->device));
      if (!can) return fail(KG_E_CUDA, "device %d cannot access peer %d", s->device, nb[k]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(nb[k]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(KG_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    s->peer_inbox[k] = nb[k]->inbox;
    s->peer_is_ipc[k] = false;
  }
  return KG_OK;
}

int kg_gridstrip_init_forest_fire(kg_gridstrip* s, float density, uint64_t seed) {
  KG_TRY(gsuse(s));
  const int32_t own = s->x1 - s->x0;
  GSLAUNCH(s, gs_init_forest_kernel, 148 * 8, 256, s->buf[s->read], s->x0, own, s->height, density, seed);
  s->prepared = s->nranks == 1;
  return gs_sync_check(s);
}

int kg_gridstrip_upload(kg_gridstrip* s, const uint8_t* own_rows) {
  KG_TRY(gsuse(s));
  if (!own_rows) return fail(KG_E_INVALID, "null input");
  const size_t bytes = (size_t)(s->x1 - s->x0) * (size_t)s->height;
  KG_CUDA(cudaMemcpyAsync(s->buf[s->read], own_rows, bytes, cudaMemcpyHostToDevice, s->stream));
  s->prepared = s->nranks == 1;
  return gs_sync_check(s);
}

int kg_gridstrip_download(kg_gridstrip* s, uint8_t* own_rows) {
  KG_TRY(gsuse(s));
  if (!own_rows) return fail(KG_E_INVALID, "null output");
  const size_t bytes = (size_t)(s->x1 - s->x0) * (size_t)s->height;
  KG_CUDA(cudaMemcpyAsync(own_rows, s->buf[s->read], bytes, cudaMemcpyDeviceToHost, s->stream));
  return gs_sync_check(s);
}

int kg_gridstrip_prepare(kg_gridstrip* s) {
  KG_TRY(gsuse(s));
  if (s->nranks > 1 && ((has_left(s) && !s->peer_inbox[0]) || (has_right(s) && !s->peer_inbox[1])))
    return fail(KG_E_INVALID, "grid strip is not connected to its neighbours");
  // flags published before this point describe an older state: move the epoch base past them
  // (by two, keeping the slot parity) — every rank calls prepare the same number of times
  s->steps_done += 2;
  const unsigned long long t = s->steps_done;
  const int rp = (int)(t & 1);
  const int32_t own = s->x1 - s->x0;
  const uint8_t* rd = s->buf[s->read];
  const int32_t nb = std::min(kFFHalo, own);  // boundary rows handed over per side
  if (has_left(s)) {  // my rows 0 .. nb-1 = the left neighbour's rows own .. own+nb-1
    Slot out = slot_of(s->peer_inbox[0], s->slot_bytes, 1, rp);
    GSLAUNCH(s, gs_push_row_kernel, 1, 256, rd, out.row, (int64_t)nb * s->height, out.flag, t + 1);
  }
  if (has_right(s)) {  // my rows own-nb .. own-1 = the right neighbour's rows -nb .. -1
    Slot out = slot_of(s->peer_inbox[1], s->slot_bytes, 0, rp);
    GSLAUNCH(s, gs_push_row_kernel, 1, 256, rd + (size_t)(own - nb) * s->height,
             out.row + (size_t)(kFFHalo - nb) * s->height, (int64_t)nb * s->height, out.flag, t + 1);
  }
  s->prepared = true;
  return gs_sync_check(s);
}

int kg_gridstrip_run_stencil(kg_gridstrip* s, int rule, uint64_t nsteps) {
  KG_TRY(gsuse(s));
  if (rule != KG_RULE_FOREST_FIRE) return fail(KG_E_INVALID, "unknown stencil rule %d", rule);
  return gs_run(s, nsteps);
}

int kg_gridstrip_run_stencil_timed(kg_gridstrip* s, int rule, uint64_t nsteps, double* ms_total) {
  KG_TRY(gsuse(s));
  if (!ms_total) return fail(KG_E_INVALID, "null argument");
  if (rule != KG_RULE_FOREST_FIRE) return fail(KG_E_INVALID, "unknown stencil rule %d", rule);
  cudaEvent_t a = nullptr, b = nullptr;
  KG_TRY(s->events.get(0, &a));
  KG_TRY(s->events.get(1, &b));
  KG_CUDA(cudaEventRecord(a, s->stream));
  KG_TRY(gs_run(s, nsteps));
  KG_CUDA(cudaEventRecord(b, s->stream));
  KG_TRY(gs_sync_check(s));
  float t = 0.f;
  KG_CUDA(cudaEventElapsedTime(&t, a, b));
  *ms_total = t;
  return KG_OK;
}

int kg_gridstrip_sync(kg_gridstrip* s) {
  KG_TRY(gsuse(s));
  return gs_sync_check(s);
}

}  // extern "C"
