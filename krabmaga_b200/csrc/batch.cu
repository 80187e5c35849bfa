// Batched independent replicas of one Flockers world for the `explore` sweeps.
//
// Replaces the replica fan-out of explore_parallel! (src/explore/model_exploration.rs:354-423:
// one State + Schedule per rayon task, simulate_explore! :160-190 inside each) for models whose
// step is the shipped boids kernel: R replicas of identical geometry and population size advance
// together, one launch per phase for the whole batch, no communication between replicas.
//
// Layout: the batch is ONE cell list over R*C cells (global cell = r*C + x*dh + y) and one pair of
// agent buffers of R*n entries.  Agents never change replica, K4 writes log entry i from sorted
// entry i, and the global cell order is replica-major, so replica r owns indices [r*n, (r+1)*n) of
// both buffers at every step: the replica of an index is i / n, no per-agent replica id is stored.
// Each replica has its own KgBoidsParams (weights, seed: the swept inputs) in a device array.
#include <algorithm>
#include <vector>

#include "boids_device.cuh"
#include "common.cuh"
#include "reduce.cuh"
#include "scan.cuh"

namespace kg {

// Flocker::init (state.rs:41-56) of every replica with its own Philox key
__global__ void batch_init_kernel(Geom g, uint32_t rep_n, uint64_t total,
                                  const KgBoidsParams* __restrict__ params, Agents wr,
                                  uint32_t* __restrict__ count, int* err) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t r = (uint32_t)(t / rep_n), id = (uint32_t)(t - (uint64_t)r * rep_n);
  const uint64_t seed = params[r].seed;
  Philox4 ph = philox4x32_10(id, 0, 0, DOMAIN_INIT, (uint32_t)seed, (uint32_t)(seed >> 32));
  float x = fmul(g.w, u01_f32(ph.v[0])), y = fmul(g.h, u01_f32(ph.v[1]));
  wr.id[t] = id;
  wr.pv[t] = make_float4(x, y, 0.f, 0.f);
  uint32_t c;
  if (flat_cell(g, x, y, &c))
    atomicAdd(&count[r * g.ncells + c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}

__global__ void batch_pack_kernel(Geom g, uint32_t rep_n, uint64_t total, SoA s, Agents d,
                                  uint32_t* __restrict__ count, int* err) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t r = (uint32_t)(t / rep_n);
  const float x = s.x[t], y = s.y[t];
  d.id[t] = s.id[t];
  d.pv[t] = make_float4(x, y, s.dx[t], s.dy[t]);
  uint32_t c;
  if (flat_cell(g, x, y, &c))
    atomicAdd(&count[r * g.ncells + c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}
__global__ void batch_unpack_kernel(Geom g, uint64_t total, Agents a, SoA d, int32_t* cell) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  float4 q = a.pv[t];
  d.id[t] = a.id[t];
  d.x[t] = q.x;
  d.y[t] = q.y;
  d.dx[t] = q.z;
  d.dy[t] = q.w;
  if (cell) {
    uint32_t c;
    flat_cell(g, q.x, q.y, &c);
    cell[t] = (int32_t)c;  // cell inside the agent's own replica
  }
}

// K3 for the batch
__global__ void __launch_bounds__(256)
batch_scatter_kernel(Geom g, uint32_t rep_n, uint64_t total, Agents src, Agents dst,
                     const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ count) {
#if !KG_SCATTER_EARLY
  grid_dep_wait();  // cell_start comes from the scan launched just before
#endif
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t r = (uint32_t)(t / rep_n);
  float4 q = src.pv[t];
  uint32_t id = src.id[t];
  uint32_t c;
  const bool in_grid = flat_cell(g, q.x, q.y, &c);
#if KG_SCATTER_EARLY
  grid_dep_wait();  // the log is older than the scan (loaded beside it); count and cell_start need the scan done
#endif
  if (!in_grid) return;
  c += r * g.ncells;
  uint32_t rank = atomicSub(&count[c], 1u) - 1u;
  uint32_t d = cell_start[c] + rank;
  dst.id[d] = id;
  dst.pv[d] = q;
}

__global__ void batch_sort_cells_kernel(uint32_t ncells, const uint32_t* __restrict__ cs, Agents a) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint32_t s = cs[c], e = cs[c + 1];
  for (uint32_t p = s + 1; p < e; ++p) {
    uint32_t id = a.id[p];
    if (a.id[p - 1] <= id) continue;
    float4 v = a.pv[p];
    uint32_t q = p;
    while (q > s && a.id[q - 1] > id) {
      a.id[q] = a.id[q - 1];
      a.pv[q] = a.pv[q - 1];
      --q;
    }
    a.id[q] = id;
    a.pv[q] = v;
  }
}

// K4 for the batch.  The packed kernels need one window size `dd` and one query kind for the whole
// batch (checked on the host; radii may still differ, each replica has its own exact-query
// threshold); otherwise the generic window walk runs with each replica's own radius and query kind.
template <int MODE>  // 0: generic walk, 1: packed relaxed query, 2: packed exact-distance query
__global__ void __launch_bounds__(128)
batch_step_kernel(Geom g, int dd, uint32_t rep_n, uint64_t total, uint64_t step,
                  const KgBoidsParams* __restrict__ params, const float* __restrict__ thresholds,
                  Agents rd,
                  const uint32_t* __restrict__ cell_start, Agents wr, uint32_t* __restrict__ count,
                  const int* __restrict__ ids_dup, int* err) {
  grid_dep_wait();  // the read buffer comes from the scatter launched just before
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t i = (uint32_t)t;
  const uint32_t r = i / rep_n;
  KgBoidsParams p = params[r];
  p.step = step;
  const uint32_t* __restrict__ cs = cell_start + (size_t)r * g.ncells;  // absolute offsets into rd
  const uint32_t id = rd.id[i];
  const ulonglong2 self = reinterpret_cast<const ulonglong2*>(rd.pv)[i];
  uint32_t c;
  bool ok;
  if (MODE != 0) {
    int ncx, ncy;
    const ulonglong2 out = boids_step_packed<MODE == 2>(g, p, dd, MODE == 2 ? thresholds[r] : 0.0f,
                                                        *ids_dup != 0, i, id, self, 0, cs, rd.id, rd.pv,
                                                        &ncx, &ncy);
    wr.id[i] = id;
    reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
    c = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
    ok = (int32_t)c >= 0 && c < g.ncells;
  } else {
    float px, py, ldx, ldy;
    unpack2(self.x, &px, &py);
    unpack2(self.y, &ldx, &ldy);
    BoidsAcc acc;
    const uint32_t* __restrict__ rid = rd.id;
    const float4* __restrict__ rpv = rd.pv;
    if (p.exact_query)
      for_each_neighbor<true>(g, cs, rpv, px, py, p.radius, [&](uint32_t k) {
        boids_pair(acc, id, px, py, rid[k], rpv[k], g.w, g.h);
      });
    else
      for_each_neighbor<false>(g, cs, rpv, px, py, p.radius, [&](uint32_t k) {
        boids_pair(acc, id, px, py, rid[k], rpv[k], g.w, g.h);
      });
    float4 out = boids_finish(acc, p, id, px, py, ldx, ldy, g.w);
    wr.id[i] = id;
    wr.pv[i] = out;
    ok = flat_cell(g, out.x, out.y, &c);
  }
  if (ok)
    atomicAdd(&count[r * g.ncells + c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}

}  // namespace kg

using namespace kg;

struct kg_batch {
  int device = 0;
  cudaStream_t stream = nullptr;
  Geom g{};               // geometry of ONE replica (g.ncells = cells per replica)
  uint32_t nrep = 0, rep_n = 0;
  uint64_t total = 0;     // nrep * rep_n agents
  uint64_t cells_total = 0;
  Agents A, B;
  uint32_t* cell_start = nullptr;
  uint32_t* count = nullptr;
  LookbackState scan;
  KgBoidsParams* d_params = nullptr;
  float* d_thresholds = nullptr;  // exact_threshold(radius) per replica
  std::vector<KgBoidsParams> h_params;
  bool params_dirty = true;
  bool populated = false;  // the read buffer holds the population
  bool logged = false;     // the write log holds a full population waiting for a rebuild
  int order = KG_ORDER_ANY;
  int* d_err = nullptr;
  int* h_err = nullptr;
  int* d_ids_dup = nullptr;  // 0: ids 0..n-1 per replica from init (index test valid); 1: compare ids
  SoA stage;
  int32_t* stage_cell = nullptr;
  bool have_stage = false;
  double* red = nullptr;  // scratch of kg_batch_reduce
  size_t red_bytes = 0;
  Stopwatch watch;
  L2Flusher flusher;
  EventPool events;
};

namespace {

constexpr int kBT = 256;
inline unsigned bblocks(uint64_t n, int t = kBT) { return (unsigned)std::max<uint64_t>(1, (n + t - 1) / t); }

int buse(kg_batch* b) {
  if (!b) return fail(KG_E_INVALID, "null batch handle");
  KG_CUDA(cudaSetDevice(b->device));
  return KG_OK;
}
#define BLAUNCH_PDL(b, kernel, grid, block, ...)                                          \
  do {                                                                                    \
    cudaError_t _le = launch_pdl(kernel, dim3(grid), dim3(block), (b)->stream, __VA_ARGS__); \
    if (_le != cudaSuccess)                                                               \
      return fail(KG_E_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_le)); \
    launch_counter().fetch_add(1, std::memory_order_relaxed);                             \
  } while (0)
#define BLAUNCH(b, kernel, grid, block, ...)                                              \
  do {                                                                                    \
    kernel<<<grid, block, 0, (b)->stream>>>(__VA_ARGS__);                                 \
    cudaError_t _le = cudaGetLastError();                                                 \
    if (_le != cudaSuccess)                                                               \
      return fail(KG_E_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_le)); \
    launch_counter().fetch_add(1, std::memory_order_relaxed);                             \
  } while (0)

int bsync_check(kg_batch* b) {
  KG_CUDA(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  KG_CUDA(cudaStreamSynchronize(b->stream));
  if (*b->h_err & DEV_ERR_OOB) {
    KG_CUDA(cudaMemsetAsync(b->d_err, 0, sizeof(int), b->stream));
    return fail(KG_E_OOB, "agent coordinate outside the bag grid (reference: index out of bounds panic)");
  }
  return KG_OK;
}

int push_params(kg_batch* b) {
  if (!b->params_dirty) return KG_OK;
  KG_CUDA(cudaMemcpyAsync(b->d_params, b->h_params.data(), sizeof(KgBoidsParams) * b->nrep,
                          cudaMemcpyHostToDevice, b->stream));
  std::vector<float> th(b->nrep);
  for (uint32_t r = 0; r < b->nrep; ++r) {
    const float rad = b->h_params[r].radius;
    th[r] = (rad > 0.0f && rad < 3.0e38f) ? exact_threshold(rad) : 0.0f;
  }
  KG_CUDA(cudaMemcpyAsync(b->d_thresholds, th.data(), sizeof(float) * b->nrep, cudaMemcpyHostToDevice,
                          b->stream));
  KG_CUDA(cudaStreamSynchronize(b->stream));  // h_params may be edited right after
  b->params_dirty = false;
  return KG_OK;
}

int ensure_stage(kg_batch* b) {
  if (b->have_stage) return KG_OK;
  uint64_t n = b->total + 64;
  KG_CUDA(cudaMalloc(&b->stage.id, n * 4));
  KG_CUDA(cudaMalloc(&b->stage.x, n * 4));
  KG_CUDA(cudaMalloc(&b->stage.y, n * 4));
  KG_CUDA(cudaMalloc(&b->stage.dx, n * 4));
  KG_CUDA(cudaMalloc(&b->stage.dy, n * 4));
  KG_CUDA(cudaMalloc(&b->stage_cell, n * 4));
  b->have_stage = true;
  return KG_OK;
}

// lazy_update of every replica's field: one scan over all R*C cells, one scatter
int batch_rebuild(kg_batch* b) {
  if (!b->logged) return fail(KG_E_INVALID, "batch: nothing to rebuild (init or upload first)");
  exclusive_scan_lookback(b->scan, b->count, b->cells_total, b->cell_start, b->stream, 0, true);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  BLAUNCH_PDL(b, batch_scatter_kernel, bblocks(b->total), kBT, b->g, b->rep_n, b->total, b->B, b->A,
              b->cell_start, b->count);
  if (b->order == KG_ORDER_CANONICAL)
    BLAUNCH(b, batch_sort_cells_kernel, bblocks(b->cells_total, 128), 128, (uint32_t)b->cells_total,
            b->cell_start, b->A);
  b->populated = true;
  b->logged = false;
  return KG_OK;
}

int batch_step(kg_batch* b, uint64_t step) {
  if (!b->populated) return fail(KG_E_INVALID, "batch: no population in the read buffer");
  if (b->logged) return fail(KG_E_INVALID, "batch: write log already full (missing lazy_update)");
  KG_TRY(push_params(b));
  // the packed kernel needs one window size for the whole batch
  int dd = 0;
  bool fast = true;
  const int exact = b->h_params[0].exact_query ? 1 : 0;
  for (uint32_t r = 0; r < b->nrep && fast; ++r) {
    int d = 0;
    fast = (b->h_params[r].exact_query ? 1 : 0) == exact &&
           k4_fast_geometry(b->g, b->h_params[r].radius, exact, &d);
    if (r == 0) dd = d;
    else if (d != dd) fast = false;
  }
  unsigned grid = bblocks(b->total, 128);
#define BATCH_STEP(MODE, DD)                                                                       \
  BLAUNCH_PDL(b, batch_step_kernel<MODE>, grid, 128, b->g, DD, b->rep_n, b->total, step,           \
              (const KgBoidsParams*)b->d_params, (const float*)b->d_thresholds, b->A,              \
              (const uint32_t*)b->cell_start, b->B, b->count, (const int*)b->d_ids_dup, b->d_err)
  if (fast && exact)
    BATCH_STEP(2, dd);
  else if (fast)
    BATCH_STEP(1, dd);
  else
    BATCH_STEP(0, 0);
#undef BATCH_STEP
  b->logged = true;
  return KG_OK;
}

}  // namespace

extern "C" {

int kg_batch_create(float w, float h, float d, int toroidal, uint32_t replicas,
                    uint32_t agents_per_replica, int device, kg_batch** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (!(w > 0.f) || !(h > 0.f) || !(d > 0.f)) return fail(KG_E_INVALID, "w, h, discretization must be > 0");
  if (replicas == 0 || agents_per_replica == 0) return fail(KG_E_INVALID, "empty batch");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_batch* b = new kg_batch();
  b->device = device;
  b->g.w = w; b->g.h = h; b->g.disc = d; b->g.toroidal = toroidal ? 1 : 0;
  b->g.max_x = (int)std::min(2147483520.0f, ceilf(w / d));  // field_2d.rs:317-318 / :487-488
  b->g.max_y = (int)std::min(2147483520.0f, ceilf(h / d));
  b->g.dw = b->g.max_x + 1;
  b->g.dh = b->g.max_y + 1;
  uint64_t nc = (uint64_t)b->g.dw * (uint64_t)b->g.dh;
  b->nrep = replicas;
  b->rep_n = agents_per_replica;
  b->total = (uint64_t)replicas * agents_per_replica;
  b->cells_total = nc * replicas;
  if (b->cells_total >= (1ull << 31) || b->total >= 0xFFFFFFF0ull) {
    delete b;
    return fail(KG_E_INVALID, "batch too large: %llu cells, %llu agents",
                (unsigned long long)(nc * replicas), (unsigned long long)replicas * agents_per_replica);
  }
  b->g.ncells = (uint32_t)nc;
  auto cleanup = [&](int code) { kg_batch_destroy(b); return code; };
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  int rc;
  if ((rc = lookback_init(b->scan, b->cells_total, b->stream)) != KG_OK) return cleanup(rc);
  if (cudaMalloc(&b->A.id, (b->total + 64) * 4) != cudaSuccess ||
      cudaMalloc(&b->A.pv, (b->total + 64) * 16) != cudaSuccess ||
      cudaMalloc(&b->B.id, (b->total + 64) * 4) != cudaSuccess ||
      cudaMalloc(&b->B.pv, (b->total + 64) * 16) != cudaSuccess ||
      cudaMalloc(&b->cell_start, (b->cells_total + 16) * 4) != cudaSuccess ||
      cudaMalloc(&b->count, (b->cells_total + 16) * 4) != cudaSuccess ||
      cudaMalloc(&b->d_params, sizeof(KgBoidsParams) * replicas) != cudaSuccess ||
      cudaMalloc(&b->d_thresholds, sizeof(float) * replicas) != cudaSuccess ||
      cudaMalloc(&b->d_err, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&b->d_ids_dup, sizeof(int)) != cudaSuccess ||
      cudaHostAlloc(&b->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess)
    return cleanup(fail(KG_E_CUDA, "batch allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaMemsetAsync(b->cell_start, 0, (b->cells_total + 16) * 4, b->stream);
  cudaMemsetAsync(b->count, 0, (b->cells_total + 16) * 4, b->stream);
  cudaMemsetAsync(b->d_err, 0, sizeof(int), b->stream);
  cudaMemsetAsync(b->d_ids_dup, 0, sizeof(int), b->stream);
  // defaults: the fixture's constants with the relaxed query, seed 42 + replica
  // (bird.rs:12-17)
  b->h_params.resize(replicas);
  for (uint32_t r = 0; r < replicas; ++r) {
    KgBoidsParams p{};
    p.cohesion = p.avoidance = p.randomness = p.consistency = p.momentum = 1.0f;
    p.jump = 0.7f;
    p.radius = 10.0f;
    p.exact_query = KG_QUERY_RELAX;
    p.seed = 42ull + r;
    b->h_params[r] = p;
  }
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return cleanup(fail(KG_E_CUDA, "batch init failed"));
  *out = b;
  return KG_OK;
}

int kg_batch_destroy(kg_batch* b) {
  if (!b) return KG_OK;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  b->watch.destroy();
  b->flusher.destroy();
  b->events.destroy();
  cudaFree(b->A.id); cudaFree(b->A.pv); cudaFree(b->B.id); cudaFree(b->B.pv);
  if (b->have_stage) {
    cudaFree(b->stage.id); cudaFree(b->stage.x); cudaFree(b->stage.y); cudaFree(b->stage.dx);
    cudaFree(b->stage.dy); cudaFree(b->stage_cell);
  }
  lookback_destroy(b->scan);
  cudaFree(b->cell_start);
  cudaFree(b->count);
  cudaFree(b->d_params);
  cudaFree(b->d_thresholds);
  cudaFree(b->d_err);
  cudaFree(b->d_ids_dup);
  if (b->h_err) cudaFreeHost(b->h_err);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
  return KG_OK;
}

int kg_batch_dims(kg_batch* b, uint32_t* replicas, uint32_t* agents_per_replica, int32_t* dw,
                  int32_t* dh) {
  if (!b) return fail(KG_E_INVALID, "null batch handle");
  if (replicas) *replicas = b->nrep;
  if (agents_per_replica) *agents_per_replica = b->rep_n;
  if (dw) *dw = b->g.dw;
  if (dh) *dh = b->g.dh;
  return KG_OK;
}

int kg_batch_set_order(kg_batch* b, int order) {
  if (!b) return fail(KG_E_INVALID, "null batch handle");
  if (order != KG_ORDER_ANY && order != KG_ORDER_CANONICAL) return fail(KG_E_INVALID, "bad order");
  b->order = order;
  return KG_OK;
}

int kg_batch_set_params(kg_batch* b, uint32_t first, uint32_t n, const KgBoidsParams* p) {
  if (!b || !p) return fail(KG_E_INVALID, "null argument");
  if ((uint64_t)first + n > b->nrep) return fail(KG_E_INVALID, "replica range out of bounds");
  for (uint32_t r = 0; r < n; ++r) b->h_params[first + r] = p[r];
  b->params_dirty = true;
  return KG_OK;
}

int kg_batch_init_flockers(kg_batch* b) {
  KG_TRY(buse(b));
  KG_TRY(push_params(b));
  if (b->logged) KG_CUDA(cudaMemsetAsync(b->count, 0, b->cells_total * 4, b->stream));
  BLAUNCH(b, batch_init_kernel, bblocks(b->total), kBT, b->g, b->rep_n, b->total, b->d_params, b->B,
          b->count, b->d_err);
  KG_CUDA(cudaMemsetAsync(b->d_ids_dup, 0, sizeof(int), b->stream));  // ids 0..n-1 per replica
  b->logged = true;
  b->populated = false;
  return bsync_check(b);
}

int kg_batch_upload(kg_batch* b, const uint32_t* id, const float* x, const float* y, const float* dx,
                    const float* dy) {
  KG_TRY(buse(b));
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null input array");
  KG_TRY(ensure_stage(b));
  cudaStream_t s = b->stream;
  const uint64_t n = b->total;
  KG_CUDA(cudaMemcpyAsync(b->stage.id, id, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(b->stage.x, x, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(b->stage.y, y, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(b->stage.dx, dx, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(b->stage.dy, dy, n * 4, cudaMemcpyHostToDevice, s));
  if (b->logged) KG_CUDA(cudaMemsetAsync(b->count, 0, b->cells_total * 4, s));
  BLAUNCH(b, batch_pack_kernel, bblocks(n), kBT, b->g, b->rep_n, n, b->stage, b->B, b->count, b->d_err);
  KG_CUDA(cudaMemsetAsync(b->d_ids_dup, 1, sizeof(int), s));  // foreign ids: compare ids (bird.rs:63)
  b->logged = true;
  b->populated = false;
  int rc = bsync_check(b);
  if (rc != KG_OK) {  // nothing of a rejected upload stays behind
    cudaMemsetAsync(b->count, 0, b->cells_total * 4, s);
    b->logged = false;
  }
  return rc;
}

int kg_batch_lazy_update(kg_batch* b) {
  KG_TRY(buse(b));
  return batch_rebuild(b);
}

int kg_batch_step_boids(kg_batch* b, uint64_t step) {
  KG_TRY(buse(b));
  return batch_step(b, step);
}

int kg_batch_run_boids(kg_batch* b, uint64_t first_step, uint64_t nsteps) {
  KG_TRY(buse(b));
  for (uint64_t i = 0; i < nsteps; ++i) {
    KG_TRY(batch_step(b, first_step + i));
    KG_TRY(batch_rebuild(b));
  }
  return KG_OK;
}

int kg_batch_run_boids_timed(kg_batch* b, uint64_t first_step, uint64_t nsteps, uint64_t flush_bytes,
                             double* ms_sum) {
  KG_TRY(buse(b));
  if (!ms_sum) return fail(KG_E_INVALID, "null argument");
  for (uint64_t i = 0; i < nsteps; ++i) {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    KG_TRY(b->events.get(2 * i, &e0));
    KG_TRY(b->events.get(2 * i + 1, &e1));
    KG_TRY(b->flusher.run(flush_bytes, b->stream));
    KG_CUDA(cudaEventRecord(e0, b->stream));
    KG_TRY(batch_step(b, first_step + i));
    KG_TRY(batch_rebuild(b));
    KG_CUDA(cudaEventRecord(e1, b->stream));
  }
  KG_TRY(bsync_check(b));
  double sum = 0;
  for (uint64_t i = 0; i < nsteps; ++i) {
    float t = 0.f;
    KG_CUDA(cudaEventElapsedTime(&t, b->events.ev[2 * i], b->events.ev[2 * i + 1]));
    sum += t;
  }
  *ms_sum = sum;
  return KG_OK;
}

int kg_batch_download(kg_batch* b, uint32_t* id, float* x, float* y, float* dx, float* dy,
                      int32_t* cell) {
  KG_TRY(buse(b));
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null output array");
  if (!b->populated) return fail(KG_E_INVALID, "batch: no population in the read buffer");
  KG_TRY(ensure_stage(b));
  const uint64_t n = b->total;
  cudaStream_t s = b->stream;
  BLAUNCH(b, batch_unpack_kernel, bblocks(n), kBT, b->g, n, b->A, b->stage, cell ? b->stage_cell : nullptr);
  KG_CUDA(cudaMemcpyAsync(id, b->stage.id, n * 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(x, b->stage.x, n * 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(y, b->stage.y, n * 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(dx, b->stage.dx, n * 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(dy, b->stage.dy, n * 4, cudaMemcpyDeviceToHost, s));
  if (cell) KG_CUDA(cudaMemcpyAsync(cell, b->stage_cell, n * 4, cudaMemcpyDeviceToHost, s));
  return bsync_check(b);
}

int kg_batch_reduce(kg_batch* b, double* out) {
  KG_TRY(buse(b));
  if (!out) return fail(KG_E_INVALID, "null out");
  if (!b->populated) return fail(KG_E_INVALID, "batch has no population in its read buffer (init / upload + lazy_update first)");
  // replica r owns entries [r * n, (r + 1) * n) of the read buffer at every step
  return reduce_segments(b->A.pv, b->nrep, b->rep_n, &b->red, &b->red_bytes, out, b->stream);
}

int kg_batch_sync(kg_batch* b) {
  KG_TRY(buse(b));
  return bsync_check(b);
}
int kg_batch_timer_start(kg_batch* b) {
  KG_TRY(buse(b));
  return b->watch.start(b->stream);
}
int kg_batch_timer_stop(kg_batch* b, double* ms) {
  KG_TRY(buse(b));
  return b->watch.stop(b->stream, ms);
}

}  // extern "C"
