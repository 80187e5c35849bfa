// 2-D block decomposition of Field2D over the GPUs of one box (precedent: the kd-tree
// blocks of src/engine/fields/kdtree_mpi.rs:211-238, which cut the field along both axes when strips of
// whole columns get too thin).  The world's cell grid is cut into nbx x nby rectangles; block (bx, by)
// owns cell columns [bx*max_x/nbx, (bx+1)*max_x/nbx) x rows [by*max_y/nby, (by+1)*max_y/nby) (the last
// block of an axis also the padding column / row) and keeps a ring of `dd` halo cells around them.
//
// Per block the layout is a small Field2D over its LOCAL cells (owned + halo): read buffer A sorted by
// local cell, write log B, count / cell_start, the same look-back scan and atomic scatter.  Positions
// stay in world coordinates, so K4 is the packed kernel's own per-agent step (boids_step_packed) with the
// local column offset and a cell_start pointer shifted by the local row offset: the candidate sequence of
// an owned agent — and with KG_ORDER_CANONICAL every f32 bit of its step — is what one GPU computes.
//
// One exchange per step, decided by the SENDER: the block that steps an agent knows its new cell, hence
// its new owner and every block whose halo ring contains that cell (all among the sender's 8 neighbours;
// halos do not wrap — toroidal fields clamp the window, SURVEY F3 — migrants do).  Each recipient gets a
// copy: into the sender's own log if it is the recipient, else into the outbox of that direction.  Ghosts
// are never stepped and never forwarded, so a block's halo is rebuilt from scratch every step.
// The step kernel stores the copies straight into the recipients' inboxes (peer memory over NVLink, double
// buffered by step parity); only the nine counts per block travel through the host, which hands them to the
// receivers' append kernels as arguments (one process drives every GPU, one stream synchronisation per block
// and step).  It is the decomposition for worlds where strips are too thin, not the fast path — strips
// (strip.cu) keep everything, flags included, on the device.
#include <algorithm>
#include <vector>

#include "boids_device.cuh"
#include "scan.cuh"

namespace kg {

struct BlockGeom {
  Geom g;  // world geometry, except g.dh = LOCAL row count (what the gather functions index cells with)
  int dd;
  int bx, by, nbx, nby;
  int own_x0, own_x1, own_y0, own_y1;  // owned cells
  int lx0, lx1, ly0, ly1;              // local cells (owned + halo ring, clipped to the world)
  int gdw, gdh;                        // the world's dw, dh
};

// first column (row) of part b of `parts` over `maxc` scanned columns; part `parts` = end incl. padding
__host__ __device__ __forceinline__ int part_begin(int b, int parts, int maxc) {
  return b >= parts ? maxc + 1 : (int)(((long long)b * maxc) / parts);
}
__host__ __device__ __forceinline__ int part_of(int c, int parts, int maxc) {
  if (c >= maxc) return parts - 1;  // the padding column / row
  int b = (int)(((long long)c * parts) / maxc);
  b = b < 0 ? 0 : (b >= parts ? parts - 1 : b);
  while (b + 1 < parts && part_begin(b + 1, parts, maxc) <= c) ++b;
  while (b > 0 && part_begin(b, parts, maxc) > c) --b;
  return b;
}
// local window of part b along one axis: [lo, hi)
__host__ __device__ __forceinline__ void part_window(int b, int parts, int maxc, int dd, int gd, int* lo, int* hi) {
  const int a = part_begin(b, parts, maxc) - dd, e = part_begin(b + 1, parts, maxc) + dd;
  *lo = a < 0 ? 0 : a;
  *hi = e > gd ? gd : e;
}

struct BlockOut {  // where the copies for the 8 directions go (slot 4 = myself, unused), xcap entries each:
  uint32_t* id[9];  // this step's segment of the RECEIVER's inbox — peer memory over NVLink when the receiver
  float4* pv[9];    // sits on another GPU (plain stores; the host's stream synchronisation orders them)
  uint32_t* count;  // [0..8]: entries sent per direction; [4] = entries appended to my own log behind the agents' own
                    // slots; [9] = holes among those slots (ghosts, which do not step, and agents that left)
  uint32_t xcap;
};

__device__ __forceinline__ bool block_local_cell(const BlockGeom& bg, int cx, int cy, uint32_t* cell) {
  if (cx < bg.lx0 || cx >= bg.lx1 || cy < bg.ly0 || cy >= bg.ly1) return false;
  *cell = (uint32_t)(cx - bg.lx0) * (uint32_t)bg.g.dh + (uint32_t)(cy - bg.ly0);
  return true;
}
// appends behind `base` (0 for uploads; the stepped agents' own slots [0, n_read) during a step)
__device__ __forceinline__ void block_log_append(const BlockGeom& bg, uint32_t id, float4 v, int cx, int cy, Agents log,
                                                 uint32_t cap, uint32_t* log_len, uint32_t* count, int* err,
                                                 uint32_t base = 0) {
  uint32_t c;
  if (!block_local_cell(bg, cx, cy, &c)) {
    atomicOr(err, 4);  // an entry outside this block's window: a protocol error, never silent
    return;
  }
  const uint32_t slot = base + atomicAdd(log_len, 1u);
  if (slot >= cap) {
    atomicOr(err, 8);
    return;
  }
  log.id[slot] = id;
  log.pv[slot] = v;
  atomicAdd(&count[c], 1u);
}

// uploads: every block sees the same agents and keeps the ones inside its window (owned or ghost)
__global__ void block_upload_kernel(BlockGeom bg, uint64_t n, const uint32_t* __restrict__ id, const float* __restrict__ x,
                                    const float* __restrict__ y, const float* __restrict__ dx, const float* __restrict__ dy,
                                    Agents log, uint32_t cap, uint32_t* log_len, uint32_t* count, int* err) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = f2i_sat(floorf(fdiv(x[i], bg.g.disc))), cy = f2i_sat(floorf(fdiv(y[i], bg.g.disc)));
  if (cx < 0 || cx >= bg.gdw || cy < 0 || cy >= bg.gdh) {
    atomicOr(err, DEV_ERR_OOB);
    return;
  }
  uint32_t c;
  if (!block_local_cell(bg, cx, cy, &c)) return;
  block_log_append(bg, id[i], make_float4(x[i], y[i], dx[i], dy[i]), cx, cy, log, cap, log_len, count, err);
}

// K4 of a block: owned agents take their step; the new state goes to every block that must see it
template <bool EXACT>
__global__ void __launch_bounds__(128)
block_step_kernel(BlockGeom bg, KgBoidsParams p, float T, uint32_t n, Agents rd, const uint32_t* __restrict__ cell_start,
                  Agents log, uint32_t cap, uint32_t* __restrict__ count, BlockOut out, int* err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ulonglong2 self = reinterpret_cast<const ulonglong2*>(rd.pv)[i];
  const Recip rdisc = recip_of(bg.g.disc);
  int cx, cy;
  cell_of2(self.x, rdisc, &cx, &cy);
  // The write log starts with one slot per entry of the read buffer: an agent that stays mine writes its own
  // slot (no counter: one shared append counter for 32 M stayers cost 5 ms per step), everything else leaves a
  // hole there (kIdNone, skipped by the scatter) and the few extra copies are appended behind those slots.
  if (cx < bg.own_x0 || cx >= bg.own_x1 || cy < bg.own_y0 || cy >= bg.own_y1) {  // a ghost: not stepped
    log.id[i] = kIdNone;
    atomicAdd(&out.count[9], 1u);
    return;
  }
  const uint32_t id = rd.id[i];
  bool stays = false;
  int ncx, ncy;
  // cells are indexed (ci - lx0) * lh + (cj - ly0): the row offset is folded into the pointer
  const ulonglong2 o = boids_step_packed<EXACT>(bg.g, p, bg.dd, T, false, i, id, self, bg.lx0, cell_start - bg.ly0,
                                                rd.id, rd.pv, &ncx, &ncy);
  if (ncx < 0 || ncx >= bg.gdw || ncy < 0 || ncy >= bg.gdh) {
    atomicOr(err, DEV_ERR_OOB);
    log.id[i] = kIdNone;
    atomicAdd(&out.count[9], 1u);
    return;
  }
  float4 v;
  {
    float a, b, c, d;
    unpack2(o.x, &a, &b);
    unpack2(o.y, &c, &d);
    v = make_float4(a, b, c, d);
  }
  const int ox = part_of(ncx, bg.nbx, bg.g.max_x), oy = part_of(ncy, bg.nby, bg.g.max_y);
  for (int rx = ox - 1; rx <= ox + 1; ++rx) {
    if (rx < 0 || rx >= bg.nbx) continue;  // halos do not wrap
    int lo, hi;
    part_window(rx, bg.nbx, bg.g.max_x, bg.dd, bg.gdw, &lo, &hi);
    if (ncx < lo || ncx >= hi) continue;
    for (int ry = oy - 1; ry <= oy + 1; ++ry) {
      if (ry < 0 || ry >= bg.nby) continue;
      part_window(ry, bg.nby, bg.g.max_y, bg.dd, bg.gdh, &lo, &hi);
      if (ncy < lo || ncy >= hi) continue;
      int ddx = rx - bg.bx, ddy = ry - bg.by;  // migrants wrap: the recipient is still one of my 8 neighbours
      if (ddx > 1) ddx -= bg.nbx;
      if (ddx < -1) ddx += bg.nbx;
      if (ddy > 1) ddy -= bg.nby;
      if (ddy < -1) ddy += bg.nby;
      if (ddx < -1 || ddx > 1 || ddy < -1 || ddy > 1) {
        atomicOr(err, 4);
        continue;
      }
      if (ddx == 0 && ddy == 0 && rx == ox && ry == oy) {  // it stays mine: its own slot
        uint32_t c;
        if (block_local_cell(bg, ncx, ncy, &c)) {
          log.id[i] = id;
          log.pv[i] = v;
          atomicAdd(&count[c], 1u);
          stays = true;
        } else {
          atomicOr(err, 4);
        }
      } else if (ddx == 0 && ddy == 0) {  // it left, but its new cell is in my halo ring: I keep a ghost of it
        block_log_append(bg, id, v, ncx, ncy, log, cap, &out.count[4], count, err, n);
      } else {
        const int d = (ddx + 1) * 3 + (ddy + 1);
        const uint32_t slot = atomicAdd(&out.count[d], 1u);
        if (slot >= out.xcap) {
          atomicOr(err, 8);
        } else {
          out.id[d][slot] = id;
          out.pv[d][slot] = v;
        }
      }
    }
  }
  if (!stays) {
    log.id[i] = kIdNone;
    atomicAdd(&out.count[9], 1u);
  }
}

struct BlockIn {
  uint32_t c[9];
};
// arrivals (migrants and ghosts alike) join the log; blockIdx.y = direction
__global__ void block_append_kernel(BlockGeom bg, BlockIn in, const uint32_t* __restrict__ in_id, const float4* __restrict__ in_pv,
                                    uint32_t xcap, Agents log, uint32_t cap, uint32_t* log_len, uint32_t* count, int* err,
                                    uint32_t base) {
  const int d = blockIdx.y;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= in.c[d]) return;
  const float4 v = in_pv[(size_t)d * xcap + k];
  const int cx = f2i_sat(floorf(fdiv(v.x, bg.g.disc))), cy = f2i_sat(floorf(fdiv(v.y, bg.g.disc)));
  block_log_append(bg, in_id[(size_t)d * xcap + k], v, cx, cy, log, cap, log_len, count, err, base);
}

__global__ void __launch_bounds__(256)
block_scatter_kernel(BlockGeom bg, uint32_t n, Agents src, Agents dst, const uint32_t* __restrict__ cell_start,
                     uint32_t* __restrict__ count) {
  grid_dep_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = src.id[i];
  if (id == kIdNone) return;  // the slot of a ghost or of an agent that left
  const float4 v = src.pv[i];
  const int cx = f2i_sat(floorf(fdiv(v.x, bg.g.disc))), cy = f2i_sat(floorf(fdiv(v.y, bg.g.disc)));
  uint32_t c;
  if (!block_local_cell(bg, cx, cy, &c)) return;  // cannot happen: the log only takes window entries
  const uint32_t rank = atomicSub(&count[c], 1u) - 1u;
  const uint32_t d = cell_start[c] + rank;
  dst.id[d] = id;
  dst.pv[d] = v;
}
// KG_ORDER_CANONICAL: every bag in ascending id (what makes the f32 sums reproducible)
__global__ void block_sort_cells_kernel(uint32_t ncells, const uint32_t* __restrict__ cs, Agents a) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t s = cs[c], e = cs[c + 1];
  for (uint32_t p = s + 1; p < e; ++p) {
    const uint32_t id = a.id[p];
    if (a.id[p - 1] <= id) continue;
    const float4 v = a.pv[p];
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
__global__ void block_owned_kernel(BlockGeom bg, uint32_t n, const float4* __restrict__ pv, uint8_t* __restrict__ owned) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int cx = f2i_sat(floorf(fdiv(pv[i].x, bg.g.disc))), cy = f2i_sat(floorf(fdiv(pv[i].y, bg.g.disc)));
  owned[i] = cx >= bg.own_x0 && cx < bg.own_x1 && cy >= bg.own_y0 && cy < bg.own_y1;
}

}  // namespace kg

using namespace kg;

struct kg_block {
  int device = 0;
  cudaStream_t stream = nullptr;
  BlockGeom bg;
  uint32_t ncells = 0;
  uint64_t capacity = 0;
  uint32_t xcap = 0;
  int order = KG_ORDER_ANY;
  Agents A, B;
  uint32_t* count = nullptr;
  uint32_t* cell_start = nullptr;
  LookbackState scan;
  BlockOut out{};
  uint32_t* in_id = nullptr;
  float4* in_pv = nullptr;
  uint32_t* h_counts = nullptr;  // pinned [9]
  int* d_err = nullptr;
  int* h_err = nullptr;
  SoA stage;
  uint64_t stage_cap = 0;
  uint8_t* owned = nullptr;
  uint64_t epoch = 0;     // steps taken; its parity selects the inbox half the neighbours write into
  bool peers_ok = false;  // peer access towards the neighbours' devices enabled
  uint32_t n_read = 0;
  uint32_t n_log_host = 0;  // entries appended by uploads since the last rebuild are counted on the device
};

namespace {

constexpr int kBT = 256;
inline unsigned bblk(uint64_t n, int t = kBT) { return (unsigned)std::max<uint64_t>(1, (n + t - 1) / t); }
int buse(kg_block* b) {
  if (!b) return fail(KG_E_INVALID, "null block handle");
  KG_CUDA(cudaSetDevice(b->device));
  return KG_OK;
}
int block_check(kg_block* b) {
  KG_CUDA(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  KG_CUDA(cudaStreamSynchronize(b->stream));
  const int e = *b->h_err;
  if (!e) return KG_OK;
  KG_CUDA(cudaMemsetAsync(b->d_err, 0, sizeof(int), b->stream));
  if (e & DEV_ERR_OOB) return fail(KG_E_OOB, "agent coordinate outside the bag grid (reference: index out of bounds panic)");
  if (e & 8) return fail(KG_E_CAPACITY, "block (%d, %d): more agents or exchange entries than its capacity", b->bg.bx, b->bg.by);
  return fail(KG_E_INVALID, "block (%d, %d): an agent reached a block that is not a neighbour", b->bg.bx, b->bg.by);
}
int block_err_code(kg_block* b, int e) {
  if (!e) return KG_OK;
  cudaMemsetAsync(b->d_err, 0, sizeof(int), b->stream);
  if (e & DEV_ERR_OOB) return fail(KG_E_OOB, "agent coordinate outside the bag grid (reference: index out of bounds panic)");
  if (e & 8) return fail(KG_E_CAPACITY, "block (%d, %d): more agents or exchange entries than its capacity", b->bg.bx, b->bg.by);
  return fail(KG_E_INVALID, "block (%d, %d): an agent reached a block that is not a neighbour", b->bg.bx, b->bg.by);
}
// log (n entries, histogrammed) -> sorted read buffer
int block_rebuild(kg_block* b, uint32_t n, uint32_t n_live) {
  if (n > b->capacity) return fail(KG_E_CAPACITY, "block (%d, %d) holds %u agents, capacity %llu", b->bg.bx, b->bg.by, n,
                                   (unsigned long long)b->capacity);
  exclusive_scan_lookback(b->scan, b->count, b->ncells, b->cell_start, b->stream, 0, false);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  if (n) {
    cudaError_t e = launch_pdl(block_scatter_kernel, dim3(bblk(n)), dim3(kBT), b->stream, b->bg, n, b->B, b->A,
                               (const uint32_t*)b->cell_start, b->count);
    if (e != cudaSuccess) return fail(KG_E_CUDA, "block scatter launch failed: %s", cudaGetErrorString(e));
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    if (b->order == KG_ORDER_CANONICAL) {
      block_sort_cells_kernel<<<bblk(b->ncells, 128), 128, 0, b->stream>>>(b->ncells, b->cell_start, b->A);
      launch_counter().fetch_add(1, std::memory_order_relaxed);
    }
  }
  b->n_read = n_live;
  return KG_OK;
}
kg_block* neighbour(kg_block** blocks, const kg_block* b, int d) {
  const int ddx = d / 3 - 1, ddy = d % 3 - 1;
  const int nx = (b->bg.bx + ddx + b->bg.nbx) % b->bg.nbx, ny = (b->bg.by + ddy + b->bg.nby) % b->bg.nby;
  return blocks[nx * b->bg.nby + ny];
}

}  // namespace

extern "C" {

int kg_block_create(float w, float h, float disc, int toroidal, float radius, int bx, int by, int nbx, int nby,
                    uint64_t capacity, uint64_t xcap, int device, kg_block** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (!(w > 0.f && h > 0.f && disc > 0.f)) return fail(KG_E_INVALID, "bad field geometry");
  if (nbx < 1 || nby < 1 || bx < 0 || bx >= nbx || by < 0 || by >= nby) return fail(KG_E_INVALID, "bad block coordinates");
  if (capacity == 0 || capacity >= (1ull << 31) || xcap == 0 || xcap >= (1ull << 28)) return fail(KG_E_INVALID, "bad capacity");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  Geom g;
  g.w = w; g.h = h; g.disc = disc; g.toroidal = toroidal;
  g.max_x = (int)std::min(2147483520.0f, ceilf(w / disc));  // field_2d.rs:487-488, as kg_field2d_create
  g.max_y = (int)std::min(2147483520.0f, ceilf(h / disc));
  g.dw = g.max_x + 1;  // :317-318
  g.dh = g.max_y + 1;
  if ((uint64_t)g.dw * (uint64_t)g.dh >= (1ull << 31)) return fail(KG_E_INVALID, "too many cells");
  g.ncells = (uint32_t)g.dw * (uint32_t)g.dh;
  int dd = 0;
  if (!k4_fast_geometry(g, radius, 0, &dd))
    return fail(KG_E_INVALID, "blocks need a toroidal field and a query window small against the world (the packed K4's geometry class)");
  if (g.max_x / nbx < dd + 1 || g.max_y / nby < dd + 1)
    return fail(KG_E_INVALID, "blocks of %d x %d cells are thinner than the window (%d + 1 cells)", g.max_x / nbx, g.max_y / nby, dd);
  KG_CUDA(cudaSetDevice(device));
  kg_block* b = new kg_block();
  b->device = device;
  BlockGeom& bg = b->bg;
  bg.dd = dd; bg.bx = bx; bg.by = by; bg.nbx = nbx; bg.nby = nby;
  bg.gdw = g.dw; bg.gdh = g.dh;
  bg.own_x0 = part_begin(bx, nbx, g.max_x); bg.own_x1 = part_begin(bx + 1, nbx, g.max_x);
  bg.own_y0 = part_begin(by, nby, g.max_y); bg.own_y1 = part_begin(by + 1, nby, g.max_y);
  part_window(bx, nbx, g.max_x, dd, g.dw, &bg.lx0, &bg.lx1);
  part_window(by, nby, g.max_y, dd, g.dh, &bg.ly0, &bg.ly1);
  bg.g = g;
  bg.g.dh = bg.ly1 - bg.ly0;  // the gather functions index cells with g.dh: local rows
  b->ncells = (uint32_t)(bg.lx1 - bg.lx0) * (uint32_t)bg.g.dh;
  bg.g.ncells = b->ncells;
  b->capacity = capacity;
  b->xcap = (uint32_t)xcap;
  auto bail = [&](int code) { kg_block_destroy(b); return code; };
  if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  const size_t cap = capacity + 64, cells = (size_t)b->ncells + 16;
  bool ok = cudaMalloc(&b->A.id, cap * 4) == cudaSuccess && cudaMalloc(&b->A.pv, cap * 16) == cudaSuccess &&
            cudaMalloc(&b->B.id, cap * 4) == cudaSuccess && cudaMalloc(&b->B.pv, cap * 16) == cudaSuccess &&
            cudaMalloc(&b->count, cells * 4) == cudaSuccess && cudaMalloc(&b->cell_start, cells * 4) == cudaSuccess &&
            cudaMalloc(&b->out.count, 16 * 4) == cudaSuccess && cudaMalloc(&b->in_id, 2 * 9 * xcap * 4) == cudaSuccess &&
            cudaMalloc(&b->in_pv, 2 * 9 * xcap * 16) == cudaSuccess && cudaMalloc(&b->d_err, 4) == cudaSuccess &&
            cudaMalloc(&b->owned, cap) == cudaSuccess && cudaMallocHost(&b->h_counts, 16 * 4) == cudaSuccess &&
            cudaMallocHost(&b->h_err, 4) == cudaSuccess;
  if (!ok) return bail(fail(KG_E_CUDA, "block allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  b->out.xcap = b->xcap;
  if (lookback_init(b->scan, b->ncells, b->stream) != KG_OK) return bail(KG_E_CUDA);
  cudaMemsetAsync(b->count, 0, cells * 4, b->stream);
  cudaMemsetAsync(b->cell_start, 0, cells * 4, b->stream);
  cudaMemsetAsync(b->out.count, 0, 16 * 4, b->stream);
  cudaMemsetAsync(b->d_err, 0, 4, b->stream);
  if (cudaStreamSynchronize(b->stream) != cudaSuccess) return bail(fail(KG_E_CUDA, "block init failed"));
  *out = b;
  return KG_OK;
}

int kg_block_destroy(kg_block* b) {
  if (!b) return KG_OK;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  cudaFree(b->A.id); cudaFree(b->A.pv); cudaFree(b->B.id); cudaFree(b->B.pv);
  cudaFree(b->count); cudaFree(b->cell_start); cudaFree(b->out.count);
  cudaFree(b->in_id); cudaFree(b->in_pv); cudaFree(b->d_err); cudaFree(b->owned);
  cudaFree(b->stage.id); cudaFree(b->stage.x); cudaFree(b->stage.y); cudaFree(b->stage.dx); cudaFree(b->stage.dy);
  if (b->h_counts) cudaFreeHost(b->h_counts);
  if (b->h_err) cudaFreeHost(b->h_err);
  lookback_destroy(b->scan);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
  return KG_OK;
}

int kg_block_cells(kg_block* b, int32_t* own /*[4]: x0, x1, y0, y1*/, int32_t* local /*[4]*/) {
  if (!b) return fail(KG_E_INVALID, "null block handle");
  if (own) { own[0] = b->bg.own_x0; own[1] = b->bg.own_x1; own[2] = b->bg.own_y0; own[3] = b->bg.own_y1; }
  if (local) { local[0] = b->bg.lx0; local[1] = b->bg.lx1; local[2] = b->bg.ly0; local[3] = b->bg.ly1; }
  return KG_OK;
}
int kg_block_set_order(kg_block* b, int order) {
  if (!b) return fail(KG_E_INVALID, "null block handle");
  if (order != KG_ORDER_ANY && order != KG_ORDER_CANONICAL) return fail(KG_E_INVALID, "bad order");
  b->order = order;
  return KG_OK;
}

/* n x set_object_location: the block keeps the agents inside its window (owned cells and halo ring) */
int kg_block_upload(kg_block* b, uint64_t n, const uint32_t* id, const float* x, const float* y, const float* dx,
                    const float* dy) {
  KG_TRY(buse(b));
  if (n == 0) return KG_OK;
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null argument");
  const uint64_t chunk = 1ull << 20;
  if (b->stage_cap == 0) {
    KG_CUDA(cudaMalloc(&b->stage.id, chunk * 4));
    KG_CUDA(cudaMalloc(&b->stage.x, chunk * 4));
    KG_CUDA(cudaMalloc(&b->stage.y, chunk * 4));
    KG_CUDA(cudaMalloc(&b->stage.dx, chunk * 4));
    KG_CUDA(cudaMalloc(&b->stage.dy, chunk * 4));
    b->stage_cap = chunk;
  }
  cudaStream_t s = b->stream;
  for (uint64_t o = 0; o < n; o += chunk) {
    const uint64_t m = std::min(chunk, n - o);
    KG_CUDA(cudaMemcpyAsync(b->stage.id, id + o, m * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(b->stage.x, x + o, m * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(b->stage.y, y + o, m * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(b->stage.dx, dx + o, m * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(b->stage.dy, dy + o, m * 4, cudaMemcpyHostToDevice, s));
    block_upload_kernel<<<bblk(m), kBT, 0, s>>>(b->bg, m, b->stage.id, b->stage.x, b->stage.y, b->stage.dx, b->stage.dy, b->B,
                                               (uint32_t)b->capacity, &b->out.count[4], b->count, b->d_err);
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    KG_CUDA(cudaStreamSynchronize(s));  // the host arrays are the caller's
  }
  return block_check(b);
}

/* Field::lazy_update after uploads: the log becomes the sorted read buffer */
int kg_block_lazy_update(kg_block* b) {
  KG_TRY(buse(b));
  KG_CUDA(cudaMemcpyAsync(b->h_counts, b->out.count, 9 * 4, cudaMemcpyDeviceToHost, b->stream));
  KG_CUDA(cudaStreamSynchronize(b->stream));
  const uint32_t n = b->h_counts[4];
  KG_TRY(block_rebuild(b, n, n));
  KG_CUDA(cudaMemsetAsync(b->out.count, 0, 10 * 4, b->stream));
  return block_check(b);
}

/* One world step of all blocks (blocks[bx * nby + by], every block of the decomposition, driven by this
 * process): K4 everywhere, one exchange, rebuild everywhere. */
int kg_blocks_step(kg_block** blocks, int nblocks, const KgBoidsParams* p) {
  if (!blocks || !p || nblocks < 1) return fail(KG_E_INVALID, "null argument");
  for (int k = 0; k < nblocks; ++k)
    if (!blocks[k]) return fail(KG_E_INVALID, "null block handle");
  if (blocks[0]->bg.nbx * blocks[0]->bg.nby != nblocks) return fail(KG_E_INVALID, "the world has %d blocks", blocks[0]->bg.nbx * blocks[0]->bg.nby);
  if (!(p->radius > 0.f) || (int)floorf(p->radius / blocks[0]->bg.g.disc) != blocks[0]->bg.dd)
    return fail(KG_E_INVALID, "the query radius does not match the halo the blocks were built for");
  const float T = p->exact_query ? exact_threshold(p->radius) : 0.0f;
  // 0. once: the step kernels store into the neighbours' inboxes, so their devices must be mapped
  for (int k = 0; k < nblocks; ++k) {
    kg_block* b = blocks[k];
    if (b->peers_ok) continue;
    KG_TRY(buse(b));
    for (int d = 0; d < 9; ++d) {
      kg_block* t = neighbour(blocks, b, d);
      if (d == 4 || t->device == b->device) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(t->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
      } else if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(KG_E_CUDA, "blocks need peer access from device %d to device %d: %s", b->device, t->device,
                    cudaGetErrorString(e));
      }
    }
    b->peers_ok = true;
  }
  const uint64_t epoch = blocks[0]->epoch;
  const size_t half = (size_t)(epoch & 1) * 9;  // the inbox half of this step
  // 1. every block steps its owned agents; copies for other blocks land in their inboxes directly
  for (int k = 0; k < nblocks; ++k) {
    kg_block* b = blocks[k];
    KG_TRY(buse(b));
    if (b->epoch != epoch) return fail(KG_E_INVALID, "the blocks of a world must be stepped together");
    KG_CUDA(cudaMemsetAsync(b->out.count, 0, 10 * 4, b->stream));
    if (b->n_read) {
      BlockOut out = b->out;
      for (int d = 0; d < 9; ++d) {
        kg_block* t = neighbour(blocks, b, d);
        if (t->xcap != b->xcap) return fail(KG_E_INVALID, "the blocks of a world share one exchange capacity");
        out.id[d] = t->in_id + (half + (size_t)(8 - d)) * t->xcap;  // slot = the direction seen from the receiver
        out.pv[d] = t->in_pv + (half + (size_t)(8 - d)) * t->xcap;
      }
      if (p->exact_query)
        block_step_kernel<true><<<bblk(b->n_read, 128), 128, 0, b->stream>>>(b->bg, *p, T, b->n_read, b->A, b->cell_start, b->B,
                                                                            (uint32_t)b->capacity, b->count, out, b->d_err);
      else
        block_step_kernel<false><<<bblk(b->n_read, 128), 128, 0, b->stream>>>(b->bg, *p, T, b->n_read, b->A, b->cell_start, b->B,
                                                                             (uint32_t)b->capacity, b->count, out, b->d_err);
      launch_counter().fetch_add(1, std::memory_order_relaxed);
    }
    KG_CUDA(cudaMemcpyAsync(b->h_counts, b->out.count, 10 * 4, cudaMemcpyDeviceToHost, b->stream));
    KG_CUDA(cudaMemcpyAsync(b->h_err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  }
  // 2. ONE host synchronisation per block and step: its stores into the neighbours' inboxes are complete and its
  // counts are on the host.  The inbox half written in step s + 2 is the one the append of step s read; that
  // append finished before the synchronisation of step s + 1 returned, i.e. before step s + 2 was launched.
  // (A first version wrote into single-buffered inboxes after synchronising only the sender: invisible with all
  // blocks on one device, 58 differing words over two real GPUs in the bench's parity leg.)
  std::vector<BlockIn> in(nblocks);
  for (auto& q : in) std::fill(q.c, q.c + 9, 0u);
  for (int k = 0; k < nblocks; ++k) {
    kg_block* b = blocks[k];
    KG_TRY(buse(b));
    KG_CUDA(cudaStreamSynchronize(b->stream));
    KG_TRY(block_err_code(b, *b->h_err));
    for (int d = 0; d < 9; ++d) {
      const uint32_t c = d == 4 ? 0u : b->h_counts[d];
      if (!c) continue;
      if (c > b->xcap) return fail(KG_E_CAPACITY, "block (%d, %d) sends %u entries to one neighbour, capacity %u", b->bg.bx, b->bg.by, c, b->xcap);
      kg_block* t = neighbour(blocks, b, d);
      in[t->bg.bx * t->bg.nby + t->bg.by].c[8 - d] += c;
    }
  }
  // 3. arrivals join the log; the log becomes the next read buffer
  for (int k = 0; k < nblocks; ++k) {
    kg_block* b = blocks[k];
    KG_TRY(buse(b));
    // log = [one slot per read-buffer entry (holes included) | my own extra ghosts | arrivals]
    uint32_t n = b->n_read + b->h_counts[4], most = 0;
    for (int d = 0; d < 9; ++d) {
      n += in[k].c[d];
      most = std::max(most, in[k].c[d]);
    }
    const uint32_t holes = b->n_read ? b->h_counts[9] : 0u;
    if (n > b->capacity) return fail(KG_E_CAPACITY, "block (%d, %d) would hold %u agents, capacity %llu", b->bg.bx, b->bg.by, n,
                                     (unsigned long long)b->capacity);
    if (most) {
      block_append_kernel<<<dim3(bblk(most), 9), kBT, 0, b->stream>>>(b->bg, in[k], b->in_id + half * b->xcap, b->in_pv + half * b->xcap, b->xcap, b->B,
                                                                     (uint32_t)b->capacity, &b->out.count[4], b->count, b->d_err,
                                                                     b->n_read);
      launch_counter().fetch_add(1, std::memory_order_relaxed);
    }
    KG_TRY(block_rebuild(b, n, n - holes));
    b->epoch += 1;
  }
  return KG_OK;
}

int kg_blocks_run(kg_block** blocks, int nblocks, const KgBoidsParams* p, uint64_t nsteps) {
  if (!p) return fail(KG_E_INVALID, "null params");
  KgBoidsParams q = *p;
  for (uint64_t i = 0; i < nsteps; ++i) {
    q.step = p->step + i;
    KG_TRY(kg_blocks_step(blocks, nblocks, &q));
  }
  for (int k = 0; k < nblocks; ++k) {
    KG_TRY(buse(blocks[k]));
    KG_TRY(block_check(blocks[k]));
  }
  return KG_OK;
}

/* the agents this block owns (ghosts left out), in read-buffer order */
int kg_block_download(kg_block* b, uint64_t cap, uint32_t* id, float* x, float* y, float* dx, float* dy, uint64_t* n_out) {
  KG_TRY(buse(b));
  if (!n_out) return fail(KG_E_INVALID, "null argument");
  *n_out = 0;
  const uint32_t n = b->n_read;
  if (n == 0) return block_check(b);
  block_owned_kernel<<<bblk(n), kBT, 0, b->stream>>>(b->bg, n, b->A.pv, b->owned);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  std::vector<uint8_t> own(n);
  std::vector<uint32_t> ids(n);
  std::vector<float4> pv(n);
  KG_CUDA(cudaMemcpyAsync(own.data(), b->owned, n, cudaMemcpyDeviceToHost, b->stream));
  KG_CUDA(cudaMemcpyAsync(ids.data(), b->A.id, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
  KG_CUDA(cudaMemcpyAsync(pv.data(), b->A.pv, (size_t)n * 16, cudaMemcpyDeviceToHost, b->stream));
  KG_TRY(block_check(b));
  uint64_t m = 0;
  for (uint32_t i = 0; i < n; ++i) m += own[i];
  *n_out = m;
  if (m > cap) return fail(KG_E_CAPACITY, "download needs room for %llu agents", (unsigned long long)m);
  uint64_t k = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (!own[i]) continue;
    if (id) id[k] = ids[i];
    if (x) x[k] = pv[i].x;
    if (y) y[k] = pv[i].y;
    if (dx) dx[k] = pv[i].z;
    if (dy) dy[k] = pv[i].w;
    ++k;
  }
  return KG_OK;
}

/* the partition rule by itself (no device needed): part `b` of `parts` over `maxc` scanned columns (or rows) owns
 * [out[0], out[1]) — the last part also the padding column — and keeps [out[2], out[3]) with its halo ring;
 * *owner = the part owning column c */
int kg_block_partition(int b, int parts, int maxc, int dd, int c, int32_t* out /*[4]*/, int32_t* owner) {
  if (parts < 1 || maxc < 1 || b < 0 || b >= parts || !out || !owner) return fail(KG_E_INVALID, "bad partition query");
  out[0] = part_begin(b, parts, maxc);
  out[1] = part_begin(b + 1, parts, maxc);
  int lo, hi;
  part_window(b, parts, maxc, dd, maxc + 1, &lo, &hi);
  out[2] = lo;
  out[3] = hi;
  *owner = part_of(c, parts, maxc);
  return KG_OK;
}

int kg_block_counts(kg_block* b, uint64_t* n_local, uint64_t* n_cells) {
  if (!b) return fail(KG_E_INVALID, "null block handle");
  if (n_local) *n_local = b->n_read;
  if (n_cells) *n_cells = b->ncells;
  return KG_OK;
}

}  // extern "C"
