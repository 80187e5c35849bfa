// DenseGrid2D<O> on the device: the object grid of krABMaga 0.6.1 on the cell-sorted layout.
//
// Replaces (reference paths relative to the krABMaga crate root, default variant):
//   DenseGrid2D::new                      src/engine/fields/dense_object_grid_2d.rs:201-214
//   set_object_location (replace-on-insert) :688-697      remove_object_location  :729-736
//   get_objects / _unbuffered  :507-520, :547-561         get_location / _unbuffered :429-441, :471-482
//   iter_objects / _unbuffered :589-608, :634-654         get_empty_bags :358-370
//   apply_to_all_values :258-328 (closure family, see krabgpu.h), incl. the y-major bag id it hands
//   the closure (calculate_indexes_bag :768-779)          Field::lazy_update :743-750
// Objects are (id, tag) pairs that compare by id (the fixture's Bird: bird.rs:168-172, tag ~ Bird.flag).
//
// HBM layout: each of the two buffers is a CSR over the width*height bags in flat order x*height + y
// (= iter_objects order): start[ncells + 1], id[], tag[].  Writes are appended to an op log
// (SET / REMOVE / PUSH, cell, id, tag) in call order; the log is folded into the write buffer's CSR
// when something reads the write side or at lazy_update: counting sort of (old bag content + log) by
// cell, each bag put back in sequence order, then the reference's sequential semantics applied per bag —
// an entry survives iff no later SET or REMOVE of an equal object follows it in its bag — and a
// compaction.  Nothing here is a hot loop of the BASELINE configs; the kernels are one thread per
// element or per bag.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace kg {

enum : uint32_t { OP_KEEP = 0, OP_SET = 1, OP_REMOVE = 2, OP_PUSH = 3 };  // KEEP = already in the bag

struct ObjCsr {
  uint32_t* start = nullptr;  // [ncells + 1]
  uint32_t* id = nullptr;
  uint32_t* tag = nullptr;
  uint32_t n = 0;
};
struct ObjOps {  // op log / combined work list
  uint32_t *op = nullptr, *cell = nullptr, *id = nullptr, *tag = nullptr;
};

__global__ void og_expand_kernel(uint32_t ncells, const uint32_t* __restrict__ start, const uint32_t* __restrict__ id,
                                 const uint32_t* __restrict__ tag, ObjOps out) {
  // old bag content -> work-list entries (op KEEP), one thread per bag
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
    out.op[k] = OP_KEEP;
    out.cell[k] = c;
    out.id[k] = id[k];
    out.tag[k] = tag[k];
  }
}
__global__ void og_hist_kernel(uint32_t m, const uint32_t* __restrict__ cell, uint32_t* __restrict__ count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) atomicAdd(&count[cell[i]], 1u);
}
// scatter by cell (any order inside a bag), remembering each entry's sequence number
__global__ void og_scatter_kernel(uint32_t m, ObjOps in, const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                  uint32_t* __restrict__ seq, ObjOps out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t c = in.cell[i];
  const uint32_t d = start[c] + atomicAdd(&cursor[c], 1u);
  seq[d] = i;
  out.op[d] = in.op[i];
  out.cell[d] = c;
  out.id[d] = in.id[i];
  out.tag[d] = in.tag[i];
}
// one thread per bag: back into sequence order, then the sequential semantics of the op log
__global__ void og_resolve_kernel(uint32_t ncells, const uint32_t* __restrict__ start, uint32_t* seq, ObjOps w,
                                  uint32_t* __restrict__ live, uint32_t* __restrict__ live_count) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t s = start[c], e = start[c + 1];
  for (uint32_t p = s + 1; p < e; ++p) {  // insertion sort by sequence number
    const uint32_t q0 = seq[p];
    if (seq[p - 1] <= q0) continue;
    const uint32_t o = w.op[p], i = w.id[p], t = w.tag[p];
    uint32_t q = p;
    while (q > s && seq[q - 1] > q0) {
      seq[q] = seq[q - 1]; w.op[q] = w.op[q - 1]; w.id[q] = w.id[q - 1]; w.tag[q] = w.tag[q - 1];
      --q;
    }
    seq[q] = q0; w.op[q] = o; w.id[q] = i; w.tag[q] = t;
  }
  uint32_t n = 0;
  for (uint32_t p = s; p < e; ++p) {
    bool alive = w.op[p] != OP_REMOVE;
    for (uint32_t q = p + 1; alive && q < e; ++q)
      if (w.id[q] == w.id[p] && (w.op[q] == OP_SET || w.op[q] == OP_REMOVE)) alive = false;  // retain(!= object)
    live[p] = alive ? 1u : 0u;
    n += alive ? 1u : 0u;
  }
  live_count[c] = n;
}
__global__ void og_compact_kernel(uint32_t ncells, const uint32_t* __restrict__ start, const uint32_t* __restrict__ live,
                                  ObjOps w, const uint32_t* __restrict__ new_start, uint32_t* __restrict__ id,
                                  uint32_t* __restrict__ tag) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint32_t d = new_start[c];
  for (uint32_t p = start[c]; p < start[c + 1]; ++p)
    if (live[p]) {
      id[d] = w.id[p];
      tag[d] = w.tag[p];
      ++d;
    }
}
__global__ void og_find_kernel(uint32_t ncells, const uint32_t* __restrict__ start, const uint32_t* __restrict__ id,
                               uint32_t target, uint32_t* first_cell) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (uint32_t k = start[c]; k < start[c + 1]; ++k)
    if (id[k] == target) {
      atomicMin(first_cell, c);
      return;
    }
}
__global__ void og_cells_kernel(uint32_t ncells, int32_t height, const uint32_t* __restrict__ start, int32_t* __restrict__ xs,
                                int32_t* __restrict__ ys) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
    xs[k] = (int32_t)(c / (uint32_t)height);
    ys[k] = (int32_t)(c % (uint32_t)height);
  }
}
__global__ void og_counts_kernel(uint32_t ncells, const uint32_t* __restrict__ start, uint32_t* __restrict__ out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncells) out[c] = start[c + 1] - start[c];
}

// ---------------------------------------------------------------- SparseGrid2D: keys -> bag slots
// SparseGrid2D<O> (sparse_object_grid_2d.rs:203-721) keeps two HashMap<Int2D, Vec<O>>: ANY Int2D is a key.
// On the device a bag is a slot of one open-addressing table (linear probing, 64-bit keys, shared by both
// buffers); everything else — op log, fold, CSR per buffer — is the dense grid's machinery with
// "cell" = slot.  A key whose bags are empty is the reference's absent key.
constexpr unsigned long long kKeyEmpty = ~0ull;
__host__ __device__ __forceinline__ unsigned long long sg_key(int32_t x, int32_t y) {
  // sign bits flipped: the all-ones pattern (the empty marker) is the single key (INT_MAX, INT_MAX)
  return ((((unsigned long long)(uint32_t)x) << 32) | (unsigned long long)(uint32_t)y) ^ 0x8000000080000000ull;
}
__host__ __device__ __forceinline__ void sg_unkey(unsigned long long k, int32_t* x, int32_t* y) {
  k ^= 0x8000000080000000ull;
  *x = (int32_t)(uint32_t)(k >> 32);
  *y = (int32_t)(uint32_t)k;
}
__device__ __forceinline__ uint32_t sg_hash(unsigned long long k, uint32_t mask) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
  return (uint32_t)k & mask;
}
constexpr uint32_t kSlotNone = 0xFFFFFFFFu;
// slot of key k; with `insert` an absent key claims the first free slot of its probe sequence
__device__ __forceinline__ uint32_t sg_slot(unsigned long long* keys, uint32_t mask, unsigned long long k, bool insert,
                                            uint32_t* used) {
  uint32_t h = sg_hash(k, mask);
  for (uint32_t probe = 0; probe <= mask; ++probe, h = (h + 1) & mask) {
    unsigned long long cur = keys[h];
    if (cur == k) return h;
    if (cur == kKeyEmpty) {
      if (!insert) return kSlotNone;
      cur = atomicCAS(&keys[h], kKeyEmpty, k);
      if (cur == kKeyEmpty) {
        atomicAdd(used, 1u);
        return h;
      }
      if (cur == k) return h;
    }
  }
  return kSlotNone;  // table full
}
__global__ void sg_slots_kernel(uint32_t n, const int32_t* __restrict__ x, const int32_t* __restrict__ y, int insert,
                                unsigned long long* keys, uint32_t mask, uint32_t* __restrict__ cell, uint32_t* used,
                                int* full) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = sg_slot(keys, mask, sg_key(x[i], y[i]), insert != 0, used);
  cell[i] = s;
  if (s == kSlotNone && insert) *full = 1;
}
// rehash: entries of a buffer move from their slot of the old table to the key's slot of the new one
__global__ void sg_remap_kernel(uint32_t m, uint32_t* __restrict__ cell, const unsigned long long* __restrict__ old_keys,
                                unsigned long long* new_keys, uint32_t mask, uint32_t* used, int* full) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t s = sg_slot(new_keys, mask, old_keys[cell[i]], true, used);
  if (s == kSlotNone) *full = 1;
  cell[i] = s == kSlotNone ? 0u : s;
}
__global__ void sg_cells_kernel(uint32_t nslots, const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ start,
                                int32_t* __restrict__ xs, int32_t* __restrict__ ys) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nslots) return;
  int32_t x, y;
  sg_unkey(keys[c], &x, &y);
  for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
    xs[k] = x;
    ys[k] = y;
  }
}
// bag sizes over the nominal width x height area (get_empty_bags :482-499)
__global__ void sg_area_sizes_kernel(int32_t width, int32_t height, unsigned long long* keys, uint32_t mask,
                                     const uint32_t* __restrict__ start, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)width * (uint32_t)height) return;
  const uint32_t s = sg_slot(keys, mask, sg_key((int32_t)(i / (uint32_t)height), (int32_t)(i % (uint32_t)height)), false, nullptr);
  out[i] = s == kSlotNone ? 0u : start[s + 1] - start[s];
}
// SparseGrid2D::apply_to_all_values :278-320.  The closure sees the bag's key; a None result is the
// reference's panic ("error on closure") and raises *none_seen.
__device__ __forceinline__ bool sg_closure(int op, uint32_t arg, unsigned long long key, uint32_t tag, uint32_t* new_tag) {
  switch (op) {
    case KG_OBJ_SET_TAG: *new_tag = arg; return true;
    case KG_OBJ_REMOVE: return false;
    case KG_OBJ_REMOVE_IF_TAG: *new_tag = tag; return tag != arg;
    default: {
      int32_t x, y;
      sg_unkey(key, &x, &y);
      *new_tag = (uint32_t)x * 65536u + (uint32_t)y;
      return true;
    }
  }
}
// READ / WRITE arms: every object of the side's bags is replaced by the closure's result
__global__ void sg_apply_inplace_kernel(uint32_t nslots, const unsigned long long* __restrict__ keys,
                                        const uint32_t* __restrict__ start, uint32_t* tag, int op, uint32_t arg,
                                        int* none_seen) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nslots) return;
  for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
    uint32_t t;
    if (sg_closure(op, arg, keys[c], tag[k], &t))
      tag[k] = t;
    else
      *none_seen = 1;
  }
}
// READWRITE arm :297-317: per read key — the write bag's objects in place if the key is in the write map,
// else a NEW write bag that ends up holding the closure's result for the LAST read object only
// (HashMap::insert replaces the bag on every iteration)
__global__ void sg_apply_readwrite_kernel(uint32_t nslots, const unsigned long long* __restrict__ keys,
                                          const uint32_t* __restrict__ rstart, const uint32_t* __restrict__ rid,
                                          const uint32_t* __restrict__ rtag, const uint32_t* __restrict__ wstart,
                                          uint32_t* wtag, int op, uint32_t arg, ObjOps log, uint32_t log_cap,
                                          uint32_t* log_n, unsigned long long* calls, int* overflow, int* none_seen) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nslots) return;
  const uint32_t rs = rstart[c], re = rstart[c + 1];
  if (rs == re) return;
  unsigned long long mine = 0;
  if (wstart[c + 1] > wstart[c]) {
    for (uint32_t k = wstart[c]; k < wstart[c + 1]; ++k) {
      uint32_t t;
      ++mine;
      if (sg_closure(op, arg, keys[c], wtag[k], &t)) wtag[k] = t; else *none_seen = 1;
    }
  } else {
    uint32_t t = 0;
    bool ok = true;
    for (uint32_t k = rs; k < re; ++k) {
      ++mine;
      ok = sg_closure(op, arg, keys[c], rtag[k], &t) && ok;
    }
    if (!ok) {
      *none_seen = 1;
    } else {
      const uint32_t slot = atomicAdd(log_n, 1u);
      if (slot >= log_cap) {
        *overflow = 1;
      } else {
        log.op[slot] = OP_PUSH;
        log.cell[slot] = c;
        log.id[slot] = rid[re - 1];
        log.tag[slot] = t;
      }
    }
  }
  atomicAdd(calls, mine);
}

// the closure family of apply_to_all_values: returns false for None
__device__ __forceinline__ bool og_closure(int op, uint32_t arg, uint32_t flat, int32_t width, uint32_t id, uint32_t tag,
                                           uint32_t* new_tag) {
  (void)id;
  switch (op) {
    case KG_OBJ_SET_TAG: *new_tag = arg; return true;
    case KG_OBJ_REMOVE: return false;
    case KG_OBJ_REMOVE_IF_TAG: *new_tag = tag; return tag != arg;
    default: {  // KG_OBJ_TAG_WITH_BAG_ID: calculate_indexes_bag(i, width, height) = (i - width*row, row)  :768-779
      const uint32_t row = flat / (uint32_t)width;
      *new_tag = (flat - (uint32_t)width * row) * 65536u + row;
      return true;
    }
  }
}
// READ arm :263-277: every read bag becomes the closure's Some(..) results; keep[] marks them
__global__ void og_apply_read_kernel(uint32_t ncells, int32_t width, const uint32_t* __restrict__ start, uint32_t* id,
                                     uint32_t* tag, int op, uint32_t arg, uint32_t* __restrict__ keep,
                                     uint32_t* __restrict__ keep_count) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint32_t n = 0;
  for (uint32_t k = start[c]; k < start[c + 1]; ++k) {
    uint32_t t;
    const bool some = og_closure(op, arg, c, width, id[k], tag[k], &t);
    if (some) tag[k] = t;
    keep[k] = some ? 1u : 0u;
    n += some ? 1u : 0u;
  }
  keep_count[c] = n;
}
__global__ void og_keep_compact_kernel(uint32_t ncells, const uint32_t* __restrict__ start, const uint32_t* __restrict__ keep,
                                       const uint32_t* __restrict__ id, const uint32_t* __restrict__ tag,
                                       const uint32_t* __restrict__ new_start, uint32_t* __restrict__ nid,
                                       uint32_t* __restrict__ ntag) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint32_t d = new_start[c];
  for (uint32_t k = start[c]; k < start[c + 1]; ++k)
    if (keep[k]) {
      nid[d] = id[k];
      ntag[d] = tag[k];
      ++d;
    }
}
// WRITE arm :278-291: results of the read bags are pushed (plain push) into the write bags = PUSH ops;
// READWRITE arm :293-326 (write side resolved): a non-empty write bag is updated in place, an empty
// one receives the read bag's results without duplicates.  log_n is the log's device-side length.
__global__ void og_apply_push_kernel(uint32_t ncells, int32_t width, int readwrite, const uint32_t* __restrict__ rstart,
                                     const uint32_t* __restrict__ rid, const uint32_t* __restrict__ rtag,
                                     const uint32_t* __restrict__ wstart, uint32_t* wtag, const uint32_t* __restrict__ wid,
                                     int op, uint32_t arg, ObjOps log, uint32_t log_cap, uint32_t* log_n,
                                     unsigned long long* calls, int* overflow) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  unsigned long long mine = 0;
  if (readwrite && wstart[c + 1] > wstart[c]) {
    for (uint32_t k = wstart[c]; k < wstart[c + 1]; ++k) {
      uint32_t t;
      ++mine;
      if (og_closure(op, arg, c, width, wid[k], wtag[k], &t)) wtag[k] = t;  // None leaves the element as it is
    }
  } else {
    const uint32_t s = rstart[c], e = rstart[c + 1];
    for (uint32_t k = s; k < e; ++k) {
      uint32_t t;
      ++mine;
      if (!og_closure(op, arg, c, width, rid[k], rtag[k], &t)) continue;
      if (readwrite) {  // `if !wlocs[i].contains(&result)`: equal objects already pushed from this bag
        bool dup = false;
        for (uint32_t q = s; q < k && !dup; ++q) {
          uint32_t tq;
          dup = rid[q] == rid[k] && og_closure(op, arg, c, width, rid[q], rtag[q], &tq);
        }
        if (dup) continue;
      }
      const uint32_t slot = atomicAdd(log_n, 1u);
      if (slot >= log_cap) {
        *overflow = 1;
        continue;
      }
      // slots are handed out in arbitrary order across bags; inside one bag this thread takes them in
      // bag order, which is all the per-bag sequence needs
      log.op[slot] = OP_PUSH;
      log.cell[slot] = c;
      log.id[slot] = rid[k];
      log.tag[slot] = t;
    }
  }
  if (mine) atomicAdd(calls, mine);
}

}  // namespace kg

using namespace kg;

struct kg_objgrid {
  int device = 0;
  cudaStream_t stream = nullptr;
  int32_t width = 0, height = 0;
  uint32_t ncells = 0;
  uint64_t capacity = 0;
  ObjCsr buf[2];
  int read = 0, write = 1;
  ObjOps log;         // pending ops on the write side, in call order
  uint32_t nlog = 0;  // host mirror of the log length
  uint32_t* d_nlog = nullptr;
  ObjOps work, sorted;
  uint32_t *seq = nullptr, *live = nullptr, *count = nullptr, *cursor = nullptr, *new_start = nullptr,
           *tile_sums = nullptr, *tmp_id = nullptr, *tmp_tag = nullptr;
  unsigned long long* d_calls = nullptr;
  int* d_flag = nullptr;
  uint32_t* d_first = nullptr;
  // SparseGrid2D mode: bag = slot of a key table (ncells = number of slots)
  bool sparse = false;
  unsigned long long *keys = nullptr, *keys_alt = nullptr;
  uint32_t tmask = 0;
  uint32_t* d_used = nullptr;  // claimed slots (device counter), host mirror below
  uint32_t used = 0;
  int32_t *d_qx = nullptr, *d_qy = nullptr;  // coordinates of one batch of calls
  uint32_t* d_slot = nullptr;
  uint32_t* area_buf = nullptr;  // scratch of bag_sizes over the nominal area, grown on demand
  size_t area_cap = 0;
};

namespace {

constexpr int kOT = 128;
inline unsigned oblk(uint64_t n) { return (unsigned)std::max<uint64_t>(1, (n + kOT - 1) / kOT); }
int ouse(kg_objgrid* g) {
  if (!g) return fail(KG_E_INVALID, "null object-grid handle");
  KG_CUDA(cudaSetDevice(g->device));
  return KG_OK;
}
int alloc_ops(ObjOps& o, uint64_t n) {
  KG_CUDA(cudaMalloc(&o.op, n * 4));
  KG_CUDA(cudaMalloc(&o.cell, n * 4));
  KG_CUDA(cudaMalloc(&o.id, n * 4));
  KG_CUDA(cudaMalloc(&o.tag, n * 4));
  return KG_OK;
}
void free_ops(ObjOps& o) {
  cudaFree(o.op); cudaFree(o.cell); cudaFree(o.id); cudaFree(o.tag);
  o = ObjOps{};
}
#define OLAUNCH(g, kernel, grid, ...)                                         \
  do {                                                                        \
    kernel<<<grid, kOT, 0, (g)->stream>>>(__VA_ARGS__);                       \
    launch_counter().fetch_add(1, std::memory_order_relaxed);                 \
  } while (0)

// fold `nlog` entries of the op log into buffer w's CSR.  `new_keys` (sparse rehash only): every entry
// first moves to its key's slot of that table.
int fold(kg_objgrid* g, ObjCsr& w, uint32_t nlog, unsigned long long* new_keys) {
  if (nlog == 0 && !new_keys) return KG_OK;
  const uint64_t m = (uint64_t)w.n + nlog;
  if (m > g->capacity) return fail(KG_E_CAPACITY, "object grid: %llu entries exceed the capacity %llu",
                                   (unsigned long long)m, (unsigned long long)g->capacity);
  cudaStream_t s = g->stream;
  // work list = old bags (KEEP) followed by the log
  if (w.n) OLAUNCH(g, og_expand_kernel, oblk(g->ncells), g->ncells, w.start, w.id, w.tag, g->work);
  if (nlog) {
    KG_CUDA(cudaMemcpyAsync(g->work.op + w.n, g->log.op, (size_t)nlog * 4, cudaMemcpyDeviceToDevice, s));
    KG_CUDA(cudaMemcpyAsync(g->work.cell + w.n, g->log.cell, (size_t)nlog * 4, cudaMemcpyDeviceToDevice, s));
    KG_CUDA(cudaMemcpyAsync(g->work.id + w.n, g->log.id, (size_t)nlog * 4, cudaMemcpyDeviceToDevice, s));
    KG_CUDA(cudaMemcpyAsync(g->work.tag + w.n, g->log.tag, (size_t)nlog * 4, cudaMemcpyDeviceToDevice, s));
  }
  if (new_keys && m)
    OLAUNCH(g, sg_remap_kernel, oblk(m), (uint32_t)m, g->work.cell, g->keys, new_keys, g->tmask, g->d_used, g->d_flag);
  KG_CUDA(cudaMemsetAsync(g->count, 0, ((size_t)g->ncells + 1) * 4, s));
  KG_CUDA(cudaMemsetAsync(g->cursor, 0, ((size_t)g->ncells + 1) * 4, s));
  if (m) OLAUNCH(g, og_hist_kernel, oblk(m), (uint32_t)m, g->work.cell, g->count);
  exclusive_scan_u32(g->count, g->ncells, g->new_start, g->tile_sums, s);  // new_start = offsets of the unresolved bags
  launch_counter().fetch_add(3, std::memory_order_relaxed);
  if (m) OLAUNCH(g, og_scatter_kernel, oblk(m), (uint32_t)m, g->work, g->new_start, g->cursor, g->seq, g->sorted);
  OLAUNCH(g, og_resolve_kernel, oblk(g->ncells), g->ncells, g->new_start, g->seq, g->sorted, g->live, g->count);
  exclusive_scan_u32(g->count, g->ncells, w.start, g->tile_sums, s);
  launch_counter().fetch_add(3, std::memory_order_relaxed);
  OLAUNCH(g, og_compact_kernel, oblk(g->ncells), g->ncells, g->new_start, g->live, g->sorted, w.start, w.id, w.tag);
  uint32_t n = 0;
  KG_CUDA(cudaMemcpyAsync(&n, w.start + g->ncells, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  w.n = n;
  return KG_OK;
}
// fold the op log into the write buffer's CSR
int resolve_write(kg_objgrid* g) {
  if (g->nlog == 0) return KG_OK;
  KG_TRY(fold(g, g->buf[g->write], g->nlog, nullptr));
  g->nlog = 0;
  KG_CUDA(cudaMemsetAsync(g->d_nlog, 0, 4, g->stream));
  return KG_OK;
}
// sparse: rebuild the key table from the keys whose bags hold something (emptied keys are the garbage)
int sg_rehash(kg_objgrid* g) {
  KG_TRY(resolve_write(g));
  cudaStream_t s = g->stream;
  KG_CUDA(cudaMemsetAsync(g->keys_alt, 0xFF, ((size_t)g->tmask + 1) * 8, s));
  KG_CUDA(cudaMemsetAsync(g->d_used, 0, 4, s));
  KG_CUDA(cudaMemsetAsync(g->d_flag, 0, 4, s));
  for (int k = 0; k < 2; ++k) KG_TRY(fold(g, g->buf[k], 0, g->keys_alt));
  std::swap(g->keys, g->keys_alt);
  int full = 0;
  KG_CUDA(cudaMemcpyAsync(&g->used, g->d_used, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(&full, g->d_flag, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  if (full) return fail(KG_E_CAPACITY, "sparse object grid: key table full");
  return KG_OK;
}
// sparse: slots of a batch of keys (host arrays), written to `cell_out` (device)
int sg_slots(kg_objgrid* g, uint64_t n, const int32_t* x, const int32_t* y, bool insert, uint32_t* cell_out) {
  for (uint64_t i = 0; i < n; ++i)
    if (sg_key(x[i], y[i]) == kKeyEmpty) return fail(KG_E_INVALID, "SparseGrid2D: (i32::MAX, i32::MAX) is reserved");
  cudaStream_t s = g->stream;
  KG_CUDA(cudaMemcpyAsync(g->d_qx, x, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(g->d_qy, y, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemsetAsync(g->d_flag, 0, 4, s));
  OLAUNCH(g, sg_slots_kernel, oblk(n), (uint32_t)n, g->d_qx, g->d_qy, insert ? 1 : 0, g->keys, g->tmask, cell_out,
          g->d_used, g->d_flag);
  int full = 0;
  KG_CUDA(cudaMemcpyAsync(&full, g->d_flag, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(&g->used, g->d_used, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  if (full) return fail(KG_E_CAPACITY, "sparse object grid: key table full");
  return KG_OK;
}
// sparse: slot of one key, kSlotNone when the key was never seen
int sg_lookup(kg_objgrid* g, int32_t x, int32_t y, uint32_t* slot) {
  KG_TRY(sg_slots(g, 1, &x, &y, false, g->d_slot));
  KG_CUDA(cudaMemcpyAsync(slot, g->d_slot, 4, cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  return KG_OK;
}

int flat_index(const kg_objgrid* g, int32_t x, int32_t y, uint32_t* out, const char* who) {
  // ((loc.x * self.height) + loc.y) as usize, bounds-checked against the Vec only (no per-axis check)
  const int64_t idx = (int64_t)x * g->height + y;
  if (idx < 0 || idx >= (int64_t)g->ncells)
    return fail(KG_E_OOB, "DenseGrid2D::%s: location (%d, %d) outside the bag Vec (reference: index out of bounds panic)",
                who, x, y);
  *out = (uint32_t)idx;
  return KG_OK;
}

int append_ops(kg_objgrid* g, uint32_t op, uint64_t n, const uint32_t* id, const uint32_t* tag, const int32_t* x,
               const int32_t* y, const char* who) {
  if (n == 0) return KG_OK;
  if (!id || !x || !y) return fail(KG_E_INVALID, "null argument");
  std::vector<uint32_t> cells(n), ops(n, op), tags(n, 0u);
  if (!g->sparse)
    for (uint64_t i = 0; i < n; ++i) KG_TRY(flat_index(g, x[i], y[i], &cells[i], who));
  if (tag) std::copy(tag, tag + n, tags.begin());
  if ((uint64_t)g->nlog + n + g->buf[g->write].n > g->capacity) {
    KG_TRY(resolve_write(g));  // replaced / removed objects free their entries
    if ((uint64_t)g->nlog + n + g->buf[g->write].n > g->capacity)
      return fail(KG_E_CAPACITY, "object grid: capacity %llu exceeded", (unsigned long long)g->capacity);
  }
  cudaStream_t s = g->stream;
  if (g->sparse) {  // the keys claim their slots on the device (a removed key that was never set gets an empty bag)
    if ((uint64_t)g->used + n > ((uint64_t)g->tmask + 1) / 2) KG_TRY(sg_rehash(g));  // folds the log first
    KG_TRY(sg_slots(g, n, x, y, true, g->log.cell + g->nlog));
  }
  else
    KG_CUDA(cudaMemcpyAsync(g->log.cell + g->nlog, cells.data(), n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(g->log.op + g->nlog, ops.data(), n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(g->log.id + g->nlog, id, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(g->log.tag + g->nlog, tags.data(), n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaStreamSynchronize(s));  // the host vectors die with this call
  g->nlog += (uint32_t)n;
  KG_CUDA(cudaMemcpyAsync(g->d_nlog, &g->nlog, 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaStreamSynchronize(s));
  return KG_OK;
}

ObjCsr* side(kg_objgrid* g, int which) { return &g->buf[which == KG_BUF_READ ? g->read : g->write]; }

}  // namespace

extern "C" {

static int objgrid_create(int32_t width, int32_t height, uint64_t capacity, int device, bool sparse, kg_objgrid** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  int64_t nc = (int64_t)width * (int64_t)height;  // the Vec length uses width*height before the abs() (:203-211)
  if (!sparse && (nc < 0 || nc >= (1ll << 31))) return fail(KG_E_INVALID, "DenseGrid2D::new: capacity overflow");
  if (capacity == 0 || capacity >= (1ull << 27)) return fail(KG_E_INVALID, "bad capacity");
  if (sparse) {  // bags = slots of the key table: a power of two >= 8 x capacity (load <= 1/2 between rehashes)
    if (nc < 0 || nc >= (1ll << 31)) return fail(KG_E_INVALID, "SparseGrid2D::new: width * height overflow");
    nc = 64;
    while ((uint64_t)nc < 8 * capacity) nc <<= 1;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_objgrid* g = new kg_objgrid();
  g->device = device;
  g->sparse = sparse;
  g->width = sparse ? width : (width < 0 ? -width : width);  // SparseGrid2D::new keeps the sign (:224-234)
  g->height = sparse ? height : (height < 0 ? -height : height);
  g->ncells = (uint32_t)nc;
  g->tmask = sparse ? (uint32_t)nc - 1u : 0u;
  g->capacity = capacity;
  auto bail = [&](int code) { kg_objgrid_destroy(g); return code; };
  if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  const size_t cells = (size_t)g->ncells + 16, cap = (size_t)capacity + 16;
  bool ok = true;
  for (int k = 0; k < 2 && ok; ++k)
    ok = cudaMalloc(&g->buf[k].start, cells * 4) == cudaSuccess && cudaMalloc(&g->buf[k].id, cap * 4) == cudaSuccess &&
         cudaMalloc(&g->buf[k].tag, cap * 4) == cudaSuccess;
  ok = ok && alloc_ops(g->log, cap) == KG_OK && alloc_ops(g->work, cap) == KG_OK && alloc_ops(g->sorted, cap) == KG_OK;
  ok = ok && cudaMalloc(&g->seq, cap * 4) == cudaSuccess && cudaMalloc(&g->live, cap * 4) == cudaSuccess &&
       cudaMalloc(&g->count, cells * 4) == cudaSuccess && cudaMalloc(&g->cursor, cells * 4) == cudaSuccess &&
       cudaMalloc(&g->new_start, cells * 4) == cudaSuccess && cudaMalloc(&g->tmp_id, cap * 4) == cudaSuccess &&
       cudaMalloc(&g->tmp_tag, cap * 4) == cudaSuccess &&
       cudaMalloc(&g->tile_sums, ((size_t)scan_num_tiles(std::max<uint64_t>(cells, cap)) + 16) * 4) == cudaSuccess &&
       cudaMalloc(&g->d_calls, 8) == cudaSuccess && cudaMalloc(&g->d_flag, 8) == cudaSuccess &&
       cudaMalloc(&g->d_first, 4) == cudaSuccess && cudaMalloc(&g->d_nlog, 4) == cudaSuccess;
  if (sparse)
    ok = ok && cudaMalloc(&g->keys, cells * 8) == cudaSuccess && cudaMalloc(&g->keys_alt, cells * 8) == cudaSuccess &&
         cudaMalloc(&g->d_used, 4) == cudaSuccess && cudaMalloc(&g->d_qx, cap * 4) == cudaSuccess &&
         cudaMalloc(&g->d_qy, cap * 4) == cudaSuccess && cudaMalloc(&g->d_slot, 16) == cudaSuccess;
  if (!ok) return bail(fail(KG_E_CUDA, "object grid allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  if (sparse) {
    cudaMemsetAsync(g->keys, 0xFF, cells * 8, g->stream);
    cudaMemsetAsync(g->d_used, 0, 4, g->stream);
  }
  for (int k = 0; k < 2; ++k) cudaMemsetAsync(g->buf[k].start, 0, cells * 4, g->stream);
  cudaMemsetAsync(g->d_nlog, 0, 4, g->stream);
  if (cudaStreamSynchronize(g->stream) != cudaSuccess) return bail(fail(KG_E_CUDA, "object grid init failed"));
  *out = g;
  return KG_OK;
}

int kg_objgrid_create(int32_t width, int32_t height, uint64_t capacity, int device, kg_objgrid** out) {
  return objgrid_create(width, height, capacity, device, false, out);
}
int kg_objgrid_create_sparse(int32_t width, int32_t height, uint64_t capacity, int device, kg_objgrid** out) {
  return objgrid_create(width, height, capacity, device, true, out);
}

int kg_objgrid_destroy(kg_objgrid* g) {
  if (!g) return KG_OK;
  cudaSetDevice(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  cudaFree(g->keys); cudaFree(g->keys_alt); cudaFree(g->d_used); cudaFree(g->d_qx); cudaFree(g->d_qy);
  cudaFree(g->d_slot); cudaFree(g->area_buf);
  for (int k = 0; k < 2; ++k) {
    cudaFree(g->buf[k].start); cudaFree(g->buf[k].id); cudaFree(g->buf[k].tag);
  }
  free_ops(g->log); free_ops(g->work); free_ops(g->sorted);
  cudaFree(g->seq); cudaFree(g->live); cudaFree(g->count); cudaFree(g->cursor); cudaFree(g->new_start);
  cudaFree(g->tmp_id); cudaFree(g->tmp_tag); cudaFree(g->tile_sums); cudaFree(g->d_calls); cudaFree(g->d_flag);
  cudaFree(g->d_first); cudaFree(g->d_nlog);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return KG_OK;
}

int kg_objgrid_set_object_locations(kg_objgrid* g, uint64_t n, const uint32_t* id, const uint32_t* tag,
                                    const int32_t* x, const int32_t* y) {
  KG_TRY(ouse(g));
  // SparseGrid2D pushes (:648-659); DenseGrid2D replaces an equal object (:688-697)
  return append_ops(g, g->sparse ? OP_PUSH : OP_SET, n, id, tag, x, y, "set_object_location");
}
int kg_objgrid_remove_object_locations(kg_objgrid* g, uint64_t n, const uint32_t* id, const int32_t* x,
                                       const int32_t* y) {
  KG_TRY(ouse(g));
  return append_ops(g, OP_REMOVE, n, id, nullptr, x, y, "remove_object_location");
}

int kg_objgrid_lazy_update(kg_objgrid* g) {
  KG_TRY(ouse(g));
  KG_TRY(resolve_write(g));
  std::swap(g->read, g->write);
  ObjCsr& w = g->buf[g->write];  // every write bag is cleared (:746-749)
  w.n = 0;
  KG_CUDA(cudaMemsetAsync(w.start, 0, ((size_t)g->ncells + 1) * 4, g->stream));
  return KG_OK;
}
int kg_objgrid_update(kg_objgrid* g) {
  KG_TRY(ouse(g));
  if (g->sparse) {  // SparseGrid2D::update :711-718 — read = clone of write, write cleared
    KG_TRY(resolve_write(g));
    ObjCsr &r = g->buf[g->read], &w = g->buf[g->write];
    cudaStream_t s = g->stream;
    KG_CUDA(cudaMemcpyAsync(r.start, w.start, ((size_t)g->ncells + 1) * 4, cudaMemcpyDeviceToDevice, s));
    if (w.n) {
      KG_CUDA(cudaMemcpyAsync(r.id, w.id, (size_t)w.n * 4, cudaMemcpyDeviceToDevice, s));
      KG_CUDA(cudaMemcpyAsync(r.tag, w.tag, (size_t)w.n * 4, cudaMemcpyDeviceToDevice, s));
    }
    r.n = w.n;
    w.n = 0;
    KG_CUDA(cudaMemsetAsync(w.start, 0, ((size_t)g->ncells + 1) * 4, s));
    return KG_OK;
  }
  return fail(KG_E_INVALID,
              "DenseGrid2D::update is not offered: the reference's Vec::insert (dense_object_grid_2d.rs:753-763) "
              "doubles the read Vec and its own apply_to_all_values then panics; use lazy_update");
}

int kg_objgrid_num_objects(kg_objgrid* g, int which, uint64_t* out) {
  KG_TRY(ouse(g));
  if (!out) return fail(KG_E_INVALID, "null out");
  if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
  *out = side(g, which)->n;
  return KG_OK;
}

int kg_objgrid_get_objects(kg_objgrid* g, int which, int32_t x, int32_t y, uint64_t cap, uint32_t* id, uint32_t* tag,
                           uint64_t* n_out) {
  KG_TRY(ouse(g));
  uint32_t c;
  if (g->sparse) {
    KG_TRY(sg_lookup(g, x, y, &c));
    if (c == kSlotNone) {  // never a key: Option::None
      if (n_out) *n_out = 0;
      return KG_OK;
    }
  } else {
    KG_TRY(flat_index(g, x, y, &c, "get_objects"));
  }
  if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
  const ObjCsr* b = side(g, which);
  uint32_t se[2];
  KG_CUDA(cudaMemcpyAsync(se, b->start + c, 8, cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  const uint64_t n = se[1] - se[0];
  if (n_out) *n_out = n;  // 0 = Option::None (:512-518)
  if (n > cap) return fail(KG_E_CAPACITY, "get_objects needs room for %llu objects", (unsigned long long)n);
  if (n && id) KG_CUDA(cudaMemcpyAsync(id, b->id + se[0], n * 4, cudaMemcpyDeviceToHost, g->stream));
  if (n && tag) KG_CUDA(cudaMemcpyAsync(tag, b->tag + se[0], n * 4, cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  return KG_OK;
}

int kg_objgrid_get_location(kg_objgrid* g, int which, uint32_t id, int32_t* x, int32_t* y, int* found) {
  KG_TRY(ouse(g));
  if (!x || !y || !found) return fail(KG_E_INVALID, "null argument");
  *found = 0;
  if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
  const ObjCsr* b = side(g, which);
  if (b->n == 0 || g->ncells == 0) return KG_OK;
  KG_CUDA(cudaMemsetAsync(g->d_first, 0xFF, 4, g->stream));
  OLAUNCH(g, og_find_kernel, oblk(g->ncells), g->ncells, b->start, b->id, id, g->d_first);
  uint32_t c = 0xFFFFFFFFu;
  KG_CUDA(cudaMemcpyAsync(&c, g->d_first, 4, cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  if (c != 0xFFFFFFFFu && g->sparse) {  // any bag holding it: the reference's HashMap order is unspecified (:346-356)
    unsigned long long key = 0;
    KG_CUDA(cudaMemcpyAsync(&key, g->keys + c, 8, cudaMemcpyDeviceToHost, g->stream));
    KG_CUDA(cudaStreamSynchronize(g->stream));
    *found = 1;
    sg_unkey(key, x, y);
  } else if (c != 0xFFFFFFFFu) {  // first bag in x-outer / y-inner order (:431-438)
    *found = 1;
    *x = (int32_t)(c / (uint32_t)g->height);
    *y = (int32_t)(c % (uint32_t)g->height);
  }
  return KG_OK;
}

int kg_objgrid_iter_objects(kg_objgrid* g, int which, uint64_t cap, int32_t* x, int32_t* y, uint32_t* id, uint32_t* tag,
                            uint64_t* n_out) {
  KG_TRY(ouse(g));
  if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
  const ObjCsr* b = side(g, which);
  if (n_out) *n_out = b->n;
  if (b->n > cap) return fail(KG_E_CAPACITY, "iter_objects needs room for %u objects", b->n);
  if (b->n == 0) return KG_OK;
  int32_t* dx = (int32_t*)g->seq;
  int32_t* dy = (int32_t*)g->live;
  if (g->sparse)
    OLAUNCH(g, sg_cells_kernel, oblk(g->ncells), g->ncells, g->keys, b->start, dx, dy);
  else
    OLAUNCH(g, og_cells_kernel, oblk(g->ncells), g->ncells, g->height, b->start, dx, dy);
  cudaStream_t s = g->stream;
  if (x) KG_CUDA(cudaMemcpyAsync(x, dx, (size_t)b->n * 4, cudaMemcpyDeviceToHost, s));
  if (y) KG_CUDA(cudaMemcpyAsync(y, dy, (size_t)b->n * 4, cudaMemcpyDeviceToHost, s));
  if (id) KG_CUDA(cudaMemcpyAsync(id, b->id, (size_t)b->n * 4, cudaMemcpyDeviceToHost, s));
  if (tag) KG_CUDA(cudaMemcpyAsync(tag, b->tag, (size_t)b->n * 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  return KG_OK;
}

int kg_objgrid_bag_sizes(kg_objgrid* g, int which, uint64_t cap, uint32_t* sizes) {
  KG_TRY(ouse(g));
  if (!sizes) return fail(KG_E_INVALID, "null out");
  if (g->sparse) {  // sizes over the nominal [0, width) x [0, height) area (get_empty_bags :482-499)
    const int64_t area = g->width > 0 && g->height > 0 ? (int64_t)g->width * g->height : 0;
    if ((int64_t)cap < area) return fail(KG_E_CAPACITY, "bag_sizes needs %lld entries", (long long)area);
    if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
    if (area == 0) return KG_OK;
    if (g->area_cap < (size_t)area) {  // the area is unrelated to the table size: a scratch of its own, kept in the handle
      if (g->area_buf) cudaFree(g->area_buf);
      g->area_buf = nullptr;
      g->area_cap = 0;
      KG_CUDA(cudaMalloc(&g->area_buf, (size_t)area * 4));
      g->area_cap = (size_t)area;
    }
    OLAUNCH(g, sg_area_sizes_kernel, oblk((uint64_t)area), g->width, g->height, g->keys, g->tmask, side(g, which)->start,
            g->area_buf);
    KG_CUDA(cudaMemcpyAsync(sizes, g->area_buf, (size_t)area * 4, cudaMemcpyDeviceToHost, g->stream));
    KG_CUDA(cudaStreamSynchronize(g->stream));
    return KG_OK;
  }
  if (cap < g->ncells) return fail(KG_E_CAPACITY, "bag_sizes needs %u entries", g->ncells);
  if (which != KG_BUF_READ) KG_TRY(resolve_write(g));
  if (g->ncells == 0) return KG_OK;
  OLAUNCH(g, og_counts_kernel, oblk(g->ncells), g->ncells, side(g, which)->start, g->count);
  KG_CUDA(cudaMemcpyAsync(sizes, g->count, (size_t)g->ncells * 4, cudaMemcpyDeviceToHost, g->stream));
  KG_CUDA(cudaStreamSynchronize(g->stream));
  return KG_OK;
}

int kg_objgrid_apply(kg_objgrid* g, int op, uint32_t arg, int option, uint64_t* calls_out) {
  KG_TRY(ouse(g));
  if (op < KG_OBJ_SET_TAG || op > KG_OBJ_TAG_WITH_BAG_ID) return fail(KG_E_INVALID, "bad apply op");
  if (option < KG_GRID_READ || option > KG_GRID_READWRITE) return fail(KG_E_INVALID, "bad GridOption");
  if (calls_out) *calls_out = 0;
  if (g->ncells == 0) return KG_OK;
  cudaStream_t s = g->stream;
  ObjCsr& r = g->buf[g->read];
  if (g->sparse) {  // sparse_object_grid_2d.rs:278-320
    KG_CUDA(cudaMemsetAsync(g->d_flag, 0, 4, s));
    KG_CUDA(cudaMemsetAsync(g->d_flag + 1, 0, 4, s));
    unsigned long long calls = 0;
    int flags[2] = {0, 0};
    if (option != KG_GRID_READ) KG_TRY(resolve_write(g));
    ObjCsr& w = g->buf[g->write];
    if (option == KG_GRID_READWRITE) {
      KG_CUDA(cudaMemsetAsync(g->d_calls, 0, 8, s));
      const uint32_t log_cap = (uint32_t)std::min<uint64_t>(g->capacity - w.n, 0xFFFFFFF0ull);
      OLAUNCH(g, sg_apply_readwrite_kernel, oblk(g->ncells), g->ncells, g->keys, r.start, r.id, r.tag, w.start, w.tag, op,
              arg, g->log, log_cap, g->d_nlog, g->d_calls, g->d_flag + 1, g->d_flag);
      KG_CUDA(cudaMemcpyAsync(&calls, g->d_calls, 8, cudaMemcpyDeviceToHost, s));
      KG_CUDA(cudaMemcpyAsync(&g->nlog, g->d_nlog, 4, cudaMemcpyDeviceToHost, s));
    } else {
      ObjCsr& b = option == KG_GRID_READ ? r : w;
      calls = b.n;
      if (b.n) OLAUNCH(g, sg_apply_inplace_kernel, oblk(g->ncells), g->ncells, g->keys, b.start, b.tag, op, arg, g->d_flag);
    }
    KG_CUDA(cudaMemcpyAsync(flags, g->d_flag, 8, cudaMemcpyDeviceToHost, s));
    KG_CUDA(cudaStreamSynchronize(s));
    if (flags[1]) {
      g->nlog = (uint32_t)std::min<uint64_t>(g->nlog, g->capacity - w.n);
      KG_CUDA(cudaMemcpyAsync(g->d_nlog, &g->nlog, 4, cudaMemcpyHostToDevice, s));
      return fail(KG_E_CAPACITY, "sparse object grid: apply_to_all_values overflowed the capacity");
    }
    if (flags[0]) return fail(KG_E_INVALID, "SparseGrid2D::apply_to_all_values: error on closure (it returned None)");
    if (calls_out) *calls_out = calls;
    return KG_OK;
  }
  if (option == KG_GRID_READ) {
    if (calls_out) *calls_out = r.n;  // one closure call per object of the read bags
    if (r.n == 0) return KG_OK;
    OLAUNCH(g, og_apply_read_kernel, oblk(g->ncells), g->ncells, g->width, r.start, r.id, r.tag, op, arg, g->live,
            g->count);
    if (op == KG_OBJ_REMOVE || op == KG_OBJ_REMOVE_IF_TAG) {  // the bags shrink: compact the CSR
      exclusive_scan_u32(g->count, g->ncells, g->new_start, g->tile_sums, s);
      launch_counter().fetch_add(3, std::memory_order_relaxed);
      OLAUNCH(g, og_keep_compact_kernel, oblk(g->ncells), g->ncells, r.start, g->live, r.id, r.tag, g->new_start,
              g->tmp_id, g->tmp_tag);
      uint32_t n = 0;
      KG_CUDA(cudaMemcpyAsync(&n, g->new_start + g->ncells, 4, cudaMemcpyDeviceToHost, s));
      KG_CUDA(cudaStreamSynchronize(s));
      KG_CUDA(cudaMemcpyAsync(r.start, g->new_start, ((size_t)g->ncells + 1) * 4, cudaMemcpyDeviceToDevice, s));
      if (n) {
        KG_CUDA(cudaMemcpyAsync(r.id, g->tmp_id, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        KG_CUDA(cudaMemcpyAsync(r.tag, g->tmp_tag, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
      }
      r.n = n;
    }
    return KG_OK;
  }
  const int readwrite = option == KG_GRID_READWRITE;
  if (readwrite) KG_TRY(resolve_write(g));  // "is the write bag empty?" needs the bags themselves
  ObjCsr& w = g->buf[g->write];
  KG_CUDA(cudaMemsetAsync(g->d_calls, 0, 8, s));
  KG_CUDA(cudaMemsetAsync(g->d_flag, 0, 4, s));
  const uint32_t log_cap = (uint32_t)std::min<uint64_t>(g->capacity - w.n, 0xFFFFFFF0ull);
  OLAUNCH(g, og_apply_push_kernel, oblk(g->ncells), g->ncells, g->width, readwrite, r.start, r.id, r.tag, w.start, w.tag,
          w.id, op, arg, g->log, log_cap, g->d_nlog, g->d_calls, g->d_flag);
  unsigned long long calls = 0;
  int overflow = 0;
  uint32_t nlog = 0;
  KG_CUDA(cudaMemcpyAsync(&calls, g->d_calls, 8, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(&overflow, g->d_flag, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaMemcpyAsync(&nlog, g->d_nlog, 4, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  if (overflow) {
    nlog = log_cap;
    KG_CUDA(cudaMemcpyAsync(g->d_nlog, &nlog, 4, cudaMemcpyHostToDevice, s));
    g->nlog = nlog;
    return fail(KG_E_CAPACITY, "object grid: apply_to_all_values overflowed the capacity %llu",
                (unsigned long long)g->capacity);
  }
  g->nlog = nlog;
  if (calls_out) *calls_out = calls;
  return KG_OK;
}

int kg_objgrid_dims(kg_objgrid* g, int32_t* width, int32_t* height, uint64_t* nbags) {
  if (!g) return fail(KG_E_INVALID, "null object-grid handle");
  if (width) *width = g->width;
  if (height) *height = g->height;
  if (nbags) *nbags = g->ncells;
  return KG_OK;
}

}  // extern "C"
