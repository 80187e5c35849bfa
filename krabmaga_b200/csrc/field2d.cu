// Field2D on the device: cell-list rebuild (K1 histogram, K2 scan, K3 scatter), neighbour
// queries, the fused neighbour-gather + boids step (K4) and the C ABI over them.
//
// Replaces, for all agents at once (reference paths relative to the krABMaga crate root):
//   set_object_location      src/engine/fields/field_2d.rs:838-846   -> append + histogram
//   lazy_update              src/engine/fields/field_2d.rs:905-921   -> scan + scatter
//   get_neighbors_within_*   src/engine/fields/field_2d.rs:386-516   -> window walk over the
//                                                                       cell-sorted buffer
//   Bird::step               tests/model/flockers/bird.rs:39-155     -> step_boids kernels
//
// HBM layout (per handle): two agent buffers A (read: sorted by flat cell x*dh+y, i.e. the
// reference's iter_objects order) and B (write: append log), each {id: u32[cap], pv: float4[cap]}
// with pv = (pos.x, pos.y, last_d.x, last_d.y) so that one 128-bit load fetches everything a
// neighbour contributes; `cell_start[C+1]` offsets into A; `count[C]` histogram of B, which the
// scatter consumes back to zero (rank = atomicSub-1) so it never needs clearing.  The C ABI speaks
// SoA (five arrays); pack/unpack kernels convert through a staging area on upload/download.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "boids_device.cuh"
#include "common.cuh"
#include "jit.cuh"
#include "jit_agent.cuh"
#include "reduce.cuh"
#include "scan.cuh"

namespace kg {

// ------------------------------------------------------------------ SoA <-> packed conversion
__global__ void pack_kernel(uint64_t n, SoA s, Agents d, uint64_t d_off) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d.id[d_off + i] = s.id[i];
  d.pv[d_off + i] = make_float4(s.x[i], s.y[i], s.dx[i], s.dy[i]);
}
__global__ void unpack_kernel(Geom g, uint64_t n, Agents a, SoA d, int32_t* __restrict__ cell) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = a.pv[i];
  d.id[i] = a.id[i];
  d.x[i] = q.x;
  d.y[i] = q.y;
  d.dx[i] = q.z;
  d.dy[i] = q.w;
  if (cell) {
    uint32_t c;
    flat_cell(g, q.x, q.y, &c);
    cell[i] = (int32_t)c;
  }
}

// e2e upload in one pass: SoA staging -> packed log entry, histogram, validation flags
// (s.id == nullptr: agent i has id i)
__global__ void pack_hist_kernel(Geom g, uint64_t n, SoA s, Agents d, uint32_t* __restrict__ count, int* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = s.id ? s.id[i] : (uint32_t)i;
  const float4 q = make_float4(s.x[i], s.y[i], s.dx[i], s.dy[i]);
  d.id[i] = id;
  d.pv[i] = q;
  uint32_t c;
  if (flat_cell(g, q.x, q.y, &c))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
  if (id == kIdNone) atomicOr(err, DEV_ERR_SENTINEL);
}
// entries [first, first + n) of a packed buffer -> the same range of the SoA staging arrays
__global__ void unpack_range_kernel(uint64_t first, uint64_t n, Agents a, SoA d) {
  uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= first + n) return;
  const float4 q = a.pv[i];
  d.id[i] = a.id[i];
  d.x[i] = q.x;
  d.y[i] = q.y;
  d.dx[i] = q.z;
  d.dy[i] = q.w;
}

// entries [0, n) of the write log (entry k = the stepped agent of read slot k) -> SoA staging arrays at the
// agent's position in the host's INPUT arrays (origin[k]; ids do not travel back)
__global__ void unpack_ordered_kernel(uint64_t n, Agents a, const uint32_t* __restrict__ origin, SoA d) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const float4 q = a.pv[k];
  const uint32_t o = origin[k];
  d.x[o] = q.x;
  d.y[o] = q.y;
  d.dx[o] = q.z;
  d.dy[o] = q.w;
}

// ------------------------------------------------------------------ K1: histogram of new entries
__global__ void hist_kernel(Geom g, uint64_t first, uint64_t n, const float4* __restrict__ pv,
                            uint32_t* __restrict__ count, int* err) {
  uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= first + n) return;
  float4 q = pv[i];
  uint32_t c;
  if (flat_cell(g, q.x, q.y, &c))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}
// validation-only pass for uploads: raises the flag before anything is appended
__global__ void check_cells_kernel(Geom g, uint64_t n, const uint32_t* __restrict__ id, const float* __restrict__ x,
                                   const float* __restrict__ y, int* err) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c;
  if (!flat_cell(g, x[i], y[i], &c)) atomicOr(err, DEV_ERR_OOB);
  if (id[i] == kIdNone) atomicOr(err, DEV_ERR_SENTINEL);
}

// ------------------------------------------------------------------ K3: scatter log -> sorted
// Two log entries per thread, a block apart (coalesced), so each thread keeps two independent
// load -> atomic -> store chains in flight: the kernel is bound by that chain's latency.
#ifndef KG_SCATTER_ITEMS
#define KG_SCATTER_ITEMS 2
#endif
constexpr int kScatterItems = KG_SCATTER_ITEMS;
// ORIGIN (the ordered e2e entry only): origin[d] = log index of the entry that landed in sorted slot d
template <bool ORIGIN = false>
__global__ void __launch_bounds__(256)
scatter_kernel(Geom g, uint64_t n, Agents src, Agents dst, const uint32_t* __restrict__ cell_start,
               uint32_t* __restrict__ count, uint32_t* __restrict__ origin = nullptr) {
#if !KG_SCATTER_EARLY
  grid_dep_wait();  // cell_start comes from the scan launched just before
#endif
  const uint64_t i0 = (uint64_t)blockIdx.x * (blockDim.x * kScatterItems) + threadIdx.x;
  float4 q[kScatterItems];
  uint32_t id[kScatterItems], c[kScatterItems], rank[kScatterItems];
  bool ok[kScatterItems];
#pragma unroll
  for (int k = 0; k < kScatterItems; ++k) {
    const uint64_t i = i0 + (uint64_t)k * blockDim.x;
    ok[k] = i < n;
    if (ok[k]) {
      q[k] = src.pv[i];
      id[k] = src.id[i];
      ok[k] = id[k] != kIdNone;  // a stopped agent's entry (dynamic population): not in the histogram
    }
  }
#if KG_SCATTER_EARLY
  // The log was written before the scan started (the scan triggers its dependents only after its own wait),
  // so the loads above ran beside the scan; `count` (read by the scan) and cell_start (its output) need it done.
  grid_dep_wait();
#endif
  // (warp-aggregated rank allocation was measured on B200 twice and lost both times: with __match_any_sync
  // on the cell 16.8 vs 12.7 us at 1M agents, 94.5 vs 70.0 at 8M; with runs of ADJACENT equal cells found
  // by one shuffle + one ballot — the log is nearly sorted — 12.7 vs 12.6 and 72.7 vs 69.9: the atomics are
  // not what bounds this kernel.  profiles/r02_k4_experiments.txt)
#pragma unroll
  for (int k = 0; k < kScatterItems; ++k) {
    ok[k] = ok[k] && flat_cell(g, q[k].x, q[k].y, &c[k]);  // out-of-grid: already flagged by the histogram
    if (ok[k]) rank[k] = atomicSub(&count[c[k]], 1u) - 1u;
  }
#pragma unroll
  for (int k = 0; k < kScatterItems; ++k) {
    if (!ok[k]) continue;
    const uint32_t d = cell_start[c[k]] + rank[k];
    dst.id[d] = id[k];
    dst.pv[d] = q[k];
    if (ORIGIN) origin[d] = (uint32_t)(i0 + (uint64_t)k * blockDim.x);
  }
}

// State::after_step of the dynamic population: parent i of the READ buffer (iter_objects order) with
// birth[i] set pushes child number scan[i] into the log: id = next_id + scan[i], the parent's read
// position, last_d = 0.
__global__ void spawn_kernel(Geom g, uint32_t n, Agents rd, const uint32_t* __restrict__ birth,
                             const uint32_t* __restrict__ scan, uint32_t next_id, Agents wr_children,
                             uint32_t* __restrict__ count, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !birth[i]) return;
  const float4 parent = rd.pv[i];
  const uint32_t k = scan[i];
  wr_children.id[k] = next_id + k;
  wr_children.pv[k] = make_float4(parent.x, parent.y, 0.f, 0.f);
  uint32_t c;
  if (flat_cell(g, parent.x, parent.y, &c))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}
__global__ void count_stopped_kernel(uint32_t n, const uint32_t* __restrict__ ids, unsigned long long* out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool gone = i < n && ids[i] == kIdNone;
  const unsigned m = __ballot_sync(0xffffffffu, gone);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

// optional K3b: ascending-id order inside every bag (KG_ORDER_CANONICAL)
template <bool ORIGIN = false>
__global__ void sort_cells_kernel(uint32_t ncells, const uint32_t* __restrict__ cs, Agents a,
                                  uint32_t* __restrict__ origin = nullptr) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  uint32_t s = cs[c], e = cs[c + 1];
  for (uint32_t p = s + 1; p < e; ++p) {
    uint32_t id = a.id[p];
    if (a.id[p - 1] <= id) continue;
    float4 v = a.pv[p];
    uint32_t og = ORIGIN ? origin[p] : 0u;
    uint32_t q = p;
    while (q > s && a.id[q - 1] > id) {
      a.id[q] = a.id[q - 1];
      a.pv[q] = a.pv[q - 1];
      if (ORIGIN) origin[q] = origin[q - 1];
      --q;
    }
    a.id[q] = id;
    a.pv[q] = v;
    if (ORIGIN) origin[q] = og;
  }
}

__global__ void count_empty_kernel(uint32_t ncells, const uint32_t* __restrict__ cs,
                                   unsigned long long* out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  bool empty = c < ncells && cs[c] == cs[c + 1];
  unsigned m = __ballot_sync(0xffffffffu, empty);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

__global__ void cell_counts_from_starts_kernel(uint32_t ncells, const uint32_t* __restrict__ cs,
                                               uint32_t* __restrict__ out) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncells) out[c] = cs[c + 1] - cs[c];
}

__global__ void num_at_locations_kernel(Geom g, uint64_t nq, const float* __restrict__ x,
                                        const float* __restrict__ y,
                                        const uint32_t* __restrict__ cs, uint32_t* out, int* err) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t c;
  if (flat_cell(g, x[q], y[q], &c))
    out[q] = cs[c + 1] - cs[c];
  else {
    out[q] = 0;
    atomicOr(err, DEV_ERR_OOB);
  }
}

// remove_object_location on the write log: keep[i] = 0 for matching entries
__global__ void mark_remove_kernel(Geom g, uint64_t n, Agents a, uint32_t target,
                                   uint32_t target_cell, uint32_t* keep, uint32_t* count) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = a.pv[i];
  uint32_t c;
  bool ok = flat_cell(g, q.x, q.y, &c);
  bool drop = ok && c == target_cell && a.id[i] == target;
  keep[i] = drop ? 0u : 1u;
  if (drop) atomicSub(&count[c], 1u);
}
__global__ void compact_kernel(uint64_t n, const uint32_t* __restrict__ keep_scan, Agents src,
                               Agents dst) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t d = keep_scan[i];
  if (keep_scan[i + 1] == d) return;
  dst.id[d] = src.id[i];
  dst.pv[d] = src.pv[i];
}

// ------------------------------------------------------------------ queries (parity / debug)
template <bool EXACT>
__global__ void query_count_kernel(Geom g, uint64_t nq, const float* __restrict__ qx,
                                   const float* __restrict__ qy, float dist,
                                   const uint32_t* __restrict__ cs, const float4* __restrict__ pv,
                                   uint32_t* __restrict__ counts) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint32_t n = 0;
  for_each_neighbor<EXACT>(g, cs, pv, qx[q], qy[q], dist, [&](uint32_t) { ++n; });
  counts[q] = n;
}
template <bool EXACT>
__global__ void query_fill_kernel(Geom g, uint64_t nq, const float* __restrict__ qx,
                                  const float* __restrict__ qy, float dist,
                                  const uint32_t* __restrict__ cs, const float4* __restrict__ pv,
                                  const uint32_t* __restrict__ rid,
                                  const uint64_t* __restrict__ offsets, uint32_t* __restrict__ ids,
                                  float4* __restrict__ agents, uint64_t cap) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  uint64_t o = offsets[q];
  for_each_neighbor<EXACT>(g, cs, pv, qx[q], qy[q], dist, [&](uint32_t k) {
    if (o < cap) {
      ids[o] = rid[k];
      if (agents) agents[o] = pv[k];  // the neighbour itself (pos, last_d): the reference returns Vec<O>
    }
    ++o;
  });
}
__global__ void split_agents_kernel(uint64_t n, const float4* __restrict__ a, float* __restrict__ x,
                                    float* __restrict__ y, float* __restrict__ dx, float* __restrict__ dy) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = a[i];
  x[i] = v.x; y[i] = v.y; dx[i] = v.z; dy[i] = v.w;
}
__global__ void widen_offsets_kernel(uint64_t nq, const uint32_t* __restrict__ scan32,
                                     uint64_t* __restrict__ out) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q <= nq) out[q] = scan32[q];
}

// ------------------------------------------------------------------ K4: fused gather + boids
// (arithmetic in boids_device.cuh)
// generic K4: any geometry, both query kinds
// LIFE: dynamic population (krabgpu.h KgLifeRule) — a stopped agent leaves a kIdNone entry in the log
// and no histogram count; birth[i] says whether agent i leaves a child.
template <bool EXACT, bool LIFE = false>
__global__ void __launch_bounds__(128)
step_boids_kernel(Geom g, KgBoidsParams p, uint32_t n, Agents rd,
                  const uint32_t* __restrict__ cell_start, Agents wr, uint32_t* __restrict__ count,
                  int* err, KgLifeRule life = KgLifeRule{}, uint32_t* __restrict__ birth = nullptr,
                  uint32_t first = 0) {
  uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;  // agents [first, n) of the read buffer
  if (i >= n) return;
  uint32_t id = rd.id[i];
  float4 self = rd.pv[i];
  BoidsAcc acc;
  const uint32_t* __restrict__ rid = rd.id;
  const float4* __restrict__ rpv = rd.pv;
  for_each_neighbor<EXACT>(g, cell_start, rpv, self.x, self.y, p.radius, [&](uint32_t k) {
    boids_pair(acc, id, self.x, self.y, rid[k], rpv[k], g.w, g.h);
  });
  float4 out = boids_finish(acc, p, id, self.x, self.y, self.z, self.w, g.w);
  if (LIFE) {
    bool stopped, child;
    life_decide(life, p, id, acc.count, &stopped, &child);
    birth[i] = child ? 1u : 0u;
    if (stopped) {
      wr.id[i] = kIdNone;
      return;
    }
  }
  wr.id[i] = id;
  wr.pv[i] = out;
  uint32_t c;
  if (flat_cell(g, out.x, out.y, &c))
    atomicAdd(&count[c], 1u);  // K1 fused: histogram of the write log
  else
    atomicOr(err, DEV_ERR_OOB);
}

// Fast K4 for the north-star geometry class: toroidal field (clamped window, F3), relaxed query,
// and a window so small against the world that toroidal_distance always takes its first branch
// ((dd+1)*disc well below dim/2, checked on the host).  Per candidate: one 128-bit load + one id
// load and ~28 FP32/INT instructions, no branches besides the loop.
__global__ void __launch_bounds__(128)
step_boids_fast_kernel(Geom g, KgBoidsParams p, int dd, uint32_t n, Agents rd,
                       const uint32_t* __restrict__ cell_start, Agents wr,
                       uint32_t* __restrict__ count, int* err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t id = rd.id[i];
  const float4 self = rd.pv[i];
  const float px = self.x, py = self.y;
  int cx = f2i_sat(floorf(fdiv(px, g.disc)));
  int cy = f2i_sat(floorf(fdiv(py, g.disc)));
  int min_i = max(0, cx - dd), max_i = min(cx + dd, g.max_x - 1);
  int min_j = max(0, cy - dd), max_j = min(cy + dd, g.max_y - 1);
  // quotient-range guard for fdiv2_shared: an agent within 2^-20 of the origin axes could form a
  // denormal-scale dx; such threads take the compiler's full division instead
  const bool safe = px >= 9.5367431640625e-7f && py >= 9.5367431640625e-7f;
  BoidsAcc acc;
  const uint32_t* __restrict__ rid = rd.id;
  const float4* __restrict__ rpv = rd.pv;
  if (min_j <= max_j) {
    for (int ci = min_i; ci <= max_i; ++ci) {
      const uint32_t s = cell_start[ci * g.dh + min_j];
      const uint32_t e = cell_start[ci * g.dh + max_j + 1];
      acc.nvec += e - s;
      if (safe)
        boids_slice<true>(acc, id, px, py, rid, rpv, s, e);
      else
        boids_slice<false>(acc, id, px, py, rid, rpv, s, e);
    }
  }
  float4 out = boids_finish(acc, p, id, px, py, self.z, self.w, g.w);
  wr.id[i] = id;
  wr.pv[i] = out;
  uint32_t c;
  if (flat_cell(g, out.x, out.y, &c))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}

// Packed fast K4: step_boids_fast_kernel with the candidate loop on FADD2/FMUL2/FFMA2
// (boids_slice2) — about 17 issue slots per candidate instead of 31, same result bits.
// `ids_dup` is a device word written by verify_ids(): 0 = the ids of the read buffer were verified
// unique, so "candidate index == my index" is the reference's "elem.id == self.id" (bird.rs:63)
// and neither the id load nor a per-candidate counter is needed; non-zero = duplicates (or ids
// too large to verify) => compare ids like the reference does.  The branch is grid-uniform.
template <bool EXACT, bool LIFE = false>
__global__ void __launch_bounds__(128, (EXACT || LIFE) ? 6 : KG_K4_MINBLOCKS)
step_boids_packed_kernel(Geom g, KgBoidsParams p, int dd, float T, uint32_t n, Agents rd,
                         const uint32_t* __restrict__ cell_start, Agents wr,
                         uint32_t* __restrict__ count, const int* __restrict__ ids_dup, int* err,
                         KgLifeRule life = KgLifeRule{}, uint32_t* __restrict__ birth = nullptr,
                         uint32_t first = 0) {
  grid_dep_wait();  // the read buffer comes from the scatter launched just before
  uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;  // agents [first, n) of the read buffer
  if (i >= n) return;
  const uint32_t id = rd.id[i];
  const ulonglong2 self = reinterpret_cast<const ulonglong2*>(rd.pv)[i];
  int ncx, ncy, cnt = 0;
  const ulonglong2 out = boids_step_packed<EXACT>(g, p, dd, T, *ids_dup != 0, i, id, self, 0, cell_start,
                                                  rd.id, rd.pv, &ncx, &ncy, LIFE ? &cnt : nullptr);
  if (LIFE) {
    bool stopped, child;
    life_decide(life, p, id, cnt, &stopped, &child);
    birth[i] = child ? 1u : 0u;
    if (stopped) {
      wr.id[i] = kIdNone;
      return;
    }
  }
  wr.id[i] = id;
  reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
  const uint32_t c = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;  // field_2d.rs:840
  if ((int32_t)c >= 0 && c < g.ncells)
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}

// Tile K4 (the north-star's design): a block owns the agents of K consecutive cells of one cell ROW
// (fixed cy, x0 <= cx < x0 + K).  For the relaxed 3x3 query the candidates of column x are ONE
// contiguous, 16-byte aligned slice of the read buffer (cells cy-1..cy+1 of that column), so the
// block stages the K + 2 slices its agents can see with one cp.async.bulk each (TMA bulk copy engine,
// completion on an mbarrier; no thread spends an instruction per staged element) laid end to end in
// x order.  The window of an agent in column x (slices x-1, x, x+1) is then ONE contiguous range of
// shared memory in exactly the reference's order (x outer, y inner, bag order; field_2d.rs:502-512):
// one candidate loop per agent instead of three, candidates come from LDS.128 instead of L1/L2, and
// the agent's own cell is known from the tile instead of from a division.  All agents of a cell share
// their window, so the block sorts its <= K non-empty cells by window length (counting sort on 64
// bins) and deals agents to lanes in that order: the 32 loops of a warp have nearly equal trip
// counts.  While the copies fly the block does that sort.  Arithmetic, order of operations and
// results are those of step_boids_packed_kernel, bit for bit.
constexpr int kTileMaxCols = 128;    // K + 2 <= 128
constexpr int kTileThreads = 256;
constexpr int kTileStageCap = 1536;  // staged candidates (24 KB); a denser tile takes the global path
constexpr int kTileMaxOwn = 512;     // owner table; beyond it agents find their cell by search
constexpr int kTileTargetOwn = 236;  // host: agents per tile aimed at (256 threads, ~2 sigma headroom)

// exclusive scan, in place, of arr[0..128) by one warp (four entries per lane); returns the total
__device__ __forceinline__ uint32_t warp_scan128(uint32_t* arr, int lane) {
  const uint4 v = reinterpret_cast<const uint4*>(arr)[lane];
  const uint32_t s = v.x + v.y + v.z + v.w;
  const uint32_t inc = warp_incl_scan(s, lane);
  uint4 o;
  o.x = inc - s;
  o.y = o.x + v.x;
  o.z = o.y + v.y;
  o.w = o.z + v.z;
  reinterpret_cast<uint4*>(arr)[lane] = o;
  return __shfl_sync(0xffffffffu, inc, 31);
}

__global__ void __launch_bounds__(kTileThreads, 5)
step_boids_tile_kernel(Geom g, KgBoidsParams p, int K, int stage_mode, Agents rd,
                       const uint32_t* __restrict__ cell_start, Agents wr, uint32_t* __restrict__ count,
                       const int* __restrict__ ids_dup, int* err) {
  __shared__ __align__(16) ulonglong2 stage[kTileStageCap];
  __shared__ __align__(16) uint32_t col_off[kTileMaxCols + 4];   // offset of a column's slice in `stage`
  __shared__ __align__(16) uint32_t sown_off[kTileMaxCols + 4];  // prefix over owned agents, sorted cells
  __shared__ __align__(16) uint32_t bins[64];
  __shared__ uint32_t col_s[kTileMaxCols];    // global index of the first staged candidate of a column
  __shared__ uint32_t own_m0[kTileMaxCols];   // global index of the first owned agent of a column
  __shared__ uint8_t sorted_col[kTileMaxCols];
  __shared__ uint8_t owner[kTileMaxOwn];      // owned agent (in sorted order) -> position of its cell
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int cy = blockIdx.x;
  const int x0 = blockIdx.y * K;
  const int ncols = K + 2;
  const int min_j = max(0, cy - 1), max_j = min(cy + 1, g.max_y - 1);
  const bool rows_ok = min_j <= max_j;
  if (tid == 0) mbar_init(&bar, 1);
  if (tid < 64) bins[tid] = 0;
  grid_dep_wait();  // the read buffer and cell_start come from the rebuild launched just before
  // ---- phase 1: slice and owned range of every column (thread t <-> column x0 - 1 + t)
  uint32_t len = 0, nown = 0, src0 = 0;
  if (tid < kTileMaxCols + 4) {
    uint32_t m0 = 0;
    const int x = x0 - 1 + tid;
    if (tid < ncols && x >= 0 && x < g.dw) {
      const uint32_t base = (uint32_t)x * (uint32_t)g.dh;
      if (rows_ok && x < g.max_x) {  // the padding column is never scanned (F4)
        src0 = cell_start[base + min_j];
        len = cell_start[base + max_j + 1] - src0;
      }
      if (tid >= 1 && tid <= K) {
        m0 = cell_start[base + cy];
        nown = cell_start[base + cy + 1] - m0;
      }
    }
    if (tid < kTileMaxCols) {
      col_s[tid] = src0;
      own_m0[tid] = m0;
    }
    col_off[tid] = len;
    sown_off[tid] = 0;
  }
  if (!__syncthreads_or(nown != 0)) return;  // nobody lives in this tile
  if (wid == 0) {
    const uint32_t total = warp_scan128(col_off, lane);
    if (lane == 0) col_off[kTileMaxCols] = total;
  }
  __syncthreads();
  const uint32_t stage_total = col_off[kTileMaxCols];
  const bool by_id = *ids_dup != 0;
  const bool staged = stage_total != 0 && stage_total <= (uint32_t)kTileStageCap && !by_id;
  // ---- phase 2: one bulk copy per column slice; the sort below runs while they are in flight
  if (staged && stage_mode == 0) {
    if (tid == 0) mbar_arrive_expect_tx(&bar, stage_total * 16u);
    if (len != 0)
      bulk_copy_g2s(&stage[col_off[tid]], reinterpret_cast<const ulonglong2*>(rd.pv) + src0, len * 16u, &bar);
  } else if (staged) {
    // lab variant (KG_TILE_STAGE=1): a warp copies a column slice with one LDG.128 / STS.128 per lane
    const ulonglong2* __restrict__ src = reinterpret_cast<const ulonglong2*>(rd.pv);
    for (int t = wid; t < ncols; t += kTileThreads / 32) {
      const uint32_t o = col_off[t], l = col_off[t + 1] - o, s0 = col_s[t];
      for (uint32_t j = lane; j < l; j += 32) stage[o + j] = src[s0 + j];
    }
  }
  // ---- phase 3: non-empty owned cells sorted by window length, longest first
  const bool has = nown != 0;
  uint32_t key = 0, slot = 0;
  if (has) {
    const uint32_t wlen = col_off[tid + 2] - col_off[tid - 1];
    key = 63u - min(63u, wlen);
    slot = atomicAdd(&bins[key], 1u);
  }
  __syncthreads();
  if (wid == 0) {
    const uint2 v = reinterpret_cast<const uint2*>(bins)[lane];
    const uint32_t s = v.x + v.y;
    const uint32_t ex = warp_incl_scan(s, lane) - s;
    reinterpret_cast<uint2*>(bins)[lane] = make_uint2(ex, ex + v.x);
  }
  __syncthreads();
  if (has) {
    const uint32_t pos = bins[key] + slot;
    sorted_col[pos] = (uint8_t)tid;
    sown_off[pos] = nown;
  }
  __syncthreads();
  if (wid == 0) {
    const uint32_t total = warp_scan128(sown_off, lane);
    if (lane == 0) sown_off[kTileMaxCols] = total;
  }
  __syncthreads();
  const uint32_t own_total = sown_off[kTileMaxCols];
  if (tid < kTileMaxCols) {
    const uint32_t b = sown_off[tid], e = min(sown_off[tid + 1], (uint32_t)kTileMaxOwn);
    for (uint32_t a = b; a < e; ++a) owner[a] = (uint8_t)tid;
  }
  __syncthreads();
  if (staged && stage_mode == 0) mbar_wait(&bar, 0);
  // ---- phase 4: the agents, one per lane, in sorted-cell order
  const Recip rdisc = recip_of(g.disc);
  for (uint32_t a = tid; a < own_total; a += kTileThreads) {
    uint32_t pos;
    if (a < (uint32_t)kTileMaxOwn) {
      pos = owner[a];
    } else {  // crowded tile: last position whose prefix is <= a
      int l = 0, r = kTileMaxCols - 1;
      while (l < r) {
        const int m = (l + r + 1) >> 1;
        if (sown_off[m] <= a) l = m; else r = m - 1;
      }
      pos = (uint32_t)l;
    }
    const int t = sorted_col[pos];
    const uint32_t kk = a - sown_off[pos];
    const uint32_t i = own_m0[t] + kk;  // index in the read buffer
    const uint32_t id = rd.id[i];
    int ncx, ncy;
    ulonglong2 out, self;
    bool fast_lane = staged;
    uint32_t self_j = 0x80000000u;  // my own slot in `stage` (none: padding row / column)
    if (staged) {
      const int x = x0 - 1 + t;
      const bool self_in = rows_ok && x < g.max_x && cy < g.max_y;
      if (self_in) self_j = col_off[t] + (own_m0[t] - col_s[t]) + kk;
      self = self_in ? stage[self_j] : reinterpret_cast<const ulonglong2*>(rd.pv)[i];
      float px, py;
      unpack2(self.x, &px, &py);
      fast_lane = px >= 9.5367431640625e-7f && py >= 9.5367431640625e-7f;  // fdiv2_shared's domain
    }
    if (fast_lane) {
      BoidsAcc2 a2;
      const uint32_t lo = col_off[t - 1], hi = col_off[t + 2];
      const ulonglong2* __restrict__ pc = stage + lo;
      uint32_t left = hi - lo;
      uint32_t rel = self_j - lo;
#pragma unroll 1
      for (; left >= 4; left -= 4, rel -= 4, pc += 4) {
        const ulonglong2 c0 = pc[0], c1 = pc[1], c2 = pc[2], c3 = pc[3];
        boids_pair2<1, 0>(a2, self.x, c0, rel, 0u, 0u);
        boids_pair2<1, 1>(a2, self.x, c1, rel, 0u, 0u);
        boids_pair2<1, 2>(a2, self.x, c2, rel, 0u, 0u);
        boids_pair2<1, 3>(a2, self.x, c3, rel, 0u, 0u);
      }
      if (left & 2u) {
        const ulonglong2 c0 = pc[0], c1 = pc[1];
        boids_pair2<1, 0>(a2, self.x, c0, rel, 0u, 0u);
        boids_pair2<1, 1>(a2, self.x, c1, rel, 0u, 0u);
        rel -= 2;
        pc += 2;
      }
      if (left & 1u) boids_pair2<1, 0>(a2, self.x, pc[0], rel, 0u, 0u);
      const uint32_t nvec = hi - lo;
      const int cnt = (int)(nvec - (self_j != 0x80000000u ? 1u : 0u));
      boids_finish_packed(a2.a, a2.c, a2.s, cnt, nvec, p, id, self.x, self.y, g.w, &out.x, &out.y);
      cell_of2(out.x, rdisc, &ncx, &ncy);
    } else {
      // dense tile, unverified ids or a near-origin agent: the per-agent path on the global arrays
      self = reinterpret_cast<const ulonglong2*>(rd.pv)[i];
      out = boids_step_packed<false>(g, p, 1, 0.0f, by_id, i, id, self, 0, cell_start, rd.id, rd.pv, &ncx,
                                     &ncy);
    }
    wr.id[i] = id;
    reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
    const uint32_t nc = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
    if ((int32_t)nc >= 0 && nc < g.ncells)
      atomicAdd(&count[nc], 1u);
    else
      atomicOr(err, DEV_ERR_OOB);
  }
}

// Column-chunk K4 (KG_K4_COLTILE; KG_ORDER_ANY only).  A block owns a chunk of <= kCtAgents CONSECUTIVE
// agents of one cell column x of the sorted read buffer (chunks of a column are equal, so no lane is
// lost to a ragged tile): its own loads, id loads and log stores are coalesced.  The candidates its
// agents can see are rows rs..re of columns x-1, x, x+1 — three contiguous slices of the read buffer —
// and the block stages them into shared memory ROW-major: [row rs: col x-1 | col x | col x+1][row rs+1:
// ...].  The destination of a staged entry follows from cell_start alone (no scan): rowoff[k] =
// sum over the three columns of cell_start[c][rs+k] - cell_start[c][rs].  In that layout the 3x3 window
// of an agent in row r is ONE contiguous range, rows r-1..r+1: one candidate loop per agent (the packed
// kernel runs three, each to its warp's longest lane), fed by LDS.128.  The chunk's agents are then
// counting-sorted by window length and dealt to lanes in that order, so the 32 loops of a warp have
// nearly equal trip counts (model: 26.3 candidate slots per lane against 33.4, tools/k4_lane_model.py).
// The window is walked y-outer / x-inner, not in field_2d.rs:502-512's x-outer order: the candidate SET
// and every per-pair operation are the reference's, the order of the f32 sums is not — which
// KG_ORDER_ANY leaves open anyway (bag order there is whatever the scatter's atomics produced).
// KG_ORDER_CANONICAL keeps the per-agent kernel.
constexpr int kCtThreads = 128;
constexpr int kCtPasses = 2;
constexpr int kCtAgents = kCtThreads * kCtPasses;  // own agents per chunk
constexpr int kCtRows = 256;                       // staged rows (own rows + 2)
constexpr int kCtStageCap = 1280;                  // staged candidates (20 KB); denser chunks take the global path

template <int STAGE>  // 0: one cp.async.bulk per (row, column) cell; 1: LDG.128 / STS.128 with the row from the position
__global__ void __launch_bounds__(kCtThreads, 8)
step_boids_coltile_kernel(Geom g, KgBoidsParams p, Agents rd, const uint32_t* __restrict__ cell_start,
                          Agents wr, uint32_t* __restrict__ count, const int* __restrict__ ids_dup,
                          int* err) {
  __shared__ __align__(16) ulonglong2 stage[kCtStageCap];
  __shared__ uint32_t rowoff[kCtRows + 4];    // first staged entry of row rs + k; [nrows] = total
  __shared__ uint32_t delta[3][kCtRows + 4];  // staged slot of (column ci, row k) entry = delta + its global index
  __shared__ uint32_t perm[kCtAgents];        // sorted position -> own agent (low 16 bits) and its row k (high)
  __shared__ uint32_t bins[64];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int x = blockIdx.x;
  grid_dep_wait();  // the read buffer and cell_start come from the rebuild launched just before
  const uint32_t colbase = (uint32_t)x * (uint32_t)g.dh;
  const uint32_t c0 = cell_start[colbase], c1 = cell_start[colbase + (uint32_t)g.dh];
  const uint32_t n_x = c1 - c0;
  if (n_x == 0) return;
  const uint32_t nch = (n_x + kCtAgents - 1) / kCtAgents;
  const uint32_t chunk = n_x / nch, extra = n_x - chunk * nch;  // the first `extra` chunks hold one agent more
  const bool by_id = *ids_dup != 0;
  const Recip rdisc = recip_of(g.disc);
  const ulonglong2* __restrict__ pv = reinterpret_cast<const ulonglong2*>(rd.pv);
  const int cmin = max(x - 1, 0), cmax = min(x + 1, g.max_x - 1);
  const int ncol = cmax - cmin + 1;
  uint32_t phase = 0;
  if (STAGE == 0) {
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
  }
  for (uint32_t j = blockIdx.y; j < nch; j += gridDim.y) {
    const uint32_t a0 = c0 + j * chunk + min(j, extra);
    const uint32_t n = chunk + (j < extra ? 1u : 0u);  // 1 .. kCtAgents
    const uint32_t a1 = a0 + n;
    // rows of the first and the last own agent (the buffer is sorted by cell: they bound the chunk)
    int r0, r1, tmp;
    cell_of2(pv[a0].x, rdisc, &tmp, &r0);
    cell_of2(pv[a1 - 1].x, rdisc, &tmp, &r1);
    const int rs = max(r0 - 1, 0), re = min(r1 + 1, g.max_y - 1);
    const int nrows = re - rs + 1;
    bool tiled = !by_id && x < g.max_x && r0 <= r1 && r0 >= 0 && nrows >= 1 && nrows <= kCtRows;
    uint32_t cb[3] = {0, 0, 0}, ln[3] = {0, 0, 0}, total = 0;
    if (tiled) {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ci < ncol) {
          const uint32_t base = (uint32_t)(cmin + ci) * (uint32_t)g.dh;
          cb[ci] = cell_start[base + rs];
          ln[ci] = cell_start[base + re + 1] - cb[ci];
          total += ln[ci];
        }
      tiled = total <= (uint32_t)kCtStageCap;
    }
    if (!tiled) {
      // unverified ids, the padding column, a sparse or a crowded chunk: the per-agent path on the global arrays
      for (uint32_t idx = tid; idx < n; idx += kCtThreads) {
        const uint32_t i = a0 + idx;
        const uint32_t id = rd.id[i];
        int ncx, ncy;
        const ulonglong2 out = boids_step_packed<false>(g, p, 1, 0.0f, by_id, i, id, pv[i], 0, cell_start, rd.id,
                                                        rd.pv, &ncx, &ncy);
        wr.id[i] = id;
        reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
        const uint32_t nc = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
        if ((int32_t)nc >= 0 && nc < g.ncells)
          atomicAdd(&count[nc], 1u);
        else
          atomicOr(err, DEV_ERR_OOB);
      }
      continue;
    }
    // ---- tables: where every (row, column) cell of the window region lands in `stage`
    if (tid < 64) bins[tid] = 0;
    if (STAGE == 0 && tid == 0) mbar_arrive_expect_tx(&bar, total * 16u);
    for (int k = tid; k <= nrows; k += kCtThreads) {
      uint32_t s[3] = {0, 0, 0}, l[3] = {0, 0, 0}, off = 0;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ci < ncol) {
          const uint32_t at = (uint32_t)(cmin + ci) * (uint32_t)g.dh + (uint32_t)(rs + k);
          s[ci] = cell_start[at];
          if (k < nrows) l[ci] = cell_start[at + 1] - s[ci];
          off += s[ci] - cb[ci];
        }
      rowoff[k] = off;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci) {
        delta[ci][k] = off - s[ci];
        // staging by the copy engine: the cell's entries are one contiguous 16-byte aligned segment
        if (STAGE == 0 && l[ci] != 0) bulk_copy_g2s(&stage[off], pv + s[ci], l[ci] * 16u, &bar);
        off += l[ci];
      }
    }
    __syncthreads();
    // ---- staging (lab variant): the three column slices, coalesced loads, each entry to its row-major slot
    if (STAGE == 1)
    for (uint32_t e = tid; e < total; e += kCtThreads) {
      const int ci = e < ln[0] ? 0 : (e < ln[0] + ln[1] ? 1 : 2);
      const uint32_t gi = ci == 0 ? cb[0] + e : (ci == 1 ? cb[1] + (e - ln[0]) : cb[2] + (e - ln[0] - ln[1]));
      const ulonglong2 v = pv[gi];
      int vx, vy;
      cell_of2(v.x, rdisc, &vx, &vy);
      const int k = min(max(vy - rs, 0), nrows - 1);
      const uint32_t d = delta[ci][k] + gi;
      if (d < total) stage[d] = v;
    }
    // ---- own agents: row, window length -> counting sort, longest window first
    uint32_t info[kCtPasses], key[kCtPasses], slot[kCtPasses];
#pragma unroll
    for (int q = 0; q < kCtPasses; ++q) {
      const uint32_t idx = tid + q * kCtThreads;
      key[q] = 0;
      slot[q] = 0;
      info[q] = 0xFFFFFFFFu;
      if (idx < n) {
        int ax, r;
        cell_of2(pv[a0 + idx].x, rdisc, &ax, &r);
        r = min(max(r, r0), r1);
        const int klo = max(r - 1, 0) - rs, khi = min(r + 1, g.max_y - 1) - rs;
        const uint32_t wl = rowoff[khi + 1] - rowoff[klo];
        key[q] = 63u - min(63u, wl);
        slot[q] = atomicAdd(&bins[key[q]], 1u);
        info[q] = idx | ((uint32_t)(r - rs + 1) << 16);  // row stored relative to rs - 1 (the padding row sits above re)
      }
    }
    __syncthreads();
    {
      const uint2 bv = reinterpret_cast<const uint2*>(bins)[lane];
      const uint32_t s = bv.x + bv.y;
      const uint32_t ex = warp_incl_scan(s, lane) - s;
#pragma unroll
      for (int q = 0; q < kCtPasses; ++q) {
        const uint32_t b0 = __shfl_sync(0xffffffffu, ex, key[q] >> 1);
        const uint32_t b1 = __shfl_sync(0xffffffffu, bv.x, key[q] >> 1);
        if (info[q] != 0xFFFFFFFFu) perm[b0 + ((key[q] & 1u) ? b1 : 0u) + slot[q]] = info[q];
      }
    }
    __syncthreads();
    if (STAGE == 0) {
      mbar_wait(&bar, phase);
      phase ^= 1u;
    }
    // ---- the agents, in sorted order; odd passes run backwards so that every warp gets long and short windows
#pragma unroll 1
    for (int q = 0; q < kCtPasses; ++q) {
      const uint32_t sp = (q & 1) ? (uint32_t)((q + 1) * kCtThreads - 1 - tid) : (uint32_t)(q * kCtThreads + tid);
      if (sp >= n) continue;
      const uint32_t pi = perm[sp];
      const uint32_t i = a0 + (pi & 0xFFFFu);
      const int r = rs - 1 + (int)(pi >> 16);
      const uint32_t id = rd.id[i];
      const ulonglong2 self = pv[i];
      float px, py;
      unpack2(self.x, &px, &py);
      ulonglong2 out;
      int ncx, ncy;
      if (px >= 9.5367431640625e-7f && py >= 9.5367431640625e-7f) {  // fdiv2_shared's domain
        const int klo = max(r - 1, 0) - rs, khi = min(r + 1, g.max_y - 1) - rs;
        const uint32_t lo = rowoff[klo], hi = rowoff[khi + 1];
        const uint32_t self_j = r <= re ? delta[x - cmin][r - rs] + i : 0x80000000u;  // none: padding row
        BoidsAcc2 a2;
        const ulonglong2* __restrict__ pc = stage + lo;
        uint32_t left = hi - lo;
        float relf = (float)(int)(self_j - lo);  // my own slot, counted from the loop's position (exact: < 2^24)
#pragma unroll 1
        for (; left >= 4; left -= 4, relf -= 4.0f, pc += 4) {
          const ulonglong2 q0 = pc[0], q1 = pc[1], q2 = pc[2], q3 = pc[3];
          const uint32_t rb = __float_as_uint(relf);
          boids_pair2<3, 0>(a2, self.x, q0, rb, 0u, 0u);
          boids_pair2<3, 1>(a2, self.x, q1, rb, 0u, 0u);
          boids_pair2<3, 2>(a2, self.x, q2, rb, 0u, 0u);
          boids_pair2<3, 3>(a2, self.x, q3, rb, 0u, 0u);
        }
        if (left & 2u) {
          const ulonglong2 q0 = pc[0], q1 = pc[1];
          const uint32_t rb = __float_as_uint(relf);
          boids_pair2<3, 0>(a2, self.x, q0, rb, 0u, 0u);
          boids_pair2<3, 1>(a2, self.x, q1, rb, 0u, 0u);
          relf -= 2.0f;
          pc += 2;
        }
        if (left & 1u) boids_pair2<3, 0>(a2, self.x, pc[0], __float_as_uint(relf), 0u, 0u);
        const uint32_t nvec = hi - lo;
        const int cnt = (int)(nvec - (self_j != 0x80000000u ? 1u : 0u));
        boids_finish_packed(a2.a, a2.c, a2.s, cnt, nvec, p, id, self.x, self.y, g.w, &out.x, &out.y);
        cell_of2(out.x, rdisc, &ncx, &ncy);
      } else {
        out = boids_step_packed<false>(g, p, 1, 0.0f, false, i, id, self, 0, cell_start, rd.id, rd.pv, &ncx, &ncy);
      }
      wr.id[i] = id;
      reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
      const uint32_t nc = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
      if ((int32_t)nc >= 0 && nc < g.ncells)
        atomicAdd(&count[nc], 1u);
      else
        atomicOr(err, DEV_ERR_OOB);
    }
    __syncthreads();  // the tables and the stage are rebuilt by the next chunk
  }
}

// Staged K4 (KG_K4_STAGED): the packed kernel's arithmetic, loops and order — bit-identical results —
// with the candidates read from shared memory.  A block owns a chunk of <= 128 consecutive agents of one
// cell column x (chunks of a column are equal).  Everything those agents can see is rows rs..re of
// columns x-1, x, x+1: THREE contiguous, 16-byte aligned slices of the read buffer, which ONE elected
// thread moves with three cp.async.bulk copies (TMA; ~2 KB each, completion on an mbarrier) while every
// thread fetches its own entry and its six slice bounds.  A bulk copy is a warp-uniform instruction
// (UBLKCP: per-lane copies are serialised by an ELECT loop, ~9 instructions per copy and lane — measured
// on the per-cell variant of the column-chunk kernel), so a few large copies are what the engine is for.
// The three candidate loops then run on LDS.128: no L1/L2 latency inside the loops.
constexpr int kSgThreads = 128;
constexpr int kSgStageCap = 640;  // staged candidates (10 KB); a denser chunk takes the global path

__global__ void __launch_bounds__(kSgThreads, 9)
step_boids_staged_kernel(Geom g, KgBoidsParams p, Agents rd, const uint32_t* __restrict__ cell_start,
                         Agents wr, uint32_t* __restrict__ count, const int* __restrict__ ids_dup, int* err) {
  __shared__ __align__(16) ulonglong2 stage[kSgStageCap];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const int x = blockIdx.x;
  if (tid == 0) mbar_init(&bar, 1);
  grid_dep_wait();  // the read buffer and cell_start come from the rebuild launched just before
  const uint32_t colbase = (uint32_t)x * (uint32_t)g.dh;
  const uint32_t c0 = cell_start[colbase], c1 = cell_start[colbase + (uint32_t)g.dh];
  const uint32_t n_x = c1 - c0;
  if (n_x == 0) return;
  const uint32_t nch = (n_x + kSgThreads - 1) / kSgThreads;
  const uint32_t chunk = n_x / nch, extra = n_x - chunk * nch;  // the first `extra` chunks hold one agent more
  const bool by_id = *ids_dup != 0;
  const Recip rdisc = recip_of(g.disc);
  const ulonglong2* __restrict__ pv = reinterpret_cast<const ulonglong2*>(rd.pv);
  const int cmin = max(x - 1, 0), cmax = min(x + 1, g.max_x - 1);
  const int ncol = cmax - cmin + 1;
  uint32_t phase = 0;
  __syncthreads();  // the mbarrier is initialised
  for (uint32_t j = blockIdx.y; j < nch; j += gridDim.y) {
    const uint32_t a0 = c0 + j * chunk + min(j, extra);
    const uint32_t n = chunk + (j < extra ? 1u : 0u);  // 1 .. kSgThreads
    int r0, r1, tmp;
    cell_of2(pv[a0].x, rdisc, &tmp, &r0);  // the buffer is sorted by cell: first and last agent bound the rows
    cell_of2(pv[a0 + n - 1].x, rdisc, &tmp, &r1);
    const int rs = max(r0 - 1, 0), re = min(r1 + 1, g.max_y - 1);
    bool staged = !by_id && x < g.max_x && r0 <= r1 && r0 >= 0 && rs <= re;
    uint32_t cb[3] = {0, 0, 0}, so[3] = {0, 0, 0}, ln[3] = {0, 0, 0}, total = 0;
    if (staged) {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ci < ncol) {
          const uint32_t base = (uint32_t)(cmin + ci) * (uint32_t)g.dh;
          cb[ci] = cell_start[base + rs];
          so[ci] = total;
          ln[ci] = cell_start[base + re + 1] - cb[ci];
          total += ln[ci];
        }
      staged = total <= (uint32_t)kSgStageCap;
    }
    if (staged && tid == 0) {
      mbar_arrive_expect_tx(&bar, total * 16u);
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ln[ci] != 0) bulk_copy_g2s(&stage[so[ci]], pv + cb[ci], ln[ci] * 16u, &bar);
    }
    const bool mine = (uint32_t)tid < n;
    const uint32_t i = a0 + (mine ? (uint32_t)tid : 0u);
    const uint32_t id = rd.id[i];
    const ulonglong2 self = pv[i];
    float px, py;
    unpack2(self.x, &px, &py);
    int cxs, cy;
    cell_of2(self.x, rdisc, &cxs, &cy);
    const int min_j = max(0, cy - 1), max_j = min(cy + 1, g.max_y - 1);
    const bool fast = staged && px >= 9.5367431640625e-7f && py >= 9.5367431640625e-7f;  // fdiv2_shared's domain
    uint32_t sb[3] = {0, 0, 0}, eb[3] = {0, 0, 0};
    if (fast && min_j <= max_j) {
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        if (ci < ncol) {
          const uint32_t base = (uint32_t)(cmin + ci) * (uint32_t)g.dh;
          sb[ci] = cell_start[base + min_j];
          eb[ci] = cell_start[base + max_j + 1];
        }
    }
    if (staged) {
      mbar_wait(&bar, phase);
      phase ^= 1u;
    }
    if (mine) {
      ulonglong2 out;
      int ncx, ncy;
      if (fast) {
        BoidsAcc2 a2;
        uint32_t nvec = 0, self_hits = 0;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          if (ci >= ncol) break;
          const uint32_t s0 = sb[ci], e0 = eb[ci];
          nvec += e0 - s0;
          // index k of the read buffer sits at stage[so + k - cb]
          const uint32_t sh = so[ci] - cb[ci], ss = s0 + sh, se = e0 + sh, sk = i + sh;
          if (i - s0 < e0 - s0) {  // my own column: leave myself out of the consistency sum
            self_hits += 1;
            boids_slice2<1>(a2, sk, id, self.x, rd.id, stage, ss, se);
          } else {
            boids_slice2<0>(a2, sk, id, self.x, rd.id, stage, ss, se);
          }
        }
        boids_finish_packed(a2.a, a2.c, a2.s, (int)(nvec - self_hits), nvec, p, id, self.x, self.y, g.w, &out.x,
                            &out.y);
        cell_of2(out.x, rdisc, &ncx, &ncy);
      } else {
        out = boids_step_packed<false>(g, p, 1, 0.0f, by_id, i, id, self, 0, cell_start, rd.id, rd.pv, &ncx, &ncy);
      }
      wr.id[i] = id;
      reinterpret_cast<ulonglong2*>(wr.pv)[i] = out;
      const uint32_t nc = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
      if ((int32_t)nc >= 0 && nc < g.ncells)
        atomicAdd(&count[nc], 1u);
      else
        atomicOr(err, DEV_ERR_OOB);
    }
    if (j + gridDim.y < nch) __syncthreads();  // the next chunk's copies overwrite the stage
  }
}

// host: cells per tile so that a tile holds about kTileTargetOwn agents, tiles of one row equal
inline int tile_cells_for(const Geom& g, uint64_t n) {
  const double cells = (double)g.max_x * (double)g.max_y;
  const double rho = cells > 0 ? (double)n / cells : 1.0;
  int target = kTileTargetOwn;
  if (const char* e = getenv("KG_TILE_TARGET")) target = std::max(16, atoi(e));  // lab hook (tools/k4_ab.py)
  int k = rho > 0 ? (int)(target / rho) : kTileMaxCols - 2;
  k = std::max(4, std::min(k, kTileMaxCols - 2));
  const int ntx = (g.dw + k - 1) / k;
  return (g.dw + ntx - 1) / ntx;
}

// self-test of fdiv2_shared against __fdiv_rn over the domain the fast kernel feeds it
__global__ void selftest_div_kernel(uint64_t n, uint64_t seed, unsigned long long* mismatches) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad = 0;
  for (; i < n; i += stride) {
    Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 7, 99, (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    // numerators: sign * 2^[-60,20) * [1,2);  denominators: sq*sq+1 with sq in 2^[-30,20)
    int ea = (int)(r.v[0] % 80u) - 60;
    float a0 = ldexpf(1.0f + u01_f32(r.v[1]), ea) * ((r.v[0] & 0x80000000u) ? -1.f : 1.f);
    int eb = (int)(r.v[2] % 80u) - 60;
    float a1 = ldexpf(1.0f + u01_f32(r.v[3]), eb) * ((r.v[2] & 0x80000000u) ? -1.f : 1.f);
    Philox4 r2 = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 8, 99, (uint32_t)seed,
                               (uint32_t)(seed >> 32));
    int es = (int)(r2.v[0] % 50u) - 30;
    float sq = ldexpf(1.0f + u01_f32(r2.v[1]), es);
    if ((r2.v[2] & 7u) == 0) {  // the kernel's own construction: den from dx, dy
      sq = fadd(fmul(a0, a0), fmul(a1, a1));
      if (!(sq < 1.0e6f)) sq = 1.0e6f;
    }
    if ((r2.v[2] & 0xF0u) == 0) a0 = 0.0f;
    float den = fadd(fmul(sq, sq), 1.0f);
    float q0, q1;
    fdiv2_shared(a0, a1, den, &q0, &q1);
    bad += (__float_as_uint(q0) != __float_as_uint(fdiv(a0, den)));
    bad += (__float_as_uint(q1) != __float_as_uint(fdiv(a1, den)));
  }
  for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

// State::init of the fixture (state.rs:41-56) with Philox draws
__global__ void init_flockers_kernel(Geom g, uint64_t first, uint64_t n, uint64_t seed, Agents wr,
                                     uint32_t* __restrict__ count, int* err) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t id = (uint32_t)t;
  Philox4 r = philox4x32_10(id, 0, 0, DOMAIN_INIT, (uint32_t)seed, (uint32_t)(seed >> 32));
  float x = fmul(g.w, u01_f32(r.v[0])), y = fmul(g.h, u01_f32(r.v[1]));
  uint64_t i = first + t;
  wr.id[i] = id;
  wr.pv[i] = make_float4(x, y, 0.f, 0.f);
  uint32_t c;
  if (flat_cell(g, x, y, &c))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(err, DEV_ERR_OOB);
}

}  // namespace kg

// ====================================================================== handle + C ABI
using namespace kg;

struct kg_field2d {
  int device = 0;
  cudaStream_t stream = nullptr;
  Geom g{};
  uint64_t capacity = 0;
  Agents A, B;  // A = read (cell-sorted), B = write (append log)
  uint64_t n_read = 0, n_write = 0;
  uint32_t* cell_start = nullptr;  // [ncells+1] offsets into A
  uint32_t* count = nullptr;       // [ncells] histogram of B (zero between rebuilds)
  LookbackState scan;
  uint32_t* tile_sums = nullptr;   // scratch for the 3-pass scan (queries / remove)
  uint32_t* scratch = nullptr;     // u32 scratch (keep flags / counts)
  uint64_t scratch_len = 0;
  SoA stage;                       // SoA staging for the ABI, allocated on first use
  int32_t* stage_cell = nullptr;
  bool have_stage = false;
  int* d_err = nullptr;
  int* h_err = nullptr;  // pinned mirror
  uint64_t nagents = 0;
  bool density_estimation_check = false;
  int order = KG_ORDER_ANY;
  int variant = KG_K4_AUTO;
  // id uniqueness of the read buffer (selects the self-exclusion test of the packed K4)
  uint32_t* id_bitmap = nullptr;
  uint64_t id_bitmap_bits = 0;
  int* d_ids_dup = nullptr;
  float* stage4 = nullptr;             // ordered e2e entry: x | y | dx | dy of n agents back to back (one copy each way)
  uint32_t* origin = nullptr;          // ordered e2e entry: input index of the agent in every sorted slot
  bool want_origin = false;            // the next rebuild() records it
  cudaStream_t copy_stream = nullptr;  // e2e entry: downloads of finished slabs run beside the next slab's K4
  cudaEvent_t slab_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t copy_done = nullptr;
  void* qbuf = nullptr;         // persistent scratch of the neighbour queries (query- and result-sized parts)
  size_t qbuf_bytes = 0;
  void* rbuf = nullptr;
  size_t rbuf_bytes = 0;
  double* red = nullptr;        // scratch of kg_field2d_reduce
  size_t red_bytes = 0;
  Agents remove_tmp;            // compaction target of remove_object_location, allocated on first use
  uint32_t next_id = 0;         // dynamic population: id of the next child
  bool log_has_holes = false;   // the write log holds kIdNone entries: the rebuild must count survivors
  bool ids_unknown = true;      // read buffer not verified since its ids last changed
  bool pending_new_ids = false; // the write log holds entries that did not come from K4
  Profiler prof;
  Stopwatch watch;
  L2Flusher flusher;
  EventPool events;
};

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(uint64_t n, int threads = kThreads) {
  return (unsigned)((n + threads - 1) / threads);
}

int alloc_agents(Agents& a, uint64_t cap) {
  uint64_t n = cap + 64;
  KG_CUDA(cudaMalloc(&a.id, n * 4));
  KG_CUDA(cudaMalloc(&a.pv, n * 16));
  return KG_OK;
}
void free_agents(Agents& a) {
  cudaFree(a.id);
  cudaFree(a.pv);
  a = Agents{};
}
int ensure_stage(kg_field2d* f) {
  if (f->have_stage) return KG_OK;
  uint64_t n = f->capacity + 64;
  KG_CUDA(cudaMalloc(&f->stage.id, n * 4));
  KG_CUDA(cudaMalloc(&f->stage.x, n * 4));
  KG_CUDA(cudaMalloc(&f->stage.y, n * 4));
  KG_CUDA(cudaMalloc(&f->stage.dx, n * 4));
  KG_CUDA(cudaMalloc(&f->stage.dy, n * 4));
  KG_CUDA(cudaMalloc(&f->stage_cell, n * 4));
  KG_CUDA(cudaMalloc(&f->origin, n * 4));
  KG_CUDA(cudaMalloc(&f->stage4, n * 16));
  f->have_stage = true;
  return KG_OK;
}
void free_stage(kg_field2d* f) {
  cudaFree(f->stage.id); cudaFree(f->stage.x); cudaFree(f->stage.y);
  cudaFree(f->stage.dx); cudaFree(f->stage.dy); cudaFree(f->stage_cell); cudaFree(f->origin); cudaFree(f->stage4);
  f->stage4 = nullptr;
  f->stage = SoA{};
  f->stage_cell = nullptr;
  f->origin = nullptr;
  f->have_stage = false;
}
int ensure_scratch(kg_field2d* f, uint64_t n) {
  if (f->scratch_len >= n) return KG_OK;
  if (f->scratch) cudaFree(f->scratch);
  f->scratch = nullptr;
  f->scratch_len = 0;
  KG_CUDA(cudaMalloc(&f->scratch, (n + 16) * sizeof(uint32_t)));
  f->scratch_len = n;
  return KG_OK;
}
// synchronise and surface deferred device errors
int sync_check(kg_field2d* f) {
  KG_CUDA(cudaMemcpyAsync(f->h_err, f->d_err, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  KG_CUDA(cudaStreamSynchronize(f->stream));
  if (*f->h_err) {
    const int e = *f->h_err;
    KG_CUDA(cudaMemsetAsync(f->d_err, 0, sizeof(int), f->stream));
    if (e & DEV_ERR_OOB)
      return fail(KG_E_OOB, "agent coordinate outside the bag grid (reference: index out of bounds panic)");
    return fail(KG_E_INVALID, "agent id 0xFFFFFFFF is reserved (it marks a stopped agent's log entry)");
  }
  return KG_OK;
}
int use(kg_field2d* f) {
  if (!f) return fail(KG_E_INVALID, "null field handle");
  KG_CUDA(cudaSetDevice(f->device));
  return KG_OK;
}
#define LAUNCH(f, kind, kernel, grid, block, ...)             \
  do {                                                        \
    (f)->prof.begin(kind, (f)->stream);                       \
    kernel<<<grid, block, 0, (f)->stream>>>(__VA_ARGS__);     \
    (f)->prof.end((f)->stream);                               \
  } while (0)

#define LAUNCH_PDL(f, kind, kernel, grid, block, ...)                                   \
  do {                                                                                  \
    (f)->prof.begin(kind, (f)->stream);                                                 \
    cudaError_t _pe = launch_pdl(kernel, dim3(grid), dim3(block), (f)->stream, __VA_ARGS__); \
    (f)->prof.end((f)->stream);                                                         \
    if (_pe != cudaSuccess)                                                             \
      return fail(KG_E_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_pe)); \
  } while (0)

// append n entries sitting in device SoA arrays (validated first: nothing lands on KG_E_OOB)
int append_soa_dev(kg_field2d* f, uint64_t n, const SoA& s) {
  LAUNCH(f, KG_K_MISC, check_cells_kernel, blocks_for(n), kThreads, f->g, n, s.id, s.x, s.y, f->d_err);
  KG_TRY(sync_check(f));
  uint64_t o = f->n_write;
  LAUNCH(f, KG_K_MISC, pack_kernel, blocks_for(n), kThreads, n, s, f->B, o);
  LAUNCH(f, KG_K_HIST, hist_kernel, blocks_for(n), kThreads, f->g, o, n, f->B.pv, f->count, f->d_err);
  f->n_write += n;
  f->pending_new_ids = true;
  if (!f->density_estimation_check) f->nagents += n;
  return KG_OK;
}

int rebuild(kg_field2d* f) {
  // lazy_update: the log B becomes the (sorted) read buffer A; B is then logically empty
  uint64_t n = f->n_write;
  if (n > 0xFFFFFFF0ull) return fail(KG_E_CAPACITY, "more than 2^32 agents");
  f->prof.begin(KG_K_SCAN, f->stream);
  exclusive_scan_lookback(f->scan, f->count, f->g.ncells, f->cell_start, f->stream, 0, true);
  f->prof.end(f->stream);
  if (n) {
    if (f->want_origin) {
      LAUNCH_PDL(f, KG_K_SCATTER, scatter_kernel<true>, blocks_for(n, kThreads * kScatterItems), kThreads, f->g, n,
                 f->B, f->A, (const uint32_t*)f->cell_start, f->count, f->origin);
      if (f->order == KG_ORDER_CANONICAL)
        LAUNCH(f, KG_K_SORTCELL, sort_cells_kernel<true>, blocks_for(f->g.ncells, 128), 128, f->g.ncells,
               f->cell_start, f->A, f->origin);
    } else {
      LAUNCH_PDL(f, KG_K_SCATTER, scatter_kernel<false>, blocks_for(n, kThreads * kScatterItems), kThreads, f->g, n,
                 f->B, f->A, (const uint32_t*)f->cell_start, f->count, (uint32_t*)nullptr);
      if (f->order == KG_ORDER_CANONICAL)
        LAUNCH(f, KG_K_SORTCELL, sort_cells_kernel<false>, blocks_for(f->g.ncells, 128), 128, f->g.ncells,
               f->cell_start, f->A, (uint32_t*)nullptr);
    }
  }
  f->n_read = n;
  if (f->log_has_holes) {  // stopped agents were not scattered: the read buffer is shorter than the log
    uint32_t total = 0;
    KG_CUDA(cudaMemcpyAsync(&total, f->cell_start + f->g.ncells, 4, cudaMemcpyDeviceToHost, f->stream));
    KG_CUDA(cudaStreamSynchronize(f->stream));
    f->n_read = total;
    f->log_has_holes = false;
  }
  f->n_write = 0;
  f->density_estimation_check = true;
  if (f->pending_new_ids) {
    f->ids_unknown = true;
    f->pending_new_ids = false;
  }
  return KG_OK;
}

bool fast_path_ok(const kg_field2d* f, const KgBoidsParams& p, int* dd_out) {
  if (f->variant == KG_K4_GENERIC) return false;
  return k4_fast_geometry(f->g, p.radius, p.exact_query, dd_out);
}

// Decide (on the device, no host round trip) whether the read buffer's ids are unique.  Runs only
// after ids entered the field from outside (uploads); K4 itself copies ids through unchanged.
int verify_ids(kg_field2d* f) {
  if (f->variant == KG_K4_PACKED_BY_ID) {  // test/bench hook: always compare ids
    KG_CUDA(cudaMemsetAsync(f->d_ids_dup, 1, sizeof(int), f->stream));  // any non-zero word
    f->ids_unknown = true;  // re-verify once the variant is switched back
    return KG_OK;
  }
  if (!f->ids_unknown) return KG_OK;
  uint32_t n = (uint32_t)f->n_read;
  KG_CUDA(cudaMemsetAsync(f->d_ids_dup, 0, sizeof(int), f->stream));
  KG_CUDA(cudaMemsetAsync(f->id_bitmap, 0, f->id_bitmap_bits / 8, f->stream));
  if (n) {
    LAUNCH(f, KG_K_MISC, ids_mark_kernel, blocks_for(n), kThreads, n, f->A.id, f->id_bitmap_bits,
           f->id_bitmap, f->d_ids_dup);
  }
  f->ids_unknown = false;
  return KG_OK;
}

// Agents [first, first + cnt) of the read buffer take their step (cnt == 0: all of them).  The whole
// population is appended to the write log by the union of the ranges; n_write moves once, with the
// call whose range ends at n_read.
int step_boids_range(kg_field2d* f, const KgBoidsParams& p, uint64_t first64, uint64_t cnt) {
  uint64_t n = f->n_read;
  if (cnt == 0) cnt = n - first64;
  const uint64_t end = first64 + cnt;
  if (end > n) return fail(KG_E_INVALID, "step range beyond the read buffer");
  if (f->n_write + n > f->capacity)
    return fail(KG_E_CAPACITY, "write buffer cannot take %llu stepped agents",
                (unsigned long long)n);
  if (n == 0) return KG_OK;
  // stepped agents are pushed behind whatever set_object_location already appended
  Agents wr = f->B;
  wr.id += f->n_write;
  wr.pv += f->n_write;
  const uint32_t first = (uint32_t)first64;
  const bool whole = first64 == 0 && end == n;
  unsigned grid = blocks_for(cnt, 128);
  int dd = 0;
  if (fast_path_ok(f, p, &dd)) {
    if (whole && f->variant == KG_K4_FAST_SCALAR && !p.exact_query) {
      LAUNCH(f, KG_K_STEP, step_boids_fast_kernel, grid, 128, f->g, p, dd, (uint32_t)n, f->A,
             f->cell_start, wr, f->count, f->d_err);
    } else {
      KG_TRY(verify_ids(f));
      if (whole && !p.exact_query && dd == 1 && f->variant == KG_K4_TILED) {
        const int K = tile_cells_for(f->g, n);
        dim3 tgrid((unsigned)f->g.dh, (unsigned)((f->g.dw + K - 1) / K));
        static const int stage_mode = getenv("KG_TILE_STAGE") ? atoi(getenv("KG_TILE_STAGE")) : 0;  // lab hook
        LAUNCH_PDL(f, KG_K_STEP, step_boids_tile_kernel, tgrid, kTileThreads, f->g, p, K, stage_mode, f->A,
                   (const uint32_t*)f->cell_start, wr, f->count, (const int*)f->d_ids_dup, f->d_err);
      } else if (whole && !p.exact_query && dd == 1 && f->variant == KG_K4_COLTILE && f->order == KG_ORDER_ANY) {
        // chunks of kCtAgents agents per column; gridDim.y covers 1.5x the mean column, the kernel loops beyond
        const double per_col = (double)n / (double)std::max(1, f->g.max_x);
        const unsigned gy = (unsigned)std::min(64.0, std::max(1.0, std::ceil(1.5 * per_col / kCtAgents)));
        dim3 cgrid((unsigned)f->g.dw, gy);
        static const int ct_stage = getenv("KG_CT_STAGE") ? atoi(getenv("KG_CT_STAGE")) : 0;  // lab hook
        if (ct_stage == 0)
          LAUNCH_PDL(f, KG_K_STEP, step_boids_coltile_kernel<0>, cgrid, kCtThreads, f->g, p, f->A,
                     (const uint32_t*)f->cell_start, wr, f->count, (const int*)f->d_ids_dup, f->d_err);
        else
          LAUNCH_PDL(f, KG_K_STEP, step_boids_coltile_kernel<1>, cgrid, kCtThreads, f->g, p, f->A,
                     (const uint32_t*)f->cell_start, wr, f->count, (const int*)f->d_ids_dup, f->d_err);
      } else if (whole && !p.exact_query && dd == 1 && f->variant == KG_K4_STAGED) {
        const double per_col = (double)n / (double)std::max(1, f->g.max_x);
        const unsigned gy = (unsigned)std::min(256.0, std::max(1.0, std::ceil(1.5 * per_col / kSgThreads)));
        dim3 cgrid((unsigned)f->g.dw, gy);
        LAUNCH_PDL(f, KG_K_STEP, step_boids_staged_kernel, cgrid, kSgThreads, f->g, p, f->A,
                   (const uint32_t*)f->cell_start, wr, f->count, (const int*)f->d_ids_dup, f->d_err);
      } else if (p.exact_query)
        LAUNCH_PDL(f, KG_K_STEP, step_boids_packed_kernel<true>, grid, 128, f->g, p, dd,
                   exact_threshold(p.radius), (uint32_t)end, f->A, (const uint32_t*)f->cell_start, wr,
                   f->count, (const int*)f->d_ids_dup, f->d_err, KgLifeRule{}, (uint32_t*)nullptr, first);
      else {
        static const int k4_block = getenv("KG_K4_BLOCK") ? atoi(getenv("KG_K4_BLOCK")) : 128;  // lab hook: 32 / 64 / 128
        LAUNCH_PDL(f, KG_K_STEP, step_boids_packed_kernel<false>, blocks_for(cnt, k4_block), k4_block, f->g, p, dd, 0.0f,
                   (uint32_t)end, f->A, (const uint32_t*)f->cell_start, wr, f->count,
                   (const int*)f->d_ids_dup, f->d_err, KgLifeRule{}, (uint32_t*)nullptr, first);
      }
    }
  } else if (p.exact_query)
    LAUNCH(f, KG_K_STEP, step_boids_kernel<true>, grid, 128, f->g, p, (uint32_t)end, f->A,
           f->cell_start, wr, f->count, f->d_err, KgLifeRule{}, (uint32_t*)nullptr, first);
  else
    LAUNCH(f, KG_K_STEP, step_boids_kernel<false>, grid, 128, f->g, p, (uint32_t)end, f->A,
           f->cell_start, wr, f->count, f->d_err, KgLifeRule{}, (uint32_t*)nullptr, first);
  if (end == n) f->n_write += n;
  return KG_OK;
}
int step_boids(kg_field2d* f, const KgBoidsParams& p) { return step_boids_range(f, p, 0, 0); }

}  // namespace

extern "C" {

const char* kg_last_error(void) { return last_error().c_str(); }
int kg_abi_version(void) { return KG_ABI_VERSION; }
int kg_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) return fail(KG_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  return n;
}
uint64_t kg_launch_count(void) { return launch_counter().load(); }
int kg_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  KG_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return KG_OK;
}
int kg_host_free(void* p) {
  KG_CUDA(cudaFreeHost(p));
  return KG_OK;
}

int kg_selftest_div(int device, uint64_t n, uint64_t seed, uint64_t* mismatches) {
  if (!mismatches) return fail(KG_E_INVALID, "null out");
  KG_CUDA(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  KG_CUDA(cudaMalloc(&d, 8));
  KG_CUDA(cudaMemset(d, 0, 8));
  selftest_div_kernel<<<kNumSMs * 8, 256>>>(n, seed, d);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) return fail(KG_E_CUDA, "selftest_div: %s", cudaGetErrorString(e));
  *mismatches = h;
  return KG_OK;
}

int kg_field2d_create(float w, float h, float d, int toroidal, uint64_t capacity, int device,
                      kg_field2d** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (!(w > 0.f) || !(h > 0.f) || !(d > 0.f)) return fail(KG_E_INVALID, "w, h, discretization must be > 0");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_field2d* f = new kg_field2d();
  f->device = device;
  f->capacity = capacity;
  // same f32 operation order as field_2d.rs:317-318 / :487-488
  f->g.w = w; f->g.h = h; f->g.disc = d; f->g.toroidal = toroidal ? 1 : 0;
  f->g.max_x = (int)std::min(2147483520.0f, ceilf(w / d));
  f->g.max_y = (int)std::min(2147483520.0f, ceilf(h / d));
  f->g.dw = f->g.max_x + 1;
  f->g.dh = f->g.max_y + 1;
  uint64_t nc = (uint64_t)f->g.dw * (uint64_t)f->g.dh;
  if (nc >= (1ull << 31)) { delete f; return fail(KG_E_INVALID, "bag grid of %llu cells is too large", (unsigned long long)nc); }
  f->g.ncells = (uint32_t)nc;
  int rc = KG_OK;
  // one bit per id up to 8x the capacity (at least 2^22): ids beyond that cannot be verified
  // unique and select the id-comparing loop of the packed K4
  const uint64_t id_bits = std::max<uint64_t>(8 * capacity, 1ull << 22) / 256 * 256 + 256;
  f->id_bitmap_bits = id_bits;
  auto cleanup = [&](int code) { kg_field2d_destroy(f); return code; };
  if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  if ((rc = alloc_agents(f->A, capacity)) != KG_OK) return cleanup(rc);
  if ((rc = alloc_agents(f->B, capacity)) != KG_OK) return cleanup(rc);
  if ((rc = lookback_init(f->scan, nc, f->stream)) != KG_OK) return cleanup(rc);
  if (cudaMalloc(&f->cell_start, (nc + 16) * 4) != cudaSuccess ||
      cudaMalloc(&f->count, (nc + 16) * 4) != cudaSuccess ||
      cudaMalloc(&f->tile_sums, ((uint64_t)scan_num_tiles(capacity + 1) + 16) * 4) != cudaSuccess ||
      cudaMalloc(&f->d_err, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&f->d_ids_dup, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&f->id_bitmap, id_bits / 8) != cudaSuccess ||
      cudaHostAlloc(&f->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess)
    return cleanup(fail(KG_E_CUDA, "device allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaMemsetAsync(f->cell_start, 0, (nc + 16) * 4, f->stream);
  cudaMemsetAsync(f->count, 0, (nc + 16) * 4, f->stream);
  cudaMemsetAsync(f->d_err, 0, sizeof(int), f->stream);
  if (cudaStreamSynchronize(f->stream) != cudaSuccess)
    return cleanup(fail(KG_E_CUDA, "init sync failed"));
  *out = f;
  return KG_OK;
}

int kg_field2d_destroy(kg_field2d* f) {
  if (!f) return KG_OK;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  f->prof.destroy();
  f->watch.destroy();
  f->flusher.destroy();
  f->events.destroy();
  free_agents(f->A);
  free_agents(f->B);
  free_agents(f->remove_tmp);
  for (auto& e : f->slab_ev)
    if (e) cudaEventDestroy(e);
  if (f->copy_done) cudaEventDestroy(f->copy_done);
  if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
  cudaFree(f->qbuf);
  cudaFree(f->rbuf);
  cudaFree(f->red);
  free_stage(f);
  lookback_destroy(f->scan);
  cudaFree(f->cell_start);
  cudaFree(f->count);
  cudaFree(f->tile_sums);
  cudaFree(f->scratch);
  cudaFree(f->d_err);
  cudaFree(f->d_ids_dup);
  cudaFree(f->id_bitmap);
  if (f->h_err) cudaFreeHost(f->h_err);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return KG_OK;
}

int kg_field2d_sync(kg_field2d* f) {
  KG_TRY(use(f));
  return sync_check(f);
}

int kg_field2d_dims(kg_field2d* f, int32_t* dw, int32_t* dh, int32_t* max_x, int32_t* max_y) {
  if (!f) return fail(KG_E_INVALID, "null field handle");
  if (dw) *dw = f->g.dw;
  if (dh) *dh = f->g.dh;
  if (max_x) *max_x = f->g.max_x;
  if (max_y) *max_y = f->g.max_y;
  return KG_OK;
}

int kg_field2d_set_order(kg_field2d* f, int order) {
  if (!f) return fail(KG_E_INVALID, "null field handle");
  if (order != KG_ORDER_ANY && order != KG_ORDER_CANONICAL) return fail(KG_E_INVALID, "bad order");
  f->order = order;
  return KG_OK;
}
int kg_field2d_set_kernel_variant(kg_field2d* f, int variant) {
  if (!f) return fail(KG_E_INVALID, "null field handle");
  if (variant < KG_K4_AUTO || variant > KG_K4_STAGED) return fail(KG_E_INVALID, "bad K4 variant");
  f->variant = variant;
  return KG_OK;
}

int kg_field2d_set_object_locations(kg_field2d* f, uint64_t n, const uint32_t* id, const float* x,
                                    const float* y, const float* dx, const float* dy) {
  KG_TRY(use(f));
  if (n == 0) return KG_OK;
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null input array");
  if (f->n_write + n > f->capacity)
    return fail(KG_E_CAPACITY, "write buffer holds %llu, +%llu exceeds capacity %llu",
                (unsigned long long)f->n_write, (unsigned long long)n,
                (unsigned long long)f->capacity);
  KG_TRY(ensure_stage(f));
  cudaStream_t s = f->stream;
  KG_CUDA(cudaMemcpyAsync(f->stage.id, id, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.x, x, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.y, y, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.dx, dx, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.dy, dy, n * 4, cudaMemcpyHostToDevice, s));
  KG_TRY(append_soa_dev(f, n, f->stage));
  uint32_t top = 0;
  for (uint64_t i = 0; i < n; ++i) top = std::max(top, id[i]);
  f->next_id = std::max(f->next_id, top + 1u);  // dynamic population: children get fresh ids
  return KG_OK;
}

int kg_field2d_set_object_locations_dev(kg_field2d* f, uint64_t n, const uint32_t* id,
                                        const float* x, const float* y, const float* dx,
                                        const float* dy) {
  KG_TRY(use(f));
  if (n == 0) return KG_OK;
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null input array");
  if (f->n_write + n > f->capacity)
    return fail(KG_E_CAPACITY, "write buffer holds %llu, +%llu exceeds capacity %llu",
                (unsigned long long)f->n_write, (unsigned long long)n,
                (unsigned long long)f->capacity);
  SoA s;
  s.id = const_cast<uint32_t*>(id);
  s.x = const_cast<float*>(x);
  s.y = const_cast<float*>(y);
  s.dx = const_cast<float*>(dx);
  s.dy = const_cast<float*>(dy);
  return append_soa_dev(f, n, s);
}

int kg_field2d_remove_object_location(kg_field2d* f, uint32_t id, float x, float y) {
  KG_TRY(use(f));
  // discretize on the host with the same f32 ops as the device (division + floor are exact IEEE)
  int cx = (int)floorf(x / f->g.disc), cy = (int)floorf(y / f->g.disc);
  uint32_t cell = (uint32_t)cx * (uint32_t)f->g.dh + (uint32_t)cy;
  if ((int32_t)cell < 0 || cell >= f->g.ncells)
    return fail(KG_E_OOB, "remove_object_location: location outside the bag grid");
  uint64_t n = f->n_write;
  if (n == 0) return KG_OK;
  KG_TRY(ensure_scratch(f, 2 * (n + 1) + 64));
  uint32_t* keep = f->scratch;
  uint32_t* keep_scan = f->scratch + ((n + 1 + 15) / 16) * 16;
  LAUNCH(f, KG_K_MISC, mark_remove_kernel, blocks_for(n), kThreads, f->g, n, f->B, id, cell, keep,
         f->count);
  exclusive_scan_u32(keep, n, keep_scan, f->tile_sums, f->stream);
  launch_counter().fetch_add(3, std::memory_order_relaxed);
  if (!f->remove_tmp.id) KG_TRY(alloc_agents(f->remove_tmp, f->capacity));  // once per handle
  // compact into the spare buffer, then make it the log (pointer swap: no copy back, one sync for the count)
  LAUNCH(f, KG_K_MISC, compact_kernel, blocks_for(n), kThreads, n, keep_scan, f->B, f->remove_tmp);
  uint32_t kept = 0;
  KG_CUDA(cudaMemcpyAsync(&kept, keep_scan + n, 4, cudaMemcpyDeviceToHost, f->stream));
  KG_CUDA(cudaStreamSynchronize(f->stream));
  std::swap(f->B, f->remove_tmp);
  if (!f->density_estimation_check) f->nagents -= (n - kept);
  f->n_write = kept;
  return KG_OK;
}

int kg_field2d_lazy_update(kg_field2d* f) {
  KG_TRY(use(f));
  return rebuild(f);
}
int kg_field2d_update(kg_field2d* f) {
  KG_TRY(use(f));
  return KG_OK;
}

int kg_field2d_nagents(kg_field2d* f, uint64_t* out) {
  if (!f || !out) return fail(KG_E_INVALID, "null argument");
  *out = f->nagents;
  return KG_OK;
}
int kg_field2d_num_objects(kg_field2d* f, int which, uint64_t* out) {
  if (!f || !out) return fail(KG_E_INVALID, "null argument");
  *out = which == KG_BUF_READ ? f->n_read : f->n_write;
  return KG_OK;
}

int kg_field2d_download(kg_field2d* f, int which, uint64_t cap, uint32_t* id, float* x, float* y,
                        float* dx, float* dy, int32_t* cell, uint64_t* n_out) {
  KG_TRY(use(f));
  const Agents& a = which == KG_BUF_READ ? f->A : f->B;
  uint64_t n = which == KG_BUF_READ ? f->n_read : f->n_write;
  if (n_out) *n_out = n;
  if (n > cap) return fail(KG_E_CAPACITY, "download needs room for %llu agents", (unsigned long long)n);
  cudaStream_t st = f->stream;
  if (n) {
    KG_TRY(ensure_stage(f));
    LAUNCH(f, KG_K_MISC, unpack_kernel, blocks_for(n), kThreads, f->g, n, a, f->stage,
           cell ? f->stage_cell : nullptr);
    if (id) KG_CUDA(cudaMemcpyAsync(id, f->stage.id, n * 4, cudaMemcpyDeviceToHost, st));
    if (x) KG_CUDA(cudaMemcpyAsync(x, f->stage.x, n * 4, cudaMemcpyDeviceToHost, st));
    if (y) KG_CUDA(cudaMemcpyAsync(y, f->stage.y, n * 4, cudaMemcpyDeviceToHost, st));
    if (dx) KG_CUDA(cudaMemcpyAsync(dx, f->stage.dx, n * 4, cudaMemcpyDeviceToHost, st));
    if (dy) KG_CUDA(cudaMemcpyAsync(dy, f->stage.dy, n * 4, cudaMemcpyDeviceToHost, st));
    if (cell) KG_CUDA(cudaMemcpyAsync(cell, f->stage_cell, n * 4, cudaMemcpyDeviceToHost, st));
  }
  return sync_check(f);
}

int kg_field2d_cell_counts(kg_field2d* f, int which, uint64_t cap, uint32_t* counts) {
  KG_TRY(use(f));
  if (!counts) return fail(KG_E_INVALID, "null output");
  uint64_t nc = f->g.ncells;
  if (cap < nc) return fail(KG_E_CAPACITY, "cell_counts needs %llu entries", (unsigned long long)nc);
  if (which == KG_BUF_WRITE) {
    KG_CUDA(cudaMemcpyAsync(counts, f->count, nc * 4, cudaMemcpyDeviceToHost, f->stream));
  } else {
    KG_TRY(ensure_scratch(f, nc));
    LAUNCH(f, KG_K_MISC, cell_counts_from_starts_kernel, blocks_for(nc), kThreads, (uint32_t)nc,
           f->cell_start, f->scratch);
    KG_CUDA(cudaMemcpyAsync(counts, f->scratch, nc * 4, cudaMemcpyDeviceToHost, f->stream));
  }
  return sync_check(f);
}

int kg_field2d_num_objects_at_locations(kg_field2d* f, uint64_t nq, const float* x, const float* y,
                                        uint32_t* out) {
  KG_TRY(use(f));
  if (nq == 0) return KG_OK;
  if (!x || !y || !out) return fail(KG_E_INVALID, "null argument");
  KG_TRY(ensure_scratch(f, 3 * nq + 64));
  float* dx = (float*)f->scratch;
  float* dy = dx + nq;
  uint32_t* dout = f->scratch + 2 * nq;
  KG_CUDA(cudaMemcpyAsync(dx, x, nq * 4, cudaMemcpyHostToDevice, f->stream));
  KG_CUDA(cudaMemcpyAsync(dy, y, nq * 4, cudaMemcpyHostToDevice, f->stream));
  LAUNCH(f, KG_K_QUERY, num_at_locations_kernel, blocks_for(nq), kThreads, f->g, nq, dx, dy,
         f->cell_start, dout, f->d_err);
  KG_CUDA(cudaMemcpyAsync(out, dout, nq * 4, cudaMemcpyDeviceToHost, f->stream));
  return sync_check(f);
}

int kg_field2d_get_objects(kg_field2d* f, int which, float x, float y, uint64_t cap, uint32_t* ids,
                           uint64_t* n_out) {
  KG_TRY(use(f));
  int cx = (int)floorf(x / f->g.disc), cy = (int)floorf(y / f->g.disc);
  uint32_t cell = (uint32_t)cx * (uint32_t)f->g.dh + (uint32_t)cy;
  if ((int32_t)cell < 0 || cell >= f->g.ncells)
    return fail(KG_E_OOB, "get_objects: location outside the bag grid");
  if (which == KG_BUF_READ) {
    uint32_t se[2];
    KG_CUDA(cudaMemcpyAsync(se, f->cell_start + cell, 8, cudaMemcpyDeviceToHost, f->stream));
    KG_CUDA(cudaStreamSynchronize(f->stream));
    uint64_t n = se[1] - se[0];
    if (n_out) *n_out = n;
    if (n > cap) return fail(KG_E_CAPACITY, "get_objects needs room for %llu ids", (unsigned long long)n);
    if (n && ids) KG_CUDA(cudaMemcpyAsync(ids, f->A.id + se[0], n * 4, cudaMemcpyDeviceToHost, f->stream));
    return sync_check(f);
  }
  // write buffer: the log is unsorted; filter it on the host side of the boundary (debug API)
  uint64_t n = f->n_write;
  std::vector<uint32_t> hid(n);
  std::vector<int32_t> hcell(n);
  uint64_t got = 0;
  KG_TRY(kg_field2d_download(f, KG_BUF_WRITE, n, hid.data(), nullptr, nullptr, nullptr, nullptr,
                             hcell.data(), &got));
  uint64_t k = 0;
  for (uint64_t i = 0; i < n; ++i)
    if ((uint32_t)hcell[i] == cell) {
      if (k < cap && ids) ids[k] = hid[i];
      ++k;
    }
  if (n_out) *n_out = k;
  if (k > cap) return fail(KG_E_CAPACITY, "get_objects needs room for %llu ids", (unsigned long long)k);
  return KG_OK;
}

int kg_field2d_num_empty_bags(kg_field2d* f, uint64_t* out) {
  KG_TRY(use(f));
  if (!out) return fail(KG_E_INVALID, "null output");
  KG_TRY(ensure_scratch(f, 16));
  unsigned long long* d = (unsigned long long*)f->scratch;
  KG_CUDA(cudaMemsetAsync(d, 0, 8, f->stream));
  LAUNCH(f, KG_K_MISC, count_empty_kernel, blocks_for(f->g.ncells), kThreads, f->g.ncells,
         f->cell_start, d);
  unsigned long long h = 0;
  KG_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, f->stream));
  KG_TRY(sync_check(f));
  *out = h;
  return KG_OK;
}

// grow-only device scratch owned by the handle: queries allocate nothing per call
static int ensure_bytes(void** buf, size_t* have, size_t want) {
  if (*have >= want) return KG_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *have = 0;
  want += want / 2 + 4096;
  KG_CUDA(cudaMalloc(buf, want));
  *have = want;
  return KG_OK;
}

// get_neighbors_within_{relax_,}distance for nq query points in one launch pair (count, fill).  With
// x/y/ldx/ldy non-null the neighbours themselves are returned (the reference returns Vec<O>), not
// only their ids.  One host synchronisation (the total), no allocation in steady state.
int kg_field2d_neighbors_agents(kg_field2d* f, uint64_t nq, const float* qx, const float* qy, float dist, int mode,
                                uint64_t* offsets, uint32_t* ids, float* x, float* y, float* ldx, float* ldy,
                                uint64_t cap, uint64_t* total_out) {
  KG_TRY(use(f));
  if (!offsets) return fail(KG_E_INVALID, "null offsets");
  offsets[0] = 0;
  if (total_out) *total_out = 0;
  if (nq == 0) return KG_OK;
  if (!qx || !qy) return fail(KG_E_INVALID, "null query array");
  if (mode != KG_QUERY_RELAX && mode != KG_QUERY_EXACT) return fail(KG_E_INVALID, "bad query mode");
  const bool payload = x && y && ldx && ldy;
  cudaStream_t s = f->stream;
  // query-sized part: qx, qy, counts, scan, 64-bit offsets, scan tiles
  const size_t a4 = ((nq + 32) * 4 + 15) / 16 * 16;  // the scan reads and writes uint4: keep every part 16-byte aligned
  const size_t qbytes = 4 * a4 + (nq + 2) * 8 + ((size_t)scan_num_tiles(nq) + 16) * 4 + 256;
  KG_TRY(ensure_bytes(&f->qbuf, &f->qbuf_bytes, qbytes));
  char* q = (char*)f->qbuf;
  float* dqx = (float*)q;
  float* dqy = (float*)(q + a4);
  uint32_t* dcnt = (uint32_t*)(q + 2 * a4);
  uint32_t* dscan = (uint32_t*)(q + 3 * a4);
  uint64_t* doff = (uint64_t*)(q + 4 * a4);
  uint32_t* dtiles = (uint32_t*)(q + 4 * a4 + (nq + 2) * 8);
  KG_CUDA(cudaMemcpyAsync(dqx, qx, nq * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(dqy, qy, nq * 4, cudaMemcpyHostToDevice, s));
  if (mode == KG_QUERY_EXACT)
    LAUNCH(f, KG_K_QUERY, query_count_kernel<true>, blocks_for(nq, 128), 128, f->g, nq, dqx, dqy, dist,
           f->cell_start, f->A.pv, dcnt);
  else
    LAUNCH(f, KG_K_QUERY, query_count_kernel<false>, blocks_for(nq, 128), 128, f->g, nq, dqx, dqy, dist,
           f->cell_start, f->A.pv, dcnt);
  exclusive_scan_u32(dcnt, nq, dscan, dtiles, s);
  launch_counter().fetch_add(3, std::memory_order_relaxed);
  LAUNCH(f, KG_K_MISC, widen_offsets_kernel, blocks_for(nq + 1), kThreads, nq, dscan, doff);
  KG_CUDA(cudaMemcpyAsync(offsets, doff, (nq + 1) * 8, cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  const uint64_t total = offsets[nq];
  if (total_out) *total_out = total;
  if (total > cap) return fail(KG_E_CAPACITY, "neighbour list needs %llu entries", (unsigned long long)total);
  if (total && ids) {
    // result-sized part: ids, and for the payload one float4 per neighbour plus its SoA split
    const size_t r4 = (total + 32) * 4;
    KG_TRY(ensure_bytes(&f->rbuf, &f->rbuf_bytes, r4 + (payload ? (total + 2) * 16 + 4 * r4 : 0) + 256));
    char* r = (char*)f->rbuf;
    float4* dag = payload ? (float4*)r : nullptr;
    char* rest = r + (payload ? (total + 2) * 16 : 0);
    uint32_t* dids = (uint32_t*)rest;
    if (mode == KG_QUERY_EXACT)
      LAUNCH(f, KG_K_QUERY, query_fill_kernel<true>, blocks_for(nq, 128), 128, f->g, nq, dqx, dqy, dist,
             f->cell_start, f->A.pv, f->A.id, doff, dids, dag, total);
    else
      LAUNCH(f, KG_K_QUERY, query_fill_kernel<false>, blocks_for(nq, 128), 128, f->g, nq, dqx, dqy, dist,
             f->cell_start, f->A.pv, f->A.id, doff, dids, dag, total);
    KG_CUDA(cudaMemcpyAsync(ids, dids, total * 4, cudaMemcpyDeviceToHost, s));
    if (payload) {
      float* sx = (float*)(rest + r4);
      float* sy = (float*)(rest + 2 * r4);
      float* sdx = (float*)(rest + 3 * r4);
      float* sdy = (float*)(rest + 4 * r4);
      LAUNCH(f, KG_K_MISC, split_agents_kernel, blocks_for(total), kThreads, total, (const float4*)dag, sx, sy, sdx, sdy);
      KG_CUDA(cudaMemcpyAsync(x, sx, total * 4, cudaMemcpyDeviceToHost, s));
      KG_CUDA(cudaMemcpyAsync(y, sy, total * 4, cudaMemcpyDeviceToHost, s));
      KG_CUDA(cudaMemcpyAsync(ldx, sdx, total * 4, cudaMemcpyDeviceToHost, s));
      KG_CUDA(cudaMemcpyAsync(ldy, sdy, total * 4, cudaMemcpyDeviceToHost, s));
    }
  }
  return sync_check(f);
}

int kg_field2d_neighbors(kg_field2d* f, uint64_t nq, const float* qx, const float* qy, float dist,
                         int mode, uint64_t* offsets, uint32_t* ids, uint64_t cap,
                         uint64_t* total_out) {
  return kg_field2d_neighbors_agents(f, nq, qx, qy, dist, mode, offsets, ids, nullptr, nullptr, nullptr, nullptr,
                                     cap, total_out);
}

int kg_field2d_step_boids(kg_field2d* f, const KgBoidsParams* p) {
  KG_TRY(use(f));
  if (!p) return fail(KG_E_INVALID, "null params");
  return step_boids(f, *p);
}

int kg_jit_agent_source(const char* pair, const char* finish, int may_stop, char* out, uint64_t cap, uint64_t* need) {
  if (!pair || !finish || !need) return fail(KG_E_INVALID, "null argument");
  const std::string src = jit::agent_step_source(pair, finish, may_stop != 0);
  *need = src.size() + 1;
  if (out && cap >= *need) memcpy(out, src.c_str(), *need);
  return KG_OK;
}

int kg_field2d_step_custom(kg_field2d* f, const KgCustomStep* cs) {
  KG_TRY(use(f));
  if (!cs || !cs->pair || !cs->finish) return fail(KG_E_INVALID, "null custom step");
  if (cs->nconsts < 0 || cs->nconsts > 16) return fail(KG_E_INVALID, "at most 16 constants");
  const uint64_t n = f->n_read;
  if (f->n_write + n > f->capacity)
    return fail(KG_E_CAPACITY, "write buffer cannot take %llu stepped agents", (unsigned long long)n);
  KG_CUDA(cudaFree(nullptr));  // the runtime's primary context is current for the driver calls below
  CUfunction fn;
  KG_TRY(jit::get_kernel(f->device, jit::agent_step_source(cs->pair, cs->finish, cs->may_stop != 0), "kg_agent_step", &fn));
  if (n == 0) return KG_OK;
  struct { float c[16]; } consts;
  for (int k = 0; k < 16; ++k) consts.c[k] = k < cs->nconsts ? cs->consts[k] : 0.0f;
  Geom g = f->g;
  uint32_t n32 = (uint32_t)n;
  const uint32_t* rid = f->A.id;
  const float4* rpv = f->A.pv;
  const uint32_t* cell_start = f->cell_start;
  uint32_t* wid = f->B.id + f->n_write;
  float4* wpv = f->B.pv + f->n_write;
  uint32_t* count = f->count;
  int* err = f->d_err;
  float dist = cs->radius;
  int exact = cs->exact_query;
  unsigned long long seed = cs->seed, step = cs->step;
  void* args[] = {&g, &n32, &rid, &rpv, &cell_start, &wid, &wpv, &count, &err, &dist, &exact, &seed, &step, &consts};
  f->prof.begin(KG_K_STEP, f->stream);
  const int rc = jit::launch(fn, blocks_for(n, 128), 128, f->stream, args);
  f->prof.end(f->stream);
  KG_TRY(rc);
  f->n_write += n;
  if (cs->may_stop) f->log_has_holes = true;
  return KG_OK;
}

int kg_field2d_run_boids(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps) {
  KG_TRY(use(f));
  if (!p) return fail(KG_E_INVALID, "null params");
  KgBoidsParams q = *p;
  for (uint64_t i = 0; i < nsteps; ++i) {
    q.step = p->step + i;
    KG_TRY(step_boids(f, q));
    KG_TRY(rebuild(f));
  }
  return KG_OK;
}

int kg_field2d_init_flockers(kg_field2d* f, uint64_t n, uint64_t seed) {
  KG_TRY(use(f));
  if (n == 0) return KG_OK;
  if (f->n_write + n > f->capacity) return fail(KG_E_CAPACITY, "init_flockers exceeds capacity");
  LAUNCH(f, KG_K_MISC, init_flockers_kernel, blocks_for(n), kThreads, f->g, f->n_write, n, seed, f->B,
         f->count, f->d_err);
  f->n_write += n;
  f->pending_new_ids = true;
  f->next_id = std::max<uint32_t>(f->next_id, (uint32_t)n);
  if (!f->density_estimation_check) f->nagents += n;
  return KG_OK;
}

int kg_field2d_set_next_id(kg_field2d* f, uint32_t next_id) {
  if (!f) return fail(KG_E_INVALID, "null field handle");
  f->next_id = next_id;
  return KG_OK;
}

int kg_field2d_step_boids_life(kg_field2d* f, const KgBoidsParams* p, const KgLifeRule* life,
                               uint64_t* n_stopped, uint64_t* n_born) {
  KG_TRY(use(f));
  if (!p || !life) return fail(KG_E_INVALID, "null argument");
  if (!(life->death_prob >= 0.f && life->death_prob <= 1.f && life->birth_prob >= 0.f && life->birth_prob <= 1.f))
    return fail(KG_E_INVALID, "death_prob and birth_prob must lie in [0, 1]");
  if (n_stopped) *n_stopped = 0;
  if (n_born) *n_born = 0;
  const uint64_t n = f->n_read;
  if (f->n_write + n > f->capacity)
    return fail(KG_E_CAPACITY, "write buffer cannot take %llu stepped agents", (unsigned long long)n);
  if (n == 0) return KG_OK;
  // scratch: birth flags [n], their exclusive scan [n + 1], one 64-bit counter
  KG_TRY(ensure_scratch(f, 2 * (n + 32) + 64));
  uint32_t* birth = f->scratch;
  uint32_t* scan = f->scratch + ((n + 31) / 16) * 16;
  unsigned long long* d_stopped = (unsigned long long*)(scan + ((n + 1 + 31) / 16) * 16);
  Agents wr = f->B;
  wr.id += f->n_write;
  wr.pv += f->n_write;
  const unsigned grid = blocks_for(n, 128);
  int dd = 0;
  if (fast_path_ok(f, *p, &dd)) {
    KG_TRY(verify_ids(f));
    if (p->exact_query)
      LAUNCH_PDL(f, KG_K_STEP, (step_boids_packed_kernel<true, true>), grid, 128, f->g, *p, dd,
                 exact_threshold(p->radius), (uint32_t)n, f->A, (const uint32_t*)f->cell_start, wr, f->count,
                 (const int*)f->d_ids_dup, f->d_err, *life, birth, 0u);
    else
      LAUNCH_PDL(f, KG_K_STEP, (step_boids_packed_kernel<false, true>), grid, 128, f->g, *p, dd, 0.0f,
                 (uint32_t)n, f->A, (const uint32_t*)f->cell_start, wr, f->count, (const int*)f->d_ids_dup,
                 f->d_err, *life, birth, 0u);
  } else if (p->exact_query) {
    LAUNCH(f, KG_K_STEP, (step_boids_kernel<true, true>), grid, 128, f->g, *p, (uint32_t)n, f->A, f->cell_start,
           wr, f->count, f->d_err, *life, birth);
  } else {
    LAUNCH(f, KG_K_STEP, (step_boids_kernel<false, true>), grid, 128, f->g, *p, (uint32_t)n, f->A, f->cell_start,
           wr, f->count, f->d_err, *life, birth);
  }
  // births: ranks of the parents in read-buffer (iter_objects) order
  exclusive_scan_u32(birth, n, scan, f->tile_sums, f->stream);
  launch_counter().fetch_add(3, std::memory_order_relaxed);
  KG_CUDA(cudaMemsetAsync(d_stopped, 0, 8, f->stream));
  LAUNCH(f, KG_K_MISC, count_stopped_kernel, blocks_for(n), kThreads, (uint32_t)n, (const uint32_t*)wr.id,
         d_stopped);
  uint32_t nb = 0;
  unsigned long long ns = 0;
  KG_CUDA(cudaMemcpyAsync(&nb, scan + n, 4, cudaMemcpyDeviceToHost, f->stream));
  KG_CUDA(cudaMemcpyAsync(&ns, d_stopped, 8, cudaMemcpyDeviceToHost, f->stream));
  KG_TRY(sync_check(f));
  f->n_write += n;
  f->log_has_holes = f->log_has_holes || ns != 0;
  if (nb) {
    if (f->n_write + nb > f->capacity)
      return fail(KG_E_CAPACITY, "%u births do not fit the field's capacity of %llu agents", nb,
                  (unsigned long long)f->capacity);
    if ((uint64_t)f->next_id + nb >= (uint64_t)kIdNone) return fail(KG_E_CAPACITY, "agent ids exhausted");
    Agents ch = f->B;
    ch.id += f->n_write;
    ch.pv += f->n_write;
    LAUNCH(f, KG_K_MISC, spawn_kernel, blocks_for(n), kThreads, f->g, (uint32_t)n, f->A, (const uint32_t*)birth,
           (const uint32_t*)scan, f->next_id, ch, f->count, f->d_err);
    f->n_write += nb;
    f->next_id += nb;
    f->pending_new_ids = true;
  }
  if (n_stopped) *n_stopped = ns;
  if (n_born) *n_born = nb;
  return KG_OK;
}

int kg_field2d_step_boids_host(kg_field2d* f, const KgBoidsParams* p, uint64_t n,
                               const uint32_t* id_in, const float* x_in, const float* y_in,
                               const float* dx_in, const float* dy_in, uint32_t* id_out,
                               float* x_out, float* y_out, float* dx_out, float* dy_out) {
  KG_TRY(use(f));
  if (!p) return fail(KG_E_INVALID, "null params");
  if (n && (!id_in || !x_in || !y_in || !dx_in || !dy_in || !id_out || !x_out || !y_out || !dx_out || !dy_out))
    return fail(KG_E_INVALID, "null host array");
  if (n > f->capacity) return fail(KG_E_CAPACITY, "%llu agents exceed the capacity %llu", (unsigned long long)n,
                                   (unsigned long long)f->capacity);
  // start from empty buffers: this entry point owns the whole state for the call
  f->n_write = 0;
  f->n_read = 0;
  f->log_has_holes = false;
  cudaStream_t s = f->stream;
  KG_CUDA(cudaMemsetAsync(f->count, 0, (size_t)f->g.ncells * 4, s));
  if (n == 0) return sync_check(f);
  if (!f->copy_stream) {
    KG_CUDA(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
    for (auto& e : f->slab_ev) KG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    KG_CUDA(cudaEventCreateWithFlags(&f->copy_done, cudaEventDisableTiming));
  }
  KG_TRY(ensure_stage(f));
  // 1. upload; one kernel packs, histograms and validates (flags are read at the end: no sync here)
  KG_CUDA(cudaMemcpyAsync(f->stage.id, id_in, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.x, x_in, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.y, y_in, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.dx, dx_in, n * 4, cudaMemcpyHostToDevice, s));
  KG_CUDA(cudaMemcpyAsync(f->stage.dy, dy_in, n * 4, cudaMemcpyHostToDevice, s));
  LAUNCH(f, KG_K_MISC, pack_hist_kernel, blocks_for(n), kThreads, f->g, n, f->stage, f->B, f->count, f->d_err);
  f->n_write = n;
  f->pending_new_ids = true;
  // 2. lazy_update: the uploaded log becomes the sorted read buffer
  KG_TRY(rebuild(f));
  // 3. every agent's step, slab by slab of the read buffer; a finished slab is unpacked and travels to
  //    the host on the copy stream while the next slab computes.  The result is the write log, i.e.
  //    the agents in the cell order of their INPUT positions (ids travel with them) — the closing
  //    lazy_update would only reorder what the host receives, so it is left to the next call's upload.
  // slabs of at least 2^20 agents (4 MB per host array): smaller copies waste PCIe time on their fixed
  // cost — measured at 1M agents: 8 slabs 0.96 ms per call against 0.93 ms unslabbed
  const int nslab = (int)std::max<uint64_t>(1, std::min<uint64_t>(8, n >> 20));
  const uint64_t per = ((n + nslab - 1) / nslab + 127) / 128 * 128;
  for (int k = 0; k < nslab; ++k) {
    const uint64_t a0 = std::min<uint64_t>(n, (uint64_t)k * per), a1 = std::min<uint64_t>(n, a0 + per);
    if (a1 == a0) continue;
    KG_TRY(step_boids_range(f, *p, a0, a1 - a0));
    LAUNCH(f, KG_K_MISC, unpack_range_kernel, blocks_for(a1 - a0), kThreads, a0, a1 - a0, f->B, f->stage);
    KG_CUDA(cudaEventRecord(f->slab_ev[k], s));
    KG_CUDA(cudaStreamWaitEvent(f->copy_stream, f->slab_ev[k], 0));
    const size_t off = (size_t)a0, bytes = (size_t)(a1 - a0) * 4;
    KG_CUDA(cudaMemcpyAsync(id_out + off, f->stage.id + off, bytes, cudaMemcpyDeviceToHost, f->copy_stream));
    KG_CUDA(cudaMemcpyAsync(x_out + off, f->stage.x + off, bytes, cudaMemcpyDeviceToHost, f->copy_stream));
    KG_CUDA(cudaMemcpyAsync(y_out + off, f->stage.y + off, bytes, cudaMemcpyDeviceToHost, f->copy_stream));
    KG_CUDA(cudaMemcpyAsync(dx_out + off, f->stage.dx + off, bytes, cudaMemcpyDeviceToHost, f->copy_stream));
    KG_CUDA(cudaMemcpyAsync(dy_out + off, f->stage.dy + off, bytes, cudaMemcpyDeviceToHost, f->copy_stream));
  }
  KG_CUDA(cudaEventRecord(f->copy_done, f->copy_stream));
  KG_CUDA(cudaStreamWaitEvent(s, f->copy_done, 0));  // the handle's stream (and its timers) see the copies end
  return sync_check(f);
}

int kg_field2d_step_boids_host_ordered(kg_field2d* f, const KgBoidsParams* p, uint64_t n,
                                       const uint32_t* id_in, const float* x_in, const float* y_in,
                                       const float* dx_in, const float* dy_in, float* x_out, float* y_out,
                                       float* dx_out, float* dy_out) {
  KG_TRY(use(f));
  if (!p) return fail(KG_E_INVALID, "null params");
  if (n && (!x_in || !y_in || !dx_in || !dy_in || !x_out || !y_out || !dx_out || !dy_out))
    return fail(KG_E_INVALID, "null host array");
  if (n > f->capacity) return fail(KG_E_CAPACITY, "%llu agents exceed the capacity %llu", (unsigned long long)n,
                                   (unsigned long long)f->capacity);
  if (!id_in && n > 0xFFFFFFFEull) return fail(KG_E_INVALID, "implicit ids need n < 2^32 - 1");
  f->n_write = 0;
  f->n_read = 0;
  f->log_has_holes = false;
  cudaStream_t s = f->stream;
  KG_CUDA(cudaMemsetAsync(f->count, 0, (size_t)f->g.ncells * 4, s));
  if (n == 0) return sync_check(f);
  KG_TRY(ensure_stage(f));
  // 1. upload (ids only when the caller has its own), pack + histogram + validation in one kernel.
  //    Host arrays that lie back to back (x | y | dx | dy in one block) travel as ONE copy each way.
  SoA in = f->stage;
  in.x = f->stage4;
  in.y = in.x + n;
  in.dx = in.y + n;
  in.dy = in.dx + n;
  if (id_in) KG_CUDA(cudaMemcpyAsync(f->stage.id, id_in, n * 4, cudaMemcpyHostToDevice, s));
  else in.id = nullptr;
  if (y_in == x_in + n && dx_in == y_in + n && dy_in == dx_in + n) {
    KG_CUDA(cudaMemcpyAsync(in.x, x_in, n * 16, cudaMemcpyHostToDevice, s));
  } else {
    KG_CUDA(cudaMemcpyAsync(in.x, x_in, n * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(in.y, y_in, n * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(in.dx, dx_in, n * 4, cudaMemcpyHostToDevice, s));
    KG_CUDA(cudaMemcpyAsync(in.dy, dy_in, n * 4, cudaMemcpyHostToDevice, s));
  }
  LAUNCH(f, KG_K_MISC, pack_hist_kernel, blocks_for(n), kThreads, f->g, n, in, f->B, f->count, f->d_err);
  f->n_write = n;
  f->pending_new_ids = id_in != nullptr;
  // 2. lazy_update; the scatter (and the canonical in-bag sort) also record where every input ended up
  f->want_origin = true;
  const int rc = rebuild(f);
  f->want_origin = false;
  KG_TRY(rc);
  if (!id_in) {  // ids 0 .. n-1 are unique by construction: no bitmap pass (verify_ids) before K4
    KG_CUDA(cudaMemsetAsync(f->d_ids_dup, 0, sizeof(int), s));
    f->ids_unknown = false;
  }
  // 3. every agent's step, then each result goes to its agent's place in the input arrays
  KG_TRY(step_boids(f, *p));
  LAUNCH(f, KG_K_MISC, unpack_ordered_kernel, blocks_for(n), kThreads, n, f->B, (const uint32_t*)f->origin, in);
  if (y_out == x_out + n && dx_out == y_out + n && dy_out == dx_out + n) {
    KG_CUDA(cudaMemcpyAsync(x_out, in.x, n * 16, cudaMemcpyDeviceToHost, s));
  } else {
    KG_CUDA(cudaMemcpyAsync(x_out, in.x, n * 4, cudaMemcpyDeviceToHost, s));
    KG_CUDA(cudaMemcpyAsync(y_out, in.y, n * 4, cudaMemcpyDeviceToHost, s));
    KG_CUDA(cudaMemcpyAsync(dx_out, in.dx, n * 4, cudaMemcpyDeviceToHost, s));
    KG_CUDA(cudaMemcpyAsync(dy_out, in.dy, n * 4, cudaMemcpyDeviceToHost, s));
  }
  return sync_check(f);
}

int kg_field2d_reduce(kg_field2d* f, double* out) {
  KG_TRY(use(f));
  if (!out) return fail(KG_E_INVALID, "null out");
  for (int k = 0; k < kRedVals; ++k) out[k] = 0.0;
  if (f->n_read == 0) return KG_OK;
  return reduce_segments(f->A.pv, 1, f->n_read, &f->red, &f->red_bytes, out, f->stream);
}

int kg_field2d_run_boids_series(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps, uint64_t every,
                                double* out, uint64_t out_rows) {
  KG_TRY(use(f));
  if (!p || !out) return fail(KG_E_INVALID, "null argument");
  if (every == 0) return fail(KG_E_INVALID, "every == 0");
  const uint64_t rows = nsteps / every;
  if (out_rows < rows) return fail(KG_E_CAPACITY, "series needs %llu rows", (unsigned long long)rows);
  if (f->log_has_holes) return fail(KG_E_INVALID, "series over a dynamic population: use kg_field2d_reduce per step");
  // scratch: chunk partials of one reduction + every row of the series, all on the device until the end
  const uint32_t nchunk = (uint32_t)std::max<uint64_t>(1, (f->capacity + kRedChunk - 1) / kRedChunk);
  const size_t need = ((size_t)nchunk + rows) * kRedVals * sizeof(double);
  if (f->red_bytes < need) {
    if (f->red) cudaFree(f->red);
    f->red = nullptr;
    f->red_bytes = 0;
    KG_CUDA(cudaMalloc(&f->red, need));
    f->red_bytes = need;
  }
  double* series = f->red + (size_t)nchunk * kRedVals;
  KgBoidsParams q = *p;
  uint64_t row = 0;
  for (uint64_t i = 0; i < nsteps; ++i) {
    q.step = p->step + i;
    KG_TRY(step_boids(f, q));
    KG_TRY(rebuild(f));
    if ((i + 1) % every == 0) {
      if (f->n_read)
        KG_TRY(reduce_segments_dev(f->A.pv, 1, f->n_read, f->red, series + row * kRedVals, f->stream));
      else
        KG_CUDA(cudaMemsetAsync(series + row * kRedVals, 0, kRedVals * sizeof(double), f->stream));
      ++row;
    }
  }
  if (rows) KG_CUDA(cudaMemcpyAsync(out, series, rows * kRedVals * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
  return sync_check(f);
}

int kg_field2d_l2_flush(kg_field2d* f, uint64_t bytes) {
  KG_TRY(use(f));
  return f->flusher.run(bytes, f->stream);
}

int kg_field2d_run_boids_timed(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps,
                               uint64_t flush_bytes, double* ms_sum) {
  KG_TRY(use(f));
  if (!p || !ms_sum) return fail(KG_E_INVALID, "null argument");
  KgBoidsParams q = *p;
  for (uint64_t i = 0; i < nsteps; ++i) {
    cudaEvent_t a = nullptr, b = nullptr;
    KG_TRY(f->events.get(2 * i, &a));
    KG_TRY(f->events.get(2 * i + 1, &b));
    KG_TRY(f->flusher.run(flush_bytes, f->stream));
    q.step = p->step + i;
    KG_CUDA(cudaEventRecord(a, f->stream));
    KG_TRY(step_boids(f, q));
    KG_TRY(rebuild(f));
    KG_CUDA(cudaEventRecord(b, f->stream));
  }
  KG_TRY(sync_check(f));
  double sum = 0;
  for (uint64_t i = 0; i < nsteps; ++i) {
    float t = 0.f;
    KG_CUDA(cudaEventElapsedTime(&t, f->events.ev[2 * i], f->events.ev[2 * i + 1]));
    sum += t;
  }
  *ms_sum = sum;
  return KG_OK;
}

int kg_field2d_timer_start(kg_field2d* f) {
  KG_TRY(use(f));
  return f->watch.start(f->stream);
}
int kg_field2d_timer_stop(kg_field2d* f, double* ms) {
  KG_TRY(use(f));
  return f->watch.stop(f->stream, ms);
}

int kg_field2d_profile(kg_field2d* f, int enable) {
  KG_TRY(use(f));
  f->prof.drain();
  f->prof.enabled = enable != 0;
  return KG_OK;
}
int kg_field2d_profile_read(kg_field2d* f, double* ms, uint64_t* launches, int reset) {
  KG_TRY(use(f));
  KG_CUDA(cudaStreamSynchronize(f->stream));
  f->prof.drain();
  for (int k = 0; k < KG_K_COUNT; ++k) {
    if (ms) ms[k] = f->prof.ms[k];
    if (launches) launches[k] = f->prof.launches[k];
    if (reset) {
      f->prof.ms[k] = 0;
      f->prof.launches[k] = 0;
    }
  }
  return KG_OK;
}

}  // extern "C"
