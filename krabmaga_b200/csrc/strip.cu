// Multi-GPU Field2D: one x-strip of whole cell columns per GPU, per-step agent migration and
// halo exchange over NVLink peer memory.
//
// The reference's precedent is src/engine/fields/kdtree_mpi.rs (block per MPI rank, halo regions
// of width `distance`, two-phase count-then-payload exchange :705-790).  Here a step has ONE
// exchange, fused into the step's own kernels: K4 stages the agents that leave (migrants) and the
// agents that now sit in a boundary column (ghosts = the neighbour's next halo); one push kernel
// writes both straight into the neighbour GPU's inbox with peer stores and publishes count + epoch
// in one flag word behind a system-scope fence; the neighbour's append / halo-build kernels park
// on that flag (bounded spin).  No host round trip, no count phase, no NCCL on the data path.
// Parity double buffering of the inboxes is enough because neighbours can be at most one step apart.
//
// Layout per rank: cell columns [own_x0, own_x1) are owned (the last rank also owns the padding
// column max_x, F4); local columns = halo_l + owned + halo_r with local cell index
// (cx - x_off)*dh + cy.  Toroidal fields clamp the query window (F3), so halos exist only
// between adjacent strips (a line), while migration is a ring because toroidal_transform wraps
// positions (bird.rs:146).  The sorted read buffer is [left halo | owned | right halo] with the
// owned part starting at the fixed offset `hcap`, the left halo right-aligned before it.
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "boids_device.cuh"
#include "common.cuh"
#include "scan.cuh"

namespace kg {

enum : int {
  SERR_OOB = 1,           // coordinate outside the bag grid
  SERR_MIG_OVERFLOW = 2,  // more migrants in one step than the outbox holds
  SERR_MIG_FAR = 4,       // an agent jumped past the neighbouring strip
  SERR_HALO_OVERFLOW = 8, // boundary columns hold more agents than the halo inbox
  SERR_TIMEOUT = 16,      // a neighbour's flag did not arrive
  SERR_CAPACITY = 32      // strip holds more agents than its capacity
};

struct StripGeom {
  Geom g;              // global geometry; g.ncells = number of LOCAL cells
  int x_off;           // global column of local column 0
  int own_x0, own_x1;  // owned global columns
  int left_x0, left_x1, right_x0, right_x1;  // ring neighbours' owned columns (migration)
  int halo_l, halo_r;  // halo columns present on each side (0 or dd)
  int dd;              // halo width in columns = floor(radius / disc) the strip was built for
  int ncols;           // local columns
};

struct SlotHeader {
  // (epoch << 32) | count of the last completed push: ONE word, so the receiver needs one
  // system-scope ordering point per push, not one for the count and another for the flag
  unsigned long long flag;
  uint32_t pad[2];
};
__device__ __forceinline__ uint32_t slot_count(const SlotHeader* h) {
  return (uint32_t)*(const volatile unsigned long long*)&h->flag;
}

// one inbox slot (direction x parity): migration part and halo part
struct SlotPtrs {
  SlotHeader* mig_hdr;
  uint32_t* mig_id;
  float4* mig_pv;
  SlotHeader* halo_hdr;
  uint32_t* halo_cnt;  // per-cell counts of the dd*dh halo cells
  uint32_t* halo_id;
  float4* halo_pv;
};

struct SlotLayout {
  size_t mig_hdr, mig_id, mig_pv, halo_hdr, halo_cnt, halo_id, halo_pv, bytes;
};
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline SlotLayout make_layout(uint64_t mcap, uint64_t hcap, uint64_t halo_cells) {
  SlotLayout L;
  size_t o = 0;
  L.mig_hdr = o; o += 256;
  L.mig_id = o; o += align_up(mcap * 4, 256);
  L.mig_pv = o; o += align_up(mcap * 16, 256);
  L.halo_hdr = o; o += 256;
  L.halo_cnt = o; o += align_up((halo_cells + 16) * 4, 256);
  L.halo_id = o; o += align_up(hcap * 4, 256);
  L.halo_pv = o; o += align_up(hcap * 16, 256);
  L.bytes = o;
  return L;
}
inline SlotPtrs slot_ptrs(void* base, const SlotLayout& L, int dir, int parity) {
  char* p = (char*)base + (size_t)(dir * 2 + parity) * L.bytes;
  SlotPtrs s;
  s.mig_hdr = (SlotHeader*)(p + L.mig_hdr);
  s.mig_id = (uint32_t*)(p + L.mig_id);
  s.mig_pv = (float4*)(p + L.mig_pv);
  s.halo_hdr = (SlotHeader*)(p + L.halo_hdr);
  s.halo_cnt = (uint32_t*)(p + L.halo_cnt);
  s.halo_id = (uint32_t*)(p + L.halo_id);
  s.halo_pv = (float4*)(p + L.halo_pv);
  return s;
}

// device-resident bookkeeping of one strip
struct StripState {
  uint32_t n_owned;     // owned agents in the sorted buffer (from hcap)
  uint32_t n_log;       // entries in the write log
  uint32_t out_count[2];  // migrants staged for the left / right neighbour this step
  uint32_t gout_count[2]; // ghosts (my agents now in my first / last dd columns) staged for the line neighbours
  uint32_t gself_count[2];  // my migrants that landed in the left / right neighbour's boundary columns
  uint32_t push_done[4];  // completion counters of the multi-block pushes
  int err;
  uint32_t mig_in_total;  // statistics: migrants received so far
  uint32_t mig_out_total;
  uint32_t halo_in[2];    // last halo sizes
  int ids_dup;            // != 0: ids of [halo|owned|halo] not verified unique => K4 compares ids
  uint32_t k4_done;       // fused push: blocks of the step kernel that have finished (reset by the last one)
  uint32_t ghost_begin;   // fold mode: log entries from here on are ghosts (halo content), not owned agents
  uint32_t scan_sub;      // fold mode: ghosts of the left halo = how far the sorted buffer starts before hcap
};

__device__ __forceinline__ int global_col(const Geom& g, float x) {
  return f2i_sat(floorf(fdiv(x, g.disc)));
}
__device__ __forceinline__ bool local_cell(const StripGeom& sg, float x, float y, uint32_t* cell,
                                           int* col) {
  int cx = f2i_sat(floorf(fdiv(x, sg.g.disc)));
  int cy = f2i_sat(floorf(fdiv(y, sg.g.disc)));
  *col = cx;
  int lx = cx - sg.x_off;
  *cell = (uint32_t)(lx * sg.g.dh + cy);
  return lx >= 0 && lx < sg.ncols && cy >= 0 && cy < sg.g.dh;
}
__device__ __forceinline__ bool owns(const StripGeom& sg, int col) {
  return col >= sg.own_x0 && col < sg.own_x1;
}

// Philox init (state.rs:41-56): every rank walks all ids and keeps the agents it owns
__global__ void strip_init_kernel(StripGeom sg, uint64_t n_global, uint64_t seed, Agents log,
                                  uint64_t cap, uint32_t* __restrict__ count, StripState* st) {
  grid_dep_wait();
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_global) return;
  uint32_t id = (uint32_t)t;
  Philox4 r = philox4x32_10(id, 0, 0, DOMAIN_INIT, (uint32_t)seed, (uint32_t)(seed >> 32));
  float x = fmul(sg.g.w, u01_f32(r.v[0])), y = fmul(sg.g.h, u01_f32(r.v[1]));
  uint32_t c;
  int col;
  bool ok = local_cell(sg, x, y, &c, &col);
  if (!owns(sg, col)) return;
  if (!ok) {
    atomicOr(&st->err, SERR_OOB);
    return;
  }
  uint32_t slot = atomicAdd(&st->n_log, 1u);
  if (slot >= cap) {
    atomicOr(&st->err, SERR_CAPACITY);
    return;
  }
  log.id[slot] = id;
  log.pv[slot] = make_float4(x, y, 0.f, 0.f);
  atomicAdd(&count[c], 1u);
}

// upload path: entries [first, first+n) of the log were packed by the host call
__global__ void strip_hist_kernel(StripGeom sg, uint64_t first, uint64_t n, Agents log,
                                  uint32_t* __restrict__ count, StripState* st) {
  grid_dep_wait();
  uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= first + n) return;
  float4 q = log.pv[i];
  uint32_t c;
  int col;
  bool ok = local_cell(sg, q.x, q.y, &c, &col);
  if (ok && owns(sg, col))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(&st->err, SERR_OOB);
}
__global__ void strip_pack_kernel(uint64_t n, const uint32_t* id, const float* x, const float* y,
                                  const float* dx, const float* dy, Agents d, uint64_t off) {
  grid_dep_wait();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d.id[off + i] = id[i];
  d.pv[off + i] = make_float4(x[i], y[i], dx[i], dy[i]);
}

// K4 for a strip: the fast boids kernel over the owned agents, then classification of the new
// position: owned -> histogram; neighbour's -> staged in the outbox for that direction.
// staging areas of the one-exchange-per-step scheme (strip_step): ghosts for the neighbours and
// my own migrants that stay visible to me
struct GhostBufs {
  int on;
  Agents gout[2];
  Agents gself[2];
  // fused push (KG_STRIP_PUSH != kernel): migrants and ghosts for the neighbours are stored straight into
  // their inbox slots over NVLink by the step kernel itself, and the last block to finish publishes the
  // (epoch, count) flags — the transfer overlaps the boids arithmetic and the push launch disappears
  int fused;
  SlotPtrs peer[2];            // this step's slot in the left / right neighbour's inbox
  unsigned long long epoch;
};
// one slot of a staging list; the lanes of a warp that emit together share ONE atomic (the
// boundary columns are whole warps of emitters, all bumping the same counter)
__device__ __forceinline__ uint32_t take_slot(uint32_t* counter) {
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}
__device__ __forceinline__ void emit_ghost(Agents dst, uint32_t* counter, uint32_t cap, uint32_t id, float4 v,
                                           StripState* st) {
  const uint32_t slot = take_slot(counter);
  if (slot >= cap) {
    atomicOr(&st->err, SERR_HALO_OVERFLOW);
    return;
  }
  dst.id[slot] = id;
  dst.pv[slot] = v;
}
// the same into a neighbour's inbox (peer memory): a plain store over NVLink; publish_flags_kernel,
// launched behind the step kernel, orders it before the flag
__device__ __forceinline__ void emit_remote(uint32_t* rid, float4* rpv, uint32_t* counter, uint32_t cap, int errbit,
                                            uint32_t id, float4 v, StripState* st) {
  const uint32_t slot = take_slot(counter);
  if (slot >= cap) {
    atomicOr(&st->err, errbit);
    return;
  }
  rid[slot] = id;
  rpv[slot] = v;
}
// one agent of a strip's K4 (thread i of the step kernel)
template <bool EXACT>
__device__ __forceinline__ void strip_step_agent(const StripGeom& sg, const KgBoidsParams& p, float T, uint32_t hcap,
                                                 Agents rd, const uint32_t* __restrict__ cell_start, Agents log,
                                                 uint32_t* __restrict__ count, Agents out_l, Agents out_r,
                                                 uint32_t mcap, const GhostBufs& gx, StripState* st, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Geom& g = sg.g;
  const uint32_t a = hcap + i;
  const uint32_t id = rd.id[a];
  const ulonglong2 self = reinterpret_cast<const ulonglong2*>(rd.pv)[a];
  int col, ncy;
  const ulonglong2 outp = boids_step_packed<EXACT>(g, p, sg.dd, T, st->ids_dup != 0, a, id, self, sg.x_off,
                                                   cell_start, rd.id, rd.pv, &col, &ncy);
  const float4 out = *reinterpret_cast<const float4*>(&outp);
  log.id[i] = id;
  log.pv[i] = out;
  const int lx = col - sg.x_off;
  const uint32_t c = (uint32_t)(lx * g.dh + ncy);
  const bool ok = lx >= 0 && lx < sg.ncols && ncy >= 0 && ncy < g.dh;  // == local_cell()
  if (owns(sg, col)) {
    if (ok)
      atomicAdd(&count[c], 1u);
    else
      atomicOr(&st->err, SERR_OOB);
    // it stays mine; if it now sits in my first / last dd columns it is part of that line
    // neighbour's next halo: hand it over together with this step's migrants
    if (gx.on) {
      if (sg.halo_l > 0 && col < sg.own_x0 + sg.dd) {
        if (gx.fused)
          emit_remote(gx.peer[0].halo_id, gx.peer[0].halo_pv, &st->gout_count[0], hcap, SERR_HALO_OVERFLOW, id, out, st);
        else
          emit_ghost(gx.gout[0], &st->gout_count[0], hcap, id, out, st);
      }
      if (sg.halo_r > 0 && col >= sg.own_x1 - sg.dd) {
        if (gx.fused)
          emit_remote(gx.peer[1].halo_id, gx.peer[1].halo_pv, &st->gout_count[1], hcap, SERR_HALO_OVERFLOW, id, out, st);
        else
          emit_ghost(gx.gout[1], &st->gout_count[1], hcap, id, out, st);
      }
    }
    return;
  }
  // it leaves; if it lands in a line neighbour's boundary columns it is in MY next halo
  if (gx.on) {
    if (sg.halo_l > 0 && col >= sg.own_x0 - sg.dd && col < sg.own_x0)
      emit_ghost(gx.gself[0], &st->gself_count[0], mcap, id, out, st);
    if (sg.halo_r > 0 && col >= sg.own_x1 && col < sg.own_x1 + sg.dd)
      emit_ghost(gx.gself[1], &st->gself_count[1], mcap, id, out, st);
  }
  int dir;
  if (col >= sg.right_x0 && col < sg.right_x1)
    dir = 1;
  else if (col >= sg.left_x0 && col < sg.left_x1)
    dir = 0;
  else {
    atomicOr(&st->err, SERR_MIG_FAR);
    return;
  }
  if (gx.fused) {
    if (dir == 0)
      emit_remote(gx.peer[0].mig_id, gx.peer[0].mig_pv, &st->out_count[0], mcap, SERR_MIG_OVERFLOW, id, out, st);
    else
      emit_remote(gx.peer[1].mig_id, gx.peer[1].mig_pv, &st->out_count[1], mcap, SERR_MIG_OVERFLOW, id, out, st);
    return;
  }
  uint32_t slot;
  if (dir == 0)
    slot = take_slot(&st->out_count[0]);
  else
    slot = take_slot(&st->out_count[1]);
  if (slot >= mcap) {
    atomicOr(&st->err, SERR_MIG_OVERFLOW);
    return;
  }
  Agents o = dir ? out_r : out_l;
  o.id[slot] = id;
  o.pv[slot] = out;
}


// EXACT: get_neighbors_within_distance (the query the reference's own fixture calls, bird.rs:41) on
// the packed exact path (any window: columns walked in groups of three cells), T = exact_threshold(radius)
template <bool EXACT>
__global__ void __launch_bounds__(128, EXACT ? 6 : 10)
strip_step_kernel(StripGeom sg, KgBoidsParams p, float T, uint32_t hcap, Agents rd,
                  const uint32_t* __restrict__ cell_start, Agents log,
                  uint32_t* __restrict__ count, Agents out_l, Agents out_r, uint32_t mcap, GhostBufs gx,
                  StripState* st) {
  grid_dep_wait();
  const uint32_t n = st->n_owned;
  if (blockIdx.x * blockDim.x >= n) return;  // whole block beyond the population: not part of the count below
  strip_step_agent<EXACT>(sg, p, T, hcap, rd, cell_start, log, count, out_l, out_r, mcap, gx, st, n);
}

// Fused push, second half: the step kernel has stored this step's migrants and ghosts into the
// neighbours' inboxes; one thread publishes the (epoch, count) flags.  The kernel boundary orders the
// step kernel's peer stores before this thread's system-scope fence and flag stores.
__device__ __forceinline__ void wait_flag_thread(const SlotHeader* h, unsigned long long epoch, StripState* st);

__global__ void publish_flags_kernel(StripGeom sg, GhostBufs gx, uint32_t mcap, uint32_t hcap, StripState* st,
                                     SlotPtrs in_l, SlotPtrs in_r, int wait_in) {
  grid_dep_wait();
  __threadfence_system();
  for (int d = 0; d < 2; ++d) {
    const uint32_t nm = min(st->out_count[d], mcap);
    *(volatile unsigned long long*)&gx.peer[d].mig_hdr->flag = (gx.epoch << 32) | nm;
    const bool halo = d == 0 ? sg.halo_l > 0 : sg.halo_r > 0;
    if (halo) {
      const uint32_t ng = min(st->gout_count[d], hcap);
      *(volatile unsigned long long*)&gx.peer[d].halo_hdr->flag = (gx.epoch << 32) | ng;
      st->gout_count[d] = 0;
    }
    st->mig_out_total += nm;
    st->out_count[d] = 0;
  }
  // Having published, this one thread also waits for the neighbours' flags of the same step, so that the
  // append kernel behind it (hundreds of blocks) starts with the arrivals in place instead of every block
  // parking on four flags behind a system fence of its own.
  if (wait_in) {
    wait_flag_thread(in_l.mig_hdr, gx.epoch, st);
    wait_flag_thread(in_r.mig_hdr, gx.epoch, st);
    if (sg.halo_l > 0) wait_flag_thread(in_l.halo_hdr, gx.epoch, st);
    if (sg.halo_r > 0) wait_flag_thread(in_r.halo_hdr, gx.epoch, st);
    __threadfence_system();
  }
}

// A block parks on a neighbour's flag until it reaches `epoch` (bounded: ~4 s, then SERR_TIMEOUT).
__device__ __forceinline__ void wait_flag_block(const SlotHeader* h, unsigned long long epoch, StripState* st) {
  if (threadIdx.x == 0 && h != nullptr) {
    const volatile unsigned long long* f = &h->flag;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((*f >> 32) < epoch) {
      __nanosleep(100);
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 4000000000ull) {
        atomicOr(&st->err, SERR_TIMEOUT);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

__device__ __forceinline__ void wait_flag_thread(const SlotHeader* h, unsigned long long epoch, StripState* st) {
  const volatile unsigned long long* f = &h->flag;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while ((*f >> 32) < epoch) {
    __nanosleep(100);
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > 4000000000ull) {
      atomicOr(&st->err, SERR_TIMEOUT);
      break;
    }
  }
}

// Push the staged migrants of both directions (blockIdx.y: 0 = to the left ring neighbour, 1 = to
// the right) into the neighbours' inbox slots with peer stores; the last block of a direction
// publishes count and epoch flag behind a system-scope fence.
// The same launch carries the ghosts (gsrc: my agents that now sit in the boundary columns facing
// that neighbour, unsorted) into the halo part of the same slot; its flag is the halo header's.
struct PushMigArgs {
  Agents src[2];
  uint32_t* src_count[2];
  Agents gsrc[2];
  uint32_t* gsrc_count[2];  // nullptr: no ghosts in this direction (ring-only neighbour / prepare)
  SlotPtrs dst[2];
  uint32_t* done[2];
};
__global__ void push_migrants_kernel(PushMigArgs pa, uint32_t mcap, uint32_t hcap, unsigned long long epoch,
                                     StripState* st) {
  grid_dep_wait();
  const int d = blockIdx.y;
  const Agents src = pa.src[d];
  const SlotPtrs dst = pa.dst[d];
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t n = min(*pa.src_count[d], mcap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    dst.mig_id[i] = src.id[i];
    dst.mig_pv[i] = src.pv[i];
  }
  uint32_t ng = 0;
  if (pa.gsrc_count[d] != nullptr) {
    const Agents g = pa.gsrc[d];
    ng = min(*pa.gsrc_count[d], hcap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ng; i += stride) {
      dst.halo_id[i] = g.id[i];
      dst.halo_pv[i] = g.pv[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t prev = atomicAdd(pa.done[d], 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();  // every block fenced its stores before its atomic; order mine after them
      *(volatile unsigned long long*)&dst.mig_hdr->flag = (epoch << 32) | n;
      if (pa.gsrc_count[d] != nullptr) {
        *(volatile unsigned long long*)&dst.halo_hdr->flag = (epoch << 32) | ng;
        *pa.gsrc_count[d] = 0;
      }
      atomicAdd(&st->mig_out_total, n);
      *pa.src_count[d] = 0;
      *pa.done[d] = 0;
    }
  }
}

// Push the strip's boundary columns (each a contiguous slice of the sorted buffer) as the line
// neighbours' halos: agents, per-cell counts, then count + flag.  blockIdx.y: 0 = my first dd
// columns -> left neighbour, 1 = my last dd columns -> right neighbour.
struct PushHaloArgs {
  uint32_t first_cell[2];
  int have[2];
  SlotPtrs dst[2];
  uint32_t* done[2];
};
__global__ void push_halo_kernel(Agents a, const uint32_t* __restrict__ cell_start, PushHaloArgs pa,
                                 uint32_t ncells, uint32_t hcap, unsigned long long epoch, StripState* st) {
  grid_dep_wait();
  const int d = blockIdx.y;
  if (!pa.have[d]) return;
  const uint32_t first_cell = pa.first_cell[d];
  const SlotPtrs dst = pa.dst[d];
  const uint32_t s = cell_start[first_cell], e = cell_start[first_cell + ncells];
  uint32_t n = e - s;
  if (n > hcap) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&st->err, SERR_HALO_OVERFLOW);
    n = hcap;
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    dst.halo_id[i] = a.id[s + i];
    dst.halo_pv[i] = a.pv[s + i];
  }
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += stride)
    dst.halo_cnt[c] = cell_start[first_cell + c + 1] - cell_start[first_cell + c];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t prev = atomicAdd(pa.done[d], 1u);
    if (prev == gridDim.x - 1) {
      __threadfence_system();
      *(volatile unsigned long long*)&dst.halo_hdr->flag = (epoch << 32) | n;
      *pa.done[d] = 0;
    }
  }
}

// Append the migrants found in both inbox slots to the write log and histogram them.
// Every block first parks on the two neighbours' flags of this step (no separate wait launch).
__global__ void append_migrants_kernel(StripGeom sg, SlotPtrs in_l, SlotPtrs in_r, int have_l,
                                       int have_r, unsigned long long epoch, Agents log, uint64_t cap,
                                       uint32_t* __restrict__ count, StripState* st) {
  grid_dep_wait();
  wait_flag_block(have_l ? in_l.mig_hdr : nullptr, epoch, st);
  wait_flag_block(have_r ? in_r.mig_hdr : nullptr, epoch, st);
  const uint32_t nl = have_l ? slot_count(in_l.mig_hdr) : 0u;
  const uint32_t nr = have_r ? slot_count(in_r.mig_hdr) : 0u;
  const uint32_t base = st->n_owned;  // K4 wrote log[0, n_owned)
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    st->n_log = base + nl + nr;
    st->mig_in_total += nl + nr;
  }
  if (i >= nl + nr) return;
  uint32_t id;
  float4 q;
  if (i < nl) {
    id = __ldcg(&in_l.mig_id[i]);
    q = __ldcg(&in_l.mig_pv[i]);
  } else {
    id = __ldcg(&in_r.mig_id[i - nl]);
    q = __ldcg(&in_r.mig_pv[i - nl]);
  }
  if ((uint64_t)base + i >= cap) {
    atomicOr(&st->err, SERR_CAPACITY);
    return;
  }
  log.id[base + i] = id;
  log.pv[base + i] = q;
  uint32_t c;
  int col;
  bool ok = local_cell(sg, q.x, q.y, &c, &col);
  if (ok && owns(sg, col))
    atomicAdd(&count[c], 1u);
  else
    atomicOr(&st->err, SERR_OOB);
}
// leaving step mode (upload / prepare / clear after fold-mode steps): the log is empty, nothing in it is a ghost
__global__ void strip_reset_log_kernel(StripState* st) {
  st->n_log = 0;
  st->ghost_begin = 0xFFFFFFFFu;
}
__global__ void set_log_len_kernel(StripState* st) {
  grid_dep_wait();
  st->n_log = st->n_owned;
}

// Fold mode of the one-exchange step: the migrants AND the ghosts that make up both halos (arrived
// from the line neighbours, plus my own migrants that landed in their boundary columns) are appended
// to the write log and histogrammed into their (owned or halo) cells, so that the ONE scan + scatter
// of the rebuild sorts the halos into place together with the owned agents — no separate halo sort.
// Log layout: [K4 outputs | migrants l, r | left ghosts in, own | right ghosts in, own].
__global__ void append_all_kernel(StripGeom sg, SlotPtrs in_l, SlotPtrs in_r, unsigned long long epoch,
                                  Agents gself_l, Agents gself_r, uint32_t hcap, Agents log, uint64_t cap,
                                  uint32_t* __restrict__ count, StripState* st, int prewaited) {
  grid_dep_wait();
  if (!prewaited) {  // otherwise publish_flags_kernel, launched just before, has seen all four flags
    wait_flag_block(in_l.mig_hdr, epoch, st);
    wait_flag_block(in_r.mig_hdr, epoch, st);
    wait_flag_block(sg.halo_l > 0 ? in_l.halo_hdr : nullptr, epoch, st);
    wait_flag_block(sg.halo_r > 0 ? in_r.halo_hdr : nullptr, epoch, st);
  }
  const uint32_t nl = slot_count(in_l.mig_hdr), nr = slot_count(in_r.mig_hdr);
  uint32_t gl = sg.halo_l > 0 ? min(slot_count(in_l.halo_hdr), hcap) : 0u;
  uint32_t gr = sg.halo_r > 0 ? min(slot_count(in_r.halo_hdr), hcap) : 0u;
  uint32_t sl = sg.halo_l > 0 ? st->gself_count[0] : 0u, sr = sg.halo_r > 0 ? st->gself_count[1] : 0u;
  bool overflow = false;
  if (gl + sl > hcap) { sl = hcap - gl; overflow = true; }
  if (gr + sr > hcap) { sr = hcap - gr; overflow = true; }
  const uint32_t base = st->n_owned;  // K4 wrote log[0, n_owned)
  const uint32_t nmig = nl + nr, nall = nmig + gl + sl + gr + sr;
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    st->n_log = base + nall;
    st->ghost_begin = base + nmig;
    st->scan_sub = gl + sl;
    st->mig_in_total += nmig;
    if (overflow) atomicOr(&st->err, SERR_HALO_OVERFLOW);
  }
  if (i >= nall) return;
  uint32_t id;
  float4 q;
  bool ghost = true;
  uint32_t k = i;
  if (k < nl) {
    id = __ldcg(&in_l.mig_id[k]); q = __ldcg(&in_l.mig_pv[k]); ghost = false;
  } else if ((k -= nl) < nr) {
    id = __ldcg(&in_r.mig_id[k]); q = __ldcg(&in_r.mig_pv[k]); ghost = false;
  } else if ((k -= nr) < gl) {
    id = __ldcg(&in_l.halo_id[k]); q = __ldcg(&in_l.halo_pv[k]);
  } else if ((k -= gl) < sl) {
    id = gself_l.id[k]; q = gself_l.pv[k];
  } else if ((k -= sl) < gr) {
    id = __ldcg(&in_r.halo_id[k]); q = __ldcg(&in_r.halo_pv[k]);
  } else {
    k -= gr;
    id = gself_r.id[k]; q = gself_r.pv[k];
  }
  if ((uint64_t)base + i >= cap) {
    atomicOr(&st->err, SERR_CAPACITY);
    return;
  }
  log.id[base + i] = id;
  log.pv[base + i] = q;
  uint32_t c;
  int col;
  const bool ok = local_cell(sg, q.x, q.y, &c, &col);
  if (ok && owns(sg, col) != ghost)   // a migrant lands in my columns, a ghost in a halo column
    atomicAdd(&count[c], 1u);
  else
    atomicOr(&st->err, SERR_OOB);
}
// after the fold-mode scatter: the bookkeeping halo_build_kernel used to leave behind
__global__ void strip_finish_kernel(StripGeom sg, uint32_t hcap, const uint32_t* __restrict__ cell_start,
                                    StripState* st) {
  grid_dep_wait();
  const uint32_t own_end = (uint32_t)((sg.halo_l + (sg.own_x1 - sg.own_x0)) * sg.g.dh);
  const uint32_t n_owned = cell_start[own_end] - hcap;
  st->n_owned = n_owned;
  st->n_log = 0;
  st->halo_in[0] = hcap - cell_start[0];
  st->halo_in[1] = cell_start[sg.g.ncells] - (hcap + n_owned);
  st->gself_count[0] = 0;
  st->gself_count[1] = 0;
  st->ghost_begin = 0xFFFFFFFFu;
}

// K3 for a strip: owned entries of the log go to their cell slot, migrants that left are skipped
__global__ void __launch_bounds__(256)
strip_scatter_kernel(StripGeom sg, Agents src, Agents dst, const uint32_t* __restrict__ cell_start,
                     uint32_t* __restrict__ count, StripState* st, uint32_t hcap, int finish) {
  grid_dep_wait();
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (finish && i == 0) {
    // fold mode: the step's bookkeeping (what strip_finish_kernel did in a launch of its own).  The scan is
    // complete, nobody in this kernel reads these words.  n_log and ghost_begin ARE read here, so they keep
    // this step's values until the next append overwrites them (strip_reset_log when stepping stops).
    const uint32_t own_end = (uint32_t)((sg.halo_l + (sg.own_x1 - sg.own_x0)) * sg.g.dh);
    const uint32_t n_owned = cell_start[own_end] - hcap;
    st->n_owned = n_owned;
    st->halo_in[0] = hcap - cell_start[0];
    st->halo_in[1] = cell_start[sg.g.ncells] - (hcap + n_owned);
    st->gself_count[0] = 0;
    st->gself_count[1] = 0;
  }
  if (i >= st->n_log) return;
  float4 q = src.pv[i];
  uint32_t id = src.id[i];
  uint32_t c;
  int col;
  bool ok = local_cell(sg, q.x, q.y, &c, &col);
  // owned agents go to their cell; log entries from ghost_begin on (fold mode) are halo content
  if (!ok || owns(sg, col) == (i >= st->ghost_begin)) return;
  uint32_t rank = atomicSub(&count[c], 1u) - 1u;
  uint32_t d = cell_start[c] + rank;
  dst.id[d] = id;
  dst.pv[d] = q;
}

__global__ void strip_sort_cells_kernel(uint32_t first, uint32_t ncells,
                                        const uint32_t* __restrict__ cs, Agents a) {
  grid_dep_wait();
  uint32_t c = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= first + ncells) return;
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

// After the scatter: record n_owned; place the received halos around the owned block and give the
// halo cells their cell_start entries.  blockIdx.y = side (0 left, 1 right); block (0, side) scans
// that side's per-cell counts, every block copies a share of the agents.
__global__ void __launch_bounds__(256)
unpack_halo_kernel(StripGeom sg, uint32_t hcap, SlotPtrs in_l, SlotPtrs in_r, unsigned long long epoch,
                   Agents a, uint32_t* cell_start, StripState* st) {
  grid_dep_wait();
  {  // park on this side's flag of the current rebuild (no separate wait launch)
    const int sd = blockIdx.y;
    const bool hv = sd == 0 ? sg.halo_l > 0 : sg.halo_r > 0;
    wait_flag_block(hv ? (sd == 0 ? in_l.halo_hdr : in_r.halo_hdr) : nullptr, epoch, st);
  }
  const int side = blockIdx.y;
  const int dh = sg.g.dh;
  const uint32_t own_end = (uint32_t)((sg.halo_l + (sg.own_x1 - sg.own_x0)) * dh);
  const uint32_t n_owned = cell_start[own_end] - hcap;
  const int have = side == 0 ? sg.halo_l > 0 : sg.halo_r > 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && side == 0) {
    st->n_owned = n_owned;
    st->n_log = 0;
  }
  const uint32_t ncells = (uint32_t)(sg.dd * dh);
  if (!have) {
    // no halo on this side: the right edge still needs its closing cell_start entry
    if (side == 1 && blockIdx.x == 0 && threadIdx.x == 0) {
      st->halo_in[1] = 0;
    }
    if (side == 0 && blockIdx.x == 0 && threadIdx.x == 0) st->halo_in[0] = 0;
    return;
  }
  const SlotPtrs& in = side == 0 ? in_l : in_r;
  uint32_t n = slot_count(in.halo_hdr);
  if (n > hcap) n = hcap;
  const uint32_t dst0 = side == 0 ? hcap - n : hcap + n_owned;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    a.id[dst0 + i] = __ldcg(&in.halo_id[i]);
    a.pv[dst0 + i] = __ldcg(&in.halo_pv[i]);
  }
  if (blockIdx.x == 0) {
    // exclusive scan of the per-cell counts, chunked by the block
    const uint32_t cell0 = side == 0 ? 0u : own_end;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < ncells; b0 += blockDim.x) {
      uint32_t c = b0 + threadIdx.x;
      uint32_t v = c < ncells ? __ldcg(&in.halo_cnt[c]) : 0u;
      uint32_t total;
      uint32_t ex = block_excl_scan(v, &total);
      if (c < ncells) cell_start[cell0 + c] = dst0 + carry + ex;
      carry += total;
    }
    if (threadIdx.x == 0) {
      if (side == 1) cell_start[own_end + ncells] = dst0 + carry;  // closing entry
      st->halo_in[side] = n;
    }
  }
}

// One-exchange scheme: build a side's halo from the UNSORTED ghosts that arrived with this step's
// migrants (inbox) plus my own migrants that landed in that neighbour's boundary columns (gself):
// a counting sort by halo cell into the halo region next to the owned block, and the halo cells'
// cell_start entries.  One block per side (blockIdx.x = side); `hist`/`cursor` are [2][ncells+1]
// scratch words, left zeroed for the next step.
__global__ void __launch_bounds__(1024)
halo_build_kernel(StripGeom sg, uint32_t hcap, SlotPtrs in_l, SlotPtrs in_r, unsigned long long epoch,
                  Agents gself_l, Agents gself_r, Agents a, uint32_t* cell_start, uint32_t* hist_g,
                  uint32_t* cursor_g, int use_smem, StripState* st) {
  extern __shared__ uint32_t hb_smem[];  // [2][ncells] when it fits (use_smem), else the global scratch
  __shared__ uint32_t s_carry;
  grid_dep_wait();
  const int side = blockIdx.x;
  const int dh = sg.g.dh;
  const bool have = side == 0 ? sg.halo_l > 0 : sg.halo_r > 0;
  const SlotPtrs& in = side == 0 ? in_l : in_r;
  wait_flag_block(have ? in.halo_hdr : nullptr, epoch, st);
  const uint32_t own_end = (uint32_t)((sg.halo_l + (sg.own_x1 - sg.own_x0)) * dh);
  const uint32_t n_owned = cell_start[own_end] - hcap;
  if (side == 0 && threadIdx.x == 0) {
    st->n_owned = n_owned;
    st->n_log = 0;
  }
  if (!have) {
    if (threadIdx.x == 0) st->halo_in[side] = 0;
    return;
  }
  const uint32_t ncells = (uint32_t)(sg.dd * dh);
  const Agents gs = side == 0 ? gself_l : gself_r;
  uint32_t nin = slot_count(in.halo_hdr);
  if (nin > hcap) nin = hcap;
  uint32_t ns = st->gself_count[side];
  uint32_t n = nin + ns;
  if (n > hcap) {
    if (threadIdx.x == 0) atomicOr(&st->err, SERR_HALO_OVERFLOW);
    ns = hcap - nin;
    n = hcap;
  }
  uint32_t* h = use_smem ? hb_smem : hist_g + (size_t)side * (ncells + 1);
  uint32_t* cur = use_smem ? hb_smem + ncells : cursor_g + (size_t)side * (ncells + 1);
  if (use_smem) {
    for (uint32_t c = threadIdx.x; c < ncells; c += blockDim.x) h[c] = 0;
    __syncthreads();
  }
  // first global column of this halo, and the index of its first cell in cell_start
  const int col0 = side == 0 ? sg.own_x0 - sg.dd : sg.own_x1;
  const uint32_t cell0 = side == 0 ? 0u : own_end;
  auto ghost = [&](uint32_t i, uint32_t* id, float4* pv) {
    if (i < nin) {
      *id = __ldcg(&in.halo_id[i]);
      *pv = __ldcg(&in.halo_pv[i]);
    } else {
      *id = gs.id[i - nin];
      *pv = gs.pv[i - nin];
    }
  };
  auto cell_of = [&](float4 pv) -> uint32_t {
    const int cx = f2i_sat(floorf(fdiv(pv.x, sg.g.disc))), cy = f2i_sat(floorf(fdiv(pv.y, sg.g.disc)));
    const int lx = cx - col0;
    if (lx < 0 || lx >= sg.dd || cy < 0 || cy >= dh) return 0xFFFFFFFFu;
    return (uint32_t)(lx * dh + cy);
  };
  // 1. histogram (the first kHbKeep ghosts of a thread stay in registers for the scatter)
  constexpr int kHbKeep = 8;
  uint32_t kid[kHbKeep], kcell[kHbKeep];
  float4 kpv[kHbKeep];
#pragma unroll
  for (int k = 0; k < kHbKeep; ++k) {
    const uint32_t i = threadIdx.x + (uint32_t)k * blockDim.x;
    kcell[k] = 0xFFFFFFFFu;
    if (i < n) {
      ghost(i, &kid[k], &kpv[k]);
      kcell[k] = cell_of(kpv[k]);
      if (kcell[k] == 0xFFFFFFFFu)
        atomicOr(&st->err, SERR_OOB);
      else
        atomicAdd(&h[kcell[k]], 1u);
    }
  }
  for (uint32_t i = threadIdx.x + kHbKeep * blockDim.x; i < n; i += blockDim.x) {
    uint32_t id;
    float4 pv;
    ghost(i, &id, &pv);
    const uint32_t c = cell_of(pv);
    if (c == 0xFFFFFFFFu)
      atomicOr(&st->err, SERR_OOB);
    else
      atomicAdd(&h[c], 1u);
  }
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  // 2. exclusive scan of the cell counts -> cell_start of the halo cells and scatter cursors
  const uint32_t dst0 = side == 0 ? hcap - n : hcap + n_owned;
  for (uint32_t b0 = 0; b0 < ncells; b0 += blockDim.x) {
    const uint32_t c = b0 + threadIdx.x;
    const uint32_t v = c < ncells ? h[c] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan_t<1024>(v, &total);
    const uint32_t carry = s_carry;
    if (c < ncells) {
      cell_start[cell0 + c] = dst0 + carry + ex;
      cur[c] = dst0 + carry + ex;
      if (!use_smem) h[c] = 0;  // the global scratch stays zero between steps
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (side == 1) cell_start[own_end + ncells] = dst0 + s_carry;  // closing entry
    st->halo_in[side] = n;
    st->gself_count[side] = 0;
  }
  // 3. scatter
#pragma unroll
  for (int k = 0; k < kHbKeep; ++k) {
    if (kcell[k] == 0xFFFFFFFFu) continue;
    const uint32_t pos = atomicAdd(&cur[kcell[k]], 1u);
    a.id[pos] = kid[k];
    a.pv[pos] = kpv[k];
  }
  for (uint32_t i = threadIdx.x + kHbKeep * blockDim.x; i < n; i += blockDim.x) {
    uint32_t id;
    float4 pv;
    ghost(i, &id, &pv);
    const uint32_t c = cell_of(pv);
    if (c == 0xFFFFFFFFu) continue;
    const uint32_t pos = atomicAdd(&cur[c], 1u);
    a.id[pos] = id;
    a.pv[pos] = pv;
  }
}

// id-uniqueness check over everything K4 can see: [left halo | owned | right halo]
__global__ void strip_ids_mark_kernel(uint32_t hcap, const uint32_t* __restrict__ ids, uint64_t nbits,
                                      uint32_t* __restrict__ bitmap, StripState* st) {
  grid_dep_wait();
  const uint32_t lo = hcap - st->halo_in[0], hi = hcap + st->n_owned + st->halo_in[1];
  uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const uint32_t id = ids[i], bit = 1u << (id & 31);
  if ((uint64_t)id >= nbits) {  // cannot verify: fall back to the id comparison
    st->ids_dup = 1;
    return;
  }
  if (atomicOr(&bitmap[id >> 5], bit) & bit) st->ids_dup = 1;
}
__global__ void strip_ids_reset_kernel(StripState* st) {
  grid_dep_wait();
  st->ids_dup = 0;
}

__global__ void strip_unpack_kernel(uint64_t n_cap, uint32_t hcap, Agents a, const StripState* st,
                                    uint32_t* id, float* x, float* y, float* dx, float* dy) {
  grid_dep_wait();
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st->n_owned || i >= n_cap) return;
  float4 q = a.pv[hcap + i];
  id[i] = a.id[hcap + i];
  x[i] = q.x;
  y[i] = q.y;
  dx[i] = q.z;
  dy[i] = q.w;
}

}  // namespace kg

// ====================================================================== handle + C ABI
using namespace kg;

struct kg_strip {
  bool log_dirty = false;  // fold-mode steps leave n_log / ghost_begin of the last step behind (strip_reset_log)
  int device = 0, rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  StripGeom sg{};
  uint64_t capacity = 0;  // owned agents
  uint32_t hcap = 0;      // halo agents per side
  uint32_t mcap = 0;      // migrants per direction per step
  Agents A, B;            // A: [hcap | capacity | hcap], B: log [capacity]
  Agents out[2];          // staged migrants (left, right)
  Agents gout[2];         // staged ghosts for the left / right line neighbour (hcap each)
  Agents gself[2];        // my migrants visible in my left / right halo (mcap each)
  uint32_t* halo_hist = nullptr;    // [2][dd*dh + 1] scratch of halo_build_kernel (kept zero)
  uint32_t* halo_cursor = nullptr;  // [2][dd*dh + 1]
  bool halo_smem_optin = false;
  uint32_t* cell_start = nullptr;
  uint32_t* count = nullptr;
  LookbackState scan;
  StripState* st = nullptr;  // device
  StripState* h_st = nullptr;  // pinned mirror
  SlotLayout layout{};
  void* inbox = nullptr;       // my inbox block: 4 slots (from_left/from_right x parity)
  void* peer_inbox[2] = {nullptr, nullptr};  // neighbours' inbox blocks (left, right)
  bool peer_is_ipc[2] = {false, false};
  int left_rank = -1, right_rank = -1;  // ring neighbours (-1: none)
  unsigned long long xchg_epoch = 0;  // one per exchange (prepare's halo push, each step's push); parity = epoch & 1
  uint64_t steps_done = 0;
  int order = KG_ORDER_ANY;
  bool prepared = false;
  SoA stage;
  bool have_stage = false;
  Stopwatch watch;
  L2Flusher flusher;
  EventPool events;
  uint64_t launches = 0;
  // optional per-kernel device timing (KG_STRIP_PROF=1): CUDA events around every launch
  uint64_t log_cap = 0;  // entries the write log B can hold
  bool fold = true;      // halos sorted by the rebuild's own scan + scatter (KG_STRIP_HALO=build: separate halo sort)
  bool fused_push = true;  // K4 stores migrants / ghosts into the peers' inboxes itself (KG_STRIP_PUSH=kernel: own launch)
  bool prof_on = false;
  std::vector<std::pair<const char*, std::pair<cudaEvent_t, cudaEvent_t>>> prof_ev;
  // ids written by kg_strip_init_flockers are unique by construction; uploaded ids are verified
  // after every rebuild (halos and migrants may bring a duplicate next to its twin at any step)
  bool ids_trusted = true;
  uint32_t* id_bitmap = nullptr;
  uint64_t id_bitmap_bits = 0;
};

namespace {

constexpr int kT = 256;
inline unsigned nblk(uint64_t n, int t = kT) { return (unsigned)std::max<uint64_t>(1, (n + t - 1) / t); }

int suse(kg_strip* s) {
  if (!s) return fail(KG_E_INVALID, "null strip handle");
  KG_CUDA(cudaSetDevice(s->device));
  return KG_OK;
}
// every kernel of the step chain is a programmatic dependent of its predecessor (common.cuh):
// each one starts with grid_dep_wait(), so the ~11 launches of a strip step overlap their launch
// latency with the previous kernel's tail
#define SLAUNCH(s, kernel, grid, block, ...)                                              \
  do {                                                                                    \
    cudaEvent_t _e0 = nullptr, _e1 = nullptr;                                             \
    if ((s)->prof_on) {                                                                   \
      cudaEventCreate(&_e0);                                                              \
      cudaEventCreate(&_e1);                                                              \
      cudaEventRecord(_e0, (s)->stream);                                                  \
    }                                                                                     \
    cudaError_t _le = launch_pdl(kernel, dim3(grid), dim3(block), (s)->stream, __VA_ARGS__); \
    if ((s)->prof_on) {                                                                   \
      cudaEventRecord(_e1, (s)->stream);                                                  \
      (s)->prof_ev.push_back({#kernel, {_e0, _e1}});                                      \
    }                                                                                     \
    if (_le != cudaSuccess)                                                               \
      return fail(KG_E_CUDA, "launch of %s failed: %s", #kernel, cudaGetErrorString(_le)); \
    launch_counter().fetch_add(1, std::memory_order_relaxed);                             \
    (s)->launches += 1;                                                                   \
  } while (0)

int strip_reset_log(kg_strip* s) {
  if (!s->log_dirty) return KG_OK;
  strip_reset_log_kernel<<<1, 1, 0, s->stream>>>(s->st);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  s->log_dirty = false;
  return KG_OK;
}
int strip_sync_check(kg_strip* s) {
  KG_CUDA(cudaMemcpyAsync(s->h_st, s->st, sizeof(StripState), cudaMemcpyDeviceToHost, s->stream));
  KG_CUDA(cudaStreamSynchronize(s->stream));
  int e = s->h_st->err;
  if (e) {
    KG_CUDA(cudaMemsetAsync(&s->st->err, 0, sizeof(int), s->stream));
    if (e & SERR_TIMEOUT) return fail(KG_E_CUDA, "strip %d: neighbour flag timed out (peer not stepping?)", s->rank);
    if (e & SERR_MIG_OVERFLOW) return fail(KG_E_CAPACITY, "strip %d: migration outbox overflow", s->rank);
    if (e & SERR_HALO_OVERFLOW) return fail(KG_E_CAPACITY, "strip %d: halo inbox overflow", s->rank);
    if (e & SERR_CAPACITY) return fail(KG_E_CAPACITY, "strip %d: more agents than capacity", s->rank);
    if (e & SERR_MIG_FAR) return fail(KG_E_INVALID, "strip %d: an agent moved past the neighbouring strip", s->rank);
    return fail(KG_E_OOB, "strip %d: agent coordinate outside the bag grid", s->rank);
  }
  return KG_OK;
}

int alloc_agents_n(Agents& a, uint64_t n) {
  KG_CUDA(cudaMalloc(&a.id, (n + 64) * 4));
  KG_CUDA(cudaMalloc(&a.pv, (n + 64) * 16));
  return KG_OK;
}

// CUDA loads kernels lazily and the first launch of a function may wait for the device to go
// idle — which never happens while a neighbour strip of the same process (or this strip's own
// stream) is parked on a neighbour's flag.  Touch every kernel of the step path up front.
int preload_kernels() {
  cudaFuncAttributes a;
  KG_CUDA(cudaFuncGetAttributes(&a, strip_init_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_hist_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_pack_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_step_kernel<false>));
  KG_CUDA(cudaFuncGetAttributes(&a, push_migrants_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, push_halo_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, append_migrants_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, set_log_len_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_scatter_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_sort_cells_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, unpack_halo_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, halo_build_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, append_all_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, publish_flags_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_finish_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_unpack_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_ids_reset_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, strip_ids_mark_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, scan_lookback_kernel));
  KG_CUDA(cudaFuncGetAttributes(&a, l2_flush_kernel));
  return KG_OK;
}

// neighbours' owned column range
void col_range(const kg_strip* s, int r, int* x0, int* x1) {
  int max_x = s->sg.g.max_x, G = s->nranks;
  *x0 = (int)((int64_t)r * max_x / G);
  *x1 = r == G - 1 ? s->sg.g.dw : (int)((int64_t)(r + 1) * max_x / G);
}

// lazy_update of a strip: scan -> scatter -> halos.
//   from_step = false (kg_strip_prepare): the boundary columns are pushed to the line neighbours
//     and theirs awaited — an exchange of its own (epoch `xchg_epoch` + 1);
//   from_step = true: the halos' content already arrived with this step's migrants (ghosts,
//     epoch `epoch`) and is only sorted into place — the step has ONE exchange, not two.
int strip_rebuild(kg_strip* s, bool from_step, unsigned long long epoch) {
  const StripGeom& sg = s->sg;
  const uint32_t own_cols = (uint32_t)(sg.own_x1 - sg.own_x0);
  const bool fold = from_step && s->fold && s->nranks > 1;
  // fold mode: the left halo's ghosts are part of this sort, so the sorted buffer starts that many
  // entries before hcap and the owned block still begins exactly at hcap
  exclusive_scan_lookback(s->scan, s->count, sg.g.ncells, s->cell_start, s->stream, s->hcap, true,
                          fold ? &s->st->scan_sub : nullptr);
  launch_counter().fetch_add(1, std::memory_order_relaxed);
  s->launches += 1;
  static const int fused_finish_env = getenv("KG_STRIP_FINISH") ? atoi(getenv("KG_STRIP_FINISH")) : 1;  // 0: own launch (A/B)
  const int fused_finish = fold && fused_finish_env ? 1 : 0;
  SLAUNCH(s, strip_scatter_kernel, nblk(fold ? s->log_cap : s->capacity), kT, sg, s->B, s->A, s->cell_start,
          s->count, s->st, s->hcap, fused_finish);
  if (fused_finish) s->log_dirty = true;
  if (!from_step) {
    if (s->order == KG_ORDER_CANONICAL)
      SLAUNCH(s, strip_sort_cells_kernel, nblk((uint64_t)own_cols * sg.g.dh, 128), 128,
              (uint32_t)(sg.halo_l * sg.g.dh), own_cols * (uint32_t)sg.g.dh, s->cell_start, s->A);
    s->xchg_epoch += 1;
    epoch = s->xchg_epoch;
  }
  const int parity = (int)(epoch & 1);
  SlotPtrs in_l = slot_ptrs(s->inbox, s->layout, 0, parity);
  SlotPtrs in_r = slot_ptrs(s->inbox, s->layout, 1, parity);
  if (!from_step) {
    const uint32_t hcells = (uint32_t)(sg.dd * sg.g.dh);
    // my first dd columns -> left neighbour's "from right" slot; my last dd -> right's "from left"
    if (sg.halo_l > 0 || sg.halo_r > 0) {
      PushHaloArgs pa{};
      pa.have[0] = sg.halo_l > 0;
      pa.have[1] = sg.halo_r > 0;
      pa.first_cell[0] = (uint32_t)(sg.halo_l * sg.g.dh);
      pa.first_cell[1] = (uint32_t)((sg.halo_l + (int)own_cols - sg.dd) * sg.g.dh);
      if (pa.have[0]) pa.dst[0] = slot_ptrs(s->peer_inbox[0], s->layout, 1, parity);
      if (pa.have[1]) pa.dst[1] = slot_ptrs(s->peer_inbox[1], s->layout, 0, parity);
      pa.done[0] = &s->st->push_done[2];
      pa.done[1] = &s->st->push_done[3];
      SLAUNCH(s, push_halo_kernel, dim3(64, 2), kT, s->A, (const uint32_t*)s->cell_start, pa, hcells, s->hcap,
              epoch, s->st);
    }
    dim3 grid(16, 2);
    SLAUNCH(s, unpack_halo_kernel, grid, kT, sg, s->hcap, in_l, in_r, epoch, s->A, s->cell_start, s->st);
  } else if (fold) {
    if (!fused_finish) SLAUNCH(s, strip_finish_kernel, 1, 1, sg, s->hcap, (const uint32_t*)s->cell_start, s->st);
    if (s->order == KG_ORDER_CANONICAL)
      SLAUNCH(s, strip_sort_cells_kernel, nblk(sg.g.ncells, 128), 128, 0u, sg.g.ncells, s->cell_start, s->A);
  } else {
    {
      // histogram + cursors of one side in shared memory when they fit (they do for every
      // BASELINE geometry: 2 x dd*dh words), else in the global scratch
      const size_t need = 2 * (size_t)sg.dd * sg.g.dh * 4;
      const int use_smem = need <= 160 * 1024;
      if (use_smem && need > 40 * 1024 && !s->halo_smem_optin) {
        KG_CUDA(cudaFuncSetAttribute(halo_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     160 * 1024));
        s->halo_smem_optin = true;
      }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(1024);
      cfg.dynamicSmemBytes = use_smem ? need : 0;
      cfg.stream = s->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      cudaError_t le = cudaLaunchKernelEx(&cfg, halo_build_kernel, sg, s->hcap, in_l, in_r, epoch, s->gself[0],
                                          s->gself[1], s->A, s->cell_start, s->halo_hist, s->halo_cursor,
                                          use_smem, s->st);
      if (le != cudaSuccess) return fail(KG_E_CUDA, "launch of halo_build_kernel failed: %s", cudaGetErrorString(le));
      launch_counter().fetch_add(1, std::memory_order_relaxed);
      s->launches += 1;
    }
    // ghosts arrive unsorted: in canonical order every local cell, halo cells included, is sorted by id
    if (s->order == KG_ORDER_CANONICAL)
      SLAUNCH(s, strip_sort_cells_kernel, nblk(sg.g.ncells, 128), 128, 0u, sg.g.ncells, s->cell_start, s->A);
  }
  if (!s->ids_trusted) {
    const uint64_t span = s->capacity + 2ull * s->hcap;
    KG_CUDA(cudaMemsetAsync(s->id_bitmap, 0, s->id_bitmap_bits / 8, s->stream));
    SLAUNCH(s, strip_ids_reset_kernel, 1, 1, s->st);
    SLAUNCH(s, strip_ids_mark_kernel, nblk(span), kT, s->hcap, s->A.id, s->id_bitmap_bits, s->id_bitmap,
            s->st);
  }
  return KG_OK;
}

int strip_step(kg_strip* s, const KgBoidsParams& p) {
  const StripGeom& sg = s->sg;
  if (!s->prepared) return fail(KG_E_INVALID, "strip not prepared (call kg_strip_prepare on every rank)");
  if (!(p.radius > 0.f)) return fail(KG_E_INVALID, "strips need a positive query radius");
  int dd = (int)floorf(p.radius / sg.g.disc);
  if (dd != sg.dd) return fail(KG_E_INVALID, "radius gives a %d-column window, strip was built for %d", dd, sg.dd);
  if (p.exact_query && !(p.radius < 3.0e38f)) return fail(KG_E_INVALID, "strips need a finite query radius");
  const bool ring = s->nranks > 1;
  // ONE exchange per step: migrants (ring) and the ghosts that make up the neighbours' next halos (line)
  s->xchg_epoch += 1;
  const unsigned long long epoch = s->xchg_epoch;
  const int parity = (int)(epoch & 1);
  GhostBufs gx{};
  gx.on = ring ? 1 : 0;
  for (int k = 0; k < 2; ++k) {
    gx.gout[k] = s->gout[k];
    gx.gself[k] = s->gself[k];
  }
  gx.fused = ring && s->fused_push ? 1 : 0;
  if (ring) {
    gx.peer[0] = slot_ptrs(s->peer_inbox[0], s->layout, 1, parity);  // my slot "from the right" in the left neighbour
    gx.peer[1] = slot_ptrs(s->peer_inbox[1], s->layout, 0, parity);
    gx.epoch = epoch;
  }
  if (p.exact_query)
    SLAUNCH(s, strip_step_kernel<true>, nblk(s->capacity, 128), 128, sg, p, exact_threshold(p.radius), s->hcap,
            s->A, s->cell_start, s->B, s->count, s->out[0], s->out[1], s->mcap, gx, s->st);
  else
    SLAUNCH(s, strip_step_kernel<false>, nblk(s->capacity, 128), 128, sg, p, 0.0f, s->hcap, s->A, s->cell_start,
            s->B, s->count, s->out[0], s->out[1], s->mcap, gx, s->st);
  if (ring) {
    SlotPtrs in_l = slot_ptrs(s->inbox, s->layout, 0, parity);
    SlotPtrs in_r = slot_ptrs(s->inbox, s->layout, 1, parity);
    static const int prewait_env = getenv("KG_STRIP_PREWAIT") ? atoi(getenv("KG_STRIP_PREWAIT")) : 1;  // 0: every append block waits (A/B)
    const int prewait = gx.fused && s->fold && prewait_env ? 1 : 0;
    if (gx.fused) {
      SLAUNCH(s, publish_flags_kernel, 1, 1, sg, gx, s->mcap, s->hcap, s->st, in_l, in_r, prewait);
    } else {
      PushMigArgs pm{};
      pm.src[0] = s->out[0];
      pm.src[1] = s->out[1];
      pm.src_count[0] = &s->st->out_count[0];
      pm.src_count[1] = &s->st->out_count[1];
      pm.gsrc[0] = s->gout[0];
      pm.gsrc[1] = s->gout[1];
      pm.gsrc_count[0] = sg.halo_l > 0 ? &s->st->gout_count[0] : nullptr;
      pm.gsrc_count[1] = sg.halo_r > 0 ? &s->st->gout_count[1] : nullptr;
      pm.dst[0] = gx.peer[0];
      pm.dst[1] = gx.peer[1];
      pm.done[0] = &s->st->push_done[0];
      pm.done[1] = &s->st->push_done[1];
      SLAUNCH(s, push_migrants_kernel, dim3(32, 2), kT, pm, s->mcap, s->hcap, epoch, s->st);
    }
    if (s->fold)
      SLAUNCH(s, append_all_kernel, nblk(2 * (uint64_t)s->mcap + 2 * (uint64_t)s->hcap), kT, sg, in_l, in_r, epoch,
              s->gself[0], s->gself[1], s->hcap, s->B, s->log_cap, s->count, s->st, prewait);
    else
      SLAUNCH(s, append_migrants_kernel, nblk(2 * (uint64_t)s->mcap), kT, sg, in_l, in_r, 1, 1, epoch, s->B,
              s->capacity, s->count, s->st);
  } else {
    SLAUNCH(s, set_log_len_kernel, 1, 1, s->st);
  }
  KG_TRY(strip_rebuild(s, true, epoch));
  s->steps_done += 1;
  return KG_OK;
}

}  // namespace

extern "C" {

int kg_strip_create(float w, float h, float disc, int toroidal, float radius, int rank, int nranks,
                    uint64_t capacity, uint64_t halo_capacity, uint64_t migrate_capacity, int device,
                    kg_strip** out) {
  if (!out) return fail(KG_E_INVALID, "null out");
  *out = nullptr;
  if (!(w > 0.f) || !(h > 0.f) || !(disc > 0.f) || !(radius > 0.f)) return fail(KG_E_INVALID, "w, h, disc, radius must be > 0");
  if (!toroidal) return fail(KG_E_INVALID, "strips need a toroidal field (clamped window, SURVEY F3)");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(KG_E_INVALID, "bad rank/nranks");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(KG_E_CUDA, "no CUDA device (%s); libkrabgpu has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return fail(KG_E_INVALID, "device %d out of range", device);
  KG_CUDA(cudaSetDevice(device));
  kg_strip* s = new kg_strip();
  s->device = device; s->rank = rank; s->nranks = nranks;
  s->prof_on = getenv("KG_STRIP_PROF") != nullptr;
  s->capacity = capacity;
  s->hcap = (uint32_t)std::max<uint64_t>(halo_capacity, 16);
  s->mcap = (uint32_t)std::max<uint64_t>(migrate_capacity, 16);
  StripGeom& sg = s->sg;
  sg.g.w = w; sg.g.h = h; sg.g.disc = disc; sg.g.toroidal = 1;
  sg.g.max_x = (int)std::min(2147483520.0f, ceilf(w / disc));
  sg.g.max_y = (int)std::min(2147483520.0f, ceilf(h / disc));
  sg.g.dw = sg.g.max_x + 1;
  sg.g.dh = sg.g.max_y + 1;
  sg.dd = (int)floorf(radius / disc);
  col_range(s, rank, &sg.own_x0, &sg.own_x1);
  s->left_rank = nranks > 1 ? (rank + nranks - 1) % nranks : -1;
  s->right_rank = nranks > 1 ? (rank + 1) % nranks : -1;
  if (nranks > 1) {
    col_range(s, s->left_rank, &sg.left_x0, &sg.left_x1);
    col_range(s, s->right_rank, &sg.right_x0, &sg.right_x1);
  } else {
    sg.left_x0 = sg.left_x1 = sg.right_x0 = sg.right_x1 = -1;
    sg.own_x0 = 0; sg.own_x1 = sg.g.dw;
  }
  sg.halo_l = rank > 0 ? sg.dd : 0;
  sg.halo_r = rank < nranks - 1 ? sg.dd : 0;
  int own_cols = sg.own_x1 - sg.own_x0;
  auto bail = [&](int code) { kg_strip_destroy(s); return code; };
  if (nranks > 1 && (own_cols < std::max(sg.dd, 1) + 1 || sg.dd > 64 || sg.dd < 1))
    return bail(fail(KG_E_INVALID, "strip of %d columns is too narrow for a %d-column halo", own_cols, sg.dd));
  double span = ((double)sg.dd + 1.0) * disc * 1.01 + 1e-3;
  if (span > 0.5 * std::min(w, h) || span > 1024.0 || std::max(w, h) > 1048576.0f)
    return bail(fail(KG_E_INVALID, "geometry outside the fast K4's domain (window must be < half the world)"));
  sg.x_off = sg.own_x0 - sg.halo_l;
  sg.ncols = sg.halo_l + own_cols + sg.halo_r;
  uint64_t nc = (uint64_t)sg.ncols * sg.g.dh;
  if (nc >= (1ull << 31) || capacity + 2ull * s->hcap >= (1ull << 32))
    return bail(fail(KG_E_INVALID, "strip too large for 32-bit indices"));
  sg.g.ncells = (uint32_t)nc;
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(KG_E_CUDA, "cudaStreamCreate failed"));
  int rc;
  if ((rc = preload_kernels()) != KG_OK) return bail(rc);
  if ((rc = alloc_agents_n(s->A, capacity + 2ull * s->hcap)) != KG_OK) return bail(rc);
  s->log_cap = capacity + 2ull * s->mcap + 2ull * s->hcap;  // stepped agents + migrants + (fold mode) ghosts
  if ((rc = alloc_agents_n(s->B, s->log_cap)) != KG_OK) return bail(rc);
  s->fold = getenv("KG_STRIP_HALO") ? strcmp(getenv("KG_STRIP_HALO"), "build") != 0 : true;
  s->fused_push = getenv("KG_STRIP_PUSH") ? strcmp(getenv("KG_STRIP_PUSH"), "kernel") != 0 : true;
  if ((rc = alloc_agents_n(s->out[0], s->mcap)) != KG_OK) return bail(rc);
  if ((rc = alloc_agents_n(s->out[1], s->mcap)) != KG_OK) return bail(rc);
  for (int k = 0; k < 2; ++k) {
    if ((rc = alloc_agents_n(s->gout[k], s->hcap)) != KG_OK) return bail(rc);
    if ((rc = alloc_agents_n(s->gself[k], s->mcap)) != KG_OK) return bail(rc);
  }
  if ((rc = lookback_init(s->scan, nc, s->stream)) != KG_OK) return bail(rc);
  s->layout = make_layout(s->mcap, s->hcap, (uint64_t)sg.dd * sg.g.dh);
  s->id_bitmap_bits = std::max<uint64_t>(8 * (capacity + 2ull * s->hcap), 1ull << 22) / 256 * 256 + 256;
  if (cudaMalloc(&s->cell_start, (nc + 16) * 4) != cudaSuccess ||
      cudaMalloc(&s->count, (nc + 16) * 4) != cudaSuccess ||
      cudaMalloc(&s->st, sizeof(StripState)) != cudaSuccess ||
      cudaMalloc(&s->halo_hist, 2 * ((size_t)sg.dd * sg.g.dh + 1) * 4) != cudaSuccess ||
      cudaMalloc(&s->halo_cursor, 2 * ((size_t)sg.dd * sg.g.dh + 1) * 4) != cudaSuccess ||
      cudaHostAlloc(&s->h_st, sizeof(StripState), cudaHostAllocDefault) != cudaSuccess ||
      cudaMalloc(&s->inbox, 4 * s->layout.bytes) != cudaSuccess ||
      cudaMalloc(&s->id_bitmap, s->id_bitmap_bits / 8) != cudaSuccess)
    return bail(fail(KG_E_CUDA, "strip allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
  cudaMemsetAsync(s->cell_start, 0, (nc + 16) * 4, s->stream);
  cudaMemsetAsync(s->count, 0, (nc + 16) * 4, s->stream);
  cudaMemsetAsync(s->st, 0, sizeof(StripState), s->stream);
  {
    const uint32_t none = 0xFFFFFFFFu;  // no ghosts in the log
    cudaMemcpyAsync(&s->st->ghost_begin, &none, 4, cudaMemcpyHostToDevice, s->stream);
  }
  cudaMemsetAsync(s->halo_hist, 0, 2 * ((size_t)sg.dd * sg.g.dh + 1) * 4, s->stream);
  cudaMemsetAsync(s->inbox, 0, 4 * s->layout.bytes, s->stream);
  if (cudaStreamSynchronize(s->stream) != cudaSuccess) return bail(fail(KG_E_CUDA, "strip init failed"));
  *out = s;
  return KG_OK;
}

int kg_strip_destroy(kg_strip* s) {
  if (!s) return KG_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->prof_on && !s->prof_ev.empty()) {  // KG_STRIP_PROF=1: per-kernel totals of the whole run
    std::vector<std::pair<std::string, std::pair<double, int>>> tot;
    for (auto& e : s->prof_ev) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e.second.first, e.second.second);
      cudaEventDestroy(e.second.first);
      cudaEventDestroy(e.second.second);
      bool found = false;
      for (auto& t : tot)
        if (t.first == e.first) { t.second.first += ms; t.second.second += 1; found = true; }
      if (!found) tot.push_back({e.first, {ms, 1}});
    }
    for (auto& t : tot)
      fprintf(stderr, "[strip %d] %-28s n=%5d  mean %8.1f us\n", s->rank, t.first.c_str(), t.second.second,
              1e3 * t.second.first / t.second.second);
  }
  for (int k = 0; k < 2; ++k)
    if (s->peer_inbox[k] && s->peer_is_ipc[k]) {
      if (k == 1 && s->peer_inbox[1] == s->peer_inbox[0]) continue;
      cudaIpcCloseMemHandle(s->peer_inbox[k]);
    }
  s->watch.destroy();
  s->flusher.destroy();
  s->events.destroy();
  cudaFree(s->A.id); cudaFree(s->A.pv); cudaFree(s->B.id); cudaFree(s->B.pv);
  for (int k = 0; k < 2; ++k) {
    cudaFree(s->out[k].id); cudaFree(s->out[k].pv);
    cudaFree(s->gout[k].id); cudaFree(s->gout[k].pv);
    cudaFree(s->gself[k].id); cudaFree(s->gself[k].pv);
  }
  cudaFree(s->halo_hist);
  cudaFree(s->halo_cursor);
  if (s->have_stage) {
    cudaFree(s->stage.id); cudaFree(s->stage.x); cudaFree(s->stage.y); cudaFree(s->stage.dx);
    cudaFree(s->stage.dy);
  }
  lookback_destroy(s->scan);
  cudaFree(s->cell_start);
  cudaFree(s->count);
  cudaFree(s->st);
  cudaFree(s->inbox);
  cudaFree(s->id_bitmap);
  if (s->h_st) cudaFreeHost(s->h_st);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return KG_OK;
}

int kg_strip_columns(kg_strip* s, int32_t* own_x0, int32_t* own_x1, int32_t* halo_l, int32_t* halo_r,
                     int32_t* dh) {
  if (!s) return fail(KG_E_INVALID, "null strip handle");
  if (own_x0) *own_x0 = s->sg.own_x0;
  if (own_x1) *own_x1 = s->sg.own_x1;
  if (halo_l) *halo_l = s->sg.halo_l;
  if (halo_r) *halo_r = s->sg.halo_r;
  if (dh) *dh = s->sg.g.dh;
  return KG_OK;
}

int kg_strip_set_order(kg_strip* s, int order) {
  if (!s) return fail(KG_E_INVALID, "null strip handle");
  s->order = order;
  return KG_OK;
}

int kg_strip_ipc_export(kg_strip* s, void* handle64) {
  KG_TRY(suse(s));
  if (!handle64) return fail(KG_E_INVALID, "null handle buffer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "KG_IPC_HANDLE_BYTES");
  cudaIpcMemHandle_t h;
  KG_CUDA(cudaIpcGetMemHandle(&h, s->inbox));
  memcpy(handle64, &h, sizeof(h));
  return KG_OK;
}

int kg_strip_connect_ipc(kg_strip* s, const void* left_handle64, const void* right_handle64) {
  KG_TRY(suse(s));
  if (s->nranks == 1) return KG_OK;
  if (!left_handle64 || !right_handle64) return fail(KG_E_INVALID, "both ring neighbours are required");
  cudaIpcMemHandle_t hl, hr;
  memcpy(&hl, left_handle64, 64);
  memcpy(&hr, right_handle64, 64);
  KG_CUDA(cudaIpcOpenMemHandle(&s->peer_inbox[0], hl, cudaIpcMemLazyEnablePeerAccess));
  s->peer_is_ipc[0] = true;
  if (s->left_rank == s->right_rank) {
    s->peer_inbox[1] = s->peer_inbox[0];  // two ranks: both neighbours are the same GPU
  } else {
    KG_CUDA(cudaIpcOpenMemHandle(&s->peer_inbox[1], hr, cudaIpcMemLazyEnablePeerAccess));
  }
  s->peer_is_ipc[1] = true;
  return KG_OK;
}

int kg_strip_connect_local(kg_strip* s, kg_strip* left, kg_strip* right) {
  KG_TRY(suse(s));
  if (s->nranks == 1) return KG_OK;
  if (!left || !right) return fail(KG_E_INVALID, "both ring neighbours are required");
  kg_strip* nb[2] = {left, right};
  for (int k = 0; k < 2; ++k) {
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

int kg_strip_init_flockers(kg_strip* s, uint64_t n_global, uint64_t seed) {
  KG_TRY(suse(s));
  if (n_global >= (1ull << 32)) return fail(KG_E_CAPACITY, "ids are 32-bit");
  SLAUNCH(s, strip_init_kernel, nblk(n_global), kT, s->sg, n_global, seed, s->B, s->capacity, s->count, s->st);
  s->prepared = false;
  return strip_sync_check(s);
}

int kg_strip_upload(kg_strip* s, uint64_t n, const uint32_t* id, const float* x, const float* y,
                    const float* dx, const float* dy) {
  KG_TRY(suse(s));
  if (n == 0) return KG_OK;
  if (!id || !x || !y || !dx || !dy) return fail(KG_E_INVALID, "null input array");
  KG_TRY(strip_reset_log(s));
  KG_TRY(strip_sync_check(s));
  uint64_t have = s->h_st->n_log;
  if (have + n > s->capacity) return fail(KG_E_CAPACITY, "strip upload exceeds capacity");
  if (!s->have_stage) {
    uint64_t m = s->capacity + 64;
    KG_CUDA(cudaMalloc(&s->stage.id, m * 4)); KG_CUDA(cudaMalloc(&s->stage.x, m * 4));
    KG_CUDA(cudaMalloc(&s->stage.y, m * 4)); KG_CUDA(cudaMalloc(&s->stage.dx, m * 4));
    KG_CUDA(cudaMalloc(&s->stage.dy, m * 4));
    s->have_stage = true;
  }
  cudaStream_t st = s->stream;
  KG_CUDA(cudaMemcpyAsync(s->stage.id, id, n * 4, cudaMemcpyHostToDevice, st));
  KG_CUDA(cudaMemcpyAsync(s->stage.x, x, n * 4, cudaMemcpyHostToDevice, st));
  KG_CUDA(cudaMemcpyAsync(s->stage.y, y, n * 4, cudaMemcpyHostToDevice, st));
  KG_CUDA(cudaMemcpyAsync(s->stage.dx, dx, n * 4, cudaMemcpyHostToDevice, st));
  KG_CUDA(cudaMemcpyAsync(s->stage.dy, dy, n * 4, cudaMemcpyHostToDevice, st));
  SLAUNCH(s, strip_pack_kernel, nblk(n), kT, n, s->stage.id, s->stage.x, s->stage.y, s->stage.dx,
          s->stage.dy, s->B, have);
  SLAUNCH(s, strip_hist_kernel, nblk(n), kT, s->sg, have, n, s->B, s->count, s->st);
  uint32_t nl = (uint32_t)(have + n);
  KG_CUDA(cudaMemcpyAsync(&s->st->n_log, &nl, 4, cudaMemcpyHostToDevice, st));
  s->prepared = false;
  s->ids_trusted = false;  // verified on the device after every rebuild from now on
  return strip_sync_check(s);
}

int kg_strip_clear(kg_strip* s) {
  KG_TRY(suse(s));
  // forget every agent (owned, logged, staged); cell counts are already zero between rebuilds
  KG_TRY(strip_reset_log(s));
  KG_CUDA(cudaMemsetAsync(s->st, 0, offsetof(StripState, err), s->stream));
  KG_CUDA(cudaMemsetAsync(s->count, 0, (size_t)s->sg.g.ncells * 4, s->stream));
  KG_CUDA(cudaMemsetAsync(s->halo_hist, 0, 2 * ((size_t)s->sg.dd * s->sg.g.dh + 1) * 4, s->stream));
  SLAUNCH(s, strip_ids_reset_kernel, 1, 1, s->st);
  s->ids_trusted = true;
  s->prepared = false;
  return KG_OK;
}

int kg_strip_prepare(kg_strip* s) {
  KG_TRY(suse(s));
  if (s->nranks > 1 && (!s->peer_inbox[0] || !s->peer_inbox[1]))
    return fail(KG_E_INVALID, "strip is not connected to its neighbours");
  KG_TRY(strip_reset_log(s));
  KG_TRY(strip_rebuild(s, false, 0));
  s->prepared = true;
  return KG_OK;
}

int kg_strip_step_boids(kg_strip* s, const KgBoidsParams* p) {
  KG_TRY(suse(s));
  if (!p) return fail(KG_E_INVALID, "null params");
  return strip_step(s, *p);
}

int kg_strip_run_boids(kg_strip* s, const KgBoidsParams* p, uint64_t nsteps) {
  KG_TRY(suse(s));
  if (!p) return fail(KG_E_INVALID, "null params");
  KgBoidsParams q = *p;
  for (uint64_t i = 0; i < nsteps; ++i) {
    q.step = p->step + i;
    KG_TRY(strip_step(s, q));
  }
  return KG_OK;
}

int kg_strip_run_boids_timed(kg_strip* s, const KgBoidsParams* p, uint64_t nsteps,
                             uint64_t flush_bytes, double* ms_sum) {
  KG_TRY(suse(s));
  if (!p || !ms_sum) return fail(KG_E_INVALID, "null argument");
  KgBoidsParams q = *p;
  for (uint64_t i = 0; i < nsteps; ++i) {
    cudaEvent_t a = nullptr, b = nullptr;
    KG_TRY(s->events.get(2 * i, &a));
    KG_TRY(s->events.get(2 * i + 1, &b));
    KG_TRY(s->flusher.run(flush_bytes, s->stream));
    q.step = p->step + i;
    KG_CUDA(cudaEventRecord(a, s->stream));
    KG_TRY(strip_step(s, q));
    KG_CUDA(cudaEventRecord(b, s->stream));
  }
  KG_TRY(strip_sync_check(s));
  double sum = 0;
  for (uint64_t i = 0; i < nsteps; ++i) {
    float t = 0.f;
    KG_CUDA(cudaEventElapsedTime(&t, s->events.ev[2 * i], s->events.ev[2 * i + 1]));
    sum += t;
  }
  *ms_sum = sum;
  return KG_OK;
}

int kg_strip_sync(kg_strip* s) {
  KG_TRY(suse(s));
  return strip_sync_check(s);
}

int kg_strip_stats(kg_strip* s, uint64_t* n_owned, uint64_t* migrants_in, uint64_t* migrants_out,
                   uint64_t* halo_left, uint64_t* halo_right, uint64_t* launches) {
  KG_TRY(suse(s));
  KG_TRY(strip_sync_check(s));
  if (n_owned) *n_owned = s->h_st->n_owned;
  if (migrants_in) *migrants_in = s->h_st->mig_in_total;
  if (migrants_out) *migrants_out = s->h_st->mig_out_total;
  if (halo_left) *halo_left = s->h_st->halo_in[0];
  if (halo_right) *halo_right = s->h_st->halo_in[1];
  if (launches) *launches = s->launches;
  return KG_OK;
}

int kg_strip_download(kg_strip* s, uint64_t cap, uint32_t* id, float* x, float* y, float* dx,
                      float* dy, uint64_t* n_out) {
  KG_TRY(suse(s));
  KG_TRY(strip_sync_check(s));
  uint64_t n = s->h_st->n_owned;
  if (n_out) *n_out = n;
  if (n > cap) return fail(KG_E_CAPACITY, "download needs room for %llu agents", (unsigned long long)n);
  if (n == 0) return KG_OK;
  if (!s->have_stage) {
    uint64_t m = s->capacity + 64;
    KG_CUDA(cudaMalloc(&s->stage.id, m * 4)); KG_CUDA(cudaMalloc(&s->stage.x, m * 4));
    KG_CUDA(cudaMalloc(&s->stage.y, m * 4)); KG_CUDA(cudaMalloc(&s->stage.dx, m * 4));
    KG_CUDA(cudaMalloc(&s->stage.dy, m * 4));
    s->have_stage = true;
  }
  SLAUNCH(s, strip_unpack_kernel, nblk(n), kT, n, s->hcap, s->A, s->st, s->stage.id, s->stage.x,
          s->stage.y, s->stage.dx, s->stage.dy);
  cudaStream_t st = s->stream;
  if (id) KG_CUDA(cudaMemcpyAsync(id, s->stage.id, n * 4, cudaMemcpyDeviceToHost, st));
  if (x) KG_CUDA(cudaMemcpyAsync(x, s->stage.x, n * 4, cudaMemcpyDeviceToHost, st));
  if (y) KG_CUDA(cudaMemcpyAsync(y, s->stage.y, n * 4, cudaMemcpyDeviceToHost, st));
  if (dx) KG_CUDA(cudaMemcpyAsync(dx, s->stage.dx, n * 4, cudaMemcpyDeviceToHost, st));
  if (dy) KG_CUDA(cudaMemcpyAsync(dy, s->stage.dy, n * 4, cudaMemcpyDeviceToHost, st));
  return strip_sync_check(s);
}

int kg_strip_timer_start(kg_strip* s) {
  KG_TRY(suse(s));
  return s->watch.start(s->stream);
}
int kg_strip_timer_stop(kg_strip* s, double* ms) {
  KG_TRY(suse(s));
  return s->watch.stop(s->stream, ms);
}

}  // extern "C"
