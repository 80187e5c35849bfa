// K5: the Forest-Fire class stencil on u8 DenseNumberGrid2D cells, shared by the single-GPU grid
// (grid.cu) and the multi-GPU row strips (gridstrip.cu).
#pragma once
#include "common.cuh"

namespace kg {

enum : uint32_t { FF_GREEN = 1, FF_BURNING = 2, FF_BURNED = 3 };

// Everything a strip of rows needs beyond its own buffers; all-zero for a whole grid on one GPU.
// Rows x = -1 and x = width of a strip live in its inbox, written by the line neighbours with
// peer stores; a neighbour's completed push is announced by an epoch flag.
struct FFExchange {
  // An inbox slot holds TWO rows, `height` bytes apart: rows -2, -1 (from the left neighbour) or
  // rows width, width+1 (from the right one).  One step per pass needs one of them, the two-step
  // pass both; every pass pushes both, so that either kind of pass can follow.
  const uint8_t* halo_lo = nullptr;   // row x = -1 (nullptr: outside the world => no fire)
  const uint8_t* halo_lo2 = nullptr;  // row x = -2
  const uint8_t* halo_hi = nullptr;   // row x = width
  const uint8_t* halo_hi2 = nullptr;  // row x = width + 1
  const unsigned long long* flag_lo = nullptr;  // wait until *flag >= wait_epoch before reading halo
  const unsigned long long* flag_hi = nullptr;
  unsigned long long wait_epoch = 0;
  uint8_t* push_lo = nullptr;  // left neighbour's slot: my new row x (0 or 1) goes to push_lo + x * height
  uint8_t* push_hi = nullptr;  // right neighbour's slot: my new row x (width-2 or width-1) to push_hi + (x - (width-2)) * height
  unsigned long long* push_flag_lo = nullptr;
  unsigned long long* push_flag_hi = nullptr;
  unsigned long long push_epoch = 0;
  uint32_t* done = nullptr;  // [2] completion counters of the blocks that push (lo, hi)
  int* err = nullptr;        // bit 0: a neighbour's flag timed out
};

// A lane owns 16 consecutive y cells (one uint4) and marches down `rows` consecutive x rows with a
// three-row sliding window held in registers; a warp therefore streams 512 contiguous bytes per
// row.  Per row the lane derives the byte-parallel mask "a burning cell is at y-1, y or y+1"
// (neighbour lanes supply the two halo bytes by shuffle, the warp's outer halo by two byte loads);
// OR-ing the masks of rows x-1, x, x+1 gives the Moore-8 test.  Cells hold 1, 2, 3 or 0xFF, so
// burning = bit1 & ~bit0 and green = bit0 & ~bit1 in every byte.
struct Row {
  uint32_t v[4];   // the 16 cells
  uint32_t hm[4];  // 0x01 in every byte whose y-1 / y / y+1 neighbour (same row) is burning
};

// `row` = first byte of grid row x (nullptr: the row does not exist).  All loads are ld.global.cg
// (L2, no L1): the data is streamed once, and the halo rows of a strip were written by another GPU
// during this kernel's lifetime, so they must not come from a non-coherent cache.
__device__ __forceinline__ void load_row(Row& r, const uint8_t* __restrict__ row, int32_t height, int64_t y0,
                                         int lane, bool in_y) {
  const uint32_t M = 0x01010101u;
  uint32_t b[4] = {0, 0, 0, 0};
  uint32_t left = 0, right = 0;  // burning flag (bit 0) of the cell just below / above our chunk
  if (row != nullptr && in_y) {
    const uint8_t* p = row + y0;
    uint4 q = __ldcg(reinterpret_cast<const uint4*>(p));
    r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
#pragma unroll
    for (int k = 0; k < 4; ++k) b[k] = (r.v[k] >> 1) & ~r.v[k] & M;
    // outer halo of the warp's 512-byte span
    if (lane == 0 && y0 > 0) left = (__ldcg(p - 1) == FF_BURNING);
    if (lane == 31 && y0 + 16 < height) right = (__ldcg(p + 16) == FF_BURNING);
  } else {
    r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0xFFFFFFFFu;
  }
  uint32_t from_below = __shfl_up_sync(0xffffffffu, b[3] >> 24, 1);
  uint32_t from_above = __shfl_down_sync(0xffffffffu, b[0] & 1u, 1);
  if (lane != 0) left = from_below;
  if (lane != 31) right = from_above;
  // up[k]: flags moved one cell towards higher y (neighbour y-1), dn[k]: towards lower y (y+1)
  uint32_t up0 = (b[0] << 8) | left;
  uint32_t up1 = __funnelshift_l(b[0], b[1], 8);
  uint32_t up2 = __funnelshift_l(b[1], b[2], 8);
  uint32_t up3 = __funnelshift_l(b[2], b[3], 8);
  uint32_t dn0 = __funnelshift_r(b[0], b[1], 8);
  uint32_t dn1 = __funnelshift_r(b[1], b[2], 8);
  uint32_t dn2 = __funnelshift_r(b[2], b[3], 8);
  uint32_t dn3 = (b[3] >> 8) | (right << 24);
  r.hm[0] = b[0] | up0 | dn0;
  r.hm[1] = b[1] | up1 | dn1;
  r.hm[2] = b[2] | up2 | dn2;
  r.hm[3] = b[3] | up3 | dn3;
}

// next state of row x (cur) from rows x-1, x, x+1; stores it and, for a strip's boundary rows,
// also into the line neighbour's inbox (peer stores)
template <bool WRITE_NONE>
__device__ __forceinline__ void ff_emit_row(const Row& prev, const Row& cur, const Row& next,
                                            uint8_t* __restrict__ wr, int32_t x, int32_t width,
                                            int32_t height, int64_t y0, const FFExchange& ex) {
  const uint32_t M = 0x01010101u;
  uint32_t o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t v = cur.v[k];
    uint32_t s = v >> 1;
    uint32_t burning = s & ~v & M;
    uint32_t green = v & ~s & M;
    uint32_t fire = prev.hm[k] | cur.hm[k] | next.hm[k];
    o[k] = v + (green & fire) + burning;  // 1->2 on fire, 2->3, 3 and 0xFF unchanged
  }
  uint8_t* q = wr + (uint64_t)x * (uint64_t)height + y0;
  if (!WRITE_NONE) {
    // only live cells may be written: merge with what the write buffer already holds
    uint4 old = *reinterpret_cast<const uint4*>(q);
    uint32_t ov[4] = {old.x, old.y, old.z, old.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t v = cur.v[k];
      // byte is None iff all 8 bits set: AND-fold the bits into bit 0
      uint32_t t = v & (v >> 4);
      t &= t >> 2;
      t &= t >> 1;
      uint32_t none_mask = (t & M) * 0xFFu;
      o[k] = (o[k] & ~none_mask) | (ov[k] & none_mask);
    }
  }
  const uint4 ov4 = make_uint4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<uint4*>(q) = ov4;
  if (x <= 1 && ex.push_lo) *reinterpret_cast<uint4*>(ex.push_lo + (int64_t)x * height + y0) = ov4;
  if (x >= width - 2 && ex.push_hi)
    *reinterpret_cast<uint4*>(ex.push_hi + (int64_t)(x - (width - 2)) * height + y0) = ov4;
}

// one thread parks on a neighbour's flag (bounded: ~4 s, then err bit 0)
__device__ __forceinline__ void ff_wait_flag(const unsigned long long* flag, unsigned long long epoch,
                                             int* err) {
  if (threadIdx.x == 0) {
    const volatile unsigned long long* f = flag;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < epoch) {
      __nanosleep(100);
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > 4000000000ull) {
        atomicOr(err, 1);
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}

// the last of the `nblocks` pushing blocks publishes the epoch behind a system-scope fence
__device__ __forceinline__ void ff_publish(uint32_t* done, uint32_t nblocks, unsigned long long* flag,
                                           unsigned long long epoch) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t prev = atomicAdd(done, 1u);
    if (prev == nblocks - 1) {
      *done = 0;
      __threadfence_system();
      *(volatile unsigned long long*)flag = epoch;
      __threadfence_system();
    }
  }
}

// 12 resident CTAs per SM (40 registers): measured best on B200 — the kernel is bound by HBM
// latency x bandwidth, so resident warps matter more than registers per thread
#ifndef KG_FF_MINB
#define KG_FF_MINB 12
#endif
template <bool WRITE_NONE>
__global__ void __launch_bounds__(128, KG_FF_MINB)
forest_fire_u8_kernel(const uint8_t* __restrict__ rd, uint8_t* __restrict__ wr, int32_t width,
                      int32_t height, int32_t rows_per_strip, FFExchange ex) {
  grid_dep_wait();  // the read buffer is the previous step's output (dependent launch, common.cuh)
  int lane = threadIdx.x & 31;
  int warp_in_block = threadIdx.x >> 5;
  // blockIdx.x: 512-byte span of y (4 warps per block stack 4 spans), blockIdx.y: strip of rows
  int64_t y0 = ((int64_t)(blockIdx.x * 4 + warp_in_block) * 32 + lane) * 16;
  bool in_y = y0 < height;  // warp-uniform per 512-byte span except the ragged last span
  // Blocks are dispatched in blockIdx order: give the LAST row tile the second slot, so that both
  // boundary rows of a strip are produced — and their flags published — at the start of the kernel,
  // not at its end, where every neighbour would already be waiting for them.
  const uint32_t ntiles = gridDim.y;
  const uint32_t tile = blockIdx.y == 0 ? 0u : (blockIdx.y == 1 ? ntiles - 1 : blockIdx.y - 1);
  int32_t x_begin = (int32_t)tile * rows_per_strip;
  int32_t x_end = min(width, x_begin + rows_per_strip);
  if (x_begin >= width) return;
  const bool first = x_begin == 0, last = x_end == width;
  if (first && ex.flag_lo) ff_wait_flag(ex.flag_lo, ex.wait_epoch, ex.err);
  if (last && ex.flag_hi) ff_wait_flag(ex.flag_hi, ex.wait_epoch, ex.err);
  auto rowp = [&](int32_t x) -> const uint8_t* {
    if (x >= 0 && x < width) return rd + (uint64_t)x * (uint64_t)height;
    return x < 0 ? ex.halo_lo : ex.halo_hi;  // nullptr outside the world
  };
  // rows a, b, c rotate through the roles (x-1, x, x+1): unrolled by three, no register moves
  Row a, b, c;
  load_row(a, rowp(x_begin - 1), height, y0, lane, in_y);
  load_row(b, rowp(x_begin), height, y0, lane, in_y);
  for (int32_t x = x_begin; x < x_end; x += 3) {
    load_row(c, rowp(x + 1), height, y0, lane, in_y);
    if (in_y) ff_emit_row<WRITE_NONE>(a, b, c, wr, x, width, height, y0, ex);
    if (x + 1 >= x_end) break;
    load_row(a, rowp(x + 2), height, y0, lane, in_y);
    if (in_y) ff_emit_row<WRITE_NONE>(b, c, a, wr, x + 1, width, height, y0, ex);
    if (x + 2 >= x_end) break;
    load_row(b, rowp(x + 3), height, y0, lane, in_y);
    if (in_y) ff_emit_row<WRITE_NONE>(c, a, b, wr, x + 2, width, height, y0, ex);
  }
  if (first && ex.push_lo) ff_publish(ex.done + 0, gridDim.x, ex.push_flag_lo, ex.push_epoch);
  if (last && ex.push_hi) ff_publish(ex.done + 1, gridDim.x, ex.push_flag_hi, ex.push_epoch);
}

// ------------------------------------------------------------------------------------------
// Two steps per pass (temporal blocking).  K5 above sits at ~81 % of the HBM roofline, so the only
// way to more cell-updates per second is fewer bytes per update: this kernel reads step t once and
// writes step t+2, keeping step t+1 in registers — 1 B read + 1 B written per cell for TWO updates.
// The rule is deterministic and None cells never change, so two fused steps equal two launches of
// K5 bit for bit (tests/test_gpu_grid.py runs both against the oracle).
//
// Same register-window march as K5, with a second three-row window one time level up.  A result
// cell of step t+2 depends on step t within two cells, so a warp that loads 32 x 16 cells of y per
// row can only vouch for the inner 30 x 16 = 480: lanes 0 and 31 feed their neighbours' first step
// and store nothing (6 % redundant work, and no byte-wide halo loads at all); a block of
// `rows_per_tile` rows likewise reads two extra rows above and below.

// burning-neighbour mask of a row held in registers (any time level); the cells beyond the warp's
// span count as not burning — only lanes 0 and 31 can see the difference, and only in cells whose
// second step is discarded
__device__ __forceinline__ void ff_row_mask(Row& r, int lane) {
  const uint32_t M = 0x01010101u;
  uint32_t b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) b[k] = (r.v[k] >> 1) & ~r.v[k] & M;
  uint32_t left = __shfl_up_sync(0xffffffffu, b[3] >> 24, 1);
  uint32_t right = __shfl_down_sync(0xffffffffu, b[0] & 1u, 1);
  if (lane == 0) left = 0;
  if (lane == 31) right = 0;
  const uint32_t up0 = (b[0] << 8) | left;
  const uint32_t up1 = __funnelshift_l(b[0], b[1], 8);
  const uint32_t up2 = __funnelshift_l(b[1], b[2], 8);
  const uint32_t up3 = __funnelshift_l(b[2], b[3], 8);
  const uint32_t dn0 = __funnelshift_r(b[0], b[1], 8);
  const uint32_t dn1 = __funnelshift_r(b[1], b[2], 8);
  const uint32_t dn2 = __funnelshift_r(b[2], b[3], 8);
  const uint32_t dn3 = (b[3] >> 8) | (right << 24);
  r.hm[0] = b[0] | up0 | dn0;
  r.hm[1] = b[1] | up1 | dn1;
  r.hm[2] = b[2] | up2 | dn2;
  r.hm[3] = b[3] | up3 | dn3;
}

// next state of the cells of `cur` given the masks of the rows above and below (ff_emit_row's rule)
__device__ __forceinline__ void ff_next(uint32_t (&o)[4], const Row& prev, const Row& cur, const Row& next) {
  const uint32_t M = 0x01010101u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t v = cur.v[k];
    const uint32_t s = v >> 1;
    const uint32_t burning = s & ~v & M;
    const uint32_t green = v & ~s & M;
    const uint32_t fire = prev.hm[k] | cur.hm[k] | next.hm[k];
    o[k] = v + (green & fire) + burning;  // 1->2 on fire, 2->3, 3 and 0xFF unchanged
  }
}

#ifndef KG_FF2_MINB
#define KG_FF2_MINB 7
#endif
#ifndef KG_FF2_RING
#define KG_FF2_RING 0  // rows in flight per warp through a cp.async ring (0: one row ahead, held in registers; measured faster: the kernel is bound by the integer pipe, not by load latency)
#endif
constexpr int kFF2Span = 480;  // cells of y a warp produces per row

// `width` = rows of this grid / strip.  A strip's rows -2, -1, width, width+1 come from its inbox and
// its new rows 0, 1, width-2, width-1 also go to the line neighbours' inboxes, as in K5; the host
// guarantees that both pushed rows of a side lie in the first / last row tile (the tiles that publish).
static __global__ void __launch_bounds__(128, KG_FF2_MINB)
forest_fire_u8_x2_kernel(const uint8_t* __restrict__ rd, uint8_t* __restrict__ wr, int32_t width,
                         int32_t height, int32_t rows_per_tile, FFExchange ex) {
  grid_dep_wait();  // the read buffer is the previous pass's output (dependent launch, common.cuh)
  const int lane = threadIdx.x & 31;
  const int64_t span = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t y0 = span * kFF2Span + (int64_t)(lane - 1) * 16;
  const bool live = span * kFF2Span < height;  // warp-uniform
  const bool in_y = y0 >= 0 && y0 < height;
  const bool owner = in_y && lane != 0 && lane != 31;
  // boundary tiles first (see K5): tile 0, then the last one, then the interior
  const uint32_t ntiles = gridDim.y;
  const uint32_t tile = blockIdx.y == 0 ? 0u : (blockIdx.y == 1 ? ntiles - 1 : blockIdx.y - 1);
  const int32_t x_begin = (int32_t)tile * rows_per_tile;
  const int32_t x_end = min(width, x_begin + rows_per_tile);
  if (x_begin >= width) return;
  const bool first = x_begin == 0, last = x_end == width;
  if (first && ex.flag_lo) ff_wait_flag(ex.flag_lo, ex.wait_epoch, ex.err);
  if (last && ex.flag_hi) ff_wait_flag(ex.flag_hi, ex.wait_epoch, ex.err);

  auto rowp = [&](int32_t x) -> const uint8_t* {  // step-t row x as stored; nullptr where the grid has no row
    if (x >= 0 && x < width) return rd + (uint64_t)x * (uint64_t)height;
    if (x == -1) return ex.halo_lo;
    if (x == -2) return ex.halo_lo2;
    if (x == width) return ex.halo_hi;
    if (x == width + 1) return ex.halo_hi2;
    return nullptr;
  };
  const uint4 kNone4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
#if KG_FF2_RING
  // Rows travel global -> shared by cp.async (16 bytes per lane, L2 only), KG_FF2_RING rows ahead of the
  // arithmetic: the bytes in flight per warp no longer cost registers, and at 28-32 resident warps per SM
  // it takes three to four rows in flight per warp to cover HBM latency at full bandwidth.  A lane only
  // ever reads back the 16 bytes it copied itself, so the only synchronisation is its own wait_group.
  __shared__ uint4 ring[4][KG_FF2_RING][32];
  uint4* const my_ring = &ring[threadIdx.x >> 5][0][lane];
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(my_ring);
  auto issue = [&](int32_t x, int32_t k) {  // row x into slot k % KG_FF2_RING; always commits a group
    const uint8_t* row = rowp(x);
    if (row != nullptr && in_y && x <= x_end + 1)  // nothing beyond the tile's last input row
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + (uint32_t)(k % KG_FF2_RING) * 512u),
                   "l"(row + y0)
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto take = [&](int32_t x, int32_t k) -> uint4 {  // the oldest row in flight
    asm volatile("cp.async.wait_group %0;" ::"n"(KG_FF2_RING - 1) : "memory");
    const uint4 q = my_ring[(k % KG_FF2_RING) * 32];
    return (rowp(x) != nullptr && in_y) ? q : kNone4;
  };
#else
  auto fetch = [&](int32_t x) -> uint4 {
    const uint8_t* row = rowp(x);
    if (row != nullptr && in_y) return __ldcg(reinterpret_cast<const uint4*>(row + y0));
    return kNone4;
  };
#endif
  if (live) {
#if KG_FF2_RING
    // arrival number k <-> row x_begin - 2 + k
#pragma unroll
    for (int k = 0; k < KG_FF2_RING; ++k) issue(x_begin - 2 + k, k);
#else
    uint4 pre = fetch(x_begin);  // always one row ahead of the arithmetic
#endif
    // (a0, b0, c0): step-t rows xin-2, xin-1, xin;  (a1, b1, c1): step-t+1 rows xin-3, xin-2, xin-1
    auto stage = [&](Row& a0, Row& b0, Row& c0, Row& a1, Row& b1, Row& c1, int32_t xin) {
#if KG_FF2_RING
      const int32_t k = xin - x_begin + 2;
      const uint4 q = take(xin, k);
#else
      const uint4 q = pre;
      pre = fetch(xin + 1);
#endif
      c0.v[0] = q.x; c0.v[1] = q.y; c0.v[2] = q.z; c0.v[3] = q.w;
      ff_row_mask(c0, lane);
      ff_next(c1.v, a0, b0, c0);  // step t+1 of row xin-1 (a row outside the world stays None)
      ff_row_mask(c1, lane);
      const int32_t xo = xin - 2;  // step t+2 of row xin-2
      if (xo >= x_begin) {
        uint32_t o[4];
        ff_next(o, a1, b1, c1);
        if (owner) {
          const uint4 ov4 = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(wr + (uint64_t)xo * (uint64_t)height + y0) = ov4;
          if (xo <= 1 && ex.push_lo) *reinterpret_cast<uint4*>(ex.push_lo + (int64_t)xo * height + y0) = ov4;
          if (xo >= width - 2 && ex.push_hi)
            *reinterpret_cast<uint4*>(ex.push_hi + (int64_t)(xo - (width - 2)) * height + y0) = ov4;
        }
      }
#if KG_FF2_RING
      issue(xin + KG_FF2_RING, k + KG_FF2_RING);  // into the slot just read (c0 is in registers by now)
#endif
    };
    Row A0, B0, C0, A1 = {}, B1 = {}, C1 = {};
    {
#if KG_FF2_RING
      const uint4 qa = take(x_begin - 2, 0);
      issue(x_begin - 2 + KG_FF2_RING, KG_FF2_RING);
      const uint4 qb = take(x_begin - 1, 1);
      issue(x_begin - 1 + KG_FF2_RING, 1 + KG_FF2_RING);
#else
      const uint4 qa = fetch(x_begin - 2), qb = fetch(x_begin - 1);
#endif
      A0.v[0] = qa.x; A0.v[1] = qa.y; A0.v[2] = qa.z; A0.v[3] = qa.w;
      B0.v[0] = qb.x; B0.v[1] = qb.y; B0.v[2] = qb.z; B0.v[3] = qb.w;
      ff_row_mask(A0, lane);
      ff_row_mask(B0, lane);
    }
    // rows x_begin .. x_end+1 arrive; the three roles rotate through the registers, no moves
    const int32_t x_last = x_end + 1;
    for (int32_t xin = x_begin;; xin += 3) {
      stage(A0, B0, C0, A1, B1, C1, xin);
      if (xin >= x_last) break;
      stage(B0, C0, A0, B1, C1, A1, xin + 1);
      if (xin + 1 >= x_last) break;
      stage(C0, A0, B0, C1, A1, B1, xin + 2);
      if (xin + 2 >= x_last) break;
    }
  }
  if (first && ex.push_lo) ff_publish(ex.done + 0, gridDim.x, ex.push_flag_lo, ex.push_epoch);
  if (last && ex.push_hi) ff_publish(ex.done + 1, gridDim.x, ex.push_flag_hi, ex.push_epoch);
}

}  // namespace kg
