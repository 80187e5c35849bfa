// K5: the Forest-Fire class stencil on u8 DenseNumberGrid2D cells, shared by the single-GPU grid
// (grid.cu) and the multi-GPU row strips (gridstrip.cu).
#pragma once
#include "common.cuh"

namespace kg {

enum : uint32_t { FF_GREEN = 1, FF_BURNING = 2, FF_BURNED = 3 };

// Everything a strip of rows needs beyond its own buffers; all-zero for a whole grid on one GPU.
// Rows x = -1 and x = width of a strip live in its inbox, written by the line neighbours with
// peer stores; a neighbour's completed push is announced by an epoch flag.
constexpr int kFFHalo = 8;  // rows per inbox slot = the most steps one pass may fuse

struct FFExchange {
  // An inbox slot holds kFFHalo rows, `height` bytes apart: rows -kFFHalo .. -1 (from the left
  // neighbour; row x at index kFFHalo + x) or rows width .. width + kFFHalo - 1 (from the right one;
  // row x at index x - width).  One step per pass needs one row of each, a pass of T steps needs T;
  // every pass pushes all kFFHalo boundary rows it owns, so that any kind of pass can follow.
  const uint8_t* slot_lo = nullptr;  // nullptr: outside the world => no fire
  const uint8_t* slot_hi = nullptr;
  const unsigned long long* flag_lo = nullptr;  // wait until *flag >= wait_epoch before reading the slot
  const unsigned long long* flag_hi = nullptr;
  unsigned long long wait_epoch = 0;
  uint8_t* push_lo = nullptr;  // left neighbour's slot_hi: my new row x < kFFHalo goes to index x
  uint8_t* push_hi = nullptr;  // right neighbour's slot_lo: my new row x >= width - kFFHalo to index kFFHalo + x - width
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
    // (the same bit test as inside the span, so that bytes outside the model's alphabet — which the
    // fused passes below step by their two low bits — behave alike on both sides of a warp seam)
    if (lane == 0 && y0 > 0) left = (__ldcg(p - 1) & 3u) == FF_BURNING;
    if (lane == 31 && y0 + 16 < height) right = (__ldcg(p + 16) & 3u) == FF_BURNING;
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
  if (x < kFFHalo && ex.push_lo) *reinterpret_cast<uint4*>(ex.push_lo + (int64_t)x * height + y0) = ov4;
  if (x >= width - kFFHalo && ex.push_hi)
    *reinterpret_cast<uint4*>(ex.push_hi + (int64_t)(kFFHalo + x - width) * height + y0) = ov4;
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
    if (x < 0) return ex.slot_lo ? ex.slot_lo + (int64_t)(kFFHalo - 1) * height : nullptr;  // row -1
    return ex.slot_hi;  // row width; nullptr outside the world
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
// T steps per pass (temporal blocking) on bit planes.  K5 above sits at ~81 % of the HBM roofline,
// so the only way to more cell-updates per second is fewer bytes per update: this kernel reads
// step t once and writes step t+T, keeping the T-1 steps between in registers — 1 B read + 1 B
// written per cell for T updates.  The rule is deterministic and None cells never change, so T
// fused steps equal T launches of K5 bit for bit (tests/test_gpu_grid.py, test_gpu_gridstrips.py).
//
// A first version kept K5's byte-parallel arithmetic for two steps and ran into the integer pipe
// (LOP3/SHF issue at half rate on sm_100): 0.278 ms per step at 32768^2 against K5's 0.405, with
// HBM at 60 %.  The rule only looks at two bits of a cell — GREEN 01, BURNING 10, BURNED 11, and
// None (0xFF) behaves like BURNED: inert — so here a lane turns its 32 cells of a row into two
// 32-bit planes (one multiply per four cells gathers a bit of each byte into a nibble), steps
// those, and only the row leaving the pipeline is spread back to bytes:
//   burning = p1 & ~p0;  mask = burning | burning << 1 | burning >> 1  (neighbour lanes supply the
//   edge bits by shuffle);  fire = mask(x-1) | mask(x) | mask(x+1);  ignite = p0 & ~p1 & fire;
//   p1' = p1 | ignite;  p0' = (p0 & ~ignite) | burning.
// One time level costs 9 integer instructions per 32 cells and six registers, so eight levels fit.
// The original bytes wait in a shared-memory ring (T+1 rows per warp) for their row to come out:
// out = (byte & 0xFC) | code restores None and leaves every other bit to the planes.
//
// A result cell of step t+T depends on step t within T cells, so a warp that loads 32 x 32 cells
// of y per row vouches for the inner 30 x 32 = 960: lanes 0 and 31 only feed their neighbours
// (6 % redundant work, no byte-wide halo loads), and a tile of rows reads T extra rows above and
// below.

// bit `b` of the four bytes of each word -> one nibble; word k lands in bits 4k .. 4k+3
template <int BIT>
__device__ __forceinline__ uint32_t ff_plane(const uint32_t (&w)[8]) {
  // (x & 0x01010101) * 0x10204080 puts byte j's bit 0 at bit 28 + j (no two partial products
  // meet); for bit 1 the operand is twice that and the multiplier half
  const uint32_t M = 0x01010101u << BIT, K = 0x10204080u >> BIT;
  uint32_t p = 0;
#pragma unroll
  for (int k = 7; k >= 0; --k) p = __funnelshift_l((w[k] & M) * K, p, 4);  // p = p << 4 | top nibble
  return p;
}

// planes -> the bytes of word k: (nibble * 0x00204081) & 0x01010101 spreads bit j to byte j
__device__ __forceinline__ uint32_t ff_spread(uint32_t v, uint32_t p0, uint32_t p1, int k) {
  const uint32_t e0 = (((p0 >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
  const uint32_t e1 = (((p1 >> (4 * k)) & 0xFu) * 0x00408102u) & 0x02020202u;
  return (v & 0xFCFCFCFCu) | e0 | e1;  // live cells are 1, 2, 3 (upper bits clear); None is 0xFC | 3
}

template <int T>
struct FFPipe {
  // stepper L takes rows from time level L to L+1.  At a stage of parity Q, with row r arriving:
  // hm[L][Q] = mask of row r-2, hm[L][Q^1] = mask of row r-1, (p0, p1)[L][Q] = planes of row r-1;
  // the roles swap with the parity, so nothing is ever moved between registers.
  uint32_t hm[T][2], p0[T][2], p1[T][2];
};

template <int T, int Q>
__device__ __forceinline__ void ff_pipe_stage(FFPipe<T>& pp, uint32_t& a0, uint32_t& a1) {
#pragma unroll
  for (int L = 0; L < T; ++L) {
    const uint32_t b = a1 & ~a0;  // burning cells of the arriving row
    const uint32_t bl = __shfl_up_sync(0xffffffffu, b, 1), br = __shfl_down_sync(0xffffffffu, b, 1);
    // cells beyond the warp's span: lanes 0 and 31 see their own bits there — garbage that moves one
    // cell per level and never leaves those two lanes (T <= 32)
    const uint32_t hnew = b | __funnelshift_l(bl, b, 1) | __funnelshift_r(b, br, 1);
    const uint32_t fire = pp.hm[L][Q] | pp.hm[L][Q ^ 1] | hnew;
    const uint32_t c0 = pp.p0[L][Q], c1 = pp.p1[L][Q];
    const uint32_t ignite = c0 & ~c1 & fire;
    const uint32_t n1 = c1 | ignite;
    const uint32_t n0 = (c0 & ~ignite) | (c1 & ~c0);
    pp.hm[L][Q] = hnew;
    pp.p0[L][Q ^ 1] = a0;
    pp.p1[L][Q ^ 1] = a1;
    a0 = n0;
    a1 = n1;
  }
}

#ifndef KG_FFT_MINB
#define KG_FFT_MINB 6  // resident 128-thread blocks per SM the T <= 4 kernels are compiled for
#endif
#ifndef KG_FFT_MINB8
#define KG_FFT_MINB8 5  // ... and the eight-level kernel (its pipeline alone is 48 registers)
#endif
constexpr int kFFTSpan = 960;  // cells of y a warp produces per row

// `width` = rows of this grid / strip.  A strip's rows -T .. -1 and width .. width+T-1 come from its
// inbox and its new boundary rows also go to the line neighbours' inboxes, as in K5; the host
// guarantees that all pushed rows of a side lie in the first / last row tile (the tiles that publish).
template <int T>
static __global__ void __launch_bounds__(128, T == 8 ? KG_FFT_MINB8 : KG_FFT_MINB)
forest_fire_u8_multi_kernel(const uint8_t* __restrict__ rd, uint8_t* __restrict__ wr, int32_t width,
                            int32_t height, int32_t rows_per_tile, FFExchange ex) {
  static_assert(T >= 2 && T <= kFFHalo && T % 2 == 0, "an even number of steps, at most the slot depth");
  __shared__ uint4 ring[4][T + 1][2][32];  // the bytes of the rows inside the pipeline
  grid_dep_wait();  // the read buffer is the previous pass's output (dependent launch, common.cuh)
  const int lane = threadIdx.x & 31;
  const int64_t span = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  const int64_t y0 = span * kFFTSpan + (int64_t)(lane - 1) * 32;
  const bool live = span * kFFTSpan < height;  // warp-uniform
  const bool in0 = y0 >= 0 && y0 < height, in1 = y0 >= 0 && y0 + 16 < height;
  const bool owner = lane != 0 && lane != 31;
  // boundary tiles first (see K5): tile 0, then the last one, then the interior
  const uint32_t ntiles = gridDim.y;
  const uint32_t tile = blockIdx.y == 0 ? 0u : (blockIdx.y == 1 ? ntiles - 1 : blockIdx.y - 1);
  const int32_t x_begin = (int32_t)tile * rows_per_tile;
  const int32_t x_end = min(width, x_begin + rows_per_tile);
  if (x_begin >= width) return;
  const bool first = x_begin == 0, last = x_end == width;
  if (first && ex.flag_lo) ff_wait_flag(ex.flag_lo, ex.wait_epoch, ex.err);
  if (last && ex.flag_hi) ff_wait_flag(ex.flag_hi, ex.wait_epoch, ex.err);

  auto rowp = [&](int32_t x) -> const uint8_t* {  // step-t row x as stored; nullptr where the world has no row
    if (x >= 0 && x < width) return rd + (uint64_t)x * (uint64_t)height;
    if (x < 0) return (ex.slot_lo && x >= -kFFHalo) ? ex.slot_lo + (int64_t)(kFFHalo + x) * height : nullptr;
    return (ex.slot_hi && x - width < kFFHalo) ? ex.slot_hi + (int64_t)(x - width) * height : nullptr;
  };
  const uint4 kNone4 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
  if (live) {
    uint4 (*my_ring)[2][32] = ring[threadIdx.x >> 5];
    const int32_t x_first = x_begin - T, x_last = x_end - 1 + T;  // rows that arrive
    uint4 pre0 = kNone4, pre1 = kNone4;  // one row ahead of the arithmetic
    auto fetch = [&](int32_t x) {
      const uint8_t* row = x <= x_last ? rowp(x) : nullptr;
      pre0 = (row != nullptr && in0) ? __ldcg(reinterpret_cast<const uint4*>(row + y0)) : kNone4;
      pre1 = (row != nullptr && in1) ? __ldcg(reinterpret_cast<const uint4*>(row + y0 + 16)) : kNone4;
    };
    FFPipe<T> pp;
#pragma unroll
    for (int L = 0; L < T; ++L) {  // rows before the first one: inert, and outside every stored cell's cone
      pp.hm[L][0] = pp.hm[L][1] = 0u;
      pp.p0[L][0] = pp.p0[L][1] = pp.p1[L][0] = pp.p1[L][1] = 0xFFFFFFFFu;
    }
    int slot = 0;  // ring slot of the arriving row = (xin - x_first) mod (T + 1)
    auto stage = [&](auto parity, int32_t xin) {
      constexpr int Q = decltype(parity)::value;
      const uint4 q0 = pre0, q1 = pre1;
      fetch(xin + 1);
      my_ring[slot][0][lane] = q0;
      my_ring[slot][1][lane] = q1;
      const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      uint32_t a0 = ff_plane<0>(w), a1 = ff_plane<1>(w);
      ff_pipe_stage<T, Q>(pp, a0, a1);  // (a0, a1): row xin - T at step t + T
      slot = slot == T ? 0 : slot + 1;  // now the slot of row xin - T (the next to be overwritten)
      const int32_t xo = xin - T;
      if (xo >= x_begin && xo < x_end && owner) {
        const uint4 v0 = my_ring[slot][0][lane], v1 = my_ring[slot][1][lane];
        const uint4 o0 = make_uint4(ff_spread(v0.x, a0, a1, 0), ff_spread(v0.y, a0, a1, 1),
                                    ff_spread(v0.z, a0, a1, 2), ff_spread(v0.w, a0, a1, 3));
        const uint4 o1 = make_uint4(ff_spread(v1.x, a0, a1, 4), ff_spread(v1.y, a0, a1, 5),
                                    ff_spread(v1.z, a0, a1, 6), ff_spread(v1.w, a0, a1, 7));
        const uint64_t off = (uint64_t)xo * (uint64_t)height + y0;
        if (in0) *reinterpret_cast<uint4*>(wr + off) = o0;
        if (in1) *reinterpret_cast<uint4*>(wr + off + 16) = o1;
        if (xo < kFFHalo && ex.push_lo) {
          uint8_t* q = ex.push_lo + (int64_t)xo * height + y0;
          if (in0) *reinterpret_cast<uint4*>(q) = o0;
          if (in1) *reinterpret_cast<uint4*>(q + 16) = o1;
        }
        if (xo >= width - kFFHalo && ex.push_hi) {
          uint8_t* q = ex.push_hi + (int64_t)(kFFHalo + xo - width) * height + y0;
          if (in0) *reinterpret_cast<uint4*>(q) = o0;
          if (in1) *reinterpret_cast<uint4*>(q + 16) = o1;
        }
      }
    };
    fetch(x_first);
    // an even number of rows arrives (rows of the tile + 2T) unless the tile is odd: the second half
    // of the last pair then sees an inert row and stores nothing
    for (int32_t xin = x_first; xin <= x_last; xin += 2) {
      stage(std::integral_constant<int, 0>{}, xin);
      stage(std::integral_constant<int, 1>{}, xin + 1);
    }
  }
  if (first && ex.push_lo) ff_publish(ex.done + 0, gridDim.x, ex.push_flag_lo, ex.push_epoch);
  if (last && ex.push_hi) ff_publish(ex.done + 1, gridDim.x, ex.push_flag_hi, ex.push_epoch);
}

// Rows per tile of a T-step pass over `own` rows of `height` cells.  A tile of r rows costs r + 2T row
// stages (T extra rows read above and below); the blocks run in waves of kNumSMs x resident blocks.  Pick
// the r in [4T, 32T] with the fewest stages on the critical path, waves x (r + 2T); ties go to the larger
// tile.  Measured on B200 (tools/ff_tile_probe.py): 4096 x 32768 (one of eight strips): 32 rows 0.0111,
// 48 0.0126 (just over one wave), 64 0.0100 ms per step; 32768^2: 64 rows 0.0656, 107 0.0639, 128 0.0662.
inline int ff_multi_rows_per_tile(int T, int64_t own, int64_t height) {
  const int64_t spans = (height + kFFTSpan - 1) / kFFTSpan;
  const int64_t bx = (spans + 3) / 4;
  const int64_t capacity = (int64_t)kNumSMs * (T == 8 ? KG_FFT_MINB8 : KG_FFT_MINB);
  int best = 4 * T;
  int64_t best_cost = INT64_MAX;
  for (int r = 4 * T; r <= 32 * T; ++r) {
    const int64_t blocks = bx * ((own + r - 1) / r);
    const int64_t cost = ((blocks + capacity - 1) / capacity) * (r + 2 * T);
    if (cost <= best_cost) {
      best_cost = cost;
      best = r;
    }
  }
  return best;
}

}  // namespace kg
