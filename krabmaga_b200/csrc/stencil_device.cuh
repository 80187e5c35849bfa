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
  const uint8_t* halo_lo = nullptr;  // row x = -1 (nullptr: outside the world => no fire)
  const uint8_t* halo_hi = nullptr;  // row x = width
  const unsigned long long* flag_lo = nullptr;  // wait until *flag >= wait_epoch before reading halo
  const unsigned long long* flag_hi = nullptr;
  unsigned long long wait_epoch = 0;
  uint8_t* push_lo = nullptr;  // neighbour inbox row that receives my new row 0
  uint8_t* push_hi = nullptr;  // ... my new row width-1
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
  if (x == 0 && ex.push_lo) *reinterpret_cast<uint4*>(ex.push_lo + y0) = ov4;
  if (x == width - 1 && ex.push_hi) *reinterpret_cast<uint4*>(ex.push_hi + y0) = ov4;
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

}  // namespace kg
