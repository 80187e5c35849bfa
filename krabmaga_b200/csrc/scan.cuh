// Exclusive prefix sum over u32 cell counts (K2 of the cell-list rebuild).
// Reduce-then-scan in three launches: per-tile sums -> scan of tile sums -> per-tile scan.
// HBM-bound: 4 B read (reduce) + 4 B read + 4 B write per cell.
#pragma once
#include "common.cuh"

namespace kg {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;  // per thread, as 4 x uint4
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (NT threads); returns exclusive prefix, total in *total
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan_t(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[NT / 32];
  __shared__ uint32_t block_total;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < NT / 32 ? warp_sums[lane] : 0;
    uint32_t wi = warp_incl_scan(w, lane);
    if (lane < NT / 32) warp_sums[lane] = wi - w;
    if (lane == NT / 32 - 1) block_total = wi;
  }
  __syncthreads();
  uint32_t r = incl - v + warp_sums[wid];
  if (total) *total = block_total;
  __syncthreads();
  return r;
}
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  return block_excl_scan_t<kScanThreads>(v, total);
}

static __global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ tile_sums) {
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems / 4; ++k) {
    uint64_t i = base + ((uint64_t)k * kScanThreads + threadIdx.x) * 4;
    if (i + 3 < n) {
      uint4 v = *reinterpret_cast<const uint4*>(in + i);
      s += v.x + v.y + v.z + v.w;
    } else {
      for (int j = 0; j < 4; ++j)
        if (i + j < n) s += in[i + j];
    }
  }
  uint32_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of the tile sums (any count), total appended at [nt]
static __global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(uint32_t* __restrict__ tile_sums, uint32_t nt) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nt; base += kScanThreads) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nt ? tile_sums[i] : 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(v, &total);
    if (i < nt) tile_sums[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) tile_sums[nt] = carry;
}

// out[i] = exclusive prefix of in[0..i); out[n] = total.  in and out may alias exactly.
static __global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const uint32_t* in, uint64_t n, const uint32_t* __restrict__ tile_sums,
                  uint32_t* out) {
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  // thread owns 16 consecutive items so that the local scan is sequential in registers
  uint64_t i0 = base + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems / 4; ++k) {
    uint64_t i = i0 + k * 4;
    if (i + 3 < n) {
      uint4 q = *reinterpret_cast<const uint4*>(in + i);
      v[k * 4 + 0] = q.x; v[k * 4 + 1] = q.y; v[k * 4 + 2] = q.z; v[k * 4 + 3] = q.w;
    } else {
      for (int j = 0; j < 4; ++j) v[k * 4 + j] = (i + j < n) ? in[i + j] : 0;
    }
  }
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) s += v[k];
  uint32_t ex = block_excl_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems / 4; ++k) {
    uint64_t i = i0 + k * 4;
    uint4 q;
    q.x = ex; ex += v[k * 4 + 0];
    q.y = ex; ex += v[k * 4 + 1];
    q.z = ex; ex += v[k * 4 + 2];
    q.w = ex; ex += v[k * 4 + 3];
    if (i + 3 < n) {
      *reinterpret_cast<uint4*>(out + i) = q;
    } else {
      if (i + 0 < n) out[i + 0] = q.x;
      if (i + 1 < n) out[i + 1] = q.y;
      if (i + 2 < n) out[i + 2] = q.z;
    }
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_sums[gridDim.x];
}

inline uint32_t scan_num_tiles(uint64_t n) { return (uint32_t)((n + kScanTile - 1) / kScanTile); }

// in: n counts (16-byte aligned), out: n+1 offsets, tile_sums: scan_num_tiles(n)+1 scratch
inline void exclusive_scan_u32(const uint32_t* in, uint64_t n, uint32_t* out, uint32_t* tile_sums,
                               cudaStream_t s) {
  uint32_t nt = scan_num_tiles(n);
  scan_reduce_kernel<<<nt, kScanThreads, 0, s>>>(in, n, tile_sums);
  scan_tiles_kernel<<<1, kScanThreads, 0, s>>>(tile_sums, nt);
  scan_apply_kernel<<<nt, kScanThreads, 0, s>>>(in, n, tile_sums, out);
}

// ------------------------------------------------------------------------------------------
// Single-pass exclusive scan with decoupled look-back (Merrill & Garland 2016): one launch,
// 4 B read + 4 B write per cell.  Tiles are handed out in arrival order by an atomic ticket so a
// tile only ever waits on tiles that already started.  The per-tile status words carry an epoch
// so neither they nor the ticket counter need clearing between launches:
//   status[t] = (epoch << 34) | (kind << 32) | value,  kind 1 = tile aggregate, 2 = inclusive prefix
#ifndef KG_LB_THREADS
#define KG_LB_THREADS 256
#endif
constexpr int kLbThreads = KG_LB_THREADS;
#ifndef KG_LB_ITEMS
#define KG_LB_ITEMS 16  // cells per thread of the look-back scan (8, 16, 32 measure 12.0, 10.1, 10.6 us at 361k cells)
#endif
constexpr int kLbItems = KG_LB_ITEMS;
constexpr int kLbTile = kLbThreads * kLbItems;

struct LookbackState {
  unsigned long long* status = nullptr;  // [max tiles]
  uint32_t* ticket = nullptr;            // monotonically increasing across launches
  uint32_t ticket_base = 0;              // host mirror: tickets consumed so far
  uint32_t epoch = 0;
  uint32_t max_tiles = 0;
};

static __global__ void __launch_bounds__(kLbThreads)
scan_lookback_kernel(const uint32_t* in, uint64_t n, uint32_t* out,
                     volatile unsigned long long* status, uint32_t* ticket, uint32_t ticket_base,
                     uint32_t epoch, uint32_t num_tiles, uint32_t base, const uint32_t* base_sub) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_prefix;
  grid_dep_wait();  // `in` is the previous kernel's histogram
#if KG_SCATTER_EARLY
  grid_dep_launch();  // the scatter behind this scan may already load the log (everything older than the scan is done)
#endif
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u) - ticket_base;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t i0 = (uint64_t)tile * kLbTile + (uint64_t)threadIdx.x * kLbItems;
  uint32_t v[kLbItems];
#pragma unroll
  for (int k = 0; k < kLbItems / 4; ++k) {
    uint64_t i = i0 + k * 4;
    if (i + 3 < n) {
      uint4 q = *reinterpret_cast<const uint4*>(in + i);
      v[k * 4 + 0] = q.x; v[k * 4 + 1] = q.y; v[k * 4 + 2] = q.z; v[k * 4 + 3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[k * 4 + j] = (i + j < n) ? in[i + j] : 0;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kLbItems; ++k) s += v[k];
  uint32_t total;
  uint32_t ex = block_excl_scan_t<kLbThreads>(s, &total);
  const unsigned long long tag = (unsigned long long)epoch << 34;
  if (threadIdx.x == 0) {
    if (tile == 0) {
      status[0] = tag | (2ull << 32) | total;
      s_prefix = 0;
    } else {
      status[tile] = tag | (1ull << 32) | total;
    }
  }
  if (tile != 0 && threadIdx.x < 32) {
    // warp-wide look-back: lane l inspects predecessor (tile-1-l) of the current window
    uint32_t prefix = 0;
    int32_t base = (int32_t)tile - 1;
    for (;;) {
      int32_t t = base - (int32_t)threadIdx.x;
      unsigned long long w = 0;
      uint32_t kind = 2;  // tiles before 0 count as a finished prefix of 0
      uint32_t val = 0;
      if (t >= 0) {
        do {
          w = status[t];
        } while ((w >> 34) != epoch || ((w >> 32) & 3ull) == 0);
        kind = (uint32_t)((w >> 32) & 3ull);
        val = (uint32_t)w;
      }
      unsigned done = __ballot_sync(0xffffffffu, kind == 2);
      // lanes up to and including the first finished prefix contribute
      int first = done ? __ffs(done) - 1 : 32;
      uint32_t c = ((int)threadIdx.x <= first) ? val : 0;
      for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      prefix += c;
      if (done) break;
      base -= 32;
    }
    if (threadIdx.x == 0) {
      status[tile] = tag | (2ull << 32) | (unsigned long long)(prefix + total);
      s_prefix = prefix;
    }
  }
  __syncthreads();
  ex += s_prefix + base - (base_sub ? *base_sub : 0u);  // base_sub: a device word written by an earlier kernel
#pragma unroll
  for (int k = 0; k < kLbItems / 4; ++k) {
    uint64_t i = i0 + k * 4;
    uint4 q;
    q.x = ex; ex += v[k * 4 + 0];
    q.y = ex; ex += v[k * 4 + 1];
    q.z = ex; ex += v[k * 4 + 2];
    q.w = ex; ex += v[k * 4 + 3];
    if (i + 3 < n) {
      *reinterpret_cast<uint4*>(out + i) = q;
    } else {
      if (i + 0 < n) out[i + 0] = q.x;
      if (i + 1 < n) out[i + 1] = q.y;
      if (i + 2 < n) out[i + 2] = q.z;
    }
  }
  if (tile == num_tiles - 1 && threadIdx.x == kLbThreads - 1) out[n] = ex;  // grand total
}

inline uint32_t lookback_num_tiles(uint64_t n) { return (uint32_t)((n + kLbTile - 1) / kLbTile); }

inline int lookback_init(LookbackState& st, uint64_t max_n, cudaStream_t s) {
  st.max_tiles = lookback_num_tiles(max_n) + 1;
  KG_CUDA(cudaMalloc(&st.status, (size_t)st.max_tiles * 8));
  KG_CUDA(cudaMalloc(&st.ticket, 4));
  KG_CUDA(cudaMemsetAsync(st.status, 0, (size_t)st.max_tiles * 8, s));
  KG_CUDA(cudaMemsetAsync(st.ticket, 0, 4, s));
  st.ticket_base = 0;
  st.epoch = 0;
  return KG_OK;
}
inline void lookback_destroy(LookbackState& st) {
  cudaFree(st.status);
  cudaFree(st.ticket);
  st = LookbackState{};
}
// out[i] = base + sum in[0..i), out[n] = base + total; in/out may alias exactly
// `dependent` launches it as a programmatic dependent of the stream's previous kernel
// `base_sub` (optional): device word subtracted from `base` at run time
inline void exclusive_scan_lookback(LookbackState& st, const uint32_t* in, uint64_t n, uint32_t* out,
                                    cudaStream_t s, uint32_t base = 0, bool dependent = false,
                                    const uint32_t* base_sub = nullptr) {
  uint32_t nt = lookback_num_tiles(n);
  st.epoch = (st.epoch + 1) & 0x3FFFFFFFu;
  if (st.epoch == 0) st.epoch = 1;
  if (dependent)
    launch_pdl(scan_lookback_kernel, dim3(nt), dim3(kLbThreads), s, in, n, out, st.status, st.ticket,
               st.ticket_base, st.epoch, nt, base, base_sub);
  else
    scan_lookback_kernel<<<nt, kLbThreads, 0, s>>>(in, n, out, st.status, st.ticket, st.ticket_base,
                                                   st.epoch, nt, base, base_sub);
  st.ticket_base += nt;
}

}  // namespace kg
