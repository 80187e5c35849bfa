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

// block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  __shared__ uint32_t block_total;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint32_t wi = warp_incl_scan(w, lane);
    if (lane < kScanThreads / 32) warp_sums[lane] = wi - w;
    if (lane == kScanThreads / 32 - 1) block_total = wi;
  }
  __syncthreads();
  uint32_t r = incl - v + warp_sums[wid];
  if (total) *total = block_total;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads)
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
__global__ void __launch_bounds__(kScanThreads)
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
__global__ void __launch_bounds__(kScanThreads)
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

}  // namespace kg
