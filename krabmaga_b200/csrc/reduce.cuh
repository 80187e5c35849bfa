// Device-side reductions over agent segments: what a model's output columns (`explore`'s FrameRow
// fields, src/explore/model_exploration.rs:160-190, src/lib.rs:1781-1800 write_csv) or a plot!
// series are computed from, without downloading the population.  Deterministic: every block sums a
// fixed chunk in a fixed order (f64), a second kernel adds the chunk partials sequentially.
#pragma once
#include "common.cuh"

namespace kg {

constexpr int kRedVals = 8;       // sum x, sum y, sum last_d.x, sum last_d.y, sum |last_d|, sum x^2, sum y^2, (spare)
constexpr int kRedThreads = 256;
constexpr uint32_t kRedChunk = 8192;  // agents per block

// blockIdx.y = segment (a replica, or the whole field), blockIdx.x = chunk of that segment
static __global__ void __launch_bounds__(kRedThreads)
reduce_partial_kernel(const float4* __restrict__ pv, uint64_t seg_len, uint32_t nchunk, double* __restrict__ partial) {
  __shared__ double sm[kRedThreads / 32][kRedVals];
  const uint64_t seg0 = (uint64_t)blockIdx.y * seg_len;
  const uint64_t c0 = (uint64_t)blockIdx.x * kRedChunk;
  const uint64_t c1 = min(seg_len, c0 + (uint64_t)kRedChunk);
  double v[kRedVals] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (uint64_t i = c0 + threadIdx.x; i < c1; i += kRedThreads) {
    const float4 q = pv[seg0 + i];
    v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
    v[4] += sqrt((double)q.z * q.z + (double)q.w * q.w);
    v[5] += (double)q.x * q.x; v[6] += (double)q.y * q.y;
  }
#pragma unroll
  for (int k = 0; k < kRedVals; ++k)
    for (int o = 16; o; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < kRedVals; ++k) sm[wid][k] = v[k];
  __syncthreads();
  if (threadIdx.x < kRedVals) {
    double s = 0;
    for (int w = 0; w < kRedThreads / 32; ++w) s += sm[w][threadIdx.x];
    partial[((uint64_t)blockIdx.y * nchunk + blockIdx.x) * kRedVals + threadIdx.x] = s;
  }
}
static __global__ void reduce_final_kernel(uint32_t nseg, uint32_t nchunk, const double* __restrict__ partial,
                                           double* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg * kRedVals) return;
  const uint32_t seg = t / kRedVals, k = t % kRedVals;
  double s = 0;
  for (uint32_t c = 0; c < nchunk; ++c) s += partial[((uint64_t)seg * nchunk + c) * kRedVals + k];
  out[t] = s;
}

// the same two kernels writing their nseg x kRedVals sums to a DEVICE row (no copy, no sync): one point
// of a plot! series recorded while the steps keep running
inline int reduce_segments_dev(const float4* pv, uint32_t nseg, uint64_t seg_len, double* partial, double* out_dev,
                               cudaStream_t s) {
  const uint32_t nchunk = (uint32_t)std::max<uint64_t>(1, (seg_len + kRedChunk - 1) / kRedChunk);
  reduce_partial_kernel<<<dim3(nchunk, nseg), kRedThreads, 0, s>>>(pv, seg_len, nchunk, partial);
  reduce_final_kernel<<<(nseg * kRedVals + 127) / 128, 128, 0, s>>>(nseg, nchunk, partial, out_dev);
  launch_counter().fetch_add(2, std::memory_order_relaxed);
  return KG_OK;
}

// out_host[nseg][kRedVals]; scratch is grown on demand and owned by the caller's handle
inline int reduce_segments(const float4* pv, uint32_t nseg, uint64_t seg_len, double** scratch, size_t* scratch_bytes,
                           double* out_host, cudaStream_t s) {
  if (nseg == 0) return KG_OK;
  const uint32_t nchunk = (uint32_t)std::max<uint64_t>(1, (seg_len + kRedChunk - 1) / kRedChunk);
  const size_t need = ((size_t)nseg * nchunk + nseg) * kRedVals * sizeof(double);
  if (*scratch_bytes < need) {
    if (*scratch) cudaFree(*scratch);
    *scratch = nullptr;
    *scratch_bytes = 0;
    KG_CUDA(cudaMalloc(scratch, need));
    *scratch_bytes = need;
  }
  double* partial = *scratch;
  double* out = partial + (size_t)nseg * nchunk * kRedVals;
  reduce_partial_kernel<<<dim3(nchunk, nseg), kRedThreads, 0, s>>>(pv, seg_len, nchunk, partial);
  reduce_final_kernel<<<(nseg * kRedVals + 127) / 128, 128, 0, s>>>(nseg, nchunk, partial, out);
  launch_counter().fetch_add(2, std::memory_order_relaxed);
  KG_CUDA(cudaMemcpyAsync(out_host, out, (size_t)nseg * kRedVals * sizeof(double), cudaMemcpyDeviceToHost, s));
  KG_CUDA(cudaStreamSynchronize(s));
  return KG_OK;
}

}  // namespace kg
