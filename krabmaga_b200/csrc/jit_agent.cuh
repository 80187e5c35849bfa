// A model's own Agent::step on the device: the body of `fn step(&mut self, state)` (src/engine/agent.rs:7-16;
// the Flockers fixture's is tests/model/flockers/bird.rs:39-155) given as two CUDA C snippets — what happens per
// neighbour the field query returns, and what turns the sums into the agent's new state — compiled at run time
// (jit.cuh) around the library's window walk, random stream and write log.  The generated kernel is the generic K4
// (step_boids_kernel) with the snippets in place of boids_pair / boids_finish: reference-shaped walk
// (field_2d.rs:401-437 / :485-514), any geometry, both query kinds.
//
// The device-side helpers are restated here as source text (NVRTC cannot include this library's headers, which
// pull in host headers); tests/test_gpu_custom_step.py pins the restatement: Bird::step written as snippets must
// reproduce the built-in generic kernel bit for bit.
#pragma once
#include <string>

namespace kg {
namespace jit {

inline std::string agent_step_source(const char* pair, const char* finish, bool may_stop) {
  static const char* prelude = R"SRC(
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
struct Geom { float w, h, disc; int toroidal; int max_x, max_y, dw, dh; uint32_t ncells; };
struct Consts { float c[16]; };
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ int f2i_sat(float v) { return __float2int_rz(v); }
__device__ __forceinline__ int t_transform(int n, int size) { return n >= 0 ? n % size : (n % size) + size; }
__device__ __forceinline__ float toroidal_transform(float v, float dim) {
  if (v >= 0.0f && v < dim) return v;
  float r = fmodf(v, dim);
  if (r < 0.0f) r = fadd(r, dim);
  return r;
}
__device__ __forceinline__ float toroidal_distance(float a, float b, float dim) {
  float d0 = fsub(a, b);
  if (fabsf(d0) <= fmul(dim, 0.5f)) return d0;
  float d = fsub(toroidal_transform(a, dim), toroidal_transform(b, dim));
  if (fmul(d, 2.0f) > dim) return fsub(d, dim);
  if (fmul(d, 2.0f) < -dim) return fadd(d, dim);
  return d;
}
__device__ __forceinline__ float distance(float ax, float ay, float bx, float by, const Geom& g) {
  float dx, dy;
  if (g.toroidal) { dx = toroidal_distance(ax, bx, g.w); dy = toroidal_distance(ay, by, g.h); }
  else { dx = fsub(ax, bx); dy = fsub(ay, by); }
  return fsqrt(fadd(fmul(dx, dx), fmul(dy, dy)));
}
__device__ __forceinline__ int check_circle(int bx, int by, const Geom& g, float lx, float ly, float dis) {
  float nwx = fmul((float)bx, g.disc), nwy = fmul((float)by, g.disc);
  float ney = fminf(fadd(nwy, g.disc), g.h);
  float swx = fminf(fadd(nwx, g.disc), g.w);
  float d0 = distance(nwx, nwy, lx, ly, g), d1 = distance(nwx, ney, lx, ly, g);
  float d2 = distance(swx, nwy, lx, ly, g), d3 = distance(swx, ney, lx, ly, g);
  if (d0 <= dis && d1 <= dis && d2 <= dis && d3 <= dis) return 1;
  if (d0 > dis && d1 > dis && d2 > dis && d3 > dis) return -1;
  return 0;
}
struct Philox4 { uint32_t v[4]; };
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}
__device__ __forceinline__ float u01_f32(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }
)SRC";
  std::string s = may_stop ? "#define KG_MAY_STOP 1\n" : "#define KG_MAY_STOP 0\n";
  s += prelude;
  // the snippets become the bodies of two inline functions
  s += "struct Self { uint32_t sid; float sx, sy, sa, sb; };\n";
  s += "__device__ __forceinline__ void kg_pair(const Geom& g, const Consts& K, const Self& S, uint32_t oid, float ox, float oy,\n"
       "    float oa, float ob, float dx, float dy, float* acc, int& cnt) {\n"
       "  const float* c = K.c; const float w = g.w, h = g.h; const uint32_t sid = S.sid;\n"
       "  const float sx = S.sx, sy = S.sy, sa = S.sa, sb = S.sb;\n"
       "  (void)c; (void)w; (void)h; (void)sid; (void)sx; (void)sy; (void)sa; (void)sb; (void)oid; (void)ox; (void)oy; (void)oa; (void)ob;\n"
       "  { ";
  s += pair;
  s += " }\n}\n";
  s += "__device__ __forceinline__ void kg_finish(const Geom& g, const Consts& K, const Self& S, const float* acc, int cnt,\n"
       "    uint32_t nvec, float u0, float u1, float& nx, float& ny, float& na, float& nb, bool& stopped) {\n"
       "  const float* c = K.c; const float w = g.w, h = g.h; const uint32_t sid = S.sid;\n"
       "  const float sx = S.sx, sy = S.sy, sa = S.sa, sb = S.sb;\n"
       "  (void)c; (void)w; (void)h; (void)sid; (void)sx; (void)sy; (void)sa; (void)sb; (void)acc; (void)cnt; (void)nvec; (void)u0; (void)u1;\n"
       "  { ";
  s += finish;
  s += " }\n}\n";
  s += R"SRC(
extern "C" __global__ void __launch_bounds__(128)
kg_agent_step(Geom g, uint32_t n, const uint32_t* __restrict__ rid, const float4* __restrict__ rpv,
              const uint32_t* __restrict__ cs, uint32_t* __restrict__ wid, float4* __restrict__ wpv,
              uint32_t* __restrict__ count, int* err, float dist, int exact, uint64_t seed, uint64_t step, Consts K) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 me = rpv[i];
  Self S; S.sid = rid[i]; S.sx = me.x; S.sy = me.y; S.sa = me.z; S.sb = me.w;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int cnt = 0;
  uint32_t nvec = 0;
  // the window walk of field_2d.rs:401-437 (exact) / :485-514 (relaxed), in the reference's order
  if (!(dist <= 0.0f)) {  // field_2d.rs:393 / :481 (NaN falls through, as in the reference)
    int dd = f2i_sat(floorf(fdiv(dist, g.disc)));
    int cx = f2i_sat(floorf(fdiv(me.x, g.disc)));
    int cy = f2i_sat(floorf(fdiv(me.y, g.disc)));
    int min_i = cx - dd, max_i = cx + dd, min_j = cy - dd, max_j = cy + dd;
    if (g.toroidal) {
      min_i = max(0, min_i); max_i = min(max_i, g.max_x - 1);
      min_j = max(0, min_j); max_j = min(max_j, g.max_y - 1);
    }
    if (!exact && g.toroidal) {
      // clamped window: every column's cells min_j..max_j are ONE contiguous slice of the sorted buffer
      if (min_j <= max_j)
        for (int ci = min_i; ci <= max_i; ++ci)
          for (uint32_t k = cs[ci * g.dh + min_j]; k < cs[ci * g.dh + max_j + 1]; ++k) {
            const float4 o = rpv[k];
            ++nvec;
            kg_pair(g, K, S, rid[k], o.x, o.y, o.z, o.w, toroidal_distance(me.x, o.x, g.w), toroidal_distance(me.y, o.y, g.h),
                    acc, cnt);
          }
    } else
    for (int ci = min_i; ci <= max_i; ++ci) {
      const int bx = t_transform(ci, g.max_x);
      for (int cj = min_j; cj <= max_j; ++cj) {
        const int by = t_transform(cj, g.max_y);
        const int check = exact ? check_circle(bx, by, g, me.x, me.y, dist) : 1;
        if (check < 0) continue;
        const uint32_t cell = (uint32_t)(bx * g.dh + by);
        for (uint32_t k = cs[cell]; k < cs[cell + 1]; ++k) {
          const float4 o = rpv[k];
          if (check == 0 && !(distance(me.x, me.y, o.x, o.y, g) <= dist)) continue;
          ++nvec;
          kg_pair(g, K, S, rid[k], o.x, o.y, o.z, o.w, toroidal_distance(me.x, o.x, g.w), toroidal_distance(me.y, o.y, g.h),
                  acc, cnt);
        }
      }
    }
  }
  const Philox4 r = philox4x32_10(S.sid, (uint32_t)step, (uint32_t)(step >> 32), 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
  float nx = me.x, ny = me.y, na = me.z, nb = me.w;
  bool stopped = false;
  kg_finish(g, K, S, acc, cnt, nvec, u01_f32(r.v[0]), u01_f32(r.v[1]), nx, ny, na, nb, stopped);
  if (KG_MAY_STOP && stopped) {            // Agent::is_stopped (agent.rs:18): not rescheduled, gone after lazy_update
    wid[i] = 0xFFFFFFFFu;
    return;
  }
  wid[i] = S.sid;
  wpv[i] = make_float4(nx, ny, na, nb);
  const int ncx = f2i_sat(floorf(fdiv(nx, g.disc))), ncy = f2i_sat(floorf(fdiv(ny, g.disc)));
  const uint32_t c = (uint32_t)ncx * (uint32_t)g.dh + (uint32_t)ncy;
  if ((int)c >= 0 && c < g.ncells) atomicAdd(&count[c], 1u);   // K1 fused: histogram of the write log
  else atomicOr(err, 1);
}
)SRC";
  return s;
}

}  // namespace jit
}  // namespace kg
