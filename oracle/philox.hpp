// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under krabmaga_b200/ may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker.
//
// Philox4x32-10 counter-based generator (Salmon et al., SC'11; Random123 v1.14
// `philox4x32_R(10, ctr, key)`).  The reference draws from `rand::rng()`
// (tests/model/flockers/bird.rs:113-117, state.rs:42-45), which is OS-seeded and
// not reproducible; BASELINE.json's north_star replaces it on both sides with a
// Philox stream keyed by (seed, agent id, step).  Pinned against the Random123
// known-answer vectors in tests/test_oracle_philox.py.
//
// Stream layout shared with krabmaga_b200/csrc/philox.cuh (independent code):
//   key = (seed_lo, seed_hi)
//   ctr = (agent_id_or_cell_lo, step_lo, step_hi | cell_hi<<?, domain)   see DESIGN.md §RNG
//   domain 0: State::init positions  (r1 = out[0], r2 = out[1])
//   domain 1: Agent::step randomness (r1 = out[0], r2 = out[1])
//   domain 2: DenseNumberGrid2D initial fill (out[0])
// u32 -> f32 follows rand 0.9.2 `StandardUniform for f32`: 24 high bits * 2^-24.
#pragma once
#include <cstdint>

namespace oracle {

struct Philox4 {
  uint32_t v[4];
};

inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                             uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// rand 0.9 StandardUniform<f32>: (u >> 8) as f32 * 2^-24  in [0,1)
inline float u01_f32(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

enum PhiloxDomain : uint32_t { DOMAIN_INIT = 0, DOMAIN_STEP = 1, DOMAIN_GRID = 2, DOMAIN_LIFE = 3 };

}  // namespace oracle
