//! `test-harness` feature: the counter-based random stream shared with libkrabgpu and its oracle
//!.  Philox4x32-10 (Salmon et al., SC'11), key = (seed lo, seed hi),
//! counter = (agent id, step lo, step hi, domain); domain 0 = State::init, 1 = Agent::step,
//! 2 = grid cells.  u32 -> f32 as rand 0.9.2's StandardUniform: 24 high bits * 2^-24.
//!
//! NOT COMPILED IN THIS REPOSITORY (no Rust toolchain in the build image).  The same function is
//! implemented in oracle/philox.hpp and krabmaga_b200/csrc/common.cuh and pinned there to the
//! Random123 known-answer vectors that the unit test below repeats.

const M0: u32 = 0xD251_1F53;
const M1: u32 = 0xCD9E_8D57;
const W0: u32 = 0x9E37_79B9;
const W1: u32 = 0xBB67_AE85;

pub const DOMAIN_INIT: u32 = 0;
pub const DOMAIN_STEP: u32 = 1;
pub const DOMAIN_GRID: u32 = 2;

#[inline]
pub fn philox4x32_10(mut c: [u32; 4], mut k: [u32; 2]) -> [u32; 4] {
    for _ in 0..10 {
        let p0 = (M0 as u64) * (c[0] as u64);
        let p1 = (M1 as u64) * (c[2] as u64);
        c = [
            ((p1 >> 32) as u32) ^ c[1] ^ k[0],
            p1 as u32,
            ((p0 >> 32) as u32) ^ c[3] ^ k[1],
            p0 as u32,
        ];
        k = [k[0].wrapping_add(W0), k[1].wrapping_add(W1)];
    }
    c
}

/// rand 0.9.2 `StandardUniform` for f32
#[inline]
pub fn u01(u: u32) -> f32 {
    (u >> 8) as f32 * (1.0 / 16_777_216.0)
}

/// The two draws of `Bird::step` (tests/model/flockers/bird.rs:113-117)
#[inline]
pub fn step_draws(seed: u64, id: u32, step: u64) -> (f32, f32) {
    let r = philox4x32_10([id, step as u32, (step >> 32) as u32, DOMAIN_STEP], [seed as u32, (seed >> 32) as u32]);
    (u01(r[0]), u01(r[1]))
}

/// The two draws of `Flocker::init` (tests/model/flockers/state.rs:42-45)
#[inline]
pub fn init_draws(seed: u64, id: u32) -> (f32, f32) {
    let r = philox4x32_10([id, 0, 0, DOMAIN_INIT], [seed as u32, (seed >> 32) as u32]);
    (u01(r[0]), u01(r[1]))
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn random123_known_answers() {
        assert_eq!(philox4x32_10([0; 4], [0; 2]), [0x6627_e8d5, 0xe169_c58d, 0xbc57_ac4c, 0x9b00_dbd8]);
        assert_eq!(philox4x32_10([0xffff_ffff; 4], [0xffff_ffff; 2]),
                   [0x408f_276d, 0x41c8_3b0e, 0xa20b_c7c6, 0x6d54_51fd]);
        assert_eq!(philox4x32_10([0x243f_6a88, 0x85a3_08d3, 0x1319_8a2e, 0x0370_7344], [0xa409_3822, 0x299f_31d0]),
                   [0xd16c_fe09, 0x94fd_cceb, 0x5001_e420, 0x2412_6ea1]);
    }
}
