//! `krabmaga::engine::fields::gpu` — the third sibling of the two `Field2D` variants selected in
//! `src/engine/fields/field_2d.rs:26-27` / `:267`, backed by libkrabgpu (include/krabgpu.h).
//!
//! NOT COMPILED IN THIS REPOSITORY: the build image has no Rust toolchain.  The file
//! is the binding a krABMaga maintainer adds under `#[cfg(feature = "gpu")]`; it is kept in sync
//! with the header by hand and reviewed against `krabmaga_b200/_abi.py`, which binds the same
//! symbols and IS exercised by the test-suite.
//!
//! Conventions preserved from the reference: `&self` + interior mutability for per-step reads and
//! writes, `&mut self` only for `update`/`lazy_update` (field_2d.rs:838 vs :905); failures panic;
//! values are returned by copy.
#![allow(non_camel_case_types)]
use crate::engine::fields::field::Field;
use crate::engine::location::{Int2D, Real2D};
use std::cell::Cell;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct kg_field2d(c_void);
#[repr(C)]
pub struct kg_grid(c_void);
#[repr(C)]
pub struct kg_batch(c_void);
#[repr(C)]
pub struct kg_gridstrip(c_void);
#[repr(C)]
pub struct kg_strip(c_void);

/// `KgBoidsParams` of include/krabgpu.h (tests/model/flockers/bird.rs:12-17, :41)
#[repr(C)]
#[derive(Clone, Copy)]
pub struct KgBoidsParams {
    pub cohesion: f32,
    pub avoidance: f32,
    pub randomness: f32,
    pub consistency: f32,
    pub momentum: f32,
    pub jump: f32,
    pub radius: f32,
    pub exact_query: i32,
    pub seed: u64,
    pub step: u64,
}

/// `KgLifeRule` of include/krabgpu.h: Agent::is_stopped (agent.rs:18) + the births of State::after_step
#[repr(C)]
#[derive(Clone, Copy)]
pub struct KgLifeRule {
    pub death_prob: f32,
    pub birth_prob: f32,
    pub crowd_limit: u32,
    pub reserved: u32,
}

#[link(name = "krabgpu")]
extern "C" {
    fn kg_last_error() -> *const c_char;
    fn kg_field2d_create(w: f32, h: f32, d: f32, toroidal: c_int, capacity: u64, device: c_int,
                         out: *mut *mut kg_field2d) -> c_int;
    fn kg_field2d_destroy(f: *mut kg_field2d) -> c_int;
    fn kg_field2d_sync(f: *mut kg_field2d) -> c_int;
    fn kg_field2d_set_object_locations(f: *mut kg_field2d, n: u64, id: *const u32, x: *const f32,
                                       y: *const f32, ldx: *const f32, ldy: *const f32) -> c_int;
    fn kg_field2d_remove_object_location(f: *mut kg_field2d, id: u32, x: f32, y: f32) -> c_int;
    fn kg_field2d_lazy_update(f: *mut kg_field2d) -> c_int;
    fn kg_field2d_nagents(f: *mut kg_field2d, out: *mut u64) -> c_int;
    fn kg_field2d_num_objects(f: *mut kg_field2d, which: c_int, out: *mut u64) -> c_int;
    fn kg_field2d_download(f: *mut kg_field2d, which: c_int, cap: u64, id: *mut u32, x: *mut f32,
                           y: *mut f32, ldx: *mut f32, ldy: *mut f32, cell: *mut i32,
                           n_out: *mut u64) -> c_int;
    fn kg_field2d_neighbors(f: *mut kg_field2d, nq: u64, qx: *const f32, qy: *const f32, dist: f32,
                            mode: c_int, offsets: *mut u64, ids: *mut u32, cap: u64,
                            total: *mut u64) -> c_int;
    fn kg_field2d_neighbors_agents(f: *mut kg_field2d, nq: u64, qx: *const f32, qy: *const f32, dist: f32,
                                   mode: c_int, offsets: *mut u64, ids: *mut u32, x: *mut f32, y: *mut f32,
                                   last_dx: *mut f32, last_dy: *mut f32, cap: u64, total_out: *mut u64) -> c_int;
    fn kg_field2d_num_objects_at_locations(f: *mut kg_field2d, nq: u64, x: *const f32,
                                           y: *const f32, out: *mut u32) -> c_int;
    fn kg_field2d_step_boids(f: *mut kg_field2d, p: *const KgBoidsParams) -> c_int;
    fn kg_field2d_step_boids_host_ordered(f: *mut kg_field2d, p: *const KgBoidsParams, n: u64,
                                          id_in: *const u32, x_in: *const f32, y_in: *const f32,
                                          dx_in: *const f32, dy_in: *const f32, x_out: *mut f32,
                                          y_out: *mut f32, dx_out: *mut f32, dy_out: *mut f32) -> c_int;
    fn kg_field2d_init_flockers(f: *mut kg_field2d, n: u64, seed: u64) -> c_int;
    fn kg_field2d_set_next_id(f: *mut kg_field2d, next_id: u32) -> c_int;
    fn kg_field2d_step_boids_life(f: *mut kg_field2d, p: *const KgBoidsParams, life: *const KgLifeRule,
                                  n_stopped: *mut u64, n_born: *mut u64) -> c_int;
    fn kg_grid_create(w: i32, h: i32, elem: c_int, none: u32, device: c_int,
                      out: *mut *mut kg_grid) -> c_int;
    fn kg_grid_destroy(g: *mut kg_grid) -> c_int;
    fn kg_grid_set_values(g: *mut kg_grid, n: u64, x: *const i32, y: *const i32,
                          v: *const c_void) -> c_int;
    fn kg_grid_get_values(g: *mut kg_grid, which: c_int, n: u64, x: *const i32, y: *const i32,
                          out: *mut c_void) -> c_int;
    fn kg_grid_lazy_update(g: *mut kg_grid) -> c_int;
    fn kg_grid_update(g: *mut kg_grid) -> c_int;
    fn kg_grid_step_stencil(g: *mut kg_grid, rule: c_int) -> c_int;
    // batched replicas for explore_parallel! (src/explore/model_exploration.rs:354-423)
    fn kg_batch_create(w: f32, h: f32, d: f32, toroidal: c_int, replicas: u32, agents_per_replica: u32,
                       device: c_int, out: *mut *mut kg_batch) -> c_int;
    fn kg_batch_destroy(b: *mut kg_batch) -> c_int;
    fn kg_batch_set_params(b: *mut kg_batch, first: u32, n: u32, p: *const KgBoidsParams) -> c_int;
    fn kg_batch_init_flockers(b: *mut kg_batch) -> c_int;
    fn kg_batch_lazy_update(b: *mut kg_batch) -> c_int;
    fn kg_batch_run_boids(b: *mut kg_batch, first_step: u64, nsteps: u64) -> c_int;
    fn kg_batch_download(b: *mut kg_batch, id: *mut u32, x: *mut f32, y: *mut f32, ldx: *mut f32,
                         ldy: *mut f32, cell: *mut i32) -> c_int;
    // row strips of a DenseNumberGrid2D<u8> over the GPUs of one box
    fn kg_gridstrip_create(width: i32, height: i32, rank: c_int, nranks: c_int, device: c_int,
                           out: *mut *mut kg_gridstrip) -> c_int;
    fn kg_gridstrip_destroy(s: *mut kg_gridstrip) -> c_int;
    fn kg_gridstrip_connect_local(s: *mut kg_gridstrip, left: *mut kg_gridstrip,
                                  right: *mut kg_gridstrip) -> c_int;
    fn kg_gridstrip_init_forest_fire(s: *mut kg_gridstrip, density: f32, seed: u64) -> c_int;
    fn kg_gridstrip_prepare(s: *mut kg_gridstrip) -> c_int;
    fn kg_gridstrip_run_stencil(s: *mut kg_gridstrip, rule: c_int, nsteps: u64) -> c_int;
    fn kg_gridstrip_download(s: *mut kg_gridstrip, own_rows: *mut u8) -> c_int;
    // x-strips of one Field2D world over the GPUs of one box (precedent: kdtree_mpi.rs:705-790)
    fn kg_strip_create(w: f32, h: f32, d: f32, toroidal: c_int, radius: f32, rank: c_int,
                       nranks: c_int, capacity: u64, halo_capacity: u64, migrate_capacity: u64,
                       device: c_int, out: *mut *mut kg_strip) -> c_int;
    fn kg_strip_destroy(s: *mut kg_strip) -> c_int;
    fn kg_strip_columns(s: *mut kg_strip, own_x0: *mut i32, own_x1: *mut i32, halo_l: *mut i32,
                        halo_r: *mut i32, dh: *mut i32) -> c_int;
    fn kg_strip_connect_local(s: *mut kg_strip, left: *mut kg_strip, right: *mut kg_strip) -> c_int;
    fn kg_strip_ipc_export(s: *mut kg_strip, handle: *mut c_void) -> c_int;
    fn kg_strip_connect_ipc(s: *mut kg_strip, left: *const c_void, right: *const c_void) -> c_int;
    fn kg_strip_init_flockers(s: *mut kg_strip, n_global: u64, seed: u64) -> c_int;
    fn kg_strip_upload(s: *mut kg_strip, n: u64, id: *const u32, x: *const f32, y: *const f32,
                       ldx: *const f32, ldy: *const f32) -> c_int;
    fn kg_strip_prepare(s: *mut kg_strip) -> c_int;
    fn kg_strip_step_boids(s: *mut kg_strip, p: *const KgBoidsParams) -> c_int;
    fn kg_strip_run_boids(s: *mut kg_strip, p: *const KgBoidsParams, nsteps: u64) -> c_int;
    fn kg_strip_sync(s: *mut kg_strip) -> c_int;
    fn kg_strip_stats(s: *mut kg_strip, n_owned: *mut u64, migrants_in: *mut u64,
                      migrants_out: *mut u64, halo_left: *mut u64, halo_right: *mut u64,
                      launches: *mut u64) -> c_int;
    fn kg_strip_download(s: *mut kg_strip, cap: u64, id: *mut u32, x: *mut f32, y: *mut f32,
                         ldx: *mut f32, ldy: *mut f32, n_out: *mut u64) -> c_int;
}

/// Non-zero status -> panic, matching the reference's `expect`/index panics.
fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(kg_last_error()) }.to_string_lossy().into_owned();
        panic!("krabgpu error {rc}: {msg}");
    }
}

/// The agent payload the shipped kernels understand: the Flockers `Bird`
/// (tests/model/flockers/bird.rs:19-25).
pub trait BoidLike: Copy {
    fn id(&self) -> u32;
    fn pos(&self) -> Real2D;
    fn last_d(&self) -> Real2D;
    fn from_parts(id: u32, pos: Real2D, last_d: Real2D) -> Self;
}

/// GPU `Field2D`: same method names as `Field2D<O>` (field_2d.rs:269-921).
pub struct Field2D<O: BoidLike> {
    h: *mut kg_field2d,
    pub width: f32,
    pub height: f32,
    pub discretization: f32,
    pub toroidal: bool,
    pending: std::cell::RefCell<Vec<(O, Real2D)>>, // set_object_location calls of the current step
    step: Cell<u64>,
}
// One handle is used by one thread at a time (State: Send, state.rs:45)
unsafe impl<O: BoidLike> Send for Field2D<O> {}

impl<O: BoidLike> Field2D<O> {
    /// `Field2D::new(w, h, d, t)` (field_2d.rs:304-322) + device capacity
    pub fn new(w: f32, h: f32, d: f32, t: bool, capacity: u64, device: i32) -> Self {
        let mut h_: *mut kg_field2d = std::ptr::null_mut();
        check(unsafe { kg_field2d_create(w, h, d, t as c_int, capacity, device, &mut h_) });
        Field2D { h: h_, width: w, height: h, discretization: d, toroidal: t,
                  pending: Default::default(), step: Cell::new(0) }
    }
    /// field_2d.rs:838-846 — buffered on the host, flushed as ONE boundary crossing
    pub fn set_object_location(&self, object: O, loc: Real2D) {
        self.pending.borrow_mut().push((object, loc));
    }
    fn flush(&self) {
        let p = std::mem::take(&mut *self.pending.borrow_mut());
        if p.is_empty() { return; }
        // `loc` decides the bag (field_2d.rs:839), as in the reference; the payload travels with it
        let id: Vec<u32> = p.iter().map(|(o, _)| o.id()).collect();
        let x: Vec<f32> = p.iter().map(|(_, l)| l.x).collect();
        let y: Vec<f32> = p.iter().map(|(_, l)| l.y).collect();
        let dx: Vec<f32> = p.iter().map(|(o, _)| o.last_d().x).collect();
        let dy: Vec<f32> = p.iter().map(|(o, _)| o.last_d().y).collect();
        check(unsafe { kg_field2d_set_object_locations(self.h, p.len() as u64, id.as_ptr(),
              x.as_ptr(), y.as_ptr(), dx.as_ptr(), dy.as_ptr()) });
    }
    /// field_2d.rs:885-898
    pub fn remove_object_location(&self, object: O, loc: Real2D) {
        self.flush();
        check(unsafe { kg_field2d_remove_object_location(self.h, object.id(), loc.x, loc.y) });
    }
    /// Many query points in ONE boundary crossing; the neighbours themselves come back (the
    /// reference returns `Vec<O>` by copy, field_2d.rs:386, :472).  The handle keeps its device
    /// scratch between calls, so a model that still queries per agent pays no allocation.
    pub fn neighbors_batch(&self, locs: &[Real2D], dist: f32, exact: bool) -> Vec<Vec<O>> {
        let nq = locs.len();
        let qx: Vec<f32> = locs.iter().map(|l| l.x).collect();
        let qy: Vec<f32> = locs.iter().map(|l| l.y).collect();
        let mut offs = vec![0u64; nq + 1];
        let mut cap = (64 * nq.max(4)) as u64;
        loop {
            let m = cap as usize;
            let mut ids = vec![0u32; m];
            let (mut x, mut y, mut dx, mut dy) = (vec![0f32; m], vec![0f32; m], vec![0f32; m], vec![0f32; m]);
            let mut total = 0u64;
            let rc = unsafe { kg_field2d_neighbors_agents(self.h, nq as u64, qx.as_ptr(), qy.as_ptr(), dist,
                              exact as c_int, offs.as_mut_ptr(), ids.as_mut_ptr(), x.as_mut_ptr(),
                              y.as_mut_ptr(), dx.as_mut_ptr(), dy.as_mut_ptr(), cap, &mut total) };
            if rc == -3 && total > cap { cap = total; continue; } // KG_E_CAPACITY: retry with room
            check(rc);
            return (0..nq).map(|q| {
                (offs[q] as usize..offs[q + 1] as usize).map(|k| {
                    O::from_parts(ids[k], Real2D { x: x[k], y: y[k] }, Real2D { x: dx[k], y: dy[k] })
                }).collect()
            }).collect();
        }
    }
    /// field_2d.rs:386-440
    pub fn get_neighbors_within_distance(&self, loc: Real2D, dist: f32) -> Vec<O> {
        self.neighbors_batch(&[loc], dist, true).pop().unwrap()
    }
    /// field_2d.rs:472-516
    pub fn get_neighbors_within_relax_distance(&self, loc: Real2D, dist: f32) -> Vec<O> {
        self.neighbors_batch(&[loc], dist, false).pop().unwrap()
    }
    /// field_2d.rs:806-811
    pub fn num_objects_at_location(&self, loc: Real2D) -> usize {
        let mut out = 0u32;
        check(unsafe { kg_field2d_num_objects_at_locations(self.h, 1, &loc.x, &loc.y, &mut out) });
        out as usize
    }
    /// Field2D.nagents (field_2d.rs:277)
    pub fn nagents(&self) -> usize {
        let mut n = 0u64;
        check(unsafe { kg_field2d_nagents(self.h, &mut n) });
        n as usize
    }
    /// iter_objects order (field_2d.rs:594-626)
    pub fn objects(&self) -> Vec<O> {
        let mut n = 0u64;
        check(unsafe { kg_field2d_num_objects(self.h, 0, &mut n) });
        let m = n as usize;
        let (mut id, mut x, mut y, mut dx, mut dy) =
            (vec![0u32; m], vec![0f32; m], vec![0f32; m], vec![0f32; m], vec![0f32; m]);
        check(unsafe { kg_field2d_download(self.h, 0, n, id.as_mut_ptr(), x.as_mut_ptr(),
              y.as_mut_ptr(), dx.as_mut_ptr(), dy.as_mut_ptr(), std::ptr::null_mut(), &mut n) });
        (0..m).map(|i| O::from_parts(id[i], Real2D { x: x[i], y: y[i] },
                                     Real2D { x: dx[i], y: dy[i] })).collect()
    }
    /// All agents' `Agent::step` of the Flockers model (bird.rs:39-155) in one launch.
    pub fn step_boids(&self, mut p: KgBoidsParams, schedule_step: u64) {
        self.flush();
        p.step = schedule_step;
        check(unsafe { kg_field2d_step_boids(self.h, &p) });
    }
    /// One whole step for a model that keeps its birds in a `Vec` on the host (the schedule's agent list):
    /// positions and last directions go up, every `Bird::step` runs, and the results come back at the same
    /// indices (bird i has id i, as `State::init` numbers them, state.rs:47) — 16 bytes per bird each way.
    pub fn step_boids_in_place(&self, mut p: KgBoidsParams, schedule_step: u64, x: &mut [f32], y: &mut [f32],
                               last_dx: &mut [f32], last_dy: &mut [f32]) {
        let n = x.len();
        assert!(y.len() == n && last_dx.len() == n && last_dy.len() == n);
        p.step = schedule_step;
        check(unsafe { kg_field2d_step_boids_host_ordered(self.h, &p, n as u64, std::ptr::null(), x.as_ptr(),
              y.as_ptr(), last_dx.as_ptr(), last_dy.as_ptr(), x.as_mut_ptr(), y.as_mut_ptr(),
              last_dx.as_mut_ptr(), last_dy.as_mut_ptr()) });
    }
    /// Dynamic population: every Bird's `step` + `Agent::is_stopped` (agent.rs:18), then the births
    /// of `State::after_step`; returns (stopped, born).  `lazy_update` compacts the dead away.
    pub fn step_boids_life(&self, mut p: KgBoidsParams, life: KgLifeRule, schedule_step: u64) -> (u64, u64) {
        self.flush();
        p.step = schedule_step;
        let (mut stopped, mut born) = (0u64, 0u64);
        check(unsafe { kg_field2d_step_boids_life(self.h, &p, &life, &mut stopped, &mut born) });
        (stopped, born)
    }
    pub fn set_next_id(&self, next_id: u32) { check(unsafe { kg_field2d_set_next_id(self.h, next_id) }); }
    /// `State::init` of the fixture (state.rs:41-56) on the device
    pub fn init_flockers(&self, n: u64, seed: u64) {
        check(unsafe { kg_field2d_init_flockers(self.h, n, seed) });
    }
    pub fn sync(&self) { check(unsafe { kg_field2d_sync(self.h) }); }
}

impl<O: BoidLike> Field for Field2D<O> {
    fn update(&mut self) {}
    /// field_2d.rs:905-921
    fn lazy_update(&mut self) {
        self.flush();
        check(unsafe { kg_field2d_lazy_update(self.h) });
        self.step.set(self.step.get() + 1);
    }
}
impl<O: BoidLike> Drop for Field2D<O> {
    fn drop(&mut self) { unsafe { kg_field2d_destroy(self.h) }; }
}

/// GPU `DenseNumberGrid2D<u8>` (dense_number_grid_2d.rs:90-561); `None` is 0xFF on the device.
pub struct DenseNumberGrid2D {
    h: *mut kg_grid,
    pub width: i32,
    pub height: i32,
}
unsafe impl Send for DenseNumberGrid2D {}
impl DenseNumberGrid2D {
    pub fn new(width: i32, height: i32, device: i32) -> Self {
        let mut h: *mut kg_grid = std::ptr::null_mut();
        check(unsafe { kg_grid_create(width, height, 1, 0xFF, device, &mut h) });
        DenseNumberGrid2D { h, width: width.abs(), height: height.abs() }
    }
    /// :350-354
    pub fn get_value(&self, loc: &Int2D) -> Option<u8> {
        let mut v = 0xFFu8;
        check(unsafe { kg_grid_get_values(self.h, 0, 1, &loc.x, &loc.y, &mut v as *mut u8 as *mut c_void) });
        if v == 0xFF { None } else { Some(v) }
    }
    /// :492-495
    pub fn set_value_location(&self, value: u8, loc: &Int2D) {
        check(unsafe { kg_grid_set_values(self.h, 1, &loc.x, &loc.y, &value as *const u8 as *const c_void) });
    }
    /// one Forest-Fire step for every live cell (get_value + set_value_location per cell)
    pub fn step_forest_fire(&self) { check(unsafe { kg_grid_step_stencil(self.h, 0) }); }
}
impl Field for DenseNumberGrid2D {
    fn lazy_update(&mut self) { check(unsafe { kg_grid_lazy_update(self.h) }); }
    fn update(&mut self) { check(unsafe { kg_grid_update(self.h) }); }
}
impl Drop for DenseNumberGrid2D {
    fn drop(&mut self) { unsafe { kg_grid_destroy(self.h) }; }
}

/// Proxy agent: `Schedule::step` (schedule.rs:347-413) needs at least one scheduled agent or it
/// takes the empty-queue branch (:357-365).  Its `step` is every Bird's `step`.
#[derive(Clone)]
pub struct Flock {
    pub params: KgBoidsParams,
}
// impl Agent for Flock {
//     fn step(&mut self, state: &mut dyn State) {
//         let s = state.as_any().downcast_ref::<Flocker>().unwrap();
//         s.field1.step_boids(self.params, s.step);
//     }
// }
// and in the model's State (tests/model/flockers/state.rs:41-60):
//     fn init(&mut self, schedule: &mut Schedule) {
//         self.field1.init_flockers(self.initial_flockers as u64, SEED);
//         schedule.schedule_repeating(Box::new(Flock { params }), 0., 0);
//     }
//     fn update(&mut self, step: u64) { self.step = step; self.field1.lazy_update(); }

/// R independent Flockers replicas advanced together — what one rayon task of `explore_parallel!`
/// (src/explore/model_exploration.rs:387-420) does for one configuration, for all of them at once.
pub struct FlockerBatch {
    h: *mut kg_batch,
    pub replicas: u32,
    pub agents: u32,
}
unsafe impl Send for FlockerBatch {}
impl FlockerBatch {
    pub fn new(dim: (f32, f32), disc: f32, replicas: u32, agents: u32, params: &[KgBoidsParams],
               device: i32) -> Self {
        let mut h: *mut kg_batch = std::ptr::null_mut();
        check(unsafe { kg_batch_create(dim.0, dim.1, disc, 1, replicas, agents, device, &mut h) });
        check(unsafe { kg_batch_set_params(h, 0, params.len() as u32, params.as_ptr()) });
        FlockerBatch { h, replicas, agents }
    }
    /// `State::init` + the first `State::update` of every replica, then `nstep` x `Schedule::step`
    pub fn simulate(&self, nstep: u64) {
        check(unsafe { kg_batch_init_flockers(self.h) });
        check(unsafe { kg_batch_lazy_update(self.h) });
        check(unsafe { kg_batch_run_boids(self.h, 0, nstep) });
    }
}
impl Drop for FlockerBatch {
    fn drop(&mut self) { unsafe { kg_batch_destroy(self.h) }; }
}

/// One Field2D world cut into x-strips, one per GPU of the box, all owned by this process (the
/// multi-process form wires the same handles with `kg_strip_ipc_export`/`kg_strip_connect_ipc`
/// after an MPI/`torchrun`-style rendezvous).  Neighbour strips exchange halo agents and migrants
/// by peer stores; there is no host round trip and no collective.
pub struct StripWorld {
    strips: Vec<*mut kg_strip>,
}
unsafe impl Send for StripWorld {}
impl StripWorld {
    /// `n_global` sizes the buffers for a roughly uniform population (what
    /// `krabmaga_b200/strips.py::default_capacities` does): agents per cell column x the widest
    /// strip, x `disc_dist` columns x 4 for a halo, x 2 columns for one step's migrants, each with
    /// 50 % slack.  Overflow is reported as `KG_E_CAPACITY` at the next sync, never silently.
    pub fn new(w: f32, h: f32, d: f32, radius: f32, devices: &[i32], n_global: u64) -> Self {
        let g = devices.len() as u64;
        let max_x = (w / d).ceil().max(1.0) as u64;
        let dd = ((radius / d).floor() as u64).max(1);
        let per_col = n_global as f64 / max_x as f64;
        let widest = (max_x + g - 1) / g + 1; // the last strip also owns the padding column
        let capacity = ((per_col * widest as f64 * 1.5) as u64 + 1024).min(n_global.max(1024) + 1024);
        let halo = (per_col * (dd * 4) as f64 * 1.5) as u64 + 1024;
        let migrate = (per_col * 2.0 * 1.5) as u64 + 1024;
        let mut strips = Vec::new();
        for (r, dev) in devices.iter().enumerate() {
            let mut s: *mut kg_strip = std::ptr::null_mut();
            check(unsafe { kg_strip_create(w, h, d, 1, radius, r as c_int, g as c_int, capacity, halo,
                                           migrate, *dev, &mut s) });
            strips.push(s);
        }
        for r in 0..strips.len() {
            let left = strips[(r + strips.len() - 1) % strips.len()];
            let right = strips[(r + 1) % strips.len()];
            check(unsafe { kg_strip_connect_local(strips[r], left, right) });
        }
        StripWorld { strips }
    }
    /// `State::init` of the fixture for ids 0..n: every strip keeps the agents it owns
    pub fn init_flockers(&self, n: u64, seed: u64) {
        for s in &self.strips { check(unsafe { kg_strip_init_flockers(*s, n, seed) }); }
        for s in &self.strips { check(unsafe { kg_strip_prepare(*s) }); }
    }
    /// One `Schedule::step` of the whole world: step `step` is issued for every strip before the
    /// next one for any (the strips of one process share the host thread).
    pub fn step_boids(&self, mut p: KgBoidsParams, step: u64) {
        p.step = step;
        for s in &self.strips { check(unsafe { kg_strip_step_boids(*s, &p) }); }
    }
    pub fn sync(&self) {
        for s in &self.strips { check(unsafe { kg_strip_sync(*s) }); }
    }
    /// Owned agents of every strip, strips in x order (= `iter_objects` order of the whole field)
    pub fn objects<O: BoidLike>(&self) -> Vec<O> {
        let mut all = Vec::new();
        for s in &self.strips {
            // six distinct out-parameters: one `&mut` each (no aliasing)
            let (mut n, mut mig_in, mut mig_out, mut halo_l, mut halo_r, mut launches) =
                (0u64, 0u64, 0u64, 0u64, 0u64, 0u64);
            check(unsafe { kg_strip_stats(*s, &mut n, &mut mig_in, &mut mig_out, &mut halo_l, &mut halo_r,
                                          &mut launches) });
            let m = n as usize;
            let (mut id, mut x, mut y, mut dx, mut dy) =
                (vec![0u32; m], vec![0f32; m], vec![0f32; m], vec![0f32; m], vec![0f32; m]);
            check(unsafe { kg_strip_download(*s, n, id.as_mut_ptr(), x.as_mut_ptr(), y.as_mut_ptr(),
                                             dx.as_mut_ptr(), dy.as_mut_ptr(), &mut n) });
            all.extend((0..m).map(|i| O::from_parts(id[i], Real2D { x: x[i], y: y[i] },
                                                    Real2D { x: dx[i], y: dy[i] })));
        }
        all
    }
}
impl Drop for StripWorld {
    fn drop(&mut self) { for s in &self.strips { unsafe { kg_strip_destroy(*s) }; } }
}
