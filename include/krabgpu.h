/* krabgpu.h — C ABI of libkrabgpu.so, the B200 (sm_100a) implementation of krABMaga's
 * agent-step hot path.  This is the boundary a `krabmaga::engine::fields::gpu` Rust module
 * binds with `extern "C"` (INTEGRATION.md shows the binding); the same symbols are driven
 * from Python (ctypes, krabmaga_b200/_abi.py) and C++ (krabmaga_b200/host/krabmaga_gpu.hpp).
 *
 * The reference (krABMaga 0.6.1, pure Rust) has no FFI layer: every entry point below names
 * the reference method it replaces (paths relative to the reference crate root).
 *
 * Conventions
 *   - every call returns int: KG_OK (0) or a negative KG_E_* code; kg_last_error() gives the
 *     text of the calling thread's last failure.  Nothing aborts or throws across the boundary;
 *     the Rust shim turns non-zero into panic! to match the reference's failure behaviour.
 *   - handles are opaque, own all their device memory and one CUDA stream, may be moved between
 *     threads (`Send`) but calls on one handle must be serialised by the caller; distinct
 *     handles are independent.
 *   - host pointers are borrowed for the duration of the call only.  `*_dev` variants take
 *     device pointers on the handle's device.
 *   - device work is asynchronous on the handle's stream; calls that return data to the host
 *     (download, queries, counts) and kg_*_sync() synchronise and report deferred device errors
 *     (e.g. KG_E_OOB raised by a kernel).
 *   - there is no CPU fallback: without a CUDA device every create call fails with KG_E_CUDA.
 */
#ifndef KRABGPU_H
#define KRABGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_ABI_VERSION 1

enum {
  KG_OK = 0,
  KG_E_CUDA = -1,     /* CUDA runtime failure (no device, launch error, out of memory) */
  KG_E_INVALID = -2,  /* bad argument / bad handle state */
  KG_E_CAPACITY = -3, /* more agents than the handle was created for, or output buffer too small */
  KG_E_OOB = -4       /* coordinate outside the bag grid: the reference would index-panic
                         (field_2d.rs:840-842, dense_number_grid_2d.rs:493-494) */
};

/* which buffer of a double-buffered field */
enum { KG_BUF_READ = 0, KG_BUF_WRITE = 1 };
/* neighbour query kind */
enum {
  KG_QUERY_RELAX = 0, /* Field2D::get_neighbors_within_relax_distance  field_2d.rs:472-516 */
  KG_QUERY_EXACT = 1  /* Field2D::get_neighbors_within_distance        field_2d.rs:386-440 */
};
/* order of agents inside one bag of the read buffer after lazy_update */
enum {
  KG_ORDER_ANY = 0,      /* whatever the scatter's atomics produced (fastest) */
  KG_ORDER_CANONICAL = 1 /* ascending id: makes f32 sums reproducible bit for bit */
};

const char* kg_last_error(void);
int kg_abi_version(void);
/* number of CUDA devices visible, or KG_E_CUDA */
int kg_device_count(void);
/* device self-test: n random (a0, a1, den) triples through the shared-reciprocal division of the
 * fast K4 versus IEEE division; *mismatches must come back 0 */
int kg_selftest_div(int device, uint64_t n, uint64_t seed, uint64_t* mismatches);
/* page-locked host memory for the e2e path (cudaHostAlloc / cudaFreeHost) */
int kg_host_alloc(size_t bytes, void** out);
int kg_host_free(void* p);

/* ------------------------------------------------------------------------------------------
 * Field2D  (src/engine/fields/field_2d.rs:269-921, default variant)
 * Agent payload = the Flockers `Bird` (tests/model/flockers/bird.rs:19-25) as SoA:
 * id u32, pos (x,y) f32, last_d (dx,dy) f32.
 * ------------------------------------------------------------------------------------------ */
typedef struct kg_field2d kg_field2d;

/* Field2D::new(w,h,d,t)  field_2d.rs:304-322.  `capacity` = most agents either buffer holds. */
int kg_field2d_create(float w, float h, float discretization, int toroidal, uint64_t capacity,
                      int device, kg_field2d** out);
int kg_field2d_destroy(kg_field2d* f);
int kg_field2d_sync(kg_field2d* f);
/* dw, dh (field_2d.rs:317-318) and max_x, max_y (:487-488) */
int kg_field2d_dims(kg_field2d* f, int32_t* dw, int32_t* dh, int32_t* max_x, int32_t* max_y);
int kg_field2d_set_order(kg_field2d* f, int order);
/* Which K4 (fused neighbour gather + Bird::step) kg_field2d_step_boids launches.  All variants but
 * KG_K4_COLTILE return identical bits; the parity tests and bench.py switch between them.
 *   KG_K4_AUTO          packed kernel (FADD2/FMUL2/FFMA2 candidate loop) when the geometry allows
 *                       (toroidal + relaxed query + window << world), else the generic kernel
 *   KG_K4_GENERIC       generic window walk (any geometry, both query kinds)
 *   KG_K4_FAST_SCALAR   the scalar fast kernel (one f32 lane per instruction)
 *   KG_K4_PACKED_BY_ID  packed kernel, self exclusion by id comparison even when ids are unique
 *   KG_K4_TILED         block per run of cells of one cell row; the candidates' column slices are
 *                       staged in shared memory by cp.async.bulk (TMA) copies, agents dealt to
 *                       lanes by window length (relaxed 3x3 query only, else the packed kernel)
 *   KG_K4_COLTILE       block per chunk of consecutive agents of one cell column; the window region is
 *                       staged row-major in shared memory so that a 3x3 window is one contiguous range,
 *                       agents dealt to lanes by window length.  Same candidate set and per-pair
 *                       arithmetic; the window is summed y-outer instead of x-outer, so results agree
 *                       with the other variants to rounding (1e-5), not bit for bit.  KG_ORDER_ANY +
 *                       relaxed 3x3 query only, else the packed kernel
 *   KG_K4_STAGED        the packed kernel's loops fed from shared memory: a block owns <= 128 consecutive
 *                       agents of one cell column and one thread stages the three column slices they can
 *                       see with three cp.async.bulk (TMA) copies.  Bit-identical to the packed kernel */
enum {
  KG_K4_AUTO = 0, KG_K4_GENERIC = 1, KG_K4_FAST_SCALAR = 2, KG_K4_PACKED_BY_ID = 3, KG_K4_TILED = 4,
  KG_K4_COLTILE = 5, KG_K4_STAGED = 6
};
int kg_field2d_set_kernel_variant(kg_field2d* f, int variant);

/* n x Field2D::set_object_location  field_2d.rs:838-846: append to the WRITE buffer.
 * Out-of-grid coordinates -> KG_E_OOB (reference: Vec index panic), nothing is appended. */
int kg_field2d_set_object_locations(kg_field2d* f, uint64_t n, const uint32_t* id, const float* x,
                                    const float* y, const float* last_dx, const float* last_dy);
int kg_field2d_set_object_locations_dev(kg_field2d* f, uint64_t n, const uint32_t* id,
                                        const float* x, const float* y, const float* last_dx,
                                        const float* last_dy);
/* Field2D::remove_object_location(object, loc)  field_2d.rs:885-898: drop every entry with this
 * id from the write-buffer bag that `loc` discretizes to. */
int kg_field2d_remove_object_location(kg_field2d* f, uint32_t id, float x, float y);
/* Field::lazy_update  field_2d.rs:905-921: swap read/write, make the new read buffer queryable
 * (cell-list rebuild: histogram -> scan -> scatter), write buffer becomes empty. */
int kg_field2d_lazy_update(kg_field2d* f);
/* Field::update  field_2d.rs:903 (a no-op for Field2D) */
int kg_field2d_update(kg_field2d* f);

/* Field2D.nagents (field_2d.rs:277, counted only until the first lazy_update :843-845) */
int kg_field2d_nagents(kg_field2d* f, uint64_t* out);
/* number of agent copies currently in a buffer */
int kg_field2d_num_objects(kg_field2d* f, int which, uint64_t* out);
/* Copy a buffer to the host (any pointer may be NULL).  READ buffer: iter_objects order
 * (field_2d.rs:594-626: x outer, y inner, bag order) with the flat cell index x*dh+y per agent.
 * WRITE buffer: append order.  `cap` = length of the output arrays. */
int kg_field2d_download(kg_field2d* f, int which, uint64_t cap, uint32_t* id, float* x, float* y,
                        float* last_dx, float* last_dy, int32_t* cell, uint64_t* n_out);
/* per-cell occupancy (dw*dh entries) of a buffer */
int kg_field2d_cell_counts(kg_field2d* f, int which, uint64_t cap, uint32_t* counts);
/* Field2D::num_objects_at_location  field_2d.rs:806-811 (read buffer) for nq locations */
int kg_field2d_num_objects_at_locations(kg_field2d* f, uint64_t nq, const float* x, const float* y,
                                        uint32_t* out);
/* Field2D::get_objects / get_objects_unbuffered  field_2d.rs:546-575: ids in the bag of `loc` */
int kg_field2d_get_objects(kg_field2d* f, int which, float x, float y, uint64_t cap, uint32_t* ids,
                           uint64_t* n_out);
/* Field2D::get_empty_bags().len()  field_2d.rs:718-730 */
int kg_field2d_num_empty_bags(kg_field2d* f, uint64_t* out);

/* Batched neighbour query against the READ buffer: for query q the ids are
 * ids[offsets[q] .. offsets[q+1]) in the reference's order (x asc, y asc, bag order).
 * offsets has nq+1 entries and is always complete; when the total exceeds `cap` the call
 * returns KG_E_CAPACITY with *total_out set so the caller can retry. */
int kg_field2d_neighbors(kg_field2d* f, uint64_t nq, const float* qx, const float* qy, float dist,
                         int mode, uint64_t* offsets, uint32_t* ids, uint64_t cap,
                         uint64_t* total_out);
/* The same with the neighbours themselves (pos, last_d) beside their ids: what the reference's
 * queries return is Vec<O> (field_2d.rs:386, :472).  x/y/last_dx/last_dy may all be NULL (ids only).
 * The handle keeps its scratch between calls: no allocation in steady state, one synchronisation. */
int kg_field2d_neighbors_agents(kg_field2d* f, uint64_t nq, const float* qx, const float* qy, float dist,
                                int mode, uint64_t* offsets, uint32_t* ids, float* x, float* y,
                                float* last_dx, float* last_dy, uint64_t cap, uint64_t* total_out);

/* Flockers per-agent step parameters (tests/model/flockers/bird.rs:12-17, :41) */
typedef struct KgBoidsParams {
  float cohesion, avoidance, randomness, consistency, momentum; /* weights  bird.rs:12-16 */
  float jump;                                                   /* bird.rs:17 */
  float radius;                                                 /* bird.rs:41 */
  int32_t exact_query; /* KG_QUERY_EXACT (fixture) or KG_QUERY_RELAX (north-star geometry) */
  uint64_t seed;       /* Philox key */
  uint64_t step;       /* Schedule::step at the time of the call: Philox counter word */
} KgBoidsParams;

/* A model's OWN Agent::step (src/engine/agent.rs:7-16) for every agent of the field: the body of `step` as two
 * CUDA C snippets, compiled at run time for sm_100a (NVRTC, cached per device and source) around the library's
 * window walk, random stream and write log — the generic K4 with the snippets where Bird::step's arithmetic sits.
 * An agent is (id, x, y, a, b): its position and two floats of its own (Bird: last_d).
 *   pair    statements run for every neighbour the query returns, the agent itself included (bird.rs:62-81):
 *           reads sid, sx, sy, sa, sb (self), oid, ox, oy, oa, ob (the neighbour), dx, dy =
 *           toroidal_distance(self, neighbour) per axis (field_2d.rs:988-1002), w, h, c[] (the constants);
 *           updates `float acc[8]` and `int cnt`
 *   finish  statements run once per agent afterwards (bird.rs:83-153): reads acc[], cnt, nvec (neighbours returned),
 *           u0, u1 (this agent's two uniform [0,1) draws of the step), self, c[]; assigns nx, ny (the new position,
 *           inside the grid), na, nb; with may_stop it may set `stopped = true` (Agent::is_stopped, agent.rs:18)
 * Helpers visible to the snippets: toroidal_transform, toroidal_distance, fsqrt, fadd/fsub/fmul/fdiv; ordinary
 * operators are IEEE too (no FMA contraction).  KG_E_INVALID with the compiler's log if a snippet does not compile. */
typedef struct KgCustomStep {
  const char* pair;
  const char* finish;
  float consts[16];
  int32_t nconsts;
  float radius;        /* query distance */
  int32_t exact_query; /* KG_QUERY_EXACT or KG_QUERY_RELAX */
  int32_t may_stop;    /* the finish snippet may set `stopped` */
  uint64_t seed, step; /* Philox key and counter word, as in KgBoidsParams */
} KgCustomStep;
int kg_field2d_step_custom(kg_field2d* f, const KgCustomStep* s);
/* the CUDA source kg_field2d_step_custom compiles for these snippets (no device needed; *need = bytes incl. NUL) */
int kg_jit_agent_source(const char* pair, const char* finish, int may_stop, char* out, uint64_t cap, uint64_t* need);

/* All agents' Agent::step (bird.rs:39-155) for one Schedule::step: reads the READ buffer, pushes
 * every agent's new copy into the WRITE buffer (set_object_location :151-153). */
int kg_field2d_step_boids(kg_field2d* f, const KgBoidsParams* p);
/* nsteps x { step_boids(step = p->step + i); lazy_update } without returning to the host
 * (the body of simulate!'s inner loop, lib.rs:1167-1172, for a Flockers state). */
int kg_field2d_run_boids(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps);
/* Dynamic population: Flockers whose agents die and reproduce, a model of this
 * repo built only from the reference's mechanisms —
 *   death  Agent::is_stopped (src/engine/agent.rs:18): Schedule::step does not reschedule a stopped
 *          agent (src/engine/schedule.rs:401-407); a dying bird does not push itself into the write
 *          buffer, so it is gone from the field after lazy_update
 *   birth  State::after_step(schedule) walks the READ buffer in iter_objects order (bags by index,
 *          each bag in stored order) and, for every parent that draws a birth, pushes a child
 *          (id = next_id++, the parent's position at the start of the step, last_d = 0) into the
 *          write buffer and schedules it (schedule_repeating, schedule.rs:295-303)
 * Draws: Philox(seed; id, step, domain 3): v[0] < death_prob -> stopped, v[1] < birth_prob -> child;
 * crowd_limit > 0 also stops an agent whose query returned >= crowd_limit neighbours other than
 * itself (bird.rs:80's `count`).  With KG_ORDER_CANONICAL the children's ids equal the oracle's.
 * Id 0xFFFFFFFF is reserved (it marks a stopped agent's log entry). */
typedef struct KgLifeRule {
  float death_prob, birth_prob;
  uint32_t crowd_limit;
  uint32_t reserved;
} KgLifeRule;
/* id the next child gets (default: init_flockers' n, or 1 + the largest id uploaded from host arrays) */
int kg_field2d_set_next_id(kg_field2d* f, uint32_t next_id);
/* kg_field2d_step_boids + is_stopped for every agent, then the births of State::after_step; the
 * following kg_field2d_lazy_update compacts the dead away.  *n_stopped / *n_born: this step's. */
int kg_field2d_step_boids_life(kg_field2d* f, const KgBoidsParams* p, const KgLifeRule* life,
                               uint64_t* n_stopped, uint64_t* n_born);

/* State::init of the Flockers fixture (state.rs:41-56) on the device: agent i gets
 * pos = (w*r1, h*r2), last_d = 0 with (r1,r2) = Philox(seed; i, 0, 0, domain 0), pushed into
 * the WRITE buffer. */
int kg_field2d_init_flockers(kg_field2d* f, uint64_t n, uint64_t seed);

/* One e2e step with HOST buffers: upload n agents (n x set_object_location), lazy_update, every
 * agent's step; the stepped agents come back in the cell order of their INPUT positions (ids travel
 * with them), slab by slab while the next slab still computes.  The handle is left with the stepped
 * population in its write buffer (a following kg_field2d_lazy_update makes it the read buffer); the
 * next call starts from empty buffers again.  in/out arrays should be page-locked (kg_host_alloc)
 * for full PCIe rate; out-of-grid input -> KG_E_OOB. */
int kg_field2d_step_boids_host(kg_field2d* f, const KgBoidsParams* p, uint64_t n,
                               const uint32_t* id_in, const float* x_in, const float* y_in,
                               const float* dx_in, const float* dy_in, uint32_t* id_out,
                               float* x_out, float* y_out, float* dx_out, float* dy_out);
/* The same step for a host that keeps its agents in arrays and finds them again BY POSITION: the result of
 * input agent i comes back at index i (no ids travel back), and id_in may be NULL, meaning "agent i has id i"
 * (Flockers: Bird::new(i, ..), state.rs:47) — 16 bytes per agent each way instead of 20.  Same arithmetic,
 * same random stream (keyed by the id) as kg_field2d_step_boids_host; x_out etc. may alias the inputs. */
int kg_field2d_step_boids_host_ordered(kg_field2d* f, const KgBoidsParams* p, uint64_t n,
                                       const uint32_t* id_in, const float* x_in, const float* y_in,
                                       const float* dx_in, const float* dy_in, float* x_out, float* y_out,
                                       float* dx_out, float* dy_out);

/* Device-side reductions over the READ buffer: what a model's output columns
 * (`explore`'s FrameRow fields, src/explore/model_exploration.rs:160-190; write_csv src/lib.rs:1781-1800)
 * or a plot! series are computed from, without downloading the population.  out[KG_RED_COUNT] =
 * sum x, sum y, sum last_d.x, sum last_d.y, sum |last_d|, sum x^2, sum y^2, 0 — f64, deterministic
 * (fixed chunking and order). */
enum { KG_RED_SUM_X = 0, KG_RED_SUM_Y = 1, KG_RED_SUM_LDX = 2, KG_RED_SUM_LDY = 3, KG_RED_SUM_SPEED = 4,
       KG_RED_SUM_XX = 5, KG_RED_SUM_YY = 6, KG_RED_COUNT = 8 };
int kg_field2d_reduce(kg_field2d* f, double* out /*[KG_RED_COUNT]*/);
/* A plot! series (src/lib.rs:1202-1240: plot!(name, series, x, y) called from after_step) recorded on the
 * device: nsteps steps as kg_field2d_run_boids, and after every `every`-th one the reductions above are
 * written to a device row; the nsteps / every rows come back with ONE copy at the end (out[rows][KG_RED_COUNT],
 * row r = state after step (r + 1) * every).  No host round trip inside the loop. */
int kg_field2d_run_boids_series(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps, uint64_t every,
                                double* out, uint64_t out_rows);

/* kernel-time instrumentation: accumulated device milliseconds and launch counts per kernel
 * family since the last reset (CUDA events on the handle's stream; enable=0 turns it off). */
enum { KG_K_STEP = 0, KG_K_HIST = 1, KG_K_SCAN = 2, KG_K_SCATTER = 3, KG_K_SORTCELL = 4,
       KG_K_QUERY = 5, KG_K_MISC = 6, KG_K_STENCIL = 7, KG_K_COUNT = 8 };
/* Measurement helpers (bench.py).  l2_flush overwrites a private scratch buffer of `bytes` on the
 * handle's stream so that the next kernel starts with a cold L2.  run_boids_timed = run_boids with
 * every {step_boids; lazy_update} bracketed by CUDA events on the stream and an untimed l2_flush
 * between steps (flush_bytes = 0: none); *ms_sum = sum of the per-step device times. */
int kg_field2d_l2_flush(kg_field2d* f, uint64_t bytes);
int kg_field2d_run_boids_timed(kg_field2d* f, const KgBoidsParams* p, uint64_t nsteps,
                               uint64_t flush_bytes, double* ms_sum);
/* device-side stopwatch on the handle's stream: start records an event, stop records another,
 * synchronises and returns the milliseconds between them (CUDA events, not host clocks). */
int kg_field2d_timer_start(kg_field2d* f);
int kg_field2d_timer_stop(kg_field2d* f, double* ms);
int kg_field2d_profile(kg_field2d* f, int enable);
int kg_field2d_profile_read(kg_field2d* f, double* ms /*[KG_K_COUNT]*/,
                            uint64_t* launches /*[KG_K_COUNT]*/, int reset);
/* total kernel launches issued by this library in this process */
uint64_t kg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU Field2D: the world is cut into x-strips of whole cell columns, one kg_strip per GPU
 * (one per process under torchrun, or several in one process).  Per step every strip runs the
 * fused boids kernel on its own agents, hands agents that crossed a strip boundary to the ring
 * neighbour (toroidal_transform wraps positions, bird.rs:146) and refreshes the halo columns of
 * its line neighbours (the toroidal query window is clamped, field_2d.rs:495-500).  Both travel
 * in ONE exchange per step: peer stores over NVLink into the neighbour's inbox.  The reference precedent is
 * src/engine/fields/kdtree_mpi.rs:705-790 (halo regions + per-step p2p exchange over MPI).
 * Toroidal fields, relaxed query only.
 * ------------------------------------------------------------------------------------------ */
typedef struct kg_strip kg_strip;
#define KG_IPC_HANDLE_BYTES 64

/* `radius` fixes the halo width floor(radius/disc); `capacity` = most agents the strip may own,
 * `halo_capacity` = most agents in one neighbour's boundary columns, `migrate_capacity` = most
 * agents leaving in one direction in one step. */
int kg_strip_create(float w, float h, float discretization, int toroidal, float radius, int rank,
                    int nranks, uint64_t capacity, uint64_t halo_capacity, uint64_t migrate_capacity,
                    int device, kg_strip** out);
int kg_strip_destroy(kg_strip* s);
/* owned global cell columns [own_x0, own_x1), halo columns per side, cells per column */
int kg_strip_columns(kg_strip* s, int32_t* own_x0, int32_t* own_x1, int32_t* halo_l, int32_t* halo_r,
                     int32_t* dh);
int kg_strip_set_order(kg_strip* s, int order);
/* wiring, multi-process: export my inbox as a CUDA IPC handle, open the ring neighbours' */
int kg_strip_ipc_export(kg_strip* s, void* handle /*[KG_IPC_HANDLE_BYTES]*/);
int kg_strip_connect_ipc(kg_strip* s, const void* left_handle, const void* right_handle);
/* wiring, single process: neighbours are other handles of this process (peer access is enabled) */
int kg_strip_connect_local(kg_strip* s, kg_strip* left, kg_strip* right);
/* State::init of the Flockers fixture for ids 0..n_global-1: the strip keeps what it owns */
int kg_strip_init_flockers(kg_strip* s, uint64_t n_global, uint64_t seed);
/* n x set_object_location for agents this strip owns (anything else: KG_E_OOB) */
int kg_strip_upload(kg_strip* s, uint64_t n, const uint32_t* id, const float* x, const float* y,
                    const float* last_dx, const float* last_dy);
/* drop every agent of the strip (the e2e path re-uploads the population each step) */
int kg_strip_clear(kg_strip* s);
/* first lazy_update incl. halo exchange; every rank must call it before stepping */
int kg_strip_prepare(kg_strip* s);
/* one Schedule::step for the strip: K4 + migration + lazy_update + halo refresh (asynchronous).
 * With several strips in ONE process issue step s for every strip before step s+1 for any. */
int kg_strip_step_boids(kg_strip* s, const KgBoidsParams* p);
/* multi-process only (each process owns one strip) */
int kg_strip_run_boids(kg_strip* s, const KgBoidsParams* p, uint64_t nsteps);
int kg_strip_run_boids_timed(kg_strip* s, const KgBoidsParams* p, uint64_t nsteps,
                             uint64_t flush_bytes, double* ms_sum);
int kg_strip_sync(kg_strip* s);
int kg_strip_stats(kg_strip* s, uint64_t* n_owned, uint64_t* migrants_in, uint64_t* migrants_out,
                   uint64_t* halo_left, uint64_t* halo_right, uint64_t* launches);
/* owned agents in iter_objects order */
int kg_strip_download(kg_strip* s, uint64_t cap, uint32_t* id, float* x, float* y, float* last_dx,
                      float* last_dy, uint64_t* n_out);
int kg_strip_timer_start(kg_strip* s);
int kg_strip_timer_stop(kg_strip* s, double* ms);

/* ------------------------------------------------------------------------------------------
 * DenseNumberGrid2D<T>  (src/engine/fields/dense_number_grid_2d.rs:90-561, default variant)
 * Flat index x*height + y.  `Option<T>` is stored as T with one reserved value `none` = None.
 * ------------------------------------------------------------------------------------------ */
typedef struct kg_grid kg_grid;
enum { KG_GRID_READ = 0, KG_GRID_WRITE = 1, KG_GRID_READWRITE = 2 }; /* grid_option.rs:3-10 */
enum { KG_APPLY_CONST = 0, KG_APPLY_ADD = 1 };                       /* closure family */
enum {
  KG_RULE_FOREST_FIRE = 0 /* Moore-8: GREEN(1)->BURNING(2) if any neighbour burns;
                             BURNING->BURNED(3); BURNED stays; None stays None */
};

/* DenseNumberGrid2D::new(width,height)  :112-126.  elem_size in {1,2,4}. */
int kg_grid_create(int32_t width, int32_t height, int elem_size, uint32_t none, int device,
                   kg_grid** out);
int kg_grid_destroy(kg_grid* g);
int kg_grid_sync(kg_grid* g);
/* n x set_value_location(value, loc) :492-495 / remove_value_location(loc) :526-530.
 * A value equal to `none` (the reserved encoding of Option::None) -> KG_E_INVALID. */
int kg_grid_set_values(kg_grid* g, uint64_t n, const int32_t* x, const int32_t* y,
                       const void* values);
int kg_grid_remove_values(kg_grid* g, uint64_t n, const int32_t* x, const int32_t* y);
/* n x get_value :350-354 (which = KG_BUF_READ) / get_value_unbuffered :376-380; None -> `none` */
int kg_grid_get_values(kg_grid* g, int which, uint64_t n, const int32_t* x, const int32_t* y,
                       void* out);
/* whole buffer, x-major, None = `none` */
int kg_grid_upload(kg_grid* g, int which, const void* cells);
int kg_grid_download(kg_grid* g, int which, void* cells);
/* apply_to_all_values(closure, option) :155-195 for closure in {|_| c, |v| v + c}.  v + c wraps
 * like a Rust release build; a result (or a constant c) equal to `none` cannot be stored as Some(..)
 * and is reported as KG_E_INVALID (by the call for a constant, at the next sync for v + c). */
int kg_grid_apply(kg_grid* g, int op, uint32_t operand, int option);
/* apply_to_all_values(closure, option) :155-195 for an ARBITRARY closure: `expr` is the closure's body as
 * a CUDA C expression of the cell type in the variable `v` (the cell's value; also visible: the cell's
 * int x, y), e.g. "v - 1" for the doc example's |x| x - 1 (:152), "v == 2 ? 0 : v * 3", "(v + x) % 7".
 * It is compiled at run time for sm_100a (NVRTC, cached per device and source) into the same kernel as
 * kg_grid_apply: GridOption semantics, None cells and the reserved-value check are identical.
 * KG_E_INVALID with the compiler's log if the expression does not compile; KG_E_CUDA without NVRTC. */
int kg_grid_apply_expr(kg_grid* g, const char* expr, int option);
/* get_location / get_location_unbuffered :204-229: first Some(value) in x-outer/y-inner order;
 * *found = 0 when absent (always for value == `none`: empty cells never match, :222) */
int kg_grid_get_location(kg_grid* g, int which, uint32_t value, int32_t* x, int32_t* y, int* found);
/* get_empty_bags().len() :236-247 */
int kg_grid_num_empty(kg_grid* g, uint64_t* out);
/* Field::lazy_update :537-545 (swap; write := None) and Field::update :553-561 */
int kg_grid_lazy_update(kg_grid* g);
int kg_grid_update(kg_grid* g);
/* one model step through the field API: every live cell of the READ buffer writes its next state
 * into the WRITE buffer (get_value + set_value_location), no swap */
int kg_grid_step_stencil(kg_grid* g, int rule);
/* the same step for a rule given as a CUDA C expression (compiled at run time like kg_grid_apply_expr):
 * `v` = the cell's value, `at(dx, dy)` = the value of the cell at (x + dx, y + dy) of the READ buffer, or
 * NONE outside the grid and for an empty cell, `x`, `y`, `NONE`.  Every live cell writes the expression's
 * value into the WRITE buffer, a None cell stays None; no swap.  Forest Fire reads
 *   "v == 1 ? (at(-1,-1)==2 || at(-1,0)==2 || at(-1,1)==2 || at(0,-1)==2 || at(0,1)==2 || at(1,-1)==2 ||
 *              at(1,0)==2 || at(1,1)==2 ? 2 : 1) : (v == 2 ? 3 : v)"
 * (the generic kernel: one thread per cell, no register window — the shipped rule stays the fast path). */
int kg_grid_step_expr(kg_grid* g, const char* expr);
/* nsteps x { step_stencil; lazy_update } on the device.  For u8 grids (none = 0xFF, height a multiple
 * of 16) runs of 8 / 4 / 2 Forest-Fire steps are computed by ONE pass over the grid (step t in, step
 * t+T out, the levels between in registers: stencil_device.cuh); the buffers afterwards hold exactly what
 * nsteps single steps leave: read buffer = step t+nsteps, write buffer all None. */
int kg_grid_run_stencil(kg_grid* g, int rule, uint64_t nsteps);
/* Forest-Fire initial state on the device: tree with probability `density`
 * (Philox(seed; cell, domain 2)), trees in column x == 0 burning; then lazy_update. */
int kg_grid_init_forest_fire(kg_grid* g, float density, uint64_t seed);
int kg_grid_run_stencil_timed(kg_grid* g, int rule, uint64_t nsteps, double* ms_sum);
int kg_grid_timer_start(kg_grid* g);
int kg_grid_timer_stop(kg_grid* g, double* ms);
int kg_grid_profile(kg_grid* g, int enable);
int kg_grid_profile_read(kg_grid* g, double* ms, uint64_t* launches, int reset);

/* ------------------------------------------------------------------------------------------
 * Batched independent replicas for parameter sweeps: the device-side form of explore_parallel!
 * (src/explore/model_exploration.rs:354-423 — one State + Schedule per rayon task, the body of
 * simulate_explore! :160-190 inside each).  `replicas` Flockers worlds of the same geometry and
 * population advance together, one launch per phase for the whole batch, no communication between
 * replicas.  Every replica has its own KgBoidsParams (the swept inputs: weights, radius, seed);
 * defaults are the fixture's constants (bird.rs:12-17), relaxed query, seed 42 + replica index.
 * Array arguments of upload/download are replica-major: element r * agents_per_replica + k.
 * Sharding a sweep over several GPUs is one batch per device (replica i -> device i % G, as
 * explore/mpi/model_exploration.rs:206 assigns configurations to ranks); there is no exchange. */
typedef struct kg_batch kg_batch;
int kg_batch_create(float w, float h, float discretization, int toroidal, uint32_t replicas,
                    uint32_t agents_per_replica, int device, kg_batch** out);
int kg_batch_destroy(kg_batch* b);
int kg_batch_dims(kg_batch* b, uint32_t* replicas, uint32_t* agents_per_replica, int32_t* dw,
                  int32_t* dh);
int kg_batch_set_order(kg_batch* b, int order); /* KG_ORDER_* as for kg_field2d */
/* parameters of replicas [first, first + n); `step` fields are ignored */
int kg_batch_set_params(kg_batch* b, uint32_t first, uint32_t n, const KgBoidsParams* p);
/* State::init (state.rs:41-56) of every replica with its own seed, into the WRITE buffers */
int kg_batch_init_flockers(kg_batch* b);
/* n x set_object_location for every replica at once (replaces the whole population) */
int kg_batch_upload(kg_batch* b, const uint32_t* id, const float* x, const float* y,
                    const float* last_dx, const float* last_dy);
/* Field::lazy_update of every replica's field */
int kg_batch_lazy_update(kg_batch* b);
/* every replica's Schedule::step agent phase (all Bird::step) with Philox counter word `step` */
int kg_batch_step_boids(kg_batch* b, uint64_t step);
/* nsteps x { step_boids(first_step + i); lazy_update }: simulate_explore!'s loop for all replicas */
int kg_batch_run_boids(kg_batch* b, uint64_t first_step, uint64_t nsteps);
int kg_batch_run_boids_timed(kg_batch* b, uint64_t first_step, uint64_t nsteps, uint64_t flush_bytes,
                             double* ms_sum);
/* read buffers, replica-major, each replica in its iter_objects order; cell may be NULL */
int kg_batch_download(kg_batch* b, uint32_t* id, float* x, float* y, float* last_dx, float* last_dy,
                      int32_t* cell);
/* kg_field2d_reduce for every replica: out[replicas][KG_RED_COUNT] — an explore sweep's output rows
 * cost replicas x 64 bytes of download instead of the whole population */
int kg_batch_reduce(kg_batch* b, double* out);
int kg_batch_sync(kg_batch* b);
int kg_batch_timer_start(kg_batch* b);
int kg_batch_timer_stop(kg_batch* b, double* ms);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU DenseNumberGrid2D<u8> for stencil models (Forest Fire): strips of whole x rows, one
 * kg_gridstrip per GPU (rank r owns rows [r*W/G, (r+1)*W/G)).  Each step is ONE kernel per GPU:
 * the blocks computing a strip's first / last row also store it into the line neighbour's inbox
 * (peer stores over NVLink) and publish an epoch flag; the blocks that need the neighbour's row
 * wait on that flag.  No collective, no host round trip.  Protocol: create on every rank ->
 * exchange inbox handles (ipc_export / connect_ipc between processes, connect_local inside one)
 * -> init_forest_fire or upload -> [barrier] -> prepare on every rank -> [barrier] -> run.
 * Cells are u8 with Option::None = 0xFF; height must be a multiple of 16. */
typedef struct kg_gridstrip kg_gridstrip;
int kg_gridstrip_create(int32_t width, int32_t height, int rank, int nranks, int device,
                        kg_gridstrip** out);
int kg_gridstrip_destroy(kg_gridstrip* s);
int kg_gridstrip_rows(kg_gridstrip* s, int32_t* x0, int32_t* x1);
/* Host logic of kg_gridstrip_run_stencil, callable without a device: the passes nsteps are cut into
 * (steps_of_pass[k] in {8, 4, 2, 1}, never more than the rows of the smallest strip, the same on every rank)
 * and the rows per tile of each pass for the smallest strip (its last tile never holds fewer than the eight
 * rows a pass hands to a neighbour).  Fills up to `cap` entries, *npasses = the number of passes. */
int kg_gridstrip_pass_plan(int32_t width, int32_t height, int nranks, uint64_t nsteps, int32_t* steps_of_pass,
                           int32_t* rows_per_tile, uint64_t cap, uint64_t* npasses);
int kg_gridstrip_ipc_export(kg_gridstrip* s, void* handle /*[KG_IPC_HANDLE_BYTES]*/);
/* handles of the line neighbours; NULL where there is none (rank 0 / rank G-1) */
int kg_gridstrip_connect_ipc(kg_gridstrip* s, const void* left_handle, const void* right_handle);
int kg_gridstrip_connect_local(kg_gridstrip* s, kg_gridstrip* left, kg_gridstrip* right);
/* same cells as kg_grid_init_forest_fire on the whole grid (Philox counter = global cell index),
 * already readable (no lazy_update needed) */
int kg_gridstrip_init_forest_fire(kg_gridstrip* s, float density, uint64_t seed);
/* the strip's own rows, (x1-x0)*height bytes, to / from the READ buffer */
int kg_gridstrip_upload(kg_gridstrip* s, const uint8_t* own_rows);
int kg_gridstrip_download(kg_gridstrip* s, uint8_t* own_rows);
/* hands the current boundary rows to the neighbours; required after init / upload */
int kg_gridstrip_prepare(kg_gridstrip* s);
/* nsteps x { every live cell's get_value + set_value_location; lazy_update }.  Passes of 8 / 4 / 2 / 1
 * steps as for kg_grid_run_stencil (T <= the rows of the smallest strip); every pass hands its first and
 * last eight rows to the line neighbours, one flag round per pass.  Every rank must call it with the same
 * nsteps. */
int kg_gridstrip_run_stencil(kg_gridstrip* s, int rule, uint64_t nsteps);
/* same, bracketed by two CUDA events on the strip's stream: *ms_total = device time of the run */
int kg_gridstrip_run_stencil_timed(kg_gridstrip* s, int rule, uint64_t nsteps, double* ms_total);
int kg_gridstrip_sync(kg_gridstrip* s);

/* ------------------------------------------------------------------------------------------
 * 2-D block decomposition of Field2D (precedent: the kd-tree blocks of
 * src/engine/fields/kdtree_mpi.rs:211-238): the cell grid cut into nbx x nby rectangles, one kg_block per
 * rectangle (any device), each with a ring of halo cells.  For worlds where strips of whole columns are
 * too thin; one process drives every block (the step kernels store migrants and ghosts straight into the
 * neighbours' inboxes — peer access between the devices is required —, the counts go through the host).  With
 * KG_ORDER_CANONICAL the blocks reproduce one GPU bit for bit.  Toroidal fields of the packed K4's
 * geometry class, relaxed and exact query.
 * ------------------------------------------------------------------------------------------ */
typedef struct kg_block kg_block;
/* capacity = most agents (owned + ghosts) the block holds; xcap = most entries it sends one neighbour per step */
int kg_block_create(float w, float h, float disc, int toroidal, float radius, int bx, int by, int nbx, int nby,
                    uint64_t capacity, uint64_t xcap, int device, kg_block** out);
int kg_block_destroy(kg_block* b);
int kg_block_cells(kg_block* b, int32_t* own /*[4]: x0, x1, y0, y1*/, int32_t* local /*[4], halo ring included*/);
int kg_block_set_order(kg_block* b, int order);
/* n x set_object_location (:838-846): hand every block the same agents, it keeps those inside its window */
int kg_block_upload(kg_block* b, uint64_t n, const uint32_t* id, const float* x, const float* y, const float* dx,
                    const float* dy);
/* Field::lazy_update after uploads */
int kg_block_lazy_update(kg_block* b);
/* one step of the whole world: blocks[bx * nby + by], all nbx * nby of them */
int kg_blocks_step(kg_block** blocks, int nblocks, const KgBoidsParams* p);
int kg_blocks_run(kg_block** blocks, int nblocks, const KgBoidsParams* p, uint64_t nsteps);
/* the agents the block owns (its ghosts left out) */
int kg_block_download(kg_block* b, uint64_t cap, uint32_t* id, float* x, float* y, float* dx, float* dy,
                      uint64_t* n_out);
int kg_block_counts(kg_block* b, uint64_t* n_local /*owned + ghosts*/, uint64_t* n_cells);
/* the partition rule alone (host only): part b of `parts` over maxc scanned columns owns [out[0], out[1]) (the last
 * part also the padding column) and keeps [out[2], out[3]) with its halo ring of dd cells; *owner = part owning c */
int kg_block_partition(int b, int parts, int maxc, int dd, int c, int32_t* out /*[4]*/, int32_t* owner);

/* ------------------------------------------------------------------------------------------
 * DenseGrid2D<O>  (src/engine/fields/dense_object_grid_2d.rs:175-779, default variant) — SURVEY §8f-2
 * Objects are (id, tag) pairs that compare by id (the fixture's Bird, bird.rs:168-172; tag ~ Bird.flag).
 * Bags are addressed by the flat index x*height + y and bounds-checked against the Vec only, like
 * the reference (an out-of-range index -> KG_E_OOB where it panics).  which = KG_BUF_READ / KG_BUF_WRITE
 * selects the buffered / *_unbuffered method.
 * ------------------------------------------------------------------------------------------ */
typedef struct kg_objgrid kg_objgrid;
/* DenseGrid2D::new(width, height) :201-214; `capacity` = most objects one buffer (plus pending writes) holds */
int kg_objgrid_create(int32_t width, int32_t height, uint64_t capacity, int device, kg_objgrid** out);
/* SparseGrid2D<O>::new(width, height)  (src/engine/fields/sparse_object_grid_2d.rs:203-234) behind the same
 * verbs.  Two HashMap<Int2D, Vec<O>> in the reference: ANY (x, y) is a key (no bounds, no KG_E_OOB; only
 * (i32::MAX, i32::MAX) is reserved), a key whose bag is empty does not exist.  On the device a bag is a slot
 * of an open-addressing key table shared by both buffers (rebuilt from the live keys when half full).
 * Differences from the dense grid, all the reference's own:
 *   set_object_location  :648-659  pushes without replacing an equal object
 *   remove_object_location :690-699, lazy_update :705-708, update :711-718 (read = copy of write; offered)
 *   get_objects :421-423 None for a key without objects; get_location :346-356 any bag holding the object
 *   iter_objects :561-574 unspecified key order; bag_sizes = the nominal width x height area (get_empty_bags
 *   :482-499); apply_to_all_values :278-320 — the closure sees the bag's key, a None result is the reference's
 *   panic (KG_E_INVALID), READWRITE gives a key absent from the write map ONE object (the last read one). */
int kg_objgrid_create_sparse(int32_t width, int32_t height, uint64_t capacity, int device, kg_objgrid** out);
int kg_objgrid_destroy(kg_objgrid* g);
int kg_objgrid_dims(kg_objgrid* g, int32_t* width, int32_t* height, uint64_t* nbags);
/* n x set_object_location(object, loc) :688-697, in array order: an equal object already in that
 * write bag is replaced (removed, the new one pushed last) */
int kg_objgrid_set_object_locations(kg_objgrid* g, uint64_t n, const uint32_t* id, const uint32_t* tag,
                                    const int32_t* x, const int32_t* y);
/* n x remove_object_location(object, loc) :729-736 */
int kg_objgrid_remove_object_locations(kg_objgrid* g, uint64_t n, const uint32_t* id, const int32_t* x,
                                       const int32_t* y);
/* Field::lazy_update :743-750: swap, every write bag cleared */
int kg_objgrid_lazy_update(kg_objgrid* g);
/* Field::update :753-763 is NOT offered (KG_E_INVALID): the reference's Vec::insert doubles the read
 * Vec and breaks its own apply_to_all_values; the oracle pins that quirk, the device refuses it. */
int kg_objgrid_update(kg_objgrid* g);
int kg_objgrid_num_objects(kg_objgrid* g, int which, uint64_t* out);
/* get_objects :507-520 / get_objects_unbuffered :547-561, bag order; *n_out == 0 is Option::None */
int kg_objgrid_get_objects(kg_objgrid* g, int which, int32_t x, int32_t y, uint64_t cap, uint32_t* id,
                           uint32_t* tag, uint64_t* n_out);
/* get_location :429-441 / get_location_unbuffered :471-482: first bag in x-outer/y-inner order */
int kg_objgrid_get_location(kg_objgrid* g, int which, uint32_t id, int32_t* x, int32_t* y, int* found);
/* iter_objects :589-608 / iter_objects_unbuffered :634-654: x outer, y inner, bag order */
int kg_objgrid_iter_objects(kg_objgrid* g, int which, uint64_t cap, int32_t* x, int32_t* y, uint32_t* id,
                            uint32_t* tag, uint64_t* n_out);
/* objects per bag in flat order (get_empty_bags :358-370 = the zeros) */
int kg_objgrid_bag_sizes(kg_objgrid* g, int which, uint64_t cap, uint32_t* sizes);
/* apply_to_all_values(closure, option) :258-328 for the closure family
 *   KG_OBJ_SET_TAG           |_, o| Some(o with tag = arg)
 *   KG_OBJ_REMOVE            |_, _| None
 *   KG_OBJ_REMOVE_IF_TAG     |_, o| if o.tag == arg { None } else { Some(o) }
 *   KG_OBJ_TAG_WITH_BAG_ID   |bag, o| Some(o with tag = bag.x * 65536 + bag.y)   (bag = calculate_indexes_bag,
 *                            :768-779 — y-major, which is the reference's own quirk)
 * with GridOption READ / WRITE / READWRITE semantics as in the reference; *calls = closure calls. */
enum { KG_OBJ_SET_TAG = 0, KG_OBJ_REMOVE = 1, KG_OBJ_REMOVE_IF_TAG = 2, KG_OBJ_TAG_WITH_BAG_ID = 3 };
int kg_objgrid_apply(kg_objgrid* g, int op, uint32_t arg, int option, uint64_t* calls);

#ifdef __cplusplus
}
#endif
#endif /* KRABGPU_H */
