// ORACLE — TEST INFRASTRUCTURE ONLY (see krabmaga_oracle.hpp header).
// Flat C entry points over the restated reference so that tests/ (ctypes), smoke() and
// bench.py's cpu_baseline leg can drive it.  A restated Rust panic surfaces as return
// code -1 with the text available from okg_last_error().
#include <chrono>
#include <cstring>
#include <thread>

#include "krabmaga_oracle.hpp"
#include "object_grid.hpp"

using namespace oracle;

static thread_local std::string g_err;
#define OKG_TRY try {
#define OKG_CATCH                   \
  }                                 \
  catch (const std::exception& e) { \
    g_err = e.what();               \
    return -1;                      \
  }                                 \
  return 0;

extern "C" {

const char* okg_last_error() { return g_err.c_str(); }

void okg_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  std::memcpy(out, r.v, sizeof(r.v));
}
float okg_u01_f32(uint32_t u) { return u01_f32(u); }
float okg_toroidal_transform(float v, float dim) { return toroidal_transform(v, dim); }
float okg_toroidal_distance(float a, float b, float dim) { return toroidal_distance(a, b, dim); }
int okg_t_transform(int n, int size) { return t_transform(n, size); }

// ------------------------------------------------------------------ Field2D<Bird>
typedef Field2D<Bird> F2;
void* okg_field2d_new(float w, float h, float d, int tor) { return new F2(w, h, d, tor != 0); }
void okg_field2d_free(void* f) { delete (F2*)f; }
void okg_field2d_dims(void* f, int* dw, int* dh, uint64_t* nbags_read, uint64_t* nbags_write) {
  F2* p = (F2*)f;
  *dw = p->dw;
  *dh = p->dh;
  *nbags_read = p->bags[p->read].size();
  *nbags_write = p->bags[p->write].size();
}
uint64_t okg_field2d_nagents(void* f) { return ((F2*)f)->nagents; }
int okg_field2d_discretize(void* f, float x, float y, int* cx, int* cy) {
  Int2D c = ((F2*)f)->discretize(Real2D{x, y});
  *cx = c.x;
  *cy = c.y;
  return 0;
}
int okg_field2d_set_object_location(void* f, uint32_t id, float x, float y, float ldx, float ldy) {
  OKG_TRY((F2*)f)->set_object_location(Bird{id, Real2D{x, y}, Real2D{ldx, ldy}, false},
                                       Real2D{x, y});
  OKG_CATCH
}
int okg_field2d_set_object_locations(void* f, uint64_t n, const uint32_t* id, const float* x,
                                     const float* y, const float* ldx, const float* ldy) {
  OKG_TRY for (uint64_t i = 0; i < n; ++i)((F2*)f)
      ->set_object_location(Bird{id[i], Real2D{x[i], y[i]}, Real2D{ldx[i], ldy[i]}, false},
                            Real2D{x[i], y[i]});
  OKG_CATCH
}
int okg_field2d_remove_object_location(void* f, uint32_t id, float x, float y) {
  OKG_TRY((F2*)f)->remove_object_location(Bird{id, Real2D{x, y}, Real2D{0, 0}, false},
                                          Real2D{x, y});
  OKG_CATCH
}
void okg_field2d_lazy_update(void* f) { ((F2*)f)->lazy_update(); }
void okg_field2d_update(void* f) { ((F2*)f)->update(); }

static int64_t copy_ids(const std::vector<Bird>& v, uint32_t* out, uint64_t cap) {
  for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i].id;
  return (int64_t)v.size();
}
// mode 0 = relax, 1 = exact.  Returns the neighbour count (ids in reference order), -1 on panic.
int64_t okg_field2d_neighbors(void* f, float x, float y, float dist, int mode, uint32_t* out,
                              uint64_t cap) {
  try {
    F2* p = (F2*)f;
    auto v = mode ? p->get_neighbors_within_distance(Real2D{x, y}, dist)
                  : p->get_neighbors_within_relax_distance(Real2D{x, y}, dist);
    return copy_ids(v, out, cap);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// batched: CSR offsets[nq+1]; ids truncated at cap but offsets always exact
int okg_field2d_neighbors_batch(void* f, uint64_t nq, const float* qx, const float* qy, float dist,
                                int mode, uint64_t* offsets, uint32_t* ids, uint64_t cap) {
  OKG_TRY F2* p = (F2*)f;
  uint64_t total = 0;
  offsets[0] = 0;
  for (uint64_t q = 0; q < nq; ++q) {
    auto v = mode ? p->get_neighbors_within_distance(Real2D{qx[q], qy[q]}, dist)
                  : p->get_neighbors_within_relax_distance(Real2D{qx[q], qy[q]}, dist);
    for (const Bird& b : v) {
      if (total < cap) ids[total] = b.id;
      ++total;
    }
    offsets[q + 1] = total;
  }
  OKG_CATCH
}
int64_t okg_field2d_get_objects(void* f, float x, float y, int unbuffered, uint32_t* out,
                                uint64_t cap) {
  try {
    F2* p = (F2*)f;
    auto v = unbuffered ? p->get_objects_unbuffered(Real2D{x, y}) : p->get_objects(Real2D{x, y});
    return copy_ids(v, out, cap);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int64_t okg_field2d_num_objects_at_location(void* f, float x, float y) {
  try {
    return (int64_t)((F2*)f)->num_objects_at_location(Real2D{x, y});
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int64_t okg_field2d_get_empty_bags(void* f, float* ox, float* oy, uint64_t cap) {
  try {
    auto v = ((F2*)f)->get_empty_bags();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
      ox[i] = v[i].x;
      oy[i] = v[i].y;
    }
    return (int64_t)v.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// iter_objects order (field_2d.rs:594-660: x outer, y inner, bag order): per object its
// cell origin (not_discretize), id, pos, last_d and flat cell index.
int64_t okg_field2d_iter_objects(void* f, int unbuffered, uint64_t cap, uint32_t* id, float* x,
                                 float* y, float* ldx, float* ldy, int32_t* cell, float* ox,
                                 float* oy) {
  try {
    F2* p = (F2*)f;
    const auto& bg = p->bags[unbuffered ? p->write : p->read];
    uint64_t n = 0;
    for (int32_t i = 0; i < p->dw; ++i)
      for (int32_t j = 0; j < p->dh; ++j) {
        size_t index = (size_t)(i * p->dh + j);
        if (index >= bg.size()) rust_panic("iter_objects index");
        Real2D rp = p->not_discretize(Int2D{i, j});
        for (const Bird& b : bg[index]) {
          if (n < cap) {
            if (id) id[n] = b.id;
            if (x) x[n] = b.pos.x;
            if (y) y[n] = b.pos.y;
            if (ldx) ldx[n] = b.last_d.x;
            if (ldy) ldy[n] = b.last_d.y;
            if (ox) ox[n] = rp.x;
            if (oy) oy[n] = rp.y;
            if (cell) cell[n] = (int32_t)index;
          }
          ++n;
        }
      }
    return (int64_t)n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// per-cell occupancy of the chosen buffer (length min(cap, nbags))
int64_t okg_field2d_cell_counts(void* f, int unbuffered, uint32_t* counts, uint64_t cap) {
  F2* p = (F2*)f;
  const auto& b = p->bags[unbuffered ? p->write : p->read];
  for (size_t i = 0; i < b.size() && i < cap; ++i) counts[i] = (uint32_t)b[i].size();
  return (int64_t)b.size();
}

// ------------------------------------------------------------------ Flockers model + Schedule
struct FlockSim {
  Flocker state;
  Schedule schedule;
  std::vector<Bird> preset;
  FlockSim(float w, float h, uint32_t n, float disc, bool tor, BoidsParams p)
      : state(w, h, n, disc, tor, p) {}
};

struct OkgBoidsParams {  // mirrors include/krabgpu.h KgBoidsParams field for field
  float cohesion, avoidance, randomness, consistency, momentum, jump, radius;
  int32_t exact_query;
  uint64_t seed;
  uint64_t step;  // ignored here: the oracle takes the step index from its Schedule
};
static BoidsParams to_params(const OkgBoidsParams* q) {
  BoidsParams p;
  p.cohesion = q->cohesion;
  p.avoidance = q->avoidance;
  p.randomness = q->randomness;
  p.consistency = q->consistency;
  p.momentum = q->momentum;
  p.jump = q->jump;
  p.radius = q->radius;
  p.exact_query = q->exact_query;
  p.seed = q->seed;
  return p;
}

void* okg_flockers_new(float w, float h, float disc, int tor, uint32_t n, const OkgBoidsParams* q,
                       int canonical_order) {
  FlockSim* s = new FlockSim(w, h, n, disc, tor != 0, to_params(q));
  s->state.canonical_order = canonical_order != 0;
  return s;
}
void okg_flockers_free(void* s) { delete (FlockSim*)s; }
// Optional: place explicit agents instead of the Philox init (must precede okg_flockers_init)
void okg_flockers_preset(void* sp, uint64_t n, const uint32_t* id, const float* x, const float* y,
                         const float* ldx, const float* ldy) {
  FlockSim* s = (FlockSim*)sp;
  s->preset.clear();
  for (uint64_t i = 0; i < n; ++i)
    s->preset.push_back(Bird{id[i], Real2D{x[i], y[i]}, Real2D{ldx[i], ldy[i]}, false});
  s->state.preset = &s->preset;
}
int okg_flockers_init(void* sp) {
  OKG_TRY FlockSim* s = (FlockSim*)sp;
  s->schedule = Schedule();
  s->state.init(s->schedule);
  OKG_CATCH
}
int okg_flockers_step(void* sp, uint64_t nsteps) {
  OKG_TRY FlockSim* s = (FlockSim*)sp;
  for (uint64_t i = 0; i < nsteps; ++i) s->schedule.step_once(s->state);
  OKG_CATCH
}
// dynamic population (LifeRule); `next_id` = id given to the first child
void okg_flockers_set_life(void* sp, float death_prob, float birth_prob, uint32_t crowd_limit, uint32_t next_id) {
  FlockSim* s = (FlockSim*)sp;
  s->state.life_on = true;
  s->state.life.death_prob = death_prob;
  s->state.life.birth_prob = birth_prob;
  s->state.life.crowd_limit = crowd_limit;
  s->state.next_id = next_id;
}
// the scheduled agents (any ids), in the store's order; returns how many there are
uint64_t okg_flockers_population(void* sp, uint64_t cap, uint32_t* id, float* x, float* y, float* ldx, float* ldy,
                                 uint64_t* born, uint64_t* died) {
  FlockSim* s = (FlockSim*)sp;
  uint64_t k = 0;
  for (const auto& e : s->schedule.events.map) {
    const BirdAgent* a = static_cast<const BirdAgent*>(e.first.agent.get());
    if (k < cap) {
      id[k] = a->b.id;
      x[k] = a->b.pos.x;
      y[k] = a->b.pos.y;
      ldx[k] = a->b.last_d.x;
      ldy[k] = a->b.last_d.y;
    }
    ++k;
  }
  if (born) *born = s->state.born;
  if (died) *died = s->state.died;
  return k;
}
uint64_t okg_flockers_schedule_step(void* sp) { return ((FlockSim*)sp)->schedule.step; }
void* okg_flockers_field(void* sp) { return &((FlockSim*)sp)->state.field1; }
// agents' own copies held by the schedule, written at index == id (ids are 0..n-1 here)
int okg_flockers_agents(void* sp, uint64_t n, float* x, float* y, float* ldx, float* ldy) {
  OKG_TRY FlockSim* s = (FlockSim*)sp;
  for (const auto& e : s->schedule.events.map) {
    const BirdAgent* a = static_cast<const BirdAgent*>(e.first.agent.get());
    if (a->b.id >= n) rust_panic("okg_flockers_agents: id beyond output");
    x[a->b.id] = a->b.pos.x;
    y[a->b.id] = a->b.pos.y;
    ldx[a->b.id] = a->b.last_d.x;
    ldy[a->b.id] = a->b.last_d.y;
  }
  OKG_CATCH
}
// order in which the next Schedule::step will run the agents (pop order of the queue)
int okg_flockers_pop_order(void* sp, uint32_t* ids, uint64_t cap) {
  OKG_TRY FlockSim* s = (FlockSim*)sp;
  // replay on a copy of the index structure (agents themselves are not cloned)
  PriorityQueue q;
  for (const auto& e : s->schedule.events.map)
    q.push(AgentImpl{e.first.id, nullptr, e.first.repeating}, e.second);
  // NB: a fresh queue built in store order reproduces heap state only when all priorities are
  // equal (the Flockers case); good enough for the order test.
  uint64_t k = 0;
  while (!q.is_empty()) {
    auto it = q.pop();
    if (k < cap) ids[k] = it.first.id;
    ++k;
  }
  OKG_CATCH
}
// timed run for the CPU baseline: returns seconds spent in `nsteps` Schedule::step calls
double okg_flockers_time_steps(void* sp, uint64_t nsteps) {
  FlockSim* s = (FlockSim*)sp;
  auto t0 = std::chrono::steady_clock::now();
  for (uint64_t i = 0; i < nsteps; ++i) s->schedule.step_once(s->state);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// replica sweep on host threads, mirroring explore_parallel! (model_exploration.rs:387: one
// State + Schedule per task, no sharing).  Returns wall seconds; agent_steps = total work done.
double okg_flockers_sweep(float w, float h, float disc, int tor, uint32_t n,
                          const OkgBoidsParams* q, uint32_t replicas, uint64_t nsteps,
                          uint32_t threads, uint64_t* agent_steps) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (uint32_t t = 0; t < threads; ++t)
    pool.emplace_back([=]() {
      for (uint32_t r = t; r < replicas; r += threads) {
        OkgBoidsParams qq = *q;
        qq.seed = q->seed + r;
        FlockSim s(w, h, n, disc, tor != 0, to_params(&qq));
        s.state.init(s.schedule);
        for (uint64_t i = 0; i < nsteps; ++i) s.schedule.step_once(s.state);
      }
    });
  for (auto& th : pool) th.join();
  *agent_steps = (uint64_t)replicas * n * nsteps;
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ------------------------------------------------------------------ Schedule known-answer hooks
struct NopAgent : Agent {  // tests/utils/mynode.rs: a no-op agent
  uint32_t tag;
  explicit NopAgent(uint32_t t) : tag(t) {}
  void step(State&) override {}
  std::unique_ptr<Agent> clone() const override { return std::make_unique<NopAgent>(tag); }
};
void* okg_schedule_new() { return new Schedule(); }
void okg_schedule_free(void* s) { delete (Schedule*)s; }
int okg_schedule_repeating(void* s, uint32_t tag, float t, int ordering, uint32_t* id_out) {
  auto r = ((Schedule*)s)->distributed_schedule_repeating(std::make_unique<NopAgent>(tag), t, ordering);
  *id_out = r.first;
  return r.second ? 1 : 0;
}
int64_t okg_schedule_events(void* s, uint32_t* tags, uint64_t cap) {
  auto v = ((Schedule*)s)->get_all_events();
  for (size_t i = 0; i < v.size() && i < cap; ++i) tags[i] = static_cast<const NopAgent*>(v[i])->tag;
  return (int64_t)v.size();
}
int okg_schedule_dequeue(void* s, uint32_t id) { return ((Schedule*)s)->dequeue(id) ? 1 : 0; }

// ------------------------------------------------------------------ DenseNumberGrid2D<u16> / <u8>
typedef DenseNumberGrid2D<uint16_t> G16;
void* okg_grid_new(int w, int h) {
  try {
    return new G16(w, h);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void okg_grid_free(void* g) { delete (G16*)g; }
int okg_grid_set_value_location(void* g, uint16_t v, int x, int y) {
  OKG_TRY((G16*)g)->set_value_location(v, Int2D{x, y});
  OKG_CATCH
}
int okg_grid_remove_value_location(void* g, int x, int y) {
  OKG_TRY((G16*)g)->remove_value_location(Int2D{x, y});
  OKG_CATCH
}
// returns 1 = Some (value in *v), 0 = None, -1 = panic
int okg_grid_get_value(void* g, int x, int y, int unbuffered, uint16_t* v) {
  try {
    auto r = unbuffered ? ((G16*)g)->get_value_unbuffered(Int2D{x, y}) : ((G16*)g)->get_value(Int2D{x, y});
    if (!r) return 0;
    *v = *r;
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int okg_grid_get_location(void* g, uint16_t v, int unbuffered, int* x, int* y) {
  auto r = ((G16*)g)->get_location(v, unbuffered != 0);
  if (!r) return 0;
  *x = r->x;
  *y = r->y;
  return 1;
}
int64_t okg_grid_num_empty_bags(void* g) { return (int64_t)((G16*)g)->get_empty_bags().size(); }
void okg_grid_lazy_update(void* g) { ((G16*)g)->lazy_update(); }
void okg_grid_update(void* g) { ((G16*)g)->update(); }
// closure family for apply_to_all_values: op 0 -> constant c, op 1 -> value + c
void okg_grid_apply(void* g, int op, uint16_t c, int option) {
  GridOption o = option == 0 ? GridOption::READ : option == 1 ? GridOption::WRITE : GridOption::READWRITE;
  if (op == 0)
    ((G16*)g)->apply_to_all_values([c](uint16_t) { return c; }, o);
  else
    ((G16*)g)->apply_to_all_values([c](uint16_t v) { return (uint16_t)(v + c); }, o);
}
// dump a buffer x-major; None -> `none`
void okg_grid_dump(void* g, int unbuffered, uint16_t none, uint16_t* out) {
  G16* p = (G16*)g;
  const auto& v = p->locs[unbuffered ? p->write : p->read];
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i] ? *v[i] : none;
}

// ------------------------------------------------------------------ Forest Fire
void* okg_ff_new(int w, int h) { return new ForestFire(w, h); }
void okg_ff_free(void* f) { delete (ForestFire*)f; }
void okg_ff_init(void* f, float density, uint64_t seed) { ((ForestFire*)f)->init(density, seed); }
// explicit initial state, x-major bytes; `none` marks empty cells
void okg_ff_load(void* f, const uint8_t* cells, uint8_t none) {
  ForestFire* p = (ForestFire*)f;
  for (int32_t x = 0; x < p->grid.width; ++x)
    for (int32_t y = 0; y < p->grid.height; ++y) {
      uint8_t v = cells[(size_t)x * p->grid.height + y];
      if (v != none) p->grid.set_value_location(v, Int2D{x, y});
    }
  p->grid.lazy_update();
}
void okg_ff_step(void* f, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) ((ForestFire*)f)->step();
}
double okg_ff_time_steps(void* f, uint64_t n) {
  auto t0 = std::chrono::steady_clock::now();
  for (uint64_t i = 0; i < n; ++i) ((ForestFire*)f)->step();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
void okg_ff_dump(void* f, uint8_t none, uint8_t* out) {
  ForestFire* p = (ForestFire*)f;
  const auto& v = p->grid.locs[p->grid.read];
  for (size_t i = 0; i < v.size(); ++i) out[i] = v[i] ? *v[i] : none;
}

// ------------------------------------------------------------------ DenseGrid2D<GridObj>
// Objects are (id, tag) pairs that compare by id, like the fixture's Bird (bird.rs:168-172); the
// tag plays the part of Bird.flag in tests/engine/dense_object_grid_2d.rs.
struct GridObj {
  uint32_t id, tag;
  bool operator==(const GridObj& o) const { return id == o.id; }
};
typedef DenseGrid2D<GridObj> OG;
void* okg_ogrid_new(int w, int h) {
  try {
    return new OG(w, h);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void okg_ogrid_free(void* g) { delete (OG*)g; }
int okg_ogrid_set_object_location(void* g, uint32_t id, uint32_t tag, int x, int y) {
  OKG_TRY((OG*)g)->set_object_location(GridObj{id, tag}, Int2D{x, y});
  OKG_CATCH
}
int okg_ogrid_remove_object_location(void* g, uint32_t id, int x, int y) {
  OKG_TRY((OG*)g)->remove_object_location(GridObj{id, 0}, Int2D{x, y});
  OKG_CATCH
}
int okg_ogrid_lazy_update(void* g) {
  OKG_TRY((OG*)g)->lazy_update();
  OKG_CATCH
}
int okg_ogrid_update(void* g) {
  OKG_TRY((OG*)g)->update();
  OKG_CATCH
}
// number of bags the read / write Vec holds (update() grows the read Vec)
uint64_t okg_ogrid_nbags(void* g, int unbuffered) {
  OG* p = (OG*)g;
  return p->locs[unbuffered ? p->write : p->read].size();
}
// -1 = panic, -2 = None (empty bag), else the number of objects (written up to cap)
int64_t okg_ogrid_get_objects(void* g, int unbuffered, int x, int y, uint32_t* ids, uint32_t* tags, uint64_t cap) {
  try {
    auto v = ((OG*)g)->get_objects(Int2D{x, y}, unbuffered != 0);
    if (!v) return -2;
    for (size_t i = 0; i < v->size() && i < cap; ++i) {
      ids[i] = (*v)[i].id;
      tags[i] = (*v)[i].tag;
    }
    return (int64_t)v->size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// 1 = found (x, y written), 0 = None, -1 = panic
int okg_ogrid_get_location(void* g, int unbuffered, uint32_t id, int* x, int* y) {
  try {
    auto l = ((OG*)g)->get_location(GridObj{id, 0}, unbuffered != 0);
    if (!l) return 0;
    *x = l->x;
    *y = l->y;
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int64_t okg_ogrid_get_empty_bags(void* g, int* xs, int* ys, uint64_t cap) {
  try {
    auto v = ((OG*)g)->get_empty_bags();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
      xs[i] = v[i].x;
      ys[i] = v[i].y;
    }
    return (int64_t)v.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// iter_objects / iter_objects_unbuffered flattened: one (x, y, id, tag) row per closure call
int64_t okg_ogrid_iter_objects(void* g, int unbuffered, int* xs, int* ys, uint32_t* ids, uint32_t* tags,
                               uint64_t cap) {
  try {
    uint64_t n = 0;
    ((OG*)g)->iter_objects(
        [&](const Int2D& loc, const GridObj& o) {
          if (n < cap) {
            xs[n] = loc.x;
            ys[n] = loc.y;
            ids[n] = o.id;
            tags[n] = o.tag;
          }
          ++n;
        },
        unbuffered != 0);
    return (int64_t)n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// apply_to_all_values with a closure family: op 0 = Some(obj with tag = arg), 1 = None,
// 2 = None if tag == arg else Some(obj), 3 = Some(obj with tag = bag_id.x * 65536 + bag_id.y)
// (exposes the bag id the closure is handed).  Returns the number of closure calls, -1 on panic.
int64_t okg_ogrid_apply(void* g, int op, uint32_t arg, int option) {
  try {
    int64_t calls = 0;
    ((OG*)g)->apply_to_all_values(
        [&](const Int2D& bag, const GridObj& o) -> std::optional<GridObj> {
          ++calls;
          switch (op) {
            case 0: return GridObj{o.id, arg};
            case 1: return std::nullopt;
            case 2: return o.tag == arg ? std::nullopt : std::optional<GridObj>(o);
            default: return GridObj{o.id, (uint32_t)bag.x * 65536u + (uint32_t)bag.y};
          }
        },
        (GridOption)option);
    return calls;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ------------------------------------------------------------------ SparseGrid2D<GridObj>
typedef SparseGrid2D<GridObj> SG;
void* okg_sgrid_new(int w, int h) { return new SG(w, h); }
void okg_sgrid_free(void* g) { delete (SG*)g; }
int okg_sgrid_set_object_location(void* g, uint32_t id, uint32_t tag, int x, int y) {
  OKG_TRY((SG*)g)->set_object_location(GridObj{id, tag}, Int2D{x, y});
  OKG_CATCH
}
int okg_sgrid_remove_object_location(void* g, uint32_t id, int x, int y) {
  OKG_TRY((SG*)g)->remove_object_location(GridObj{id, 0}, Int2D{x, y});
  OKG_CATCH
}
int okg_sgrid_lazy_update(void* g) {
  OKG_TRY((SG*)g)->lazy_update();
  OKG_CATCH
}
int okg_sgrid_update(void* g) {
  OKG_TRY((SG*)g)->update();
  OKG_CATCH
}
int64_t okg_sgrid_get_objects(void* g, int unbuffered, int x, int y, uint32_t* ids, uint32_t* tags, uint64_t cap) {
  try {
    auto v = ((SG*)g)->get_objects(Int2D{x, y}, unbuffered != 0);
    if (!v) return -2;
    for (size_t i = 0; i < v->size() && i < cap; ++i) {
      ids[i] = (*v)[i].id;
      tags[i] = (*v)[i].tag;
    }
    return (int64_t)v->size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int okg_sgrid_get_location(void* g, int unbuffered, uint32_t id, int* x, int* y) {
  try {
    auto l = ((SG*)g)->get_location(GridObj{id, 0}, unbuffered != 0);
    if (!l) return 0;
    *x = l->x;
    *y = l->y;
    return 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int64_t okg_sgrid_get_empty_bags(void* g, int* xs, int* ys, uint64_t cap) {
  try {
    auto v = ((SG*)g)->get_empty_bags();
    for (size_t i = 0; i < v.size() && i < cap; ++i) {
      xs[i] = v[i].x;
      ys[i] = v[i].y;
    }
    return (int64_t)v.size();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
int64_t okg_sgrid_iter_objects(void* g, int unbuffered, int* xs, int* ys, uint32_t* ids, uint32_t* tags,
                               uint64_t cap) {
  try {
    uint64_t n = 0;
    ((SG*)g)->iter_objects(
        [&](const Int2D& loc, const GridObj& o) {
          if (n < cap) {
            xs[n] = loc.x;
            ys[n] = loc.y;
            ids[n] = o.id;
            tags[n] = o.tag;
          }
          ++n;
        },
        unbuffered != 0);
    return (int64_t)n;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}
// same closure family as okg_ogrid_apply (op 1 and a matching op 2 return None -> the reference panics)
int64_t okg_sgrid_apply(void* g, int op, uint32_t arg, int option) {
  try {
    int64_t calls = 0;
    ((SG*)g)->apply_to_all_values(
        [&](const Int2D& bag, const GridObj& o) -> std::optional<GridObj> {
          ++calls;
          switch (op) {
            case 0: return GridObj{o.id, arg};
            case 1: return std::nullopt;
            case 2: return o.tag == arg ? std::nullopt : std::optional<GridObj>(o);
            default: return GridObj{o.id, (uint32_t)bag.x * 65536u + (uint32_t)bag.y};
          }
        },
        (GridOption)option);
    return calls;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

unsigned okg_hardware_concurrency() { return std::thread::hardware_concurrency(); }

}  // extern "C"
