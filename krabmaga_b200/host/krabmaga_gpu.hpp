// C++ host mirror of the reference interface for the hot path, above the C ABI
// (include/krabgpu.h).  The reference is compiled code (Rust) whose toolchain is absent from this
// image, so the host side is mirrored in C++: same type and method names, argument meaning and
// error behaviour (a Rust panic == a thrown krabmaga::gpu::Panic) as
//   Field2D             src/engine/fields/field_2d.rs:269-921
//   DenseNumberGrid2D   src/engine/fields/dense_number_grid_2d.rs:90-561
//   Field               src/engine/fields/field.rs:4-9
//   Schedule/State/Agent  src/engine/schedule.rs:227-413, state.rs:45-63, agent.rs:7-40
//   Flocker / Bird      tests/model/flockers/{state,bird}.rs
//   simulate!           src/lib.rs:1158-1175
#pragma once
#include <cstdint>
#include <cstdio>
#include <chrono>
#include <cmath>
#include <memory>
#include <optional>
#include <queue>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/krabgpu.h"

namespace krabmaga {
namespace gpu {

struct Panic : std::runtime_error {
  int code;
  Panic(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != KG_OK) throw Panic(rc, kg_last_error());
}

struct Real2D { float x, y; };
struct Int2D { int32_t x, y; };

// tests/model/flockers/bird.rs:19-25
struct Bird {
  uint32_t id;
  Real2D pos;
  Real2D last_d;
  bool flag = false;
};

struct Field {  // field.rs:4-9
  virtual ~Field() = default;
  virtual void update() {}
  virtual void lazy_update() {}
};

class Field2D : public Field {
 public:
  float width, height, discretization;
  bool toroidal;
  // Field2D::new(w, h, d, t)  field_2d.rs:304-322
  Field2D(float w, float h, float d, bool t, uint64_t capacity, int device = 0)
      : width(w), height(h), discretization(d), toroidal(t) {
    check(kg_field2d_create(w, h, d, t ? 1 : 0, capacity, device, &h_));
  }
  ~Field2D() override { kg_field2d_destroy(h_); }
  Field2D(const Field2D&) = delete;
  Field2D& operator=(const Field2D&) = delete;

  // field_2d.rs:838-846
  void set_object_location(const Bird& o, Real2D loc) const {
    check(kg_field2d_set_object_locations(h_, 1, &o.id, &loc.x, &loc.y, &o.last_d.x, &o.last_d.y));
  }
  void set_object_locations(const std::vector<Bird>& v) const {
    std::vector<uint32_t> id;
    std::vector<float> x, y, dx, dy;
    for (const Bird& b : v) {
      id.push_back(b.id); x.push_back(b.pos.x); y.push_back(b.pos.y);
      dx.push_back(b.last_d.x); dy.push_back(b.last_d.y);
    }
    check(kg_field2d_set_object_locations(h_, v.size(), id.data(), x.data(), y.data(), dx.data(), dy.data()));
  }
  // field_2d.rs:885-898
  void remove_object_location(const Bird& o, Real2D loc) const {
    check(kg_field2d_remove_object_location(h_, o.id, loc.x, loc.y));
  }
  // field_2d.rs:386-440 / :472-516 — neighbour ids in the reference's order
  std::vector<uint32_t> get_neighbors_within_distance(Real2D loc, float dist) const {
    return neighbors(loc, dist, KG_QUERY_EXACT);
  }
  std::vector<uint32_t> get_neighbors_within_relax_distance(Real2D loc, float dist) const {
    return neighbors(loc, dist, KG_QUERY_RELAX);
  }
  // field_2d.rs:546-575
  std::vector<uint32_t> get_objects(Real2D loc, bool unbuffered = false) const {
    uint64_t n = 0, cap = 1024;
    for (;;) {
      std::vector<uint32_t> ids(cap);
      int rc = kg_field2d_get_objects(h_, unbuffered ? KG_BUF_WRITE : KG_BUF_READ, loc.x, loc.y, cap,
                                      ids.data(), &n);
      if (rc == KG_E_CAPACITY && n > cap) { cap = n; continue; }
      check(rc);
      ids.resize(n);
      return ids;
    }
  }
  std::vector<uint32_t> get_objects_unbuffered(Real2D loc) const { return get_objects(loc, true); }
  // field_2d.rs:806-811
  size_t num_objects_at_location(Real2D loc) const {
    uint32_t out = 0;
    check(kg_field2d_num_objects_at_locations(h_, 1, &loc.x, &loc.y, &out));
    return out;
  }
  size_t num_empty_bags() const {  // get_empty_bags().len()  field_2d.rs:718-730
    uint64_t n = 0;
    check(kg_field2d_num_empty_bags(h_, &n));
    return n;
  }
  size_t nagents() const {
    uint64_t n = 0;
    check(kg_field2d_nagents(h_, &n));
    return n;
  }
  // read buffer in iter_objects order (field_2d.rs:594-626)
  std::vector<Bird> objects(bool unbuffered = false) const {
    uint64_t n = 0;
    check(kg_field2d_num_objects(h_, unbuffered ? KG_BUF_WRITE : KG_BUF_READ, &n));
    std::vector<uint32_t> id(n);
    std::vector<float> x(n), y(n), dx(n), dy(n);
    check(kg_field2d_download(h_, unbuffered ? KG_BUF_WRITE : KG_BUF_READ, n, id.data(), x.data(),
                              y.data(), dx.data(), dy.data(), nullptr, &n));
    std::vector<Bird> out(n);
    for (uint64_t i = 0; i < n; ++i) out[i] = Bird{id[i], {x[i], y[i]}, {dx[i], dy[i]}, false};
    return out;
  }
  void update() override { check(kg_field2d_update(h_)); }
  void lazy_update() override { check(kg_field2d_lazy_update(h_)); }  // field_2d.rs:905-921
  // all agents' Bird::step  bird.rs:39-155
  void step_boids(KgBoidsParams p, uint64_t schedule_step) const {
    p.step = schedule_step;
    check(kg_field2d_step_boids(h_, &p));
  }
  // One whole step for a model that keeps its birds in host arrays (bird i has id i, state.rs:47): positions
  // and last directions go up, every Bird::step runs, the results come back at the same indices.
  void step_boids_in_place(KgBoidsParams p, uint64_t schedule_step, std::vector<float>& x, std::vector<float>& y,
                           std::vector<float>& last_dx, std::vector<float>& last_dy) const {
    const uint64_t n = x.size();
    if (y.size() != n || last_dx.size() != n || last_dy.size() != n) throw std::invalid_argument("array lengths differ");
    p.step = schedule_step;
    check(kg_field2d_step_boids_host_ordered(h_, &p, n, nullptr, x.data(), y.data(), last_dx.data(), last_dy.data(),
                                             x.data(), y.data(), last_dx.data(), last_dy.data()));
  }
  void init_flockers(uint64_t n, uint64_t seed) const { check(kg_field2d_init_flockers(h_, n, seed)); }
  void set_canonical_order(bool on) const { check(kg_field2d_set_order(h_, on ? KG_ORDER_CANONICAL : KG_ORDER_ANY)); }
  void sync() const { check(kg_field2d_sync(h_)); }

 private:
  std::vector<uint32_t> neighbors(Real2D loc, float dist, int mode) const {
    uint64_t offs[2] = {0, 0}, total = 0, cap = 1024;
    for (;;) {
      std::vector<uint32_t> ids(cap);
      int rc = kg_field2d_neighbors(h_, 1, &loc.x, &loc.y, dist, mode, offs, ids.data(), cap, &total);
      if (rc == KG_E_CAPACITY && total > cap) { cap = total; continue; }
      check(rc);
      ids.resize(total);
      return ids;
    }
  }
  kg_field2d* h_ = nullptr;
};

// DenseNumberGrid2D<u8>, None == 0xFF on the device
class DenseNumberGrid2D : public Field {
 public:
  int32_t width, height;
  DenseNumberGrid2D(int32_t w, int32_t h, int device = 0) : width(w < 0 ? -w : w), height(h < 0 ? -h : h) {
    check(kg_grid_create(w, h, 1, 0xFF, device, &h_));
  }
  ~DenseNumberGrid2D() override { kg_grid_destroy(h_); }
  std::optional<uint8_t> get_value(const Int2D& loc) const { return get(loc, KG_BUF_READ); }
  std::optional<uint8_t> get_value_unbuffered(const Int2D& loc) const { return get(loc, KG_BUF_WRITE); }
  void set_value_location(uint8_t v, const Int2D& loc) const { check(kg_grid_set_values(h_, 1, &loc.x, &loc.y, &v)); }
  void remove_value_location(const Int2D& loc) const { check(kg_grid_remove_values(h_, 1, &loc.x, &loc.y)); }
  size_t num_empty_bags() const {
    uint64_t n = 0;
    check(kg_grid_num_empty(h_, &n));
    return n;
  }
  void lazy_update() override { check(kg_grid_lazy_update(h_)); }
  void update() override { check(kg_grid_update(h_)); }
  void step_forest_fire() const { check(kg_grid_step_stencil(h_, KG_RULE_FOREST_FIRE)); }
  void init_forest_fire(float density, uint64_t seed) const { check(kg_grid_init_forest_fire(h_, density, seed)); }
  std::vector<uint8_t> cells(bool unbuffered = false) const {
    std::vector<uint8_t> out((size_t)width * height);
    check(kg_grid_download(h_, unbuffered ? KG_BUF_WRITE : KG_BUF_READ, out.data()));
    return out;
  }

 private:
  std::optional<uint8_t> get(const Int2D& loc, int which) const {
    uint8_t v = 0xFF;
    check(kg_grid_get_values(h_, which, 1, &loc.x, &loc.y, &v));
    return v == 0xFF ? std::nullopt : std::optional<uint8_t>(v);
  }
  kg_grid* h_ = nullptr;
};

// ---------------------------------------------------------------- engine contracts (host side)
struct State;
class Schedule;
struct Agent {  // agent.rs:7-40
  virtual ~Agent() = default;
  virtual void step(State& state) = 0;
  virtual bool is_stopped(State&) { return false; }
  virtual void before_step(State&) {}
  virtual void after_step(State&) {}
};
struct State {  // state.rs:45-63
  virtual ~State() = default;
  virtual void init(Schedule& schedule) = 0;
  virtual void reset() = 0;
  virtual void update(uint64_t step) = 0;
  virtual void before_step(Schedule&) {}
  virtual void after_step(Schedule&) {}
  virtual bool end_condition(Schedule&) { return false; }
};

// Sequential Schedule  schedule.rs:227-413 (kept on the host; one proxy agent per population)
class Schedule {
  struct Ev {
    float time;
    int32_t ordering;
    uint64_t seq;
    uint32_t id;
    std::shared_ptr<Agent> agent;
    bool repeating;
    bool operator<(const Ev& o) const {  // priority.rs:21-38: lower time, then lower ordering first
      if (time != o.time) return time > o.time;
      if (ordering != o.ordering) return ordering > o.ordering;
      return seq > o.seq;
    }
  };
  std::priority_queue<Ev> events_;
  uint64_t seq_ = 0;

 public:
  uint64_t step = 0;
  float time = 0.f;
  uint32_t agent_ids_counting = 0;
  bool schedule_repeating(std::shared_ptr<Agent> a, float t, int32_t ordering) {  // :295-303
    events_.push(Ev{t, ordering, seq_++, agent_ids_counting++, std::move(a), true});
    return true;
  }
  void step_once(State& state) {  // :347-413
    if (step == 0) state.update(step);
    state.before_step(*this);
    if (events_.empty()) {
      state.after_step(*this);
      step += 1;
      state.update(step);
      return;
    }
    time = events_.top().time;
    std::vector<Ev> cevents;
    while (!events_.empty() && !(events_.top().time > time)) {
      cevents.push_back(events_.top());
      events_.pop();
    }
    for (Ev& e : cevents) {
      e.agent->before_step(state);
      e.agent->step(state);
      e.agent->after_step(state);
      if (e.repeating && !e.agent->is_stopped(state)) {
        e.time += 1.0f;
        e.seq = seq_++;
        events_.push(e);
      }
    }
    state.after_step(*this);
    step += 1;
    state.update(step);
  }
};

// tests/model/flockers/state.rs with the field type swapped for the GPU field
struct Flocker : State {
  uint64_t step = 0;
  std::unique_ptr<Field2D> field1;
  uint32_t initial_flockers;
  float dim0, dim1, discretization;
  bool toroidal;
  KgBoidsParams params;
  bool canonical_order = false;
  struct Flock : Agent {  // proxy for every Bird::step
    void step(State& st) override {
      Flocker& s = static_cast<Flocker&>(st);
      s.field1->step_boids(s.params, s.step);
    }
  };
  Flocker(float w, float h, uint32_t n, float disc, bool tor, KgBoidsParams p)
      : initial_flockers(n), dim0(w), dim1(h), discretization(disc), toroidal(tor), params(p) {
    reset();
  }
  void reset() override {  // state.rs:35-38
    step = 0;
    field1 = std::make_unique<Field2D>(dim0, dim1, discretization, toroidal, initial_flockers ? initial_flockers : 1);
  }
  void init(Schedule& schedule) override {  // state.rs:41-56
    field1->set_canonical_order(canonical_order);
    field1->init_flockers(initial_flockers, params.seed);
    schedule.schedule_repeating(std::make_shared<Flock>(), 0.f, 0);
  }
  void update(uint64_t s) override {  // state.rs:58-60
    step = s;
    field1->lazy_update();
  }
};

// simulate!(state, steps, reps, false)  lib.rs:1158-1175
inline void simulate(State& state, uint64_t n_step, uint32_t reps) {
  for (uint32_t r = 0; r < reps; ++r) {
    Schedule schedule;
    state.init(schedule);
    for (uint64_t i = 0; i < n_step; ++i) {
      schedule.step_once(state);
      if (state.end_condition(schedule)) break;
    }
  }
}

// ---------------------------------------------------------------------------- explore
// R independent Flocker states advanced together on one GPU (kg_batch_*): what the rayon tasks of
// explore_parallel! (src/explore/model_exploration.rs:387-420) each do for one configuration.
class FlockerBatch {
 public:
  FlockerBatch(float w, float h, float disc, bool toroidal, const std::vector<KgBoidsParams>& params,
               uint32_t agents, int device = 0, bool canonical_order = false)
      : replicas((uint32_t)params.size()), agents(agents) {
    check(kg_batch_create(w, h, disc, toroidal ? 1 : 0, replicas, agents, device, &h_));
    check(kg_batch_set_params(h_, 0, replicas, params.data()));
    if (canonical_order) check(kg_batch_set_order(h_, KG_ORDER_CANONICAL));
  }
  ~FlockerBatch() { kg_batch_destroy(h_); }
  FlockerBatch(const FlockerBatch&) = delete;
  FlockerBatch& operator=(const FlockerBatch&) = delete;
  // simulate_explore! (model_exploration.rs:160-190) for every replica: init, then nstep steps
  void simulate(uint64_t nstep) {
    check(kg_batch_init_flockers(h_));
    check(kg_batch_lazy_update(h_));
    check(kg_batch_run_boids(h_, 0, nstep));
    check(kg_batch_sync(h_));
  }
  // every replica's birds in its iter_objects order, replica-major
  std::vector<Bird> objects() const {
    const size_t n = (size_t)replicas * agents;
    std::vector<uint32_t> id(n);
    std::vector<float> x(n), y(n), dx(n), dy(n);
    check(kg_batch_download(h_, id.data(), x.data(), y.data(), dx.data(), dy.data(), nullptr));
    std::vector<Bird> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = Bird{id[i], {x[i], y[i]}, {dx[i], dy[i]}};
    return out;
  }
  const uint32_t replicas, agents;

 private:
  kg_batch* h_ = nullptr;
};

// One row of the sweep's data frame (build_dataframe!, model_exploration.rs:451-540)
struct FrameRow {
  uint32_t conf_num, conf_rep;
  KgBoidsParams input;
  float output;  // flock polarisation |mean last_d| / jump
  float run_duration, step_per_sec;

  // DataFrame trait, src/lib.rs:1794-1800
  static const std::vector<std::string>& field_names() {
    static const std::vector<std::string> names = {"conf_num", "conf_rep", "cohesion", "avoidance", "randomness",
                                                   "consistency", "momentum", "jump", "radius", "seed",
                                                   "polarisation", "run_duration", "step_per_sec"};
    return names;
  }
  std::vector<std::string> to_string() const {
    auto f = [](float v) {
      char buf[32];
      std::snprintf(buf, sizeof buf, "%.9g", (double)v);
      return std::string(buf);
    };
    return {std::to_string(conf_num), std::to_string(conf_rep), f(input.cohesion), f(input.avoidance),
            f(input.randomness), f(input.consistency), f(input.momentum), f(input.jump), f(input.radius),
            std::to_string((unsigned long long)input.seed), f(output), f(run_duration), f(step_per_sec)};
  }
};

// write_csv(name, &dataframe), src/lib.rs:1781-1792: "<name>.csv", header then one record per row
template <class Row>
inline void write_csv(const std::string& name, const std::vector<Row>& dataframe) {
  const std::string path = name + ".csv";
  std::FILE* fh = std::fopen(path.c_str(), "w");
  if (!fh) throw Panic(KG_E_INVALID, "error on open the file path: " + path);
  auto record = [&](const std::vector<std::string>& cells) {
    for (size_t i = 0; i < cells.size(); ++i) std::fprintf(fh, "%s%s", i ? "," : "", cells[i].c_str());
    std::fputc('\n', fh);
  };
  record(Row::field_names());
  for (const Row& r : dataframe) record(r.to_string());
  std::fclose(fh);
}

// explore_parallel!(nstep, rep_conf, State, input {...}, output [...], ExploreMode::Matched): one
// configuration per entry of `confs`, `rep_conf` repetitions each (seed + repetition), all runs of
// the sweep as one batch
inline std::vector<FrameRow> explore_parallel(uint64_t nstep, uint32_t rep_conf, float w, float h, float disc,
                                              uint32_t agents, const std::vector<KgBoidsParams>& confs,
                                              int device = 0, bool canonical_order = false) {
  std::vector<KgBoidsParams> runs;
  for (size_t i = 0; i < confs.size(); ++i)
    for (uint32_t r = 0; r < rep_conf; ++r) {
      KgBoidsParams p = confs[i];
      p.seed += r;
      runs.push_back(p);
    }
  FlockerBatch batch(w, h, disc, true, runs, agents, device, canonical_order);
  auto t0 = std::chrono::steady_clock::now();
  batch.simulate(nstep);
  const float dt = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
  const std::vector<Bird> birds = batch.objects();
  std::vector<FrameRow> rows(runs.size());
  for (size_t k = 0; k < runs.size(); ++k) {
    double sx = 0, sy = 0;
    for (uint32_t a = 0; a < agents; ++a) {
      sx += birds[k * agents + a].last_d.x;
      sy += birds[k * agents + a].last_d.y;
    }
    const float pol = (float)(std::sqrt(sx * sx + sy * sy) / agents / runs[k].jump);
    rows[k] = FrameRow{(uint32_t)(k / rep_conf), (uint32_t)(k % rep_conf), runs[k], pol, dt, nstep / dt};
  }
  return rows;
}

}  // namespace gpu
}  // namespace krabmaga
