// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under krabmaga_b200/ may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker.
//
// CPU restatement (C++17, single-threaded) of krABMaga 0.6.1's agent-step hot path.
// The reference is Rust and cannot be built in this image (no cargo/rustc), so this
// file restates its algorithm function by function; every block cites the reference
// lines it follows (paths relative to /root/reference).
//
// PARITY STATUS
//   * Field2D / DenseNumberGrid2D / Schedule semantics: PINNED by the reference's own
//     known-answer tests (tests/engine/field_2d.rs, dense_number_grid_2d.rs,
//     schedule.rs, tests/explore/simulate.rs), ported in tests/test_oracle_*.py.
//   * Bird::step numeric output (positions): PARITY UNPINNED — the reference ships no
//     numeric golden vectors and uses an OS-seeded RNG; only this restatement pins it.
//   * priority-queue 2.0.2 (Cargo.toml:27, un-vendored): pop order restated from the
//     crate's published algorithm (index-map store + binary heap of indices); weakly
//     pinned by tests/engine/schedule.rs:23-28.
//
// Compile with -O2/-O3 -ffp-contract=off -fno-fast-math: Rust never contracts a*b+c
// into an FMA, `%` on f32 is fmodf, and float->int `as` casts saturate (NaN -> 0).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "philox.hpp"

namespace oracle {

// ---------------------------------------------------------------- Rust scalar semantics
inline int32_t f32_as_i32(float v) {  // Rust `v as i32` (saturating, NaN -> 0)
  if (v != v) return 0;
  if (v >= 2147483648.0f) return INT32_MAX;
  if (v <= -2147483648.0f) return INT32_MIN;
  return (int32_t)v;
}
inline size_t f32_as_usize(float v) {  // Rust `v as usize`
  if (v != v || v <= 0.0f) return 0;
  if (v >= 18446744073709551616.0f) return SIZE_MAX;
  return (size_t)v;
}
[[noreturn]] inline void rust_panic(const std::string& what) { throw std::out_of_range(what); }

// ---------------------------------------------------------------- location.rs:9-41
struct Real2D {
  float x, y;
};
struct Int2D {
  int32_t x, y;
};
inline bool operator==(const Real2D& a, const Real2D& b) { return a.x == b.x && a.y == b.y; }

// ---------------------------------------------------------------- field_2d.rs:926-1014 helpers
// t_transform  field_2d.rs:926-932
inline int32_t t_transform(int32_t n, int32_t size) {
  if (n >= 0) return n % size;
  return (n % size) + size;
}
// toroidal_transform  field_2d.rs:1004-1014  (f32 `%` == fmodf; may return exactly dim)
inline float toroidal_transform(float val, float dim) {
  if (val >= 0.0f && val < dim) return val;
  float v = std::fmod(val, dim);
  if (v < 0.0f) v += dim;
  return v;
}
// toroidal_distance  field_2d.rs:988-1002
inline float toroidal_distance(float val1, float val2, float dim) {
  if (std::fabs(val1 - val2) <= dim / 2.0f) return val1 - val2;
  float d = toroidal_transform(val1, dim) - toroidal_transform(val2, dim);
  if (d * 2.0f > dim) return d - dim;
  if (d * 2.0f < -dim) return d + dim;
  return d;
}
// distance  field_2d.rs:974-986
inline float distance(const Real2D& a, const Real2D& b, float dim1, float dim2, bool tor) {
  float dx, dy;
  if (tor) {
    dx = toroidal_distance(a.x, b.x, dim1);
    dy = toroidal_distance(a.y, b.y, dim2);
  } else {
    dx = a.x - b.x;
    dy = a.y - b.y;
  }
  return std::sqrt(dx * dx + dy * dy);
}
// check_circle  field_2d.rs:934-972 : 1 = all four corners inside, -1 = all outside, 0 = mixed
inline int check_circle(const Int2D& bag, float disc, float width, float height, const Real2D& loc,
                        float dis, bool tor) {
  Real2D nw{(float)bag.x * disc, (float)bag.y * disc};
  Real2D ne{nw.x, std::fmin(nw.y + disc, height)};
  Real2D sw{std::fmin(nw.x + disc, width), nw.y};
  Real2D se{sw.x, ne.y};
  float d0 = distance(nw, loc, width, height, tor), d1 = distance(ne, loc, width, height, tor);
  float d2 = distance(sw, loc, width, height, tor), d3 = distance(se, loc, width, height, tor);
  if (d0 <= dis && d1 <= dis && d2 <= dis && d3 <= dis) return 1;
  if (d0 > dis && d1 > dis && d2 > dis && d3 > dis) return -1;
  return 0;
}

// ---------------------------------------------------------------- Field2D  field_2d.rs:269-921
// Default (non-`parallel`) variant: two bag grids Vec<Vec<O>>, x-major index x*dh+y,
// (ceil(w/d)+1) x (ceil(h/d)+1) cells.  O needs `.id` and `.pos`.
template <class O>
struct Field2D {
  std::vector<std::vector<O>> bags[2];
  int read = 0, write = 1;
  size_t nagents = 0;
  float width, height, discretization;
  bool toroidal;
  int32_t dh, dw;
  size_t density_estimation = 0;
  bool density_estimation_check = false;

  // new  field_2d.rs:304-322.  NB the bag count is evaluated in f32 (:307-308) while
  // dw/dh are integers (:317-318); above 2^24 cells the two can differ by one.
  Field2D(float w, float h, float d, bool t)
      : width(w), height(h), discretization(d), toroidal(t) {
    size_t nb = f32_as_usize((std::ceil(w / d) + 1.0f) * (std::ceil(h / d) + 1.0f));
    bags[0].assign(nb, {});
    bags[1].assign(nb, {});
    dh = f32_as_i32(std::ceil(h / d)) + 1;
    dw = f32_as_i32(std::ceil(w / d)) + 1;
  }

  // discretize  field_2d.rs:328-339
  Int2D discretize(const Real2D& loc) const {
    return Int2D{f32_as_i32(std::floor(loc.x / discretization)),
                 f32_as_i32(std::floor(loc.y / discretization))};
  }
  // not_discretize  field_2d.rs:345-353
  Real2D not_discretize(const Int2D& loc) const {
    return Real2D{(float)loc.x * discretization, (float)loc.y * discretization};
  }
  // `((bag.x * self.dh) + bag.y) as usize` then Vec indexing (panics when out of range)
  size_t bag_index(const Int2D& bag, const std::vector<std::vector<O>>& b) const {
    int64_t idx = (int64_t)(int32_t)((uint32_t)bag.x * (uint32_t)dh + (uint32_t)bag.y);
    if (idx < 0 || (size_t)idx >= b.size()) rust_panic("Field2D: bag index out of bounds");
    return (size_t)idx;
  }

  struct Window {
    int32_t min_i, max_i, min_j, max_j, max_x, max_y;
  };
  // window set-up shared by both queries  field_2d.rs:401-416 / :485-500
  Window window(const Real2D& loc, float dist) const {
    int32_t disc_dist = f32_as_i32(std::floor(dist / discretization));
    Int2D dl = discretize(loc);
    Window w;
    w.max_x = f32_as_i32(std::ceil(width / discretization));
    w.max_y = f32_as_i32(std::ceil(height / discretization));
    w.min_i = dl.x - disc_dist;
    w.max_i = dl.x + disc_dist;
    w.min_j = dl.y - disc_dist;
    w.max_j = dl.y + disc_dist;
    if (toroidal) {  // CLAMP, not wrap  :411-416
      w.min_i = std::max(0, w.min_i);
      w.max_i = std::min(w.max_i, w.max_x - 1);
      w.min_j = std::max(0, w.min_j);
      w.max_j = std::min(w.max_j, w.max_y - 1);
    }
    return w;
  }

  // get_neighbors_within_distance  field_2d.rs:386-440
  std::vector<O> get_neighbors_within_distance(Real2D loc, float dist) const {
    std::vector<O> neighbors;
    if (density_estimation_check) neighbors.reserve(density_estimation * 2);
    if (dist <= 0.0f) return neighbors;
    Window w = window(loc, dist);
    const auto& rb = bags[read];
    for (int32_t i = w.min_i; i < w.max_i + 1; ++i) {
      for (int32_t j = w.min_j; j < w.max_j + 1; ++j) {
        Int2D bag_id{t_transform(i, w.max_x), t_transform(j, w.max_y)};
        int check = check_circle(bag_id, discretization, width, height, loc, dist, toroidal);
        size_t index = bag_index(bag_id, rb);
        for (const O& elem : rb[index]) {
          if ((check == 0 && distance(loc, elem.pos, width, height, toroidal) <= dist) ||
              check == 1)
            neighbors.push_back(elem);
        }
      }
    }
    return neighbors;
  }

  // get_neighbors_within_relax_distance  field_2d.rs:472-516
  std::vector<O> get_neighbors_within_relax_distance(Real2D loc, float dist) const {
    std::vector<O> neighbors;
    if (density_estimation_check) neighbors.reserve(density_estimation * 2);
    if (dist <= 0.0f) return neighbors;
    Window w = window(loc, dist);
    const auto& rb = bags[read];
    for (int32_t i = w.min_i; i < w.max_i + 1; ++i) {
      for (int32_t j = w.min_j; j < w.max_j + 1; ++j) {
        Int2D bag_id{t_transform(i, w.max_x), t_transform(j, w.max_y)};
        size_t index = bag_index(bag_id, rb);
        for (const O& elem : rb[index]) neighbors.push_back(elem);
      }
    }
    return neighbors;
  }

  // get_objects / get_objects_unbuffered  field_2d.rs:546-575
  std::vector<O> get_objects(Real2D loc) const {
    return bags[read][bag_index(discretize(loc), bags[read])];
  }
  std::vector<O> get_objects_unbuffered(Real2D loc) const {
    return bags[write][bag_index(discretize(loc), bags[write])];
  }
  // iter_objects / iter_objects_unbuffered  field_2d.rs:594-660 (x outer, y inner, cell origin)
  template <class F>
  void iter_objects(F&& f, bool unbuffered = false) const {
    const auto& b = bags[unbuffered ? write : read];
    for (int32_t i = 0; i < dw; ++i)
      for (int32_t j = 0; j < dh; ++j) {
        size_t index = (size_t)(i * dh + j);
        if (index >= b.size()) rust_panic("Field2D::iter_objects index");
        if (!b[index].empty()) {
          Real2D rp = not_discretize(Int2D{i, j});
          for (const O& o : b[index]) f(rp, o);
        }
      }
  }
  // get_empty_bags  field_2d.rs:718-730
  std::vector<Real2D> get_empty_bags() const {
    std::vector<Real2D> out;
    for (int32_t i = 0; i < dw; ++i)
      for (int32_t j = 0; j < dh; ++j) {
        size_t index = (size_t)(i * dh + j);
        if (index >= bags[read].size()) rust_panic("Field2D::get_empty_bags index");
        if (bags[read][index].empty()) out.push_back(not_discretize(Int2D{i, j}));
      }
    return out;
  }
  // num_objects_at_location  field_2d.rs:806-811
  size_t num_objects_at_location(Real2D loc) const {
    return bags[read][bag_index(discretize(loc), bags[read])].size();
  }
  // set_object_location  field_2d.rs:838-846
  void set_object_location(const O& object, Real2D loc) {
    size_t index = bag_index(discretize(loc), bags[write]);
    bags[write][index].push_back(object);
    if (!density_estimation_check) nagents += 1;
  }
  // remove_object_location  field_2d.rs:885-898
  void remove_object_location(const O& object, Real2D loc) {
    size_t index = bag_index(discretize(loc), bags[write]);
    auto& bag = bags[write][index];
    if (!bag.empty()) {
      size_t before = bag.size();
      bag.erase(std::remove_if(bag.begin(), bag.end(),
                               [&](const O& x) { return x.id == object.id; }),
                bag.end());
      size_t after = bag.size();
      if (!density_estimation_check) nagents -= before - after;
    }
  }
  // Field::update is a no-op for Field2D  field_2d.rs:903
  void update() {}
  // Field::lazy_update  field_2d.rs:905-921
  void lazy_update() {
    std::swap(read, write);
    if (!density_estimation_check) {
      density_estimation = nagents / (size_t)(dw * dh);
      density_estimation_check = true;
      bags[write].clear();
      bags[write].resize((size_t)(dw * dh));
      for (auto& b : bags[write]) b.reserve(density_estimation);
    } else {
      for (auto& b : bags[write]) b.clear();
    }
  }
};

// ---------------------------------------------------------------- DenseNumberGrid2D
// dense_number_grid_2d.rs:90-561, default variant.  Flat index x*height + y.
enum class GridOption { READ, WRITE, READWRITE };  // grid_option.rs:3-10

template <class T>
struct DenseNumberGrid2D {
  std::vector<std::optional<T>> locs[2];
  int read = 0, write = 1;
  int32_t width, height;

  // new  :112-126 (sized with the signed product, stores abs())
  DenseNumberGrid2D(int32_t w, int32_t h) {
    int64_t n = (int64_t)(int32_t)((uint32_t)w * (uint32_t)h);
    if (n < 0) rust_panic("DenseNumberGrid2D::new capacity overflow");
    locs[0].assign((size_t)n, std::nullopt);
    locs[1].assign((size_t)n, std::nullopt);
    width = std::abs(w);
    height = std::abs(h);
  }
  size_t index_of(const Int2D& loc, const std::vector<std::optional<T>>& v) const {
    int64_t idx = (int64_t)(int32_t)((uint32_t)loc.x * (uint32_t)height + (uint32_t)loc.y);
    if (idx < 0 || (size_t)idx >= v.size()) rust_panic("DenseNumberGrid2D: index out of bounds");
    return (size_t)idx;
  }
  // apply_to_all_values  :155-195
  template <class F>
  void apply_to_all_values(F&& closure, GridOption option) {
    switch (option) {
      case GridOption::READ:
        for (auto& v : locs[read]) {
          if (!v) continue;
          v = closure(*v);
        }
        break;
      case GridOption::WRITE:
        for (size_t i = 0; i < locs[read].size(); ++i) {
          if (!locs[read][i]) continue;
          locs[write][i] = closure(*locs[read][i]);
        }
        break;
      case GridOption::READWRITE:
        for (size_t i = 0; i < locs[read].size(); ++i) {
          if (locs[write][i])
            locs[write][i] = closure(*locs[write][i]);
          else if (locs[read][i])
            locs[write][i] = closure(*locs[read][i]);
        }
        break;
    }
  }
  // get_location / get_location_unbuffered  :204-229 (x outer, y inner, first match)
  std::optional<Int2D> get_location(T value, bool unbuffered = false) const {
    const auto& v = locs[unbuffered ? write : read];
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j) {
        const auto& e = v.at((size_t)(i * height + j));
        if (e && *e == value) return Int2D{i, j};
      }
    return std::nullopt;
  }
  // get_empty_bags  :236-247
  std::vector<Int2D> get_empty_bags() const {
    std::vector<Int2D> out;
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j)
        if (!locs[read].at((size_t)(i * height + j))) out.push_back(Int2D{i, j});
    return out;
  }
  // get_value :350-354 / get_value_unbuffered :376-380
  std::optional<T> get_value(const Int2D& loc) const { return locs[read][index_of(loc, locs[read])]; }
  std::optional<T> get_value_unbuffered(const Int2D& loc) const {
    return locs[write][index_of(loc, locs[write])];
  }
  // iter_values / iter_values_unbuffered  :404-452
  template <class F>
  void iter_values(F&& f, bool unbuffered = false) const {
    const auto& v = locs[unbuffered ? write : read];
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j) {
        const auto& e = v.at((size_t)(i * height + j));
        if (e) f(Int2D{i, j}, *e);
      }
  }
  // set_value_location :492-495 / remove_value_location :526-530
  void set_value_location(T value, const Int2D& loc) { locs[write][index_of(loc, locs[write])] = value; }
  void remove_value_location(const Int2D& loc) {
    locs[write][index_of(loc, locs[write])] = std::nullopt;
  }
  // Field::lazy_update :537-545 : swap, then every write cell := None
  void lazy_update() {
    std::swap(read, write);
    for (auto& v : locs[write]) v = std::nullopt;
  }
  // Field::update :553-561 : copy write -> read, clear write
  void update() {
    for (size_t i = 0; i < locs[write].size(); ++i) {
      locs[read][i] = locs[write][i];
      locs[write][i] = std::nullopt;
    }
  }
};

// ---------------------------------------------------------------- engine contracts
struct State;
struct Schedule;
// agent.rs:7-40
struct Agent {
  virtual ~Agent() = default;
  virtual void step(State& state) = 0;
  virtual bool is_stopped(State&) { return false; }
  virtual void before_step(State&) {}
  virtual void after_step(State&) {}
  virtual std::unique_ptr<Agent> clone() const = 0;
};
// state.rs:45-63
struct State {
  virtual ~State() = default;
  virtual void init(Schedule& schedule) = 0;
  virtual void reset() = 0;
  virtual void update(uint64_t step) = 0;
  virtual void before_step(Schedule&) {}
  virtual void after_step(Schedule&) {}
  virtual bool end_condition(Schedule&) { return false; }
};

// priority.rs:4-38 : lower time first, then lower ordering; cmp() returns Greater for "pops first"
struct Priority {
  float time;
  int32_t ordering;
};
inline int priority_cmp(const Priority& a, const Priority& b) {
  if (a.time < b.time) return 1;
  if (a.time > b.time) return -1;
  if (a.ordering < b.ordering) return 1;
  if (a.ordering > b.ordering) return -1;
  return 0;
}
// agentimpl.rs:15-19
struct AgentImpl {
  uint32_t id;
  std::unique_ptr<Agent> agent;
  bool repeating;
};

// priority-queue 2.0.2 `PriorityQueue<I,P>` (un-vendored dependency, Cargo.toml:27), restated
// from its published design: `map` (insertion-ordered store, swap_remove on delete), `heap`
// (binary max-heap of map indices), `qp` (map index -> heap position), id -> map index hash.
struct PriorityQueue {
  std::vector<std::pair<AgentImpl, Priority>> map;
  std::vector<size_t> heap, qp;
  std::unordered_map<uint32_t, size_t> index_of;

  bool is_empty() const { return heap.empty(); }
  size_t len() const { return heap.size(); }
  const Priority& prio_at(size_t pos) const { return map[heap[pos]].second; }
  void heap_swap(size_t a, size_t b) {
    std::swap(heap[a], heap[b]);
    qp[heap[a]] = a;
    qp[heap[b]] = b;
  }
  void bubble_up(size_t pos) {
    while (pos > 0) {
      size_t parent = (pos - 1) / 2;
      if (priority_cmp(prio_at(parent), prio_at(pos)) < 0)
        heap_swap(parent, pos), pos = parent;
      else
        break;
    }
  }
  void heapify(size_t i) {
    size_t n = heap.size();
    if (n <= 1) return;
    for (;;) {
      size_t l = 2 * i + 1, r = 2 * i + 2, largest = i;
      if (l < n && priority_cmp(prio_at(l), prio_at(largest)) > 0) largest = l;
      if (r < n && priority_cmp(prio_at(r), prio_at(largest)) > 0) largest = r;
      if (largest == i) return;
      heap_swap(i, largest);
      i = largest;
    }
  }
  // push: returns true when the item was not present before (schedule.rs:301-302 `opt.is_none()`)
  bool push(AgentImpl item, Priority p) {
    auto it = index_of.find(item.id);
    if (it != index_of.end()) {  // existing item: replace priority and restore heap order
      size_t mi = it->second;
      map[mi].second = p;
      size_t pos = qp[mi];
      bubble_up(pos);
      heapify(qp[mi]);
      return false;
    }
    size_t mi = map.size();
    index_of[item.id] = mi;
    map.emplace_back(std::move(item), p);
    qp.push_back(heap.size());
    heap.push_back(mi);
    bubble_up(heap.size() - 1);
    return true;
  }
  const std::pair<AgentImpl, Priority>* peek() const { return heap.empty() ? nullptr : &map[heap[0]]; }
  // remove the element at heap position `pos` (Store::swap_remove): heap swap_remove, then
  // map swap_remove with index fix-up of the entry that moved into the hole.
  std::pair<AgentImpl, Priority> swap_remove(size_t pos) {
    size_t head = heap[pos];
    heap[pos] = heap.back();
    heap.pop_back();
    if (pos < heap.size()) qp[heap[pos]] = pos;
    size_t last = map.size() - 1;
    std::pair<AgentImpl, Priority> out = std::move(map[head]);
    index_of.erase(out.first.id);
    if (head != last) {
      map[head] = std::move(map[last]);
      qp[head] = qp[last];
      heap[qp[head]] = head;
      index_of[map[head].first.id] = head;
    }
    map.pop_back();
    qp.pop_back();
    return out;
  }
  std::pair<AgentImpl, Priority> pop() {
    if (heap.empty()) rust_panic("Error on pop from queue");
    auto out = swap_remove(0);
    heapify(0);
    return out;
  }
  bool remove(uint32_t id) {
    auto it = index_of.find(id);
    if (it == index_of.end()) return false;
    size_t pos = qp[it->second];
    swap_remove(pos);
    if (pos < heap.size()) {
      bubble_up(pos);
      heapify(pos);
    }
    return true;
  }
};

// Schedule (sequential variant)  schedule.rs:227-413
struct Schedule {
  uint64_t step = 0;
  float time = 0.0f;
  PriorityQueue events;
  uint32_t agent_ids_counting = 0;
  bool quiet = true;  // schedule.rs:358 prints a line when the queue is empty

  // schedule_once :284-286
  void schedule_once(AgentImpl a, float t, int32_t ordering) { events.push(std::move(a), Priority{t, ordering}); }
  // schedule_repeating :295-303
  bool schedule_repeating(std::unique_ptr<Agent> agent, float t, int32_t ordering) {
    AgentImpl a{agent_ids_counting, std::move(agent), true};
    agent_ids_counting += 1;
    return events.push(std::move(a), Priority{t, ordering});
  }
  // distributed_schedule_repeating :305-313
  std::pair<uint32_t, bool> distributed_schedule_repeating(std::unique_ptr<Agent> agent, float t,
                                                           int32_t ordering) {
    bool ok = schedule_repeating(std::move(agent), t, ordering);
    return {agent_ids_counting - 1, ok};
  }
  // get_all_events :316-322 : iteration order of the store
  std::vector<const Agent*> get_all_events() const {
    std::vector<const Agent*> out;
    for (const auto& e : events.map) out.push_back(e.first.agent.get());
    return out;
  }
  // dequeue :329-341
  bool dequeue(uint32_t my_id) { return events.remove(my_id); }

  // step :347-413
  void step_once(State& state) {
    if (step == 0) state.update(step);
    state.before_step(*this);
    if (events.is_empty()) {
      if (!quiet) std::printf("No agent in the queue to schedule. Terminating.\n");
      state.after_step(*this);
      step += 1;
      state.update(step);
      return;
    }
    std::vector<std::pair<AgentImpl, Priority>> cevents;
    time = events.peek()->second.time;
    while (!events.is_empty()) {
      if (events.peek()->second.time > time) break;
      cevents.push_back(events.pop());
    }
    for (auto& item : cevents) {
      item.first.agent->before_step(state);
      item.first.agent->step(state);
      item.first.agent->after_step(state);
      if (item.first.repeating && !item.first.agent->is_stopped(state)) {
        float t = item.second.time + 1.0f;
        int32_t ord = item.second.ordering;
        schedule_once(std::move(item.first), t, ord);
      }
    }
    state.after_step(*this);
    step += 1;
    state.update(step);
  }
};

// ---------------------------------------------------------------- Flockers fixture
// tests/model/flockers/bird.rs + state.rs, with the fixture's constants lifted to parameters
//  and rand::rng() replaced by the Philox stream (test-harness feature).
struct BoidsParams {
  float cohesion = 1.0f, avoidance = 1.0f, randomness = 1.0f, consistency = 1.0f, momentum = 1.0f;
  float jump = 0.7f;      // bird.rs:12-17
  float radius = 10.0f;   // bird.rs:41
  int exact_query = 1;    // bird.rs:41 uses get_neighbors_within_distance; 0 = relax variant
  uint64_t seed = 42;
};

struct Bird {  // bird.rs:19-25
  uint32_t id;
  Real2D pos;
  Real2D last_d;
  bool flag;
};

// Dynamic population on top of the fixture (a model of this repo, like Forest Fire: the reference
// ships no model with births or deaths).  It uses only the reference's own mechanisms:
//   death  Agent::is_stopped (agent.rs:18) -> Schedule::step does not reschedule the agent
//          (schedule.rs:401-407); a dying bird does not push itself into the write buffer
//   birth  State::after_step(schedule) (state.rs / schedule.rs:409) walks the READ buffer in
//          iter_objects order (bags by index, each bag in its stored order), and for every parent
//          that draws a birth pushes a child into the write buffer and schedules it with
//          schedule_repeating (schedule.rs:295-303) for the next step
// Draws: Philox(seed; id, step, DOMAIN_LIFE): v[0] decides death, v[1] birth.
struct LifeRule {
  float death_prob = 0.0f;   // is_stopped when u_death < death_prob ...
  float birth_prob = 0.0f;   // a child when u_birth < birth_prob
  uint32_t crowd_limit = 0;  // ... or (crowd_limit > 0 and neighbours other than self >= crowd_limit)
};

struct Flocker;
struct BirdAgent : Agent {
  Bird b;
  bool stopped = false;
  explicit BirdAgent(Bird bb) : b(bb) {}
  void step(State& state) override;
  bool is_stopped(State&) override { return stopped; }
  std::unique_ptr<Agent> clone() const override { return std::make_unique<BirdAgent>(b); }
};

struct Flocker : State {  // state.rs:16-60
  uint64_t step = 0;
  Field2D<Bird> field1;
  uint32_t initial_flockers;
  float dim0, dim1, discretization;
  bool toroidal;
  BoidsParams params;
  uint64_t current_step = 0;  // mirror of Schedule::step, feeds the Philox counter
  bool canonical_order = false;  // test-harness: sort each read bag by id after the swap
  // when non-null, init() places these agents instead of drawing positions
  const std::vector<Bird>* preset = nullptr;
  bool life_on = false;  // dynamic population (LifeRule above)
  LifeRule life;
  uint32_t next_id = 0;  // id of the next child
  uint64_t born = 0, died = 0;

  Flocker(float w, float h, uint32_t n, float disc, bool tor, BoidsParams p)
      : field1(w, h, disc, tor), initial_flockers(n), dim0(w), dim1(h), discretization(disc),
        toroidal(tor), params(p) {}
  void reset() override {
    step = 0;
    field1 = Field2D<Bird>(dim0, dim1, discretization, toroidal);
  }
  // init  state.rs:41-56
  void init(Schedule& schedule) override {
    if (preset) {
      for (const Bird& b : *preset) {
        field1.set_object_location(b, b.pos);
        schedule.schedule_repeating(std::make_unique<BirdAgent>(b), 0.0f, 0);
      }
      return;
    }
    for (uint32_t bird_id = 0; bird_id < initial_flockers; ++bird_id) {
      Philox4 r = philox4x32_10(bird_id, 0, 0, DOMAIN_INIT, (uint32_t)params.seed,
                                (uint32_t)(params.seed >> 32));
      float r1 = u01_f32(r.v[0]), r2 = u01_f32(r.v[1]);
      Bird bird{bird_id, Real2D{dim0 * r1, dim1 * r2}, Real2D{0.0f, 0.0f}, false};
      field1.set_object_location(bird, bird.pos);
      schedule.schedule_repeating(std::make_unique<BirdAgent>(bird), 0.0f, 0);
    }
  }
  // births: see LifeRule.  The child starts where its parent stood at the beginning of the step.
  void after_step(Schedule& schedule) override {
    if (!life_on) return;
    const auto& bags = field1.bags[field1.read];
    std::vector<Bird> parents;
    for (const auto& bag : bags)
      for (const Bird& b : bag) parents.push_back(b);
    for (const Bird& parent : parents) {
      Philox4 r = philox4x32_10(parent.id, (uint32_t)current_step, (uint32_t)(current_step >> 32), DOMAIN_LIFE,
                                (uint32_t)params.seed, (uint32_t)(params.seed >> 32));
      if (!(u01_f32(r.v[1]) < life.birth_prob)) continue;
      Bird child{next_id++, parent.pos, Real2D{0.0f, 0.0f}, false};
      field1.set_object_location(child, child.pos);
      schedule.schedule_repeating(std::make_unique<BirdAgent>(child), schedule.time + 1.0f, 0);
      born += 1;
    }
  }
  // update  state.rs:58-60
  void update(uint64_t s) override {
    current_step = s;
    field1.lazy_update();
    if (canonical_order)
      for (auto& bag : field1.bags[field1.read])
        std::sort(bag.begin(), bag.end(), [](const Bird& a, const Bird& b) { return a.id < b.id; });
  }
};

// Bird::step  bird.rs:39-155
inline void BirdAgent::step(State& st) {
  Flocker& state = static_cast<Flocker&>(st);
  const BoidsParams& P = state.params;
  std::vector<Bird> vec = P.exact_query
                              ? state.field1.get_neighbors_within_distance(b.pos, P.radius)
                              : state.field1.get_neighbors_within_relax_distance(b.pos, P.radius);
  float width = state.dim0, height = state.dim1;
  Real2D avoidance{0, 0}, cohesion{0, 0}, randomness{0, 0}, consistency{0, 0};
  int32_t count = 0;
  if (!vec.empty()) {
    float x_avoid = 0, y_avoid = 0, x_cohe = 0, y_cohe = 0, x_cons = 0, y_cons = 0;
    for (const Bird& elem : vec) {
      if (b.id != elem.id) {
        float dx = toroidal_distance(b.pos.x, elem.pos.x, width);
        float dy = toroidal_distance(b.pos.y, elem.pos.y, height);
        count += 1;
        float square = dx * dx + dy * dy;
        x_avoid += dx / (square * square + 1.0f);
        y_avoid += dy / (square * square + 1.0f);
        x_cohe += dx;
        y_cohe += dy;
        x_cons += elem.last_d.x;
        y_cons += elem.last_d.y;
      }
    }
    if (count > 0) {
      x_avoid /= (float)count;
      y_avoid /= (float)count;
      x_cohe /= (float)count;
      y_cohe /= (float)count;
      x_cons /= (float)count;
      y_cons /= (float)count;
      consistency = Real2D{x_cons / (float)count, y_cons / (float)count};  // divided twice :88-91
    } else {
      consistency = Real2D{x_cons, y_cons};
    }
    avoidance = Real2D{400.0f * x_avoid, 400.0f * y_avoid};
    cohesion = Real2D{-x_cohe / 10.0f, -y_cohe / 10.0f};
    // randomness :113-123, rand::rng() -> Philox(seed; id, step, DOMAIN_STEP)
    Philox4 r = philox4x32_10(b.id, (uint32_t)state.current_step,
                              (uint32_t)(state.current_step >> 32), DOMAIN_STEP, (uint32_t)P.seed,
                              (uint32_t)(P.seed >> 32));
    float r1 = u01_f32(r.v[0]);
    float x_rand = r1 * 2.0f - 1.0f;
    float r2 = u01_f32(r.v[1]);
    float y_rand = r2 * 2.0f - 1.0f;
    float square = std::sqrt(x_rand * x_rand + y_rand * y_rand);
    randomness = Real2D{0.05f * x_rand / square, 0.05f * y_rand / square};
  }
  Real2D mom = b.last_d;
  float dx = P.cohesion * cohesion.x + P.avoidance * avoidance.x + P.consistency * consistency.x +
             P.randomness * randomness.x + P.momentum * mom.x;
  float dy = P.cohesion * cohesion.y + P.avoidance * avoidance.y + P.consistency * consistency.y +
             P.randomness * randomness.y + P.momentum * mom.y;
  float dis = std::sqrt(dx * dx + dy * dy);
  if (dis > 0.0f) {
    dx = dx / dis * P.jump;
    dy = dy / dis * P.jump;
  }
  b.last_d = Real2D{dx, dy};
  float loc_x = toroidal_transform(b.pos.x + dx, width);
  float loc_y = toroidal_transform(b.pos.y + dy, width);  // `width` for y too  bird.rs:147
  b.pos = Real2D{loc_x, loc_y};
  if (state.life_on) {
    Philox4 r = philox4x32_10(b.id, (uint32_t)state.current_step, (uint32_t)(state.current_step >> 32),
                              DOMAIN_LIFE, (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
    const bool dies = u01_f32(r.v[0]) < state.life.death_prob ||
                      (state.life.crowd_limit != 0 && (uint32_t)count >= state.life.crowd_limit);
    if (dies) {  // is_stopped: not rescheduled, and gone from the field after the swap
      stopped = true;
      state.died += 1;
      return;
    }
  }
  state.field1.set_object_location(b, Real2D{loc_x, loc_y});
}

// simulate! plain branch (lib.rs:1158-1175) / simulate_explore! (model_exploration.rs:160-190)
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

// ---------------------------------------------------------------- Forest Fire (row K)
// The model is not in the reference; the rule is this repo's, written only
// through DenseNumberGrid2D's public API so that the field semantics are the oracle:
// get_value (read buffer) + set_value_location (write buffer) for EVERY live cell, then
// lazy_update (swap; unwritten cells become None).  Moore-8 neighbourhood, non-toroidal.
enum : uint8_t { FF_GREEN = 1, FF_BURNING = 2, FF_BURNED = 3 };

struct ForestFire {
  DenseNumberGrid2D<uint8_t> grid;
  uint64_t steps_done = 0;
  ForestFire(int32_t w, int32_t h) : grid(w, h) {}
  // density-p trees; trees in column x==0 start burning.  Philox(seed; cell, DOMAIN_GRID)
  void init(float density, uint64_t seed) {
    for (int32_t x = 0; x < grid.width; ++x)
      for (int32_t y = 0; y < grid.height; ++y) {
        uint64_t cell = (uint64_t)x * (uint64_t)grid.height + (uint64_t)y;
        Philox4 r = philox4x32_10((uint32_t)cell, (uint32_t)(cell >> 32), 0, DOMAIN_GRID,
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
        if (u01_f32(r.v[0]) < density)
          grid.set_value_location(x == 0 ? FF_BURNING : FF_GREEN, Int2D{x, y});
      }
    grid.lazy_update();
  }
  void step() {
    for (int32_t x = 0; x < grid.width; ++x)
      for (int32_t y = 0; y < grid.height; ++y) {
        std::optional<uint8_t> v = grid.get_value(Int2D{x, y});
        if (!v) continue;
        uint8_t next = *v;
        if (*v == FF_GREEN) {
          bool fire = false;
          for (int32_t dx = -1; dx <= 1 && !fire; ++dx)
            for (int32_t dy = -1; dy <= 1; ++dy) {
              if (dx == 0 && dy == 0) continue;
              int32_t nx = x + dx, ny = y + dy;
              if (nx < 0 || ny < 0 || nx >= grid.width || ny >= grid.height) continue;
              std::optional<uint8_t> nv = grid.get_value(Int2D{nx, ny});
              if (nv && *nv == FF_BURNING) {
                fire = true;
                break;
              }
            }
          if (fire) next = FF_BURNING;
        } else if (*v == FF_BURNING) {
          next = FF_BURNED;
        }
        grid.set_value_location(next, Int2D{x, y});
      }
    grid.lazy_update();
    steps_done += 1;
  }
};

}  // namespace oracle
