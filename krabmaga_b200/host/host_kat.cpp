// The reference's own known-answer tests for the path (tests/engine/field_2d.rs,
// tests/engine/dense_number_grid_2d.rs, tests/explore/simulate.rs) restated over the C++ host
// mirror; linked against libkrabgpu.so only.  Run by tests/test_gpu_host_cpp.py on the GPU box.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "krabmaga_gpu.hpp"

using namespace krabmaga::gpu;

static int failures = 0;
#define EXPECT(cond)                                                        \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);          \
      ++failures;                                                           \
    }                                                                       \
  } while (0)

static const float WIDTH = 10.f, HEIGHT = 10.f, DISCRETIZATION = 0.5f;  // state.rs:11-14
static KgBoidsParams fixture_params() { return KgBoidsParams{1, 1, 1, 1, 1, 0.7f, 10.f, 1, 42, 0}; }
static bool contains(const std::vector<uint32_t>& v, uint32_t id) { return std::find(v.begin(), v.end(), id) != v.end(); }

static void field_2d_single_step() {  // tests/engine/field_2d.rs:31-50
  Flocker state(WIDTH, HEIGHT, 10, DISCRETIZATION, true, fixture_params());
  Schedule schedule;
  state.init(schedule);
  schedule.step_once(state);
  EXPECT(state.field1->nagents() == 10);
  EXPECT(state.field1->get_neighbors_within_distance({5, 5}, 10).size() == 10);
  EXPECT(state.field1->get_neighbors_within_relax_distance({5, 5}, 10).size() == 10);
}
static void field_2d_neighbors() {  // :58-117
  for (float fly = 5; fly < 10; fly += 1) {
    Field2D f(WIDTH, HEIGHT, DISCRETIZATION, true, 16);
    Bird b1{1, {0, 0}, {0, 0}}, b2{2, {0, 0}, {0, 0}};
    f.set_object_location(b1, b1.pos);
    f.set_object_location(b2, b2.pos);
    f.lazy_update();
    EXPECT(f.nagents() == 2);
    EXPECT(f.get_neighbors_within_distance({5, 5}, 1).empty());
    EXPECT(f.get_neighbors_within_relax_distance({5, 5}, 1).empty());
    b1.pos = {b1.pos.x + fly, b1.pos.y + fly};
    f.set_object_location(b1, b1.pos);
    f.set_object_location(b2, b2.pos);
    f.lazy_update();
    auto v = f.get_neighbors_within_distance(b1.pos, 1);
    EXPECT(v.size() == 1 && contains(v, 1));
    v = f.get_neighbors_within_distance(b2.pos, 1);
    EXPECT(v.size() == 1 && contains(v, 2));
    v = f.get_neighbors_within_distance({5, 5}, 10);
    EXPECT(v.size() == 2 && contains(v, 1) && contains(v, 2));
    v = f.get_neighbors_within_relax_distance({5, 5}, 10);
    EXPECT(v.size() == 2 && contains(v, 1) && contains(v, 2));
  }
}
static void field_2d_gets() {  // :125-178
  Field2D f(WIDTH, HEIGHT, DISCRETIZATION, true, 16);
  Bird b1{1, {0, 0}, {0, 0}}, b2{2, {5, 5}, {0, 0}}, b3{3, {5, 5}, {0, 0}};
  f.set_object_location(b1, b1.pos);
  f.set_object_location(b2, b2.pos);
  f.set_object_location(b3, b3.pos);
  f.lazy_update();
  EXPECT(f.nagents() == 3);
  auto birds = f.get_objects({5, 5});
  EXPECT(birds.size() == 2 && contains(birds, 2) && contains(birds, 3));
  EXPECT(f.get_objects({10, 0}).empty());
  EXPECT(f.num_objects_at_location({5, 5}) == 2);
  EXPECT(f.num_objects_at_location({0, 0}) == 1);
  Bird b4{4, {0, 0}, {0, 0}};
  f.set_object_location(b4, b4.pos);
  EXPECT(f.get_objects_unbuffered({0, 0}).size() == 1);
  EXPECT(f.get_objects({0, 0}).size() == 1);
  f.remove_object_location(b4, b4.pos);
  EXPECT(f.get_objects_unbuffered({0, 0}).empty());
}
static void field_2d_bags() {  // :186-210
  Field2D f(10, 10, DISCRETIZATION, true, 16);
  EXPECT(f.num_empty_bags() == 441);
  f.set_object_location(Bird{1, {0, 0}, {0, 0}}, {0, 0});
  f.set_object_location(Bird{2, {0, 0}, {0, 0}}, {0, 0});
  f.set_object_location(Bird{3, {4, 4}, {0, 0}}, {4, 4});
  f.lazy_update();
  EXPECT(f.num_empty_bags() == 439);
}
static void field_2d_panics() {  // field_2d.rs:840-842: index out of bounds
  Field2D f(10, 10, DISCRETIZATION, true, 16);
  bool threw = false;
  try {
    f.set_object_location(Bird{1, {-1, 0}, {0, 0}}, {-1, 0});
  } catch (const Panic& p) {
    threw = p.code == KG_E_OOB;
  }
  EXPECT(threw);
}
static void dense_number_grid_2d_bags() {  // tests/engine/dense_number_grid_2d.rs:97-165 (T = u8)
  DenseNumberGrid2D g(10, 10);
  EXPECT(g.num_empty_bags() == 100);
  Int2D loc{4, 2};
  g.set_value_location(10, loc);
  EXPECT(g.get_value_unbuffered(loc) == std::optional<uint8_t>(10));
  g.remove_value_location(loc);
  EXPECT(!g.get_value_unbuffered(loc));
  g.set_value_location(10, loc);
  g.update();
  EXPECT(g.num_empty_bags() == 99);
  for (int i = 0; i < 10; ++i)
    for (int j = 0; j < 10; ++j) g.set_value_location(0, {i, j});
  g.lazy_update();
  EXPECT(g.num_empty_bags() == 0);
  EXPECT(g.get_value({3, 3}) == std::optional<uint8_t>(0));
  g.lazy_update();
  EXPECT(g.num_empty_bags() == 100);  // nothing written: every cell is None again (:541-543)
}
static void simulate_kat() {  // tests/explore/simulate.rs:18-38 geometry
  Flocker state(200.f, 200.f, 100, DISCRETIZATION, true, fixture_params());
  simulate(state, 10, 1);
  EXPECT(state.step == 10);
  auto birds = state.field1->objects();
  EXPECT(birds.size() == 100);
  for (const Bird& b : birds) {
    EXPECT(b.pos.x >= 0 && b.pos.x <= 200.f && b.pos.y >= 0 && b.pos.y <= 200.f);
    float n = std::sqrt(b.last_d.x * b.last_d.x + b.last_d.y * b.last_d.y);
    EXPECT(n == 0.f || std::fabs(n - 0.7f) < 1e-5f);
  }
}

static void explore_kat() {  // explore_parallel!: rows in run order; a batch replica == the run alone
  const float w = 120.f, disc = 10.f / 1.5f;
  const uint32_t n = 800;
  std::vector<KgBoidsParams> confs;
  for (int i = 0; i < 3; ++i) {
    KgBoidsParams p{1.f + 0.5f * i, 1.f, 1.f, 1.f, 1.f, 0.7f, 10.f, 0, 42, 0};
    confs.push_back(p);
  }
  auto rows = explore_parallel(12, 2, w, w, disc, n, confs, 0, true);
  EXPECT(rows.size() == 6);
  for (size_t k = 0; k < rows.size(); ++k) {
    EXPECT(rows[k].conf_num == k / 2 && rows[k].conf_rep == k % 2);
    EXPECT(rows[k].run_duration > 0 && rows[k].step_per_sec > 0);
    EXPECT(rows[k].output >= 0.f && rows[k].output <= 1.0001f);
  }
  // write_csv: header = DataFrame::field_names, one record per run
  write_csv("/tmp/krabgpu_host_kat_explore", rows);
  {
    std::FILE* fh = std::fopen("/tmp/krabgpu_host_kat_explore.csv", "r");
    EXPECT(fh != nullptr);
    int lines = 0, commas_first = 0;
    for (int c, first = 1; fh && (c = std::fgetc(fh)) != EOF;) {
      if (c == '\n') ++lines, first = 0;
      else if (c == ',' && first) ++commas_first;
    }
    if (fh) std::fclose(fh);
    EXPECT(lines == 7 && commas_first + 1 == (int)FrameRow::field_names().size());
    std::remove("/tmp/krabgpu_host_kat_explore.csv");
  }
  // replica 3 (configuration 1, repetition 1) against the same model run on its own field
  Flocker alone(w, w, n, disc, true, rows[3].input);
  alone.canonical_order = true;
  simulate(alone, 12, 1);
  auto a = alone.field1->objects();
  FlockerBatch batch(w, w, disc, true, {rows[3].input}, n, 0, true);
  batch.simulate(12);
  auto b = batch.objects();
  EXPECT(a.size() == b.size());
  bool same = a.size() == b.size();
  for (size_t i = 0; same && i < a.size(); ++i)
    same = a[i].id == b[i].id && a[i].pos.x == b[i].pos.x && a[i].pos.y == b[i].pos.y &&
           a[i].last_d.x == b[i].last_d.x && a[i].last_d.y == b[i].last_d.y;
  EXPECT(same);
}

int main() {
  try {
    field_2d_single_step();
    field_2d_neighbors();
    field_2d_gets();
    field_2d_bags();
    field_2d_panics();
    dense_number_grid_2d_bags();
    simulate_kat();
    explore_kat();
  } catch (const Panic& p) {
    std::printf("PANIC %d: %s\n", p.code, p.what());
    return 2;
  }
  if (failures)
    std::printf("host_kat: %d failure(s)\n", failures);
  else
    std::printf("host_kat: all passed (launches %llu)\n", (unsigned long long)kg_launch_count());
  return failures ? 1 : 0;
}
