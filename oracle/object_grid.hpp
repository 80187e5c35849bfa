// ORACLE — TEST INFRASTRUCTURE ONLY (see krabmaga_oracle.hpp header).
//
// CPU restatement of DenseGrid2D<O> and SparseGrid2D<O>, the object grids of krABMaga 0.6.1
// (src/engine/fields/dense_object_grid_2d.rs:175-779, default variant — not the `parallel` /
// `visualization` one at :17-173; SparseGrid2D: sparse_object_grid_2d.rs:203-721, see below).  SURVEY §8(f) rank 2: the next field to move onto the
// cell-sorted device layout; this restatement and its known-answer tests come first.
//
// PARITY STATUS: PINNED by the reference's own tests (tests/engine/dense_object_grid_2d.rs:31-180,
// tests/engine/sparse_object_grid_2d.rs), ported in tests/test_oracle_object_grid.py.  No device counterpart exists yet.
//
// Quirks kept as they are (SURVEY appendix C policy — do not "fix" silently):
//   * apply_to_all_values hands the closure calculate_indexes_bag(i, width, height)
//     (:768-779) = (i - width*row, row), which is the cell's (x, y) only when read as
//     y-major — on a non-square grid it is not the (i / height, i % height) that every
//     other method uses;
//   * update() (:753-763) INSERTS a clone of every write bag into the read Vec instead of
//     assigning it: afterwards the read Vec holds width*height write-bag copies followed by
//     the old read bags, and apply_to_all_values (which walks 0..rlocs.len()) panics on the
//     first index calculate_indexes_bag cannot map.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <optional>
#include <unordered_map>
#include <vector>

#include "krabmaga_oracle.hpp"

namespace oracle {

// GridOption { READ, WRITE, READWRITE } (grid_option.rs:3-10) comes from krabmaga_oracle.hpp

// dense_object_grid_2d.rs:768-779
inline std::optional<Int2D> calculate_indexes_bag(int32_t index, int32_t width, int32_t height) {
  for (int32_t i = 0; i < height; ++i)
    if (index < width * i + width && index >= width * i) return Int2D{index - width * i, i};
  return std::nullopt;
}

// O needs operator== (the reference's `O: Eq`; the fixture's Bird compares ids, bird.rs:168-172)
template <class O>
struct DenseGrid2D {
  std::vector<std::vector<O>> locs[2];
  size_t read = 0, write = 1;
  int32_t width, height;

  // new  :201-214  (the Vec length uses width*height before the abs())
  DenseGrid2D(int32_t w, int32_t h) : width(w < 0 ? -w : w), height(h < 0 ? -h : h) {
    int64_t n = (int64_t)w * (int64_t)h;
    if (n < 0) rust_panic("DenseGrid2D::new capacity overflow");
    locs[0].resize((size_t)n);
    locs[1].resize((size_t)n);
  }

  size_t index_of(const Int2D& loc, const std::vector<std::vector<O>>& v, const char* who) const {
    int64_t idx = (int64_t)loc.x * height + loc.y;  // ((loc.x * self.height) + loc.y) as usize
    if (idx < 0 || (size_t)idx >= v.size()) rust_panic(std::string("DenseGrid2D::") + who + ": index out of bounds");
    return (size_t)idx;
  }

  using Closure = std::function<std::optional<O>(const Int2D&, const O&)>;
  // apply_to_all_values  :258-328
  void apply_to_all_values(const Closure& closure, GridOption option) {
    auto bag_of = [&](size_t i) {
      auto b = calculate_indexes_bag((int32_t)i, width, height);
      if (!b) rust_panic("error in calculate_indexes_bag");
      return *b;
    };
    std::vector<std::vector<O>>& rlocs = locs[read];
    std::vector<std::vector<O>>& wlocs = locs[write];
    switch (option) {
      case GridOption::READ:  // :263-277
        for (size_t i = 0; i < rlocs.size(); ++i) {
          Int2D bag_id = bag_of(i);
          if (rlocs[i].empty()) continue;
          std::vector<O> vec;
          for (const O& elem : rlocs[i])
            if (auto r = closure(bag_id, elem)) vec.push_back(*r);
          rlocs[i] = std::move(vec);
        }
        break;
      case GridOption::WRITE:  // :278-291
        for (size_t i = 0; i < rlocs.size(); ++i) {
          Int2D bag_id = bag_of(i);
          if (rlocs[i].empty()) continue;
          if (i >= wlocs.size()) rust_panic("DenseGrid2D::apply_to_all_values: write index out of bounds");
          for (const O& elem : rlocs[i])
            if (auto r = closure(bag_id, elem)) wlocs[i].push_back(*r);
        }
        break;
      case GridOption::READWRITE:  // :293-326
        for (size_t i = 0; i < rlocs.size(); ++i) {
          Int2D bag_id = bag_of(i);
          if (i >= wlocs.size()) rust_panic("DenseGrid2D::apply_to_all_values: write index out of bounds");
          if (!wlocs[i].empty()) {
            for (O& elem : wlocs[i])
              if (auto r = closure(bag_id, elem)) elem = *r;
          } else {
            if (rlocs[i].empty()) continue;
            for (const O& elem : rlocs[i])
              if (auto r = closure(bag_id, elem))
                if (std::find(wlocs[i].begin(), wlocs[i].end(), *r) == wlocs[i].end()) wlocs[i].push_back(*r);
          }
        }
        break;
    }
  }

  // get_empty_bags  :358-370
  std::vector<Int2D> get_empty_bags() const {
    std::vector<Int2D> out;
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j)
        if (locs[read][index_of(Int2D{i, j}, locs[read], "get_empty_bags")].empty()) out.push_back(Int2D{i, j});
    return out;
  }

  // get_location :429-441 / get_location_unbuffered :471-482 : first bag (x outer) holding `object`
  std::optional<Int2D> get_location(const O& object, bool unbuffered = false) const {
    const auto& v = locs[unbuffered ? write : read];
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j) {
        const auto& bag = v[index_of(Int2D{i, j}, v, "get_location")];
        if (std::find(bag.begin(), bag.end(), object) != bag.end()) return Int2D{i, j};
      }
    return std::nullopt;
  }

  // get_objects :507-520 / get_objects_unbuffered :547-561 : None for an empty bag
  std::optional<std::vector<O>> get_objects(const Int2D& loc, bool unbuffered = false) const {
    const auto& v = locs[unbuffered ? write : read];
    const auto& bag = v[index_of(loc, v, "get_objects")];
    if (bag.empty()) return std::nullopt;
    return bag;
  }

  // iter_objects :589-608 / iter_objects_unbuffered :634-654 : x outer, y inner, bag order
  template <class F>
  void iter_objects(F&& closure, bool unbuffered = false) const {
    const auto& v = locs[unbuffered ? write : read];
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j)
        for (const O& obj : v[index_of(Int2D{i, j}, v, "iter_objects")]) closure(Int2D{i, j}, obj);
  }

  // set_object_location  :688-697 : an equal object already in that write bag is replaced (moved last)
  void set_object_location(const O& object, const Int2D& loc) {
    auto& bag = locs[write][index_of(loc, locs[write], "set_object_location")];
    if (!bag.empty()) bag.erase(std::remove(bag.begin(), bag.end(), object), bag.end());
    bag.push_back(object);
  }

  // remove_object_location  :729-736
  void remove_object_location(const O& object, const Int2D& loc) {
    auto& bag = locs[write][index_of(loc, locs[write], "remove_object_location")];
    if (!bag.empty()) bag.erase(std::remove(bag.begin(), bag.end(), object), bag.end());
  }

  // Field::lazy_update  :743-750
  void lazy_update() {
    std::swap(read, write);
    for (auto& bag : locs[write]) bag.clear();
  }

  // Field::update  :753-763  (Vec::insert, see the header note)
  void update() {
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j) {
        size_t index = index_of(Int2D{i, j}, locs[write], "update");
        if (index > locs[read].size()) rust_panic("DenseGrid2D::update: insertion index out of bounds");
        locs[read].insert(locs[read].begin() + (ptrdiff_t)index, locs[write][index]);
      }
  }
};

// ---------------------------------------------------------------- SparseGrid2D
// src/engine/fields/sparse_object_grid_2d.rs:203-721, default variant: two HashMap<Int2D, Vec<O>>.
// Differences from DenseGrid2D that the restatement keeps: no bounds (any Int2D is a key),
// set_object_location pushes without replacing an equal object (:648-659), an emptied bag loses
// its key (:690-699), apply_to_all_values panics when the closure returns None (`.expect`,
// :278-320) and its READWRITE arm keeps ONE object per new write bag (each insert overwrites the
// previous one), update() is a real copy (:711-718), iteration order is the HashMap's
// (unspecified: compare as sets).  get_random_empty_bag (:523-535) draws cells until one has no
// key and never returns when every cell has one; it is not restated.
struct Int2DHash {
  size_t operator()(const Int2D& k) const { return ((uint64_t)(uint32_t)k.x << 32) ^ (uint32_t)k.y; }
};
inline bool operator==(const Int2D& a, const Int2D& b) { return a.x == b.x && a.y == b.y; }

template <class O>
struct SparseGrid2D {
  using Map = std::unordered_map<Int2D, std::vector<O>, Int2DHash>;
  Map locs[2];
  size_t read = 0, write = 1;
  int32_t width, height;

  SparseGrid2D(int32_t w, int32_t h) : width(w), height(h) {}  // :224-234 (no abs() here)

  using Closure = std::function<std::optional<O>(const Int2D&, const O&)>;
  // apply_to_all_values  :278-320
  void apply_to_all_values(const Closure& closure, GridOption option) {
    auto must = [&](const Int2D& key, const O& obj) {
      auto r = closure(key, obj);
      if (!r) rust_panic("error on closure");
      return *r;
    };
    switch (option) {
      case GridOption::READ:
        for (auto& kv : locs[read])
          for (O& obj : kv.second) obj = must(kv.first, obj);
        break;
      case GridOption::WRITE:
        for (auto& kv : locs[write])
          for (O& obj : kv.second) obj = must(kv.first, obj);
        break;
      case GridOption::READWRITE:
        for (const auto& kv : locs[read]) {
          auto w = locs[write].find(kv.first);
          if (w != locs[write].end()) {
            for (O& obj : w->second) obj = must(kv.first, obj);
          } else {
            // HashMap::insert replaces: after the loop the new bag holds the LAST object only
            for (const O& obj : kv.second) locs[write][kv.first] = std::vector<O>{must(kv.first, obj)};
          }
        }
        break;
    }
  }

  // get_location :346-356 / get_location_unbuffered :386-396 (first hit in map order)
  std::optional<Int2D> get_location(const O& object, bool unbuffered = false) const {
    for (const auto& kv : locs[unbuffered ? write : read])
      for (const O& obj : kv.second)
        if (obj == object) return kv.first;
    return std::nullopt;
  }
  // get_objects :421-423 / get_objects_unbuffered :450-452
  std::optional<std::vector<O>> get_objects(const Int2D& loc, bool unbuffered = false) const {
    const Map& m = locs[unbuffered ? write : read];
    auto it = m.find(loc);
    if (it == m.end()) return std::nullopt;
    return it->second;
  }
  // get_empty_bags  :482-499 : cells of [0,width) x [0,height) without a key or with an empty bag
  std::vector<Int2D> get_empty_bags() const {
    std::vector<Int2D> out;
    for (int32_t i = 0; i < width; ++i)
      for (int32_t j = 0; j < height; ++j) {
        auto it = locs[read].find(Int2D{i, j});
        if (it == locs[read].end() || it->second.empty()) out.push_back(Int2D{i, j});
      }
    return out;
  }
  // iter_objects :561-574 / iter_objects_unbuffered :601-615
  template <class F>
  void iter_objects(F&& closure, bool unbuffered = false) const {
    for (const auto& kv : locs[unbuffered ? write : read])
      for (const O& obj : kv.second) closure(kv.first, obj);
  }
  // set_object_location  :648-659
  void set_object_location(const O& object, const Int2D& loc) { locs[write][loc].push_back(object); }
  // remove_object_location  :690-699
  void remove_object_location(const O& object, const Int2D& loc) {
    auto it = locs[write].find(loc);
    if (it == locs[write].end()) return;
    auto& bag = it->second;
    bag.erase(std::remove(bag.begin(), bag.end(), object), bag.end());
    if (bag.empty()) locs[write].erase(it);
  }
  // Field::lazy_update  :705-708
  void lazy_update() {
    std::swap(read, write);
    locs[write].clear();
  }
  // Field::update  :711-718
  void update() {
    locs[read] = locs[write];
    locs[write].clear();
  }
};

}  // namespace oracle
