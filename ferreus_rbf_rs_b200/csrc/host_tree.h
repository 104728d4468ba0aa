// Host-side linear Morton tree and interaction lists built over device-sorted point codes.
// Reference semantics: ferreus_bbfmm/src/linear_tree.rs:20-485, morton.rs:29-373.
#pragma once
#include <cstdint>
#include <vector>

namespace fb {

struct HostTree {
  int dim = 3;
  int depth = 0;
  bool adaptive = true;
  bool sparse = true;
  double center[3] = {0, 0, 0};
  double radius = 0;
  double disp[3] = {0, 0, 0};  // center - radius

  // cells, level-major, Morton order inside a level; cell 0 is the root
  std::vector<uint64_t> prefix;   // interleaved anchor bits at the cell's level
  std::vector<int32_t> level;
  std::vector<int32_t> parent;    // -1 for root
  std::vector<int32_t> child_ptr; // ncells+1 into child_idx (children sorted by Morton suffix)
  std::vector<int32_t> child_idx;
  std::vector<uint8_t> is_leaf;
  std::vector<int32_t> pt_begin, pt_end;  // range in Morton-sorted point order (all cells)
  std::vector<int32_t> level_ptr;         // depth+2 entries: cells of level L are [level_ptr[L], level_ptr[L+1])
  std::vector<int32_t> leaves;            // cell ids of leaves in Morton (depth-first) order

  // interaction lists as CSR over cell ids (entries sorted ascending)
  std::vector<int64_t> u_ptr, v_ptr, w_ptr, x_ptr;
  std::vector<int32_t> u_idx, v_idx, w_idx, x_idx;

  size_t ncells() const { return prefix.size(); }
  uint64_t ref_key(int c) const { return (prefix[c] << 15) | (uint64_t)level[c]; }  // morton.rs:58-119
  int find(int lvl, uint64_t pfx) const;  // cell id or -1
  void anchor(int c, uint32_t a[3]) const;
  void cell_center(int c, double out[3], double &side) const;  // morton.rs:328-346
  bool adjacent(int a, int b) const;                           // morton.rs:308-325

  // codes: Morton-sorted level-16 interleaved codes of the n source points
  void build(const uint64_t *codes, size_t n, int dim_, const double center_[3], double radius_,
             size_t max_points_per_cell, bool store_empty, bool adaptive_);
  void build_lists_adaptive();  // linear_tree.rs:177-395
  void build_lists_regular();   // linear_tree.rs:397-485
};

}  // namespace fb
