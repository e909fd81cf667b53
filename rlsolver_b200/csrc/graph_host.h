// Host-side graph object behind rlsb_graph_t (private to csrc/).
#pragma once
#include <string>
#include <vector>

#include "common.cuh"

struct rlsb_graph {
  int32_t n = 0, np = 0, bidir = 0, device = -1, levels = 0, max_listed_deg = 0, max_full_deg = 0;
  int64_t m = 0;
  std::vector<int32_t> edge_u, edge_v, weight;
  std::vector<int32_t> listed_ptr, listed_col, listed_row, listed_deg, full_ptr, full_col, level_ptr, level_nodes;
  // tile-kernel structures (only when np <= kMaxTileNodes)
  bool tileable = false;
  std::vector<uint32_t> edge_pair;
  std::vector<int32_t> level_slice;
  struct Sell {
    std::vector<int32_t> off;
    std::vector<uint16_t> node, half, col;
  } sell_listed, sell_sweep;
  // weighted objective (weights other than 1 present): edges bucketed by (bit of |w|, sign), each bucket a
  // zero-padded list of u | v << 16 words; meta[k] = {first quad, quads, signed scale (+-2^bit), edges}
  bool weighted = false;
  std::vector<uint32_t> wpair;
  std::vector<int32_t> wmeta;
  std::vector<int32_t> full_w;          // weight of every full-neighbourhood slot (aligned with full_col)
  std::vector<int32_t> wdeg;            // [np] sum of the weights of a node's (non-loop) edges
  int64_t weight_sum = 0, abs_weight_sum = 0;
  void* dev_blob = nullptr;   // one allocation holding every device array
  rlsb::GraphDev dev{};
  // side stream + fork / join events for work that runs next to the caller's stream (the streaming mask
  // generator of rlsb_ls_fused_search); created with the device blob
  cudaStream_t side_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};


namespace rlsb {
void build_csr(int32_t n, const std::vector<int32_t>& src, const std::vector<int32_t>& dst, std::vector<int32_t>& ptr,
               std::vector<int32_t>& col);
void build_sell(const std::vector<int32_t>& order, const std::vector<int32_t>& group_ptr,
                const std::vector<int32_t>& ptr, const std::vector<int32_t>& col, rlsb_graph::Sell& out,
                std::vector<int32_t>* group_slice, int32_t pad_id = -1);
}  // namespace rlsb
