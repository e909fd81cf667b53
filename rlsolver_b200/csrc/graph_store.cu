// Graph store: native CSR / edge-list / sweep-level builder and its device image.
// Replaces EnvMaxcut.__init__ (rlsolver/envs/env_L2A.py:25-52), build_adjacency_indies
// (rlsolver/methods/util_read_data.py:144-187) and calc_num_nodes_in_mygraph
// (rlsolver/methods/util.py:35-40).  Host side is plain C++; the device image is a
// handful of int32 arrays that stay L2-resident (<= a few hundred KB for Gset).
#include <atomic>
#include <stdlib.h>
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "graph_host.h"
#include "ls_workspace.cuh"

namespace rlsb {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

}  // namespace rlsb

namespace rlsb {
const GraphDev* graph_dev(const rlsb_graph_t* g) { return (g && g->dev_blob && g->tileable) ? &g->dev : nullptr; }
int graph_device_id(const rlsb_graph_t* g) { return g ? g->device : -1; }

int graph_check(const rlsb_graph_t* g, const GraphDev** out, const char* what) {
  *out = nullptr;
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_INVALID, "%s: null graph", what);
  RLSB_REQUIRE(g->device >= 0, RLSB_ERR_NODEVICE, "%s: graph was built host-only (device < 0)", what);
  RLSB_REQUIRE(g->tileable, RLSB_ERR_UNSUPPORTED,
               "%s: graph outside the shared-memory tile kernels (padded nodes %d > %d or degree > 4095)", what, g->np,
               kMaxTileNodes);
  *out = &g->dev;
  return RLSB_OK;
}

int graph_side(const rlsb_graph_t* g, GraphSide* out) {
  RLSB_REQUIRE(g && g->side_stream && g->ev_fork && g->ev_join, RLSB_ERR_NODEVICE, "graph has no side stream (host-only graph)");
  out->stream = g->side_stream, out->fork = g->ev_fork, out->join = g->ev_join;
  return RLSB_OK;
}

int cut_warps_for(int64_t m, int max_warps) {
  int64_t w = (m + 1023) / 1024;          // >= 32 edges per lane before another warp joins
  if (w < 1) w = 1;
  return int(w < max_warps ? w : max_warps);
}
}  // namespace rlsb

namespace rlsb {

// counting-sort CSR: rows = src, columns sorted ascending inside each row
void build_csr(int32_t n, const std::vector<int32_t>& src, const std::vector<int32_t>& dst,
               std::vector<int32_t>& ptr, std::vector<int32_t>& col) {
  ptr.assign(n + 1, 0);
  for (int32_t s : src) ptr[s + 1]++;
  for (int32_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
  col.resize(src.size());
  std::vector<int32_t> fill(ptr.begin(), ptr.end() - 1);
  for (size_t k = 0; k < src.size(); ++k) col[fill[src[k]]++] = dst[k];
  for (int32_t i = 0; i < n; ++i) std::sort(col.begin() + ptr[i], col.begin() + ptr[i + 1]);
}

// SELL-32 over an ordered list of nodes split into groups (slices never straddle a group).
// A slice is 32 node slots; its column ids are stored in blocks of 4 rounds, lane-major inside
// a block (col[(block*32 + lane)*4 + r]), so a lane fetches 4 neighbour ids with one 8-byte load
// and a warp reads 256 contiguous bytes.  Rows shorter than the slice are padded with the node's
// own id (word ^ word == 0: padding never counts) or with pad_id when pad_id >= 0; unused slots
// are 0xFFFF and their columns 0.
// `off` counts blocks.
void build_sell(const std::vector<int32_t>& order, const std::vector<int32_t>& group_ptr,
                const std::vector<int32_t>& ptr, const std::vector<int32_t>& col, rlsb_graph::Sell& out,
                std::vector<int32_t>* group_slice, int32_t pad_id) {
  out.off.assign(1, 0);
  out.node.clear(), out.half.clear(), out.col.clear();
  if (group_slice) group_slice->assign(1, 0);
  for (size_t gi = 0; gi + 1 < group_ptr.size(); ++gi) {
    for (int32_t b = group_ptr[gi]; b < group_ptr[gi + 1]; b += 32) {
      const int32_t cnt = std::min(32, group_ptr[gi + 1] - b);
      int32_t width = 0;
      for (int32_t l = 0; l < cnt; ++l) width = std::max(width, ptr[order[b + l] + 1] - ptr[order[b + l]]);
      const int32_t blocks = (width + 3) / 4;
      const size_t base = out.col.size();
      out.col.resize(base + size_t(blocks) * 128, 0);
      for (int32_t l = 0; l < 32; ++l) {
        if (l >= cnt) {
          out.node.push_back(0xFFFF), out.half.push_back(0);
          continue;
        }
        const int32_t i = order[b + l], deg = ptr[i + 1] - ptr[i];
        out.node.push_back(uint16_t(i)), out.half.push_back(uint16_t(deg / 2));
        for (int32_t k = 0; k < blocks * 4; ++k)
          out.col[base + (size_t(k / 4) * 32 + l) * 4 + (k % 4)] = uint16_t(k < deg ? col[ptr[i] + k] : (pad_id < 0 ? i : pad_id));
      }
      out.off.push_back(out.off.back() + blocks);
    }
    if (group_slice) group_slice->push_back(int32_t(out.off.size()) - 1);
  }
}

static int debug_flags_from_env() {
  int f = 0;
  const struct { const char* name; int bit; } vars[] = {{"RLSB_LS_PLAIN_MASKS", RLSB_DEBUG_PLAIN_MASKS},
                                                        {"RLSB_LS_FULL_CUT", RLSB_DEBUG_FULL_CUT},
                                                        {"RLSB_LS_SKIP", RLSB_DEBUG_LS_SKIP},
                                                        {"RLSB_LS_TIMES", RLSB_DEBUG_LS_TIMES},
                                                        {"RLSB_CARVEOUT_DEFAULT", RLSB_DEBUG_CARVEOUT_DEFAULT},
                                                        {"RLSB_GEN_PER_DRAW", RLSB_DEBUG_GEN_PER_DRAW},
                                                        {"RLSB_PECO_WARP_PER_ENV", RLSB_DEBUG_PECO_WARP_PER_ENV},
                                                        {"RLSB_QUBO_NO_SPLITK", RLSB_DEBUG_QUBO_NO_SPLITK},
                                                        {"RLSB_LS_THRESH_PIPE", RLSB_DEBUG_THRESH_PIPE}};
  for (const auto& v : vars) {
    const char* e = getenv(v.name);
    if (e && e[0] == '1') f |= v.bit;
  }
  return f;
}
static std::atomic<int> g_debug_flags{debug_flags_from_env()};
int debug_flags() { return g_debug_flags.load(std::memory_order_relaxed); }

}  // namespace rlsb

using rlsb::build_csr;
using rlsb::build_sell;

extern "C" {

int rlsb_version(void) { return 100; }
const char* rlsb_last_error(void) { return rlsb::g_last_error.c_str(); }
int32_t rlsb_debug_flags(int32_t set_mask, int32_t clear_mask) {
  int old = rlsb::g_debug_flags.load(), now;
  do now = (old & ~clear_mask) | set_mask;
  while (!rlsb::g_debug_flags.compare_exchange_weak(old, now));
  return now;
}

int rlsb_graph_create(int32_t num_nodes, int64_t num_edges, const int32_t* h_n0, const int32_t* h_n1,
                      const int32_t* h_w, int32_t bidirectional, int32_t device, rlsb_graph_t** out) {
  using namespace rlsb;
  RLSB_REQUIRE(out != nullptr, RLSB_ERR_INVALID, "graph_create: out is null");
  *out = nullptr;
  RLSB_REQUIRE(num_edges >= 0 && num_edges < (int64_t(1) << 30), RLSB_ERR_INVALID,
               "graph_create: num_edges %lld out of range", (long long)num_edges);
  RLSB_REQUIRE(num_edges == 0 || (h_n0 && h_n1), RLSB_ERR_INVALID, "graph_create: null endpoint arrays");
  int32_t max_id = -1;
  for (int64_t k = 0; k < num_edges; ++k) {
    RLSB_REQUIRE(h_n0[k] >= 0 && h_n1[k] >= 0, RLSB_ERR_INVALID, "graph_create: negative node id at edge %lld",
                 (long long)k);
    max_id = std::max(max_id, std::max(h_n0[k], h_n1[k]));
  }
  if (num_nodes <= 0) {  // reference quirk: N = number of distinct endpoints
    std::vector<uint8_t> seen(size_t(max_id) + 1, 0);
    int32_t distinct = 0;
    for (int64_t k = 0; k < num_edges; ++k) {
      if (!seen[h_n0[k]]) seen[h_n0[k]] = 1, ++distinct;
      if (!seen[h_n1[k]]) seen[h_n1[k]] = 1, ++distinct;
    }
    num_nodes = distinct;
  }
  RLSB_REQUIRE(max_id < num_nodes, RLSB_ERR_INVALID,
               "graph_create: node id %d >= num_nodes %d (the reference raises IndexError here: isolated "
               "nodes shrink N, rlsolver/methods/util.py:35-40)", max_id, num_nodes);

  auto* g = new rlsb_graph();
  g->n = num_nodes;
  g->np = (num_nodes + 31) / 32 * 32;
  g->m = num_edges;
  g->bidir = bidirectional ? 1 : 0;
  g->device = device;
  g->edge_u.assign(h_n0, h_n0 + num_edges);
  g->edge_v.assign(h_n1, h_n1 + num_edges);
  if (h_w) g->weight.assign(h_w, h_w + num_edges);

  // listed neighbours: forward, plus reverse when bidirectional (self loops and duplicates kept)
  std::vector<int32_t> src(g->edge_u), dst(g->edge_v);
  if (g->bidir) {
    src.insert(src.end(), g->edge_v.begin(), g->edge_v.end());
    dst.insert(dst.end(), g->edge_u.begin(), g->edge_u.end());
  }
  build_csr(g->n, src, dst, g->listed_ptr, g->listed_col);
  g->listed_row.resize(g->listed_col.size());
  for (int32_t i = 0; i < g->n; ++i)
    std::fill(g->listed_row.begin() + g->listed_ptr[i], g->listed_row.begin() + g->listed_ptr[i + 1], i);

  // full undirected neighbourhood with multiplicity, self loops dropped (they never cut)
  src.clear(), dst.clear();
  for (int64_t k = 0; k < num_edges; ++k) {
    if (g->edge_u[k] == g->edge_v[k]) continue;
    src.push_back(g->edge_u[k]), dst.push_back(g->edge_v[k]);
    src.push_back(g->edge_v[k]), dst.push_back(g->edge_u[k]);
  }
  build_csr(g->n, src, dst, g->full_ptr, g->full_col);

  // weights: only materialised when some weight differs from 1 (the reference's EnvMaxcut ignores them; the
  // weighted entry points serve the callers that do not: PISCO's adjacency energy, MCPG's weighted sampler)
  for (int64_t k = 0; k < num_edges && h_w; ++k) g->weighted = g->weighted || h_w[k] != 1;
  if (g->weighted) {
    int32_t max_abs = 0;
    for (int64_t k = 0; k < num_edges; ++k) {
      const int64_t w = h_w[k];
      RLSB_REQUIRE_CLEAN(w > -(int64_t(1) << 24) && w < (int64_t(1) << 24), delete g, RLSB_ERR_UNSUPPORTED,
                         "graph_create: |weight| of edge %lld is 2^24 or more", (long long)k);
      if (g->edge_u[k] != g->edge_v[k]) g->weight_sum += w, g->abs_weight_sum += w < 0 ? -w : w;
      max_abs = std::max<int32_t>(max_abs, int32_t(w < 0 ? -w : w));
    }
    // full_w aligned with full_col: build_csr orders a row by neighbour id, stably; replay that order
    {
      std::vector<std::vector<std::pair<int32_t, int32_t>>> rows(g->n);
      for (int64_t k = 0; k < num_edges; ++k) {
        if (g->edge_u[k] == g->edge_v[k]) continue;
        rows[g->edge_u[k]].push_back({g->edge_v[k], h_w[k]});
        rows[g->edge_v[k]].push_back({g->edge_u[k], h_w[k]});
      }
      g->full_w.assign(g->full_col.size(), 0);
      g->wdeg.assign(g->np, 0);
      for (int32_t i = 0; i < g->n; ++i) {
        auto& r = rows[i];
        std::stable_sort(r.begin(), r.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (size_t t = 0; t < r.size(); ++t) {
          // duplicates of one neighbour may carry different weights: any assignment of them to the equal
          // column entries gives the same sums
          g->full_w[g->full_ptr[i] + t] = r[t].second;
          g->wdeg[i] += r[t].second;
        }
      }
    }
    for (int bit = 0; (1 << bit) <= max_abs; ++bit)
      for (int sign = 0; sign < 2; ++sign) {
        std::vector<uint32_t> list;
        for (int64_t k = 0; k < num_edges; ++k) {
          const int32_t w = h_w[k], a = w < 0 ? -w : w;
          if (((a >> bit) & 1) && (w < 0) == (sign == 1) && g->edge_u[k] != g->edge_v[k])
            list.push_back(uint32_t(g->edge_u[k]) | (uint32_t(g->edge_v[k]) << 16));
        }
        if (list.empty()) continue;
        const int32_t first = int32_t(g->wpair.size() / 4), edges = int32_t(list.size());
        list.resize((list.size() + 3) / 4 * 4, 0u);
        g->wpair.insert(g->wpair.end(), list.begin(), list.end());
        const int32_t scale = sign ? -(1 << bit) : (1 << bit);
        g->wmeta.insert(g->wmeta.end(), {first, int32_t(list.size() / 4), scale, edges});
      }
  }

  for (int32_t i = 0; i < g->n; ++i) {
    g->max_listed_deg = std::max(g->max_listed_deg, g->listed_ptr[i + 1] - g->listed_ptr[i]);
    g->max_full_deg = std::max(g->max_full_deg, g->full_ptr[i + 1] - g->full_ptr[i]);
  }

  // Dependency levels of the in-order Gauss-Seidel sweep: node i must see the decisions of
  // every neighbour j < i, so level(i) = 1 + max level(j<i); nodes of one level are pairwise
  // non-adjacent and can be decided concurrently with the exact sequential result.
  std::vector<int32_t> level(g->n, 0);
  int32_t nlev = g->n ? 1 : 0;
  for (int32_t i = 0; i < g->n; ++i) {
    int32_t lv = 0;
    for (int32_t k = g->full_ptr[i]; k < g->full_ptr[i + 1]; ++k) {
      const int32_t j = g->full_col[k];
      if (j < i) lv = std::max(lv, level[j] + 1);
    }
    level[i] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  g->levels = nlev;
  g->level_ptr.assign(nlev + 1, 0);
  for (int32_t i = 0; i < g->n; ++i) g->level_ptr[level[i] + 1]++;
  for (int32_t l = 0; l < nlev; ++l) g->level_ptr[l + 1] += g->level_ptr[l];
  g->level_nodes.resize(g->n);
  {
    std::vector<int32_t> fill(g->level_ptr.begin(), g->level_ptr.end() - 1);
    for (int32_t i = 0; i < g->n; ++i) g->level_nodes[fill[level[i]]++] = i;
  }

  g->listed_deg.assign(g->np, 0);
  for (int32_t i = 0; i < g->n; ++i) g->listed_deg[i] = g->listed_ptr[i + 1] - g->listed_ptr[i];

  g->tileable = g->np <= kMaxTileNodes && g->max_listed_deg <= 4095 && g->max_full_deg <= 4095;
  if (g->tileable) {
    g->edge_pair.assign((num_edges + 3) / 4 * 4, 0u);   // padded with (0,0): word ^ word == 0
    for (int64_t k = 0; k < num_edges; ++k) g->edge_pair[k] = uint32_t(g->edge_u[k]) | (uint32_t(g->edge_v[k]) << 16);
    // listed neighbours, natural order over the padded node range (slot == node)
    {
      std::vector<int32_t> order(g->np), gp{0, g->np}, ptr(g->listed_ptr);
      std::iota(order.begin(), order.end(), 0);
      ptr.resize(g->np + 1, g->listed_ptr.empty() ? 0 : g->listed_ptr.back());
      build_sell(order, gp, ptr, g->listed_col, g->sell_listed, nullptr);
    }
    // full neighbours in sweep order: by level, degree-descending inside a level (stable)
    {
      std::vector<int32_t> order(g->level_nodes);
      for (int32_t l = 0; l < g->levels; ++l)
        std::stable_sort(order.begin() + g->level_ptr[l], order.begin() + g->level_ptr[l + 1], [&](int32_t a, int32_t b) {
          return g->full_ptr[a + 1] - g->full_ptr[a] > g->full_ptr[b + 1] - g->full_ptr[b];
        });
      build_sell(order, g->level_ptr, g->full_ptr, g->full_col, g->sell_sweep, &g->level_slice);
    }
  }

  if (device >= 0 && g->tileable) {
    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    struct Blob { const void* src; size_t bytes; size_t off; };
    std::vector<Blob> blobs;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) {
      blobs.push_back({src, bytes, total});
      total += (bytes + 255) / 256 * 256 + 256;
      return int(blobs.size()) - 1;
    };
    auto addv32 = [&](const std::vector<int32_t>& v) { return add(v.data(), v.size() * 4); };
    auto addv16 = [&](const std::vector<uint16_t>& v) { return add(v.data(), v.size() * 2); };
    const int b_pair = add(g->edge_pair.data(), g->edge_pair.size() * 4);
    const int b_lptr = addv32(g->listed_ptr), b_lcol = addv32(g->listed_col), b_lrow = addv32(g->listed_row);
    const int b_ldeg = addv32(g->listed_deg), b_fptr = addv32(g->full_ptr), b_fcol = addv32(g->full_col);
    const int b_l_off = addv32(g->sell_listed.off), b_l_node = addv16(g->sell_listed.node);
    const int b_l_half = addv16(g->sell_listed.half), b_l_col = addv16(g->sell_listed.col);
    // the sweep structure is one contiguous blob so that a CTA can stage it with one bulk copy
    std::vector<uint8_t> sblob;
    size_t s_lvs, s_off, s_node, s_half, s_col;
    {
      auto put = [&](const void* src, size_t bytes) {
        const size_t at = sblob.size();
        sblob.resize(at + (bytes + 15) / 16 * 16, 0);
        if (bytes) memcpy(sblob.data() + at, src, bytes);
        return at;
      };
      s_lvs = put(g->level_slice.data(), g->level_slice.size() * 4);
      s_off = put(g->sell_sweep.off.data(), g->sell_sweep.off.size() * 4);
      s_node = put(g->sell_sweep.node.data(), g->sell_sweep.node.size() * 2);
      s_half = put(g->sell_sweep.half.data(), g->sell_sweep.half.size() * 2);
      s_col = put(g->sell_sweep.col.data(), g->sell_sweep.col.size() * 2);
    }
    const int b_sweep = add(sblob.data(), sblob.size());
    const int b_wpair = add(g->wpair.data(), g->wpair.size() * 4), b_wmeta = addv32(g->wmeta);
    const int b_fullw = addv32(g->full_w), b_wdeg = addv32(g->wdeg);
    if (e == cudaSuccess) e = cudaMalloc(&g->dev_blob, total);
    // the side stream never synchronises with the legacy default stream (non-blocking) and its blocks are
    // dispatched ahead of pending blocks of the caller's stream (highest priority)
    int prio_least = 0, prio_greatest = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&g->side_stream, cudaStreamNonBlocking, prio_greatest);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming);
    for (size_t a = 0; a < blobs.size() && e == cudaSuccess; ++a)
      if (blobs[a].bytes)
        e = cudaMemcpy((char*)g->dev_blob + blobs[a].off, blobs[a].src, blobs[a].bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("graph_create: CUDA error %s", cudaGetErrorString(e));
      if (g->dev_blob) cudaFree(g->dev_blob);
      cudaSetDevice(prev);
      delete g;
      return RLSB_ERR_CUDA;
    }
    cudaSetDevice(prev);
    auto i32 = [&](int b) { return reinterpret_cast<const int32_t*>((char*)g->dev_blob + blobs[b].off); };
    auto u16 = [&](int b) { return reinterpret_cast<const uint16_t*>((char*)g->dev_blob + blobs[b].off); };
    GraphDev& d = g->dev;
    d.n = g->n, d.np = g->np, d.m = int32_t(g->m), d.md = int32_t(g->listed_col.size());
    d.mf = int32_t(g->full_col.size()), d.levels = g->levels, d.bidir = g->bidir;
    d.max_listed_deg = g->max_listed_deg, d.max_full_deg = g->max_full_deg;
    d.edge_pair = reinterpret_cast<const uint32_t*>(i32(b_pair));
    d.listed_ptr = i32(b_lptr), d.listed_col = i32(b_lcol), d.listed_row = i32(b_lrow), d.listed_deg = i32(b_ldeg);
    d.full_ptr = i32(b_fptr), d.full_col = i32(b_fcol);
    d.listed = SellDev{int32_t(g->sell_listed.off.size()) - 1, i32(b_l_off), u16(b_l_node), u16(b_l_half), u16(b_l_col)};
    const char* sb = (const char*)g->dev_blob + blobs[b_sweep].off;
    d.sweep_blob = sb, d.sweep_blob_bytes = int32_t(sblob.size());
    d.sweep_lvs = int32_t(s_lvs), d.sweep_off = int32_t(s_off), d.sweep_node = int32_t(s_node);
    d.sweep_half = int32_t(s_half), d.sweep_col = int32_t(s_col);
    d.num_sweep_slices = int32_t(g->sell_sweep.off.size()) - 1;
    d.wbuckets = g->weighted ? int32_t(g->wmeta.size() / 4) : 0;
    d.wpair = g->weighted ? reinterpret_cast<const uint32_t*>(i32(b_wpair)) : nullptr;
    d.wmeta = g->weighted ? reinterpret_cast<const int4*>(i32(b_wmeta)) : nullptr;
    d.full_w = g->weighted ? i32(b_fullw) : nullptr;
    d.wdeg = g->weighted ? i32(b_wdeg) : nullptr;
    d.max_level_slices = 0;
    for (size_t l = 0; l + 1 < g->level_slice.size(); ++l)
      d.max_level_slices = std::max(d.max_level_slices, g->level_slice[l + 1] - g->level_slice[l]);
  }
  *out = g;
  return RLSB_OK;
}

int rlsb_graph_destroy(rlsb_graph_t* g) {
  if (!g) return RLSB_OK;
  if (g->ev_fork) cudaEventDestroy(g->ev_fork);
  if (g->ev_join) cudaEventDestroy(g->ev_join);
  if (g->side_stream) cudaStreamDestroy(g->side_stream);
  if (g->dev_blob) cudaFree(g->dev_blob);
  delete g;
  return RLSB_OK;
}

int32_t rlsb_graph_num_nodes(const rlsb_graph_t* g) { return g ? g->n : 0; }
int32_t rlsb_graph_padded_nodes(const rlsb_graph_t* g) { return g ? g->np : 0; }
int64_t rlsb_graph_num_edges(const rlsb_graph_t* g) { return g ? g->m : 0; }
int64_t rlsb_graph_num_listed(const rlsb_graph_t* g) { return g ? int64_t(g->listed_col.size()) : 0; }
int64_t rlsb_graph_num_full(const rlsb_graph_t* g) { return g ? int64_t(g->full_col.size()) : 0; }
int32_t rlsb_graph_num_levels(const rlsb_graph_t* g) { return g ? g->levels : 0; }
int32_t rlsb_graph_max_listed_degree(const rlsb_graph_t* g) { return g ? g->max_listed_deg : 0; }
int32_t rlsb_graph_max_full_degree(const rlsb_graph_t* g) { return g ? g->max_full_deg : 0; }

int rlsb_graph_export(const rlsb_graph_t* g, int32_t* h_listed_ptr, int32_t* h_listed_col, int32_t* h_full_ptr,
                      int32_t* h_full_col, int32_t* h_level_ptr, int32_t* h_level_nodes) {
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_INVALID, "graph_export: null graph");
  auto cp = [](int32_t* dst, const std::vector<int32_t>& v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(int32_t));
  };
  cp(h_listed_ptr, g->listed_ptr), cp(h_listed_col, g->listed_col);
  cp(h_full_ptr, g->full_ptr), cp(h_full_col, g->full_col);
  cp(h_level_ptr, g->level_ptr), cp(h_level_nodes, g->level_nodes);
  return RLSB_OK;
}

int rlsb_graph_sell_sizes(const rlsb_graph_t* g, int32_t which, int64_t* out3) {
  RLSB_REQUIRE(g != nullptr && out3 != nullptr && (which == 0 || which == 1), RLSB_ERR_INVALID, "graph_sell_sizes: bad argument");
  RLSB_REQUIRE(g->tileable, RLSB_ERR_UNSUPPORTED, "graph_sell_sizes: graph has no tile structures");
  const auto& s = which ? g->sell_sweep : g->sell_listed;
  out3[0] = int64_t(s.off.size()) - 1, out3[1] = int64_t(s.col.size()), out3[2] = int64_t(g->level_slice.size());
  return RLSB_OK;
}

int rlsb_graph_sell_export(const rlsb_graph_t* g, int32_t which, int32_t* h_off, uint16_t* h_node, uint16_t* h_half,
                           uint16_t* h_col, int32_t* h_level_slice) {
  RLSB_REQUIRE(g != nullptr && (which == 0 || which == 1), RLSB_ERR_INVALID, "graph_sell_export: bad argument");
  RLSB_REQUIRE(g->tileable, RLSB_ERR_UNSUPPORTED, "graph_sell_export: graph has no tile structures");
  const auto& s = which ? g->sell_sweep : g->sell_listed;
  if (h_off) memcpy(h_off, s.off.data(), s.off.size() * 4);
  if (h_node && !s.node.empty()) memcpy(h_node, s.node.data(), s.node.size() * 2);
  if (h_half && !s.half.empty()) memcpy(h_half, s.half.data(), s.half.size() * 2);
  if (h_col && !s.col.empty()) memcpy(h_col, s.col.data(), s.col.size() * 2);
  if (h_level_slice && !g->level_slice.empty()) memcpy(h_level_slice, g->level_slice.data(), g->level_slice.size() * 4);
  return RLSB_OK;
}

}  // extern "C"
