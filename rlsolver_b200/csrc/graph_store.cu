// Graph store: native CSR / edge-list / sweep-level builder and its device image.
// Replaces EnvMaxcut.__init__ (rlsolver/envs/env_L2A.py:25-52), build_adjacency_indies
// (rlsolver/methods/util_read_data.py:144-187) and calc_num_nodes_in_mygraph
// (rlsolver/methods/util.py:35-40).  Host side is plain C++; the device image is a
// handful of int32 arrays that stay L2-resident (<= a few hundred KB for Gset).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "common.cuh"

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

struct rlsb_graph {
  int32_t n = 0, np = 0, bidir = 0, device = -1, levels = 0, max_listed_deg = 0, max_full_deg = 0;
  int64_t m = 0;
  std::vector<int32_t> edge_u, edge_v, weight;
  std::vector<int32_t> listed_ptr, listed_col, listed_row, full_ptr, full_col, level_ptr, level_nodes;
  void* dev_blob = nullptr;   // one allocation holding every device array
  rlsb::GraphDev dev{};
};

namespace rlsb {
const GraphDev* graph_dev(const rlsb_graph_t* g) { return (g && g->dev_blob) ? &g->dev : nullptr; }
int graph_device_id(const rlsb_graph_t* g) { return g ? g->device : -1; }
}  // namespace rlsb

namespace {

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

}  // namespace

extern "C" {

int rlsb_version(void) { return 100; }
const char* rlsb_last_error(void) { return rlsb::g_last_error.c_str(); }

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

  if (device >= 0) {
    int prev = 0;
    cudaError_t e = cudaGetDevice(&prev);
    if (e == cudaSuccess) e = cudaSetDevice(device);
    constexpr int kArrays = 9;
    const std::vector<int32_t>* arrays[kArrays] = {&g->edge_u,     &g->edge_v,   &g->listed_ptr,
                                                   &g->listed_col, &g->listed_row, &g->full_ptr,
                                                   &g->full_col,   &g->level_ptr,  &g->level_nodes};
    size_t offs[kArrays], total = 0;
    for (int a = 0; a < kArrays; ++a) {
      offs[a] = total;
      total += (arrays[a]->size() * sizeof(int32_t) + 255) / 256 * 256 + 256;
    }
    if (e == cudaSuccess) e = cudaMalloc(&g->dev_blob, total);
    for (int a = 0; a < kArrays && e == cudaSuccess; ++a)
      if (!arrays[a]->empty())
        e = cudaMemcpy((char*)g->dev_blob + offs[a], arrays[a]->data(), arrays[a]->size() * sizeof(int32_t),
                       cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      set_error("graph_create: CUDA error %s", cudaGetErrorString(e));
      if (g->dev_blob) cudaFree(g->dev_blob);
      cudaSetDevice(prev);
      delete g;
      return RLSB_ERR_CUDA;
    }
    cudaSetDevice(prev);
    auto at = [&](int a) { return reinterpret_cast<const int32_t*>((char*)g->dev_blob + offs[a]); };
    g->dev = GraphDev{g->n,  g->np, int32_t(g->m), int32_t(g->listed_col.size()), int32_t(g->full_col.size()),
                      g->levels, g->bidir, at(0), at(1), at(2), at(3), at(4), at(5), at(6), at(7), at(8)};
  }
  *out = g;
  return RLSB_OK;
}

int rlsb_graph_destroy(rlsb_graph_t* g) {
  if (!g) return RLSB_OK;
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

}  // extern "C"
