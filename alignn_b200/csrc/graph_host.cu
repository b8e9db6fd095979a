// Host-side (CPU) graph structure builders of libalignn_b200.so: the sorted-CSR edge index and the line graph.
//
// The reference builds graph structure on the CPU inside DGL (`dgl.graph`, `g.line_graph(shared=True)`,
// alignn/graphs.py:544,588) and caches it; these two entry points are the native replacement for that
// structure step (SURVEY.md section 8b lists `csr_build` / `line_graph_build`).  Pure integer work on host
// arrays, O(E) counting sorts, results bit-identical to the numpy restatement in oracle/alignn_oracle.py
// (csr_by_key, line_graph) -- checked in tests/test_host_logic.py.  No CUDA calls: usable on a box without a GPU.
#include <stdint.h>
#include <vector>

#include "alignn_b200.h"

namespace {

// stable counting sort of edge ids by key; ptr[n+1], eid[E]; returns the largest bucket size
int64_t counting_sort(const int64_t* key, int64_t E, int64_t n, int32_t* ptr, int32_t* eid) {
  std::vector<int64_t> cnt((size_t)n + 1, 0);
  for (int64_t e = 0; e < E; ++e) ++cnt[(size_t)key[e] + 1];
  int64_t max_deg = 0;
  for (int64_t v = 0; v < n; ++v) {
    if (cnt[(size_t)v + 1] > max_deg) max_deg = cnt[(size_t)v + 1];
    cnt[(size_t)v + 1] += cnt[(size_t)v];
  }
  for (int64_t v = 0; v <= n; ++v) ptr[v] = (int32_t)cnt[(size_t)v];
  for (int64_t e = 0; e < E; ++e) eid[cnt[(size_t)key[e]]++] = (int32_t)e;
  return max_deg;
}

}  // namespace

extern "C" {

int alignn_b200_csr_build_host(const int64_t* src, const int64_t* dst, int64_t num_nodes, int64_t num_edges,
                               int32_t* src32, int32_t* dst32, int32_t* in_ptr, int32_t* in_eid, int32_t* out_ptr,
                               int32_t* out_eid, int32_t* dst_sorted, int32_t* max_in_degree) {
  if (num_nodes < 0 || num_edges < 0 || num_nodes >= ((int64_t)1 << 31) || num_edges >= ((int64_t)1 << 31))
    return ALIGNN_ERR_BAD_ARG;
  if (!in_ptr || !out_ptr || (num_edges > 0 && (!src || !dst || !src32 || !dst32 || !in_eid || !out_eid)))
    return ALIGNN_ERR_BAD_ARG;
  int sorted = 1;
  for (int64_t e = 0; e < num_edges; ++e) {
    if (src[e] < 0 || dst[e] < 0 || src[e] >= num_nodes || dst[e] >= num_nodes) return ALIGNN_ERR_BAD_ARG;
    if (e > 0 && dst[e] < dst[e - 1]) sorted = 0;
    src32[e] = (int32_t)src[e];
    dst32[e] = (int32_t)dst[e];
  }
  const int64_t max_in = counting_sort(dst, num_edges, num_nodes, in_ptr, in_eid);
  counting_sort(src, num_edges, num_nodes, out_ptr, out_eid);
  if (dst_sorted) *dst_sorted = sorted;
  if (max_in_degree) *max_in_degree = (int32_t)max_in;
  return ALIGNN_OK;
}

int64_t alignn_b200_line_graph_count_host(const int32_t* src, const int32_t* in_ptr, const int32_t* in_eid,
                                          int64_t num_edges) {
  if (num_edges < 0 || (num_edges > 0 && (!src || !in_ptr || !in_eid))) return -1;
  int64_t T = 0;
  for (int64_t j = 0; j < num_edges; ++j) {
    const int32_t a = src[j];
    T += in_ptr[a + 1] - in_ptr[a];
    // a bond never pairs with itself: only a self-loop bond j can appear in the in-list of its own source
    for (int32_t p = in_ptr[a]; p < in_ptr[a + 1]; ++p)
      if (in_eid[p] == j) { --T; break; }
  }
  return T;
}

int alignn_b200_line_graph_build_host(const int32_t* src, const int32_t* in_ptr, const int32_t* in_eid,
                                      int64_t num_edges, const int64_t* batch_num_edges, int64_t batch_size,
                                      int64_t capacity, int64_t* lsrc, int64_t* ldst, int64_t* l_batch_num_edges) {
  if (num_edges < 0 || capacity < 0 || batch_size < 0) return ALIGNN_ERR_BAD_ARG;
  if (num_edges > 0 && (!src || !in_ptr || !in_eid)) return ALIGNN_ERR_BAD_ARG;
  if (capacity > 0 && (!lsrc || !ldst)) return ALIGNN_ERR_BAD_ARG;
  if (batch_size > 0 && (!batch_num_edges || !l_batch_num_edges)) return ALIGNN_ERR_BAD_ARG;
  int64_t t = 0, b = 0, b_end = batch_size > 0 ? batch_num_edges[0] : num_edges;
  for (int64_t i = 0; i < batch_size; ++i) l_batch_num_edges[i] = 0;
  for (int64_t j = 0; j < num_edges; ++j) {          // destination-major: all pairs (i -> j) of bond j, i ascending
    while (batch_size > 0 && j >= b_end && b + 1 < batch_size) b_end += batch_num_edges[++b];
    const int32_t a = src[j];
    for (int32_t p = in_ptr[a]; p < in_ptr[a + 1]; ++p) {
      const int32_t i = in_eid[p];                    // bond i ends where bond j starts
      if (i == j) continue;
      if (t >= capacity) return ALIGNN_ERR_WORKSPACE;
      lsrc[t] = i;
      ldst[t] = j;
      ++t;
      if (batch_size > 0) ++l_batch_num_edges[b];
    }
  }
  return t == capacity ? ALIGNN_OK : ALIGNN_ERR_WORKSPACE;
}

}  // extern "C"
