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

// -------------------------------------------------------------------------------------------------
// Periodic radius graph (alignn/graphs.py:267-364): bond u -> v for every (atom u of the home cell, periodic image
// `c` of atom v) with atol < |x_v + shift_c - x_u| <= cutoff, emitted in (u, c, v) order -- the order torch.where
// gives on the reference's [N, images*N] mask.  `shifts` are the cartesian image offsets (cells @ lattice), computed
// by the caller so that the arithmetic matches the restatement bit for bit.  Double precision, same operation
// order as the numpy/torch restatements: d = (shift + x_v) - x_u ; dist = sqrt((dx*dx + dy*dy) + dz*dz).
// -------------------------------------------------------------------------------------------------
#include <cmath>

namespace {
template <bool kFill>
int64_t radius_scan(const double* X, const double* shifts, int64_t n, int64_t n_images, double cutoff, double atol,
                    int64_t capacity, int64_t* u_out, int64_t* v_out, int64_t* c_out, float* r_out) {
  int64_t t = 0;
  // bounding box of the home cell's atoms: an image whose shifted box is farther than the cutoff (plus a guard
  // band for rounding) from x_u cannot hold a neighbour and is skipped as a whole -- pure pruning, the bonds found
  // and their order are unchanged
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  for (int64_t v = 0; v < n; ++v)
    for (int k = 0; k < 3; ++k) {
      const double x = X[3 * v + k];
      if (v == 0 || x < lo[k]) lo[k] = x;
      if (v == 0 || x > hi[k]) hi[k] = x;
    }
  const double reach = cutoff * (1.0 + 1e-9) + 1e-9;
  for (int64_t u = 0; u < n; ++u) {
    const double xu = X[3 * u], yu = X[3 * u + 1], zu = X[3 * u + 2];
    const double pu[3] = {xu, yu, zu};
    for (int64_t c = 0; c < n_images; ++c) {
      const double sx = shifts[3 * c], sy = shifts[3 * c + 1], sz = shifts[3 * c + 2];
      double gap2 = 0.0;
      for (int k = 0; k < 3; ++k) {
        const double a = lo[k] + shifts[3 * c + k], b = hi[k] + shifts[3 * c + k];
        const double g = pu[k] < a ? a - pu[k] : (pu[k] > b ? pu[k] - b : 0.0);
        gap2 += g * g;
      }
      if (gap2 > reach * reach) continue;
      for (int64_t v = 0; v < n; ++v) {
        const double dx = (sx + X[3 * v]) - xu, dy = (sy + X[3 * v + 1]) - yu, dz = (sz + X[3 * v + 2]) - zu;
        const double dist = std::sqrt((dx * dx + dy * dy) + dz * dz);
        if (dist <= cutoff && !(std::fabs(dist) <= atol)) {
          if (kFill) {
            if (t >= capacity) return -1;
            u_out[t] = u; v_out[t] = v; c_out[t] = c;
            r_out[3 * t] = (float)dx; r_out[3 * t + 1] = (float)dy; r_out[3 * t + 2] = (float)dz;
          }
          ++t;
        }
      }
    }
  }
  return t;
}
}  // namespace

extern "C" {

int64_t alignn_b200_radius_graph_count_host(const double* cart_coords, const double* shifts, int64_t num_atoms,
                                            int64_t num_images, double cutoff, double atol) {
  if (num_atoms < 0 || num_images < 0 || (num_atoms > 0 && !cart_coords) || (num_images > 0 && !shifts)) return -1;
  return radius_scan<false>(cart_coords, shifts, num_atoms, num_images, cutoff, atol, 0, nullptr, nullptr, nullptr, nullptr);
}

int alignn_b200_radius_graph_build_host(const double* cart_coords, const double* shifts, int64_t num_atoms,
                                        int64_t num_images, double cutoff, double atol, int64_t capacity, int64_t* u,
                                        int64_t* v, int64_t* image_index, float* r) {
  if (num_atoms < 0 || num_images < 0 || capacity < 0 || (num_atoms > 0 && !cart_coords) || (num_images > 0 && !shifts))
    return ALIGNN_ERR_BAD_ARG;
  if (capacity > 0 && (!u || !v || !image_index || !r)) return ALIGNN_ERR_BAD_ARG;
  const int64_t t = radius_scan<true>(cart_coords, shifts, num_atoms, num_images, cutoff, atol, capacity, u, v, image_index, r);
  return t == capacity ? ALIGNN_OK : ALIGNN_ERR_WORKSPACE;
}

}  // extern "C"
