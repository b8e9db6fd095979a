/* STAGED WORK -- not part of libalignn_b200.so (see egc_fused.h for the status of this directory).
 *
 * Device-side twins of the host structure builders and of the ALIGNN-FF force / stress reductions
 * (SURVEY.md section 8b "csr_build, line_graph_build, pair_force_scatter" and section 8f ranks 1-2).
 * All pointers are device pointers unless noted; the caller owns every buffer including the workspace; calls only
 * enqueue on `stream`.  Results are bit-identical to the host builders (integer work) / deterministic (d=3 sums).
 */
#ifndef ALIGNN_B200_STAGED_GRAPH_DEVICE_H
#define ALIGNN_B200_STAGED_GRAPH_DEVICE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Sorted-CSR edge index on the device = alignn_b200_csr_build_host (`dgl.graph((u, v))`, alignn/graphs.py:544).
 * in_eid / out_eid: edge ids stably sorted by destination / source; flags[0] = 1 if dst is already non-decreasing,
 * flags[1] = largest in-degree. */
size_t alignn_b200_csr_build_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int alignn_b200_csr_build(const int32_t* src, const int32_t* dst, int64_t num_nodes, int64_t num_edges, int32_t* in_ptr,
                          int32_t* in_eid, int32_t* out_ptr, int32_t* out_eid, int32_t* flags, void* workspace,
                          size_t workspace_bytes, void* stream);

/* Line graph on the device = alignn_b200_line_graph_{count,build}_host (`g.line_graph(shared=True)`,
 * alignn/graphs.py:588): offsets[j] = number of pairs (i -> j') with j' < j, offsets[E] = T (read it back to size
 * lsrc / ldst); the pairs are written destination-major with ascending sources, so `offsets` is L(g)'s in_ptr. */
size_t alignn_b200_line_graph_workspace_bytes(int64_t num_edges);
int alignn_b200_line_graph_offsets(const int32_t* src, const int32_t* dst, const int32_t* in_ptr, int64_t num_edges,
                                   int32_t* offsets, void* workspace, size_t workspace_bytes, void* stream);
int alignn_b200_line_graph_fill(const int32_t* src, const int32_t* dst, const int32_t* in_ptr, const int32_t* in_eid,
                                int64_t num_edges, const int32_t* offsets, int32_t* lsrc, int32_t* ldst, void* stream);

/* Periodic radius graph on the device = alignn_b200_radius_graph_{count,build}_host (alignn/graphs.py:267-364): bonds
 * u -> v for every image `c` of v with atol < |x_v + shifts[c] - x_u| <= cutoff, in (u, c, v) order, double precision
 * with the host builder's operation order (bit-identical bond list and displacement vectors).  offsets[u] = first bond
 * of atom u, offsets[N] = number of bonds (read it back to size the outputs); fill with empty outputs is an error only
 * if bonds exist.  The caller computes `shifts = cells @ lattice` and handles the cutoff-growth retry
 * (graphs.py:347-350) exactly as alignn_b200.neighbors.radius_graph does for the host scan. */
size_t alignn_b200_radius_graph_workspace_bytes(int64_t num_atoms);
int alignn_b200_radius_graph_offsets(const double* cart_coords, const double* shifts, int64_t num_atoms, int64_t num_images,
                                     double cutoff, double atol, int32_t* offsets, void* workspace, size_t workspace_bytes,
                                     void* stream);
int alignn_b200_radius_graph_fill(const double* cart_coords, const double* shifts, int64_t num_atoms, int64_t num_images,
                                  double cutoff, double atol, const int32_t* offsets, int32_t* u, int32_t* v,
                                  int32_t* image_index, float* r, void* stream);

/* forces[v] = sum over in-edges of pair_forces - (add_reverse ? sum over out-edges : 0)   (alignn_atomwise.py:547-563:
 * update_all(copy_e, sum) on g and on dgl.reverse(g)); pair_forces [E,3], forces [Nn,3].  in_eid NULL = identity. */
int alignn_b200_pair_force_scatter(const float* pair_forces, const int32_t* in_ptr, const int32_t* in_eid,
                                   const int32_t* out_ptr, const int32_t* out_eid, int64_t num_nodes, int add_reverse,
                                   float* forces, void* stream);

/* stress[b] = multiplier * -160.21766208 * (r_b^T pair_forces_b) / V[node_offsets[b]]   (alignn_atomwise.py:610-635);
 * edge_offsets / node_offsets [B+1] int64 prefix sums of batch_num_edges / batch_num_nodes; stress [B,3,3]. */
int alignn_b200_virial_stress(const float* r, const float* pair_forces, const int64_t* edge_offsets,
                              const int64_t* node_offsets, const float* V, int64_t batch_size, float multiplier,
                              float* stress, void* stream);

#ifdef __cplusplus
}
#endif
#endif
