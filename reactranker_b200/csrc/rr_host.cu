// Host-side batch builder (no device work): the control block that rr_graph_assemble consumes, written straight into the caller's
// pinned staging buffer.  Replaces the per-step numpy arithmetic of the batch path (load_reactions.py:549-578 -> featurization.py:246-290:
// where every molecule's rows go, which padding rows and max_num_bonds its segment has) with one pass over the molecule ids.
#include "rr_common.cuh"

namespace rr {

// ctl layout (int32): [ids | a_start | b_start | W | pad_bond | pad_atom] x n_mols, then [a0 | b0 | W] x n_seg  -- what rr_graph_assemble reads.
// dims (int64[5]): n_atoms, n_bonds, n_mols, wmax, n_seg.   seg_A / seg_B / seg_W (int64[n_seg], optional): rows and max_num_bonds per segment.
int batch_build(int n_mols, const int32_t* ids, int n_store, const int32_t* st_nA, const int32_t* st_nB, const int32_t* st_maxdeg, int n_seg,
                const int64_t* seg_lens, const int64_t* W_override, int32_t* ctl, int64_t* dims, int64_t* seg_A, int64_t* seg_B, int64_t* seg_W) {
  RR_REQUIRE(n_mols >= 0 && n_seg >= 0 && (n_mols == 0 || ids) && st_nA && st_nB && st_maxdeg && (n_seg == 0 || seg_lens) && ctl && dims,
             "batch_build: NULL argument");
  int32_t* c_ids = ctl;
  int32_t* c_a = ctl + static_cast<size_t>(n_mols);
  int32_t* c_b = ctl + 2 * static_cast<size_t>(n_mols);
  int32_t* c_W = ctl + 3 * static_cast<size_t>(n_mols);
  int32_t* c_pb = ctl + 4 * static_cast<size_t>(n_mols);
  int32_t* c_pa = ctl + 5 * static_cast<size_t>(n_mols);
  int32_t* s_a0 = ctl + 6 * static_cast<size_t>(n_mols);
  int32_t* s_b0 = s_a0 + n_seg;
  int32_t* s_W = s_b0 + n_seg;
  long long a = 0, b = 0, m = 0, wmax = 1, total = 0;
  for (int s = 0; s < n_seg; ++s) total += seg_lens[s];
  if (total != n_mols) return fail(RR_ERR_INVALID, "segment lengths sum to %lld, %d molecule ids given", total, n_mols);
  for (int s = 0; s < n_seg; ++s) {
    const long long a0 = a, b0 = b, m0 = m;
    ++a;                      // the segment's padding atom / bond row (featurization.py:255-264)
    ++b;
    int deg = 0;
    for (long long j = 0; j < seg_lens[s]; ++j, ++m) {
      const int id = ids[m];
      if (id < 0 || id >= n_store) return fail(RR_ERR_INVALID, "molecule id %d outside the store (%d molecules)", id, n_store);
      c_ids[m] = id;
      c_a[m] = static_cast<int32_t>(a);
      c_b[m] = static_cast<int32_t>(b);
      c_pb[m] = static_cast<int32_t>(b0);
      c_pa[m] = static_cast<int32_t>(a0);
      a += st_nA[id];
      b += st_nB[id];
      if (st_maxdeg[id] > deg) deg = st_maxdeg[id];
    }
    const long long W_min = deg > 1 ? deg : 1;          // max_num_bonds = max(1, largest in-degree)   featurization.py:281
    long long W = W_min;
    if (W_override) {
      W = W_override[s];
      if (W < W_min) return fail(RR_ERR_INVALID, "max_num_bonds override %lld < this batch's in-degree %lld", W, W_min);
    }
    for (long long j = m0; j < m; ++j) c_W[j] = static_cast<int32_t>(W);
    s_a0[s] = static_cast<int32_t>(a0);
    s_b0[s] = static_cast<int32_t>(b0);
    s_W[s] = static_cast<int32_t>(W);
    if (W_min > wmax) wmax = W_min;
    if (seg_A) seg_A[s] = a - a0;
    if (seg_B) seg_B[s] = b - b0;
    if (seg_W) seg_W[s] = W;
    if (a > 2147483647LL || b > 2147483647LL) return fail(RR_ERR_INVALID, "batch exceeds 2^31 rows");
  }
  dims[0] = a;
  dims[1] = b;
  dims[2] = n_mols;
  dims[3] = n_seg ? wmax : 1;
  dims[4] = n_seg;
  return RR_OK;
}

}  // namespace rr

extern "C" int rr_batch_build(int n_mols, const int32_t* h_ids, int n_store, const int32_t* h_store_n_atoms, const int32_t* h_store_n_bonds,
                              const int32_t* h_store_max_degree, int n_segments, const int64_t* h_seg_lens, const int64_t* h_W_override,
                              int32_t* h_ctl, int64_t* h_dims, int64_t* h_seg_atoms, int64_t* h_seg_bonds, int64_t* h_seg_W) {
  return rr::batch_build(n_mols, h_ids, n_store, h_store_n_atoms, h_store_n_bonds, h_store_max_degree, n_segments, h_seg_lens, h_W_override, h_ctl,
                         h_dims, h_seg_atoms, h_seg_bonds, h_seg_W);
}
