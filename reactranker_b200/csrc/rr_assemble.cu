// On-device batch assembly from a device-resident molecule store.
//
// The reference rebuilds every BatchMolGraph on the host from cached MolGraph objects and ships ~36 kB per reaction over
// PCIe each step (load_reactions.py:549-578, featurization.py:246-290, mpn.py:77).  Here the per-molecule arrays live in
// HBM once (the B200 has room for tens of millions of molecules); a step ships only molecule ids and row offsets
// (a few bytes per molecule) and this kernel writes the rr_graph: contiguous row copies of the features plus the
// offset index tables and the per-atom padding records -- bit-identical to the host-packed DeviceGraph.
#include "rr_common.cuh"

namespace rr {

__global__ void __launch_bounds__(256) k_assemble(rr_mol_store st, int n_mols, const int* __restrict__ mol_ids, const int* __restrict__ a_start,
                                                  const int* __restrict__ b_start, const int* __restrict__ mol_W, const int* __restrict__ mol_pb,
                                                  const int* __restrict__ mol_pa, rr_graph g) {
  const int mol = blockIdx.x;
  if (mol >= n_mols) return;
  const int sid = __ldg(mol_ids + mol);
  const int nA = __ldg(st.n_atoms + sid), nB = __ldg(st.n_bonds + sid);
  const long long aoff = __ldg(st.atom_off + sid), boff = __ldg(st.bond_off + sid);
  const int a0 = __ldg(a_start + mol), b0 = __ldg(b_start + mol);
  const int W = __ldg(mol_W + mol), pb = __ldg(mol_pb + mol), pa = __ldg(mol_pa + mol);
  int* a2b = const_cast<int*>(g.a2b);
  int* a2r = const_cast<int*>(g.a2b_rev);
  int* a2a = const_cast<int*>(g.a2a);
  int4* meta = reinterpret_cast<int4*>(const_cast<rr_atom_meta*>(g.a_meta));
  for (int j = threadIdx.x; j < nA; j += blockDim.x) {
    const int deg = __ldg(st.deg + aoff + j);
    const int s0 = __ldg(st.a2b_start + aoff + j);
    meta[a0 + j] = make_int4(deg, W - deg, pb, pa);
    const size_t row = static_cast<size_t>(a0 + j) * g.wmax;
    for (int k = 0; k < g.wmax; ++k) {
      int vb = 0, vr = 0, va = 0;
      if (k < deg) {
        const int lb = __ldg(st.a2b_flat + boff + s0 + k);
        vb = b0 + lb;
        vr = b0 + (lb ^ 1);                      // b2revb = b xor 1 inside a molecule (featurization.py:202-210)
        va = a0 + __ldg(st.b2a + boff + lb);
      }
      a2b[row + k] = vb;
      a2r[row + k] = vr;
      a2a[row + k] = va;
    }
  }
  // feature rows: a molecule's rows are contiguous in the store and in the batch
  {
    const float4* src = reinterpret_cast<const float4*>(st.f_atoms + aoff * RR_FA_LD);
    float4* dst = reinterpret_cast<float4*>(const_cast<float*>(g.f_atoms) + static_cast<size_t>(a0) * RR_FA_LD);
    for (int i = threadIdx.x; i < nA * (RR_FA_LD / 4); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  {
    const float4* src = reinterpret_cast<const float4*>(st.f_bonds + boff * RR_FB_LD);
    float4* dst = reinterpret_cast<float4*>(const_cast<float*>(g.f_bonds) + static_cast<size_t>(b0) * RR_FB_LD);
    for (int i = threadIdx.x; i < nB * (RR_FB_LD / 4); i += blockDim.x) dst[i] = __ldg(src + i);
  }
  if (threadIdx.x == 0) {
    const_cast<int*>(g.mol_start)[mol] = a0;
    const_cast<int*>(g.mol_size)[mol] = nA;
  }
}

// padding rows of every segment: zero features, pad-atom record, zero index rows (featurization.py:255-264)
__global__ void k_assemble_pads(int n_seg, const int* __restrict__ seg_pa, const int* __restrict__ seg_pb, const int* __restrict__ seg_W, rr_graph g) {
  const int s = blockIdx.x;
  if (s >= n_seg) return;
  const int pa = __ldg(seg_pa + s), pb = __ldg(seg_pb + s), W = __ldg(seg_W + s);
  float* fa = const_cast<float*>(g.f_atoms) + static_cast<size_t>(pa) * RR_FA_LD;
  float* fb = const_cast<float*>(g.f_bonds) + static_cast<size_t>(pb) * RR_FB_LD;
  for (int i = threadIdx.x; i < RR_FA_LD; i += blockDim.x) fa[i] = 0.f;
  for (int i = threadIdx.x; i < RR_FB_LD; i += blockDim.x) fb[i] = 0.f;
  for (int k = threadIdx.x; k < g.wmax; k += blockDim.x) {
    const_cast<int*>(g.a2b)[static_cast<size_t>(pa) * g.wmax + k] = 0;
    const_cast<int*>(g.a2b_rev)[static_cast<size_t>(pa) * g.wmax + k] = 0;
    const_cast<int*>(g.a2a)[static_cast<size_t>(pa) * g.wmax + k] = 0;
  }
  if (threadIdx.x == 0) {
    reinterpret_cast<int4*>(const_cast<rr_atom_meta*>(g.a_meta))[pa] = make_int4(0x100, W, pb, pa);
    const_cast<int*>(g.pad_atoms)[s] = pa;
    const_cast<int*>(g.pad_bonds)[s] = pb;
  }
}

int graph_assemble(const rr_mol_store* st, int n_mols, const int* mol_ids, const int* a_start, const int* b_start, const int* mol_W,
                   const int* mol_pb, const int* mol_pa, int n_seg, const int* seg_pa, const int* seg_pb, const int* seg_W, const rr_graph* out,
                   cudaStream_t s) {
  ProfScope prof_scope(KC_MISC, s);
  RR_REQUIRE(st && out && n_mols >= 0 && n_seg > 0, "graph_assemble: NULL argument");
  RR_REQUIRE(st->f_atoms && st->f_bonds && st->n_atoms && st->n_bonds && st->atom_off && st->bond_off && st->deg && st->a2b_start && st->a2b_flat && st->b2a,
             "graph_assemble: molecule store is incomplete");
  RR_REQUIRE(out->f_atoms && out->f_bonds && out->a_meta && out->a2b && out->a2b_rev && out->a2a && out->mol_start && out->mol_size && out->pad_atoms &&
                 out->pad_bonds && out->wmax > 0,
             "graph_assemble: output graph buffers missing");
  RR_REQUIRE(aligned16(out->f_atoms) && aligned16(out->f_bonds) && aligned16(out->a_meta) && aligned16(st->f_atoms) && aligned16(st->f_bonds),
             "graph_assemble: buffers must be 16-byte aligned");
  k_assemble_pads<<<n_seg, 128, 0, s>>>(n_seg, seg_pa, seg_pb, seg_W, *out);
  RR_LAUNCH_CHECK("k_assemble_pads");
  if (n_mols > 0) {
    k_assemble<<<n_mols, 256, 0, s>>>(*st, n_mols, mol_ids, a_start, b_start, mol_W, mol_pb, mol_pa, *out);
    RR_LAUNCH_CHECK("k_assemble");
  }
  return RR_OK;
}

}  // namespace rr
