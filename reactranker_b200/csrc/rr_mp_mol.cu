// Atom-level neighbour sums (the a2a gathers of MPNDiff, mpn.py:201, 215, and their backward), one MOLECULE per tile.
//
// The row pipeline (rr_mp_pipe.cu) fetches, for every atom, each of its neighbours' rows: on atom rows every source row is
// fetched once per neighbour that needs it (~2x with a mean degree of 2), from L2, and the kernel saturates at the same ~36 GB/s of
// SM-level traffic per SM as the bond-level gathers -- which is why the a2a sums sat at 0.56 of the HBM roofline against 0.78
// (profiles/r02_mp_kstage.md).  A molecule's atoms are contiguous rows and its neighbours never leave it, so here a CTA copies the
// molecule's rows into shared memory ONCE (coalesced 16-byte loads, <= 32 rows = 39 KB at h = 300: five CTAs per SM hide each other's
// latency), forms every atom's sum out of shared memory and writes it: each row is read once and written once.
//   forward   out[a]  = pad_count * src[pad_atom] + sum_k src[a2a[a, k]]                       (all rows, padding atoms included)
//   backward  dsrc[a] = sum_k dout[a2a[a, k]]  (the relation is symmetric), dsrc[pad_atom] += sum_a pad_count_a * dout[a],
//             optionally fused with the ReLU / inverted-dropout backward that follows it (rr_relu_bwd) exactly like the row pipeline
// Molecules larger than the tile read their neighbours straight from global memory (correct, slower); graphs without a molecule
// scope fall back to the row pipeline (RR_ERR_UNSUPPORTED).
#include "rr_common.cuh"

namespace rr {
namespace mol {

constexpr int TILE_ROWS = 32;

struct Args {
  rr_graph g;
  const float* src;
  float* out;
  int ld, relu_src;
  const float* y;   // fused epilogue (backward only; NULL: none)
  float scale;
  int preact;
  float* acc;
  int acc_mode, skip_out;
  int cpr, lanes;   // 16-byte chunks per row, row lanes per CTA (threads = cpr * lanes rounded up to a warp)
};

__device__ __forceinline__ float4 mask_scale(float4 d, float4 y, float scale, int preact) {
  float4 o;
  if (preact) {
    o.x = y.x > 0.f ? d.x * scale : 0.f;
    o.y = y.y > 0.f ? d.y * scale : 0.f;
    o.z = y.z > 0.f ? d.z * scale : 0.f;
    o.w = y.w > 0.f ? d.w * scale : 0.f;
  } else {
    o.x = y.x != 0.f ? d.x * scale : 0.f;
    o.y = y.y != 0.f ? d.y * scale : 0.f;
    o.z = y.z != 0.f ? d.z * scale : 0.f;
    o.w = y.w != 0.f ? d.w * scale : 0.f;
  }
  return o;
}

template <bool BWD, bool FUSED>
__global__ void __launch_bounds__(512) k_mol_nbr(const Args A) {
  extern __shared__ float4 tile4[];
  float* tile = reinterpret_cast<float*>(tile4);
  const rr_graph& g = A.g;
  const int ld = A.ld, R = A.lanes;
  const int t = threadIdx.x;
  const bool live = t < A.cpr * R;
  const int c4 = (t % A.cpr) * 4, rl = t / A.cpr;
  // contiguous run of molecules per CTA: a CTA meets one or two segments, so the atomics into a segment's padding row come from few CTAs
  const int per = g.n_mols / gridDim.x, rem = g.n_mols - per * gridDim.x;
  const int m_beg = blockIdx.x * per + min(static_cast<int>(blockIdx.x), rem);
  const int m_end = m_beg + per + (static_cast<int>(blockIdx.x) < rem ? 1 : 0);
  float4 pad_acc = f4_zero(), padv = f4_zero();
  int pad_row = -1, padv_row = -1;
  const float* src_c = A.src + c4;
  float* out_c = A.out + c4;

  for (int m = m_beg; m < m_end; ++m) {
    const int a0 = __ldg(g.mol_start + m), n = __ldg(g.mol_size + m);
    const bool staged = n <= TILE_ROWS;
    if (staged && live)
      for (int r = rl; r < n; r += R) {
        float4 v = ld_f4_stream(src_c + static_cast<size_t>(a0 + r) * ld);
        if (!BWD && A.relu_src) v = f4_relu(v);
        *reinterpret_cast<float4*>(tile + r * ld + c4) = v;
      }
    __syncthreads();
    if (live)
      for (int r = rl; r < n; r += R) {
        const int a = a0 + r;
        const int4 meta = __ldg(reinterpret_cast<const int4*>(g.a_meta) + a);
        const int deg = meta.x & 0xff, pad_count = meta.y, prow = meta.w;
        const int* ib = g.a2a + static_cast<size_t>(a) * g.wmax;
        float4 sum = f4_zero();
        for (int k = 0; k < deg; ++k) {
          const int j = __ldg(ib + k) - a0;
          float4 v;
          if (staged && j >= 0 && j < n) {
            v = *reinterpret_cast<const float4*>(tile + j * ld + c4);
          } else {
            v = ld_f4(src_c + static_cast<size_t>(j + a0) * ld);
            if (!BWD && A.relu_src) v = f4_relu(v);
          }
          sum = f4_add(sum, v);
        }
        const size_t off = static_cast<size_t>(a) * ld;
        if (!BWD) {
          if (pad_count > 0) {
            if (prow != padv_row) {
              padv = ld_f4(src_c + static_cast<size_t>(prow) * ld);
              if (A.relu_src) padv = f4_relu(padv);
              padv_row = prow;
            }
            sum = f4_fma(static_cast<float>(pad_count), padv, sum);
          }
          st_f4(out_c + off, sum);
        } else {
          if (prow != pad_row) {
            if (pad_row >= 0) red_add_f4(out_c + static_cast<size_t>(pad_row) * ld, pad_acc);
            pad_acc = f4_zero();
            pad_row = prow;
          }
          if (pad_count > 0) {
            const float4 self = staged ? *reinterpret_cast<const float4*>(tile + r * ld + c4) : ld_f4(src_c + off);
            pad_acc = f4_fma(static_cast<float>(pad_count), self, pad_acc);
          }
          if (FUSED) {
            const float4 o = mask_scale(sum, ld_f4_stream(A.y + off + c4), A.scale, A.preact);
            if (!A.skip_out) st_f4(out_c + off, o);
            if (A.acc_mode == 1) st_f4(A.acc + off + c4, o);
            else if (A.acc_mode == 2) st_f4(A.acc + off + c4, f4_add(*reinterpret_cast<const float4*>(A.acc + off + c4), o));
          } else {
            st_f4(out_c + off, sum);
          }
        }
      }
    __syncthreads();
  }
  if (BWD && live && pad_row >= 0) red_add_f4(out_c + static_cast<size_t>(pad_row) * ld, pad_acc);
  // the segments' padding atoms belong to no molecule: forward out[pad] = W * src[pad]; backward dsrc[pad] += W * dout[pad]
  if (live && rl == 0)
    for (int s = blockIdx.x; s < g.n_segments; s += gridDim.x) {
      const int pa = __ldg(g.pad_atoms + s);
      const int4 meta = __ldg(reinterpret_cast<const int4*>(g.a_meta) + pa);
      float4 v = ld_f4(src_c + static_cast<size_t>(pa) * ld);
      if (!BWD && A.relu_src) v = f4_relu(v);
      v = f4_scale(static_cast<float>(meta.y), v);
      if (!BWD) st_f4(out_c + static_cast<size_t>(pa) * ld, v);
      else red_add_f4(out_c + static_cast<size_t>(pa) * ld, v);
    }
}

}  // namespace mol

// op: 0 forward, 1 backward (y != NULL: fused ReLU-backward epilogue).  RR_ERR_UNSUPPORTED (nothing launched) when the graph carries no
// molecule scope or a row does not fit the tile: the caller then uses the row pipeline.
int moltile_launch(int op, const rr_graph* g, const float* src, float* out, int ld, int relu_src, const float* y, float scale, int preact, float* acc,
                   int acc_mode, int skip_out, cudaStream_t s) {
  using namespace mol;
  if (switches().mp_moltile == 0) return RR_ERR_UNSUPPORTED;
  if (!g->mol_start || !g->mol_size || !g->pad_atoms || g->n_mols <= 0 || g->n_segments <= 0 || (ld & 3)) return RR_ERR_UNSUPPORTED;
  const int cpr = ld >> 2;
  if (cpr < 1 || cpr > 512) return RR_ERR_UNSUPPORTED;
  Args A{};
  A.g = *g;
  A.src = src;
  A.out = out;
  A.ld = ld;
  A.relu_src = relu_src;
  A.y = y;
  A.scale = scale;
  A.preact = preact;
  A.acc = acc;
  A.acc_mode = y ? acc_mode : 0;
  A.skip_out = y ? skip_out : 0;
  A.cpr = cpr;
  const int tmax = cpr <= 128 ? 256 : 512;
  A.lanes = tmax / cpr < 1 ? 1 : tmax / cpr;
  const int threads = (cpr * A.lanes + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(TILE_ROWS) * ld * sizeof(float);
  if (smem > 100 * 1024) return RR_ERR_UNSUPPORTED;
  const int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  int grid = num_sms() * (per_sm > 6 ? 6 : (per_sm < 1 ? 1 : per_sm));
  if (grid > g->n_mols) grid = g->n_mols;
#define RR_MOL_LAUNCH(B, F)                                                                                        \
  do {                                                                                                             \
    static PerDeviceOnce attr_set;                                                                                 \
    if (attr_set.need()) {                                                                                         \
      RR_CUDA(cudaFuncSetAttribute(k_mol_nbr<B, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));    \
      attr_set.mark();                                                                                             \
    }                                                                                                              \
    k_mol_nbr<B, F><<<grid, threads, smem, s>>>(A);                                                                \
  } while (0)
  if (op == 0) RR_MOL_LAUNCH(false, false);
  else if (y) RR_MOL_LAUNCH(true, true);
  else RR_MOL_LAUNCH(true, false);
#undef RR_MOL_LAUNCH
  RR_LAUNCH_CHECK("k_mol_nbr");
  return RR_OK;
}

}  // namespace rr
