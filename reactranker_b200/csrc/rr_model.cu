// Whole-model forward / backward of ReactionModel (models/base_model.py:150-171):
//   MPN(reactants), MPN(products)  ->  p - r  ->  MPNDiff over the product graph
//   -> scope-mean readout || add_features -> FFN -> head
// sequenced on one stream out of a caller-provided workspace.  Nothing is allocated here.
//
// Parameters arrive in the reference's state_dict layout and are re-packed once per call
// into zero-padded, 16-byte-aligned matrices; concatenated inputs of the reference
// ([f_atoms || a_message], [nei_message || nei_f_bonds], [diff || a_message]) are never
// materialised: their weight matrices are split column-wise and the GEMM accumulates two
// sources.  Gradients are produced in the padded layout and un-packed at the end.
#include "rr_common.cuh"

namespace rr {

static inline int pad16(int x) { return (x + 15) / 16 * 16; }
int padded(int w) { return pad16(w); }

constexpr int kMaxDepth = 16;

// ---- packed parameter layout (float offsets) ---------------------------------------------
struct PackedLayout {
  int h, hp, vp, F;
  size_t enc_Wi, enc_bi, enc_Wh, enc_bh, enc_Wo_a, enc_Wo_m, enc_bo;
  size_t dif_Wi, dif_bi, dif_Wh_m, dif_Wh_f, dif_bh, dif_Wo_d, dif_Wo_m, dif_bo;
  size_t ffn_W[RR_MAX_FFN], ffn_b[RR_MAX_FFN];
  int ffn_in[RR_MAX_FFN], ffn_out[RR_MAX_FFN];  // padded dims
  // transposed copies [in_pad, out_pad] of every weight that has a dgrad: K-major B operand of dX = dZ W on tcgen05
  size_t enc_Wh_T, enc_Wo_m_T, dif_Wi_T, dif_Wh_m_T, dif_Wo_d_T, dif_Wo_m_T, ffn_W_T[RR_MAX_FFN];
  size_t total;       // floats, transposed copies included
  size_t grad_total;  // floats of the gradient image (no transposed copies)
};

static PackedLayout make_packed(const rr_model_cfg& c) {
  PackedLayout L{};
  L.h = c.hidden;
  L.hp = pad16(c.hidden);
  L.vp = pad16(c.hidden + c.add_features);
  L.F = c.ffn_depth;
  size_t o = 0;
  auto take = [&](size_t n) {
    size_t r = o;
    o += (n + 63) / 64 * 64;
    return r;
  };
  const size_t hp = L.hp;
  L.enc_Wi = take(hp * RR_FB_LD); L.enc_bi = take(hp);
  L.enc_Wh = take(hp * hp);       L.enc_bh = take(hp);
  L.enc_Wo_a = take(hp * RR_FA_LD); L.enc_Wo_m = take(hp * hp); L.enc_bo = take(hp);
  L.dif_Wi = take(hp * hp);       L.dif_bi = take(hp);
  L.dif_Wh_m = take(hp * hp);     L.dif_Wh_f = take(hp * RR_FB_LD); L.dif_bh = take(hp);
  L.dif_Wo_d = take(hp * hp);     L.dif_Wo_m = take(hp * hp);       L.dif_bo = take(hp);
  for (int l = 0; l < L.F; ++l) {
    L.ffn_in[l] = (l == 0) ? L.vp : L.hp;
    L.ffn_out[l] = (l == L.F - 1) ? RR_OUT_LD : L.hp;
    L.ffn_W[l] = take(static_cast<size_t>(L.ffn_in[l]) * L.ffn_out[l]);
    L.ffn_b[l] = take(L.ffn_out[l]);
  }
  L.grad_total = o;
  L.enc_Wh_T = take(hp * hp);
  L.enc_Wo_m_T = take(hp * hp);
  L.dif_Wi_T = take(hp * hp);
  L.dif_Wh_m_T = take(hp * hp);
  L.dif_Wo_d_T = take(hp * hp);
  L.dif_Wo_m_T = take(hp * hp);
  for (int l = 0; l < L.F; ++l) L.ffn_W_T[l] = take(static_cast<size_t>(L.ffn_in[l]) * L.ffn_out[l]);
  L.total = o;
  return L;
}

struct PackEntry {
  float* ref;        // state_dict-layout tensor (source when packing, destination when un-packing)
  long long packed;  // float offset in the packed buffer
  int rows, cols, ref_ld, ref_col0, packed_ld;
  int transpose;     // packed[c][r] = ref[r][c]
};
constexpr int kMaxEntries = 48;
struct PackTable {
  PackEntry e[kMaxEntries];
  int n;
};

static PackTable make_table(const rr_model_cfg& c, const PackedLayout& L, const rr_params& w, bool with_transposes) {
  PackTable T{};
  const int h = c.hidden, hp = L.hp;
  auto add = [&](float* ref, size_t off, int rows, int cols, int ref_ld, int col0, int pld) {
    if (ref == nullptr) return;
    T.e[T.n++] = PackEntry{ref, static_cast<long long>(off), rows, cols, ref_ld, col0, pld, 0};
  };
  auto add_t = [&](float* ref, size_t off, int rows, int cols, int ref_ld, int col0, int pld) {
    if (ref == nullptr || !with_transposes) return;
    T.e[T.n++] = PackEntry{ref, static_cast<long long>(off), rows, cols, ref_ld, col0, pld, 1};
  };
  add(w.enc_Wi, L.enc_Wi, h, RR_FBOND_TOTAL, RR_FBOND_TOTAL, 0, RR_FB_LD);
  add(w.enc_bi, L.enc_bi, 1, h, h, 0, hp);
  add(w.enc_Wh, L.enc_Wh, h, h, h, 0, hp);
  add(w.enc_bh, L.enc_bh, 1, h, h, 0, hp);
  add(w.enc_Wo, L.enc_Wo_a, h, RR_ATOM_FDIM, RR_ATOM_FDIM + h, 0, RR_FA_LD);
  add(w.enc_Wo, L.enc_Wo_m, h, h, RR_ATOM_FDIM + h, RR_ATOM_FDIM, hp);
  add(w.enc_bo, L.enc_bo, 1, h, h, 0, hp);
  add(w.dif_Wi, L.dif_Wi, h, h, h, 0, hp);
  add(w.dif_bi, L.dif_bi, 1, h, h, 0, hp);
  add(w.dif_Wh, L.dif_Wh_m, h, h, h + RR_FBOND_TOTAL, 0, hp);
  add(w.dif_Wh, L.dif_Wh_f, h, RR_FBOND_TOTAL, h + RR_FBOND_TOTAL, h, RR_FB_LD);
  add(w.dif_bh, L.dif_bh, 1, h, h, 0, hp);
  add(w.dif_Wo, L.dif_Wo_d, h, h, 2 * h, 0, hp);
  add(w.dif_Wo, L.dif_Wo_m, h, h, 2 * h, h, hp);
  add(w.dif_bo, L.dif_bo, 1, h, h, 0, hp);
  for (int l = 0; l < L.F; ++l) {
    const int in = (l == 0) ? h + c.add_features : h;
    const int out = (l == L.F - 1) ? c.task_num : h;
    add(w.ffn_W[l], L.ffn_W[l], out, in, in, 0, L.ffn_in[l]);
    add(w.ffn_b[l], L.ffn_b[l], 1, out, out, 0, L.ffn_out[l]);
    add_t(w.ffn_W[l], L.ffn_W_T[l], out, in, in, 0, L.ffn_out[l]);
  }
  add_t(w.enc_Wh, L.enc_Wh_T, h, h, h, 0, hp);
  add_t(w.enc_Wo, L.enc_Wo_m_T, h, h, RR_ATOM_FDIM + h, RR_ATOM_FDIM, hp);
  add_t(w.dif_Wi, L.dif_Wi_T, h, h, h, 0, hp);
  add_t(w.dif_Wh, L.dif_Wh_m_T, h, h, h + RR_FBOND_TOTAL, 0, hp);
  add_t(w.dif_Wo, L.dif_Wo_d_T, h, h, 2 * h, 0, hp);
  add_t(w.dif_Wo, L.dif_Wo_m_T, h, h, 2 * h, h, hp);
  return T;
}

// direction 0: packed <- ref (pack);  1: ref <- packed (un-pack gradients)
// Packing also writes the TF32 split of every value (hi = nearest TF32, lo = v - hi) into the images hi_off / lo_off floats further on:
// the tcgen05 GEMMs then TMA-load both halves of the B operand instead of splitting the same weight tile once per row tile.
// The same values as bf16 pairs (hi = bf16(v), lo = bf16(v - hi)) go to two uint16 images (bhi / blo, same element offsets) for the
// backward GEMMs' 3 x bf16 split.
__device__ __forceinline__ uint16_t bf16_rn(float v) {
  uint16_t r;
  asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return r;
}
__global__ void k_pack(PackTable T, float* __restrict__ packed, int direction, long long hi_off, long long lo_off, uint16_t* __restrict__ bhi,
                       uint16_t* __restrict__ blo) {
  const PackEntry e = T.e[blockIdx.y];
  const long long n = static_cast<long long>(e.rows) * e.cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / e.cols), c = static_cast<int>(i - static_cast<long long>(r) * e.cols);
    float* pp = packed + e.packed + (e.transpose ? static_cast<long long>(c) * e.packed_ld + r : static_cast<long long>(r) * e.packed_ld + c);
    float* rp = e.ref + static_cast<long long>(r) * e.ref_ld + e.ref_col0 + c;
    if (direction == 0) {
      const float v = *rp;
      const float hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
      *pp = v;
      pp[hi_off] = hi;
      pp[lo_off] = v - hi;
      const uint16_t b1 = bf16_rn(v);
      bhi[pp - packed] = b1;
      blo[pp - packed] = bf16_rn(v - __uint_as_float(static_cast<uint32_t>(b1) << 16));
    } else {
      *rp = *pp;
    }
  }
}

// ---- head (base_model.py:59-108) -----------------------------------------------------------
__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // nn.Softplus(beta=1, threshold=20)
__global__ void k_head_fwd(int N, int task, int head, const float* __restrict__ z, float* __restrict__ out) {
  const int half = task >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * task; i += gridDim.x * blockDim.x) {
    const int r = i / task, c = i - r * task;
    const float* zr = z + static_cast<size_t>(r) * RR_OUT_LD;
    float v;
    if (head == RR_HEAD_RAW) v = zr[c];
    else if (head == RR_HEAD_SOFTPLUS) v = softplus_t(zr[c]);
    else if (head == RR_HEAD_SOFTPLUS_P1) v = softplus_t(zr[c]) + 1.f;
    else if (head == RR_HEAD_NIG) {  // stack((mu, lambda, alpha, beta), dim=2).view(...): column c = 4 j + k reads quarter k, entry j
      const int j = c >> 2, k = c & 3, quarter = task >> 2;
      const float zc = zr[k * quarter + j];
      v = k == 0 ? zc : (k == 2 ? softplus_t(zc) + 1e-6f + 1.f : softplus_t(zc) + 1e-6f);
    } else {  // stack((first half, f(second half)), dim=2).view(...)  -> interleaved columns
      const int j = c >> 1;
      if ((c & 1) == 0) v = head == RR_HEAD_LOGNORM ? softplus_t(zr[j]) + 1e-6f : zr[j];
      else v = softplus_t(zr[half + j]) + (head == RR_HEAD_GAUSS_SOFTPLUS ? 0.f : 1e-6f);
    }
    out[i] = v;
  }
}
__global__ void k_head_bwd(int N, int task, int head, const float* __restrict__ z, const float* __restrict__ dout, float* __restrict__ dz) {
  const int half = task >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * RR_OUT_LD; i += gridDim.x * blockDim.x) {
    const int r = i / RR_OUT_LD, c = i - r * RR_OUT_LD;
    float v = 0.f;
    if (c < task) {
      const float zc = z[i];
      const float sg = zc > 20.f ? 1.f : 1.f / (1.f + expf(-zc));
      const float* dr = dout + static_cast<size_t>(r) * task;
      if (head == RR_HEAD_RAW) v = dr[c];
      else if (head == RR_HEAD_SOFTPLUS || head == RR_HEAD_SOFTPLUS_P1) v = dr[c] * sg;
      else if (head == RR_HEAD_NIG) {
        const int quarter = task >> 2, k = c / quarter, j = c - k * quarter;
        v = dr[4 * j + k] * (k == 0 ? 1.f : sg);
      } else v = (c < half) ? dr[2 * c] * (head == RR_HEAD_LOGNORM ? sg : 1.f) : dr[2 * (c - half) + 1] * sg;
    }
    dz[i] = v;
  }
}

// ---- padding rows in exact arithmetic -------------------------------------------------------
// A segment's padding bond / padding atom row is gathered by EVERY padded neighbour slot of the segment (featurization.py:281-286), so in
// the backward pass it collects the gradient of all of them: at the c5 batch one such row weighs as much as ~10^5 ordinary rows.  Its ReLU
// masks must therefore be decided as the reference's fp32 run decides them.  The tensor-core forward (3 x TF32, RZ accumulation) leaves
// 2-5e-6 of relative error on a pre-activation, which flips the mask of an entry that lies that close to zero; on an ordinary row such a
// flip moves the gradient by one row's share, on a padding row by a macroscopic amount.  After every tensor-core forward GEMM whose output
// feeds a ReLU, the n_segments padding rows are therefore recomputed here from the unsplit fp32 weights with fp64 accumulation (one
// rounding, tighter than any fp32 GEMM) and overwritten, epilogue (bias, residual, ReLU, the same Philox dropout mask) included.
struct PadFix {
  const int* rows;
  int n_rows, n;
  const float *X1, *W1, *X2, *W2, *bias, *resid;
  int ldx1, k1, ldx2, k2, ldr;
  float* Y;
  int ldy, relu;
  float p, inv_keep;
  uint64_t seed, stream_id;
};
// One warp per output column (grid = rows x ceil(n / 8), 8 warps per block): the k loop is unrolled into independent loads and four
// fp64 partial sums, so a launch is a couple of microseconds (the first version, one warp per ~5 columns with a serial chain, took 20).
__device__ __forceinline__ double pad_dot(const float* __restrict__ x, const float* __restrict__ w, int k, int lane) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = lane;
  for (; i + 96 < k; i += 128) {
    const float x0 = x[i], x1 = x[i + 32], x2 = x[i + 64], x3 = x[i + 96];
    const float w0 = __ldg(w + i), w1 = __ldg(w + i + 32), w2 = __ldg(w + i + 64), w3 = __ldg(w + i + 96);
    a0 = fma(static_cast<double>(x0), static_cast<double>(w0), a0);
    a1 = fma(static_cast<double>(x1), static_cast<double>(w1), a1);
    a2 = fma(static_cast<double>(x2), static_cast<double>(w2), a2);
    a3 = fma(static_cast<double>(x3), static_cast<double>(w3), a3);
  }
  for (; i < k; i += 32) a0 = fma(static_cast<double>(x[i]), static_cast<double>(__ldg(w + i)), a0);
  return (a0 + a1) + (a2 + a3);
}
__global__ void __launch_bounds__(256) k_pad_rows_linear(PadFix a) {
  const size_t row = static_cast<size_t>(__ldg(a.rows + blockIdx.x));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 8 + warp;
  if (c >= a.n) return;
  double acc = pad_dot(a.X1 + row * a.ldx1, a.W1 + static_cast<size_t>(c) * a.k1, a.k1, lane);
  if (a.X2) acc += pad_dot(a.X2 + row * a.ldx2, a.W2 + static_cast<size_t>(c) * a.k2, a.k2, lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (a.bias) acc += static_cast<double>(a.bias[c]);
    if (a.resid) acc += static_cast<double>(a.resid[row * a.ldr + c]);
    float o = static_cast<float>(acc);
    if (a.relu) o = fmaxf(o, 0.f);
    if (a.p > 0.f) {   // element (c & 3) of the 4-wide Philox draw the GEMM epilogue makes for this chunk (dropout4)
      const uint4 r = philox4x32(a.seed, a.stream_id, (row * a.ldy + c) >> 2);
      const uint32_t thr = static_cast<uint32_t>(fminf(a.p, 1.f) * 4294967295.f);
      const uint32_t rv = (c & 3) == 0 ? r.x : ((c & 3) == 1 ? r.y : ((c & 3) == 2 ? r.z : r.w));
      o = rv >= thr ? o * a.inv_keep : 0.f;
    }
    a.Y[row * a.ldy + c] = o;
  }
}
// same argument order as linear_fwd; W1 / W2 are the raw (unsplit) packed weights
static int pad_rows_linear(const int* rows, int n_rows, int n, const float* X1, int ldx1, const float* W1, int k1, const float* X2, int ldx2, const float* W2,
                           int k2, const float* bias, const float* resid, int ldr, float* Y, int ldy, int flags, float p, uint64_t seed,
                           uint64_t stream_id, cudaStream_t s) {
  if (n_rows <= 0 || rows == nullptr) return RR_OK;
  ProfScope prof_scope(KC_MISC, s);
  PadFix a{};
  a.rows = rows;
  a.n_rows = n_rows;
  a.n = n;
  a.X1 = X1; a.ldx1 = ldx1; a.W1 = W1; a.k1 = k1;
  a.X2 = (X2 && k2 > 0) ? X2 : nullptr; a.ldx2 = ldx2; a.W2 = W2; a.k2 = k2;
  a.bias = bias; a.resid = resid; a.ldr = ldr;
  a.Y = Y; a.ldy = ldy;
  a.relu = flags & 1;
  a.p = (flags & 2) ? p : 0.f;
  a.inv_keep = a.p > 0.f ? 1.f / (1.f - a.p) : 1.f;
  a.seed = seed;
  a.stream_id = stream_id;
  k_pad_rows_linear<<<dim3(n_rows, (n + 7) / 8), 256, 0, s>>>(a);
  RR_LAUNCH_CHECK("k_pad_rows_linear");
  return RR_OK;
}

// ---- joint graph ------------------------------------------------------------------------------
// Index tables of the concatenation [reactant batch | product batch] in joint row space: the reactant part is copied, the product part
// is shifted by the reactant batch's row counts (valid slots only: an unused slot stays 0 and is never read, its multiplicity lives in
// rr_atom_meta).  Every segment keeps its own padding rows and max_num_bonds, so the joint launch computes exactly what two launches did.
__global__ void k_join_tables(rr_graph r, rr_graph p, rr_graph J) {
  const int A_r = r.n_atoms, B_r = r.n_bonds;
  int4* meta = reinterpret_cast<int4*>(const_cast<rr_atom_meta*>(J.a_meta));
  int* a2b = const_cast<int*>(J.a2b);
  int* a2r = const_cast<int*>(J.a2b_rev);
  int* a2a = const_cast<int*>(J.a2a);
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < J.n_atoms; a += gridDim.x * blockDim.x) {
    const bool second = a >= A_r;
    const rr_graph& g = second ? p : r;
    const int la = second ? a - A_r : a, ob = second ? B_r : 0, oa = second ? A_r : 0;
    int4 m = __ldg(reinterpret_cast<const int4*>(g.a_meta) + la);
    const int deg = m.x & 0xff;
    m.z += ob;
    m.w += oa;
    meta[a] = m;
    const size_t src = static_cast<size_t>(la) * g.wmax, dst = static_cast<size_t>(a) * J.wmax;
    for (int k = 0; k < J.wmax; ++k) {
      const bool v = k < deg;
      a2b[dst + k] = v ? __ldg(g.a2b + src + k) + ob : 0;
      a2r[dst + k] = v ? __ldg(g.a2b_rev + src + k) + ob : 0;
      a2a[dst + k] = v ? __ldg(g.a2a + src + k) + oa : 0;
    }
  }
  for (int sgm = blockIdx.x * blockDim.x + threadIdx.x; sgm < J.n_segments; sgm += gridDim.x * blockDim.x) {
    const bool second = sgm >= r.n_segments;
    const int ls = second ? sgm - r.n_segments : sgm;
    const_cast<int*>(J.pad_bonds)[sgm] = second ? __ldg(p.pad_bonds + ls) + B_r : __ldg(r.pad_bonds + ls);
    const_cast<int*>(J.pad_atoms)[sgm] = second ? __ldg(p.pad_atoms + ls) + A_r : __ldg(r.pad_atoms + ls);
  }
}

// ---- workspace layout -----------------------------------------------------------------------
struct EncBufs {
  float* inp;
  float* pre[kMaxDepth];
  float* m[kMaxDepth + 1];  // m[1..T]
  float* am;
  float* hid;
};
struct Workspace {
  PackedLayout L;
  float* packed;      // raw | TF32 hi image | TF32 lo image, L.total floats each
  ptrdiff_t hi_off, lo_off;
  uint16_t *bhi, *blo; // bf16 (hi, lo) images, L.total elements each
  float* dpacked;
  // The shared-weight encoder runs ONCE over the reactant and product batches together (mpn.py:61-108 is row-wise apart from the gathers,
  // and those stay inside a segment): every encoder buffer holds the reactant rows followed by the product rows.
  EncBufs enc[2];     // views: 0 = reactant rows, 1 = product rows of the joint buffers below
  EncBufs j;          // joint buffers [r rows | p rows]
  rr_graph J;         // the joint graph: index tables in joint row space (built by k_join_tables), features contiguous [r | p]
  bool copy_feats;    // r / p feature arrays are not adjacent in memory: copied into j_fa / j_fb once per forward
  float *j_fa, *j_fb;
  float *d, *inp2, *nf, *am2, *hid2, *vec, *zout;
  float* nm[kMaxDepth];
  float* m2[kMaxDepth + 1];
  float* x[RR_MAX_FFN];
  // backward scratch
  float *gB1, *gB2, *dinp, *gA1, *gA2, *gA3, *dD, *dI2, *dvec, *gN1, *gN2, *dzout;
  size_t bytes;
};

static int carve(const rr_model_cfg& c, const rr_graph& r, const rr_graph& p, void* base, Workspace* W) {
  RR_REQUIRE(c.hidden > 0 && c.hidden <= 4096, "hidden %d out of range", c.hidden);
  RR_REQUIRE(c.depth >= 1 && c.depth <= kMaxDepth && c.diff_depth >= 0 && c.diff_depth <= kMaxDepth, "depth %d / diff_depth %d out of range", c.depth, c.diff_depth);
  if (c.diff_depth < 1) return fail(RR_ERR_UNSUPPORTED, "mpnn_diff_depth = 0 (mpn.py:220-221) is not built");
  RR_REQUIRE(c.ffn_depth >= 1 && c.ffn_depth <= RR_MAX_FFN, "ffn_depth %d out of range", c.ffn_depth);
  RR_REQUIRE(c.task_num >= 1 && c.task_num <= RR_OUT_LD, "task_num %d out of range", c.task_num);
  RR_REQUIRE(c.add_features >= 0, "add_features < 0");
  RR_REQUIRE(c.r_atom_map != nullptr || (r.n_atoms == p.n_atoms && r.n_mols == p.n_mols),
             "reactant and product batches must have identical atom rows (p - r is atom-wise, base_model.py:168): %d vs %d", r.n_atoms, p.n_atoms);
  RR_REQUIRE(c.r_atom_map == nullptr || !(c.training && c.dropout > 0.f),
             "r_atom_map (de-duplicated reactants) is exact only without dropout: pass NULL when training with dropout > 0");
  W->L = make_packed(c);
  const size_t hp = W->L.hp, vp = W->L.vp;
  const size_t A = p.n_atoms, N = p.n_mols;
  const int T = c.depth - 1, Td = c.diff_depth - 1;
  char* cur = static_cast<char*>(base);
  size_t used = 0;
  auto take = [&](size_t floats) {
    float* ptr = base ? reinterpret_cast<float*>(cur + used) : nullptr;
    used += (floats * sizeof(float) + 255) / 256 * 256;
    return ptr;
  };
  W->packed = take(3 * W->L.total);
  W->hi_off = static_cast<ptrdiff_t>(W->L.total);
  W->lo_off = static_cast<ptrdiff_t>(2 * W->L.total);
  W->bhi = reinterpret_cast<uint16_t*>(take((W->L.total + 1) / 2));
  W->blo = reinterpret_cast<uint16_t*>(take((W->L.total + 1) / 2));
  W->dpacked = take(W->L.grad_total);
  const size_t A_r = r.n_atoms, B_r = r.n_bonds, A_J = A_r + p.n_atoms, B_J = B_r + p.n_bonds;
  {
    EncBufs& e = W->j;
    e.inp = take(B_J * hp);
    for (int t = 0; t < T; ++t) e.pre[t] = take(B_J * hp);
    for (int t = 1; t <= T; ++t) e.m[t] = take(B_J * hp);
    e.am = take(A_J * hp);
    e.hid = take(A_J * hp);
    auto part = [&](float* q, size_t rows) { return q ? q + rows * hp : nullptr; };
    W->enc[0] = e;
    EncBufs& q = W->enc[1];
    q.inp = part(e.inp, B_r);
    for (int t = 0; t < T; ++t) q.pre[t] = part(e.pre[t], B_r);
    for (int t = 1; t <= T; ++t) q.m[t] = part(e.m[t], B_r);
    q.am = part(e.am, A_r);
    q.hid = part(e.hid, A_r);
  }
  {   // joint graph tables (int32) and, when the two batches' features are not adjacent in memory, a contiguous copy of them
    rr_graph& J = W->J;
    J = rr_graph{};
    J.n_atoms = static_cast<int>(A_J);
    J.n_bonds = static_cast<int>(B_J);
    J.n_mols = r.n_mols + p.n_mols;
    J.wmax = r.wmax > p.wmax ? r.wmax : p.wmax;
    J.n_segments = r.n_segments + p.n_segments;
    J.a_meta = reinterpret_cast<const rr_atom_meta*>(take(A_J * 4));
    J.a2b = reinterpret_cast<const int*>(take(A_J * J.wmax));
    J.a2b_rev = reinterpret_cast<const int*>(take(A_J * J.wmax));
    J.a2a = reinterpret_cast<const int*>(take(A_J * J.wmax));
    J.pad_bonds = reinterpret_cast<const int*>(take(J.n_segments));
    J.pad_atoms = reinterpret_cast<const int*>(take(J.n_segments));
    W->copy_feats = !(p.f_atoms == r.f_atoms + A_r * RR_FA_LD && p.f_bonds == r.f_bonds + B_r * RR_FB_LD);
    W->j_fa = W->copy_feats ? take(A_J * RR_FA_LD) : nullptr;
    W->j_fb = W->copy_feats ? take(B_J * RR_FB_LD) : nullptr;
    J.f_atoms = W->copy_feats ? W->j_fa : r.f_atoms;
    J.f_bonds = W->copy_feats ? W->j_fb : r.f_bonds;
  }
  W->d = take(A * hp);
  W->inp2 = take(A * hp);
  W->nf = take(A * RR_FB_LD);
  for (int t = 0; t < Td; ++t) W->nm[t] = take(A * hp);
  for (int t = 1; t <= Td; ++t) W->m2[t] = take(A * hp);
  W->am2 = take(A * hp);
  W->hid2 = take(A * hp);
  W->vec = take(N * vp);
  for (int l = 0; l + 1 < c.ffn_depth; ++l) W->x[l] = take(N * hp);
  W->zout = take(N * RR_OUT_LD);
  W->gB1 = take(B_J * hp);
  W->gB2 = take(B_J * hp);
  W->dinp = take(B_J * hp);
  W->gA1 = take(A_J * hp);
  W->gA2 = take(A_J * hp);
  W->gA3 = take(A * hp);
  W->dD = take(A * hp);
  W->dI2 = take(A * hp);
  W->dvec = take(N * vp);
  W->gN1 = take(N * hp);
  W->gN2 = take(N * hp);
  W->dzout = take(N * RR_OUT_LD);
  W->bytes = used;
  return RR_OK;
}

// dX[M, k] (+)= dZ[M, n] W[n, k]: tensor cores read the transposed copy Wt[k, n] as a K-major operand, the SIMT path reads W
static int dgrad(int M, int n, int k, const float* dZ, int lddz, const float* W, int ldw, const float* Wt, int ldwt, float* dX, int lddx,
                 int accumulate, cudaStream_t s, ptrdiff_t hi_off, ptrdiff_t lo_off, const float* packed, const uint16_t* bhi, const uint16_t* blo) {
  const bool tc = g_gemm_mode.load() == 1 || g_gemm_mode.load() == 3;
  if (tc && g_bwd_bf16.load() && tc_linear_bf16_supported(M, k, n, lddz, ldwt))
    return tc_linear_bf16(M, k, dZ, lddz, bhi + (Wt - packed), blo + (Wt - packed), ldwt, n, dX, lddx, accumulate, KC_GEMM_DGRAD, s);
  if (tc && tc_supported(M, k, n, 0, lddz, 0))
    return tc_linear(M, k, dZ, lddz, Wt + hi_off, ldwt, n, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, 0, dX, lddx, 0, accumulate, 0.f, 0, 0, KC_GEMM_DGRAD, s,
                     Wt + lo_off, nullptr);
  return linear_dgrad(M, n, k, dZ, lddz, W, ldw, dX, lddx, accumulate, s);
}

// byte offset of a named forward buffer inside the workspace (tests compare per-layer activations with the oracle)
long long model_buffer_offset(const rr_model_cfg* c, const rr_graph* r, const rr_graph* p, const char* name) {
  Workspace W;
  char* base = reinterpret_cast<char*>(static_cast<uintptr_t>(4096));
  if (carve(*c, *r, *p, base, &W) != RR_OK) return -1;
  auto off = [&](const float* ptr) { return static_cast<long long>(reinterpret_cast<const char*>(ptr) - base); };
  char buf[64];
  for (int k = 0; k < 2; ++k) {
    snprintf(buf, sizeof(buf), "enc%d.inp", k);
    if (!strcmp(name, buf)) return off(W.enc[k].inp);
    snprintf(buf, sizeof(buf), "enc%d.am", k);
    if (!strcmp(name, buf)) return off(W.enc[k].am);
    snprintf(buf, sizeof(buf), "enc%d.hid", k);
    if (!strcmp(name, buf)) return off(W.enc[k].hid);
    for (int t = 0; t < c->depth - 1; ++t) {
      snprintf(buf, sizeof(buf), "enc%d.pre%d", k, t);
      if (!strcmp(name, buf)) return off(W.enc[k].pre[t]);
      snprintf(buf, sizeof(buf), "enc%d.m%d", k, t + 1);
      if (!strcmp(name, buf)) return off(W.enc[k].m[t + 1]);
    }
  }
  for (int t = 0; t < c->diff_depth - 1; ++t) {
    snprintf(buf, sizeof(buf), "nm%d", t);
    if (!strcmp(name, buf)) return off(W.nm[t]);
    snprintf(buf, sizeof(buf), "m2_%d", t + 1);
    if (!strcmp(name, buf)) return off(W.m2[t + 1]);
  }
  if (!strcmp(name, "d")) return off(W.d);
  if (!strcmp(name, "inp2")) return off(W.inp2);
  if (!strcmp(name, "nf")) return off(W.nf);
  if (!strcmp(name, "am2")) return off(W.am2);
  if (!strcmp(name, "hid2")) return off(W.hid2);
  if (!strcmp(name, "vec")) return off(W.vec);
  if (!strcmp(name, "zout")) return off(W.zout);
  for (int l = 0; l + 1 < c->ffn_depth; ++l) {
    snprintf(buf, sizeof(buf), "x%d", l);
    if (!strcmp(name, buf)) return off(W.x[l]);
  }
  fail(RR_ERR_INVALID, "unknown buffer name %s", name);
  return -1;
}

long long model_workspace_bytes(const rr_model_cfg* c, const rr_graph* r, const rr_graph* p) {
  Workspace W;
  if (carve(*c, *r, *p, nullptr, &W) != RR_OK) return -1;
  return static_cast<long long>(W.bytes);
}

static int check_model_args(const rr_model_cfg* c, const rr_params* w, const rr_graph* r, const rr_graph* p, void* ws, long long ws_bytes, Workspace* W) {
  RR_REQUIRE(c && w && r && p && ws, "model: NULL argument");
  RR_REQUIRE(aligned16(ws), "workspace must be 16-byte aligned");
  RR_TRY(carve(*c, *r, *p, ws, W));
  if (static_cast<long long>(W->bytes) > ws_bytes) return fail(RR_ERR_WORKSPACE, "workspace %lld bytes < required %zu", ws_bytes, W->bytes);
  RR_REQUIRE(w->enc_Wi && w->enc_Wo && w->enc_bo && w->dif_Wi && w->dif_Wo && w->dif_bo, "model: missing parameter pointer");
  RR_REQUIRE(c->depth == 1 || w->enc_Wh, "model: encoder.W_h missing for depth > 1");
  RR_REQUIRE(c->diff_depth == 1 || w->dif_Wh, "model: diff_encoder.W_h missing for diff_depth > 1");
  for (int l = 0; l < c->ffn_depth; ++l) RR_REQUIRE(w->ffn_W[l], "model: ffn weight %d missing", l);
  RR_REQUIRE(c->dropout >= 0.f && c->dropout < 1.f, "dropout must be in [0,1)");
  RR_REQUIRE(r->f_bonds && r->f_atoms && p->f_bonds && p->f_atoms, "model: graph features missing");
  RR_REQUIRE(c->head >= RR_HEAD_RAW && c->head <= RR_HEAD_NIG, "unknown head %d", c->head);
  if (c->head == RR_HEAD_NIG) RR_REQUIRE((c->task_num & 3) == 0, "the NIG head needs task_num %% 4 == 0");
  else if (c->head != RR_HEAD_RAW && c->head != RR_HEAD_SOFTPLUS && c->head != RR_HEAD_SOFTPLUS_P1)
    RR_REQUIRE((c->task_num & 1) == 0, "two-parameter heads need an even task_num");
  return RR_OK;
}

static int pack_params(const rr_model_cfg& c, const Workspace& W, const rr_params& w, cudaStream_t s) {
  RR_CUDA(cudaMemsetAsync(W.packed, 0, 3 * W.L.total * sizeof(float), s));
  RR_CUDA(cudaMemsetAsync(W.bhi, 0, W.L.total * sizeof(uint16_t), s));
  RR_CUDA(cudaMemsetAsync(W.blo, 0, W.L.total * sizeof(uint16_t), s));
  PackTable T = make_table(c, W.L, w, true);
  k_pack<<<dim3(32, T.n), 256, 0, s>>>(T, W.packed, 0, W.hi_off, W.lo_off, W.bhi, W.blo);
  RR_LAUNCH_CHECK("k_pack");
  return RR_OK;
}

// Fill the joint graph of this call's workspace: write the index tables; the features are copied only when the two batches
// do not already sit next to each other in memory (DeviceGraph.assemble_pair lays them out adjacent, so the product path copies nothing).
static int build_joint(const Workspace& W, const rr_graph& r, const rr_graph& p, cudaStream_t s) {
  ProfScope prof_scope(KC_MISC, s);
  const int n = W.J.n_atoms;
  int blocks = (n + 255) / 256;
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  if (blocks < 1) blocks = 1;
  k_join_tables<<<blocks, 256, 0, s>>>(r, p, W.J);
  RR_LAUNCH_CHECK("k_join_tables");
  if (W.copy_feats) {
    RR_CUDA(cudaMemcpyAsync(W.j_fa, r.f_atoms, static_cast<size_t>(r.n_atoms) * RR_FA_LD * 4, cudaMemcpyDeviceToDevice, s));
    RR_CUDA(cudaMemcpyAsync(W.j_fa + static_cast<size_t>(r.n_atoms) * RR_FA_LD, p.f_atoms, static_cast<size_t>(p.n_atoms) * RR_FA_LD * 4, cudaMemcpyDeviceToDevice, s));
    RR_CUDA(cudaMemcpyAsync(W.j_fb, r.f_bonds, static_cast<size_t>(r.n_bonds) * RR_FB_LD * 4, cudaMemcpyDeviceToDevice, s));
    RR_CUDA(cudaMemcpyAsync(W.j_fb + static_cast<size_t>(r.n_bonds) * RR_FB_LD, p.f_bonds, static_cast<size_t>(p.n_bonds) * RR_FB_LD * 4, cudaMemcpyDeviceToDevice, s));
  }
  return RR_OK;
}

int model_forward(const rr_model_cfg* c, const rr_params* w, const rr_graph* r, const rr_graph* p, const float* addf, float* scores,
                  void* ws, long long ws_bytes, cudaStream_t s) {
  Workspace W;
  RR_TRY(check_model_args(c, w, r, p, ws, ws_bytes, &W));
  RR_REQUIRE(scores != nullptr, "scores is NULL");
  RR_REQUIRE(c->add_features == 0 || addf != nullptr, "add_features is NULL but add_features_dim = %d", c->add_features);
  const PackedLayout& L = W.L;
  const int hp = L.hp, vp = L.vp;
  const int T = c->depth - 1, Td = c->diff_depth - 1;
  const float pdrop = c->training ? c->dropout : 0.f;
  const int act = 1 | (pdrop > 0.f ? 2 : 0);
  const float* P = W.packed;
  uint64_t sid = 1;
  RR_TRY(pack_params(*c, W, *w, s));

  // Y = act(X1 W1^T + X2 W2^T + bias + resid), then the segment padding rows once more in exact arithmetic when the GEMM ran on tensor cores
  const int gmode = g_gemm_mode.load();
  const bool fix_pads = (gmode == 1 || gmode == 2);
  auto lin = [&](const int* pad_rows, int n_pad, int M, int n, const float* X1, int ldx1, size_t W1, int k1, const float* X2, int ldx2, size_t W2, int k2,
                 size_t bias, const float* resid, float* Y, int ldy, int flags) -> int {
    const uint64_t id = sid++;
    RR_TRY(linear_fwd(M, n, X1, ldx1, P + W1, k1, X2, ldx2, X2 ? P + W2 : nullptr, k2, P + bias, resid, hp, Y, ldy, flags, pdrop, c->seed, id, s, W.hi_off,
                      W.lo_off, W.packed, W.bhi, W.blo));
    if (fix_pads && pad_rows)
      RR_TRY(pad_rows_linear(pad_rows, n_pad, n, X1, ldx1, P + W1, k1, X2, ldx2, X2 ? P + W2 : nullptr, k2, P + bias, resid, hp, Y, ldy, flags, pdrop, c->seed,
                             id, s));
    return RR_OK;
  };
  RR_TRY(build_joint(W, *r, *p, s));
  {  // mpn.py:61-108, once over [reactant rows | product rows]: the weights are shared (base_model.py:155-156)
    const rr_graph* g = &W.J;
    EncBufs& e = W.j;
    // f_bonds[pad] = 0: the padding rows of W_i's output are the bias exactly, on any GEMM path
    RR_TRY(lin(nullptr, 0, g->n_bonds, hp, g->f_bonds, RR_FB_LD, L.enc_Wi, RR_FB_LD, nullptr, 0, 0, 0, L.enc_bi, nullptr, e.inp, hp, 0));
    const float* src = e.inp;
    int relu_src = 1;
    for (int t = 0; t < T; ++t) {
      RR_TRY(bond_message_fwd(g, src, e.pre[t], hp, relu_src, s));
      RR_TRY(lin(g->pad_bonds, g->n_segments, g->n_bonds, hp, e.pre[t], hp, L.enc_Wh, hp, nullptr, 0, 0, 0, L.enc_bh, e.inp, e.m[t + 1], hp, act));
      src = e.m[t + 1];
      relu_src = 0;
    }
    RR_TRY(neighbor_sum_fwd(g, 0, src, e.am, hp, relu_src, s));
    RR_TRY(lin(g->pad_atoms, g->n_segments, g->n_atoms, hp, g->f_atoms, RR_FA_LD, L.enc_Wo_a, RR_FA_LD, e.am, hp, L.enc_Wo_m, hp, L.enc_bo, nullptr, e.hid, hp, act));
  }
  const int A = p->n_atoms;
  if (c->r_atom_map) RR_TRY(sub_gather(A, hp, W.enc[1].hid, W.enc[0].hid, c->r_atom_map, W.d, s));
  else RR_TRY(sub(static_cast<long long>(A) * hp, W.enc[1].hid, W.enc[0].hid, W.d, s));  // base_model.py:168

  // mpn.py:170-240 over the product graph
  RR_TRY(lin(p->pad_atoms, p->n_segments, A, hp, W.d, hp, L.dif_Wi, hp, nullptr, 0, 0, 0, L.dif_bi, nullptr, W.inp2, hp, 0));
  if (Td > 0) RR_TRY(neighbor_sum_fwd(p, 0, p->f_bonds, W.nf, RR_FB_LD, 0, s));
  {
    const float* src = W.inp2;
    int relu_src = 1;
    for (int t = 0; t < Td; ++t) {
      RR_TRY(neighbor_sum_fwd(p, 1, src, W.nm[t], hp, relu_src, s));
      RR_TRY(lin(p->pad_atoms, p->n_segments, A, hp, W.nm[t], hp, L.dif_Wh_m, hp, W.nf, RR_FB_LD, L.dif_Wh_f, RR_FB_LD, L.dif_bh, W.inp2, W.m2[t + 1], hp, act));
      src = W.m2[t + 1];
      relu_src = 0;
    }
    RR_TRY(neighbor_sum_fwd(p, 1, src, W.am2, hp, relu_src, s));
  }
  RR_TRY(lin(p->pad_atoms, p->n_segments, A, hp, W.d, hp, L.dif_Wo_d, hp, W.am2, hp, L.dif_Wo_m, hp, L.dif_bo, nullptr, W.hid2, hp, act));
  RR_TRY(readout_fwd(p, W.hid2, hp, c->hidden, addf, c->add_features, W.vec, vp, pdrop, c->seed, sid++, s));

  // base_model.py:40-60
  const int N = p->n_mols;
  const float* x = W.vec;
  int ldx = vp;
  for (int l = 0; l < c->ffn_depth; ++l) {
    const bool last = (l == c->ffn_depth - 1);
    float* y = last ? W.zout : W.x[l];
    RR_TRY(lin(nullptr, 0, N, L.ffn_out[l], x, ldx, L.ffn_W[l], L.ffn_in[l], nullptr, 0, 0, 0, L.ffn_b[l], nullptr, y, L.ffn_out[l], last ? 0 : act));
    x = y;
    ldx = L.ffn_out[l];
  }
  {
    const int n = N * c->task_num;
    k_head_fwd<<<(n + 255) / 256, 256, 0, s>>>(N, c->task_num, c->head, W.zout, scores);
    RR_LAUNCH_CHECK("k_head_fwd");
  }
  return RR_OK;
}

int model_backward(const rr_model_cfg* c, const rr_params* w, const rr_graph* r, const rr_graph* p, const float* dscores, rr_params* grads,
                   void* ws, long long ws_bytes, cudaStream_t s) {
  Workspace W;
  RR_TRY(check_model_args(c, w, r, p, ws, ws_bytes, &W));
  RR_REQUIRE(dscores && grads, "backward: NULL argument");
  const PackedLayout& L = W.L;
  const int hp = L.hp, vp = L.vp;
  const int T = c->depth - 1, Td = c->diff_depth - 1;
  const float pdrop = c->training ? c->dropout : 0.f;
  const float keep = pdrop > 0.f ? 1.f / (1.f - pdrop) : 1.f;
  const float* P = W.packed;  // still holds this step's packed weights
  float* G = W.dpacked;
  RR_CUDA(cudaMemsetAsync(G, 0, L.grad_total * sizeof(float), s));
  const int N = p->n_mols, A = p->n_atoms;

  {
    const int n = N * RR_OUT_LD;
    k_head_bwd<<<(n + 255) / 256, 256, 0, s>>>(N, c->task_num, c->head, W.zout, dscores, W.dzout);
    RR_LAUNCH_CHECK("k_head_bwd");
  }
  // FFN
  {
    const float* g = W.dzout;
    int ldg = RR_OUT_LD;
    float* scratch[2] = {W.gN1, W.gN2};
    for (int l = c->ffn_depth - 1; l >= 0; --l) {
      const float* x = (l == 0) ? W.vec : W.x[l - 1];
      RR_TRY(linear_wgrad(N, L.ffn_out[l], L.ffn_in[l], g, ldg, x, L.ffn_in[l], G + L.ffn_W[l], L.ffn_in[l], G + L.ffn_b[l], s));
      if (l > 0) {
        float* dx = scratch[l & 1];
        RR_TRY(dgrad(N, L.ffn_out[l], L.ffn_in[l], g, ldg, P + L.ffn_W[l], L.ffn_in[l], P + L.ffn_W_T[l], L.ffn_out[l], dx, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
        RR_TRY(relu_bwd(N, hp, dx, W.x[l - 1], keep, 0, dx, nullptr, 0, s));
        g = dx;
        ldg = hp;
      } else {
        RR_TRY(dgrad(N, L.ffn_out[l], L.ffn_in[l], g, ldg, P + L.ffn_W[l], L.ffn_in[l], P + L.ffn_W_T[l], L.ffn_out[l], W.dvec, vp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
      }
    }
  }
  // readout + W_o of the diff encoder
  RR_TRY(readout_bwd(p, W.dvec, vp, W.vec, W.hid2, W.gA1, hp, pdrop, s));
  RR_TRY(linear_wgrad(A, hp, hp, W.gA1, hp, W.d, hp, G + L.dif_Wo_d, hp, G + L.dif_bo, s));
  RR_TRY(linear_wgrad(A, hp, hp, W.gA1, hp, W.am2, hp, G + L.dif_Wo_m, hp, nullptr, s));
  RR_TRY(dgrad(A, hp, hp, W.gA1, hp, P + L.dif_Wo_d, hp, P + L.dif_Wo_d_T, hp, W.dD, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
  RR_TRY(dgrad(A, hp, hp, W.gA1, hp, P + L.dif_Wo_m, hp, P + L.dif_Wo_m_T, hp, W.gA2, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
  // Every gather backward is followed by exactly one ReLU (+ dropout) backward: the mask of the message it produced the gradient of.
  // Both run as one kernel (rr_mp_pipe.cu): t_next >= 1 -> mask of m2^{t_next}, result kept (gA1) and summed into dI2;
  // t_next == 0 -> mask [inp2 > 0] of m2^0 = relu(inp2), only the sum into dI2 is wanted.
  auto diff_nbr_bwd = [&](int t_next) {
    if (t_next >= 1) return neighbor_sum_bwd_act(p, 1, W.gA2, W.gA1, hp, W.m2[t_next], keep, 0, W.dI2, t_next == Td ? 1 : 2, 0, s);
    return neighbor_sum_bwd_act(p, 1, W.gA2, W.gA3, hp, W.inp2, 1.f, 1, W.dI2, Td == 0 ? 1 : 2, 1, s);
  };
  RR_TRY(diff_nbr_bwd(Td));
  for (int t = Td; t >= 1; --t) {
    RR_TRY(linear_wgrad(A, hp, hp, W.gA1, hp, W.nm[t - 1], hp, G + L.dif_Wh_m, hp, G + L.dif_bh, s));
    RR_TRY(linear_wgrad(A, hp, RR_FB_LD, W.gA1, hp, W.nf, RR_FB_LD, G + L.dif_Wh_f, RR_FB_LD, nullptr, s));
    RR_TRY(dgrad(A, hp, hp, W.gA1, hp, P + L.dif_Wh_m, hp, P + L.dif_Wh_m_T, hp, W.gA2, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
    RR_TRY(diff_nbr_bwd(t - 1));
  }
  RR_TRY(linear_wgrad(A, hp, hp, W.dI2, hp, W.d, hp, G + L.dif_Wi, hp, G + L.dif_bi, s));
  RR_TRY(dgrad(A, hp, hp, W.dI2, hp, P + L.dif_Wi, hp, P + L.dif_Wi_T, hp, W.dD, hp, 1, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));

  // shared encoder, once over [reactant rows | product rows]: d(hid_p) = +dD, d(hid_r) = -dD (base_model.py:168), each through the
  // ReLU / dropout mask of its own rows; the weight gradients of both halves come out of one launch per layer (base_model.py:155-156)
  {
    // (the joint graph's tables, and its feature copy if one was needed, are in the workspace since the forward)
    const rr_graph* g = &W.J;
    EncBufs& e = W.j;
    const int A_r = r->n_atoms, A_J = g->n_atoms, B_J = g->n_bonds;
    float* gA1_p = W.gA1 + static_cast<size_t>(A_r) * hp;
    RR_TRY(relu_bwd(A, hp, W.dD, W.enc[1].hid, keep, 0, gA1_p, nullptr, 0, s));
    if (c->r_atom_map) {   // shared reactant rows: sum the gradient over the copies first (gA2 is free until the dgrad below)
      RR_CUDA(cudaMemsetAsync(W.gA2, 0, static_cast<size_t>(A_r) * hp * sizeof(float), s));
      RR_TRY(scatter_add_rows(A, hp, W.dD, c->r_atom_map, W.gA2, s));
      RR_TRY(relu_bwd(A_r, hp, W.gA2, W.enc[0].hid, -keep, 0, W.gA1, nullptr, 0, s));
    } else {
      RR_TRY(relu_bwd(A_r, hp, W.dD, W.enc[0].hid, -keep, 0, W.gA1, nullptr, 0, s));
    }
    RR_TRY(linear_wgrad(A_J, hp, RR_FA_LD, W.gA1, hp, g->f_atoms, RR_FA_LD, G + L.enc_Wo_a, RR_FA_LD, G + L.enc_bo, s));
    RR_TRY(linear_wgrad(A_J, hp, hp, W.gA1, hp, e.am, hp, G + L.enc_Wo_m, hp, nullptr, s));
    RR_TRY(dgrad(A_J, hp, hp, W.gA1, hp, P + L.enc_Wo_m, hp, P + L.enc_Wo_m_T, hp, W.gA2, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
    if (T >= 1) RR_TRY(neighbor_sum_bwd_act(g, 0, W.gA2, W.gB1, hp, e.m[T], keep, 0, W.dinp, 1, 0, s));
    else RR_TRY(neighbor_sum_bwd_act(g, 0, W.gA2, W.gB1, hp, e.inp, 1.f, 1, W.dinp, 1, 1, s));
    for (int t = T; t >= 1; --t) {
      RR_TRY(linear_wgrad(B_J, hp, hp, W.gB1, hp, e.pre[t - 1], hp, G + L.enc_Wh, hp, G + L.enc_bh, s));
      RR_TRY(dgrad(B_J, hp, hp, W.gB1, hp, P + L.enc_Wh, hp, P + L.enc_Wh_T, hp, W.gB2, hp, 0, s, W.hi_off, W.lo_off, W.packed, W.bhi, W.blo));
      if (t - 1 >= 1) RR_TRY(bond_message_bwd_act(g, W.gB2, W.gB1, hp, e.m[t - 1], keep, 0, W.dinp, 2, 0, s));
      else RR_TRY(bond_message_bwd_act(g, W.gB2, W.gB1, hp, e.inp, 1.f, 1, W.dinp, 2, 1, s));
    }
    RR_TRY(linear_wgrad(B_J, hp, RR_FB_LD, W.dinp, hp, g->f_bonds, RR_FB_LD, G + L.enc_Wi, RR_FB_LD, G + L.enc_bi, s));
  }

  PackTable Tb = make_table(*c, L, *grads, false);
  k_pack<<<dim3(32, Tb.n), 256, 0, s>>>(Tb, G, 1, 0, 0, nullptr, nullptr);
  RR_LAUNCH_CHECK("k_pack(grads)");
  return RR_OK;
}

}  // namespace rr
