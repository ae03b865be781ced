// Shared device/host helpers for librr_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <utility>

#include "rr_sm100.h"

namespace rr {

// ---- error plumbing (no exceptions cross the C ABI) ----------------------------------
extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;   // process-wide: autograd runs backward on its own thread

inline int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define RR_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return rr::fail(RR_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define RR_LAUNCH_CHECK(name)                                                                 \
  do {                                                                                        \
    ++rr::g_launches;                                                                         \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) return rr::fail(RR_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py) ----
enum KernelClass { KC_GEMM_FWD = 0, KC_GEMM_DGRAD, KC_GEMM_WGRAD, KC_BOND_FWD, KC_BOND_BWD, KC_NBR_FWD, KC_NBR_BWD, KC_READOUT,
                   KC_ELEMENTWISE, KC_LOSS, KC_MISC, KC_COUNT };
struct ProfState {
  bool enabled = false;
  struct Rec { int cls; cudaEvent_t a, b; };
  Rec* recs = nullptr;
  int n = 0, cap = 0;
};
extern ProfState g_prof;  // process-wide, guarded by g_prof_mutex
extern std::mutex g_prof_mutex;
void prof_push(int cls, cudaEvent_t a, cudaEvent_t b);
struct ProfScope {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  int cls;
  ProfScope(int cls_, cudaStream_t s_) : s(s_), cls(cls_) {
    if (g_prof.enabled) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, s);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEventRecord(b, s);
      prof_push(cls, a, b);
    }
  }
};

// 0 = fp32 SIMT GEMMs, 1 = tcgen05 3xTF32 GEMMs where the shape allows (forward + dgrad)
extern std::atomic<int> g_gemm_mode;

// Diagnostic / experiment switches from the environment (RR_TC_EW, RR_TC_DIAG, RR_TC_FAKE_PRESPLIT, RR_WG_KT, RR_WG_TF32, RR_WG_BKR, RR_WG3_CFG,
// RR_MP_V1, RR_MP_ACC_RED, RR_MP_CONSUMERS): read ONCE, when the library is first used, not on every launch.  rr_reload_switches() reads
// them again (tests and the micro-benchmarks flip them inside one process).
struct Switches {
  int tc_ew = 16, tc_diag = 0, tc_fake_presplit = 0;
  int wg_kt = 0, wg_tf32 = 0, wg_bkr = 0, wg3_bkr = 0, wg3_raw = 6, wg3_bf = 2;   // wg3_bkr 0: 64-row stages for narrow X tiles, 32 otherwise
  int mp_v1 = 0, mp_acc_red = -1, mp_consumers = 0, mp_kstage = 0;
};
const Switches& switches();
void reload_switches();

#define RR_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return rr::fail(RR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define RR_TRY(expr)          \
  do {                        \
    int _s = (expr);          \
    if (_s != RR_OK) return _s; \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}
// cached per device: one process may drive several GPUs through the `gpu` argument of the Python API
inline int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}
// cudaFuncSetAttribute is per device (per context): remember where a kernel's dynamic shared-memory limit has been raised
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  bool need() const { return !done[current_device()]; }
  void mark() { done[current_device()] = true; }
};

// ---- programmatic dependent launch -------------------------------------------------------
// The big kernels are persistent, one CTA per SM, with a prologue of a few microseconds (barrier init, TMEM allocation, descriptor prefetch).
// Launched with cudaLaunchAttributeProgrammaticStreamSerialization, the CTAs of kernel N+1 start on every SM that kernel N has left and run
// their prologue while N's tail still executes; pdl_wait() (griddepcontrol.wait) then holds every thread until kernel N has completed and
// its writes are visible, BEFORE the first global access of N+1 (reads and writes alike).  RR_NO_PDL=1 falls back to plain launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  static const bool off = getenv("RR_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = off ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- float4 helpers ---------------------------------------------------------------------
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_fma(float s, float4 a, float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}
__device__ __forceinline__ float4 f4_scale(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float4 f4_relu(float4 a) {
  return make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
}
// streaming (read-once) and default loads of one 16-byte chunk
__device__ __forceinline__ float4 ld_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_f4_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
  atomicAdd(p + 0, v.x);
  atomicAdd(p + 1, v.y);
  atomicAdd(p + 2, v.z);
  atomicAdd(p + 3, v.w);
}

// ---- Philox4x32-7: counter-based dropout masks, regenerable from (seed, stream, index) -----
// Seven rounds is the smallest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11, table 2; ten is the default safety margin).
// The masks are drawn in the GEMM epilogue, whose ALU work competes with the warps that feed the tensor core, so rounds are not free.
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t stream_id, uint64_t idx) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(idx), c1 = static_cast<uint32_t>(idx >> 32);
  uint32_t c2 = static_cast<uint32_t>(stream_id), c3 = static_cast<uint32_t>(stream_id >> 32);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-mask scaling of 4 consecutive elements; idx4 = (linear element index) / 4
__device__ __forceinline__ float4 dropout4(float4 v, float p, float inv_keep, uint64_t seed, uint64_t stream_id, uint64_t idx4) {
  const uint4 r = philox4x32(seed, stream_id, idx4);
  const uint32_t thr = static_cast<uint32_t>(fminf(p, 1.f) * 4294967295.f);
  v.x = (r.x >= thr) ? v.x * inv_keep : 0.f;
  v.y = (r.y >= thr) ? v.y * inv_keep : 0.f;
  v.z = (r.z >= thr) ? v.z * inv_keep : 0.f;
  v.w = (r.w >= thr) ? v.w * inv_keep : 0.f;
  return v;
}

// ---- launchers shared between translation units ----------------------------------------------
int padded(int);
int bond_message_fwd(const rr_graph*, const float*, float*, int, int, cudaStream_t);
int bond_message_bwd(const rr_graph*, const float*, float*, int, cudaStream_t);
int neighbor_sum_fwd(const rr_graph*, int, const float*, float*, int, int, cudaStream_t);
int neighbor_sum_bwd(const rr_graph*, int, const float*, float*, int, cudaStream_t);
// gather backward + the ReLU / dropout backward that follows it (y: activation the mask comes from; acc_mode 0 none, 1 acc = dz, 2 acc += dz)
int bond_message_bwd_act(const rr_graph*, const float* dpre, float* dm, int hp, const float* y, float scale, int preact, float* acc, int acc_mode,
                         int skip_out, cudaStream_t);
int neighbor_sum_bwd_act(const rr_graph*, int which, const float* dout, float* dsrc, int ld, const float* y, float scale, int preact, float* acc,
                         int acc_mode, int skip_out, cudaStream_t);
int rowpipe_launch(int op, const rr_graph* g, int which, const float* src, float* out, int ld, int relu_src, const float* y, float scale, int preact,
                   float* acc, int acc_mode, int skip_out, cudaStream_t s);
int pad_rows_act(float* out, const int* rows, int n_rows, int ld, const float* y, float scale, int preact, float* acc, int acc_mode, int skip_out,
                 cudaStream_t s);
int readout_fwd(const rr_graph*, const float*, int, int, const float*, int, float*, int, float, uint64_t, uint64_t, cudaStream_t);
int readout_bwd(const rr_graph*, const float*, int, const float*, const float*, float*, int, float, cudaStream_t);
int relu_bwd(long long, int, const float*, const float*, float, int, float*, float*, int, cudaStream_t);
int sub(long long, const float*, const float*, float*, cudaStream_t);
int sub_gather(long long rows, int ld, const float* a, const float* b, const int* map, float* out, cudaStream_t);
int scatter_add_rows(long long rows, int ld, const float* src, const int* map, float* dst, cudaStream_t);
// hi_off / lo_off: float offsets from W1 / W2 to their pre-split TF32 images (0 = split on chip)
// packed / bhi / blo: base of the packed weights and of their bf16 (hi, lo) images (image of W = bhi + (W - packed)); used when g_fwd_bf16
int linear_fwd(int M, int n, const float* X1, int ldx1, const float* W1, int k1, const float* X2, int ldx2, const float* W2, int k2,
               const float* bias, const float* resid, int ldr, float* Y, int ldy, int flags, float p, uint64_t seed, uint64_t stream_id,
               cudaStream_t s, ptrdiff_t hi_off = 0, ptrdiff_t lo_off = 0, const float* packed = nullptr, const uint16_t* bhi = nullptr,
               const uint16_t* blo = nullptr);
int linear_dgrad(int, int, int, const float*, int, const float*, int, float*, int, int, cudaStream_t);
int linear_wgrad(int, int, int, const float*, int, const float*, int, float*, int, float*, cudaStream_t);
int loss_fwdbwd(int, int, int, const float*, const float*, const int*, float, float, float*, float*, cudaStream_t, int max_group = 0);
int loss_max_group();
int rank_metrics(int, int, const float*, int, const double*, const int*, int, double, double*, cudaStream_t);
int graph_assemble(const rr_mol_store*, int, const int*, const int*, const int*, const int*, const int*, const int*, int, const int*, const int*,
                   const int*, const rr_graph*, cudaStream_t);
long long model_workspace_bytes(const rr_model_cfg*, const rr_graph*, const rr_graph*);
long long model_buffer_offset(const rr_model_cfg*, const rr_graph*, const rr_graph*, const char*);
int model_forward(const rr_model_cfg*, const rr_params*, const rr_graph*, const rr_graph*, const float*, float*, void*, long long, cudaStream_t);
int model_backward(const rr_model_cfg*, const rr_params*, const rr_graph*, const rr_graph*, const float*, rr_params*, void*, long long, cudaStream_t);

// ---- tcgen05 GEMM entry points (rr_gemm_tc.cu) ---------------------------------------------
bool tc_supported(int M, int n, int k1, int k2, int ldx1, int ldx2);
int tc_linear(int M, int n, const float* X1, int ldx1, const float* W1, int ldw1, int k1, const float* X2, int ldx2, const float* W2, int ldw2, int k2,
              const float* bias, const float* resid, int ldr, float* Y, int ldy, int relu, int accumulate, float p, uint64_t seed, uint64_t stream_id,
              int kclass, cudaStream_t s, const float* W1lo = nullptr, const float* W2lo = nullptr);
bool tc_linear_bf16_supported(int M, int n, int k, int ldx, int ldw);
int tc_linear_bf16(int M, int n, const float* X, int ldx, const uint16_t* Whi, const uint16_t* Wlo, int ldw, int k, float* Y, int ldy, int accumulate,
                   int kclass, cudaStream_t s);
int tc_linear_bf16_full(int M, int n, const float* X1, int ldx1, const uint16_t* W1hi, const uint16_t* W1lo, int ldw1, int k1, const float* X2, int ldx2,
                        const uint16_t* W2hi, const uint16_t* W2lo, int ldw2, int k2, const float* bias, const float* resid, int ldr, float* Y, int ldy,
                        int relu, int accumulate, float p, uint64_t seed, uint64_t stream_id, int kclass, cudaStream_t s);
// forward GEMMs: 0 = 3 x tf32 split (default), 1 = 3 x bf16 split (RR_FWD_BF16=1 or rr_set_forward_bf16(1))
extern std::atomic<int> g_fwd_bf16;
// backward GEMMs: 1 = 3 x bf16 split (default; RR_BWD_BF16=0 or rr_set_backward_bf16(0) keeps 3 x tf32)
extern std::atomic<int> g_bwd_bf16;
long long tc_dgrad_scratch_bytes(int n, int k);
int tc_dgrad_standalone(int M, int n, int k, const float* dZ, int lddz, const float* W, int ldw, float* dX, int lddx, int accumulate, void* scratch,
                        long long scratch_bytes, cudaStream_t s);
bool tc_wgrad_supported(int M, int n, int k, int lddz, int ldx);
int tc_wgrad_trace(unsigned long long* host_out, int n);
int tc_wgrad(int M, int n, int k, const float* dZ, int lddz, const float* X, int ldx, float* dW, int lddw, float* dbias, cudaStream_t s);

}  // namespace rr
