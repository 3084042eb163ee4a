// Shared device/host helpers for libwlb200 (sm_100a only).
// PTX wrappers for mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc/mma/commit/ld) --
// hand-written; field layouts cross-checked against cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <utility>

#define WL_OK 0
#define WL_ERR_CUDA -1
#define WL_ERR_ARG -2
#define WL_ERR_STATE -3
#define WL_ERR_NOMEM -4

namespace wl {

struct Error {
  int code;
  std::string msg;
};

#define WL_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw wl::Error{WL_ERR_CUDA, _b};                                                       \
    }                                                                                         \
  } while (0)

#define WL_CHECK(cond, code, ...)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      char _b[512];                                                \
      snprintf(_b, sizeof(_b), __VA_ARGS__);                       \
      throw wl::Error{code, std::string(_b)};                      \
    }                                                              \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }
void note_launch(int n);  // kernel launch accounting (misc.cu)

// Launches issued while a PdlScope(true) is alive carry the programmatic-stream-serialization attribute (the decoder
// step: ~420 short dependent kernels per token, whose launch ramps and weight prefetch overlap the predecessor's
// tail).  WLB200_PDL=0 disables it.
bool pdl_active();
struct PdlScope {
  explicit PdlScope(bool on);
  ~PdlScope();
  bool prev;
};

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_active() ? 1 : 0;
  WL_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

// ----------------------------------------------------------------------------------- in-graph timeline (debug)
// WLB200_TIMELINE=<file>: every decode-step kernel stamps %globaltimer at entry (what=0) and right after its
// dependency wait (what=1) from one thread of block 0 into a device log (slot 0 = entry counter).  The deltas between
// consecutive "ready" stamps are the per-kernel cost INSIDE the replayed CUDA graph, which ncu cannot show
// (tools/timeline.py).  One pointer per translation unit (no relocatable device code), bound by tl_bind().
constexpr unsigned TL_CAP = 1u << 20;
static __device__ unsigned long long* wl_tl_buf = nullptr;
// Compiled in only with -DWLB200_TL=1 (python -m whisperlive_b200.build --timeline): even one thread's worth of
// stamping code is instruction-cache footprint in kernels that are launched 400 times per token.
#ifndef WLB200_TL
#define WLB200_TL 0
#endif
__device__ __forceinline__ void tl_stamp_any(int kernel_id, int what) {
  if (!WLB200_TL) return;
  unsigned long long* buf = wl_tl_buf;
  if (buf != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned idx = atomicAdd(reinterpret_cast<unsigned*>(buf), 1u);
    if (idx < TL_CAP) buf[1 + idx] = (t << 8) | ((unsigned long long)kernel_id << 2) | (unsigned long long)what;
  }
}
__device__ __forceinline__ void tl_stamp(int kernel_id, int what) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) tl_stamp_any(kernel_id, what);
}
static inline void tl_bind_tu(unsigned long long* p) { cudaMemcpyToSymbol(wl_tl_buf, &p, sizeof(p)); }
enum TlKernel : int { TL_EMBED = 1, TL_LN = 2, TL_GEMM_PART = 3, TL_SELF = 4, TL_CROSS = 5, TL_COMBINE = 6, TL_GELU = 7,
                      TL_SROWS = 8, TL_SSTREAMS = 9, TL_GEMM = 10, TL_DSTEP = 11 };

// ----------------------------------------------------------------------------------- generic
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
      "elect.sync r|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// GELU in its exact erf form, 0.5 x (1 + erf(x / sqrt 2)).  erf is evaluated with Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, i.e. fp32 rounding level): with p = poly(t) t exp(-z^2), t = 1 / (1 + 0.3275911 z),
// z = |x| / sqrt 2, gelu(x) = max(x, 0) - |x| p / 2.  Two MUFU ops (rcp.approx, ex2.approx) + 11 FP32 ops and
// no branches -- the epilogue of the 4d-wide MLP GEMM evaluates it 61 M times per encoder layer pass.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.0f)));
  const float w = ax * 0.84932180028801904272f;   // sqrt(log2(e) / 2): exp(-x^2 / 2) = 2^(-w^2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-w * w));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  return fmaxf(x, 0.f) - ax * (poly * t * e);
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// start while its predecessor in the stream is still running; it must execute pdl_wait() before it touches anything
// the predecessor wrote (weights and other long-lived data may be fetched earlier).  pdl_trigger() lets the
// successor's CTAs be scheduled as soon as every CTA of this grid has issued it.  Both are no-ops for kernels
// launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 4-D tiled load: coordinates innermost first.
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 1-D bulk copy global -> shared, bytes multiple of 16, 16B-aligned both sides.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ----------------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16/bf16 inputs, fp32 accumulate), single CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread i of the warp = lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major operand tile stored as rows of 128 bytes with the
// 128-byte swizzle (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart (SBO), LBO unused.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16: fp32 accumulate, both operands K-major.
// bits: [4,6) D fmt (1=f32) | [7,10) A fmt | [10,13) B fmt (0=f16, 1=bf16) | 15 A major | 16 B major |
//       [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, bool bf16 = false) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
#endif  // __CUDACC__

}  // namespace wl
