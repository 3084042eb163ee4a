// What makes a dependent kernel boundary inside a replayed CUDA graph cost ~3 us in the real decode step when a chain of
// identical trivial kernels costs 1.1 us?  Variants (all PDL, 2000-kernel chains, us per kernel):
//   A  one tiny kernel, 148 x 512                      E  one tiny kernel, 16 x 128 (LayerNorm-sized grid)
//   B  12 distinct tiny kernels, round robin           F  tiny kernel that owns 200 KB of dynamic smem (no co-residency)
//   C  12 distinct kernels with ~16 KB of code each    G  C + the code path is executed once BEFORE the dependency wait
//   D  one tiny kernel with 640 B of parameters        H  tiny kernel, 3 dependent L2 round trips after the wait
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct Big { float pad[160]; };

template <int ID, int CODE, int PREWARM>
__device__ __forceinline__ float body(float v, int iters) {
  // CODE x 64 dependent FMAs with distinct constants: straight-line code of about CODE KB
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CODE * 64; ++i) v = fmaf(v, 1.0f + 1e-7f * (float)(ID * 4096 + i), 1e-9f * (float)i);
  }
  return v;
}

template <int ID, int CODE, int PREWARM>
__global__ void k_chain(const float* __restrict__ in, float* __restrict__ out, int n, int iters, int zero) {
  pdl_trigger();
  float w = 0.f;
  if (PREWARM) w = body<ID, CODE, PREWARM>((float)threadIdx.x, zero + 1) * (float)zero;   // same code, result discarded (zero == 0)
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (i + blockDim.x) % n;
  float v = __ldcg(in + j) + w;
  if (CODE > 0) v = body<ID, CODE, PREWARM>(v, iters);
  out[i] = v + 1.0f;
}
__global__ void k_big(const float* __restrict__ in, float* __restrict__ out, int n, Big b0, Big b1_) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  out[i] = __ldcg(in + (i + blockDim.x) % n) + 1.0f + b0.pad[threadIdx.x % 160] * 0.f;
}
__global__ void k_smem(const float* __restrict__ in, float* __restrict__ out, int n) {
  extern __shared__ float sm[];
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  sm[threadIdx.x] = __ldcg(in + (i + blockDim.x) % n);
  out[i] = sm[threadIdx.x] + 1.0f;
}
__global__ void k_dep3(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ idx, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = (i + blockDim.x) % n;
  j = __ldcg(idx + j);
  j = __ldcg(idx + j);
  out[i] = __ldcg(in + j) + 1.0f;
}

typedef void (*LaunchFn)(cudaStream_t, const float*, float*, int, int);
static int g_sms, g_grid, g_threads;
static int* g_idx;
template <int ID, int CODE, int PRE>
static void launch_chain(cudaStream_t st, const float* in, float* out, int n, int s) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_grid); cfg.blockDim = dim3(g_threads); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_chain<ID, CODE, PRE>, in, out, n, 1, 0));
}
template <int CODE, int PRE>
static void launch_rr(cudaStream_t st, const float* in, float* out, int n, int s) {
  switch (s % 12) {
    case 0: launch_chain<0, CODE, PRE>(st, in, out, n, s); break;   case 1: launch_chain<1, CODE, PRE>(st, in, out, n, s); break;
    case 2: launch_chain<2, CODE, PRE>(st, in, out, n, s); break;   case 3: launch_chain<3, CODE, PRE>(st, in, out, n, s); break;
    case 4: launch_chain<4, CODE, PRE>(st, in, out, n, s); break;   case 5: launch_chain<5, CODE, PRE>(st, in, out, n, s); break;
    case 6: launch_chain<6, CODE, PRE>(st, in, out, n, s); break;   case 7: launch_chain<7, CODE, PRE>(st, in, out, n, s); break;
    case 8: launch_chain<8, CODE, PRE>(st, in, out, n, s); break;   case 9: launch_chain<9, CODE, PRE>(st, in, out, n, s); break;
    case 10: launch_chain<10, CODE, PRE>(st, in, out, n, s); break; default: launch_chain<11, CODE, PRE>(st, in, out, n, s); break;
  }
}
static void launch_big(cudaStream_t st, const float* in, float* out, int n, int s) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_grid); cfg.blockDim = dim3(g_threads); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  Big b; memset(&b, 0, sizeof(b));
  CK(cudaLaunchKernelEx(&cfg, k_big, in, out, n, b, b));
}
static void launch_smem(cudaStream_t st, const float* in, float* out, int n, int s) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_grid); cfg.blockDim = dim3(g_threads); cfg.stream = st; cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_smem, in, out, n));
}
static void launch_dep3(cudaStream_t st, const float* in, float* out, int n, int s) {
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_grid); cfg.blockDim = dim3(g_threads); cfg.stream = st;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k_dep3, in, out, (const int*)g_idx, n));
}

static void run(const char* name, LaunchFn fn, int grid, int threads, cudaStream_t st, float* a, float* b) {
  g_grid = grid; g_threads = threads;
  const int n = grid * threads, STAGES = 2000;
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  float *in = a, *out = b;
  for (int s = 0; s < STAGES; ++s) { fn(st, in, out, n, s); float* t = in; in = out; out = t; }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("%-72s %.3f us per kernel\n", name, 1000.f * ms / STAGES);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
}

int main() {
  CK(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0));
  const int nmax = g_sms * 512;
  float *a, *b;
  CK(cudaMalloc(&a, nmax * 4)); CK(cudaMalloc(&b, nmax * 4)); CK(cudaMalloc(&g_idx, nmax * 4));
  CK(cudaMemset(a, 0, nmax * 4)); CK(cudaMemset(b, 0, nmax * 4)); CK(cudaMemset(g_idx, 0, nmax * 4));
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  run("A one tiny kernel, 148 x 512", launch_chain<0, 0, 0>, g_sms, 512, st, a, b);
  run("B 12 distinct tiny kernels round robin, 148 x 512", launch_rr<0, 0>, g_sms, 512, st, a, b);
  run("C 12 distinct kernels, ~16 KB straight-line code each, 148 x 128", launch_rr<16, 0>, g_sms, 128, st, a, b);
  run("C1 ONE kernel, ~16 KB straight-line code, 148 x 128 (code stays warm)", launch_chain<0, 16, 0>, g_sms, 128, st, a, b);
  run("G 12 distinct kernels, 16 KB code, executed once BEFORE the wait too", launch_rr<16, 1>, g_sms, 128, st, a, b);
  run("C4 12 distinct kernels, ~4 KB code each, 148 x 128", launch_rr<4, 0>, g_sms, 128, st, a, b);
  run("D one tiny kernel with 1280 B of parameters, 148 x 512", launch_big, g_sms, 512, st, a, b);
  run("E one tiny kernel, 16 x 128", launch_chain<0, 0, 0>, 16, 128, st, a, b);
  run("E2 one tiny kernel, 80 x 256", launch_chain<0, 0, 0>, 80, 256, st, a, b);
  run("F tiny kernel owning 200 KB dynamic smem, 148 x 384", launch_smem, g_sms, 384, st, a, b);
  run("H tiny kernel, 3 dependent L2 round trips, 148 x 128", launch_dep3, g_sms, 128, st, a, b);
  return 0;
}
