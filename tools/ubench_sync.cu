// Micro-benchmark behind the round-2 decode-step design: what does one DEPENDENT stage cost on B200
//   (a) as a kernel boundary inside a CUDA graph, with and without programmatic dependent launch, and
//   (b) as a grid-wide barrier inside one persistent kernel (monotonic counter: red.release + ld.acquire polling)?
// Also checks that a conditional WHILE graph node can drive a device-terminated loop (CUDA >= 12.4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_sync tools/ubench_sync.cu && gpurun_out/ubench_sync
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// one dependent stage: every CTA reads what its neighbour wrote in the previous stage and writes its own value
__global__ void stage_kernel(const float* __restrict__ in, float* __restrict__ out, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (i + blockDim.x) % n;
  out[i] = __ldcg(in + j) + 1.0f;
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// monotonic-counter grid barrier: the k-th barrier is complete when the counter reaches k * gridDim.x
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  epoch += gridDim.x;
  if (threadIdx.x == 0) {
    red_release(ctr, 1u);
    while ((int)(ld_acquire(ctr) - epoch) < 0) {}
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512, 1) persistent_kernel(float* a, float* b, unsigned* ctr, int stages, int n) {
  unsigned epoch = 0;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (i + blockDim.x) % n;
  float* in = a;
  float* out = b;
  for (int s = 0; s < stages; ++s) {
    out[i] = __ldcg(in + j) + 1.0f;
    grid_sync(ctr, epoch);
    float* t = in; in = out; out = t;
  }
}

// ---- conditional WHILE node
__global__ void loop_body_kernel(int* counter, int* iters, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    *iters += 1;
    const int c = *counter - 1;
    *counter = c;
    cudaGraphSetConditional(h, c > 0 ? 1 : 0);
  }
}
__global__ void loop_pre_kernel(float* x) { x[threadIdx.x] += 1.f; }

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int threads = 512, n = sms * threads, STAGES = 2000;
  float *a, *b;
  unsigned* ctr;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&ctr, 256));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;

  for (int pdl = 0; pdl < 2; ++pdl) {
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    float *in = a, *out = b;
    for (int s = 0; s < STAGES; ++s) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(sms); cfg.blockDim = dim3(threads); cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      CK(cudaLaunchKernelEx(&cfg, stage_kernel, (const float*)in, out, n));
      float* t = in; in = out; out = t;
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0, st));
      CK(cudaGraphLaunch(ge, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    printf("graph of %d dependent kernels (%d CTAs x %d thr), PDL=%d: %.3f us per kernel\n", STAGES, sms, threads, pdl, 1000.f * ms / STAGES);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  }
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemsetAsync(ctr, 0, 256, st));
    CK(cudaEventRecord(e0, st));
    persistent_kernel<<<sms, threads, 0, st>>>(a, b, ctr, STAGES, n);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  printf("persistent kernel, %d stages separated by grid barriers (%d CTAs x %d thr): %.3f us per stage\n", STAGES, sms, threads, 1000.f * ms / STAGES);
  std::vector<float> h(n);
  CK(cudaMemcpy(h.data(), (STAGES % 2) ? b : a, n * 4, cudaMemcpyDeviceToHost));
  printf("  check: value %.0f (expect %d)\n", h[0], STAGES);

  // conditional WHILE loop
  {
    int *counter, *iters;
    CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&iters, 4));
    int c0 = 37, z = 0;
    CK(cudaMemcpy(counter, &c0, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(iters, &z, 4, cudaMemcpyHostToDevice));
    cudaGraph_t g;
    CK(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    CK(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams p = { cudaGraphNodeTypeConditional };
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    CK(cudaGraphAddNode(&node, g, nullptr, 0, &p));
    cudaGraph_t body = p.conditional.phGraph_out[0];
    // capture the body from a stream (PDL launches inside)
    CK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(1); cfg.blockDim = dim3(32); cfg.stream = st;
      CK(cudaLaunchKernelEx(&cfg, loop_pre_kernel, a));
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, stage_kernel, (const float*)a, b, n));
      cfg.numAttrs = 0;
      CK(cudaLaunchKernelEx(&cfg, loop_body_kernel, counter, iters, h));
    }
    cudaGraph_t cap;
    CK(cudaStreamEndCapture(st, &cap));
    cudaGraphExec_t ge;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) {
      printf("conditional WHILE graph: instantiate failed: %s\n", cudaGetErrorString(e));
    } else {
      CK(cudaEventRecord(e0, st));
      CK(cudaGraphLaunch(ge, st));
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      int it = 0;
      CK(cudaMemcpy(&it, iters, 4, cudaMemcpyDeviceToHost));
      printf("conditional WHILE graph: body ran %d times (expect 37), %.3f us per iteration (3 kernels)\n", it, 1000.f * ms / it);
    }
  }
  return 0;
}
