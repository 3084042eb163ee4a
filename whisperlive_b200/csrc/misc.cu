// Launch accounting, kernel attribute priming and the K14 attention-probability gather.
#include <atomic>
#include <cstdlib>

#include "kernels.cuh"

namespace wl {

static std::atomic<long> g_other_launches{0};
long other_launch_count() { return g_other_launches.load(); }
void note_launch(int n) { g_other_launches += n; }

static thread_local bool t_pdl = false;
static bool pdl_env() {
  static const bool on = [] { const char* e = getenv("WLB200_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}
bool pdl_active() { return t_pdl; }
PdlScope::PdlScope(bool on) : prev(t_pdl) { t_pdl = on && pdl_env(); }
PdlScope::~PdlScope() { t_pdl = prev; }

// copy the cross-attention probabilities of the alignment heads that live in `layer`
__global__ void gather_align_kernel(DecodeState s, const float* __restrict__ probs, float* __restrict__ buf,
                                    const int* __restrict__ heads, int n_heads, int layer, int rows_per_stream, int H) {
  const int b = blockIdx.y, i = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  if (s.done[b] || heads[2 * i] != layer) return;
  const int h = heads[2 * i + 1], r = b * rows_per_stream, pos = s.pos[r];
  const float* src = probs + ((long)r * H + h) * S_ENC;
  float* dst = buf + (((long)b * n_heads + i) * T_MAX + pos) * S_ENC;
  for (int k = threadIdx.x; k < S_ENC; k += blockDim.x) dst[k] = src[k];
}

void gather_align_probs(cudaStream_t st, const DecodeState& s, const float* probs, float* buf, const int* heads, int n_heads,
                        int layer, int B, int rows_per_stream, int H) {
  dim3 grid(n_heads, B);
  launch_kernel(gather_align_kernel, grid, dim3(256), 0, st, s, probs, buf, heads, n_heads, layer, rows_per_stream, H);
  note_launch(1);
}

// batched pass (K8 machinery): rows of a cross-attention chunk -> align_buf[b][head slot][pos][1500]
__global__ void gather_align_rows_kernel(const float* __restrict__ probs, const int* __restrict__ row_b, const int* __restrict__ row_pos,
                                         const int* __restrict__ row_active, float* __restrict__ buf, const int* __restrict__ heads,
                                         int n_heads, int layer, int row0, int H) {
  const int i = blockIdx.x, rl = blockIdx.y, r = row0 + rl;
  if (heads[2 * i] != layer || !row_active[r]) return;
  const int h = heads[2 * i + 1];
  const float* src = probs + ((long)rl * H + h) * S_ENC;
  float* dst = buf + (((long)row_b[r] * n_heads + i) * T_MAX + row_pos[r]) * S_ENC;
  for (int k = threadIdx.x; k < S_ENC; k += blockDim.x) dst[k] = src[k];
}
void gather_align_rows(cudaStream_t st, const float* probs, const int* row_b, const int* row_pos, const int* row_active, float* buf,
                       const int* heads, int n_heads, int layer, int row0, int n_rows, int H) {
  dim3 grid(n_heads, n_rows);
  gather_align_rows_kernel<<<grid, 256, 0, st>>>(probs, row_b, row_pos, row_active, buf, heads, n_heads, layer, row0, H);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
