// K8 batched prompt prefill: the helper kernels around the GEMMs / attention kernels the engine already has.
//
// A prompt of P tokens (sot_prev + hotwords + previous text + sot sequence, up to ~450 tokens:
// whisper_live/transcriber/transcriber_faster_whisper.py:1480-1513) used to be fed one token per decode step, i.e. P - 1
// full weight streams before the first generated token.  The prefill pass pushes all prompt positions of all streams
// through the decoder stack at once -- M = sum(P_b - 1) rows per GEMM on the encoder's tcgen05 kernel, causal
// self-attention over the cached positions, cross-attention in groups of 8 rows per K/V stream -- and leaves the
// self-attention cache and the decode state exactly where token-by-token feeding would have left them.
//
// Row layout: stream b owns rows [rowbase_b, rowbase_b + n'_b), n'_b = its prefill positions rounded up to a multiple
// of 8 (so that a cross-attention group of 8 rows never straddles two streams); padding rows are inactive.
#include "kernels.cuh"

namespace wl {

// x[i] = E[tok[i]] + P[pos[i]] (f32) for active rows, 0 for padding; also the cache indirection of the row:
// src[i][p] = wrow[i] for p <= pos[i] (all prompt positions of a stream live in its first decode row)
__global__ void __launch_bounds__(128) prefill_embed_kernel(const int* __restrict__ tok, const int* __restrict__ pos,
                                                            const int* __restrict__ active, const int* __restrict__ wrow,
                                                            const __half* __restrict__ emb, const __half* __restrict__ pos_emb,
                                                            float* __restrict__ x, short* __restrict__ src, int d) {
  const int i = blockIdx.x;
  float* xr = x + (long)i * d;
  if (!active[i]) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) xr[c] = 0.f;
    return;
  }
  const int t = tok[i], p = pos[i];
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    xr[c] = __half2float(emb[(long)t * d + c]) + __half2float(pos_emb[(long)p * d + c]);
  const short w = (short)wrow[i];
  for (int q = threadIdx.x; q <= p; q += blockDim.x) src[(long)i * T_MAX + q] = w;
}

void prefill_embed(cudaStream_t st, const int* tok, const int* pos, const int* active, const int* wrow, const __half* emb,
                   const __half* pos_emb, float* x, short* src, int M, int d) {
  prefill_embed_kernel<<<M, 128, 0, st>>>(tok, pos, active, wrow, emb, pos_emb, x, src, d);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// k, v of every prompt position -> self-attention cache rows (fp16), BEFORE the attention kernel runs: position p of a
// stream attends to positions < p that other rows of the same launch produce
__global__ void __launch_bounds__(256) prefill_kv_write_kernel(const float* __restrict__ qkv, const int* __restrict__ pos,
                                                               const int* __restrict__ active, const int* __restrict__ wrow,
                                                               __half* __restrict__ kc, __half* __restrict__ vc, long row_stride,
                                                               int H, int d) {
  const int i = blockIdx.x;
  if (!active[i]) return;
  const long base = (long)wrow[i] * row_stride + (long)pos[i] * 64;
  const float* kp = qkv + (long)i * 3 * d + d;
  const float* vp = kp + d;
  for (int c = threadIdx.x * 2; c < d; c += blockDim.x * 2) {
    const int h = c >> 6, dd = c & 63;
    const long o = base + (long)h * T_MAX * 64 + dd;
    *reinterpret_cast<__half2*>(kc + o) = __floats2half2_rn(kp[c], kp[c + 1]);
    *reinterpret_cast<__half2*>(vc + o) = __floats2half2_rn(vp[c], vp[c + 1]);
  }
}

void prefill_kv_write(cudaStream_t st, const float* qkv, const int* pos, const int* active, const int* wrow, __half* kc, __half* vc,
                      long row_stride, int M, int H, int d) {
  prefill_kv_write_kernel<<<M, 256, 0, st>>>(qkv, pos, active, wrow, kc, vc, row_stride, H, d);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// copy selected rows of x (f32 [M][d]) to dst rows 0..n-1 (the decode step's residual buffer: final LayerNorm +
// vocabulary projection then run on them through the decode kernels)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ rows, float* __restrict__ dst, int d) {
  const int j = blockIdx.x, r = rows[j];
  for (int c = threadIdx.x; c < d; c += blockDim.x) dst[(long)j * d + c] = r >= 0 ? x[(long)r * d + c] : 0.f;
}
void gather_rows(cudaStream_t st, const float* x, const int* rows, float* dst, int n, int d) {
  gather_rows_kernel<<<n, 256, 0, st>>>(x, rows, dst, d);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// One block per logits row j: p = softmax(logits[j])[target[j]] -> out[out_index[j]] (target < 0: skip).
// no-speech probability at the sot position (K13) and the teacher-forced token probabilities of align (K14).
__global__ void __launch_bounds__(512) row_prob_kernel(const float* __restrict__ logits, int vocab, int vocab_ld,
                                                       const int* __restrict__ target, const int* __restrict__ out_index,
                                                       float* __restrict__ out) {
  const int j = blockIdx.x, tid = threadIdx.x;
  const int t = target[j];
  if (t < 0) return;
  const float* lg = logits + (long)j * vocab_ld;
  __shared__ float red[2][16];
  float m = -INFINITY;
  for (int i = tid; i < vocab; i += 512) m = fmaxf(m, lg[i]);
  m = warp_max(m);
  if ((tid & 31) == 0) red[0][tid >> 5] = m;
  __syncthreads();
  m = warp_max((tid & 31) < 16 ? red[0][tid & 31] : -INFINITY);
  float sum = 0.f;
  for (int i = tid; i < vocab; i += 512) sum += __expf(lg[i] - m);
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[1][tid >> 5] = sum;
  __syncthreads();
  sum = warp_sum((tid & 31) < 16 ? red[1][tid & 31] : 0.f);
  if (tid == 0) out[out_index[j]] = __expf(lg[t] - m) / sum;
}
void row_prob(cudaStream_t st, const float* logits, int vocab, int vocab_ld, const int* target, const int* out_index, float* out, int n) {
  row_prob_kernel<<<n, 512, 0, st>>>(logits, vocab, vocab_ld, target, out_index, out);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
