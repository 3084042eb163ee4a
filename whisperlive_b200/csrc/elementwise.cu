// Row-wise helper kernels: feature layout prep, LayerNorm, softmax, decoder token embedding.
#include <algorithm>

#include "kernels.cuh"

namespace wl {

// ---------------------------------------------------------------------------- prep_features
// [B][n_mels][3000] f32 -> [B][3002][n_mels] fp16 with one zero row before and after (conv k=3, pad=1
// becomes a plain strided GEMM over overlapping rows).  32x32 smem transpose tiles.
__global__ void prep_features_kernel(const float* __restrict__ in, __half* __restrict__ out, int n_mels) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
  const float* src = in + (long)b * n_mels * 3000;
  __half* dst = out + (long)b * 3002 * n_mels;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int m = m0 + i, t = t0 + threadIdx.x;
    tile[i][threadIdx.x] = (m < n_mels && t < 3000) ? src[(long)m * 3000 + t] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int t = t0 + i, m = m0 + threadIdx.x;
    if (t < 3000 && m < n_mels) dst[(long)(t + 1) * n_mels + m] = __float2half_rn(tile[threadIdx.x][i]);
  }
  if (blockIdx.x == 0 && threadIdx.y == 0) {
    const int m = m0 + threadIdx.x;
    if (m < n_mels) {
      dst[m] = __float2half_rn(0.f);
      dst[(long)3001 * n_mels + m] = __float2half_rn(0.f);
    }
  }
}

void prep_features(cudaStream_t st, const float* feats, __half* out, int B, int n_mels) {
  dim3 grid(cdiv(3000, 32), cdiv(n_mels, 32), B), block(32, 8);
  prep_features_kernel<<<grid, block, 0, st>>>(feats, out, n_mels);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// ---------------------------------------------------------------------------- layernorm_rows
// One warp per row, the row lives in registers (d <= 1280 -> 40 values per lane), two-pass statistics.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ be, __half* __restrict__ y,
                                                        float* __restrict__ y32, long rows, int d, float* __restrict__ zero_buf,
                                                        long zero_n) {
  // optional side job: clear the fp32 buffer the next split-K GEMM accumulates into
  for (long i = blockIdx.x * 256L + threadIdx.x; i < zero_n; i += (long)gridDim.x * 256) zero_buf[i] = 0.f;
  const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * d;
  float v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < d ? xr[c] : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = i * 32 + lane;
    const float t = c < d ? v[i] - mean : 0.f;
    sq += t * t;
  }
  const float rstd = rsqrtf(warp_sum(sq) / d + 1e-5f);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = i * 32 + lane;
    if (c < d) {
      const float o = (v[i] - mean) * rstd * g[c] + be[c];
      if (y) y[row * d + c] = __float2half_rn(o);
      if (y32) y32[row * d + c] = o;
    }
  }
}

void layernorm_rows(cudaStream_t st, const float* x, const float* gamma, const float* beta, __half* y, float* y32, long rows,
                    int d, float* zero_buf, long zero_n) {
  WL_CHECK(d <= 1280 && d % 32 == 0, WL_ERR_ARG, "layernorm: unsupported width %d", d);
  int grid = cdiv(rows, 8);
  if (zero_buf && zero_n > 0) grid = std::max(grid, std::min(148, cdiv(zero_n, 256 * 16)));
  if (d <= 512) layernorm_kernel<16><<<grid, 256, 0, st>>>(x, gamma, beta, y, y32, rows, d, zero_buf, zero_n);
  else if (d <= 1024) layernorm_kernel<32><<<grid, 256, 0, st>>>(x, gamma, beta, y, y32, rows, d, zero_buf, zero_n);
  else layernorm_kernel<40><<<grid, 256, 0, st>>>(x, gamma, beta, y, y32, rows, d, zero_buf, zero_n);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// ---------------------------------------------------------------------------- gelu_cast
__global__ void gelu_cast_kernel(const float* __restrict__ in, __half* __restrict__ out, long n) {
  const long i = (blockIdx.x * 256L + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(in + i);
  __align__(8) __half2 h[2] = {__floats2half2_rn(gelu_erf(v.x), gelu_erf(v.y)), __floats2half2_rn(gelu_erf(v.z), gelu_erf(v.w))};
  *reinterpret_cast<uint2*>(out + i) = *reinterpret_cast<const uint2*>(h);
}

void gelu_cast(cudaStream_t st, const float* in, __half* out, long n) {
  WL_CHECK(n % 4 == 0, WL_ERR_ARG, "gelu_cast: n must be a multiple of 4");
  gelu_cast_kernel<<<cdiv(n / 4, 256), 256, 0, st>>>(in, out, n);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// ---------------------------------------------------------------------------- softmax_rows
// One warp per row of <= 1536 scores held in registers.
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ s, __half* __restrict__ p, long rows, int n,
                                                      int ld_in, int ld_out, float scale) {
  const long row = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* sr = s + row * ld_in;
  float v[48];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 48; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < n ? sr[c] * scale : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 48; ++i) {
    v[i] = __expf(v[i] - mx);
    sum += v[i];
  }
  const float inv = 1.f / warp_sum(sum);
  __half* pr = p + row * ld_out;
#pragma unroll
  for (int i = 0; i < 48; ++i) {
    const int c = i * 32 + lane;
    if (c < ld_out) pr[c] = __float2half_rn(c < n ? v[i] * inv : 0.f);
  }
}

void softmax_rows(cudaStream_t st, const float* s, __half* p, long rows, int n, int ld_in, int ld_out, float scale) {
  WL_CHECK(n <= 1536 && ld_out <= 1536, WL_ERR_ARG, "softmax_rows: row too long (%d)", n);
  softmax_kernel<<<cdiv(rows, 8), 256, 0, st>>>(s, p, rows, n, ld_in, ld_out, scale);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

// ---------------------------------------------------------------------------- decoder_embed
__global__ void decoder_embed_kernel(DecodeState s, const __half* __restrict__ emb, const __half* __restrict__ pos_emb,
                                     float* __restrict__ x, int d) {
  const int r = blockIdx.x;
  if (!s.active[r]) return;
  const int tok = s.tok_in[r], pos = s.pos[r];
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    x[(long)r * d + c] = __half2float(emb[(long)tok * d + c]) + __half2float(pos_emb[(long)pos * d + c]);
  if (threadIdx.x == 0) s.src[(long)r * T_MAX + pos] = (short)r;
}

void decoder_embed(cudaStream_t st, const DecodeState& s, const __half* emb, const __half* pos_emb, float* x, int R, int d) {
  decoder_embed_kernel<<<R, 128, 0, st>>>(s, emb, pos_emb, x, d);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
