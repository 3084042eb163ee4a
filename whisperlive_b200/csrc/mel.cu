// K1: fused log-mel front end (fp32).  One CTA = 16 STFT frames of one stream:
//   PCM span -> smem (reflect / zero padding resolved on load) -> Hann window + even/odd fold ->
//   201-bin DFT by direct summation against a smem twiddle table -> |.|^2 -> Slaney filterbank ->
//   log10 -> global (raw) + per-stream running max (atomic).  A second pass applies
//   max(x, gmax-8), (x+4)/4.  Semantics: faster-whisper FeatureExtractor (oracle/mel.py; reference
//   call sites transcriber_faster_whisper.py:862, batch_inference.py:258; in-repo formula
//   tensorrt_utils.py:177-190).
#include "kernels.cuh"

namespace wl {

constexpr int MEL_FT = 16;        // frames per CTA
constexpr int MEL_NFFT = 400;
constexpr int MEL_HOP = 160;
constexpr int MEL_BINS = 201;
constexpr int MEL_SPAN = (MEL_FT - 1) * MEL_HOP + MEL_NFFT;  // 2800 samples

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void __launch_bounds__(256) mel_stft_kernel(const float* __restrict__ pcm, const long* __restrict__ pcm_off,
                                                       float* __restrict__ out, const long* __restrict__ out_off,
                                                       unsigned* __restrict__ gmax, const float* __restrict__ window,
                                                       const float* __restrict__ twiddle,  // [400][2] cos,sin
                                                       const float* __restrict__ filt,     // [n_mels][201]
                                                       const int* __restrict__ filt_range, // [n_mels][2]
                                                       int n_mels) {
  __shared__ float xs[MEL_SPAN];
  __shared__ __align__(16) float EO[2 * 200 + 8][MEL_FT];
  float(*E)[MEL_FT] = EO;         // even fold  x[i] + x[400-i]
  float(*O)[MEL_FT] = EO + 200;   // odd fold   x[i] - x[400-i]
  __shared__ float tw_c[MEL_NFFT], tw_s[MEL_NFFT];
  __shared__ float red[8];
  float(*P)[MEL_FT] = EO;         // power spectrum [201][16] reuses the fold storage after the DFT

  const int b = blockIdx.y;
  const long n = pcm_off[b + 1] - pcm_off[b];
  const int T = (int)(n / MEL_HOP) + 1;
  const int f0 = blockIdx.x * MEL_FT;
  if (f0 >= T) return;
  const float* x = pcm + pcm_off[b];
  const long L = n + MEL_HOP;  // waveform padded with 160 zeros
  const int tid = threadIdx.x;

  for (int i = tid; i < MEL_SPAN; i += 256) {
    long j = (long)f0 * MEL_HOP - MEL_NFFT / 2 + i;
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    xs[i] = (j >= 0 && j < n) ? x[j] : 0.f;
  }
  for (int i = tid; i < MEL_NFFT; i += 256) {
    tw_c[i] = twiddle[2 * i];
    tw_s[i] = twiddle[2 * i + 1];
  }
  __syncthreads();
  for (int idx = tid; idx < 200 * MEL_FT; idx += 256) {
    const int i = idx / MEL_FT, f = idx % MEL_FT;
    const float a = xs[f * MEL_HOP + i] * window[i];
    if (i == 0) {
      E[0][f] = a;
      O[0][f] = xs[f * MEL_HOP + 200] * window[200];
    } else {
      const float c = xs[f * MEL_HOP + MEL_NFFT - i] * window[MEL_NFFT - i];
      E[i][f] = a + c;
      O[i][f] = a - c;
    }
  }
  __syncthreads();
  float re[MEL_FT], im[MEL_FT];
  const int k = tid;
  if (k < MEL_BINS) {
    const float sgn = (k & 1) ? -1.f : 1.f;
#pragma unroll
    for (int f = 0; f < MEL_FT; ++f) {
      re[f] = E[0][f] + sgn * O[0][f];
      im[f] = 0.f;
    }
    int ph = 0;
    for (int i = 1; i < 200; ++i) {
      ph += k;
      if (ph >= MEL_NFFT) ph -= MEL_NFFT;
      const float c = tw_c[ph], s = tw_s[ph];
#pragma unroll
      for (int f4 = 0; f4 < MEL_FT / 4; ++f4) {
        const float4 e = *reinterpret_cast<const float4*>(&E[i][4 * f4]);
        const float4 o = *reinterpret_cast<const float4*>(&O[i][4 * f4]);
        re[4 * f4 + 0] = fmaf(e.x, c, re[4 * f4 + 0]); im[4 * f4 + 0] = fmaf(o.x, s, im[4 * f4 + 0]);
        re[4 * f4 + 1] = fmaf(e.y, c, re[4 * f4 + 1]); im[4 * f4 + 1] = fmaf(o.y, s, im[4 * f4 + 1]);
        re[4 * f4 + 2] = fmaf(e.z, c, re[4 * f4 + 2]); im[4 * f4 + 2] = fmaf(o.z, s, im[4 * f4 + 2]);
        re[4 * f4 + 3] = fmaf(e.w, c, re[4 * f4 + 3]); im[4 * f4 + 3] = fmaf(o.w, s, im[4 * f4 + 3]);
      }
    }
  }
  __syncthreads();  // everyone done reading E/O
  if (k < MEL_BINS) {
#pragma unroll
    for (int f = 0; f < MEL_FT; ++f) P[k][f] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();
  float lmax = -INFINITY;
  float* o = out + out_off[b];
  for (int idx = tid; idx < n_mels * MEL_FT; idx += 256) {
    const int m = idx / MEL_FT, f = idx % MEL_FT;
    const int lo = filt_range[2 * m], hi = filt_range[2 * m + 1];
    float acc = 0.f;
    for (int kk = lo; kk < hi; ++kk) acc = fmaf(filt[m * MEL_BINS + kk], P[kk][f], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    if (f0 + f < T) {
      o[(long)m * T + f0 + f] = v;
      lmax = fmaxf(lmax, v);
    }
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float mx = red[0];
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    atomicMax(&gmax[b], f2ord(mx));
  }
}

__global__ void mel_norm_kernel(float* __restrict__ out, const long* __restrict__ out_off, const unsigned* __restrict__ gmax) {
  const int b = blockIdx.y;
  const long n = out_off[b + 1] - out_off[b];
  const float floor_v = ord2f(gmax[b]) - 8.0f;
  float* o = out + out_off[b];
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    o[i] = (fmaxf(o[i], floor_v) + 4.0f) * 0.25f;
}

void mel_forward(cudaStream_t st, const float* pcm, const long* pcm_off, float* out, const long* out_off, unsigned* gmax,
                 const MelTables& t, int B, int max_frames) {
  WL_CUDA(cudaMemsetAsync(gmax, 0, sizeof(unsigned) * B, st));
  dim3 grid(cdiv(max_frames, MEL_FT), B);
  mel_stft_kernel<<<grid, 256, 0, st>>>(pcm, pcm_off, out, out_off, gmax, t.window, t.twiddle, t.filt, t.filt_range, t.n_mels);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
  dim3 g2(cdiv((long)max_frames * t.n_mels, 256 * 4), B);
  mel_norm_kernel<<<g2, 256, 0, st>>>(out, out_off, gmax);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
