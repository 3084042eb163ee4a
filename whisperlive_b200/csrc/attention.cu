// Decoder attention kernels (fp32 math, fp16 K/V storage).
//   K10 decoder_self_attn : KV cache append + attention over <=448 positions with beam indirection
//   K11 decoder_cross_attn: all rows (beams) of a stream against the stream's persistent encoder K/V,
//       streamed HBM -> smem by a producer warp with cp.async.bulk + mbarrier ring, consumed by 4 warps.
#include <algorithm>

#include "kernels.cuh"

namespace wl {

// ============================================================================ K10 self attention
__global__ void __launch_bounds__(128) self_attn_kernel(DecodeState s, PartialSrc qkv, __half* __restrict__ kc,
                                                        __half* __restrict__ vc, long row_stride, __half* __restrict__ out,
                                                        int H, int d) {
  const int r = blockIdx.y, h = blockIdx.x, tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (!s.active[r]) return;
  const int pos = s.pos[r];
  const int n = pos + 1;
  __shared__ float q[64], knew[64], vnew[64];
  __shared__ float sc[T_MAX];
  __shared__ float red[8];
  __shared__ float opart[16][64];

  if (tid < 64) {
    // q / k / v of this (row, head): bias + the split-K partial sums, in order
    const long off = (long)r * 3 * d + h * 64 + tid;
    float qv = 0.f, kv = 0.f, vv = 0.f;
    if (qkv.bias) { qv = qkv.bias[h * 64 + tid]; kv = qkv.bias[d + h * 64 + tid]; vv = qkv.bias[2 * d + h * 64 + tid]; }
#pragma unroll 4
    for (int sp = 0; sp < qkv.nsplit; ++sp) {
      const float* row = qkv.ptr + (long)sp * qkv.stride + off;
      qv += __ldcg(row); kv += __ldcg(row + d); vv += __ldcg(row + 2 * d);
    }
    q[tid] = qv * 0.125f;
    knew[tid] = kv;
    vnew[tid] = vv;
    const long o = (long)r * row_stride + ((long)h * T_MAX + pos) * 64 + tid;
    kc[o] = __float2half_rn(kv);
    vc[o] = __float2half_rn(vv);
  }
  __syncthreads();
  const short* src = s.src + (long)r * T_MAX;
  float lmax = -INFINITY;
  for (int p = tid; p < n; p += 128) {
    float acc = 0.f;
    if (p == pos) {
#pragma unroll 8
      for (int e = 0; e < 64; ++e) acc = fmaf(q[e], knew[e], acc);
    } else {
      const uint4* kp = reinterpret_cast<const uint4*>(kc + (long)src[p] * row_stride + ((long)h * T_MAX + p) * 64);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = kp[c];
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          acc = fmaf(q[c * 8 + 2 * e], f.x, acc);
          acc = fmaf(q[c * 8 + 2 * e + 1], f.y, acc);
        }
      }
    }
    sc[p] = acc;
    lmax = fmaxf(lmax, acc);
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) red[tid >> 5] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float lsum = 0.f;
  for (int p = tid; p < n; p += 128) {
    const float e = __expf(sc[p] - mx);
    sc[p] = e;
    lsum += e;
  }
  lsum = warp_sum(lsum);
  if ((tid & 31) == 0) red[4 + (tid >> 5)] = lsum;
  __syncthreads();
  const float inv = 1.f / (red[4] + red[5] + red[6] + red[7]);

  const int c8 = tid & 7, g = tid >> 3;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int p = g; p < n; p += 16) {
    const float w = sc[p];
    if (p == pos) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, vnew[c8 * 8 + e], acc[e]);
    } else {
      const uint4 u = *reinterpret_cast<const uint4*>(vc + (long)src[p] * row_stride + ((long)h * T_MAX + p) * 64 + c8 * 8);
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        acc[2 * e] = fmaf(w, f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(w, f.y, acc[2 * e + 1]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) opart[g][c8 * 8 + e] = acc[e];
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
#pragma unroll
    for (int gg = 0; gg < 16; ++gg) t += opart[gg][tid];
    out[(long)r * d + h * 64 + tid] = __float2half_rn(t * inv);
  }
}

void decoder_self_attn(cudaStream_t st, const DecodeState& s, const PartialSrc& qkv, __half* kcache, __half* vcache,
                       long cache_row_stride, __half* out, int R, int H, int d) {
  dim3 grid(H, R);
  launch_kernel(self_attn_kernel, grid, dim3(128), 0, st, s, qkv, kcache, vcache, cache_row_stride, out, H, d);
  note_launch(1);
}

// ============================================================================ K11 cross attention
constexpr int XA_CHUNK = 128;                 // keys per pipeline stage
constexpr int XA_STAGES = 2;                  // x up to 4 CTAs per SM: 128 KB of K/V in flight per SM
constexpr int XA_STAGE_BYTES = XA_CHUNK * 128;  // 64 halves per key
constexpr int XA_NCHUNK = (S_ENC + XA_CHUNK - 1) / XA_CHUNK;  // 12

__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NQ>
__global__ void __launch_bounds__(160) cross_attn_kernel(DecodeState s, PartialSrc q,
                                                         const __half* __restrict__ kc, const __half* __restrict__ vc,
                                                         long slot_stride, float* __restrict__ part, float* __restrict__ probs,
                                                         int rows_per_stream, int H, int d, int nsplit, int cps) {
  extern __shared__ uint8_t xa_smem_raw[];
  uint8_t* base = xa_smem_raw + ((128u - (smem_u32(xa_smem_raw) & 127u)) & 127u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* stage_buf = base;                                           // XA_STAGES x 16 KB
  float* S = reinterpret_cast<float*>(base + XA_STAGES * XA_STAGE_BYTES);  // [cps*128][8]
  float* red = S + (long)cps * XA_CHUNK * 8;                            // [2][4][8] + o-reduce [4][NQ][64]
  float* ored = red + 64;
  uint64_t* full = reinterpret_cast<uint64_t*>(ored + 4 * NQ * 64);
  uint64_t* empty = full + XA_STAGES;

  const int b = blockIdx.z, h = blockIdx.y, sp = blockIdx.x;
  const int c_begin = sp * cps;
  const int c_end = min(XA_NCHUNK, c_begin + cps);
  if (c_begin >= c_end) return;
  const int nchunks = c_end - c_begin;
  const int tid = threadIdx.x, warp = tid >> 5;
  pdl_trigger();

  if (tid == 0) {
    for (int i = 0; i < XA_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 4);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ------------------------------------------------------------------ producer warp
    // The encoder K/V of the slot were written long before this decode step: the first K chunks are requested
    // before the dependency wait, i.e. while the q projection that precedes this kernel is still running.
    if (elect_one()) {
      const long head_off = (long)s.slot[b] * slot_stride + (long)h * S_ENC * 64;
      int stage = 0;
      uint32_t phase = 0;
      bool checked = false;
      for (int pass = 0; pass < 2; ++pass) {
        const __half* src = (pass == 0 ? kc : vc) + head_off;
        for (int c = c_begin; c < c_end; ++c) {
          const int nkeys = min(XA_CHUNK, S_ENC - c * XA_CHUNK);
          if (!checked && (pass == 1 || c - c_begin == XA_STAGES)) {
            // ring is full for the first time: from here on the consumers must be alive
            checked = true;
            pdl_wait();
            if (s.done[b]) {   // stream already finished: drain the requested chunks, then leave
              const int issued = pass == 1 ? nchunks : XA_STAGES;
              for (int i = 0; i < issued; ++i) mbar_wait(&full[i], 0);
              return;
            }
          }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], nkeys * 128);
          bulk_load_1d(stage_buf + stage * XA_STAGE_BYTES, src + (long)c * XA_CHUNK * 64, nkeys * 128, &full[stage]);
          if (++stage == XA_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumer warps (128 threads)
  pdl_wait();
  if (s.done[b]) return;
  const int c8 = tid & 7, g = tid >> 3;
  const int row0 = b * rows_per_stream;
  // q rows of this (stream, head): bias + split-K partial sums in range order, reduced cooperatively (coalesced,
  // 4 ranges in flight) into shared memory, then each thread picks up its 8-wide slice of every row.
  {
    float* qs = ored;   // [NQ][64], reused as the output reduction buffer at the end
    for (int idx = tid; idx < NQ * 64; idx += 128) {
      const int j = idx >> 6, dd = idx & 63;
      float a = 0.f;
      if (j < rows_per_stream) {
        a = q.bias ? __ldg(q.bias + h * 64 + dd) : 0.f;
        const float* qp = q.ptr + (long)(row0 + j) * d + h * 64 + dd;
#pragma unroll 4
        for (int sq = 0; sq < q.nsplit; ++sq) a += __ldcg(qp + (long)sq * q.stride);
      }
      qs[idx] = a * 0.125f;
    }
  }
  consumers_sync();
  float qr[NQ][8];
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const float4 a = *reinterpret_cast<const float4*>(ored + j * 64 + c8 * 8), b2 = *reinterpret_cast<const float4*>(ored + j * 64 + c8 * 8 + 4);
    qr[j][0] = a.x; qr[j][1] = a.y; qr[j][2] = a.z; qr[j][3] = a.w;
    qr[j][4] = b2.x; qr[j][5] = b2.y; qr[j][6] = b2.z; qr[j][7] = b2.w;
  }
  int stage = 0;
  uint32_t phase = 0;
  // pass 1: scores -> S[key][j]
  for (int ci = 0; ci < nchunks; ++ci) {
    const int nkeys = min(XA_CHUNK, S_ENC - (c_begin + ci) * XA_CHUNK);
    mbar_wait(&full[stage], phase);
    const uint8_t* buf = stage_buf + stage * XA_STAGE_BYTES;
    float pr[8][NQ];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = i * 16 + g;
      float kf[8];
      if (key < nkeys) {
        const uint4 u = *reinterpret_cast<const uint4*>(buf + key * 128 + c8 * 16);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          kf[2 * e] = f.x;
          kf[2 * e + 1] = f.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) kf[e] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        float a = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) a = fmaf(kf[e], qr[j][e], a);
        pr[i][j] = a;
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[stage]);  // K bytes are in registers now
    if (++stage == XA_STAGES) { stage = 0; phase ^= 1; }
    // transpose-reduce over the 8 lanes that share a key: lane c8 ends with the full dot of key i == c8
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      float v4[4], v2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float lo = pr[i][j], hi = pr[i + 4][j];
        const float send = (c8 & 4) ? lo : hi;
        const float keep = (c8 & 4) ? hi : lo;
        v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float lo = v4[i], hi = v4[i + 2];
        const float send = (c8 & 2) ? lo : hi;
        const float keep = (c8 & 2) ? hi : lo;
        v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const float send = (c8 & 1) ? v2[0] : v2[1];
      const float keep = (c8 & 1) ? v2[1] : v2[0];
      const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      const int key = c8 * 16 + g;
      S[((long)ci * XA_CHUNK + key) * 8 + j] = key < nkeys ? tot : -INFINITY;
    }
  }
  consumers_sync();
  // softmax statistics over this CTA's key range
  const int nk_pad = nchunks * XA_CHUNK;
  float mx[NQ], sm[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) mx[j] = -INFINITY;
  for (int k = tid; k < nk_pad; k += 128) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) mx[j] = fmaxf(mx[j], S[(long)k * 8 + j]);
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    mx[j] = warp_max(mx[j]);
    if ((tid & 31) == 0) red[warp * 8 + j] = mx[j];
  }
  consumers_sync();
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    mx[j] = fmaxf(fmaxf(red[j], red[8 + j]), fmaxf(red[16 + j], red[24 + j]));
    sm[j] = 0.f;
  }
  for (int k = tid; k < nk_pad; k += 128) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      const float e = __expf(S[(long)k * 8 + j] - mx[j]);
      S[(long)k * 8 + j] = e;
      sm[j] += e;
    }
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    sm[j] = warp_sum(sm[j]);
    if ((tid & 31) == 0) red[32 + warp * 8 + j] = sm[j];
  }
  consumers_sync();
#pragma unroll
  for (int j = 0; j < NQ; ++j) sm[j] = red[32 + j] + red[40 + j] + red[48 + j] + red[56 + j];
  if (probs != nullptr && nsplit == 1) {
    for (int k = tid; k < S_ENC; k += 128) {
#pragma unroll
      for (int j = 0; j < NQ; ++j)
        if (j < rows_per_stream) probs[((long)(row0 + j) * H + h) * S_ENC + k] = S[(long)k * 8 + j] / sm[j];
    }
  }
  // pass 2: o[j][dd] += p[j][key] * V[key][dd]
  float acc[NQ][8];
#pragma unroll
  for (int j = 0; j < NQ; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  for (int ci = 0; ci < nchunks; ++ci) {
    const int nkeys = min(XA_CHUNK, S_ENC - (c_begin + ci) * XA_CHUNK);
    mbar_wait(&full[stage], phase);
    const uint8_t* buf = stage_buf + stage * XA_STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = i * 16 + g;
      if (key < nkeys) {
        const uint4 u = *reinterpret_cast<const uint4*>(buf + key * 128 + c8 * 16);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
        float vf[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          vf[2 * e] = f.x;
          vf[2 * e + 1] = f.y;
        }
        const float* pk = S + ((long)ci * XA_CHUNK + key) * 8;
        const float4 pa = *reinterpret_cast<const float4*>(pk), pb = *reinterpret_cast<const float4*>(pk + 4);
        const float pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(pw[j], vf[e], acc[j][e]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[stage]);
    if (++stage == XA_STAGES) { stage = 0; phase ^= 1; }
  }
  // reduce over the 16 key groups: 4 groups inside a warp (lane bits 3,4), then 4 warps through smem
#pragma unroll
  for (int j = 0; j < NQ; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = acc[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[j][e] = v;
    }
  if ((tid & 31) < 8) {
#pragma unroll
    for (int j = 0; j < NQ; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) ored[(warp * NQ + j) * 64 + c8 * 8 + e] = acc[j][e];
  }
  consumers_sync();
  float* dst = part + (((long)b * H + h) * nsplit + sp) * MAX_ROWS_PER_STREAM * 66;
  for (int idx = tid; idx < NQ * 64; idx += 128) {
    const int j = idx >> 6, dd = idx & 63;
    if (j < rows_per_stream) {
      const float o = ored[(0 * NQ + j) * 64 + dd] + ored[(1 * NQ + j) * 64 + dd] + ored[(2 * NQ + j) * 64 + dd] +
                      ored[(3 * NQ + j) * 64 + dd];
      dst[j * 66 + 2 + dd] = o;
    }
  }
  if (tid < NQ && tid < rows_per_stream) {
    dst[tid * 66 + 0] = mx[tid];
    dst[tid * 66 + 1] = sm[tid];
  }
}

// merge the nsplit partial softmaxes of every (row, head)
__global__ void cross_attn_combine_kernel(DecodeState s, const float* __restrict__ part, __half* __restrict__ out,
                                          int rows_per_stream, int H, int d, int nsplit) {
  const int r = blockIdx.y, h = blockIdx.x, dd = threadIdx.x;
  const int b = r / rows_per_stream, j = r % rows_per_stream;
  pdl_trigger();
  pdl_wait();
  if (s.done[b]) return;
  const float* p = part + (((long)b * H + h) * nsplit) * MAX_ROWS_PER_STREAM * 66 + j * 66;
  float M = -INFINITY;
  for (int sp = 0; sp < nsplit; ++sp) M = fmaxf(M, p[(long)sp * MAX_ROWS_PER_STREAM * 66]);
  float L = 0.f, o = 0.f;
  for (int sp = 0; sp < nsplit; ++sp) {
    const float* ps = p + (long)sp * MAX_ROWS_PER_STREAM * 66;
    const float w = __expf(ps[0] - M);
    L += ps[1] * w;
    o += ps[2 + dd] * w;
  }
  out[(long)r * d + h * 64 + dd] = __float2half_rn(o / L);
}

static int xa_smem_bytes(int cps, int NQ) {
  return 128 + XA_STAGES * XA_STAGE_BYTES + cps * XA_CHUNK * 8 * 4 + (64 + 4 * NQ * 64) * 4 + 2 * XA_STAGES * 8 + 64;
}
static int xa_template_nq(int rows_per_stream) {
  return rows_per_stream == 1 ? 1 : rows_per_stream == 2 ? 2 : rows_per_stream <= 4 ? 4 : rows_per_stream == 5 ? 5 : 8;
}

// Choose how many CTAs share one (stream, head): the grid should fill whole waves of resident CTAs
// (occupancy is set by shared memory -- the score buffer shrinks with the split -- and by registers).
int cross_attn_pick_nsplit(int B, int H, int num_sms, int rows_per_stream) {
  const int NQ = xa_template_nq(rows_per_stream);
  const int regs = NQ <= 2 ? 56 : NQ == 4 ? 96 : NQ == 5 ? 120 : 156;   // ptxas -v
  const int occ_reg = std::max(1, 65536 / (regs * 160));
  int best_ns = 1;
  double best_eff = -1.0;
  for (int ns = 1; ns <= XA_NCHUNK; ++ns) {
    const int cps = (XA_NCHUNK + ns - 1) / ns;
    const int real = (XA_NCHUNK + cps - 1) / cps;
    if (real != ns) continue;
    const int occ = std::max(1, std::min(occ_reg, (227 * 1024) / (xa_smem_bytes(cps, NQ) + 1024)));
    const long slots = (long)occ * num_sms, items = (long)B * H * ns;
    const long waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots) - 0.01 * ns;   // mild preference for fewer partials
    if (eff > best_eff) { best_eff = eff; best_ns = ns; }
  }
  return best_ns;
}

template <int NQ>
static void launch_cross(cudaStream_t st, const DecodeState& s, const PartialSrc& q, const __half* kc, const __half* vc,
                         long slot_stride, const CrossAttnWorkspace& ws, int B, int rows_per_stream, int H, int d, int nsplit) {
  const int cps = (XA_NCHUNK + nsplit - 1) / nsplit;
  const int smem = xa_smem_bytes(cps, NQ);
  dim3 grid(nsplit, H, B);
  launch_kernel(cross_attn_kernel<NQ>, grid, dim3(160), (size_t)smem, st, s, q, kc, vc, slot_stride, ws.part, ws.probs, rows_per_stream, H, d, nsplit, cps);
  note_launch(1);
}

void attention_prime() {
  WL_CUDA(cudaFuncSetAttribute(cross_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  WL_CUDA(cudaFuncSetAttribute(cross_attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  WL_CUDA(cudaFuncSetAttribute(cross_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  WL_CUDA(cudaFuncSetAttribute(cross_attn_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
  WL_CUDA(cudaFuncSetAttribute(cross_attn_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
}

void decoder_cross_attn(cudaStream_t st, const DecodeState& s, const PartialSrc& q, const __half* kc, const __half* vc,
                        long slot_stride, const CrossAttnWorkspace& ws, __half* out, int B, int rows_per_stream, int H,
                        int d, int nsplit) {
  WL_CHECK(rows_per_stream >= 1 && rows_per_stream <= MAX_ROWS_PER_STREAM, WL_ERR_ARG, "rows per stream %d", rows_per_stream);
  if (rows_per_stream == 1) launch_cross<1>(st, s, q, kc, vc, slot_stride, ws, B, rows_per_stream, H, d, nsplit);
  else if (rows_per_stream == 2) launch_cross<2>(st, s, q, kc, vc, slot_stride, ws, B, rows_per_stream, H, d, nsplit);
  else if (rows_per_stream <= 4) launch_cross<4>(st, s, q, kc, vc, slot_stride, ws, B, rows_per_stream, H, d, nsplit);
  else if (rows_per_stream == 5) launch_cross<5>(st, s, q, kc, vc, slot_stride, ws, B, rows_per_stream, H, d, nsplit);
  else launch_cross<8>(st, s, q, kc, vc, slot_stride, ws, B, rows_per_stream, H, d, nsplit);
  dim3 grid(H, B * rows_per_stream);
  launch_kernel(cross_attn_combine_kernel, grid, dim3(64), 0, st, s, ws.part, out, rows_per_stream, H, d, nsplit);
  note_launch(1);
}

}  // namespace wl
