// tcgen05 GEMM: TMA (SWIZZLE_128B) -> smem ring -> tcgen05.mma (single issuing thread) -> TMEM
// accumulator -> tcgen05.ld epilogue (bias / GELU / residual / layout transforms fused).
//
// CTA = 384 threads: warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warps4-11 epilogue
// (lane i of TMEM = accumulator row i; two warps per 32-row group split the columns).  Tile 128 x BN x 64.
#include "gemm.cuh"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

namespace wl {

static std::atomic<long> g_gemm_launches{0};
long gemm_launch_count() { return g_gemm_launches.load(); }

struct GemmKParams {
  int M, N, K;
  int zn1;  // grid z = i1 + zn1 * i2
  int a_batched, b_batched;
  int a_pos[3], b_pos[3];  // tensor-map coordinate slots (1..3) of (row, i1, i2)
  GemmEpilogue e;
  int vec_ok;  // row-major output, 16-byte aligned rows: use vector stores
  int nz;           // batch entries (or K splits in accum mode): tiles = tiles_m * tiles_n * nz
  int accum;        // 1: grid z enumerates K ranges; range s stores its partial sum at out + s * part_stride
  int kb_per_split; // k-blocks per split (accum mode)
  int n_fastest;    // tile order, see tile_decode()
};

constexpr int BM = 128;
constexpr int BK = 64;  // 64 halves = 128 bytes = one swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;

// Epilogue kinds are compile-time so that every kernel instantiation carries exactly one, compact epilogue: the
// decode-step GEMMs run ~200 times per step on a few CTAs each, where instruction fetch of a fat multi-path
// epilogue costs more than its arithmetic.
enum EpiKind : int { EPI_ROW = 0, EPI_COL = 1, EPI_PART = 2, EPI_HEADSPLIT = 3 };

template <int cnt, int KIND>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, long m, int n0, const uint32_t (&v)[cnt], int i1, int i2,
                                               int split = 0, const float4* rpre = nullptr) {
  const GemmEpilogue& e = p.e;
  if (m >= p.M) return;
  if constexpr (KIND == EPI_PART) {
    // split-K: K range `split` stores its raw fp32 partial sum; whoever consumes the result adds the ranges
    // (and the bias) in a fixed order -- no atomics, bit-reproducible.
    float* dst = (float*)e.out + (long)split * e.part_stride + m * e.ldm;
#pragma unroll
    for (int i = 0; i < cnt; ++i) {
      const int n = n0 + i;
      if (n < p.N) dst[(long)n * e.ldn] = __uint_as_float(v[i]);
    }
    return;
  }
  if constexpr (KIND == EPI_HEADSPLIT && cnt >= 8) {
    // m = (b, s), n = (h, dd); one thread writes cnt (<=32) consecutive dd of one head row.  The 16-byte pieces of a
    // 128-byte key row are stored XOR-swizzled by (s & 7): the layout ldmatrix wants in the cross-attention kernel.
    const int b = (int)(m / e.hs_S), s = (int)(m % e.hs_S);
    const int h = n0 >> 6, dd = n0 & 63;
    __half* dst = (__half*)e.out + (long)e.hs_slots[b] * e.hs_slot_stride + ((long)h * e.hs_S + s) * 64;
#pragma unroll
    for (int i = 0; i < cnt; i += 8) {
      if (n0 + i >= p.N) break;
      __align__(16) __half2 h2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = __uint_as_float(v[i + 2 * j]), c = __uint_as_float(v[i + 2 * j + 1]);
        if (e.bias) {
          a += e.bias[n0 + i + 2 * j];
          c += e.bias[n0 + i + 2 * j + 1];
        }
        h2[j] = __floats2half2_rn(a, c);
      }
      *reinterpret_cast<uint4*>(dst + ((((dd + i) >> 3) ^ (s & 7)) << 3)) = *reinterpret_cast<const uint4*>(h2);
    }
    return;
  }
  const long obase = (long)i1 * e.ob1 + (long)i2 * e.ob2 + m * e.ldm;
  const long rbase = (long)i1 * e.rb1 + (long)i2 * e.rb2 + m * e.rldm;
  const float bm = (e.bias && e.bias_on_m) ? e.bias[m] : 0.f;
  if constexpr (KIND == EPI_ROW && cnt >= 8) {
#pragma unroll
    for (int i = 0; i < cnt; i += 8) {
      const int n = n0 + i;
      if (n >= p.N) break;
      if (n + 8 > p.N) {  // ragged tail of the last tile: element-wise
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (n + j < p.N) {
            float x1 = __uint_as_float(v[i + j]) + bm;
            if (e.bias && !e.bias_on_m) x1 += e.bias[n + j];
            if (e.gelu) x1 = gelu_erf(x1);
            if (e.resid) x1 += e.resid[rbase + n + j];
            if (e.out_f32) ((float*)e.out)[obase + n + j] = x1;
            else ((__half*)e.out)[obase + n + j] = __float2half_rn(x1);
          }
        }
        continue;
      }
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[i + j]) + bm;
      if (e.bias && !e.bias_on_m) {
        const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n), b1 = *reinterpret_cast<const float4*>(e.bias + n + 4);
        x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
        x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
      }
      if (e.gelu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
      }
      if (e.resid) {
        const float* r = e.resid + rbase + n;
        const float4 r0 = rpre ? rpre[i / 4] : *reinterpret_cast<const float4*>(r), r1 = rpre ? rpre[i / 4 + 1] : *reinterpret_cast<const float4*>(r + 4);
        x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
        x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
      }
      if (e.out_f32) {
        float* o = (float*)e.out + obase + n;
        *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
      } else {
        __align__(16) __half2 h2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
        *reinterpret_cast<uint4*>((__half*)e.out + obase + n) = *reinterpret_cast<const uint4*>(h2);
      }
    }
    return;
  }
  if constexpr (KIND == EPI_COL || cnt < 8) {
#pragma unroll
  for (int i = 0; i < cnt; ++i) {
    const int n = n0 + i;
    if (n >= p.N) break;
    float x = __uint_as_float(v[i]) + bm;
    if (e.bias && !e.bias_on_m) x += e.bias[n];
    if (e.gelu) x = gelu_erf(x);
    if (e.resid) x += e.resid[rbase + (long)n * e.rldn];
    const long o = obase + (long)n * e.ldn;
    if (e.out_f32) ((float*)e.out)[o] = x;
    else ((__half*)e.out)[o] = __float2half_rn(x);
  }
  }
}

// Sense-reversing grid barrier for a grid whose CTAs are all resident; called by ONE thread per CTA.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks) {
  const unsigned gen = ld_acquire_u32(bar + 1);
  __threadfence();
  if (atomicAdd(bar, 1u) == nblocks - 1u) {
    atomicExch(bar, 0u);
    __threadfence();
    atomicAdd(bar + 1, 1u);
  } else {
    while (ld_acquire_u32(bar + 1) == gen) __nanosleep(20);
  }
  __threadfence();
}
template <int NT>
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory"); }

// Row-wise consumer of the split-K partial sums, run by the NT epilogue threads of every CTA after the grid barrier
// (same arithmetic, in the same order, as layernorm_update_kernel / gelu_cast_kernel).
template <int NT>
__device__ __forceinline__ void post_op(const GemmKParams& p, int te, float* red /*[16]*/) {
  const GemmEpilogue& e = p.e;
  const int S = e.partials, R = p.N, F = p.M;   // K ranges, rows, features
  const float* part = reinterpret_cast<const float*>(e.out);
  if (e.post == GEMM_POST_LN) {
    const int n4 = F >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(e.post_g);
    const float4* b4 = reinterpret_cast<const float4*>(e.post_b);
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
      float4* x4 = reinterpret_cast<float4*>(e.post_x + (long)r * F);
      constexpr int PER = 3;   // float4 per thread: d <= 4 * PER * NT
      float4 v[PER];
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = i * NT + te;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < n4) {
          if (e.post_bias) v[i] = __ldg(reinterpret_cast<const float4*>(e.post_bias) + c);
          const float4 a = x4[c];
          v[i].x += a.x; v[i].y += a.y; v[i].z += a.z; v[i].w += a.w;
        }
      }
#pragma unroll 8
      for (int sp = 0; sp < S; ++sp) {
        const float4* p4 = reinterpret_cast<const float4*>(part + (long)sp * e.part_stride + (long)r * F);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
          const int c = i * NT + te;
          if (c < n4) {
            const float4 q = __ldcg(p4 + c);
            v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = i * NT + te;
        if (c < n4) x4[c] = v[i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      sum = warp_sum(sum);
      if ((te & 31) == 0) red[te >> 5] = sum;
      epi_sync<NT>();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tot += red[w];
      const float mean = tot / F;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        if (i * NT + te < n4) {
          const float a = v[i].x - mean, b2 = v[i].y - mean, c2 = v[i].z - mean, d2 = v[i].w - mean;
          sq += (a * a + b2 * b2) + (c2 * c2 + d2 * d2);
        }
      }
      sq = warp_sum(sq);
      if ((te & 31) == 0) red[8 + (te >> 5)] = sq;
      epi_sync<NT>();
      float tq = 0.f;
#pragma unroll
      for (int w = 0; w < NT / 32; ++w) tq += red[8 + w];
      const float rstd = rsqrtf(tq / F + 1e-5f);
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int c = i * NT + te;
        if (c < n4) {
          const float4 gg = g4[c], bb = b4[c];
          __align__(8) __half2 h[2] = {
              __floats2half2_rn((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y),
              __floats2half2_rn((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w)};
          *reinterpret_cast<uint2*>(e.post_y + (long)r * F + 4 * c) = *reinterpret_cast<const uint2*>(h);
        }
      }
      epi_sync<NT>();   // red is reused by the next row
    }
  } else if (e.post == GEMM_POST_GELU) {
    const int c4n = F >> 2;
    const long total4 = (long)R * c4n;
    const long st4 = e.part_stride >> 2;
    for (long i4 = (long)blockIdx.x * NT + te; i4 < total4; i4 += (long)gridDim.x * NT) {
      const int c = (int)(i4 % c4n);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e.post_bias) v = __ldg(reinterpret_cast<const float4*>(e.post_bias) + c);
      const float4* p4 = reinterpret_cast<const float4*>(part) + i4;
#pragma unroll 4
      for (int sp = 0; sp < S; ++sp) {
        const float4 q = __ldcg(p4 + sp * st4);
        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
      }
      __align__(8) __half2 h[2] = {__floats2half2_rn(gelu_erf(v.x), gelu_erf(v.y)), __floats2half2_rn(gelu_erf(v.z), gelu_erf(v.w))};
      *reinterpret_cast<uint2*>(e.post_y + 4 * i4) = *reinterpret_cast<const uint2*>(h);
    }
  }
}

// Tile order inside one batch entry.  m fastest: CTAs running side by side share the B tile (right when B is the big
// operand).  n fastest (p.n_fastest): they share the A tile and sweep B -- right when B is a weight matrix that stays
// in L2 anyway and A is a large activation (FC2 of the encoder: A = 123 MB would otherwise be re-read per n tile).
__device__ __forceinline__ void tile_decode(const GemmKParams& p, int t, int tiles_m, int tiles_n, int& tile_m, int& tile_n, int& zz) {
  if (p.n_fastest) {
    tile_n = t % tiles_n;
    const int r = t / tiles_n;
    tile_m = r % tiles_m;
    zz = r / tiles_m;
  } else {
    tile_m = t % tiles_m;
    const int r = t / tiles_m;
    tile_n = r % tiles_n;
    zz = r / tiles_n;
  }
}

// Persistent: each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  (tile_m fastest, so CTAs
// running side by side share the B (weight) tile in L2).  Two TMEM accumulator stages: the epilogue warps drain
// tile i while the MMA warp already accumulates tile i+1.
template <int BN, int STAGES, int MIN_CTAS, int KIND>
__global__ void __launch_bounds__(384, MIN_CTAS)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  uint8_t* sA = base;
  uint8_t* sB = base + STAGES * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* post_red = reinterpret_cast<float*>(tmem_slot + 2);   // [16]

  const int warp = threadIdx.x >> 5;
  const int tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int total_tiles = tiles_m * tiles_n * p.nz;
  const int total_kb = (p.K + BK - 1) / BK;
  pdl_trigger();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], BN >= 64 ? 8 : 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      // coordinate slot s (1..3) of a tensor map holds whichever of (row, i1, i2) was sorted there
      auto slot = [](const int (&pos)[3], int s, int row, int j1, int j2) {
        return pos[0] == s ? row : (pos[1] == s ? j1 : (pos[2] == s ? j2 : 0));
      };
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int tile_m, tile_n, zz;
        tile_decode(p, t, tiles_m, tiles_n, tile_m, tile_n, zz);
        const int z = KIND == EPI_PART ? 0 : zz, split = KIND == EPI_PART ? zz : 0;
        const int i1 = z % p.zn1, i2 = z / p.zn1;
        const int kb0 = KIND == EPI_PART ? split * p.kb_per_split : 0;
        const int num_kb = KIND == EPI_PART ? min(p.kb_per_split, total_kb - kb0) : total_kb;
        const int a1 = p.a_batched ? i1 : 0, a2 = p.a_batched ? i2 : 0;
        const int b1 = p.b_batched ? i1 : 0, b2 = p.b_batched ? i2 : 0;
        const int ca1 = slot(p.a_pos, 1, tile_m * BM, a1, a2), ca2 = slot(p.a_pos, 2, tile_m * BM, a1, a2),
                  ca3 = slot(p.a_pos, 3, tile_m * BM, a1, a2);
        const int cb1 = slot(p.b_pos, 1, tile_n * BN, b1, b2), cb2 = slot(p.b_pos, 2, tile_n * BN, b1, b2),
                  cb3 = slot(p.b_pos, 3, tile_n * BN, b1, b2);
        int pre = 0;
        if (first) {
          // First tile of this CTA.  When A is a weight matrix (decode: swap-AB, A = W) its k-blocks are requested
          // BEFORE the dependency wait, so the weight stream overlaps the tail of the kernel that produces B.
          if (p.e.a_static) {
            pre = min(num_kb, STAGES);
            for (int kb = 0; kb < pre; ++kb) {
              mbar_expect_tx(&full[kb], A_STAGE_BYTES + B_STAGE_BYTES);
              tma_load_4d(sA + kb * A_STAGE_BYTES, &tmA, &full[kb], (kb0 + kb) * BK, ca1, ca2, ca3);
            }
          }
          if (blockIdx.x == 0) tl_stamp_any(KIND == EPI_PART ? TL_GEMM_PART : TL_GEMM, 0);
          pdl_wait();
          if (blockIdx.x == 0) tl_stamp_any(KIND == EPI_PART ? TL_GEMM_PART : TL_GEMM, 1);
          first = false;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          const int k0 = (kb0 + kb) * BK;
          if (kb < pre) {
            tma_load_4d(sB + stage * B_STAGE_BYTES, &tmB, &full[stage], k0, cb1, cb2, cb3);
          } else {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], A_STAGE_BYTES + B_STAGE_BYTES);
            tma_load_4d(sA + stage * A_STAGE_BYTES, &tmA, &full[stage], k0, ca1, ca2, ca3);
            tma_load_4d(sB + stage * B_STAGE_BYTES, &tmB, &full[stage], k0, cb1, cb2, cb3);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int zz = t / (tiles_m * tiles_n);
        const int kb0 = KIND == EPI_PART ? zz * p.kb_per_split : 0;
        const int num_kb = KIND == EPI_PART ? min(p.kb_per_split, total_kb - kb0) : total_kb;
        mbar_wait(&acc_empty[as], aphase ^ 1);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t acc = tmem_acc + as * ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 halves = 32 bytes along K inside the 128-byte swizzle row: +2 in (addr>>4) units
            umma_f16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4 && (BN >= 64 || warp < 8)) {
    // 8 epilogue warps (2 per SM sub-partition): warps 4-7 drain the left half of the accumulator columns,
    // warps 8-11 the right half; narrow tiles (BN < 64) use warps 4-7 only.
    const int q = warp & 3;
    constexpr int NH = BN >= 64 ? 2 : 1, HC = BN / NH;   // column halves, columns per half
    const int c_lo = ((warp - 4) >> 2) * HC;
    int as = 0;
    uint32_t aphase = 0;
    pdl_wait();   // the residual / output buffers belong to the preceding kernels
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int tile_m, tile_n, zz;
        tile_decode(p, t, tiles_m, tiles_n, tile_m, tile_n, zz);
      const int z = KIND == EPI_PART ? 0 : zz, split = KIND == EPI_PART ? zz : 0;
      const int i1 = z % p.zn1, i2 = z / p.zn1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const long m = (long)tile_m * BM + q * 32 + lane_id();
      const uint32_t lane_addr = tmem_acc + as * ACC_COLS + ((uint32_t)(q * 32) << 16);
      if constexpr (BN >= 32) {
#pragma unroll 1
        for (int c = c_lo; c < c_lo + HC; c += 32) {
          uint32_t v[32];
          // fp32 residual of this thread's 32 columns: requested before the accumulator read so that the 8 loads
          // are in flight together (out may alias resid, which keeps the compiler from hoisting them itself)
          float4 rr[8];
          bool rr_ok = false;
          if constexpr (KIND == EPI_ROW) {
            const int n0 = tile_n * BN + c;
            if (p.e.resid != nullptr && m < p.M && n0 + 32 <= p.N) {
              const float4* r4 = reinterpret_cast<const float4*>(p.e.resid + (long)i1 * p.e.rb1 + (long)i2 * p.e.rb2 + m * p.e.rldm + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) rr[j] = r4[j];
              rr_ok = true;
            }
          }
          tmem_ld_32x32(lane_addr + c, v);
          tmem_ld_wait();
          if (c + 32 >= c_lo + HC) {  // this warp's columns are in registers: hand the TMEM stage back before the stores
            tc_fence_before();
            __syncwarp();
            if (lane_id() == 0) mbar_arrive(&acc_empty[as]);
          }
          if constexpr (KIND == EPI_ROW) {
            if (rr_ok) epilogue_chunk<32, KIND>(p, m, tile_n * BN + c, v, i1, i2, split, rr);
            else epilogue_chunk<32, KIND>(p, m, tile_n * BN + c, v, i1, i2, split);
          } else {
            epilogue_chunk<32, KIND>(p, m, tile_n * BN + c, v, i1, i2, split);
          }
        }
      } else {
        uint32_t v[16];
        tmem_ld_32x16(lane_addr, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) mbar_arrive(&acc_empty[as]);
        epilogue_chunk<16, KIND>(p, m, tile_n * BN, v, i1, i2, split);
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if constexpr (KIND == EPI_PART) {
      if (p.e.post != GEMM_POST_NONE) {
        constexpr int NT = BN >= 64 ? 256 : 128;   // epilogue threads of this configuration
        const int te = threadIdx.x - 128;
        __threadfence();                           // this thread's partial sums are visible device-wide
        epi_sync<NT>();
        if (te == 0) grid_barrier(p.e.post_bar, gridDim.x);
        epi_sync<NT>();
        post_op<NT>(p, te, post_red);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_acc, TMEM_COLS);
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  });
  WL_CHECK(fn != nullptr, WL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

struct TmapKey {
  const void* ptr;
  long rows, k, ld, s1, s2;
  int n1, n2, box_rows, box_k;
  bool operator<(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) < 0; }
};

// Tensor map + where the (row, i1, i2) coordinates sit among its dims 1..3.  TMA wants strides in
// ascending order (each a multiple of the previous), so the view's dims are sorted by stride: e.g. the
// per-head Q/K operand of attention is {k, head (128 B), row (2*ld B), batch}.
static TmapInfo get_tmap(const GemmOperand& op, int box_rows, int box_k = BK);
TmapInfo make_tmap(const GemmOperand& op, int box_rows, int box_k) { return get_tmap(op, box_rows, box_k); }

static TmapInfo get_tmap(const GemmOperand& op, int box_rows, int box_k) {
  static std::map<TmapKey, TmapInfo> cache;
  static std::mutex mu;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = op.ptr; key.rows = op.rows; key.k = op.k; key.ld = op.ld; key.s1 = op.s1; key.s2 = op.s2;
  key.n1 = op.n1; key.n2 = op.n2; key.box_rows = box_rows; key.box_k = box_k;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  WL_CHECK(((uintptr_t)op.ptr & 15) == 0, WL_ERR_ARG, "GEMM operand pointer must be 16-byte aligned");
  struct D { long size, stride; int role; };
  std::vector<D> real, single;
  const D all[3] = {{op.rows, op.ld, 0}, {op.n1, op.s1, 1}, {op.n2, op.s2, 2}};
  for (const D& d : all) ((d.role == 0 || d.size > 1) ? real : single).push_back(d);
  std::stable_sort(real.begin(), real.end(), [](const D& x, const D& y) { return x.stride < y.stride; });
  std::vector<D> order = real;
  for (D d : single) {
    d.stride = order.back().stride * order.back().size;
    order.push_back(d);
  }
  TmapInfo info;
  cuuint64_t dims[4] = {(cuuint64_t)op.k, 1, 1, 1};
  cuuint64_t strides[3];
  cuuint32_t box[4] = {(cuuint32_t)box_k, 1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    WL_CHECK(order[i].stride > 0 && (order[i].stride * 2) % 16 == 0, WL_ERR_ARG,
             "GEMM operand stride %ld elements (dim role %d) is not a positive multiple of 16 bytes", order[i].stride, order[i].role);
    dims[1 + i] = (cuuint64_t)order[i].size;
    strides[i] = (cuuint64_t)order[i].stride * 2;
    if (order[i].role == 0) box[1 + i] = (cuuint32_t)box_rows;
    info.pos[order[i].role] = 1 + i;
  }
  CUresult r = encode_fn()(&info.tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (void*)op.ptr, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  WL_CHECK(r == CUDA_SUCCESS, WL_ERR_CUDA,
           "cuTensorMapEncodeTiled failed (%d) dims={%llu,%llu,%llu,%llu} strides={%llu,%llu,%llu} box_rows=%d", (int)r,
           (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
           (unsigned long long)strides[0], (unsigned long long)strides[1], (unsigned long long)strides[2], box_rows);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, info);
  return info;
}

template <int BN, int STAGES, int MIN_CTAS, int KIND>
static void launch_cfg(cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p, int Z) {
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 512;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  GemmKParams q = p;
  q.nz = Z;
  const long tiles = (long)cdiv(p.N, BN) * cdiv(p.M, BM) * Z;
  const int grid = (int)std::min<long>(tiles, (long)sms * MIN_CTAS);
  launch_kernel(gemm_tn_kernel<BN, STAGES, MIN_CTAS, KIND>, dim3(grid), dim3(384), (size_t)smem, stream, ta, tb, q);
  g_gemm_launches++;
}

template <int BN, int STAGES, int MIN_CTAS, int KIND>
static void prime_cfg() {
  constexpr int smem = STAGES * (A_STAGE_BYTES + BN * BK * 2) + 1024 + 512;
  WL_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, STAGES, MIN_CTAS, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
}

// opt-in shared memory sizes must be set outside stream capture: done once from wl_init
template <int KIND>
static void prime_kind() {
  prime_cfg<16, 8, 1, KIND>();
  prime_cfg<32, 8, 1, KIND>();
  prime_cfg<64, 8, 1, KIND>();
  prime_cfg<128, 5, 1, KIND>();
  prime_cfg<256, 4, 1, KIND>();
}
void gemm_tl_bind(unsigned long long* p) { tl_bind_tu(p); }
void gemm_prime() {
  prime_kind<EPI_ROW>();
  prime_kind<EPI_COL>();
  prime_kind<EPI_PART>();
  prime_kind<EPI_HEADSPLIT>();
}

template <int KIND>
static void launch_kind(int bn, cudaStream_t stream, const CUtensorMap& ta, const CUtensorMap& tb, const GemmKParams& p, int Z) {
  switch (bn) {
    case 16: launch_cfg<16, 8, 1, KIND>(stream, ta, tb, p, Z); break;
    case 32: launch_cfg<32, 8, 1, KIND>(stream, ta, tb, p, Z); break;
    case 64: launch_cfg<64, 8, 1, KIND>(stream, ta, tb, p, Z); break;
    case 128: launch_cfg<128, 5, 1, KIND>(stream, ta, tb, p, Z); break;
    case 256: launch_cfg<256, 4, 1, KIND>(stream, ta, tb, p, Z); break;
    default: WL_CHECK(false, WL_ERR_ARG, "unsupported BN %d", bn);
  }
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

static GemmKParams make_params(const GemmOperand& A, const GemmOperand& B, int M, int N, int K, const GemmEpilogue& epi,
                               int* Z) {
  WL_CHECK(M > 0 && N > 0 && K > 0, WL_ERR_ARG, "gemm_tn: empty problem %dx%dx%d", M, N, K);
  const int za = A.n1 * A.n2, zb = B.n1 * B.n2;
  WL_CHECK(za == 1 || zb == 1 || (A.n1 == B.n1 && A.n2 == B.n2), WL_ERR_ARG, "gemm_tn: batch shapes differ");
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.a_batched = za > 1; p.b_batched = zb > 1;
  p.zn1 = za > 1 ? A.n1 : B.n1;
  *Z = za > zb ? za : zb;
  p.e = epi;
  p.vec_ok = 0;
  p.accum = 0;
  // B (N x K halves) small enough to live in L2 while A is larger than B: sweep n fastest
  p.n_fastest = (zb == 1 && (long)N * K * 2 <= (24L << 20) && (long)M * K > (long)N * K) ? 1 : 0;
  p.kb_per_split = 0;
  if (epi.mode == GEMM_STORE && epi.ldn == 1) {
    const int a = epi.out_f32 ? 4 : 8;  // elements per 16 bytes
    bool ok = ((uintptr_t)epi.out & 15) == 0 && epi.ldm % 8 == 0 && epi.ob1 % 8 == 0 && epi.ob2 % 8 == 0;
    (void)a;
    if (epi.bias && !epi.bias_on_m) ok = ok && ((uintptr_t)epi.bias & 15) == 0;
    if (epi.resid)
      ok = ok && epi.rldn == 1 && ((uintptr_t)epi.resid & 15) == 0 && epi.rldm % 4 == 0 && epi.rb1 % 4 == 0 && epi.rb2 % 4 == 0;
    p.vec_ok = ok ? 1 : 0;
  }
  if (epi.mode == GEMM_HEADSPLIT) {
    WL_CHECK(!epi.out_f32 && !epi.gelu && !epi.resid && !epi.bias_on_m && N % 64 == 0 && epi.hs_slots, WL_ERR_ARG,
             "gemm_tn: bad head-split epilogue");
  }
  return p;
}

// Number of K ranges for a weight-streaming (swap-AB) GEMM so that tiles x ranges fills the SMs; always a value
// gemm_tn accepts (every range non-empty).
int gemm_split_plan(int M, int N, int K) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int bn = N <= 16 ? 16 : N <= 32 ? 32 : N <= 64 ? 64 : 128;
  const int tiles = cdiv(N, bn) * cdiv(M, BM), total_kb = cdiv(K, BK);
  int s = std::max(1, std::min(std::min(total_kb, 8), sms / std::max(1, tiles)));
  const int kbs = cdiv(total_kb, s);
  return cdiv(total_kb, kbs);
}

void gemm_tn(cudaStream_t stream, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, const GemmEpilogue& epi) {
  static const int force_simt = env_int("WLB200_GEMM_SIMT", 0);
  if (force_simt) return gemm_tn_simt(stream, A, B, M, N, K, epi);
  int Z;
  GemmKParams p = make_params(A, B, M, N, K, epi, &Z);
  static const int force_bn = env_int("WLB200_BN", 0);
  int bn;
  if (force_bn) bn = force_bn;
  else if (N <= 16) bn = 16;
  else if (N <= 32) bn = 32;
  else if (N <= 64) bn = 64;
  else if (N >= 512 && M >= 512 && epi.ldn == 1) bn = 256;
  else bn = 128;
  if (epi.mode == GEMM_HEADSPLIT && bn < 64) bn = 64;
  if (epi.partials > 0) {
    WL_CHECK(Z == 1 && epi.out_f32 && !epi.gelu && !epi.resid && !epi.bias && epi.mode == GEMM_STORE, WL_ERR_ARG,
             "gemm_tn: split-K partial output must be plain fp32 without bias");
  } else {
    WL_CHECK(epi.post == GEMM_POST_NONE, WL_ERR_ARG, "gemm_tn: a fused post-op needs the split-K partial output");
  }
  if (epi.partials > 0) {
    if (epi.post != GEMM_POST_NONE) {
      WL_CHECK(epi.post_bar && epi.post_y && M % 4 == 0 && epi.part_stride % 4 == 0 && epi.ldm == 1 && epi.ldn == M, WL_ERR_ARG,
               "gemm_tn: fused post-op needs the [range][row][feature] partial layout");
      WL_CHECK(epi.post != GEMM_POST_LN || (epi.post_x && epi.post_g && epi.post_b && M <= 4 * 3 * 128), WL_ERR_ARG,
               "gemm_tn: fused LayerNorm supports rows of at most 1536 features");
    }
    const int total_kb = cdiv(K, BK);
    p.accum = 1;
    p.kb_per_split = cdiv(total_kb, epi.partials);
    Z = cdiv(total_kb, p.kb_per_split);
    WL_CHECK(Z == epi.partials, WL_ERR_ARG, "gemm_tn: %d K ranges cannot be formed from %d k-blocks (use gemm_split_plan)", epi.partials, total_kb);
  }
  const TmapInfo ia = get_tmap(A, BM), ib = get_tmap(B, bn);
  const CUtensorMap& ta = ia.tm;
  const CUtensorMap& tb = ib.tm;
  for (int i = 0; i < 3; ++i) { p.a_pos[i] = ia.pos[i]; p.b_pos[i] = ib.pos[i]; }
  if (p.accum) launch_kind<EPI_PART>(bn, stream, ta, tb, p, Z);
  else if (epi.mode == GEMM_HEADSPLIT) launch_kind<EPI_HEADSPLIT>(bn, stream, ta, tb, p, Z);
  else if (p.vec_ok) launch_kind<EPI_ROW>(bn, stream, ta, tb, p, Z);
  else launch_kind<EPI_COL>(bn, stream, ta, tb, p, Z);
}

// ------------------------------------------------------------------------------------ SIMT reference
__global__ void gemm_tn_simt_kernel(GemmOperand A, GemmOperand B, GemmKParams p) {
  const long n = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long m = blockIdx.y;
  const int z = blockIdx.z;
  if (n >= p.N) return;
  const int i1 = z % p.zn1, i2 = z / p.zn1;
  const __half* a = A.ptr + (p.a_batched ? (long)i1 * A.s1 + (long)i2 * A.s2 : 0) + m * A.ld;
  const __half* b = B.ptr + (p.b_batched ? (long)i1 * B.s1 + (long)i2 * B.s2 : 0) + n * B.ld;
  float acc = 0.f;
  const int ka = (int)(A.k < p.K ? A.k : p.K), kb = (int)(B.k < p.K ? B.k : p.K);
  const int kk = ka < kb ? ka : kb;
  for (int k = 0; k < kk; ++k) acc = fmaf(__half2float(a[k]), __half2float(b[k]), acc);
  uint32_t v[1] = {__float_as_uint(acc)};
  if (p.e.partials > 0) {  // same contract as the split-K path: range 0 carries the sum, the others zero
    for (int sp = 0; sp < p.e.partials; ++sp)
      ((float*)p.e.out)[(long)sp * p.e.part_stride + m * p.e.ldm + n * p.e.ldn] = sp == 0 ? acc : 0.f;
    return;
  }
  GemmKParams q = p;
  q.vec_ok = 0;
  if (q.e.mode == GEMM_HEADSPLIT) {
    const GemmEpilogue& e = q.e;
    const int bb = (int)(m / e.hs_S), s = (int)(m % e.hs_S);
    float x = acc + (e.bias ? e.bias[n] : 0.f);
    ((__half*)e.out)[(long)e.hs_slots[bb] * e.hs_slot_stride + ((long)(n >> 6) * e.hs_S + s) * 64 + (((((int)n & 63) >> 3) ^ (s & 7)) << 3) + (n & 7)] = __float2half_rn(x);
    return;
  }
  epilogue_chunk<1, EPI_COL>(q, m, (int)n, v, i1, i2);
}

void gemm_tn_simt(cudaStream_t stream, const GemmOperand& A, const GemmOperand& B, int M, int N, int K, const GemmEpilogue& epi) {
  int Z;
  GemmKParams p = make_params(A, B, M, N, K, epi, &Z);
  dim3 grid(cdiv(N, 128), M, Z);
  gemm_tn_simt_kernel<<<grid, 128, 0, stream>>>(A, B, p);
  WL_CUDA(cudaGetLastError());
  g_gemm_launches++;
}

}  // namespace wl
