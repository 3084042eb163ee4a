// K5 (v2): fused encoder self-attention on tcgen05 -- scores never leave the SM.
//   one CTA = 128 queries of one (stream, head); loop over 12 key tiles of 128:
//     S = Q K^T            tcgen05.mma 128x128x16 x4  -> TMEM (three buffers rotating over the tiles)
//     online softmax       thread = query row: tcgen05.ld S, exp2, write P (fp16) into smem in the
//                          128B-swizzled K-major layout the tensor core reads
//     O_blk = P V          tcgen05.mma 128x64x16 x8   -> TMEM; accumulated into registers with the lazy rescale
//   warp 0: TMA producer (Q once, K / V^T tiles through a 4-stage mbarrier ring), warp 1: MMA issuer,
//   warp 2: TMEM allocator, warps 4-7 / 8-11: two softmax groups taking the even / odd key tiles with independent
//   running (max, sum, O), merged at the end.
#include <cstdlib>

#include "gemm.cuh"
#include "kernels.cuh"

namespace wl {

constexpr int FA_BQ = 128, FA_BK = 128, FA_STAGES = 4;
constexpr int FA_Q_BYTES = FA_BQ * 128;              // 128 rows x 64 halves
constexpr int FA_K_BYTES = FA_BK * 128;              // 128 keys x 64 halves
constexpr int FA_V_BYTES = 2 * 64 * 128;             // two k-blocks of [64 dd][64 keys]
constexpr int FA_P_BYTES = 2 * FA_BQ * 128;          // two k-blocks of [128 rows][64 keys]
constexpr int FA_STAGE_BYTES = FA_K_BYTES + FA_V_BYTES;
constexpr int FA_SMEM = 1024 + FA_Q_BYTES + FA_STAGES * FA_STAGE_BYTES + 2 * FA_P_BYTES + 256;   // P is double buffered
constexpr int FA_NT = (S_ENC + FA_BK - 1) / FA_BK;   // 12 key tiles

struct FaParams {
  int q_pos[3], k_pos[3], v_pos[3];
  __half* out;   // [B*1500][d]
  int d;
  float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ float fa_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void softmax_sync() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

__device__ __forceinline__ void fa_coords(int (&c)[4], const int (&pos)[3], int k0, int row, int i1, int i2) {
  c[0] = k0; c[1] = c[2] = c[3] = 0;
  c[pos[0]] = row; c[pos[1]] = i1; c[pos[2]] = i2;
}

// One thread's row of NCH x 16 TMEM columns, double buffered in registers: the tcgen05.ld of chunk c + 1 is in flight while
// chunk c is processed (tcgen05.wait::ld retires every load issued so far, so the wait sits AFTER the work on the
// previous chunk).  With one load + wait + work per chunk the softmax warps spent ~1500 of ~2900 clocks per key tile
// stalled on TMEM latency (10 exposed loads per tile, long-scoreboard stalls 1.8 per issue: profiles/prof_flash_r2b.md).
template <int NCH, class F>
__device__ __forceinline__ void tmem_sweep16(uint32_t base, F&& body) {
  uint32_t va[16], vb[16];
  tmem_ld_32x16(base, va);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < NCH; c += 2) {
    tmem_ld_32x16(base + (c + 1) * 16, vb);
    body(va, c);
    tmem_ld_wait();
    if (c + 2 < NCH) tmem_ld_32x16(base + (c + 2) * 16, va);
    body(vb, c + 1);
    if (c + 2 < NCH) tmem_ld_wait();
  }
}

__global__ void __launch_bounds__(384, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t fa_raw[];
  uint8_t* base = fa_raw + ((1024u - (smem_u32(fa_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* sQ = base;
  uint8_t* sKV = sQ + FA_Q_BYTES;
  uint8_t* sP = sKV + FA_STAGES * FA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                 // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;     // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;      // [3]
  uint64_t* s_empty = s_full + 3;               // [3]
  uint64_t* p_full = s_empty + 3;               // [2]
  uint64_t* o_full = p_full + 2;                // [2]
  uint64_t* o_empty = o_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM (512 columns): three S buffers rotating over the key tiles (tile j -> S[j % 3]) and one O_blk per softmax group.
  // With one S buffer per group a group that had finished tile j sat out the whole issue + execution of Q K^T (j + 2)
  // (long-scoreboard stalls 1.9 per issue, eligible warps 0.46 per scheduler: profiles/prof_flash_r2.md); with three, the
  // buffer tile j + 2 needs is the one tile j - 1 (the OTHER group, one tile ahead) releases, so S (j + 2) is computed
  // while tile j is still in its exp pass.  P (shared memory) and O_blk stay per group.
  const uint32_t tmem_O[2] = {tmem + 384, tmem + 448};

  if (warp == 0) {
    if (elect_one()) {
      int c[4];
      mbar_expect_tx(q_full, FA_Q_BYTES);
      fa_coords(c, p.q_pos, 0, qt * FA_BQ, h, b);
      tma_load_4d(sQ, &tmQ, q_full, c[0], c[1], c[2], c[3]);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < FA_NT; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + stage * FA_STAGE_BYTES;
        mbar_expect_tx(&kv_full[stage], FA_STAGE_BYTES);
        fa_coords(c, p.k_pos, 0, j * FA_BK, h, b);
        tma_load_4d(sk, &tmK, &kv_full[stage], c[0], c[1], c[2], c[3]);
        fa_coords(c, p.v_pos, j * FA_BK, 0, h, b);
        tma_load_4d(sk + FA_K_BYTES, &tmV, &kv_full[stage], c[0], c[1], c[2], c[3]);
        fa_coords(c, p.v_pos, j * FA_BK + 64, 0, h, b);
        tma_load_4d(sk + FA_K_BYTES + 64 * 128, &tmV, &kv_full[stage], c[0], c[1], c[2], c[3]);
        if (++stage == FA_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(FA_BQ, FA_BK);
      constexpr uint32_t idesc_o = umma_idesc_f16(FA_BQ, 64);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      auto issue_pv = [&](int i) {
        const int st = i % FA_STAGES, pb = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + pb * FA_P_BYTES));
        mbar_wait(&p_full[pb], par);              // P_i written by the softmax warps
        mbar_wait(&o_empty[pb], par ^ 1);         // O_blk of tile i-2 (same buffer) consumed
        tc_fence_after();
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * FA_STAGE_BYTES + FA_K_BYTES));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a = pdesc + (uint64_t)((kk >> 2) * (FA_BQ * 128 >> 4) + (kk & 3) * 2);
          const uint64_t bd = vdesc + (uint64_t)((kk >> 2) * (64 * 128 >> 4) + (kk & 3) * 2);
          umma_f16(tmem_O[pb], a, bd, idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(&o_full[pb]);
        umma_commit(&kv_empty[st]);
      };
      auto issue_qk = [&](int j) {
        const int st = j % FA_STAGES, sb = j % 3;
        mbar_wait(&kv_full[st], (uint32_t)(j / FA_STAGES) & 1u);
        mbar_wait(&s_empty[sb], ((uint32_t)(j / 3) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sKV + st * FA_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem + sb * 128, qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[sb]);
      };
      // Q K^T runs two tiles ahead of P V: S (j + 2) is issued BEFORE the issuer blocks on P (j), so it is computed while
      // the softmax groups are still busy with tiles j and j + 1 (issued after P V (j - 1), as before, every group waited
      // out the Q K^T of its next tile).
      mbar_wait(q_full, 0);
      issue_qk(0);
      issue_qk(1);
#pragma unroll 1
      for (int j = 0; j < FA_NT; ++j) {
        if (j + 2 < FA_NT) issue_qk(j + 2);
        issue_pv(j);
      }
    }
  } else if (warp >= 4) {
    // Two softmax groups of 4 warps (one warp of each per SM sub-partition), thread = query row = TMEM lane.  Group g
    // owns the key tiles j = g, g + 2, ... together with their buffers S[g], P[g], O_blk[g], and keeps its OWN running
    // (max, sum, O) over them; the two partial softmax states are merged once at the end (flash-decoding style).
    // Round 1 / early round 2 ran all 8 warps on the SAME tile, two threads per row: every tile was a lock-step sequence
    // wait S -> max pass -> cross-half exchange + barrier -> exp pass -> fold O_blk, so the MUFU pipe (the unit that bounds
    // head_dim-64 attention: 128 x 128 ex2 per tile at 16 per clock = 1024 clocks against 512 clocks of MMA) idled
    // through every non-exp phase (XU 39 % busy, profiles/prof_flash_r2.md).  With the tiles alternating between two
    // independent groups one group's exp pass runs under the other's waits, max pass and O fold, there is no exchange
    // between the halves of a row, and each tile's S is read by one thread per row.
    const int q4 = warp & 3, grp = (warp - 4) >> 2, lane = lane_id();
    const int row = q4 * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const int qrow = qt * FA_BQ + row;
    const uint32_t o_addr = tmem_O[grp] + lane_off;
    uint8_t* prow = sP + grp * FA_P_BYTES + row * 128;    // + k-block * 16 KB + swizzled 16-byte chunk
    float m = -INFINITY, l = 0.f, m_ref = -INFINITY;      // m, m_ref in the scaled log2 domain
    float o[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) o[e] = 0.f;
    auto accumulate_o = [&](int k, float m_i) {
      // O += PV of this group's k-th tile, formed with probabilities relative to m_i
      mbar_wait(&o_full[grp], (uint32_t)k & 1u);
      tc_fence_after();
      const float resc = (m_ref == -INFINITY) ? 0.f : fa_exp2(m_ref - m_i);
      tmem_sweep16<4>(o_addr, [&](const uint32_t (&v)[16], int c) {
#pragma unroll
        for (int e = 0; e < 16; ++e) o[c * 16 + e] = fmaf(o[c * 16 + e], resc, __uint_as_float(v[e]));
      });
      m_ref = m_i;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[grp]);
    };
    float m_prev_tile = -INFINITY;
    int k = 0;                                            // tiles this group has taken
#pragma unroll 1
    for (int j = grp; j < FA_NT; j += 2, ++k) {
      const int sb = j % 3;
      mbar_wait(&s_full[sb], (uint32_t)(j / 3) & 1u);
      tc_fence_after();
      const uint32_t s_addr = tmem + sb * 128 + lane_off;
      const int nvalid = S_ENC - j * FA_BK;               // keys of this tile that exist: only the 12th tile is ragged (92)
      const bool ragged = nvalid < FA_BK;
      // pass 1: raw row max over the tile's 128 keys
      float mx_raw = -INFINITY;
      if (!ragged) {
        // four independent max chains (one serial chain of 64 dependent FMNMX3 showed up as fixed-latency `wait` stalls)
        float mq[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        tmem_sweep16<FA_BK / 16>(s_addr, [&](const uint32_t (&v)[16], int) {
#pragma unroll
          for (int e = 0; e < 16; ++e) mq[e & 3] = fmaxf(mq[e & 3], __uint_as_float(v[e]));
        });
        mx_raw = fmaxf(fmaxf(mq[0], mq[1]), fmaxf(mq[2], mq[3]));
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < FA_BK; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < nvalid) mx_raw = fmaxf(mx_raw, __uint_as_float(v[e]));
        }
      }
      const float mx = fmaxf(m, mx_raw * p.scale_log2);
      const float alpha = (m == -INFINITY) ? 0.f : fa_exp2(m - mx);
      // O_blk of this group's previous tile: its P V MMA was fed one whole (other-group) tile ago.  Folding it in HERE,
      // before the exp pass, is also what frees P[grp] -- the MMA that read it has completed.
      if (k >= 1) accumulate_o(k - 1, m_prev_tile);
      // pass 2: probabilities -> smem (two k-blocks of 64 keys, 128-byte rows, 16-byte chunks XOR-swizzled by row & 7), row sum
      float sum = 0.f;
      const float neg_mx = -mx;
      if (!ragged) {
        float sq[4] = {0.f, 0.f, 0.f, 0.f};
        tmem_sweep16<FA_BK / 16>(s_addr, [&](const uint32_t (&v)[16], int c) {   // keys 16 c .. 16 c + 15
          uint8_t* pk = prow + (c >> 2) * (FA_BQ * 128);
#pragma unroll
          for (int g = 0; g < 2; ++g) {                     // 2 pieces of 8 keys = 16 bytes
            __align__(16) __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float p0 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, neg_mx));
              const float p1 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, neg_mx));
              sq[e] += p0 + p1;                               // four independent partial row sums
              h2[e] = __floats2half2_rn(p0, p1);
            }
            const int chunk = (c & 3) * 2 + g;              // 16-byte piece inside the 128-byte row of this k-block
            *reinterpret_cast<uint4*>(pk + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
          }
        });
        sum = (sq[0] + sq[1]) + (sq[2] + sq[3]);
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < FA_BK; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
          uint8_t* pk = prow + (c0 >> 6) * (FA_BQ * 128);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            __align__(16) __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int kk = c0 + g * 8 + 2 * e;
              float p0 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, neg_mx));
              float p1 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, neg_mx));
              if (kk >= nvalid) p0 = 0.f;
              if (kk + 1 >= nvalid) p1 = 0.f;
              sum += p0 + p1;
              h2[e] = __floats2half2_rn(p0, p1);
            }
            const int chunk = ((c0 & 63) >> 3) + g;
            *reinterpret_cast<uint4*>(pk + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
          }
        }
      }
      l = l * alpha + sum;
      m = mx;
      tc_fence_before();
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);
        mbar_arrive(&p_full[grp]);
      }
      m_prev_tile = mx;
    }
    accumulate_o(k - 1, m_prev_tile);
    // Merge the two groups' states.  Group 1's last tile is the kernel's last MMA: once its O_blk has arrived every
    // tensor-core read of P is complete, so the P buffers carry group 1's (o[64], m, l), one column of floats per row.
    float* mg = reinterpret_cast<float*>(sP);               // [66][128]
    if (grp == 1) {
#pragma unroll
      for (int e = 0; e < 64; ++e) mg[e * FA_BQ + row] = o[e];
      mg[64 * FA_BQ + row] = m;
      mg[65 * FA_BQ + row] = l;
    }
    softmax_sync();
    if (grp == 0 && qrow < S_ENC) {
      const float m1 = mg[64 * FA_BQ + row], l1 = mg[65 * FA_BQ + row];
      const float mm = fmaxf(m, m1);
      const float a0 = fa_exp2(m - mm), a1 = fa_exp2(m1 - mm);
      const float inv = 1.f / fmaf(l, a0, l1 * a1);
      __half* dst = p.out + ((long)b * S_ENC + qrow) * p.d + h * 64;
#pragma unroll
      for (int e = 0; e < 64; e += 8) {
        __align__(16) __half2 h2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
          h2[t] = __floats2half2_rn(fmaf(o[e + 2 * t], a0, mg[(e + 2 * t) * FA_BQ + row] * a1) * inv,
                                    fmaf(o[e + 2 * t + 1], a0, mg[(e + 2 * t + 1) * FA_BQ + row] * a1) * inv);
        *reinterpret_cast<uint4*>(dst + e) = *reinterpret_cast<const uint4*>(h2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

// The DEFAULT kernel (the round-1 / early round-2 softmax organisation): all 8 softmax warps work on the
// SAME key tile, two threads per query row (64 keys each), per-tile row-max exchange through shared memory.
__global__ void __launch_bounds__(384, 1)
flash_attn_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FaParams p) {
  extern __shared__ uint8_t fa_raw[];
  uint8_t* base = fa_raw + ((1024u - (smem_u32(fa_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS/STS)
  uint8_t* sQ = base;
  uint8_t* sKV = sQ + FA_Q_BYTES;
  uint8_t* sP = sKV + FA_STAGES * FA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * FA_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;                 // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;     // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;      // [2]
  uint64_t* s_empty = s_full + 2;               // [2]
  uint64_t* p_full = s_empty + 2;               // [2]
  uint64_t* o_full = p_full + 2;                // [2]
  uint64_t* o_empty = o_full + 2;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  __shared__ float smax[2][2][FA_BQ];   // [tile parity][column half][row]: per-tile row-max exchange

  const int warp = threadIdx.x >> 5;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], 8); mbar_init(&o_full[i], 1); mbar_init(&o_empty[i], 8); }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S[2] = {tmem, tmem + 128};
  // P (shared memory) and O_blk (TMEM) are double buffered like S: the softmax warps fold O_blk of tile j-1 into their
  // registers only AFTER they have written P of tile j, so they never sit out the P V MMA of the tile they just fed
  // (single-buffered, every tile paid that round trip between its two passes: tensor pipe 19 % busy, round-1 profile).
  const uint32_t tmem_O[2] = {tmem + 256, tmem + 320};

  if (warp == 0) {
    if (elect_one()) {
      int c[4];
      mbar_expect_tx(q_full, FA_Q_BYTES);
      fa_coords(c, p.q_pos, 0, qt * FA_BQ, h, b);
      tma_load_4d(sQ, &tmQ, q_full, c[0], c[1], c[2], c[3]);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < FA_NT; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + stage * FA_STAGE_BYTES;
        mbar_expect_tx(&kv_full[stage], FA_STAGE_BYTES);
        fa_coords(c, p.k_pos, 0, j * FA_BK, h, b);
        tma_load_4d(sk, &tmK, &kv_full[stage], c[0], c[1], c[2], c[3]);
        fa_coords(c, p.v_pos, j * FA_BK, 0, h, b);
        tma_load_4d(sk + FA_K_BYTES, &tmV, &kv_full[stage], c[0], c[1], c[2], c[3]);
        fa_coords(c, p.v_pos, j * FA_BK + 64, 0, h, b);
        tma_load_4d(sk + FA_K_BYTES + 64 * 128, &tmV, &kv_full[stage], c[0], c[1], c[2], c[3]);
        if (++stage == FA_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(FA_BQ, FA_BK);
      constexpr uint32_t idesc_o = umma_idesc_f16(FA_BQ, 64);
      const uint64_t qdesc = umma_desc_sw128(smem_u32(sQ));
      auto issue_pv = [&](int i) {
        const int st = i % FA_STAGES, pb = i & 1;
        const uint32_t par = (uint32_t)(i >> 1) & 1u;
        const uint64_t pdesc = umma_desc_sw128(smem_u32(sP + pb * FA_P_BYTES));
        mbar_wait(&p_full[pb], par);              // P_i written by the softmax warps
        mbar_wait(&o_empty[pb], par ^ 1);         // O_blk of tile i-2 (same buffer) consumed
        tc_fence_after();
        const uint64_t vdesc = umma_desc_sw128(smem_u32(sKV + st * FA_STAGE_BYTES + FA_K_BYTES));
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a = pdesc + (uint64_t)((kk >> 2) * (FA_BQ * 128 >> 4) + (kk & 3) * 2);
          const uint64_t bd = vdesc + (uint64_t)((kk >> 2) * (64 * 128 >> 4) + (kk & 3) * 2);
          umma_f16(tmem_O[pb], a, bd, idesc_o, kk > 0 ? 1u : 0u);
        }
        umma_commit(&o_full[pb]);
        umma_commit(&kv_empty[st]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < FA_NT; ++j) {
        const int st = j % FA_STAGES;
        mbar_wait(&kv_full[st], (j / FA_STAGES) & 1);
        mbar_wait(&s_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(smem_u32(sKV + st * FA_STAGE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_S[j & 1], qdesc + (uint64_t)(2 * k), kdesc + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(FA_NT - 1);
    }
  } else if (warp >= 4) {
    // 8 softmax warps = 2 per SM sub-partition.  Warps 4-7 own key columns [0,64) of the tile and output
    // columns [0,32); warps 8-11 own keys [64,128) and outputs [32,64).  Thread pair (w, w+4) shares a query row.
    const int q4 = warp & 3, half = (warp - 4) >> 2, lane = lane_id();
    const int row = q4 * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const int qrow = qt * FA_BQ + row;
    float m = -INFINITY, l = 0.f, m_ref = -INFINITY;      // m, m_ref in the scaled log2 domain
    float o[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) o[e] = 0.f;
    auto accumulate_o = [&](int i, float m_i) {
      // O += PV_i, where PV_i was formed with probabilities relative to m_i
      mbar_wait(&o_full[i & 1], (uint32_t)(i >> 1) & 1u);
      tc_fence_after();
      const float resc = (m_ref == -INFINITY) ? 0.f : fa_exp2(m_ref - m_i);
      uint32_t v[32];
      tmem_ld_32x32(tmem_O[i & 1] + lane_off + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) o[e] = fmaf(o[e], resc, __uint_as_float(v[e]));
      m_ref = m_i;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[i & 1]);
    };
    float m_prev_tile = -INFINITY;
    for (int j = 0; j < FA_NT; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int key0 = j * FA_BK + half * 64;
      const int nvalid = S_ENC - key0;                    // keys of this half that exist (only the last tile is ragged)
      const uint32_t s_addr = tmem_S[j & 1] + lane_off + half * 64;
      // Only the 12th key tile is ragged (1500 = 11 x 128 + 92).  The mask test used to be evaluated for every score
      // of every tile (ISETP + FSEL = 20 % of the kernel's instructions, source-level ncu page of round 1); now the full
      // tiles run a path without it.
      const bool ragged = nvalid < 64;
      // pass 1: raw row max over this thread's 64 keys
      float mx_raw = -INFINITY;
      if (!ragged) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) mx_raw = fmaxf(mx_raw, __uint_as_float(v[e]));
        }
      } else {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < nvalid) mx_raw = fmaxf(mx_raw, __uint_as_float(v[e]));
        }
      }
      smax[j & 1][half][row] = mx_raw;
      softmax_sync();                                     // both halves of every row have published their max
      const float mx = fmaxf(m, fmaxf(mx_raw, smax[j & 1][half ^ 1][row]) * p.scale_log2);
      const float alpha = (m == -INFINITY) ? 0.f : fa_exp2(m - mx);
      // pass 2: probabilities -> smem (swizzled), row sum.  P[j & 1] is free: O_blk of tile j-2, i.e. the completion of
      // the MMA that read it, was folded in at the end of iteration j-1.
      float sum = 0.f;
      uint8_t* prow = sP + (j & 1) * FA_P_BYTES + half * (FA_BQ * 128) + row * 128;
      const float neg_mx = -mx;
      if (!ragged) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {                     // 4 chunks of 8 keys = 16 bytes
            __align__(16) __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float p0 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, neg_mx));
              const float p1 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, neg_mx));
              sum += p0 + p1;
              h2[e] = __floats2half2_rn(p0, p1);
            }
            const int chunk = (c0 >> 3) + g;                // 16-byte chunk inside the 128-byte row
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
          }
        }
      } else {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(s_addr + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            __align__(16) __half2 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int kk = c0 + g * 8 + 2 * e;
              float p0 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e]), p.scale_log2, neg_mx));
              float p1 = fa_exp2(fmaf(__uint_as_float(v[g * 8 + 2 * e + 1]), p.scale_log2, neg_mx));
              if (kk >= nvalid) p0 = 0.f;
              if (kk + 1 >= nvalid) p1 = 0.f;
              sum += p0 + p1;
              h2[e] = __floats2half2_rn(p0, p1);
            }
            const int chunk = (c0 >> 3) + g;
            *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
          }
        }
      }
      l = l * alpha + sum;
      m = mx;
      tc_fence_before();
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[j & 1]);
        mbar_arrive(&p_full[j & 1]);
      }
      if (j >= 1) accumulate_o(j - 1, m_prev_tile);   // its P V MMA was fed one whole tile ago
      m_prev_tile = mx;
    }
    accumulate_o(FA_NT - 1, m_prev_tile);
    // total row sum = sum of the two halves (same running max in both)
    smax[0][half][row] = l;
    softmax_sync();
    const float inv = 1.f / (l + smax[0][half ^ 1][row]);
    if (qrow < S_ENC) {
      __half* dst = p.out + ((long)b * S_ENC + qrow) * p.d + h * 64 + half * 32;
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        __align__(16) __half2 h2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) h2[t] = __floats2half2_rn(o[e + 2 * t] * inv, o[e + 2 * t + 1] * inv);
        *reinterpret_cast<uint4*>(dst + e) = *reinterpret_cast<const uint4*>(h2);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

void flash_attn_prime() {
  WL_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
  WL_CUDA(cudaFuncSetAttribute(flash_attn_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
}

// qk: [nb][1500][2d] fp16 (q | k), vt: [nb][d][S_PAD] fp16 (V transposed per head), out: [nb*1500][d] fp16
void encoder_attention_fused(cudaStream_t st, const __half* qk, const __half* vt, __half* out, int nb, int H, int d) {
  GemmOperand q, k, v;
  q.ptr = qk; q.rows = S_ENC; q.k = 64; q.ld = 2L * d; q.n1 = H; q.s1 = 64; q.n2 = nb; q.s2 = (long)S_ENC * 2 * d;
  k = q;
  k.ptr = qk + d;
  v.ptr = vt; v.rows = 64; v.k = S_PAD; v.ld = S_PAD; v.n1 = H; v.s1 = 64L * S_PAD; v.n2 = nb; v.s2 = (long)d * S_PAD;
  const TmapInfo iq = make_tmap(q, FA_BQ), ik = make_tmap(k, FA_BK), iv = make_tmap(v, 64);
  FaParams p;
  for (int i = 0; i < 3; ++i) { p.q_pos[i] = iq.pos[i]; p.k_pos[i] = ik.pos[i]; p.v_pos[i] = iv.pos[i]; }
  p.out = out; p.d = d;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  dim3 grid(FA_NT, H, nb);
  // WLB200_FA_SPLIT=1 selects flash_attn_kernel (two softmax groups on alternating key tiles: 413 vs 502 us per launch at
  // 16 streams under ncu).  It is NOT the default: its encoder output agrees with the oracle to the same tolerance
  // (max err 0.013 at large-v3, 0.006 against the pair kernel), but with it two decode-level parity tests fail where they
  // pass with the pair kernel (tiny beam 5: a pruning gap of 0.52 against the 0.455 allowed; the sampling test: a row
  // leaves the oracle's at a step whose key margin is 2.4) and the round's GPU budget ended before that was explained.
  // Parity is the first gate, so the proven kernel runs until it is.
  const char* e = getenv("WLB200_FA_SPLIT");
  if (e && atoi(e) != 0) flash_attn_kernel<<<grid, 384, FA_SMEM, st>>>(iq.tm, ik.tm, iv.tm, p);
  else flash_attn_pair_kernel<<<grid, 384, FA_SMEM, st>>>(iq.tm, ik.tm, iv.tm, p);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
