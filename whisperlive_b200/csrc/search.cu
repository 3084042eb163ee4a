// K12: logits processors + log-softmax + candidate selection (search_rows), and the per-stream
// search state machine: prompt feeding, CT2-style beam search, greedy / Gumbel-max sampling
// (search_streams).  Everything stays on the device; the host only polls a finished-stream counter.
// Semantics: oracle/search.py (restated CTranslate2 behaviour; reference call site
// transcriber_faster_whisper.py:1394-1407).
#include "kernels.cuh"

namespace wl {

struct MaskCtx {
  const unsigned* suppress;
  int first, suppress_blank, use_ts, last_is_ts, penult_is_ts, has_ts, ts_cutoff, max_initial;
  int eot, no_timestamps, ts_begin, blank;
};

__device__ __forceinline__ bool tok_masked(int t, const MaskCtx& c) {
  if (c.suppress[t >> 5] & (1u << (t & 31))) return true;
  if (c.first && c.suppress_blank && (t == c.blank || t == c.eot)) return true;
  if (c.use_ts) {
    if (t == c.no_timestamps) return true;
    if (c.first) {
      if (t < c.ts_begin || t > c.ts_begin + c.max_initial) return true;
    } else {
      if (c.last_is_ts) {
        if (c.penult_is_ts) {
          if (t >= c.ts_begin) return true;
        } else if (t < c.eot) {
          return true;
        }
      }
      if (c.has_ts && t >= c.ts_begin && t < c.ts_cutoff) return true;
    }
  }
  return false;
}

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}

__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  if (m2 == -INFINITY) return;
  if (m == -INFINITY) { m = m2; s = s2; return; }
  if (m2 > m) { s = s * __expf(m - m2) + s2; m = m2; }
  else s += s2 * __expf(m2 - m);
}

__device__ void block_lse(float& m, float& s, float* red /*[16]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[w] = m; red[8 + w] = s; }
  __syncthreads();
  m = red[0]; s = red[8];
  for (int i = 1; i < 8; ++i) lse_merge(m, s, red[i], red[8 + i]);
  __syncthreads();
}

constexpr int SR_THREADS = 256;

__global__ void __launch_bounds__(SR_THREADS) search_rows_kernel(DecodeState s, const float* __restrict__ logits, SearchOpts o,
                                                                 VocabIds v) {
  const int r = blockIdx.x, tid = threadIdx.x;
  const int b = r / o.rows_per_stream;
  pdl_trigger();
  pdl_wait();
  if (!s.active[r] || s.done[b]) return;
  __shared__ float red[16];
  __shared__ float lkey[MAX_CAND][SR_THREADS];   // per-thread sorted lists, thread index fastest (no bank conflicts)
  __shared__ int ltok[MAX_CAND][SR_THREADS];
  __shared__ float wkey[8];
  __shared__ int wtok[8], wtid[8];
  __shared__ int win_tid;

  const float* lg = logits + (long)r * v.vocab_ld;
  const int fed = s.fed[b], P = s.prompt_len[b];
  if (fed == s.sot_index[b] && r == b * o.rows_per_stream) {
    float m = -INFINITY, sm = 0.f;
    for (int t = tid; t < v.vocab; t += SR_THREADS) lse_merge(m, sm, lg[t], 1.f);
    block_lse(m, sm, red);
    if (tid == 0) s.nospeech_row[r] = __expf(lg[v.no_speech] - m) / sm;
  }
  if (s.force_len[b] > 0) {
    // teacher-forced pass: probability of the NEXT forced token under the raw distribution (K14 text_token_probs)
    if (fed + 1 < P && r == b * o.rows_per_stream) {
      float m = -INFINITY, sm = 0.f;
      for (int t = tid; t < v.vocab; t += SR_THREADS) lse_merge(m, sm, lg[t], 1.f);
      block_lse(m, sm, red);
      if (tid == 0) s.force_prob[(long)b * T_MAX + fed] = __expf(lg[s.prompt[(long)b * T_MAX + fed + 1]] - m) / sm;
    }
    return;
  }
  if (fed < P - 1) return;

  const int glen = s.gen_len[r];
  const int* hist = s.hist + (long)r * T_MAX;
  MaskCtx c;
  c.suppress = o.suppress_mask;
  c.first = glen == 0;
  c.suppress_blank = o.suppress_blank;
  c.use_ts = s.use_ts[b];
  c.max_initial = o.max_initial_ts;
  c.eot = v.eot; c.no_timestamps = v.no_timestamps; c.ts_begin = v.ts_begin; c.blank = v.blank;
  const int last = glen > 0 ? hist[glen - 1] : -1;
  c.last_is_ts = glen > 0 && last >= v.ts_begin;
  c.penult_is_ts = glen < 2 || hist[glen - 2] >= v.ts_begin;
  const int lts = s.last_ts[r];
  c.has_ts = lts >= 0;
  c.ts_cutoff = (c.last_is_ts && !c.penult_is_ts) ? lts : lts + 1;

  // pass 1: softmax statistics of the text part and the timestamp part after masks a-d
  float mt = -INFINITY, st = 0.f, mz = -INFINITY, sz = 0.f;
  const float4* lg4 = reinterpret_cast<const float4*>(lg);
  const int n4 = v.vocab_ld >> 2;
  auto stat = [&](int t, float x) {
    if (t >= v.vocab || tok_masked(t, c)) return;
    if (t < v.ts_begin) lse_merge(mt, st, x, 1.f);
    else lse_merge(mz, sz, x, 1.f);
  };
  for (int i0 = tid; i0 < n4; i0 += 4 * SR_THREADS) {
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = (i0 + u * SR_THREADS < n4) ? lg4[i0 + u * SR_THREADS] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * (i0 + u * SR_THREADS);
      if (t < v.vocab) { stat(t, q[u].x); stat(t + 1, q[u].y); stat(t + 2, q[u].z); stat(t + 3, q[u].w); }
    }
  }
  block_lse(mt, st, red);
  block_lse(mz, sz, red);
  const float lse_text = st > 0.f ? mt + logf(st) : -INFINITY;
  const float lse_ts = sz > 0.f ? mz + logf(sz) : -INFINITY;
  const bool text_off = c.use_ts && !c.first && lse_ts > mt;  // rule e: mass on timestamps beats the best text token
  float lse;
  if (text_off || lse_text == -INFINITY) lse = lse_ts;
  else if (lse_ts == -INFINITY) lse = lse_text;
  else {
    const float hi = fmaxf(lse_text, lse_ts), lo = fminf(lse_text, lse_ts);
    lse = hi + log1pf(expf(lo - hi));
  }

  // pass 2: thread-local sorted top-NC, then NC rounds of block arg-max over the list heads
  const int NC = o.beam > 1 ? 2 * o.beam : 1;
  int cnt = 0;
  uint32_t gkey = 0;
  if (o.sampling) gkey = hash_u32((o.seed * 0x9E3779B1u) ^ hash_u32((uint32_t)((b * 64 + (r - b * o.rows_per_stream)) * 65537 + s.step[b])));
  auto consider = [&](int t, float x) {
    if (t >= v.vocab || tok_masked(t, c)) return;
    if (text_off && t < v.ts_begin) return;
    const float lp = x - lse;
    float key = lp;
    if (o.sampling) {
      const double u = ((double)hash_u32((uint32_t)t ^ gkey) + 0.5) / 4294967296.0;
      key = __fdiv_rn(lp, o.temperature) + (float)(-log(-log(u)));
    }
    if (cnt == NC && !(key > lkey[NC - 1][tid])) return;
    int i = cnt < NC ? cnt : NC - 1;
    while (i > 0 && key > lkey[i - 1][tid]) {
      lkey[i][tid] = lkey[i - 1][tid]; ltok[i][tid] = ltok[i - 1][tid];
      --i;
    }
    lkey[i][tid] = key; ltok[i][tid] = t;
    if (cnt < NC) ++cnt;
  };
  for (int i0 = tid; i0 < n4; i0 += 4 * SR_THREADS) {
    float4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = (i0 + u * SR_THREADS < n4) ? lg4[i0 + u * SR_THREADS] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * (i0 + u * SR_THREADS);
      if (t < v.vocab) { consider(t, q[u].x); consider(t + 1, q[u].y); consider(t + 2, q[u].z); consider(t + 3, q[u].w); }
    }
  }
  int hd = 0;
  for (int round = 0; round < NC; ++round) {
    float k = hd < cnt ? lkey[hd][tid] : -INFINITY;
    int t = hd < cnt ? ltok[hd][tid] : 0x7fffffff;
    int who = tid;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float k2 = __shfl_xor_sync(0xffffffffu, k, off);
      const int t2 = __shfl_xor_sync(0xffffffffu, t, off), w2 = __shfl_xor_sync(0xffffffffu, who, off);
      if (k2 > k || (k2 == k && t2 < t)) { k = k2; t = t2; who = w2; }
    }
    if ((tid & 31) == 0) { wkey[tid >> 5] = k; wtok[tid >> 5] = t; wtid[tid >> 5] = who; }
    __syncthreads();
    if (tid == 0) {
      float bk = wkey[0]; int bt = wtok[0], bw = wtid[0];
      for (int w = 1; w < 8; ++w)
        if (wkey[w] > bk || (wkey[w] == bk && wtok[w] < bt)) { bk = wkey[w]; bt = wtok[w]; bw = wtid[w]; }
      win_tid = bt == 0x7fffffff ? -1 : bw;
      if (win_tid < 0) { s.cand_val[(long)r * MAX_CAND + round] = -INFINITY; s.cand_tok[(long)r * MAX_CAND + round] = -1; }
    }
    __syncthreads();
    if (tid == win_tid) {
      s.cand_val[(long)r * MAX_CAND + round] = lg[ltok[hd][tid]] - lse;
      s.cand_tok[(long)r * MAX_CAND + round] = ltok[hd][tid];
      ++hd;
    }
    __syncthreads();
  }
}

void search_rows(cudaStream_t st, const DecodeState& s, const float* logits, const SearchOpts& o, const VocabIds& v, int R) {
  launch_kernel(search_rows_kernel, dim3(R), dim3(SR_THREADS), 0, st, s, logits, o, v);
  note_launch(1);
}

// ============================================================================ per-stream state machine
__device__ void finish_stream(DecodeState& s, int b, int Kr) {
  s.done[b] = 1;
  for (int j = 0; j < Kr; ++j) s.active[b * Kr + j] = 0;
  atomicAdd(s.n_done, 1);
}

__global__ void __launch_bounds__(128) search_streams_kernel(DecodeState s, SearchOpts o, VocabIds v) {
  const int b = blockIdx.x, tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (s.done[b]) return;
  const int Kr = o.rows_per_stream, row0 = b * Kr;
  const int P = s.prompt_len[b], fed = s.fed[b];
  __shared__ int sh_hist[MAX_ROWS_PER_STREAM][T_MAX];
  __shared__ short sh_src[MAX_ROWS_PER_STREAM][T_MAX];
  __shared__ int new_parent[MAX_ROWS_PER_STREAM], new_tok[MAX_ROWS_PER_STREAM];
  __shared__ float new_cum[MAX_ROWS_PER_STREAM];
  __shared__ int hyp_parent[MAX_ROWS_PER_STREAM], hyp_extra[MAX_ROWS_PER_STREAM], hyp_slot[MAX_ROWS_PER_STREAM];
  __shared__ int n_newalive, n_newhyp, finished;

  if (tid == 0) {
    s.steps_run[b] += 1;
    if (fed == s.sot_index[b]) s.no_speech[b] = s.nospeech_row[row0];
  }
  // ---- teacher-forced feeding only (detect_language / align / logits hook)
  if (s.force_len[b] > 0) {
    if (tid == 0) {
      if (fed + 1 < P) {
        s.tok_in[row0] = s.prompt[(long)b * T_MAX + fed + 1];
        s.pos[row0] = fed + 1;
        s.fed[b] = fed + 1;
      } else {
        finish_stream(s, b, Kr);
      }
    }
    return;
  }
  // ---- prompt feeding
  if (fed < P - 1) {
    const int nf = fed + 1;
    if (tid == 0) {
      s.tok_in[row0] = s.prompt[(long)b * T_MAX + nf];
      s.pos[row0] = nf;
      s.fed[b] = nf;
    }
    if (nf == P - 1 && o.beam == 1 && Kr > 1) {
      // independent sampling rows all start from the prompt cache of row 0
      for (int j = 1; j < Kr; ++j) {
        for (int p = tid; p < nf; p += 128) s.src[(long)(row0 + j) * T_MAX + p] = s.src[(long)row0 * T_MAX + p];
        if (tid == 0) {
          s.tok_in[row0 + j] = s.prompt[(long)b * T_MAX + nf];
          s.pos[row0 + j] = nf;
          s.active[row0 + j] = 1;
        }
      }
    }
    return;
  }
  const int step = s.step[b];
  const bool last_step = step + 1 >= s.n_new[b];

  // ---- greedy / sampling: every row is an independent hypothesis
  if (o.beam == 1) {
    if (tid < Kr && !s.row_done[row0 + tid]) {
      const int r = row0 + tid;
      const int tok = s.cand_tok[(long)r * MAX_CAND];
      const float val = s.cand_val[(long)r * MAX_CAND];
      int len = s.gen_len[r];
      bool fin = false;
      if (tok < 0) fin = true;
      else {
        s.cum[r] += val;
        if (tok == v.eot) fin = true;
        else {
          s.hist[(long)r * T_MAX + len] = tok;
          s.gen_len[r] = ++len;
          if (tok >= v.ts_begin) s.last_ts[r] = tok;
          if (last_step) fin = true;
        }
      }
      if (fin) { s.row_done[r] = 1; s.active[r] = 0; }
      else { s.tok_in[r] = tok; s.pos[r] += 1; }
    }
    __syncthreads();
    if (tid == 0) {
      s.step[b] = step + 1;
      bool all = true;
      for (int j = 0; j < Kr; ++j) all = all && s.row_done[row0 + j];
      finished = all ? 1 : 0;
      if (all) {
        s.hyp_count[b] = Kr;
        for (int j = 0; j < Kr; ++j) {
          s.hyp_cum[b * MAX_HYPS + j] = s.cum[row0 + j];
          s.hyp_len[b * MAX_HYPS + j] = s.gen_len[row0 + j];
        }
        finish_stream(s, b, Kr);
      }
    }
    __syncthreads();
    if (finished) {
      for (int j = 0; j < Kr; ++j)
        for (int p = tid; p < s.gen_len[row0 + j]; p += 128)
          s.hyp_tok[((long)b * MAX_HYPS + j) * T_MAX + p] = s.hist[(long)(row0 + j) * T_MAX + p];
    }
    return;
  }

  // ---- beam search (CT2 walk, see oracle/search.py)
  const int K = o.beam;
  const int n_alive = s.n_alive[b];
  const int pos_old = s.pos[row0];
  const int len_old = s.gen_len[row0];
  for (int j = 0; j < n_alive; ++j) {
    for (int p = tid; p < len_old; p += 128) sh_hist[j][p] = s.hist[(long)(row0 + j) * T_MAX + p];
    for (int p = tid; p <= pos_old; p += 128) sh_src[j][p] = s.src[(long)(row0 + j) * T_MAX + p];
  }
  if (tid == 0) {
    // merged top-2K over the alive rows (each row's list is sorted): (total desc, row asc, list order)
    int idx[MAX_ROWS_PER_STREAM];
    for (int j = 0; j < MAX_ROWS_PER_STREAM; ++j) idx[j] = 0;
    int cb[MAX_CAND], ct[MAX_CAND];
    float cs[MAX_CAND];
    int nc = 0;
    const int NC = 2 * K;
    for (int k = 0; k < NC; ++k) {
      int bj = -1;
      float bs = -INFINITY;
      for (int j = 0; j < n_alive; ++j) {
        if (idx[j] >= NC) continue;
        const float cv = s.cand_val[(long)(row0 + j) * MAX_CAND + idx[j]];
        if (cv == -INFINITY) continue;
        const float tot = s.cum[row0 + j] + cv;
        if (tot > bs) { bs = tot; bj = j; }
      }
      if (bj < 0) break;
      cb[nc] = bj; ct[nc] = s.cand_tok[(long)(row0 + bj) * MAX_CAND + idx[bj]]; cs[nc] = bs;
      ++idx[bj]; ++nc;
    }
    int na = 0, nh = 0, hc = s.hyp_count[b];
    int secondary = K;
    for (int k = 0; k < K && k < nc; ++k) {
      int beam = cb[k], tok = ct[k];
      float sc = cs[k];
      if (tok == v.eot || last_step) {
        if (hc + nh < MAX_HYPS) {
          hyp_parent[nh] = beam;
          hyp_extra[nh] = tok == v.eot ? -1 : tok;
          hyp_slot[nh] = hc + nh;
          s.hyp_cum[b * MAX_HYPS + hc + nh] = sc;
          s.hyp_len[b * MAX_HYPS + hc + nh] = len_old + (tok == v.eot ? 0 : 1);
          ++nh;
        }
        if (last_step) continue;
        bool found = false;
        while (secondary < nc) {
          const int b2 = cb[secondary], t2 = ct[secondary];
          const float s2 = cs[secondary];
          ++secondary;
          if (t2 != v.eot) { beam = b2; tok = t2; sc = s2; found = true; break; }
        }
        if (!found) continue;
      }
      new_parent[na] = beam; new_tok[na] = tok; new_cum[na] = sc;
      ++na;
    }
    s.hyp_count[b] = hc + nh;
    s.step[b] = step + 1;
    n_newalive = na; n_newhyp = nh;
    finished = (hc + nh >= o.max_cand || last_step || na == 0) ? 1 : 0;
  }
  __syncthreads();
  for (int i = 0; i < n_newhyp; ++i) {
    int* dst = s.hyp_tok + ((long)b * MAX_HYPS + hyp_slot[i]) * T_MAX;
    for (int p = tid; p < len_old; p += 128) dst[p] = sh_hist[hyp_parent[i]][p];
    if (tid == 0 && hyp_extra[i] >= 0) dst[len_old] = hyp_extra[i];
  }
  if (finished) {
    if (tid == 0) finish_stream(s, b, Kr);
    return;
  }
  for (int j = 0; j < n_newalive; ++j) {
    const int r = row0 + j, pj = new_parent[j];
    for (int p = tid; p < len_old; p += 128) s.hist[(long)r * T_MAX + p] = sh_hist[pj][p];
    for (int p = tid; p <= pos_old; p += 128) s.src[(long)r * T_MAX + p] = sh_src[pj][p];
  }
  __syncthreads();
  if (tid < Kr) {
    const int r = row0 + tid;
    if (tid < n_newalive) {
      const int tok = new_tok[tid];
      const int plts = s.last_ts[row0 + new_parent[tid]];
      s.hist[(long)r * T_MAX + len_old] = tok;
      s.gen_len[r] = len_old + 1;
      s.cum[r] = new_cum[tid];
      s.tok_in[r] = tok;
      s.pos[r] = pos_old + 1;
      s.active[r] = 1;
      // last_ts of the parent must be read before any row overwrites it: stage through registers + barrier
      new_tok[tid] = tok >= v.ts_begin ? tok : plts;
    } else {
      s.active[r] = 0;
    }
  }
  __syncthreads();
  if (tid < n_newalive) s.last_ts[row0 + tid] = new_tok[tid];
  if (tid == 0) s.n_alive[b] = n_newalive;
}

void search_streams(cudaStream_t st, const DecodeState& s, const SearchOpts& o, const VocabIds& v, int B) {
  launch_kernel(search_streams_kernel, dim3(B), dim3(128), 0, st, s, o, v);
  note_launch(1);
}

// ============================================================================ init
__global__ void decode_init_kernel(DecodeState s, SearchOpts o, VocabIds v) {
  const int b = blockIdx.x, tid = threadIdx.x;
  const int Kr = o.rows_per_stream, row0 = b * Kr;
  const int P = s.prompt_len[b];
  if (tid < Kr) {
    const int r = row0 + tid;
    const bool on = tid == 0 || (P == 1 && o.beam == 1 && s.force_len[b] == 0);
    s.tok_in[r] = s.prompt[(long)b * T_MAX];
    s.pos[r] = 0;
    s.active[r] = on ? 1 : 0;
    s.cum[r] = 0.f;
    s.gen_len[r] = 0;
    s.last_ts[r] = -1;
    s.row_done[r] = 0;
    s.nospeech_row[r] = 0.f;
  }
  if (tid == 0) {
    s.fed[b] = 0;
    s.step[b] = 0;
    s.done[b] = 0;
    s.n_alive[b] = 1;
    s.no_speech[b] = 0.f;
    s.hyp_count[b] = 0;
    s.steps_run[b] = 0;
    if (b == 0) *s.n_done = 0;
  }
}

void decode_init(cudaStream_t st, const DecodeState& s, const SearchOpts& o, const VocabIds& v, int B, int R) {
  decode_init_kernel<<<B, 32, 0, st>>>(s, o, v);
  WL_CUDA(cudaGetLastError());
  note_launch(1);
}

}  // namespace wl
