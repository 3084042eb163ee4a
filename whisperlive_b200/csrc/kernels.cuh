// Launchers for the non-GEMM kernels of libwlb200 (K1, LN/softmax/prep, K10, K11, K12, K14).
#pragma once
#include "common.cuh"

namespace wl {

constexpr int T_MAX = 448;      // decoder positions
constexpr int S_ENC = 1500;     // encoder positions
constexpr int S_PAD = 1536;     // padded key dimension for materialised attention scores
constexpr int MAX_ROWS_PER_STREAM = 8;
constexpr int MAX_HYPS = 16;
constexpr int MAX_CAND = 16;    // 2 * beam, beam <= 8

// ---------------------------------------------------------------------------- K1 mel
struct MelTables {
  const float* window;      // [400]
  const float* twiddle;     // [400][2]
  const float* filt;        // [n_mels][201]
  const int* filt_range;    // [n_mels][2]
  int n_mels;
};
void mel_forward(cudaStream_t st, const float* pcm, const long* pcm_off, float* out, const long* out_off, unsigned* gmax,
                 const MelTables& t, int B, int max_frames);

// A value produced by a split-K GEMM: v(r, c) = bias[c] + sum_s ptr[s * stride + r * ld + c]  (fixed order).
// nsplit == 1 with bias == nullptr is a plain buffer; nsplit == 0 means "nothing pending".
struct PartialSrc {
  const float* ptr = nullptr;
  int nsplit = 0;
  long stride = 0;
  const float* bias = nullptr;
};

// ---------------------------------------------------------------------------- elementwise / normalisation
// features f32 [B][n_mels][3000] -> fp16 [B][3002][n_mels] (rows 0 and 3001 are zero: conv padding)
void prep_features(cudaStream_t st, const float* feats, __half* out, int B, int n_mels);
// weight upload: fp32 [a][b][k] -> fp16 [a][k][b] (conv kernels; b = k = 1 is a plain cast)
void cast_weight_f16(cudaStream_t st, const float* in, __half* out, long a, long b, long k);
// window gather from the resident log-mel of wl_mel_device: feat[w][m][t] = t < len[w] ? mel[off[stream[w]] + m * frames[stream[w]] + seek[w] + t] : 0
// (the reference slices features[:, seek : seek + segment_size] and zero-pads to 3000 frames on the host: transcriber_faster_whisper.py:1115-1127)
void gather_windows(cudaStream_t st, const float* mel, const long* mel_off, const int* frames, const int* win_stream, const int* win_seek,
                    const int* win_len, float* feat, int n_windows, int n_mels);
// y = LayerNorm(x) * gamma + beta ; x f32 [rows][d] -> y fp16 [rows][d] (and optionally f32 copy)
void layernorm_rows(cudaStream_t st, const float* x, const float* gamma, const float* beta, __half* y, float* y32,
                    long rows, int d);
// L2 prefetch request riding on a decode-step kernel: pull the encoder K/V (one layer) of the first n_streams LIVE
// streams into L2 while the latency-bound kernels that precede the cross-attention leave HBM idle (engine.cu).
struct L2Prefetch {
  const __half* k = nullptr;   // layer base of the cross K pool [slot][H][1500][64]
  const __half* v = nullptr;
  const int* slot = nullptr;   // [B] pool slot of each stream
  const int* done = nullptr;   // [B]
  int B = 0, n_streams = 0;
  long slot_bytes = 0;         // bytes of one stream's K (or V) for one layer
};
// decode step: x[r] += upd(r, :) (residual update pending from a split-K GEMM), then y = LayerNorm(x) as fp16
void layernorm_update_rows(cudaStream_t st, float* x, const PartialSrc& upd, const float* gamma, const float* beta, __half* y,
                           int rows, int d, const L2Prefetch* pf = nullptr);
// out[r][c] = fp16(gelu(in(r, c))), rows x cols (cols % 4 == 0)
void gelu_cast(cudaStream_t st, const PartialSrc& in, __half* out, int rows, int cols);
// scores f32 [rows][ld_in] (first n valid) -> softmax(scale * s) as fp16 [rows][ld_out], columns >= n zeroed
void softmax_rows(cudaStream_t st, const float* s, __half* p, long rows, int n, int ld_in, int ld_out, float scale);

// ---------------------------------------------------------------------------- decoder state (device resident)
struct DecodeState {
  // per row
  int* tok_in;       // [R]
  int* pos;          // [R] position of tok_in == tokens already cached for the row
  int* active;       // [R]
  float* cum;        // [R]
  int* gen_len;      // [R]
  int* last_ts;      // [R] last generated timestamp token or -1
  int* row_done;     // [R]
  int* hist;         // [R][T_MAX] generated tokens
  short* src;        // [R][T_MAX] physical cache row holding position p of this row's sequence
  int* wrow;         // optional [R]: physical cache row the self-attention kernel WRITES row r's new k/v to (null: r itself;
                     //   the batched prefill runs every prompt position as its own row, all writing to the stream's first row)
  // per-row candidates produced by search_rows
  float* cand_val;   // [R][MAX_CAND]
  int* cand_tok;     // [R][MAX_CAND]
  float* nospeech_row;  // [R] softmax(raw logits)[no_speech] of the row (valid when computed)
  // per stream
  int* slot;         // [B]
  int* prompt;       // [B][T_MAX]
  int* prompt_len;   // [B]
  int* fed;          // [B] index of the prompt token fed at the current step
  int* sot_index;    // [B] or -1
  int* use_ts;       // [B]
  int* pre_n;        // [B] prompt tokens after the sot sequence (a ``prefix``): CT2 treats them as already-sampled text
  int* pre_last;     // [B] last / second-to-last of those tokens (-1 when absent) and the last timestamp among them:
  int* pre_penult;   // [B]   seeds of the timestamp rules' history
  int* pre_lts;      // [B]
  int* n_new;        // [B] max new tokens
  int* step;         // [B] generation steps done
  int* done;         // [B]
  int* n_alive;      // [B]
  float* no_speech;  // [B]
  int* hyp_count;    // [B]
  float* hyp_cum;    // [B][MAX_HYPS]
  int* hyp_len;      // [B][MAX_HYPS]
  int* hyp_tok;      // [B][MAX_HYPS][T_MAX]
  int* steps_run;    // [B] decoder steps executed for the stream (diagnostics)
  int* n_done;       // [1] number of finished streams
  int* steps_left;   // [1] decode steps the device-side loop may still run (conditional WHILE graph)
  unsigned* seed;    // [1] sampling seed of this generate call (device scalar: the captured graph does not depend on it)
  int* brk;          // [2] decode sessions (step-level admission): [0] != 0 -> the device-side loop also ends as soon as a
                     //   stream finishes (so the host can hand its result out and refill the index); [1] = n_done at launch
  // teacher-forced mode (detect_language / align / logits test hook)
  int* force_len;    // [B] 0 = normal search; >0 = feed prompt only, then stop
  float* force_prob; // [B][T_MAX] P(prompt[i+1] | prompt[..i]) in teacher-forced mode
};

struct SearchOpts {
  int beam;            // K (1 = greedy / sampling)
  int rows_per_stream; // Kr
  int max_cand;        // round(K * patience)
  int suppress_blank;
  int max_initial_ts;
  int sampling;        // 1: Gumbel-max sampling over the full distribution
  float temperature;
  unsigned seed;
  const unsigned* suppress_mask;  // device bitmask over the vocabulary
};

struct VocabIds {
  int vocab, vocab_ld, eot, sot, no_speech, no_timestamps, ts_begin, blank;
};

// x[r] = E[tok_in[r]] + P[pos[r]]  (f32) for active rows; also src[r][pos[r]] = r
void decoder_embed(cudaStream_t st, const DecodeState& s, const __half* emb, const __half* pos_emb, float* x, int R, int d);

// K10: self attention over the KV cache with beam indirection. qkv f32 [R][3d]; out fp16 [R][d].
void decoder_self_attn(cudaStream_t st, const DecodeState& s, const PartialSrc& qkv, __half* kcache, __half* vcache,
                       long cache_row_stride, __half* out, int R, int H, int d);

// K11: cross attention of the rows of each stream over its persistent encoder K/V.
//   q f32 [R][d]; K/V caches [slot][H][1500][64] fp16 (layer base pointers); out fp16 [R][d]
struct CrossAttnWorkspace {
  float* part;     // [B][H][nsplit][MAX_ROWS_PER_STREAM][66]  (m, l, o[64])
  float* probs;    // optional [R][H][1500] f32 attention probabilities (align mode) or nullptr
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // profiling: recorded right before / after the main kernel when set
};
void decoder_cross_attn(cudaStream_t st, const DecodeState& s, const PartialSrc& q, const __half* kc, const __half* vc,
                        long slot_stride, const CrossAttnWorkspace& ws, __half* out, int B, int rows_per_stream, int H,
                        int d, int nsplit);
int cross_attn_pick_nsplit(int B, int H, int num_sms, int rows_per_stream);

// K12: per-row masked log-softmax + top candidates, then per-stream beam / greedy update.
void search_rows(cudaStream_t st, const DecodeState& s, const float* logits, const SearchOpts& o, const VocabIds& v, int R);
void search_streams(cudaStream_t st, const DecodeState& s, const SearchOpts& o, const VocabIds& v, int B);

// Last node of the loop body of the conditional WHILE graph: keep iterating while some stream is still decoding and the
// step budget is not used up (one thread; it runs after search_streams, so n_done is final for this step).
void loop_condition(cudaStream_t st, const DecodeState& s, cudaGraphConditionalHandle h, int B);

// initialise the state for a generate call (prompts already uploaded).  prefilled = 1: positions 0 .. P-2 of every prompt
// are already in the self-attention cache (K8 batched prefill): start at the last prompt token.
// index != null (device, B entries): initialise only those state indices of a running decode session -- their `done`
// flag was 1 and is counted in n_done, which drops by one per admitted stream instead of being reset.
void decode_init(cudaStream_t st, const DecodeState& s, const SearchOpts& o, const VocabIds& v, int B, int R, int prefilled = 0,
                 const int* index = nullptr);

// ---------------------------------------------------------------------------- K8 batched prefill helpers (prefill.cu)
void prefill_embed(cudaStream_t st, const int* tok, const int* pos, const int* active, const int* wrow, const __half* emb,
                   const __half* pos_emb, float* x, short* src, int M, int d);
void prefill_kv_write(cudaStream_t st, const float* qkv, const int* pos, const int* active, const int* wrow, __half* kc, __half* vc,
                      long row_stride, int M, int H, int d);
void gather_rows(cudaStream_t st, const float* x, const int* rows, float* dst, int n, int d);
void row_prob(cudaStream_t st, const float* logits, int vocab, int vocab_ld, const int* target, const int* out_index, float* out, int n);

}  // namespace wl
