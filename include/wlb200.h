/* libwlb200 -- C ABI of the B200-native Whisper hot path behind WhisperLive's transcriber.
 *
 * Every entry point replaces one call the reference makes into its native engine
 * (ctranslate2.models.Whisper + faster_whisper.FeatureExtractor; neither is in the reference tree,
 * the citations are the reference-side CALL SITES, /root/reference/whisper_live/...):
 *
 *   wl_init / wl_load_tensor / wl_finalize_weights
 *        <- ctranslate2.models.Whisper(model_path, device, device_index, compute_type, ...)
 *           transcriber/transcriber_faster_whisper.py:634-643 (model load; backend/faster_whisper_backend.py:173-178)
 *   wl_mel               <- FeatureExtractor.__call__      transcriber_faster_whisper.py:862, :1759; batch_inference.py:258
 *   wl_encode            <- Whisper.encode                 transcriber_faster_whisper.py:1339-1348; batch_inference.py:271
 *   wl_generate          <- Whisper.generate               transcriber_faster_whisper.py:1394-1407; batch_inference.py:355
 *   wl_detect_language   <- Whisper.detect_language        transcriber_faster_whisper.py:1140, :1771; batch_inference.py:283
 *   wl_align             <- Whisper.align                  transcriber_faster_whisper.py:1657-1663
 *   wl_slots_release     <- StorageView lifetime           transcriber_faster_whisper.py:1055, :1820-1823
 *
 * Conventions: plain pointers and sizes only; the caller owns every host buffer; the library owns
 * device memory, streams, CUDA graphs.  Every function returns 0 or a negative WL_ERR_* code and
 * never throws / aborts; wl_last_error() returns the message of the last failure on that context.
 * A context is driven by one thread at a time (the scheduler thread); ctypes releases the GIL.
 */
#ifndef WLB200_H
#define WLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WL_ABI_VERSION 3

typedef struct wl_ctx wl_ctx;

typedef struct wl_config {
  int32_t abi_version;   /* WL_ABI_VERSION */
  int32_t device;        /* CUDA ordinal */
  /* architecture */
  int32_t d_model, n_heads, enc_layers, dec_layers, n_mels, vocab;
  /* vocabulary ids the engine needs (ctranslate2 reads them from the model vocabulary) */
  int32_t eot, sot, no_speech, no_timestamps, timestamp_begin, blank;
  int32_t lang_begin, n_lang; /* language token ids are [lang_begin, lang_begin + n_lang) */
  /* capacity */
  int32_t max_streams;   /* streams per encode/generate call */
  int32_t max_beam;      /* decoder rows per stream (beam_size or num_hypotheses), <= 8 */
  int32_t enc_slots;     /* encoder-output / cross-KV slots in the pool (>= max_streams) */
  /* word alignment heads: pairs (layer, head) */
  int32_t n_align_heads;
  const int32_t* align_heads;
} wl_config;

typedef struct wl_gen_opts {
  int32_t beam_size;                 /* 1 = greedy / sampling */
  float patience;
  int32_t num_hypotheses;
  float length_penalty;
  int32_t max_length;                /* CT2 max_length (448) */
  int32_t suppress_blank;
  int32_t max_initial_timestamp_index;
  int32_t sampling_topk;             /* 1 = arg-max, 0 = sample from the full distribution */
  float sampling_temperature;
  uint32_t seed;
  const int32_t* suppress_tokens;
  int32_t n_suppress;
  int32_t use_cuda_graph;            /* 1: capture the decoder step once per call shape */
  const int32_t* max_length_per_stream; /* optional [B]: overrides max_length per stream (ragged max_new_tokens) */
  int32_t prefill;                   /* K8: 0 = library default (batched prefill of the prompt), 1 = on, 2 = off (one decode
                                        step per prompt token; the parity tests compare the two) */
} wl_gen_opts;

int wl_init(const wl_config* cfg, wl_ctx** out);
void wl_destroy(wl_ctx* ctx);
const char* wl_last_error(wl_ctx* ctx);   /* ctx may be NULL: last wl_init failure */

/* Weights: float32 host tensors under HF WhisperForConditionalGeneration names
 * ("model.encoder.conv1.weight", ...), plus "mel_filters" [n_mels, 201]. */
int wl_load_tensor(wl_ctx* ctx, const char* name, const float* data, const int64_t* shape, int32_t ndim);
int wl_finalize_weights(wl_ctx* ctx);

/* K1. pcm: B waveforms concatenated, offsets[B+1] in samples.  out: per stream [n_mels, n_b/160 + 1]
 * float32 row-major, concatenated at out_offsets[B+1] (in floats).  Host pointers. */
int wl_mel(wl_ctx* ctx, const float* pcm, const int64_t* offsets, int32_t B, float* out, const int64_t* out_offsets);

/* K2-K7. features: host [B, n_mels, 3000] float32.  slots_out[B] receives pool slots that hold the
 * encoder output and the cross-attention K/V of each stream until released. */
int wl_encode(wl_ctx* ctx, const float* features, int32_t B, int32_t* slots_out);
int wl_slots_release(wl_ctx* ctx, const int32_t* slots, int32_t n);
int wl_slots_free_count(wl_ctx* ctx);
/* encoder output of one slot as float32 [1500, d_model] (host) */
int wl_encoder_output(wl_ctx* ctx, int32_t slot, float* out);

/* K8-K13. prompts: B token lists concatenated, prompt_off[B+1].  Outputs (host):
 *   out_ids   [B, num_hypotheses, 448] int32      out_len  [B, num_hypotheses]
 *   out_score [B, num_hypotheses] (cum_logprob / len^length_penalty)
 *   out_no_speech [B]                              out_steps [B] decoder steps executed */
int wl_generate(wl_ctx* ctx, const int32_t* slots, int32_t B, const int32_t* prompts, const int32_t* prompt_off,
                const wl_gen_opts* opts, int32_t* out_ids, int32_t* out_len, float* out_score, float* out_no_speech,
                int32_t* out_steps);

/* N2. Decode session: step-level continuous batching -- replaces the run-to-completion batches of the reference's
 * BatchInferenceWorker._process_multi (whisper_live/batch_inference.py:155-187; its gaps :259, :334-339).
 *   wl_session_open    fixes the search options (wl_generate's, sampling excluded) and the number of stream indices
 *                      (<= max_streams); every index starts idle.  Re-opening is allowed once nothing is decoding.
 *   wl_session_admit   puts n streams into idle indices: prompts prefilled in one batched pass (K8), search state
 *                      initialised; the streams already decoding are untouched.  max_length[n] like wl_generate's.
 *   wl_session_run     runs the device-side token loop over every admitted stream for at most max_steps steps; with
 *                      break_on_finish it also returns as soon as some stream has finished.  done_out[capacity]: 1 for
 *                      indices whose stream is finished and not yet collected.  steps_ran: token steps executed.
 *   wl_session_collect hypotheses of one finished index (outputs like one stream of wl_generate); the index goes idle.
 *   wl_session_close   drops whatever is still in flight.
 * The session has its own decode state and self-attention cache: wl_generate / wl_align / wl_detect_language /
 * wl_encode may be called between two wl_session_run calls (temperature-fallback retries, word alignment of a finished
 * window, the encoder pass of a stream about to be admitted). */
int wl_session_open(wl_ctx* ctx, const wl_gen_opts* opts, int32_t capacity);
int wl_session_admit(wl_ctx* ctx, int32_t n, const int32_t* index, const int32_t* slots, const int32_t* prompts,
                     const int32_t* prompt_off, const int32_t* max_length);
int wl_session_run(wl_ctx* ctx, int32_t max_steps, int32_t break_on_finish, int32_t* done_out, int32_t* steps_ran);
int wl_session_collect(wl_ctx* ctx, int32_t index, int32_t* out_ids, int32_t* out_len, float* out_score, float* out_no_speech,
                       int32_t* out_steps);
int wl_session_close(wl_ctx* ctx);

/* K13. probs [B, n_lang] softmax over the language tokens after feeding <|startoftranscript|>. */
int wl_detect_language(wl_ctx* ctx, const int32_t* slots, int32_t B, float* probs);

/* K14. Teacher-forced pass over start_seq + <|notimestamps|> + text + <|endoftext|> per stream.
 *   text/text_off[B+1]; num_frames[B]; pairs_out [cap_pairs][2] (text_idx, time_idx) concatenated at
 *   pair_off[B+1] (written); tok_probs concatenated like text. */
int wl_align(wl_ctx* ctx, const int32_t* slots, int32_t B, const int32_t* start_seq, int32_t n_start, const int32_t* text,
             const int32_t* text_off, const int32_t* num_frames, int32_t median_width, int32_t* pairs_out,
             int32_t cap_pairs, int32_t* pair_off, float* tok_probs);

/* Diagnostics / parity hooks (used by tests and bench.py, not by the reference-facing path) */
/* teacher-forced logits: tokens concatenated at tok_off[B+1]; logits_out [sum T, vocab] float32 */
int wl_decode_logits(wl_ctx* ctx, const int32_t* slots, int32_t B, const int32_t* tokens, const int32_t* tok_off,
                     float* logits_out);
/* C[z] = A[z] (MxK) * B[z]^T (NxK) (+bias[n]) on the tcgen05 path (use_simt=0) or the CUDA-core checker */
int wl_test_gemm(wl_ctx* ctx, const uint16_t* a_f16, const uint16_t* b_f16, const float* bias, float* c, int32_t M, int32_t N,
                 int32_t K, int32_t batch, int32_t transposed_store, int32_t gelu, int32_t use_simt);
/* test hook of the small-batch decode GEMM (csrc/wgemm.cu, R <= 32): out[R][n_out] = X[R][K] W[n_out][K]^T with the fused
 * epilogue `mode` -- 0: + bias; 1: out += acc + bias (residual in place); 2: gelu(acc + bias) through fp16;
 * 3: split-K partial sums (K > 1280), summed by the hook; mode | 8 (8, 9, 10): the cluster split-K GEMM (cgemm, any R)
 * with epilogue 0, 1, 2 */
int wl_test_wgemm(wl_ctx* ctx, const uint16_t* w_f16, const uint16_t* x_f16, const float* bias, float* out, int32_t R,
                  int32_t n_out, int32_t K, int32_t mode);
/* device-resident timing of the GEMM kernel: C = A(MxK) * B(NxK)^T, `iters` launches between CUDA events;
 * bn = 0 picks the tile like the engine does. ms_out = average milliseconds per launch. */
int wl_bench_gemm(wl_ctx* ctx, int32_t M, int32_t N, int32_t K, int32_t batch, int32_t iters, int32_t flags,
                  float* ms_out); /* flags: 1 transposed store, 2 bias, 4 GELU, 8 fp32 output + fp32 residual */
/* launches of library kernels since wl_init (gpu_launches accounting in bench.py) */
int64_t wl_kernel_launches(wl_ctx* ctx);
/* time (ms, CUDA events on the library stream) of the last wl_mel / wl_encode / wl_generate device work;
 * which = 3: average ms per cross-attention kernel launch since wl_profile_cross_attn(ctx, 1), 4: launches timed;
 * 2 also holds the last wl_session_run, 5 the last wl_session_admit */
float wl_last_device_ms(wl_ctx* ctx, int32_t which /*0 mel, 1 encode, 2 generate, 3/4 cross-attention profile*/);
/* enable = 1: wl_generate calls made WITHOUT a CUDA graph bracket every cross-attention launch (K11, the dominant
 * decode kernel) with CUDA events on the library stream -- bench.py's live roofline measurement.  Resets the sums. */
int wl_profile_cross_attn(wl_ctx* ctx, int32_t enable);
/* K1 with the features kept in HBM (replaces FeatureExtractor + pad_or_trim + the feature upload of encode inside
 * B200WhisperModel.transcribe_batch; reference call sites transcriber_faster_whisper.py:862, :1115-1127, :1348):
 * wl_mel_device computes the log-mel of B waveforms (frames_out[b] = len/160 + 1, like wl_mel) and keeps it resident
 * until the next wl_mel_device call; wl_encode_windows encodes B windows cut from it -- window w is frames
 * [seek, seek + len) of stream win_stream[w], zero-padded to 3000 on the device -- into fresh encoder slots. */
int wl_mel_device(wl_ctx* ctx, const float* pcm, const int64_t* offsets, int32_t B, int32_t* frames_out);
int wl_encode_windows(wl_ctx* ctx, int32_t B, const int32_t* win_stream, const int32_t* win_seek, const int32_t* win_len,
                      int32_t* slots_out);
/* resident-input variants for bench.py `value`: inputs already uploaded by the previous call of the
 * host variant are reused (no H2D, no D2H) */
int wl_mel_resident(wl_ctx* ctx);
int wl_encode_resident(wl_ctx* ctx, int32_t B, const int32_t* slots);

#ifdef __cplusplus
}
#endif
#endif /* WLB200_H */
