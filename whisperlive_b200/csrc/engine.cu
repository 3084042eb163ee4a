// libwlb200 engine: context, weights, encoder (K2-K7), decoder loop (K8-K13), alignment (K14) and
// the C ABI declared in include/wlb200.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/wlb200.h"
#include "gemm.cuh"
#include "kernels.cuh"

using namespace wl;

namespace wl {
void gemm_prime();
void attention_prime();
void search_prime();
void flash_attn_prime();
void encoder_attention_fused(cudaStream_t st, const __half* qk, const __half* vt, __half* out, int nb, int H, int d);
long other_launch_count();
void gemm_tl_bind(unsigned long long* p);
void attention_tl_bind(unsigned long long* p);
void elementwise_tl_bind(unsigned long long* p);
void search_tl_bind(unsigned long long* p);
}  // namespace wl

static std::string g_init_error;

struct EncLayer {
  __half *w_qk, *w_v, *w_o, *w_fc1, *w_fc2;
  float *b_qk, *b_v, *b_o, *b_fc1, *b_fc2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};
struct DecLayer {
  __half *w_qkv, *w_o, *w_qc, *w_kc, *w_vc, *w_oc, *w_fc1, *w_fc2;
  float *b_qkv, *b_o, *b_qc, *b_vc, *b_oc, *b_fc1, *b_fc2;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  long kernels = 0;  // kernel nodes per replay
};

struct wl_ctx {
  wl_config cfg;
  std::vector<int32_t> align_heads;
  std::string err;
  cudaStream_t st = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms[6] = {0, 0, 0, 0, 0, 0};   // [0] mel, [1] encode, [2] generate / session run, [5] session admit (prefill)
  // per-kernel profiling of the dominant decode kernel (bench.py roofline): events around every cross-attention launch
  unsigned* post_bar = nullptr;   // grid-barrier words of the fused split-K consumers
  int prof_cross = 0;
  cudaEvent_t pev0 = nullptr, pev1 = nullptr;
  double prof_cross_ms = 0.0;
  long prof_cross_n = 0;
  int num_sms = 148;
  int d, H, Le, Ld, n_mels, V, Vld, Bm, Km, Rm, NS;
  bool finalized = false;
  std::vector<void*> allocs;
  std::map<std::string, void*> dev;                 // raw uploaded tensors (fp16 for ndim>=2, f32 for 1-D)
  std::map<std::string, std::vector<int64_t>> shape;
  long graph_launched = 0;   // kernels executed through graph replays
  long capture_counted = 0;  // launcher calls that were captured, not executed

  // weights
  __half *w_conv1 = nullptr, *w_conv2 = nullptr, *emb = nullptr, *pos_dec = nullptr;
  float *b_conv1, *b_conv2, *pos_enc, *lnp_g, *lnp_b, *lnf_g, *lnf_b;
  std::vector<EncLayer> enc;
  std::vector<DecLayer> dec;
  // mel
  float *mel_window, *mel_twiddle, *mel_filt;
  int* mel_range;
  float* mel_pcm = nullptr;
  float* mel_out = nullptr;
  long *mel_off = nullptr, *mel_ooff = nullptr;
  unsigned* mel_gmax = nullptr;
  long mel_pcm_cap = 0, mel_out_cap = 0;
  int mel_last_B = 0, mel_last_frames = 0;
  // encoder workspaces
  int EB, AB;
  float* feat32;
  __half *feat16, *conv1o, *xn, *qk, *vt, *probs16, *attn, *hbuf;
  float *x, *scores;
  int* enc_slots_dev;
  // slot pool
  __half* enc16;   // [NS][1500][d]
  __half* ckv;     // [Ld][2][NS][H][1500][64]
  std::vector<int> slot_free;
  std::vector<char> slot_used;
  // decoder workspaces
  float *dx, *part1, *part2, *logits;
  __half *dxn, *datt, *dh, *kcache, *vcache;
  long cache_row_stride, cache_layer_stride;
  CrossAttnWorkspace xws;
  float* align_probs = nullptr;   // [R][H][1500]
  float* align_buf = nullptr;     // [B][nh][T_MAX][1500]
  long align_buf_cap = 0;
  int* align_heads_dev = nullptr;
  DecodeState ds;
  unsigned* suppress_mask;
  // pinned host staging
  int* h_int = nullptr;     // generic int staging
  float* h_flt = nullptr;
  size_t h_int_cap = 0, h_flt_cap = 0;
  std::map<std::string, GraphEntry> graphs;
  // WLB200_TIMELINE: in-graph per-kernel timestamps (common.cuh), dumped after every wl_generate
  unsigned long long* tl_dev = nullptr;
  std::string tl_path;
  // resident log-mel of the last wl_mel_device call (features never leave the GPU between mel and encoder)
  float *res_pcm = nullptr, *res_mel = nullptr;
  long res_pcm_cap = 0, res_mel_cap = 0;
  long *res_off = nullptr, *res_ooff = nullptr;
  int *res_frames = nullptr, *win_meta = nullptr;
  std::vector<int> res_frames_h;
  // K8 batched prefill workspaces (allocated on first use, grown on demand)
  struct Prefill {
    long cap_rows = 0;
    int *tok = nullptr, *pos = nullptr, *active = nullptr, *wrow = nullptr, *vslot = nullptr, *vdone = nullptr, *sel = nullptr;
    short* src = nullptr;
    float *x = nullptr, *qkv = nullptr, *qc = nullptr, *xpart = nullptr;
    __half *xn = nullptr, *att = nullptr, *h = nullptr;
    long rows_done = 0, calls = 0;   // statistics
    // K14 (align through the batched pass)
    int* row_b = nullptr;
    long row_b_cap = 0;
    float *aprobs = nullptr, *mat = nullptr, *tokp = nullptr;
    int *aT = nullptr, *anf = nullptr, *path = nullptr, *path_len = nullptr;
    long tokp_cap = 0;
  } pf;
  float* stage_f32 = nullptr;   // wl_load_tensor staging (freed by wl_finalize_weights)
  size_t stage_cap = 0;
  // Decode session (N2, step-level continuous batching): a second decode state + self-attention cache whose stream
  // indices are admitted, decoded for a bounded number of token steps and collected independently of each other.
  // One-shot calls (wl_generate / wl_align / wl_detect_language) keep using `ds` / `kcache`, so they may run between two
  // wl_session_run calls without disturbing the streams in flight.
  struct Session {
    bool allocated = false, open = false;
    int cap = 0, K = 1, Kr = 1, NH = 1, nsplit = 1, use_graph = 1;
    float length_penalty = 1.f;
    SearchOpts so;
    DecodeState ds;
    __half *kcache = nullptr, *vcache = nullptr;
    unsigned* mask = nullptr;
    int* idx_dev = nullptr;
    std::vector<int> hp, meta;          // host shadows: prompts [cap][T_MAX], per-stream metadata [10][cap]
    std::vector<char> used, finished;   // index holds an admitted stream / that stream has finished decoding
    int live = 0;                       // admitted and still decoding
    long steps = 0, runs = 0, admitted = 0;
  } sess;
};

#define API_BEGIN(ctx)                                          \
  if (!(ctx)) return WL_ERR_ARG;                                \
  try {                                                         \
    WL_CUDA(cudaSetDevice((ctx)->cfg.device));
#define API_END(ctx)                                            \
  }                                                             \
  catch (const wl::Error& e) {                                  \
    (ctx)->err = e.msg;                                         \
    return e.code;                                              \
  }                                                             \
  catch (const std::exception& e) {                             \
    (ctx)->err = e.what();                                      \
    return WL_ERR_STATE;                                        \
  }                                                             \
  return WL_OK;

template <class T>
static T* dalloc(wl_ctx* c, size_t n, bool zero = true) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) {
    char b[256];
    snprintf(b, sizeof(b), "cudaMalloc of %.1f MB failed: %s", n * sizeof(T) / 1048576.0, cudaGetErrorString(e));
    throw wl::Error{WL_ERR_NOMEM, b};
  }
  c->allocs.push_back(p);
  if (zero) WL_CUDA(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
  return (T*)p;
}

static void ensure_host(wl_ctx* c, size_t n_int, size_t n_flt) {
  if (n_int > c->h_int_cap) {
    if (c->h_int) cudaFreeHost(c->h_int);
    WL_CUDA(cudaMallocHost((void**)&c->h_int, n_int * sizeof(int)));
    c->h_int_cap = n_int;
  }
  if (n_flt > c->h_flt_cap) {
    if (c->h_flt) cudaFreeHost(c->h_flt);
    WL_CUDA(cudaMallocHost((void**)&c->h_flt, n_flt * sizeof(float)));
    c->h_flt_cap = n_flt;
  }
}

// device-resident decode state for Bm streams x Km rows (one per context, plus one per decode session)
static void alloc_decode_state(wl_ctx* c, DecodeState& s) {
  const size_t R = c->Rm, B = c->Bm;
  s.tok_in = dalloc<int>(c, R); s.pos = dalloc<int>(c, R); s.active = dalloc<int>(c, R); s.cum = dalloc<float>(c, R);
  s.gen_len = dalloc<int>(c, R); s.last_ts = dalloc<int>(c, R); s.row_done = dalloc<int>(c, R);
  s.hist = dalloc<int>(c, R * T_MAX); s.src = dalloc<short>(c, R * T_MAX);
  s.cand_val = dalloc<float>(c, R * MAX_CAND); s.cand_tok = dalloc<int>(c, R * MAX_CAND);
  s.nospeech_row = dalloc<float>(c, R);
  s.slot = dalloc<int>(c, B); s.prompt = dalloc<int>(c, B * T_MAX); s.prompt_len = dalloc<int>(c, B);
  s.fed = dalloc<int>(c, B); s.sot_index = dalloc<int>(c, B); s.use_ts = dalloc<int>(c, B); s.n_new = dalloc<int>(c, B);
  s.step = dalloc<int>(c, B); s.done = dalloc<int>(c, B); s.n_alive = dalloc<int>(c, B); s.no_speech = dalloc<float>(c, B);
  s.hyp_count = dalloc<int>(c, B); s.hyp_cum = dalloc<float>(c, B * MAX_HYPS); s.hyp_len = dalloc<int>(c, B * MAX_HYPS);
  s.hyp_tok = dalloc<int>(c, B * MAX_HYPS * T_MAX); s.steps_run = dalloc<int>(c, B); s.n_done = dalloc<int>(c, 1);
  s.force_len = dalloc<int>(c, B); s.force_prob = dalloc<float>(c, B * T_MAX);
  s.seed = dalloc<unsigned>(c, 1); s.steps_left = dalloc<int>(c, 1);
  s.pre_n = dalloc<int>(c, B); s.pre_last = dalloc<int>(c, B); s.pre_penult = dalloc<int>(c, B); s.pre_lts = dalloc<int>(c, B);
  s.brk = dalloc<int>(c, 2);
}

// ------------------------------------------------------------------------------------------ init
extern "C" int wl_init(const wl_config* cfg, wl_ctx** out) {
  if (!cfg || !out) return WL_ERR_ARG;
  wl_ctx* c = new wl_ctx();
  try {
    WL_CHECK(cfg->abi_version == WL_ABI_VERSION, WL_ERR_ARG, "ABI version mismatch: header %d, caller %d", WL_ABI_VERSION,
             cfg->abi_version);
    c->cfg = *cfg;
    c->align_heads.assign(cfg->align_heads, cfg->align_heads + 2 * cfg->n_align_heads);
    c->cfg.align_heads = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    WL_CHECK(e == cudaSuccess && ndev > 0, WL_ERR_CUDA, "no CUDA device available (%s): libwlb200 has no CPU fallback",
             cudaGetErrorString(e));
    WL_CHECK(cfg->device >= 0 && cfg->device < ndev, WL_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    WL_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    WL_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    WL_CHECK(prop.major == 10, WL_ERR_CUDA, "libwlb200 is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
    c->num_sms = prop.multiProcessorCount;
    c->d = cfg->d_model; c->H = cfg->n_heads; c->Le = cfg->enc_layers; c->Ld = cfg->dec_layers;
    c->n_mels = cfg->n_mels; c->V = cfg->vocab; c->Vld = (cfg->vocab + 3) / 4 * 4;
    c->Bm = cfg->max_streams; c->Km = cfg->max_beam; c->Rm = c->Bm * c->Km; c->NS = cfg->enc_slots;
    WL_CHECK(c->d % 64 == 0 && c->H * 64 == c->d && c->d <= 1280, WL_ERR_ARG, "d_model %d / heads %d unsupported", c->d, c->H);
    WL_CHECK(c->n_mels % 8 == 0 && c->n_mels <= 128, WL_ERR_ARG, "n_mels %d unsupported", c->n_mels);
    WL_CHECK(c->Km >= 1 && c->Km <= MAX_ROWS_PER_STREAM, WL_ERR_ARG, "max_beam %d must be in [1,%d]", c->Km, MAX_ROWS_PER_STREAM);
    WL_CHECK(c->Bm >= 1 && c->NS >= c->Bm, WL_ERR_ARG, "enc_slots %d must be >= max_streams %d", c->NS, c->Bm);
    WL_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    WL_CUDA(cudaEventCreate(&c->ev0));
    WL_CUDA(cudaEventCreate(&c->ev1));
    gemm_prime();
    dec_gemm_prime();
    wgemm_prime();
    cgemm_prime();
    attention_prime();
    search_prime();
    flash_attn_prime();
    if (const char* tl = getenv("WLB200_TIMELINE")) {
      c->tl_path = tl;
      WL_CUDA(cudaMalloc((void**)&c->tl_dev, (size_t)(TL_CAP + 1) * 8));
      WL_CUDA(cudaMemset(c->tl_dev, 0, (size_t)(TL_CAP + 1) * 8));
      c->allocs.push_back(c->tl_dev);
      gemm_tl_bind(c->tl_dev); dec_gemm_tl_bind(c->tl_dev); wgemm_tl_bind(c->tl_dev); attention_tl_bind(c->tl_dev); elementwise_tl_bind(c->tl_dev); search_tl_bind(c->tl_dev);
    }
    c->enc.resize(c->Le);
    c->dec.resize(c->Ld);
    for (int i = c->NS - 1; i >= 0; --i) c->slot_free.push_back(i);
    c->slot_used.assign(c->NS, 0);
  } catch (const wl::Error& e) {
    g_init_error = e.msg;
    delete c;
    return e.code;
  }
  *out = c;
  return WL_OK;
}

extern "C" void wl_destroy(wl_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->st);
  for (auto& g : c->graphs)
    if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  for (void* p : c->allocs) cudaFree(p);
  if (c->stage_f32) cudaFree(c->stage_f32);
  if (c->h_int) cudaFreeHost(c->h_int);
  if (c->h_flt) cudaFreeHost(c->h_flt);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->pev0) cudaEventDestroy(c->pev0);
  if (c->pev1) cudaEventDestroy(c->pev1);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

extern "C" const char* wl_last_error(wl_ctx* c) { return c ? c->err.c_str() : g_init_error.c_str(); }
extern "C" int64_t wl_kernel_launches(wl_ctx* c) {
  return c ? gemm_launch_count() + dec_gemm_launch_count() + wgemm_launch_count() + cgemm_launch_count() + other_launch_count() - c->capture_counted + c->graph_launched : 0;
}
extern "C" float wl_last_device_ms(wl_ctx* c, int32_t which) {
  if (!c) return -1.f;
  if (which >= 0 && which < 3) return c->last_ms[which];
  if (which == 3) return c->prof_cross_n > 0 ? (float)(c->prof_cross_ms / (double)c->prof_cross_n) : -1.f;   // avg ms per cross-attention launch
  if (which == 4) return (float)c->prof_cross_n;
  if (which == 5) return c->last_ms[5];   // last wl_session_admit (upload + batched prefill + init)
  return -1.f;
}
extern "C" int wl_profile_cross_attn(wl_ctx* c, int32_t enable) {
  API_BEGIN(c)
  if (enable && !c->pev0) {
    WL_CUDA(cudaEventCreate(&c->pev0));
    WL_CUDA(cudaEventCreate(&c->pev1));
  }
  c->prof_cross = enable ? 1 : 0;
  c->prof_cross_ms = 0.0;
  c->prof_cross_n = 0;
  API_END(c)
}

// ------------------------------------------------------------------------------------------ weights
extern "C" int wl_load_tensor(wl_ctx* c, const char* name, const float* data, const int64_t* shape, int32_t ndim) {
  API_BEGIN(c)
  WL_CHECK(name && data && shape && ndim >= 1 && ndim <= 3, WL_ERR_ARG, "wl_load_tensor: bad arguments");
  WL_CHECK(!c->finalized, WL_ERR_STATE, "weights already finalized");
  std::string nm(name);
  size_t n = 1;
  std::vector<int64_t> sh(shape, shape + ndim);
  for (auto s : sh) n *= (size_t)s;
  const bool as_f32 = ndim == 1 || nm == "model.encoder.embed_positions.weight" || nm == "mel_filters";
  if (as_f32) {
    float* p = dalloc<float>(c, n, false);
    WL_CUDA(cudaMemcpy(p, data, n * sizeof(float), cudaMemcpyHostToDevice));
    c->dev[nm] = p;
  } else {
    // fp32 -> fp16 (and the conv re-layout) on the device: the host only hands over its buffer.  A large-v3 load
    // is 1.5 G values; converting them on one host thread took longer than everything else in wl_init together.
    if (n > c->stage_cap) {
      if (c->stage_f32) cudaFree(c->stage_f32);
      c->stage_f32 = nullptr;
      c->stage_cap = 0;
      WL_CUDA(cudaMalloc((void**)&c->stage_f32, n * sizeof(float)));
      c->stage_cap = n;
    }
    WL_CUDA(cudaMemcpyAsync(c->stage_f32, data, n * sizeof(float), cudaMemcpyHostToDevice, c->st));
    __half* p = dalloc<__half>(c, n, false);
    if (ndim == 3) cast_weight_f16(c->st, c->stage_f32, p, sh[0], sh[1], sh[2]);   // [co][ci][k] -> [co][k][ci]
    else cast_weight_f16(c->st, c->stage_f32, p, (long)n, 1, 1);
    WL_CUDA(cudaStreamSynchronize(c->st));
    c->dev[nm] = p;
  }
  c->shape[nm] = sh;
  API_END(c)
}

static void* need(wl_ctx* c, const std::string& nm, std::initializer_list<int64_t> want) {
  auto it = c->dev.find(nm);
  WL_CHECK(it != c->dev.end(), WL_ERR_STATE, "missing weight tensor '%s'", nm.c_str());
  const auto& sh = c->shape[nm];
  std::vector<int64_t> w(want);
  WL_CHECK(sh == w, WL_ERR_ARG, "weight '%s' has the wrong shape", nm.c_str());
  return it->second;
}
static __half* concat_h(wl_ctx* c, std::vector<std::pair<__half*, size_t>> parts) {
  size_t tot = 0;
  for (auto& p : parts) tot += p.second;
  __half* out = dalloc<__half>(c, tot, false);
  size_t o = 0;
  for (auto& p : parts) {
    WL_CUDA(cudaMemcpy(out + o, p.first, p.second * sizeof(__half), cudaMemcpyDeviceToDevice));
    o += p.second;
  }
  return out;
}
static float* concat_f(wl_ctx* c, std::vector<std::pair<float*, size_t>> parts) {
  size_t tot = 0;
  for (auto& p : parts) tot += p.second;
  float* out = dalloc<float>(c, tot, true);
  size_t o = 0;
  for (auto& p : parts) {
    if (p.first) WL_CUDA(cudaMemcpy(out + o, p.first, p.second * sizeof(float), cudaMemcpyDeviceToDevice));
    o += p.second;
  }
  return out;
}

static void build_mel_tables(wl_ctx* c) {
  std::vector<float> win(400), tw(800);
  const double PI = 3.14159265358979323846;
  for (int i = 0; i < 400; ++i) {
    win[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / 400.0));
    tw[2 * i] = (float)cos(2.0 * PI * i / 400.0);
    tw[2 * i + 1] = (float)sin(2.0 * PI * i / 400.0);
  }
  c->mel_window = dalloc<float>(c, 400);
  c->mel_twiddle = dalloc<float>(c, 800);
  WL_CUDA(cudaMemcpy(c->mel_window, win.data(), 400 * 4, cudaMemcpyHostToDevice));
  WL_CUDA(cudaMemcpy(c->mel_twiddle, tw.data(), 800 * 4, cudaMemcpyHostToDevice));
  c->mel_filt = (float*)need(c, "mel_filters", {c->n_mels, 201});
  std::vector<float> f((size_t)c->n_mels * 201);
  WL_CUDA(cudaMemcpy(f.data(), c->mel_filt, f.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<int> rg(2 * c->n_mels);
  for (int m = 0; m < c->n_mels; ++m) {
    int lo = 201, hi = 0;
    for (int k = 0; k < 201; ++k)
      if (f[(size_t)m * 201 + k] != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
    if (lo > hi) lo = hi = 0;
    rg[2 * m] = lo; rg[2 * m + 1] = hi;
  }
  c->mel_range = dalloc<int>(c, rg.size());
  WL_CUDA(cudaMemcpy(c->mel_range, rg.data(), rg.size() * 4, cudaMemcpyHostToDevice));
  c->mel_gmax = dalloc<unsigned>(c, c->Bm);
  c->res_off = dalloc<long>(c, c->Bm + 1);
  c->res_ooff = dalloc<long>(c, c->Bm + 1);
  c->res_frames = dalloc<int>(c, c->Bm);
  c->win_meta = dalloc<int>(c, 3 * (size_t)c->Bm);
  c->mel_off = dalloc<long>(c, c->Bm + 1);
  c->mel_ooff = dalloc<long>(c, c->Bm + 1);
}

extern "C" int wl_finalize_weights(wl_ctx* c) {
  API_BEGIN(c)
  WL_CHECK(!c->finalized, WL_ERR_STATE, "weights already finalized");
  const int64_t d = c->d, ff = 4 * c->d, nm = c->n_mels, V = c->V;
  const size_t dd = (size_t)d * d;
  const std::string E = "model.encoder.", D = "model.decoder.";
  c->w_conv1 = (__half*)need(c, E + "conv1.weight", {d, nm, 3});
  c->b_conv1 = (float*)need(c, E + "conv1.bias", {d});
  c->w_conv2 = (__half*)need(c, E + "conv2.weight", {d, d, 3});
  c->b_conv2 = (float*)need(c, E + "conv2.bias", {d});
  c->pos_enc = (float*)need(c, E + "embed_positions.weight", {S_ENC, d});
  c->lnp_g = (float*)need(c, E + "layer_norm.weight", {d});
  c->lnp_b = (float*)need(c, E + "layer_norm.bias", {d});
  for (int l = 0; l < c->Le; ++l) {
    const std::string p = E + "layers." + std::to_string(l) + ".";
    EncLayer& L = c->enc[l];
    L.w_qk = concat_h(c, {{(__half*)need(c, p + "self_attn.q_proj.weight", {d, d}), dd},
                          {(__half*)need(c, p + "self_attn.k_proj.weight", {d, d}), dd}});
    L.b_qk = concat_f(c, {{(float*)need(c, p + "self_attn.q_proj.bias", {d}), (size_t)d}, {nullptr, (size_t)d}});
    L.w_v = (__half*)need(c, p + "self_attn.v_proj.weight", {d, d});
    L.b_v = (float*)need(c, p + "self_attn.v_proj.bias", {d});
    L.w_o = (__half*)need(c, p + "self_attn.out_proj.weight", {d, d});
    L.b_o = (float*)need(c, p + "self_attn.out_proj.bias", {d});
    L.ln1_g = (float*)need(c, p + "self_attn_layer_norm.weight", {d});
    L.ln1_b = (float*)need(c, p + "self_attn_layer_norm.bias", {d});
    L.w_fc1 = (__half*)need(c, p + "fc1.weight", {ff, d});
    L.b_fc1 = (float*)need(c, p + "fc1.bias", {ff});
    L.w_fc2 = (__half*)need(c, p + "fc2.weight", {d, ff});
    L.b_fc2 = (float*)need(c, p + "fc2.bias", {d});
    L.ln2_g = (float*)need(c, p + "final_layer_norm.weight", {d});
    L.ln2_b = (float*)need(c, p + "final_layer_norm.bias", {d});
  }
  c->emb = (__half*)need(c, D + "embed_tokens.weight", {V, d});
  c->pos_dec = (__half*)need(c, D + "embed_positions.weight", {T_MAX, d});
  c->lnf_g = (float*)need(c, D + "layer_norm.weight", {d});
  c->lnf_b = (float*)need(c, D + "layer_norm.bias", {d});
  for (int l = 0; l < c->Ld; ++l) {
    const std::string p = D + "layers." + std::to_string(l) + ".";
    DecLayer& L = c->dec[l];
    L.w_qkv = concat_h(c, {{(__half*)need(c, p + "self_attn.q_proj.weight", {d, d}), dd},
                           {(__half*)need(c, p + "self_attn.k_proj.weight", {d, d}), dd},
                           {(__half*)need(c, p + "self_attn.v_proj.weight", {d, d}), dd}});
    L.b_qkv = concat_f(c, {{(float*)need(c, p + "self_attn.q_proj.bias", {d}), (size_t)d},
                           {nullptr, (size_t)d},
                           {(float*)need(c, p + "self_attn.v_proj.bias", {d}), (size_t)d}});
    L.w_o = (__half*)need(c, p + "self_attn.out_proj.weight", {d, d});
    L.b_o = (float*)need(c, p + "self_attn.out_proj.bias", {d});
    L.ln1_g = (float*)need(c, p + "self_attn_layer_norm.weight", {d});
    L.ln1_b = (float*)need(c, p + "self_attn_layer_norm.bias", {d});
    L.w_qc = (__half*)need(c, p + "encoder_attn.q_proj.weight", {d, d});
    L.b_qc = (float*)need(c, p + "encoder_attn.q_proj.bias", {d});
    L.w_kc = (__half*)need(c, p + "encoder_attn.k_proj.weight", {d, d});
    L.w_vc = (__half*)need(c, p + "encoder_attn.v_proj.weight", {d, d});
    L.b_vc = (float*)need(c, p + "encoder_attn.v_proj.bias", {d});
    L.w_oc = (__half*)need(c, p + "encoder_attn.out_proj.weight", {d, d});
    L.b_oc = (float*)need(c, p + "encoder_attn.out_proj.bias", {d});
    L.ln2_g = (float*)need(c, p + "encoder_attn_layer_norm.weight", {d});
    L.ln2_b = (float*)need(c, p + "encoder_attn_layer_norm.bias", {d});
    L.w_fc1 = (__half*)need(c, p + "fc1.weight", {ff, d});
    L.b_fc1 = (float*)need(c, p + "fc1.bias", {ff});
    L.w_fc2 = (__half*)need(c, p + "fc2.weight", {d, ff});
    L.b_fc2 = (float*)need(c, p + "fc2.bias", {d});
    L.ln3_g = (float*)need(c, p + "final_layer_norm.weight", {d});
    L.ln3_b = (float*)need(c, p + "final_layer_norm.bias", {d});
  }
  build_mel_tables(c);
  if (c->stage_f32) { cudaFree(c->stage_f32); c->stage_f32 = nullptr; c->stage_cap = 0; }

  // ---- encoder workspaces (EB streams per pass, AB streams per attention sub-pass)
  const int H = c->H;
  // streams per encoder pass: 16 x 1500 = 24000 rows fill the 148 SMs' tile waves better than 12000 (7 waves at 91 % vs
  // 13 at 98 % for the 2560-wide projection); the workspaces are ~1 GB at large-v3, nothing next to 180 GB
  static const int enc_batch = [] { const char* e = getenv("WLB200_ENC_BATCH"); return e ? std::max(1, atoi(e)) : 16; }();
  c->EB = std::min(c->Bm, enc_batch);
  c->AB = std::min(c->EB, d >= 1024 ? 2 : 4);
  const size_t M = (size_t)c->EB * S_ENC;
  c->feat32 = dalloc<float>(c, (size_t)c->Bm * nm * 3000);  // all streams of a call stay resident (wl_encode_resident)
  c->feat16 = dalloc<__half>(c, (size_t)c->EB * 3002 * nm + 4096);
  c->conv1o = dalloc<__half>(c, (size_t)c->EB * 3002 * d + 4096);
  c->x = dalloc<float>(c, M * d);
  c->xn = dalloc<__half>(c, M * d);
  c->qk = dalloc<__half>(c, M * 2 * d);
  c->vt = dalloc<__half>(c, (size_t)c->EB * d * S_PAD);
  c->scores = dalloc<float>(c, (size_t)c->AB * H * S_ENC * S_PAD);
  c->probs16 = dalloc<__half>(c, (size_t)c->AB * H * S_ENC * S_PAD);
  c->attn = dalloc<__half>(c, M * d);
  c->hbuf = dalloc<__half>(c, M * ff);
  c->enc_slots_dev = dalloc<int>(c, c->Bm);
  // ---- slot pool
  c->enc16 = dalloc<__half>(c, (size_t)c->NS * S_ENC * d, false);
  c->ckv = dalloc<__half>(c, (size_t)c->Ld * 2 * c->NS * S_ENC * d, false);
  // ---- decoder workspaces
  const size_t R = c->Rm, Rp = (R + 15) / 16 * 16;
  c->dx = dalloc<float>(c, R * d);
  // split-K partial sums: up to 16 K ranges of a d-wide output, 4 of the 3d-wide QKV, 3 of the 4d-wide MLP
  c->part1 = dalloc<float>(c, R * 16 * (size_t)d + R * 4 * (size_t)ff);
  c->part2 = dalloc<float>(c, R * 16 * (size_t)d);
  c->logits = dalloc<float>(c, R * c->Vld);
  c->dxn = dalloc<__half>(c, Rp * d);
  c->datt = dalloc<__half>(c, Rp * d);
  c->dh = dalloc<__half>(c, Rp * ff);
  c->cache_row_stride = (long)H * T_MAX * 64;
  c->cache_layer_stride = (long)R * c->cache_row_stride;
  c->kcache = dalloc<__half>(c, (size_t)c->Ld * c->cache_layer_stride, false);
  c->vcache = dalloc<__half>(c, (size_t)c->Ld * c->cache_layer_stride, false);
  c->xws.part = dalloc<float>(c, (size_t)c->Bm * H * 12 * MAX_ROWS_PER_STREAM * 66);
  c->xws.probs = nullptr;
  c->post_bar = dalloc<unsigned>(c, 2);
  c->suppress_mask = dalloc<unsigned>(c, (V + 31) / 32 + 1);
  if (!c->align_heads.empty()) {
    c->align_heads_dev = dalloc<int>(c, c->align_heads.size());
    WL_CUDA(cudaMemcpy(c->align_heads_dev, c->align_heads.data(), c->align_heads.size() * 4, cudaMemcpyHostToDevice));
  }
  alloc_decode_state(c, c->ds);
  WL_CUDA(cudaDeviceSynchronize());
  c->finalized = true;
  API_END(c)
}

// ------------------------------------------------------------------------------------------ K1 mel
static void mel_run(wl_ctx* c) {
  MelTables t{c->mel_window, c->mel_twiddle, c->mel_filt, c->mel_range, c->n_mels};
  WL_CUDA(cudaEventRecord(c->ev0, c->st));
  mel_forward(c->st, c->mel_pcm, c->mel_off, c->mel_out, c->mel_ooff, c->mel_gmax, t, c->mel_last_B, c->mel_last_frames);
  WL_CUDA(cudaEventRecord(c->ev1, c->st));
}

extern "C" int wl_mel(wl_ctx* c, const float* pcm, const int64_t* offsets, int32_t B, float* out, const int64_t* out_offsets) {
  API_BEGIN(c)
  WL_CHECK(c->finalized, WL_ERR_STATE, "weights not finalized");
  WL_CHECK(pcm && offsets && out && out_offsets && B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_mel: bad arguments (B=%d, max %d)", B, c->Bm);
  const long total = offsets[B] - offsets[0];
  int max_frames = 0;
  std::vector<long> off(B + 1), ooff(B + 1);
  for (int b = 0; b <= B; ++b) off[b] = offsets[b] - offsets[0];
  for (int b = 0; b < B; ++b) {
    const long n = off[b + 1] - off[b];
    WL_CHECK(n > 0, WL_ERR_ARG, "wl_mel: empty waveform for stream %d", b);
    const int T = (int)(n / 160) + 1;
    max_frames = std::max(max_frames, T);
    WL_CHECK(out_offsets[b + 1] - out_offsets[b] == (long)T * c->n_mels, WL_ERR_ARG, "wl_mel: out_offsets do not match frames");
  }
  for (int b = 0; b <= B; ++b) ooff[b] = out_offsets[b] - out_offsets[0];
  if (total > c->mel_pcm_cap) {
    c->mel_pcm = dalloc<float>(c, total + total / 4, false);
    c->mel_pcm_cap = total + total / 4;
  }
  if (ooff[B] > c->mel_out_cap) {
    c->mel_out = dalloc<float>(c, ooff[B] + ooff[B] / 4, false);
    c->mel_out_cap = ooff[B] + ooff[B] / 4;
  }
  WL_CUDA(cudaMemcpyAsync(c->mel_pcm, pcm + offsets[0], total * sizeof(float), cudaMemcpyHostToDevice, c->st));
  WL_CUDA(cudaMemcpyAsync(c->mel_off, off.data(), (B + 1) * sizeof(long), cudaMemcpyHostToDevice, c->st));
  WL_CUDA(cudaMemcpyAsync(c->mel_ooff, ooff.data(), (B + 1) * sizeof(long), cudaMemcpyHostToDevice, c->st));
  c->mel_last_B = B;
  c->mel_last_frames = max_frames;
  mel_run(c);
  WL_CUDA(cudaMemcpyAsync(out + out_offsets[0], c->mel_out, ooff[B] * sizeof(float), cudaMemcpyDeviceToHost, c->st));
  WL_CUDA(cudaStreamSynchronize(c->st));
  WL_CUDA(cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1));
  API_END(c)
}

// K1 with the result kept on the device: PCM up, log-mel stays in HBM until the next wl_mel_device call; the windows the
// encoder consumes are gathered from it on the device (wl_encode_windows).  Round 1 copied the features to the host,
// zero-padded them there and copied them back (110 MB of PCIe traffic and ~28 ms per 32-stream step).
extern "C" int wl_mel_device(wl_ctx* c, const float* pcm, const int64_t* offsets, int32_t B, int32_t* frames_out) {
  API_BEGIN(c)
  WL_CHECK(c->finalized, WL_ERR_STATE, "weights not finalized");
  WL_CHECK(pcm && offsets && frames_out && B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_mel_device: bad arguments (B=%d, max %d)", B, c->Bm);
  const long total = offsets[B] - offsets[0];
  int max_frames = 0;
  std::vector<long> off(B + 1), ooff(B + 1, 0);
  c->res_frames_h.assign(B, 0);
  for (int b = 0; b <= B; ++b) off[b] = offsets[b] - offsets[0];
  for (int b = 0; b < B; ++b) {
    const long n = off[b + 1] - off[b];
    WL_CHECK(n > 0, WL_ERR_ARG, "wl_mel_device: empty waveform for stream %d", b);
    const int T = (int)(n / 160) + 1;
    max_frames = std::max(max_frames, T);
    c->res_frames_h[b] = T;
    frames_out[b] = T;
    ooff[b + 1] = ooff[b] + (long)T * c->n_mels;
  }
  if (total > c->res_pcm_cap) {
    c->res_pcm = dalloc<float>(c, total + total / 4, false);
    c->res_pcm_cap = total + total / 4;
  }
  if (ooff[B] > c->res_mel_cap) {
    c->res_mel = dalloc<float>(c, ooff[B] + ooff[B] / 4, false);
    c->res_mel_cap = ooff[B] + ooff[B] / 4;
  }
  cudaStream_t st = c->st;
  WL_CUDA(cudaMemcpyAsync(c->res_pcm, pcm + offsets[0], total * sizeof(float), cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(c->res_off, off.data(), (B + 1) * sizeof(long), cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(c->res_ooff, ooff.data(), (B + 1) * sizeof(long), cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(c->res_frames, c->res_frames_h.data(), B * sizeof(int), cudaMemcpyHostToDevice, st));
  MelTables t{c->mel_window, c->mel_twiddle, c->mel_filt, c->mel_range, c->n_mels};
  WL_CUDA(cudaEventRecord(c->ev0, st));
  mel_forward(st, c->res_pcm, c->res_off, c->res_mel, c->res_ooff, c->mel_gmax, t, B, max_frames);
  WL_CUDA(cudaEventRecord(c->ev1, st));
  WL_CUDA(cudaStreamSynchronize(st));   // the caller's PCM and the staging vectors above may go away
  WL_CUDA(cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1));
  API_END(c)
}

static void encode_impl(wl_ctx* c, const float* features_host, int B, const int* slots, bool resident);

extern "C" int wl_encode_windows(wl_ctx* c, int32_t B, const int32_t* win_stream, const int32_t* win_seek, const int32_t* win_len,
                                 int32_t* slots_out) {
  API_BEGIN(c)
  WL_CHECK(c->finalized && win_stream && win_seek && win_len && slots_out && B >= 1 && B <= c->Bm, WL_ERR_ARG,
           "wl_encode_windows: bad arguments (B=%d, max %d)", B, c->Bm);
  const int ns = (int)c->res_frames_h.size();
  for (int b = 0; b < B; ++b) {
    WL_CHECK(win_stream[b] >= 0 && win_stream[b] < ns, WL_ERR_ARG, "wl_encode_windows: window %d names stream %d (wl_mel_device holds %d)", b, win_stream[b], ns);
    WL_CHECK(win_seek[b] >= 0 && win_len[b] >= 0 && win_len[b] <= 3000 && win_seek[b] + win_len[b] <= c->res_frames_h[win_stream[b]],
             WL_ERR_ARG, "wl_encode_windows: window %d [%d, +%d) is outside the %d resident frames", b, win_seek[b], win_len[b],
             c->res_frames_h[win_stream[b]]);
  }
  WL_CHECK((int)c->slot_free.size() >= B, WL_ERR_NOMEM, "wl_encode_windows: %d encoder slots requested, %d free", B, (int)c->slot_free.size());
  for (int b = 0; b < B; ++b) {
    slots_out[b] = c->slot_free.back();
    c->slot_free.pop_back();
    c->slot_used[slots_out[b]] = 1;
  }
  try {
    std::vector<int> meta(3 * (size_t)B);
    for (int b = 0; b < B; ++b) { meta[b] = win_stream[b]; meta[B + b] = win_seek[b]; meta[2 * B + b] = win_len[b]; }
    WL_CUDA(cudaMemcpyAsync(c->win_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
    gather_windows(c->st, c->res_mel, c->res_ooff, c->res_frames, c->win_meta, c->win_meta + B, c->win_meta + 2 * B, c->feat32, B, c->n_mels);
    WL_CUDA(cudaStreamSynchronize(c->st));   // meta goes out of scope
    encode_impl(c, nullptr, B, slots_out, true);
  } catch (...) {
    for (int b = 0; b < B; ++b) { c->slot_used[slots_out[b]] = 0; c->slot_free.push_back(slots_out[b]); }
    throw;
  }
  API_END(c)
}

extern "C" int wl_mel_resident(wl_ctx* c) {
  API_BEGIN(c)
  WL_CHECK(c->mel_last_B > 0, WL_ERR_STATE, "wl_mel_resident: call wl_mel first");
  mel_run(c);
  WL_CUDA(cudaStreamSynchronize(c->st));
  WL_CUDA(cudaEventElapsedTime(&c->last_ms[0], c->ev0, c->ev1));
  API_END(c)
}

// ------------------------------------------------------------------------------------------ K2-K7 encoder
static GemmOperand opnd(const __half* p, long rows, long k, long ld, int n1 = 1, long s1 = 0, int n2 = 1, long s2 = 0) {
  GemmOperand o;
  o.ptr = p; o.rows = rows; o.k = k; o.ld = ld; o.n1 = n1; o.s1 = s1; o.n2 = n2; o.s2 = s2;
  return o;
}

static void encoder_pass(wl_ctx* c, int nb, const int* slots_host, const float* feat_dev) {
  const int d = c->d, H = c->H, nm = c->n_mels, ff = 4 * c->d;
  const long M = (long)nb * S_ENC;
  cudaStream_t st = c->st;
  prep_features(st, feat_dev, c->feat16, nb, nm);
  {  // conv1 + GELU -> conv1o rows 1..3000 (rows 0 / 3001 stay zero)
    GemmEpilogue e;
    e.out = c->conv1o + d; e.out_f32 = 0; e.ldm = d; e.ldn = 1; e.ob1 = 3002L * d; e.bias = c->b_conv1; e.gelu = 1;
    gemm_tn(st, opnd(c->feat16, 3000, 3 * nm, nm, nb, 3002L * nm), opnd(c->w_conv1, d, 3 * nm, 3 * nm), 3000, d, 3 * nm, e);
  }
  {  // conv2 (stride 2) + GELU + positional table -> residual stream x (f32)
    GemmEpilogue e;
    e.out = c->x; e.out_f32 = 1; e.ldm = d; e.ldn = 1; e.ob1 = (long)S_ENC * d; e.bias = c->b_conv2; e.gelu = 1;
    e.resid = c->pos_enc; e.rldm = d; e.rldn = 1; e.rb1 = 0;
    gemm_tn(st, opnd(c->conv1o, S_ENC, 3 * d, 2 * d, nb, 3002L * d), opnd(c->w_conv2, d, 3 * d, 3 * d), S_ENC, d, 3 * d, e);
  }
  for (int l = 0; l < c->Le; ++l) {
    const EncLayer& L = c->enc[l];
    layernorm_rows(st, c->x, L.ln1_g, L.ln1_b, c->xn, nullptr, M, d);
    {
      GemmEpilogue e;
      e.out = c->qk; e.ldm = 2 * d; e.bias = L.b_qk;
      gemm_tn(st, opnd(c->xn, M, d, d), opnd(L.w_qk, 2 * d, d, d), (int)M, 2 * d, d, e);
    }
    {  // V^T[b] = Wv * xn[b]^T  (swap-AB) so that P*V reads V K-major
      GemmEpilogue e;
      e.out = c->vt; e.ldm = S_PAD; e.ldn = 1; e.ob1 = (long)d * S_PAD; e.bias = L.b_v; e.bias_on_m = 1;
      gemm_tn(st, opnd(L.w_v, d, d, d), opnd(c->xn, S_ENC, d, d, nb, (long)S_ENC * d), d, S_ENC, d, e);
    }
    static const bool fused_attn = [] { const char* e = getenv("WLB200_FUSED_ATTN"); return e ? atoi(e) != 0 : true; }();
    if (fused_attn) encoder_attention_fused(st, c->qk, c->vt, c->attn, nb, H, d);
    else for (int b0 = 0; b0 < nb; b0 += c->AB) {
      const int ab = std::min(c->AB, nb - b0);
      const __half* qb = c->qk + (long)b0 * S_ENC * 2 * d;
      {
        GemmEpilogue e;
        e.out = c->scores; e.out_f32 = 1; e.ldm = S_PAD; e.ob1 = (long)S_ENC * S_PAD; e.ob2 = (long)H * S_ENC * S_PAD;
        gemm_tn(st, opnd(qb, S_ENC, 64, 2 * d, H, 64, ab, (long)S_ENC * 2 * d),
                opnd(qb + d, S_ENC, 64, 2 * d, H, 64, ab, (long)S_ENC * 2 * d), S_ENC, S_ENC, 64, e);
      }
      softmax_rows(st, c->scores, c->probs16, (long)ab * H * S_ENC, S_ENC, S_PAD, S_PAD, 0.125f);
      {
        GemmEpilogue e;
        e.out = c->attn + (long)b0 * S_ENC * d; e.ldm = d; e.ob1 = 64; e.ob2 = (long)S_ENC * d;
        gemm_tn(st, opnd(c->probs16, S_ENC, S_PAD, S_PAD, H, (long)S_ENC * S_PAD, ab, (long)H * S_ENC * S_PAD),
                opnd(c->vt + (long)b0 * d * S_PAD, 64, S_PAD, S_PAD, H, 64L * S_PAD, ab, (long)d * S_PAD), S_ENC, 64, S_PAD, e);
      }
    }
    {
      GemmEpilogue e;
      e.out = c->x; e.out_f32 = 1; e.ldm = d; e.bias = L.b_o; e.resid = c->x; e.rldm = d;
      gemm_tn(st, opnd(c->attn, M, d, d), opnd(L.w_o, d, d, d), (int)M, d, d, e);
    }
    layernorm_rows(st, c->x, L.ln2_g, L.ln2_b, c->xn, nullptr, M, d);
    {
      GemmEpilogue e;
      e.out = c->hbuf; e.ldm = ff; e.bias = L.b_fc1; e.gelu = 1;
      gemm_tn(st, opnd(c->xn, M, d, d), opnd(L.w_fc1, ff, d, d), (int)M, ff, d, e);
    }
    {
      GemmEpilogue e;
      e.out = c->x; e.out_f32 = 1; e.ldm = d; e.bias = L.b_fc2; e.resid = c->x; e.rldm = d;
      gemm_tn(st, opnd(c->hbuf, M, ff, ff), opnd(L.w_fc2, d, ff, ff), (int)M, d, ff, e);
    }
  }
  layernorm_rows(st, c->x, c->lnp_g, c->lnp_b, c->xn, nullptr, M, d);
  for (int b = 0; b < nb; ++b)
    WL_CUDA(cudaMemcpyAsync(c->enc16 + (long)slots_host[b] * S_ENC * d, c->xn + (long)b * S_ENC * d,
                            (size_t)S_ENC * d * sizeof(__half), cudaMemcpyDeviceToDevice, st));
  // K7: cross-attention K/V of every decoder layer straight into the slot pool ([slot][H][1500][64])
  const long slot_sz = (long)S_ENC * d;
  for (int l = 0; l < c->Ld; ++l) {
    const DecLayer& L = c->dec[l];
    GemmEpilogue e;
    e.mode = GEMM_HEADSPLIT; e.hs_S = S_ENC; e.hs_H = H; e.hs_slot_stride = slot_sz; e.hs_slots = c->enc_slots_dev;
    e.out = c->ckv + ((long)l * 2 + 0) * c->NS * slot_sz;
    gemm_tn(st, opnd(c->xn, M, d, d), opnd(L.w_kc, d, d, d), (int)M, d, d, e);
    e.out = c->ckv + ((long)l * 2 + 1) * c->NS * slot_sz;
    e.bias = L.b_vc;
    gemm_tn(st, opnd(c->xn, M, d, d), opnd(L.w_vc, d, d, d), (int)M, d, d, e);
  }
}

static void encode_impl(wl_ctx* c, const float* features_host, int B, const int* slots, bool resident) {
  const size_t per = (size_t)c->n_mels * 3000;
  float ms_total = 0.f;
  for (int b0 = 0; b0 < B; b0 += c->EB) {
    const int nb = std::min(c->EB, B - b0);
    if (!resident)
      WL_CUDA(cudaMemcpyAsync(c->feat32 + b0 * per, features_host + b0 * per, nb * per * sizeof(float), cudaMemcpyHostToDevice, c->st));
    WL_CUDA(cudaMemcpyAsync(c->enc_slots_dev, slots + b0, nb * sizeof(int), cudaMemcpyHostToDevice, c->st));
    WL_CUDA(cudaEventRecord(c->ev0, c->st));
    encoder_pass(c, nb, slots + b0, c->feat32 + b0 * per);
    WL_CUDA(cudaEventRecord(c->ev1, c->st));
    WL_CUDA(cudaStreamSynchronize(c->st));
    float ms;
    WL_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    ms_total += ms;
  }
  c->last_ms[1] = ms_total;
}

extern "C" int wl_encode(wl_ctx* c, const float* features, int32_t B, int32_t* slots_out) {
  API_BEGIN(c)
  WL_CHECK(c->finalized, WL_ERR_STATE, "weights not finalized");
  WL_CHECK(features && slots_out && B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_encode: bad arguments (B=%d, max %d)", B, c->Bm);
  WL_CHECK((int)c->slot_free.size() >= B, WL_ERR_NOMEM, "wl_encode: %d encoder slots requested, %d free", B, (int)c->slot_free.size());
  for (int b = 0; b < B; ++b) {
    slots_out[b] = c->slot_free.back();
    c->slot_free.pop_back();
    c->slot_used[slots_out[b]] = 1;
  }
  try {
    encode_impl(c, features, B, slots_out, false);
  } catch (...) {
    for (int b = 0; b < B; ++b) { c->slot_used[slots_out[b]] = 0; c->slot_free.push_back(slots_out[b]); }
    throw;
  }
  API_END(c)
}

extern "C" int wl_encode_resident(wl_ctx* c, int32_t B, const int32_t* slots) {
  API_BEGIN(c)
  WL_CHECK(B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_encode_resident: B must be <= %d", c->Bm);
  for (int b = 0; b < B; ++b) WL_CHECK(slots[b] >= 0 && slots[b] < c->NS && c->slot_used[slots[b]], WL_ERR_ARG, "bad slot");
  encode_impl(c, nullptr, B, slots, true);
  API_END(c)
}

extern "C" int wl_slots_release(wl_ctx* c, const int32_t* slots, int32_t n) {
  API_BEGIN(c)
  for (int i = 0; i < n; ++i) {
    WL_CHECK(slots[i] >= 0 && slots[i] < c->NS && c->slot_used[slots[i]], WL_ERR_ARG, "wl_slots_release: slot %d is not in use", slots[i]);
    c->slot_used[slots[i]] = 0;
    c->slot_free.push_back(slots[i]);
  }
  API_END(c)
}
extern "C" int wl_slots_free_count(wl_ctx* c) { return c ? (int)c->slot_free.size() : -1; }

extern "C" int wl_encoder_output(wl_ctx* c, int32_t slot, float* out) {
  API_BEGIN(c)
  WL_CHECK(out && slot >= 0 && slot < c->NS && c->slot_used[slot], WL_ERR_ARG, "wl_encoder_output: bad slot %d", slot);
  const size_t n = (size_t)S_ENC * c->d;
  std::vector<__half> h(n);
  WL_CUDA(cudaMemcpy(h.data(), c->enc16 + (long)slot * n, n * sizeof(__half), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) out[i] = __half2float(h[i]);
  API_END(c)
}

// ------------------------------------------------------------------------------------------ decoder step
namespace wl {
void gather_align_probs(cudaStream_t st, const DecodeState& s, const float* probs, float* buf, const int* heads, int n_heads,
                        int layer, int B, int rows_per_stream, int H);
}

// Switches that are read at every call and are part of the graph key, so that one process can sweep them
// (tools/sweep_prefetch.py): WLB200_XA_PREFETCH, WLB200_CGEMM
static int xa_prefetch_streams() {
  const char* e = getenv("WLB200_XA_PREFETCH");
  return e ? atoi(e) : 0;
}
// Measured on B200 (bench.py, large-v3, beam 4): the cluster split-K GEMM (cgemm) made the token step SLOWER than split-K
// partials summed by the consumers -- 3.81 vs 2.87 ms at 32 streams, 262 vs 206 ms per 8-stream batch -- so it is off by
// default; what it costs is the serial DSMEM reduction and two cluster barriers after the MMAs.
static bool cgemm_enabled() {
  const char* e = getenv("WLB200_CGEMM");
  return e ? atoi(e) != 0 : false;
}

static void decode_step(wl_ctx* c, int B, int Kr, const SearchOpts& so, const VocabIds& vi, int nsplit, bool align_mode) {
  const int d = c->d, H = c->H, ff = 4 * c->d, R = B * Kr;
  cudaStream_t st = c->st;
  const DecodeState& s = c->ds;
  const long slot_sz = (long)S_ENC * d;
  PdlScope pdl(true);   // every kernel of the step is launched as a programmatic dependent of its predecessor
  decoder_embed(st, s, c->emb, c->pos_dec, c->dx, R, d);
  auto swap_gemm = [&](const __half* W, int n_out, int K, const __half* X, GemmEpilogue e) {
    e.ldm = 1;
    e.bias_on_m = 1;
    gemm_tn(st, opnd(W, n_out, K, K), opnd(X, R, K, K), n_out, R, K, e);
  };
  static const bool splitk = [] { const char* e = getenv("WLB200_SPLITK"); return e ? atoi(e) != 0 : true; }();
  // Decode GEMMs are weight-streaming (M = out features, N = rows <= 256): K is split over enough CTAs to fill the
  // SMs; every K range stores its raw fp32 partial sum and the CONSUMER (LayerNorm, attention, GELU) adds the
  // ranges and the bias in a fixed order -- no atomics, bit-reproducible, and no separate reduction kernel.
  // part1 holds activations (qkv, q_cross, fc1), part2 the residual updates (out-proj, fc2) until the next LayerNorm.
  // WLB200_FUSE_POST=1 (default OFF): run the LayerNorm-update after out-proj / FC2 and the GELU-cast after FC1 inside
  // the producing split-K GEMM behind a grid barrier (13 -> 9 launches per layer).  Measured slower on B200 than the
  // separate kernels chained by programmatic dependent launch (32 streams: 116 vs 99 ms per 26 tokens; 4 streams:
  // 59 vs 49 ms): the barrier gates every CTA on the slowest one and the row work then runs on 70 CTAs instead of
  // 128.  Kept as a switch for the round-2 persistent-layer work.
  static const bool fuse_env = [] { const char* e = getenv("WLB200_FUSE_POST"); return e ? atoi(e) != 0 : false; }();
  static const bool simt_env = [] { const char* e = getenv("WLB200_GEMM_SIMT"); return e && atoi(e) != 0; }();
  const bool fuse = fuse_env && splitk && !simt_env;
  struct Post { int kind = GEMM_POST_NONE; const float* g = nullptr; const float* b = nullptr; };
  auto part_gemm = [&](const __half* W, int n_out, int K, const __half* X, float* buf, const float* bias, Post post,
                       int max_split = 8) -> PartialSrc {
    GemmEpilogue e;
    e.out = buf; e.out_f32 = 1; e.ldn = n_out; e.ldm = 1;
    e.a_static = 1;
    PartialSrc ps;
    ps.ptr = buf; ps.bias = bias; ps.stride = (long)c->Rm * n_out;
    if (splitk) {
      int s2 = std::min(gemm_split_plan(n_out, R, K), max_split);
      const int total_kb = cdiv(K, 64);
      while (s2 > 1 && cdiv(total_kb, cdiv(total_kb, s2)) != s2) --s2;   // every K range must be non-empty
      ps.nsplit = s2;
      e.partials = ps.nsplit; e.part_stride = ps.stride;
    } else {
      ps.nsplit = 1;   // single pass, bias still added by the consumer
    }
    static const bool compact = [] { const char* e2 = getenv("WLB200_DEC_GEMM"); return e2 ? atoi(e2) != 0 : true; }();
    if (compact && splitk && post.kind == GEMM_POST_NONE && !simt_env) {
      ps.nsplit = dec_gemm_split_plan(n_out, R, K, max_split);
      dec_gemm(st, W, n_out, K, X, R, buf, n_out, ps.stride, ps.nsplit);
      // WLB200_DUP=1 (experiment): launch every decode GEMM twice (idempotent) -- the second launch finds its code in
      // the instruction caches; the in-graph timeline shows what a warm launch of the same kernel costs
      static const bool dup = [] { const char* e2 = getenv("WLB200_DUP"); return e2 && atoi(e2) != 0; }();
      if (dup) dec_gemm(st, W, n_out, K, X, R, buf, n_out, ps.stride, ps.nsplit);
      return ps;
    }
    if (post.kind != GEMM_POST_NONE) {
      e.post = post.kind;
      e.post_bias = bias;
      e.post_bar = c->post_bar;
      if (post.kind == GEMM_POST_LN) { e.post_x = c->dx; e.post_g = post.g; e.post_b = post.b; e.post_y = c->dxn; }
      else e.post_y = c->dh;
    }
    gemm_tn(st, opnd(W, n_out, K, K), opnd(X, R, K, K), n_out, R, K, e);
    return ps;
  };
  auto ln_post = [&](const float* g, const float* b) { Post p; if (fuse) { p.kind = GEMM_POST_LN; p.g = g; p.b = b; } return p; };
  PartialSrc pending;   // residual update not yet folded into x (unfused path)
  auto cross = [&](int l, const PartialSrc& qc) {
    CrossAttnWorkspace ws = c->xws;
    ws.probs = align_mode ? c->align_probs : nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    const bool prof = c->prof_cross && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone;
    if (prof) { ws.ev0 = c->pev0; ws.ev1 = c->pev1; }
    decoder_cross_attn(st, s, qc, c->ckv + ((long)l * 2 + 0) * c->NS * slot_sz, c->ckv + ((long)l * 2 + 1) * c->NS * slot_sz,
                       slot_sz, ws, c->datt, B, Kr, H, d, nsplit);
    if (prof) {   // profiling pass only: serialises the host with the device
      float ms = 0.f;
      WL_CUDA(cudaEventSynchronize(c->pev1));
      WL_CUDA(cudaEventElapsedTime(&ms, c->pev0, c->pev1));
      c->prof_cross_ms += ms;
      c->prof_cross_n += 1;
    }
    if (align_mode)
      gather_align_probs(st, s, c->align_probs, c->align_buf, c->align_heads_dev, (int)c->align_heads.size() / 2, l, B, Kr, H);
  };
  // Small batches (R <= 32 decoder rows, e.g. 4-8 streams per GPU with beam 4): every linear layer is one wgemm launch
  // whose epilogue writes FINAL values (bias, residual, GELU fused), so LayerNorm / attention read one value instead of
  // summing partials and the GELU-cast launch is gone: 12 launches per layer instead of 13, each a fraction of the code.
  static const bool wg_env = [] { const char* e = getenv("WLB200_WGEMM"); return e ? atoi(e) != 0 : true; }();
  // (one m16 tile only: with two, every CTA re-reads 82 KB of X from L2 and the launch costs 8 us -- measured; rows
  // 17..32 take the tcgen05 path below until the K split moves into a cluster)
  static const int wg_max_rows = [] { const char* e = getenv("WLB200_WGEMM_ROWS"); return e ? atoi(e) : 16; }();
  const bool use_wg = wg_env && R <= wg_max_rows && wgemm_supported(R, d) && wgemm_supported(R, ff);
  // Above that: split-K partials summed by the consumers (13 launches per layer), or with WLB200_CGEMM=1 cgemm, the
  // tcgen05 pipeline with the K split inside a cluster (dec_gemm.cu) -- same fused epilogues as wgemm, 12 launches.
  const bool cg_env = cgemm_enabled();
  const bool small = !simt_env && !fuse && (use_wg || cg_env);
  auto plain = [](const float* ptr) { PartialSrc ps; ps.ptr = ptr; ps.nsplit = 1; ps.stride = 0; ps.bias = nullptr; return ps; };
  // mode 0: out_f32 = X W^T + bias; 1: out_f32 += X W^T + bias; 2: out_f16 = gelu(X W^T + bias)
  auto lin = [&](const __half* W, int n_out, int K, const __half* X, const float* bias, int mode, float* of32, __half* of16) {
    if (use_wg) wgemm(st, W, n_out, K, X, R, bias, mode, of32, of16, 0);
    else cgemm(st, W, n_out, K, X, R, bias, mode, of32, of16);
  };
  // WLB200_XA_PREFETCH=n: the layer's first LayerNorm also asks L2 for the encoder K/V of the first n live streams --
  // the six latency-bound kernels between it and the cross-attention leave HBM idle, the cross-attention is HBM-bound
  const int xa_pf = xa_prefetch_streams();
  L2Prefetch pf;
  if (xa_pf > 0 && !align_mode) {
    pf.slot = s.slot; pf.done = s.done; pf.B = B;
    pf.slot_bytes = slot_sz * 2;
    const long per_region = (pf.slot_bytes + 32 * 1024 - 1) / (32 * 1024);
    pf.n_streams = (int)std::min<long>(std::min(xa_pf, B), (long)R * 32 / (2 * per_region));   // one 32 KB piece per lane of warp 0
  }
  for (int l = 0; small && l < c->Ld; ++l) {
    const DecLayer& L = c->dec[l];
    pf.k = c->ckv + ((long)l * 2 + 0) * c->NS * slot_sz;
    pf.v = c->ckv + ((long)l * 2 + 1) * c->NS * slot_sz;
    layernorm_update_rows(st, c->dx, pending, L.ln1_g, L.ln1_b, c->dxn, R, d, &pf);
    lin(L.w_qkv, 3 * d, d, c->dxn, L.b_qkv, 0, c->part1, nullptr);
    decoder_self_attn(st, s, plain(c->part1), c->kcache + (long)l * c->cache_layer_stride, c->vcache + (long)l * c->cache_layer_stride,
                      c->cache_row_stride, c->datt, R, H, d);
    lin(L.w_o, d, d, c->datt, L.b_o, 1, c->dx, nullptr);
    layernorm_update_rows(st, c->dx, PartialSrc(), L.ln2_g, L.ln2_b, c->dxn, R, d);
    lin(L.w_qc, d, d, c->dxn, L.b_qc, 0, c->part1, nullptr);
    cross(l, plain(c->part1));
    lin(L.w_oc, d, d, c->datt, L.b_oc, 1, c->dx, nullptr);
    layernorm_update_rows(st, c->dx, PartialSrc(), L.ln3_g, L.ln3_b, c->dxn, R, d);
    lin(L.w_fc1, ff, d, c->dxn, L.b_fc1, 2, nullptr, c->dh);
    const int ks2 = use_wg ? wgemm_ksplit(ff) : 1;
    if (ks2 == 1) {
      lin(L.w_fc2, d, ff, c->dh, L.b_fc2, 1, c->dx, nullptr);
      pending = PartialSrc();
    } else {   // wgemm with K = 4d split over CTAs: the next LayerNorm folds the ranges (+ bias) into x
      wgemm(st, L.w_fc2, d, ff, c->dh, R, nullptr, 3, c->part2, nullptr, (long)c->Rm * d);
      pending.ptr = c->part2; pending.nsplit = ks2; pending.stride = (long)c->Rm * d; pending.bias = L.b_fc2;
    }
  }
  for (int l = 0; !small && l < c->Ld; ++l) {
    const DecLayer& L = c->dec[l];
    const bool last = l + 1 == c->Ld;
    pf.k = c->ckv + ((long)l * 2 + 0) * c->NS * slot_sz;
    pf.v = c->ckv + ((long)l * 2 + 1) * c->NS * slot_sz;
    if (!fuse || l == 0) layernorm_update_rows(st, c->dx, pending, L.ln1_g, L.ln1_b, c->dxn, R, d, &pf);
    const PartialSrc qkv = part_gemm(L.w_qkv, 3 * d, d, c->dxn, c->part1, L.b_qkv, Post());
    decoder_self_attn(st, s, qkv, c->kcache + (long)l * c->cache_layer_stride, c->vcache + (long)l * c->cache_layer_stride,
                      c->cache_row_stride, c->datt, R, H, d);
    pending = part_gemm(L.w_o, d, d, c->datt, c->part2, L.b_o, ln_post(L.ln2_g, L.ln2_b));
    if (!fuse) layernorm_update_rows(st, c->dx, pending, L.ln2_g, L.ln2_b, c->dxn, R, d);
    const PartialSrc qc = part_gemm(L.w_qc, d, d, c->dxn, c->part1, L.b_qc, Post(), 4);   // cross-attention sums <= 4 ranges
    cross(l, qc);
    pending = part_gemm(L.w_oc, d, d, c->datt, c->part2, L.b_oc, ln_post(L.ln3_g, L.ln3_b));
    if (!fuse) layernorm_update_rows(st, c->dx, pending, L.ln3_g, L.ln3_b, c->dxn, R, d);
    Post gp;
    if (fuse) gp.kind = GEMM_POST_GELU;
    const PartialSrc h1 = part_gemm(L.w_fc1, ff, d, c->dxn, c->part1, L.b_fc1, gp);
    if (!fuse) gelu_cast(st, h1, c->dh, R, ff);
    // FC2's fused LayerNorm is the NEXT layer's ln1 (or the final LayerNorm after the last layer)
    pending = part_gemm(L.w_fc2, d, ff, c->dh, c->part2, L.b_fc2,
                        last ? ln_post(c->lnf_g, c->lnf_b) : ln_post(c->dec[l + 1].ln1_g, c->dec[l + 1].ln1_b));
  }
  if (!fuse) layernorm_update_rows(st, c->dx, pending, c->lnf_g, c->lnf_b, c->dxn, R, d);
  {
    static const bool compact = [] { const char* e2 = getenv("WLB200_DEC_GEMM"); return e2 ? atoi(e2) != 0 : true; }();
    if (compact && !simt_env) {
      dec_gemm(st, c->emb, c->V, d, c->dxn, R, c->logits, c->Vld, 0, 1);
    } else {
      GemmEpilogue e;
      e.out = c->logits; e.out_f32 = 1; e.ldn = c->Vld;
      e.ldm = 1;
      e.a_static = 1;
      gemm_tn(st, opnd(c->emb, c->V, d, d), opnd(c->dxn, R, d, d), c->V, R, d, e);
    }
  }
  search_rows(st, s, c->logits, so, vi, R);
  search_streams(st, s, so, vi, B);
}

// ------------------------------------------------------------------------------------------ K8 batched prefill
constexpr int PF_VCHUNK = 128;   // cross-attention groups (8 rows each) per launch

static void prefill_reserve(wl_ctx* c, long rows) {
  wl_ctx::Prefill& f = c->pf;
  if (rows <= f.cap_rows) return;
  const long cap = (rows + 1023) / 1024 * 1024;
  const int d = c->d, ff = 4 * c->d;
  // (earlier, smaller buffers stay in c->allocs until wl_destroy: growth happens a handful of times per process)
  f.tok = dalloc<int>(c, cap); f.pos = dalloc<int>(c, cap); f.active = dalloc<int>(c, cap); f.wrow = dalloc<int>(c, cap);
  f.vslot = dalloc<int>(c, cap / 8 + PF_VCHUNK); f.vdone = dalloc<int>(c, cap / 8 + PF_VCHUNK);
  if (!f.sel) f.sel = dalloc<int>(c, 3 * (size_t)c->Rm + 16);
  f.src = dalloc<short>(c, cap * T_MAX, false);
  f.x = dalloc<float>(c, cap * d, false); f.qkv = dalloc<float>(c, cap * 3 * d, false); f.qc = dalloc<float>(c, cap * d, false);
  f.xn = dalloc<__half>(c, cap * d, false); f.att = dalloc<__half>(c, cap * d, false); f.h = dalloc<__half>(c, cap * ff, false);
  if (!f.xpart) f.xpart = dalloc<float>(c, (size_t)PF_VCHUNK * c->H * 12 * MAX_ROWS_PER_STREAM * 66, false);
  f.cap_rows = cap;
}

namespace wl {
void gather_align_rows(cudaStream_t st, const float* probs, const int* row_b, const int* row_pos, const int* row_active, float* buf,
                       const int* heads, int n_heads, int layer, int row0, int n_rows, int H);
void align_postprocess(cudaStream_t st, float* buf, float* mat, const int* Tn, const int* nfn, int B, int nh, int width, int n_start,
                       int max_T, int* path_out, int path_cap, int* path_len);
}

// Row layout of a batched decoder pass: stream b runs positions 0 .. ntok[b]-1 as rows [rowbase[b], rowbase[b] + n8)
struct PfRows {
  std::vector<int> tok, pos, act, wrow, vslot, rowbase, row_b;
  long M = 0;
  int NV = 0;
};

// index (optional, [B]): the decode-state index of list entry b (a decode session admits streams into arbitrary free
// indices; a one-shot call uses b itself) -- selects the hp row and the cache row the positions are written to.
static PfRows pf_rows(int B, int Kr, const int* toks, const int* tok_off /*B+1 or null: hp rows of T_MAX*/, const int* ntok,
                      const int32_t* slots, const int32_t* index = nullptr) {
  PfRows r;
  r.rowbase.assign(B, 0);
  for (int b = 0; b < B; ++b) {
    r.rowbase[b] = (int)r.tok.size();
    const int n = ntok[b], n8 = (n + 7) / 8 * 8;
    const int sb = index ? index[b] : b;
    const int* src = tok_off ? toks + tok_off[b] : toks + (size_t)sb * T_MAX;
    for (int i = 0; i < n8; ++i) {
      r.tok.push_back(i < n ? src[i] : 0);
      r.pos.push_back(i < n ? i : 0);
      r.act.push_back(i < n ? 1 : 0);
      r.wrow.push_back(sb * Kr);
      r.row_b.push_back(b);
    }
    for (int g = 0; g < n8 / 8; ++g) r.vslot.push_back(slots[b]);
  }
  r.M = (long)r.tok.size();
  r.NV = (int)r.vslot.size();
  return r;
}

// The decoder stack over M rows (every prompt / teacher-forced position of every stream at once): uploads the row
// tables, runs embed + Ld layers.  align = true additionally captures the cross-attention probabilities of the
// alignment heads into c->align_buf [B][nh][T_MAX][1500].  On return f.x holds the final residual stream of every row.
static void pf_stack(wl_ctx* c, const PfRows& r, bool align) {
  const int d = c->d, H = c->H, ff = 4 * c->d;
  cudaStream_t st = c->st;
  const long M = r.M;
  const int NV = r.NV;
  prefill_reserve(c, M);
  wl_ctx::Prefill& f = c->pf;
  WL_CUDA(cudaMemcpyAsync(f.tok, r.tok.data(), M * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(f.pos, r.pos.data(), M * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(f.active, r.act.data(), M * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(f.wrow, r.wrow.data(), M * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(f.vslot, r.vslot.data(), NV * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemsetAsync(f.vdone, 0, NV * 4, st));
  const int nh = (int)c->align_heads.size() / 2;
  if (align) {
    if (!f.row_b_cap || f.row_b_cap < f.cap_rows) { f.row_b = dalloc<int>(c, f.cap_rows); f.row_b_cap = f.cap_rows; }
    WL_CUDA(cudaMemcpyAsync(f.row_b, r.row_b.data(), M * 4, cudaMemcpyHostToDevice, st));
    if (!f.aprobs) f.aprobs = dalloc<float>(c, (size_t)PF_VCHUNK * 8 * H * S_ENC, false);
  }
  prefill_embed(st, f.tok, f.pos, f.active, f.wrow, c->emb, c->pos_dec, f.x, f.src, (int)M, d);
  DecodeState sv = c->ds;
  sv.active = f.active; sv.pos = f.pos; sv.src = f.src; sv.wrow = f.wrow;
  auto plain = [](const float* ptr) { PartialSrc ps; ps.ptr = ptr; ps.nsplit = 1; ps.stride = 0; ps.bias = nullptr; return ps; };
  const long slot_sz = (long)S_ENC * d;
  for (int l = 0; l < c->Ld; ++l) {
    const DecLayer& L = c->dec[l];
    __half* kc = c->kcache + (long)l * c->cache_layer_stride;
    __half* vc = c->vcache + (long)l * c->cache_layer_stride;
    layernorm_rows(st, f.x, L.ln1_g, L.ln1_b, f.xn, nullptr, M, d);
    {
      GemmEpilogue e;
      e.out = f.qkv; e.out_f32 = 1; e.ldm = 3 * d; e.bias = L.b_qkv;
      gemm_tn(st, opnd(f.xn, M, d, d), opnd(L.w_qkv, 3 * d, d, d), (int)M, 3 * d, d, e);
    }
    prefill_kv_write(st, f.qkv, f.pos, f.active, f.wrow, kc, vc, c->cache_row_stride, (int)M, H, d);
    decoder_self_attn(st, sv, plain(f.qkv), kc, vc, c->cache_row_stride, f.att, (int)M, H, d);
    {
      GemmEpilogue e;
      e.out = f.x; e.out_f32 = 1; e.ldm = d; e.bias = L.b_o; e.resid = f.x; e.rldm = d;
      gemm_tn(st, opnd(f.att, M, d, d), opnd(L.w_o, d, d, d), (int)M, d, d, e);
    }
    layernorm_rows(st, f.x, L.ln2_g, L.ln2_b, f.xn, nullptr, M, d);
    {
      GemmEpilogue e;
      e.out = f.qc; e.out_f32 = 1; e.ldm = d; e.bias = L.b_qc;
      gemm_tn(st, opnd(f.xn, M, d, d), opnd(L.w_qc, d, d, d), (int)M, d, d, e);
    }
    bool capture = false;
    if (align)
      for (int i = 0; i < nh; ++i) capture = capture || c->align_heads[2 * i] == l;
    for (int v0 = 0; v0 < NV; v0 += PF_VCHUNK) {   // groups of 8 rows against their stream's encoder K/V
      const int Bv = std::min(PF_VCHUNK, NV - v0);
      DecodeState sx = c->ds;
      sx.done = f.vdone + v0; sx.slot = f.vslot + v0;
      CrossAttnWorkspace ws;
      ws.part = f.xpart; ws.probs = capture ? f.aprobs : nullptr;
      const int nsp = capture ? 1 : cross_attn_pick_nsplit(Bv, H, c->num_sms, MAX_ROWS_PER_STREAM);   // exact probabilities need the whole key range
      decoder_cross_attn(st, sx, plain(f.qc + (long)v0 * 8 * d), c->ckv + ((long)l * 2 + 0) * c->NS * slot_sz,
                         c->ckv + ((long)l * 2 + 1) * c->NS * slot_sz, slot_sz, ws, f.att + (long)v0 * 8 * d, Bv, MAX_ROWS_PER_STREAM, H, d, nsp);
      if (capture)
        gather_align_rows(st, f.aprobs, f.row_b, f.pos, f.active, c->align_buf, c->align_heads_dev, nh, l, v0 * 8, Bv * 8, H);
    }
    {
      GemmEpilogue e;
      e.out = f.x; e.out_f32 = 1; e.ldm = d; e.bias = L.b_oc; e.resid = f.x; e.rldm = d;
      gemm_tn(st, opnd(f.att, M, d, d), opnd(L.w_oc, d, d, d), (int)M, d, d, e);
    }
    layernorm_rows(st, f.x, L.ln3_g, L.ln3_b, f.xn, nullptr, M, d);
    {
      GemmEpilogue e;
      e.out = f.h; e.ldm = ff; e.bias = L.b_fc1; e.gelu = 1;
      gemm_tn(st, opnd(f.xn, M, d, d), opnd(L.w_fc1, ff, d, d), (int)M, ff, d, e);
    }
    {
      GemmEpilogue e;
      e.out = f.x; e.out_f32 = 1; e.ldm = d; e.bias = L.b_fc2; e.resid = f.x; e.rldm = d;
      gemm_tn(st, opnd(f.h, M, ff, ff), opnd(L.w_fc2, d, ff, ff), (int)M, d, ff, e);
    }
  }
  WL_CUDA(cudaStreamSynchronize(st));   // the row tables of `r` were uploaded asynchronously
  f.rows_done += M;
  f.calls += 1;
}

// softmax(logits of row sel[j])[tgt[j]] -> out[oidx[j]] for n selected rows of f.x: final LayerNorm + vocabulary
// projection through the decode step's own kernels, Rm rows at a time
static void pf_row_probs(wl_ctx* c, const std::vector<int>& sel, const std::vector<int>& tgt, const std::vector<int>& oidx, float* out_dev) {
  cudaStream_t st = c->st;
  wl_ctx::Prefill& f = c->pf;
  const int n = (int)sel.size(), d = c->d;
  for (int i0 = 0; i0 < n; i0 += c->Rm) {
    const int m = std::min(c->Rm, n - i0);
    std::vector<int> up(3 * (size_t)m);
    for (int i = 0; i < m; ++i) { up[i] = sel[i0 + i]; up[m + i] = tgt[i0 + i]; up[2 * m + i] = oidx[i0 + i]; }
    WL_CUDA(cudaMemcpyAsync(f.sel, up.data(), up.size() * 4, cudaMemcpyHostToDevice, st));
    gather_rows(st, f.x, f.sel, c->dx, m, d);
    layernorm_update_rows(st, c->dx, PartialSrc(), c->lnf_g, c->lnf_b, c->dxn, m, d);
    dec_gemm(st, c->emb, c->V, d, c->dxn, m, c->logits, c->Vld, 0, 1);
    row_prob(st, c->logits, c->V, c->Vld, f.sel + m, f.sel + 2 * m, out_dev, m);
    WL_CUDA(cudaStreamSynchronize(st));
  }
}

// K8: all prompt positions but the last of every stream through the decoder stack in one pass.  hp = prompts [B][T_MAX]
// (pinned host), P / sot / slots per stream.  Leaves the self-attention cache filled for positions 0 .. P-2 in the
// stream's first decode row (b * Kr) and the no-speech probability of streams whose sot lies inside the prompt.
static void prefill_forward(wl_ctx* c, int B, int Kr, const int* hp, const int* P, const int* sot, const int32_t* slots,
                            const int32_t* index = nullptr) {
  // (with `index`: hp rows and P / sot columns are addressed by state index, list entry b is state index index[b])
  auto at = [&](int b) { return index ? index[b] : b; };
  if (!index) WL_CUDA(cudaMemsetAsync(c->ds.no_speech, 0, B * sizeof(float), c->st));
  else for (int b = 0; b < B; ++b) WL_CUDA(cudaMemsetAsync(c->ds.no_speech + index[b], 0, sizeof(float), c->st));
  std::vector<int> ntok(B);
  for (int b = 0; b < B; ++b) ntok[b] = P[at(b)] - 1;
  const PfRows r = pf_rows(B, Kr, hp, nullptr, ntok.data(), slots, index);
  if (r.M == 0) return;
  pf_stack(c, r, false);
  std::vector<int> sel, tgt, oidx;
  for (int b = 0; b < B; ++b) {
    const int sb = at(b);
    if (sot[sb] >= 0 && sot[sb] < P[sb] - 1) { sel.push_back(r.rowbase[b] + sot[sb]); tgt.push_back(c->cfg.no_speech); oidx.push_back(sb); }
  }
  if (!sel.empty()) pf_row_probs(c, sel, tgt, oidx, c->ds.no_speech);
}

static VocabIds vocab_ids(wl_ctx* c) {
  VocabIds v;
  v.vocab = c->V; v.vocab_ld = c->Vld; v.eot = c->cfg.eot; v.sot = c->cfg.sot; v.no_speech = c->cfg.no_speech;
  v.no_timestamps = c->cfg.no_timestamps; v.ts_begin = c->cfg.timestamp_begin; v.blank = c->cfg.blank;
  return v;
}

// Per-stream metadata of a decode call (column b of the [10][B] table `meta`; tokens into hp_row[T_MAX]).  Returns the
// decode steps the stream may need without prefill; *n_new_out = the new tokens it may emit.
static int stream_meta(wl_ctx* c, int b, int slot, const int32_t* prompt, int P, int ml, bool forced, int* hp_row, int* meta, int col,
                       int ncol, int* n_new_out) {
  WL_CHECK(P >= 1 && P <= T_MAX, WL_ERR_ARG, "stream %d: prompt length %d out of range", b, P);
  WL_CHECK(slot >= 0 && slot < c->NS && c->slot_used[slot], WL_ERR_ARG, "stream %d: bad encoder slot %d", b, slot);
  int sot = -1;
  for (int i = 0; i < P; ++i) {
    const int t = prompt[i];
    WL_CHECK(t >= 0 && t < c->V, WL_ERR_ARG, "stream %d: token id %d out of range", b, t);
    hp_row[i] = t;
    if (t == c->cfg.sot && sot < 0) sot = i;
  }
  int n_new = 0, steps = P;
  if (!forced) {
    WL_CHECK(ml >= 2 && ml <= T_MAX, WL_ERR_ARG, "stream %d: max_length %d out of range", b, ml);
    n_new = std::min(ml / 2, ml - P);
    WL_CHECK(n_new >= 1, WL_ERR_ARG, "stream %d: prompt of %d tokens leaves no room under max_length %d", b, P, ml);
    steps = P - 1 + n_new;
  }
  // end of the sot sequence = CT2's prompt length: sot, then every following id in [sot, no_timestamps]
  // (language, task, notimestamps); the tokens after it are a prefix that counts as sampled text
  int sb = P;
  if (sot >= 0) {
    sb = sot + 1;
    while (sb < P && prompt[sb] >= c->cfg.sot && prompt[sb] <= c->cfg.no_timestamps) ++sb;
  }
  const int npre = P - sb;
  int lts = -1;
  for (int i = sb; i < P; ++i)
    if (prompt[i] >= c->cfg.timestamp_begin) lts = prompt[i];
  meta[0 * ncol + col] = slot;
  meta[1 * ncol + col] = P;
  meta[2 * ncol + col] = sot;
  meta[3 * ncol + col] = sb >= 1 ? prompt[sb - 1] != c->cfg.no_timestamps : 1;
  meta[4 * ncol + col] = n_new;
  meta[5 * ncol + col] = forced ? P : 0;
  meta[6 * ncol + col] = forced ? 0 : npre;
  meta[7 * ncol + col] = npre >= 1 ? prompt[P - 1] : -1;
  meta[8 * ncol + col] = npre >= 2 ? prompt[P - 2] : -1;
  meta[9 * ncol + col] = forced ? -1 : lts;
  if (n_new_out) *n_new_out = n_new;
  return steps;
}

// prompts [B][T_MAX] + the [10][B] metadata table (pinned host) -> the device state
static void upload_state_tables(wl_ctx* c, const int* hp, const int* meta, int B) {
  const DecodeState& s = c->ds;
  cudaStream_t st = c->st;
  WL_CUDA(cudaMemcpyAsync(s.prompt, hp, (size_t)B * T_MAX * 4, cudaMemcpyHostToDevice, st));
  int* dst[10] = {s.slot, s.prompt_len, s.sot_index, s.use_ts, s.n_new, s.force_len, s.pre_n, s.pre_last, s.pre_penult, s.pre_lts};
  for (int k = 0; k < 10; ++k) WL_CUDA(cudaMemcpyAsync(dst[k], meta + (size_t)k * B, B * 4, cudaMemcpyHostToDevice, st));
}

// upload prompts & per-stream metadata; returns max steps
static int upload_streams(wl_ctx* c, const int32_t* slots, int B, const int32_t* prompts, const int32_t* off, int max_length,
                          bool forced, const int32_t* max_len_ps = nullptr, int* max_new_out = nullptr) {
  ensure_host(c, (size_t)B * (T_MAX + 16), 16);
  int* hp = c->h_int;                      // [B][T_MAX]
  int* meta = c->h_int + (size_t)B * T_MAX;  // slot, len, sot_index, use_ts, n_new, force_len, pre_n, pre_last, pre_penult, pre_lts
  int max_steps = 0, max_new = 0;
  for (int b = 0; b < B; ++b) {
    int n_new = 0;
    const int steps = stream_meta(c, b, slots[b], prompts + off[b], off[b + 1] - off[b], max_len_ps ? max_len_ps[b] : max_length,
                                  forced, hp + (size_t)b * T_MAX, meta, b, B, &n_new);
    max_steps = std::max(max_steps, steps);
    max_new = std::max(max_new, n_new);
  }
  upload_state_tables(c, hp, meta, B);
  if (max_new_out) *max_new_out = max_new;
  return max_steps;
}

// finished hypotheses of one stream (device order) -> the NH best by cum_logprob / len^length_penalty, like CT2
static void emit_hyps(int NH, float length_penalty, int count, const int* h_len, const float* h_cum, const int* h_tok, int32_t* out_ids,
                      int32_t* out_len, float* out_score) {
  const int cnt = std::min(count, MAX_HYPS);
  std::vector<int> order(cnt);
  std::vector<float> score(cnt);
  for (int i = 0; i < cnt; ++i) {
    order[i] = i;
    score[i] = length_penalty == 0.f ? h_cum[i] : h_cum[i] / powf((float)std::max(h_len[i], 1), length_penalty);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int bb) { return score[a] > score[bb]; });
  for (int hh = 0; hh < NH; ++hh) {
    int* dst = out_ids + (size_t)hh * T_MAX;
    if (hh < cnt) {
      const int i = order[hh];
      memcpy(dst, h_tok + (size_t)i * T_MAX, h_len[i] * sizeof(int));
      out_len[hh] = h_len[i];
      out_score[hh] = score[i];
    } else {
      out_len[hh] = -1;
      out_score[hh] = 0.f;
    }
  }
}

// The captured decode step for one call shape (cached per context).  loop_graph: a conditional WHILE node whose body is
// the step + loop_condition (the whole token loop is one launch); else the plain step.  `tag` separates the graphs of the
// one-shot state ("g") from those of the decode session ("s"): the captures bake the state's device pointers in.
static cudaGraphExec_t decode_graph(wl_ctx* c, const char* tag, int B, int Kr, int K, const SearchOpts& so, const VocabIds& vi,
                                    int nsplit, bool loop_graph, long* kernels) {
  cudaStream_t st = c->st;
  char key[160];
  snprintf(key, sizeof(key), "%s/%d/%d/%d/%d/%d/%d/%d/%08x/%d/%d", tag, B, Kr, K, so.max_cand, so.suppress_blank, so.max_initial_ts,
           so.sampling, *(const unsigned*)&so.temperature, loop_graph ? 1 : 0, xa_prefetch_streams() * 2 + (cgemm_enabled() ? 1 : 0));
  GraphEntry& ge = c->graphs[key];
  if (!ge.exec) {
    const long before = gemm_launch_count() + dec_gemm_launch_count() + wgemm_launch_count() + cgemm_launch_count() + other_launch_count();
    cudaGraph_t g = nullptr, cap = nullptr;
    if (loop_graph) {
      WL_CUDA(cudaGraphCreate(&g, 0));
      cudaGraphConditionalHandle h;
      WL_CUDA(cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
      cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
      np.conditional.handle = h;
      np.conditional.type = cudaGraphCondTypeWhile;
      np.conditional.size = 1;
      cudaGraphNode_t node;
      WL_CUDA(cudaGraphAddNode(&node, g, nullptr, 0, &np));
      cudaGraph_t body = np.conditional.phGraph_out[0];
      WL_CUDA(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
      try {
        decode_step(c, B, Kr, so, vi, nsplit, false);
        loop_condition(st, c->ds, h, B);
      } catch (...) {
        cudaStreamEndCapture(st, &cap);
        cudaGraphDestroy(g);
        throw;
      }
      WL_CUDA(cudaStreamEndCapture(st, &cap));
    } else {
      WL_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
      try {
        decode_step(c, B, Kr, so, vi, nsplit, false);
      } catch (...) {
        cudaStreamEndCapture(st, &g);
        throw;
      }
      WL_CUDA(cudaStreamEndCapture(st, &g));
    }
    ge.kernels = gemm_launch_count() + dec_gemm_launch_count() + wgemm_launch_count() + cgemm_launch_count() + other_launch_count() - before;
    c->capture_counted += ge.kernels;
    WL_CUDA(cudaGraphInstantiate(&ge.exec, g, 0));
    cudaGraphDestroy(g);
  }
  *kernels = ge.kernels;
  return ge.exec;
}

extern "C" int wl_generate(wl_ctx* c, const int32_t* slots, int32_t B, const int32_t* prompts, const int32_t* prompt_off,
                           const wl_gen_opts* o, int32_t* out_ids, int32_t* out_len, float* out_score, float* out_no_speech,
                           int32_t* out_steps) {
  API_BEGIN(c)
  WL_CHECK(c->finalized, WL_ERR_STATE, "weights not finalized");
  WL_CHECK(slots && prompts && prompt_off && o && out_ids && out_len && out_score, WL_ERR_ARG, "wl_generate: null argument");
  WL_CHECK(B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_generate: B=%d exceeds max_streams=%d", B, c->Bm);
  WL_CHECK(o->beam_size >= 1 && o->num_hypotheses >= 1, WL_ERR_ARG, "wl_generate: beam_size / num_hypotheses must be >= 1");
  const int K = o->beam_size;
  const int Kr = K > 1 ? K : o->num_hypotheses;
  WL_CHECK(Kr <= c->Km, WL_ERR_ARG, "wl_generate: %d rows per stream exceed max_beam=%d", Kr, c->Km);
  WL_CHECK(K == 1 || o->num_hypotheses <= MAX_HYPS, WL_ERR_ARG, "too many hypotheses");
  WL_CHECK(o->sampling_topk == 0 || o->sampling_topk == 1, WL_ERR_ARG, "sampling_topk must be 0 (full) or 1 (arg-max)");
  WL_CHECK(o->max_length >= 2 && o->max_length <= T_MAX, WL_ERR_ARG, "max_length %d out of range", o->max_length);
  const int R = B * Kr;
  SearchOpts so;
  so.beam = K; so.rows_per_stream = Kr;
  so.max_cand = std::max(1, std::min(MAX_HYPS, (int)lroundf(K * o->patience)));
  so.suppress_blank = o->suppress_blank; so.max_initial_ts = o->max_initial_timestamp_index;
  so.sampling = (K == 1 && o->sampling_topk == 0 && o->sampling_temperature > 0.f) ? 1 : 0;
  so.temperature = o->sampling_temperature; so.seed = o->seed; so.suppress_mask = c->suppress_mask;
  const VocabIds vi = vocab_ids(c);
  cudaStream_t st = c->st;
  // suppress bitmask
  const int nwords = (c->V + 31) / 32 + 1;
  std::vector<unsigned> mask(nwords, 0u);
  for (int i = 0; i < o->n_suppress; ++i) {
    const int t = o->suppress_tokens[i];
    if (t >= 0 && t < c->V) mask[t >> 5] |= 1u << (t & 31);
  }
  int max_new = 0;
  int max_steps = upload_streams(c, slots, B, prompts, prompt_off, o->max_length, false, o->max_length_per_stream, &max_new);
  // K8: every prompt position but the last goes through the decoder in ONE batched pass (WLB200_PREFILL=0: one decode
  // step per prompt token, the round-1 behaviour)
  static const bool prefill_env = [] { const char* e = getenv("WLB200_PREFILL"); return e ? atoi(e) != 0 : true; }();
  const bool prefilled = o->prefill == 1 || (o->prefill == 0 && prefill_env);
  if (prefilled) max_steps = max_new;
  WL_CUDA(cudaMemcpyAsync(c->suppress_mask, mask.data(), nwords * 4, cudaMemcpyHostToDevice, st));
  const unsigned seed_host = o->seed;
  WL_CUDA(cudaMemcpyAsync(c->ds.seed, &seed_host, sizeof(unsigned), cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaEventRecord(c->ev0, st));
  if (prefilled) {
    // upload_streams staged the prompts in pinned host memory: h_int = [B][T_MAX] tokens, then the per-stream metadata
    WL_CUDA(cudaStreamSynchronize(st));
    const int* hp = c->h_int;
    const int* meta = c->h_int + (size_t)B * T_MAX;
    prefill_forward(c, B, Kr, hp, meta + 1 * B, meta + 2 * B, slots);
  }
  decode_init(st, c->ds, so, vi, B, R, prefilled ? 1 : 0);
  const int nsplit = cross_attn_pick_nsplit(B, c->H, c->num_sms, Kr);

  // The whole token loop is ONE graph launch: a conditional WHILE node whose body is the captured decode step; the
  // body's last kernel (loop_condition) keeps the loop alive while some stream is still decoding and the step budget
  // lasts.  No host round trip per token (round 1 synchronised every 4 steps), no wasted steps after the last EOT.
  // WLB200_LOOP_GRAPH=0 falls back to one graph launch per step with a host check every 4 steps.
  static const bool loop_graph = [] { const char* e = getenv("WLB200_LOOP_GRAPH"); return e ? atoi(e) != 0 : true; }();
  cudaGraphExec_t exec = nullptr;
  long graph_kernels = 0;
  bool is_loop = false;
  if (o->use_cuda_graph) {
    exec = decode_graph(c, "g", B, Kr, K, so, vi, nsplit, loop_graph, &graph_kernels);
    is_loop = loop_graph;
  }
  ensure_host(c, (size_t)B * (T_MAX + 16) + (size_t)B * MAX_HYPS * (T_MAX + 2) + 64, (size_t)B * (MAX_HYPS + 2));
  int* h_done = c->h_int;  // reuse (prompts are already on the device: the copies above are stream-ordered)
  WL_CUDA(cudaStreamSynchronize(st));
  if (is_loop) {
    h_done[0] = max_steps;
    WL_CUDA(cudaMemcpyAsync(c->ds.steps_left, h_done, sizeof(int), cudaMemcpyHostToDevice, st));
    WL_CUDA(cudaGraphLaunch(exec, st));
    WL_CUDA(cudaMemcpyAsync(h_done, c->ds.steps_left, sizeof(int), cudaMemcpyDeviceToHost, st));
    WL_CUDA(cudaStreamSynchronize(st));
    c->graph_launched += graph_kernels * (long)(max_steps - h_done[0]);
  } else {
    int ran = 0;
    const int check_every = 4;
    while (ran < max_steps) {
      const int n = std::min(check_every, max_steps - ran);
      for (int i = 0; i < n; ++i) {
        if (exec) {
          WL_CUDA(cudaGraphLaunch(exec, st));
          c->graph_launched += graph_kernels;
        } else {
          decode_step(c, B, Kr, so, vi, nsplit, false);
        }
      }
      ran += n;
      WL_CUDA(cudaMemcpyAsync(h_done, c->ds.n_done, sizeof(int), cudaMemcpyDeviceToHost, st));
      WL_CUDA(cudaStreamSynchronize(st));
      if (*h_done >= B) break;
    }
  }
  WL_CUDA(cudaEventRecord(c->ev1, st));
  // results
  const DecodeState& s = c->ds;
  int* h_cnt = c->h_int;
  int* h_len = h_cnt + B;
  int* h_steps = h_len + (size_t)B * MAX_HYPS;
  int* h_tok = h_steps + B;
  float* h_cum = c->h_flt;
  float* h_ns = h_cum + (size_t)B * MAX_HYPS;
  WL_CUDA(cudaMemcpyAsync(h_cnt, s.hyp_count, B * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_len, s.hyp_len, (size_t)B * MAX_HYPS * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_steps, s.steps_run, B * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_tok, s.hyp_tok, (size_t)B * MAX_HYPS * T_MAX * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_cum, s.hyp_cum, (size_t)B * MAX_HYPS * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_ns, s.no_speech, B * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaStreamSynchronize(st));
  WL_CUDA(cudaEventElapsedTime(&c->last_ms[2], c->ev0, c->ev1));
  if (c->tl_dev) {   // dump + reset the in-graph timeline of this call
    std::vector<unsigned long long> h(TL_CAP + 1);
    WL_CUDA(cudaMemcpy(h.data(), c->tl_dev, h.size() * 8, cudaMemcpyDeviceToHost));
    const size_t n = std::min<size_t>((size_t)(h[0] & 0xffffffffu), TL_CAP);
    if (FILE* f = fopen(c->tl_path.c_str(), "wb")) { fwrite(h.data() + 1, 8, n, f); fclose(f); }
    WL_CUDA(cudaMemset(c->tl_dev, 0, 8));
  }
  for (int b = 0; b < B; ++b) {
    emit_hyps(o->num_hypotheses, o->length_penalty, h_cnt[b], h_len + (size_t)b * MAX_HYPS, h_cum + (size_t)b * MAX_HYPS,
              h_tok + (size_t)b * MAX_HYPS * T_MAX, out_ids + (size_t)b * o->num_hypotheses * T_MAX, out_len + (size_t)b * o->num_hypotheses,
              out_score + (size_t)b * o->num_hypotheses);
    if (out_no_speech) out_no_speech[b] = h_ns[b];
    if (out_steps) out_steps[b] = h_steps[b];
  }
  API_END(c)
}

// ------------------------------------------------------------------------------------------ N2 decode session
// Step-level continuous batching (replaces the run-to-completion batches of the reference's BatchInferenceWorker,
// whisper_live/batch_inference.py:155-187, :259, :334-339): the decode state has `cap` stream indices; a stream is
// admitted into a free index at any token-step boundary (prompt prefilled in one pass, K8), the device-side loop runs
// for a bounded number of steps or until some stream finishes, and a finished stream is collected -- and its index
// refilled -- while the others keep decoding.  Idle indices carry done = 1: every kernel of the step skips them.
struct SessScope {   // the session's state / cache / suppress mask stand in for the one-shot ones inside a session call
  wl_ctx* c;
  explicit SessScope(wl_ctx* ctx) : c(ctx) { swap(); }
  ~SessScope() { swap(); }
  void swap() {
    std::swap(c->ds, c->sess.ds);
    std::swap(c->kcache, c->sess.kcache);
    std::swap(c->vcache, c->sess.vcache);
    std::swap(c->suppress_mask, c->sess.mask);
  }
};

extern "C" int wl_session_open(wl_ctx* c, const wl_gen_opts* o, int32_t capacity) {
  API_BEGIN(c)
  WL_CHECK(c->finalized, WL_ERR_STATE, "weights not finalized");
  WL_CHECK(o, WL_ERR_ARG, "wl_session_open: null options");
  wl_ctx::Session& ss = c->sess;
  WL_CHECK(!ss.open || ss.live == 0, WL_ERR_STATE, "wl_session_open: %d streams of the open session are still decoding", ss.live);
  WL_CHECK(capacity >= 1 && capacity <= c->Bm, WL_ERR_ARG, "wl_session_open: capacity %d exceeds max_streams=%d", capacity, c->Bm);
  WL_CHECK(o->beam_size >= 1 && o->num_hypotheses >= 1, WL_ERR_ARG, "wl_session_open: beam_size / num_hypotheses must be >= 1");
  const int K = o->beam_size, Kr = K > 1 ? K : o->num_hypotheses;
  WL_CHECK(Kr <= c->Km, WL_ERR_ARG, "wl_session_open: %d rows per stream exceed max_beam=%d", Kr, c->Km);
  WL_CHECK(K == 1 || o->num_hypotheses <= MAX_HYPS, WL_ERR_ARG, "too many hypotheses");
  // sampling draws are keyed by (call seed, batch position): a session has neither, so the temperature-fallback rungs stay
  // on wl_generate (the transcriber routes them there)
  WL_CHECK(!(K == 1 && o->sampling_topk == 0 && o->sampling_temperature > 0.f), WL_ERR_ARG, "wl_session_open: sampling is not supported in a decode session");
  if (!ss.allocated) {
    alloc_decode_state(c, ss.ds);
    ss.kcache = dalloc<__half>(c, (size_t)c->Ld * c->cache_layer_stride, false);
    ss.vcache = dalloc<__half>(c, (size_t)c->Ld * c->cache_layer_stride, false);
    ss.mask = dalloc<unsigned>(c, (c->V + 31) / 32 + 1);
    ss.idx_dev = dalloc<int>(c, c->Bm);
    ss.allocated = true;
  }
  ss.cap = capacity; ss.K = K; ss.Kr = Kr; ss.NH = o->num_hypotheses; ss.length_penalty = o->length_penalty;
  ss.use_graph = o->use_cuda_graph;
  SearchOpts& so = ss.so;
  so.beam = K; so.rows_per_stream = Kr;
  so.max_cand = std::max(1, std::min(MAX_HYPS, (int)lroundf(K * o->patience)));
  so.suppress_blank = o->suppress_blank; so.max_initial_ts = o->max_initial_timestamp_index;
  so.sampling = 0; so.temperature = o->sampling_temperature; so.seed = o->seed; so.suppress_mask = ss.mask;
  ss.nsplit = cross_attn_pick_nsplit(capacity, c->H, c->num_sms, Kr);
  const int nwords = (c->V + 31) / 32 + 1;
  std::vector<unsigned> mask(nwords, 0u);
  for (int i = 0; i < o->n_suppress; ++i) {
    const int t = o->suppress_tokens[i];
    if (t >= 0 && t < c->V) mask[t >> 5] |= 1u << (t & 31);
  }
  const int cap = capacity, R = cap * Kr;
  ss.hp.assign((size_t)cap * T_MAX, 0);
  ss.meta.assign((size_t)10 * cap, 0);
  for (int b = 0; b < cap; ++b) { ss.meta[1 * cap + b] = 1; ss.meta[2 * cap + b] = -1; ss.meta[4 * cap + b] = 1; }   // harmless idle values
  ss.used.assign(cap, 0);
  ss.finished.assign(cap, 0);
  ss.live = 0;
  cudaStream_t st = c->st;
  const DecodeState& s = ss.ds;
  std::vector<int> ones(cap, 1);
  const int nd = cap, brk0[2] = {0, 0};
  const unsigned seed_host = o->seed;
  WL_CUDA(cudaMemcpyAsync(ss.mask, mask.data(), nwords * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(s.done, ones.data(), cap * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemsetAsync(s.active, 0, (size_t)R * 4, st));
  WL_CUDA(cudaMemsetAsync(s.hyp_count, 0, (size_t)cap * 4, st));
  WL_CUDA(cudaMemcpyAsync(s.n_done, &nd, 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(s.brk, brk0, 8, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(s.seed, &seed_host, 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaStreamSynchronize(st));   // the host vectors above go out of scope
  ss.open = true;
  API_END(c)
}

extern "C" int wl_session_admit(wl_ctx* c, int32_t n, const int32_t* index, const int32_t* slots, const int32_t* prompts,
                                const int32_t* prompt_off, const int32_t* max_length) {
  API_BEGIN(c)
  wl_ctx::Session& ss = c->sess;
  WL_CHECK(ss.open, WL_ERR_STATE, "wl_session_admit: no open session");
  WL_CHECK(n >= 1 && index && slots && prompts && prompt_off && max_length, WL_ERR_ARG, "wl_session_admit: bad arguments");
  const int cap = ss.cap;
  for (int i = 0; i < n; ++i) {
    WL_CHECK(index[i] >= 0 && index[i] < cap, WL_ERR_ARG, "wl_session_admit: index %d outside the session capacity %d", index[i], cap);
    WL_CHECK(!ss.used[index[i]], WL_ERR_STATE, "wl_session_admit: index %d still holds a stream", index[i]);
    for (int j = 0; j < i; ++j) WL_CHECK(index[j] != index[i], WL_ERR_ARG, "wl_session_admit: index %d listed twice", index[i]);
  }
  // validate + stage everything before touching the session (a bad prompt must not leave a half-admitted stream)
  std::vector<int> hp = ss.hp, meta = ss.meta;
  for (int i = 0; i < n; ++i)
    stream_meta(c, i, slots[i], prompts + prompt_off[i], prompt_off[i + 1] - prompt_off[i], max_length[i], false,
                hp.data() + (size_t)index[i] * T_MAX, meta.data(), index[i], cap, nullptr);
  ss.hp.swap(hp);
  ss.meta.swap(meta);
  ensure_host(c, (size_t)cap * (T_MAX + 16) + n, 16);
  int* php = c->h_int;
  int* pmeta = php + (size_t)cap * T_MAX;
  int* pidx = pmeta + (size_t)10 * cap;
  memcpy(php, ss.hp.data(), ss.hp.size() * 4);
  memcpy(pmeta, ss.meta.data(), ss.meta.size() * 4);
  memcpy(pidx, index, (size_t)n * 4);
  SessScope scope(c);
  cudaStream_t st = c->st;
  WL_CUDA(cudaEventRecord(c->ev0, st));
  // the tables of the streams in flight are rewritten with the values they already hold (nothing runs between two calls)
  upload_state_tables(c, php, pmeta, cap);
  WL_CUDA(cudaMemcpyAsync(ss.idx_dev, pidx, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaStreamSynchronize(st));
  prefill_forward(c, n, ss.Kr, ss.hp.data(), ss.meta.data() + 1 * cap, ss.meta.data() + 2 * cap, slots, index);
  decode_init(st, c->ds, ss.so, vocab_ids(c), n, cap * ss.Kr, 1, ss.idx_dev);
  WL_CUDA(cudaEventRecord(c->ev1, st));
  WL_CUDA(cudaStreamSynchronize(st));
  WL_CUDA(cudaEventElapsedTime(&c->last_ms[5], c->ev0, c->ev1));
  for (int i = 0; i < n; ++i) { ss.used[index[i]] = 1; ss.finished[index[i]] = 0; }
  ss.live += n;
  ss.admitted += n;
  API_END(c)
}

extern "C" int wl_session_run(wl_ctx* c, int32_t max_steps, int32_t break_on_finish, int32_t* done_out, int32_t* steps_ran) {
  API_BEGIN(c)
  wl_ctx::Session& ss = c->sess;
  WL_CHECK(ss.open, WL_ERR_STATE, "wl_session_run: no open session");
  WL_CHECK(max_steps >= 1 && done_out, WL_ERR_ARG, "wl_session_run: bad arguments");
  const int cap = ss.cap;
  int ran = 0;
  c->last_ms[2] = 0.f;
  if (ss.live > 0) {
    SessScope scope(c);
    cudaStream_t st = c->st;
    const VocabIds vi = vocab_ids(c);
    ensure_host(c, (size_t)cap + 16, 16);
    int* h = c->h_int;
    h[0] = max_steps; h[1] = break_on_finish ? 1 : 0; h[2] = cap - ss.live;
    WL_CUDA(cudaMemcpyAsync(c->ds.steps_left, h, 4, cudaMemcpyHostToDevice, st));
    WL_CUDA(cudaMemcpyAsync(c->ds.brk, h + 1, 8, cudaMemcpyHostToDevice, st));
    WL_CUDA(cudaEventRecord(c->ev0, st));
    if (ss.use_graph) {
      long kernels = 0;
      cudaGraphExec_t exec = decode_graph(c, "s", cap, ss.Kr, ss.K, ss.so, vi, ss.nsplit, true, &kernels);
      WL_CUDA(cudaGraphLaunch(exec, st));
      WL_CUDA(cudaMemcpyAsync(h + 4, c->ds.steps_left, 4, cudaMemcpyDeviceToHost, st));
      WL_CUDA(cudaMemcpyAsync(h + 8, c->ds.done, (size_t)cap * 4, cudaMemcpyDeviceToHost, st));
      WL_CUDA(cudaEventRecord(c->ev1, st));
      WL_CUDA(cudaStreamSynchronize(st));
      ran = max_steps - h[4];
      c->graph_launched += kernels * (long)ran;
    } else {   // graph-less (profiling / bisecting): the same loop condition evaluated on the host after every step
      for (;;) {
        decode_step(c, cap, ss.Kr, ss.so, vi, ss.nsplit, false);
        ++ran;
        WL_CUDA(cudaMemcpyAsync(h + 5, c->ds.n_done, 4, cudaMemcpyDeviceToHost, st));
        WL_CUDA(cudaStreamSynchronize(st));
        if (ran >= max_steps || h[5] >= cap || (break_on_finish && h[5] > h[2])) break;
      }
      WL_CUDA(cudaMemcpyAsync(h + 8, c->ds.done, (size_t)cap * 4, cudaMemcpyDeviceToHost, st));
      WL_CUDA(cudaEventRecord(c->ev1, st));
      WL_CUDA(cudaStreamSynchronize(st));
    }
    WL_CUDA(cudaEventElapsedTime(&c->last_ms[2], c->ev0, c->ev1));
    for (int b = 0; b < cap; ++b)
      if (ss.used[b] && !ss.finished[b] && h[8 + b]) { ss.finished[b] = 1; ss.live -= 1; }
    ss.steps += ran;
    ss.runs += 1;
  }
  for (int b = 0; b < cap; ++b) done_out[b] = (ss.used[b] && ss.finished[b]) ? 1 : 0;
  if (steps_ran) *steps_ran = ran;
  API_END(c)
}

extern "C" int wl_session_collect(wl_ctx* c, int32_t index, int32_t* out_ids, int32_t* out_len, float* out_score, float* out_no_speech,
                                  int32_t* out_steps) {
  API_BEGIN(c)
  wl_ctx::Session& ss = c->sess;
  WL_CHECK(ss.open, WL_ERR_STATE, "wl_session_collect: no open session");
  WL_CHECK(index >= 0 && index < ss.cap && out_ids && out_len && out_score, WL_ERR_ARG, "wl_session_collect: bad arguments");
  WL_CHECK(ss.used[index] && ss.finished[index], WL_ERR_STATE, "wl_session_collect: stream index %d has not finished", index);
  const DecodeState& s = ss.ds;
  cudaStream_t st = c->st;
  ensure_host(c, (size_t)MAX_HYPS * (T_MAX + 2) + 16, MAX_HYPS + 2);
  int* h_cnt = c->h_int;
  int* h_steps = h_cnt + 1;
  int* h_len = h_cnt + 8;
  int* h_tok = h_len + MAX_HYPS;
  float* h_cum = c->h_flt;
  float* h_ns = h_cum + MAX_HYPS;
  WL_CUDA(cudaMemcpyAsync(h_cnt, s.hyp_count + index, 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_steps, s.steps_run + index, 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_len, s.hyp_len + (size_t)index * MAX_HYPS, MAX_HYPS * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_tok, s.hyp_tok + (size_t)index * MAX_HYPS * T_MAX, (size_t)MAX_HYPS * T_MAX * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_cum, s.hyp_cum + (size_t)index * MAX_HYPS, MAX_HYPS * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(h_ns, s.no_speech + index, 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaStreamSynchronize(st));
  emit_hyps(ss.NH, ss.length_penalty, h_cnt[0], h_len, h_cum, h_tok, out_ids, out_len, out_score);
  if (out_no_speech) *out_no_speech = h_ns[0];
  if (out_steps) *out_steps = h_steps[0];
  ss.used[index] = 0;
  ss.finished[index] = 0;
  API_END(c)
}

extern "C" int wl_session_close(wl_ctx* c) {
  API_BEGIN(c)
  wl_ctx::Session& ss = c->sess;
  if (ss.open) {   // streams still in flight are dropped: their indices go idle again
    const int cap = ss.cap;
    std::vector<int> ones(cap, 1);
    const int nd = cap;
    WL_CUDA(cudaMemcpy(ss.ds.done, ones.data(), (size_t)cap * 4, cudaMemcpyHostToDevice));
    WL_CUDA(cudaMemset(ss.ds.active, 0, (size_t)cap * ss.Kr * 4));
    WL_CUDA(cudaMemcpy(ss.ds.n_done, &nd, 4, cudaMemcpyHostToDevice));
    ss.used.assign(cap, 0);
    ss.finished.assign(cap, 0);
    ss.live = 0;
    ss.open = false;
  }
  API_END(c)
}

// teacher-forced driver shared by wl_decode_logits / wl_detect_language / wl_align
static void forced_run(wl_ctx* c, const int32_t* slots, int B, const int32_t* tokens, const int32_t* off, bool align_mode,
                       float* logits_out_dev /* [sumT][V] or null */, const std::vector<long>& row_base) {
  SearchOpts so;
  memset(&so, 0, sizeof(so));
  so.beam = 1; so.rows_per_stream = 1; so.max_cand = 1; so.suppress_mask = c->suppress_mask;
  const VocabIds vi = vocab_ids(c);
  cudaStream_t st = c->st;
  const int max_steps = upload_streams(c, slots, B, tokens, off, T_MAX, true);
  decode_init(st, c->ds, so, vi, B, B);
  const int nsplit = align_mode ? 1 : cross_attn_pick_nsplit(B, c->H, c->num_sms, 1);
  for (int i = 0; i < max_steps; ++i) {
    decode_step(c, B, 1, so, vi, nsplit, align_mode);
    if (logits_out_dev)
      for (int b = 0; b < B; ++b)
        if (i < off[b + 1] - off[b])
          WL_CUDA(cudaMemcpyAsync(logits_out_dev + (row_base[b] + i) * c->V, c->logits + (long)b * c->Vld, (size_t)c->V * 4,
                                  cudaMemcpyDeviceToDevice, st));
  }
  WL_CUDA(cudaStreamSynchronize(st));
}

extern "C" int wl_decode_logits(wl_ctx* c, const int32_t* slots, int32_t B, const int32_t* tokens, const int32_t* tok_off,
                                float* logits_out) {
  API_BEGIN(c)
  WL_CHECK(c->finalized && slots && tokens && tok_off && logits_out && B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_decode_logits: bad arguments");
  std::vector<long> base(B);
  long tot = 0;
  for (int b = 0; b < B; ++b) { base[b] = tot; tot += tok_off[b + 1] - tok_off[b]; }
  float* dev = nullptr;
  WL_CUDA(cudaMalloc((void**)&dev, (size_t)tot * c->V * 4));
  try {
    forced_run(c, slots, B, tokens, tok_off, false, dev, base);
    WL_CUDA(cudaMemcpy(logits_out, dev, (size_t)tot * c->V * 4, cudaMemcpyDeviceToHost));
  } catch (...) {
    cudaFree(dev);
    throw;
  }
  cudaFree(dev);
  API_END(c)
}

extern "C" int wl_detect_language(wl_ctx* c, const int32_t* slots, int32_t B, float* probs) {
  API_BEGIN(c)
  WL_CHECK(c->finalized && slots && probs && B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_detect_language: bad arguments");
  WL_CHECK(c->cfg.n_lang > 0, WL_ERR_STATE, "detect_language can only be called on multilingual models");
  std::vector<int32_t> toks(B, c->cfg.sot), off(B + 1);
  std::vector<long> base(B);
  for (int b = 0; b <= B; ++b) off[b] = b;
  forced_run(c, slots, B, toks.data(), off.data(), false, nullptr, base);
  const int nl = c->cfg.n_lang;
  std::vector<float> lg((size_t)B * nl);
  for (int b = 0; b < B; ++b)
    WL_CUDA(cudaMemcpy(lg.data() + (size_t)b * nl, c->logits + (long)b * c->Vld + c->cfg.lang_begin, nl * 4, cudaMemcpyDeviceToHost));
  for (int b = 0; b < B; ++b) {
    float mx = -INFINITY;
    for (int i = 0; i < nl; ++i) mx = std::max(mx, lg[(size_t)b * nl + i]);
    double sum = 0;
    for (int i = 0; i < nl; ++i) sum += exp((double)lg[(size_t)b * nl + i] - mx);
    for (int i = 0; i < nl; ++i) probs[(size_t)b * nl + i] = (float)(exp((double)lg[(size_t)b * nl + i] - mx) / sum);
  }
  API_END(c)
}

// ------------------------------------------------------------------------------------------ K14 align
extern "C" int wl_align(wl_ctx* c, const int32_t* slots, int32_t B, const int32_t* start_seq, int32_t n_start, const int32_t* text,
                        const int32_t* text_off, const int32_t* num_frames, int32_t median_width, int32_t* pairs_out,
                        int32_t cap_pairs, int32_t* pair_off, float* tok_probs) {
  API_BEGIN(c)
  WL_CHECK(c->finalized && slots && start_seq && text && text_off && num_frames && pairs_out && pair_off && tok_probs, WL_ERR_ARG,
           "wl_align: null argument");
  WL_CHECK(B >= 1 && B <= c->Bm, WL_ERR_ARG, "wl_align: B=%d exceeds max_streams", B);
  const int nh = (int)c->align_heads.size() / 2;
  WL_CHECK(nh > 0, WL_ERR_STATE, "wl_align: no alignment heads configured");
  std::vector<int32_t> toks, off(B + 1);
  std::vector<long> base(B, 0);
  int maxT = 0;
  off[0] = 0;
  for (int b = 0; b < B; ++b) {
    for (int i = 0; i < n_start; ++i) toks.push_back(start_seq[i]);
    toks.push_back(c->cfg.no_timestamps);
    for (int i = text_off[b]; i < text_off[b + 1]; ++i) toks.push_back(text[i]);
    toks.push_back(c->cfg.eot);
    off[b + 1] = (int)toks.size();
    maxT = std::max(maxT, off[b + 1] - off[b]);
  }
  WL_CHECK(maxT <= T_MAX, WL_ERR_ARG, "wl_align: sequence of %d tokens exceeds %d", maxT, T_MAX);
  const long need_buf = (long)B * nh * T_MAX * S_ENC;
  if (need_buf > c->align_buf_cap) {
    c->align_buf = dalloc<float>(c, need_buf, false);
    c->align_buf_cap = need_buf;
  }
  for (int b = 0; b < B; ++b)
    WL_CHECK(slots[b] >= 0 && slots[b] < c->NS && c->slot_used[slots[b]], WL_ERR_ARG, "wl_align: stream %d: bad encoder slot %d", b, slots[b]);
  // ---- teacher-forced pass: every position of every stream at once (the K8 machinery), attention probabilities of
  // the alignment heads captured on the way
  std::vector<int> ntok(B);
  for (int b = 0; b < B; ++b) ntok[b] = off[b + 1] - off[b];
  const PfRows r = pf_rows(B, 1, toks.data(), off.data(), ntok.data(), slots);
  pf_stack(c, r, true);
  wl_ctx::Prefill& f = c->pf;
  cudaStream_t st = c->st;
  // ---- P(text token | prefix): the logits row that predicts it
  const int n_text = text_off[B] - text_off[0];
  if (n_text > 0) {
    if (n_text > f.tokp_cap) { f.tokp = dalloc<float>(c, (size_t)n_text + 256); f.tokp_cap = n_text + 256; }
    std::vector<int> sel, tgt, oidx;
    for (int b = 0; b < B; ++b)
      for (int i = 0; i < text_off[b + 1] - text_off[b]; ++i) {
        sel.push_back(r.rowbase[b] + n_start + i);
        tgt.push_back(toks[off[b] + n_start + 1 + i]);
        oidx.push_back(text_off[b] - text_off[0] + i);
      }
    pf_row_probs(c, sel, tgt, oidx, f.tokp);
    WL_CUDA(cudaMemcpy(tok_probs + text_off[0], f.tokp, (size_t)n_text * 4, cudaMemcpyDeviceToHost));
  }
  // ---- standardise / median filter / mean over heads / DTW on the device
  if (!f.mat) {
    f.mat = dalloc<float>(c, (size_t)c->Bm * T_MAX * S_ENC, false);
    f.aT = dalloc<int>(c, c->Bm); f.anf = dalloc<int>(c, c->Bm);
    f.path = dalloc<int>(c, (size_t)c->Bm * (T_MAX + S_ENC + 2) * 2, false);
    f.path_len = dalloc<int>(c, c->Bm);
  }
  const int path_cap = T_MAX + S_ENC + 2;
  std::vector<int> hT(B), hnf(B);
  for (int b = 0; b < B; ++b) {
    hT[b] = ntok[b];
    hnf[b] = std::max(1, std::min(num_frames[b] / 2, (int)S_ENC));
  }
  WL_CUDA(cudaMemcpyAsync(f.aT, hT.data(), B * 4, cudaMemcpyHostToDevice, st));
  WL_CUDA(cudaMemcpyAsync(f.anf, hnf.data(), B * 4, cudaMemcpyHostToDevice, st));
  align_postprocess(st, c->align_buf, f.mat, f.aT, f.anf, B, nh, median_width, n_start, maxT, f.path, path_cap, f.path_len);
  std::vector<int> hlen(B), hpath((size_t)B * path_cap * 2);
  WL_CUDA(cudaMemcpyAsync(hlen.data(), f.path_len, B * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaMemcpyAsync(hpath.data(), f.path, hpath.size() * 4, cudaMemcpyDeviceToHost, st));
  WL_CUDA(cudaStreamSynchronize(st));
  int np_total = 0;
  pair_off[0] = 0;
  for (int b = 0; b < B; ++b) {
    const int len = hlen[b];
    WL_CHECK(np_total + len <= cap_pairs, WL_ERR_ARG, "wl_align: pairs_out capacity %d too small", cap_pairs);
    const int* pp = hpath.data() + (size_t)b * path_cap * 2;
    for (int k = len - 1; k >= 0; --k) {   // the device wrote the path end -> start
      pairs_out[2 * np_total] = pp[2 * k];
      pairs_out[2 * np_total + 1] = pp[2 * k + 1];
      ++np_total;
    }
    pair_off[b + 1] = np_total;
  }
  API_END(c)
}

// ------------------------------------------------------------------------------------------ test hook
extern "C" int wl_test_gemm(wl_ctx* c, const uint16_t* a_f16, const uint16_t* b_f16, const float* bias, float* cc, int32_t M,
                            int32_t N, int32_t K, int32_t batch, int32_t transposed_store, int32_t gelu, int32_t use_simt) {
  API_BEGIN(c)
  __half *da = nullptr, *db = nullptr;
  float *dbias = nullptr, *dc = nullptr;
  const size_t na = (size_t)batch * M * K, nb = (size_t)batch * N * K, nc = (size_t)batch * M * N;
  WL_CUDA(cudaMalloc((void**)&da, na * 2));
  WL_CUDA(cudaMalloc((void**)&db, nb * 2));
  WL_CUDA(cudaMalloc((void**)&dc, nc * 4));
  WL_CUDA(cudaMemcpy(da, a_f16, na * 2, cudaMemcpyHostToDevice));
  WL_CUDA(cudaMemcpy(db, b_f16, nb * 2, cudaMemcpyHostToDevice));
  WL_CUDA(cudaMemset(dc, 0, nc * 4));
  if (bias) {
    WL_CUDA(cudaMalloc((void**)&dbias, (size_t)std::max(M, N) * 4));
    WL_CUDA(cudaMemcpy(dbias, bias, (size_t)(transposed_store ? M : N) * 4, cudaMemcpyHostToDevice));
  }
  // the copies / memset above ran on the legacy default stream, the GEMM runs on the library's non-blocking stream:
  // without this the kernel may overtake the memset of its own output buffer
  WL_CUDA(cudaDeviceSynchronize());
  GemmEpilogue e;
  e.out = dc; e.out_f32 = 1; e.gelu = gelu; e.bias = dbias;
  if (transposed_store) { e.ldm = 1; e.ldn = M; e.bias_on_m = 1; }   // C^T stored: [N][M]
  else { e.ldm = N; e.ldn = 1; }
  e.ob1 = (long)M * N;
  try {
    GemmOperand A = opnd(da, M, K, K, batch, (long)M * K), Bo = opnd(db, N, K, K, batch, (long)N * K);
    if (use_simt) gemm_tn_simt(c->st, A, Bo, M, N, K, e);
    else gemm_tn(c->st, A, Bo, M, N, K, e);
    WL_CUDA(cudaStreamSynchronize(c->st));
    WL_CUDA(cudaMemcpy(cc, dc, nc * 4, cudaMemcpyDeviceToHost));
  } catch (...) {
    cudaFree(da); cudaFree(db); cudaFree(dc); if (dbias) cudaFree(dbias);
    throw;
  }
  cudaFree(da); cudaFree(db); cudaFree(dc); if (dbias) cudaFree(dbias);
  API_END(c)
}

// mode 0/1/2/3 of wgemm (see gemm.cuh); out holds [R][n_out] floats (mode 1: the residual on input, the sum on output;
// mode 2: gelu as fp32; mode 3: the K ranges summed on the host side of this hook)
extern "C" int wl_test_wgemm(wl_ctx* c, const uint16_t* w_f16, const uint16_t* x_f16, const float* bias, float* out, int32_t R,
                             int32_t n_out, int32_t K, int32_t mode) {
  API_BEGIN(c)
  const bool clustered = (mode & 8) != 0;   // modes 8, 9, 10: the cluster split-K GEMM (cgemm) with epilogue 0, 1, 2
  mode &= 7;
  WL_CHECK(w_f16 && x_f16 && out && mode >= 0 && mode <= (clustered ? 2 : 3), WL_ERR_ARG, "wl_test_wgemm: bad arguments");
  WL_CHECK(clustered || wgemm_supported(R, K), WL_ERR_ARG, "wl_test_wgemm: unsupported shape R=%d K=%d", R, K);
  __half *dw = nullptr, *dx = nullptr, *dh = nullptr;
  float *db = nullptr, *dout = nullptr;
  const size_t nw = (size_t)n_out * K, nx = (size_t)R * K, no = (size_t)R * n_out;
  const int ks = clustered ? 1 : wgemm_ksplit(K);
  WL_CUDA(cudaMalloc((void**)&dw, nw * 2));
  WL_CUDA(cudaMalloc((void**)&dx, nx * 2));
  WL_CUDA(cudaMalloc((void**)&dh, no * 2));
  WL_CUDA(cudaMalloc((void**)&dout, no * 4 * (size_t)std::max(1, ks)));
  WL_CUDA(cudaMalloc((void**)&db, (size_t)n_out * 4));
  try {
    WL_CUDA(cudaMemcpy(dw, w_f16, nw * 2, cudaMemcpyHostToDevice));
    WL_CUDA(cudaMemcpy(dx, x_f16, nx * 2, cudaMemcpyHostToDevice));
    WL_CUDA(cudaMemset(dout, 0, no * 4 * (size_t)std::max(1, ks)));
    if (mode == 1) WL_CUDA(cudaMemcpy(dout, out, no * 4, cudaMemcpyHostToDevice));
    if (bias) WL_CUDA(cudaMemcpy(db, bias, (size_t)n_out * 4, cudaMemcpyHostToDevice));
    WL_CUDA(cudaDeviceSynchronize());
    if (clustered) cgemm(c->st, dw, n_out, K, dx, R, bias ? db : nullptr, mode, dout, dh);
    else wgemm(c->st, dw, n_out, K, dx, R, bias ? db : nullptr, mode, dout, dh, (long)no);
    WL_CUDA(cudaStreamSynchronize(c->st));
    if (mode == 2) {
      std::vector<__half> h(no);
      WL_CUDA(cudaMemcpy(h.data(), dh, no * 2, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < no; ++i) out[i] = __half2float(h[i]);
    } else if (mode == 3) {
      std::vector<float> h(no * ks);
      WL_CUDA(cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < no; ++i) { float a = 0.f; for (int q = 0; q < ks; ++q) a += h[(size_t)q * no + i]; out[i] = a; }
    } else {
      WL_CUDA(cudaMemcpy(out, dout, no * 4, cudaMemcpyDeviceToHost));
    }
  } catch (...) {
    cudaFree(dw); cudaFree(dx); cudaFree(dh); cudaFree(dout); cudaFree(db);
    throw;
  }
  cudaFree(dw); cudaFree(dx); cudaFree(dh); cudaFree(dout); cudaFree(db);
  API_END(c)
}

extern "C" int wl_bench_gemm(wl_ctx* c, int32_t M, int32_t N, int32_t K, int32_t batch, int32_t iters, int32_t flags,
                             float* ms_out) {
  // flags: 1 transposed (swap-AB) store, 2 bias, 4 GELU, 8 fp32 output with fp32 residual (in place)
  API_BEGIN(c)
  WL_CHECK(ms_out && M > 0 && N > 0 && K > 0 && batch > 0 && iters > 0, WL_ERR_ARG, "wl_bench_gemm: bad arguments");
  __half *da = nullptr, *db = nullptr;
  void* dc = nullptr;
  float* dbias = nullptr;
  const bool tr = flags & 1, f32 = flags & 8;
  const size_t na = (size_t)batch * M * K, nb = (size_t)batch * N * K, nc = (size_t)batch * M * N;
  WL_CUDA(cudaMalloc((void**)&da, na * 2));
  WL_CUDA(cudaMalloc((void**)&db, nb * 2));
  WL_CUDA(cudaMalloc(&dc, nc * (f32 ? 4 : 2)));
  WL_CUDA(cudaMalloc((void**)&dbias, (size_t)std::max(M, N) * 4));
  WL_CUDA(cudaMemset(da, 0x11, na * 2));
  WL_CUDA(cudaMemset(db, 0x11, nb * 2));
  WL_CUDA(cudaMemset(dc, 0, nc * (f32 ? 4 : 2)));
  WL_CUDA(cudaMemset(dbias, 0, (size_t)std::max(M, N) * 4));
  WL_CUDA(cudaDeviceSynchronize());
  GemmEpilogue e;
  e.out = dc; e.out_f32 = f32 ? 1 : 0;
  if (tr) { e.ldm = 1; e.ldn = M; } else { e.ldm = N; e.ldn = 1; }
  e.ob1 = (long)M * N;
  if (flags & 2) { e.bias = dbias; e.bias_on_m = tr ? 1 : 0; }
  if (flags & 4) e.gelu = 1;
  if (f32) { e.resid = (const float*)dc; e.rldm = e.ldm; e.rldn = e.ldn; e.rb1 = e.ob1; }
  try {
    GemmOperand A = opnd(da, M, K, K, batch, (long)M * K), Bo = opnd(db, N, K, K, batch, (long)N * K);
    for (int i = 0; i < 3; ++i) gemm_tn(c->st, A, Bo, M, N, K, e);
    WL_CUDA(cudaEventRecord(c->ev0, c->st));
    for (int i = 0; i < iters; ++i) gemm_tn(c->st, A, Bo, M, N, K, e);
    WL_CUDA(cudaEventRecord(c->ev1, c->st));
    WL_CUDA(cudaStreamSynchronize(c->st));
    float ms;
    WL_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_out = ms / iters;
  } catch (...) {
    cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dbias);
    throw;
  }
  cudaFree(da); cudaFree(db); cudaFree(dc); cudaFree(dbias);
  API_END(c)
}
