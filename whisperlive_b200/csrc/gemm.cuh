// tcgen05 / TMA / TMEM GEMM used by every dense contraction on the hot path (K2-K9, K12).
//   C[z][m][n] = epilogue( sum_k A[z][m][k] * B[z][n][k] )      fp16 in, fp32 accumulate
// Both operands are K-major ("TN"): activations [rows, K] and nn.Linear weights [out, K].
#pragma once
#include "common.cuh"

namespace wl {

// One operand as a strided 4-D view (k fastest): element (k, r, i1, i2) at ptr[k + r*ld + i1*s1 + i2*s2].
struct GemmOperand {
  const __half* ptr = nullptr;
  long rows = 0;      // M (for A) or N (for B)
  long k = 0;         // contraction length seen by TMA (reads beyond it are zero-filled)
  long ld = 0;        // elements between consecutive rows (may be < k: overlapping rows, conv-as-GEMM)
  int n1 = 1;         // batch dims; grid z = i1 + n1*i2
  long s1 = 0;
  int n2 = 1;
  long s2 = 0;
};

enum GemmPost : int { GEMM_POST_NONE = 0, GEMM_POST_LN = 1, GEMM_POST_GELU = 2 };

enum GemmMode : int {
  GEMM_STORE = 0,      // out[z][m*ldm + n*ldn]
  GEMM_HEADSPLIT = 1,  // cross-KV cache layout: row m=(b,s), col n=(h,dd) -> out[slot[b]][h][s][dd]
};

struct GemmEpilogue {
  void* out = nullptr;
  int out_f32 = 0;           // 1: float output, 0: __half
  long ldm = 0, ldn = 1;     // element strides of the output (ldm=1 -> transposed / swap-AB store)
  long ob1 = 0, ob2 = 0;     // output batch strides
  const float* bias = nullptr;
  int bias_on_m = 0;         // bias indexed by m (swap-AB) instead of n
  int gelu = 0;              // exact erf GELU after bias
  const float* resid = nullptr;  // fp32 residual added last; same indexing scheme as out
  long rldm = 0, rldn = 1, rb1 = 0, rb2 = 0;
  int partials = 0;          // >0: split K into `partials` ranges; range s stores its raw fp32 partial sum at
  long part_stride = 0;      //     out + s*part_stride (no bias); the consumer adds them in order (deterministic)
  // Fused consumer of a split-K result (partials > 0 only): after its partial stores every CTA passes a grid-wide
  // barrier (all CTAs are resident: the kernel is persistent) and the epilogue warps run the row-wise operation that
  // would otherwise be the next kernel.  GEMM_POST_LN: x[r] += bias + sum_s partial_s[r]; y[r] = LayerNorm(x[r]) (fp16).
  // GEMM_POST_GELU: y = gelu(bias + sum_s partial_s) (fp16).  Rows are the GEMM's n index (swap-AB), features its m.
  int post = 0;
  float* post_x = nullptr;
  const float* post_bias = nullptr;
  const float* post_g = nullptr;
  const float* post_b = nullptr;
  __half* post_y = nullptr;
  unsigned* post_bar = nullptr;   // {arrival count, generation}, zero-initialised, one pair per context
  int a_static = 0;          // A is a weight matrix: under programmatic dependent launch its first k-blocks are
                             //     fetched before waiting for the preceding kernel (which only produces B)
  int mode = GEMM_STORE;
  // GEMM_HEADSPLIT parameters
  int hs_S = 0, hs_H = 0;
  long hs_slot_stride = 0;
  const int* hs_slots = nullptr;  // device: slot index per stream b
};

// Launch on `stream`. M/N/K are the logical sizes per batch entry. Throws wl::Error.
void gemm_tn(cudaStream_t stream, const GemmOperand& A, const GemmOperand& B, int M, int N, int K,
             const GemmEpilogue& epi);

// Plain CUDA-core reference of the same contract (debug/bisect aid on the GPU box, WLB200_GEMM=simt;
// also what the GEMM unit test compares against on-device).
void gemm_tn_simt(cudaStream_t stream, const GemmOperand& A, const GemmOperand& B, int M, int N, int K,
                  const GemmEpilogue& epi);

long gemm_launch_count();
int gemm_split_plan(int M, int N, int K);

// Compact decode-step GEMM (dec_gemm.cu): out[s][r][ldn] (s < nsplit) = partial sums over K range s of
// W[n_out, K] x X[R, K]^T, fp32, no bias.  nsplit from dec_gemm_split_plan (1 for the vocabulary projection).
void dec_gemm(cudaStream_t st, const __half* W, int n_out, int K, const __half* X, int R, float* out, int ldn, long part_stride,
              int nsplit);
int dec_gemm_split_plan(int n_out, int R, int K, int max_split = 8);
long dec_gemm_launch_count();
void dec_gemm_prime();
void dec_gemm_tl_bind(unsigned long long* p);
// The same pipeline with the K split inside a thread-block cluster and the reduction through distributed shared memory:
// final values with the epilogue fused (mode 0: + bias; 1: out_f32 += acc + bias; 2: out_f16 = gelu(acc + bias)).
void cgemm(cudaStream_t st, const __half* W, int n_out, int K, const __half* X, int R, const float* bias, int mode, float* out_f32,
           __half* out_f16);
int cgemm_split_plan(int n_out, int R, int K);
long cgemm_launch_count();
void cgemm_prime();

// Small-batch decode GEMM with fused epilogue (wgemm.cu, R <= 32 rows, mma.sync + bulk-copied weight slices).
// mode 0: out_f32 = acc + bias; 1: out_f32 += acc + bias (in place); 2: out_f16 = gelu(acc + bias);
// 3: out_f32[ks] = raw partial sum of K range ks (only when K > 1280)
// prefetch_ptr / prefetch_bytes: the weights of the next linear layer, requested into L2 by this launch (optional)
void wgemm(cudaStream_t st, const __half* W, int n_out, int K, const __half* X, int R, const float* bias, int mode, float* out_f32,
           __half* out_f16, long part_stride, const void* prefetch_ptr = nullptr, long prefetch_bytes = 0);
bool wgemm_supported(int R, int K);
int wgemm_ksplit(int K);
long wgemm_launch_count();
void wgemm_prime();
void wgemm_tl_bind(unsigned long long* p);

// Tensor map over an operand view (dims sorted by stride) + the coordinate slots of (row, i1, i2).
struct TmapInfo {
  CUtensorMap tm;
  int pos[3];
};
TmapInfo make_tmap(const GemmOperand& op, int box_rows, int box_k = 64);

}  // namespace wl
