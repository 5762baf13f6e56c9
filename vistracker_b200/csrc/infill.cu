// HVOP-Net: the conditional motion in-filler that predicts the object rotation of occluded frames from the
// SMPL-T motion and the visible object rotations (model/infill/mfiller_cond.py:17-104, model/transformers/former_deci.py:31-175,
// model/transformers/posi_embed.py:35-66) and its autoregressive clip loop (interp/test_infill_autoreg.py:34-174,
// interp/test_cinfill_autoreg.py:32-51).
//
// A clip is 180 tokens and the loop is serial (clip i+1 is seeded by the prediction of clip i), so the stage is latency-bound: the design
// goal is few launches, no host round trip per clip, and fp32 FFMA math (the matrices are 32..160 wide -- far below one tensor-core tile's
// worth of work per token block).  One encoder layer is two launches:
//
//   vt_infill_attn   softmax(q k^T + key padding mask) v for (8 queries, head, clip) per CTA; K of the head staged in shared memory with an
//                    odd pitch (one key per lane), V read coalesced from L2
//   vt_infill_tail   the per-token rest of a layer AND the per-token start of the next one: out-projection + residual, LayerNorm 2,
//                    Linear -> activation -> Linear + residual, then (next layer) LayerNorm 1, q = k = (h + pos) Wq|Wk, v = h Wv
//   vt_infill_head   the start of an encoder's first layer, with the input feature projection fused in front
//   vt_infill_mlp    the predictor (Linear + LeakyReLU chain)
//
// plus the clip gather / commit of the autoregressive loop (vt_infill_pack_clip, vt_infill_commit_clip), which keep the trajectory and the
// running prediction in device memory.  Every layer in the reference is built pre-norm (former_deci.py:139-143 passes pre_norm=True to the
// layer whatever the option says); the option only decides whether a final LayerNorm exists (handled as `final_ln`).
//
// Weights are fp32, k-major ([in][out], so that consecutive threads read consecutive addresses), one pack per layer:
//   ln1.w ln1.b | Wqkv^T [D][3D] bqkv [3D] | Wo^T [D][D] bo [D] | ln2.w ln2.b | W1^T [D][F] b1 [F] | W2^T [F][D] b2 [D]
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int IF_TB = 4;            // tokens per CTA in the per-token kernels
constexpr int IF_THREADS = 256;
constexpr int IF_KU = 32;           // weight loads in flight per thread
constexpr int IF_QB = 8;            // queries per CTA in the attention kernel (one per warp)
constexpr int IF_MAX_T = 256;       // keys per clip (8 per lane)
constexpr int IF_MAX_D = 256;
constexpr int IF_MAX_F = 1024;
constexpr float IF_LN_EPS = 1e-5f;  // nn.LayerNorm default

struct IfLayer {                    // offsets (floats) into a layer pack
  int ln1, wqkv, bqkv, wo, bo, ln2, w1, b1, w2, b2, total;
};
__host__ __device__ inline int if_pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ inline IfLayer if_layer(int D, int F) {
  IfLayer l;
  int o = 0;
  l.ln1 = o; o += 2 * D;
  l.wqkv = o; o += D * 3 * D;
  l.bqkv = o; o += 3 * D;
  l.wo = o; o += D * D;
  l.bo = o; o += D;
  l.ln2 = o; o += 2 * D;
  l.w1 = o; o += D * F;
  l.b1 = o; o += F;
  l.w2 = o; o += F * D;
  l.b2 = o; o += D;
  l.total = o;
  return l;
}

__device__ __forceinline__ float if_act(float x, int act) {
  if (act == 0) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));      // F.gelu (exact)
  if (act == 1) return fmaxf(x, 0.f);                                              // F.relu
  return x > 0.f ? x : 0.01f * x;                                                  // F.leaky_relu / nn.LeakyReLU()
}

// out(o, acc[t]) for o in [0, N): acc[t] = bias[o] + sum_k xs[t][k] * Wt[k][o] (row pitch ldw); xs in shared memory (broadcast reads), Wt from L2
template <typename Epi>
__device__ __forceinline__ void if_dense(const float* xs, int ldx, int K, const float* __restrict__ Wt, int ldw, const float* __restrict__ bias, int N, Epi epi) {
  for (int o = threadIdx.x; o < N; o += IF_THREADS) {
    float acc[IF_TB];
    const float b = bias ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int t = 0; t < IF_TB; ++t) acc[t] = b;
    const float* w = Wt + o;
    // the stage is latency-bound (a handful of CTAs, every one streaming the whole matrix out of L2): keep IF_KU loads in flight per thread;
    // activations are read four k at a time (rows are 16-byte aligned and zero-padded to a multiple of 4)
    for (int k = 0; k < K; k += IF_KU) {
      float wv[IF_KU];
#pragma unroll
      for (int u = 0; u < IF_KU; ++u) wv[u] = k + u < K ? __ldg(w + (size_t)(k + u) * ldw) : 0.f;
#pragma unroll
      for (int u = 0; u < IF_KU; u += 4) {
        if (k + u < K) {
#pragma unroll
          for (int t = 0; t < IF_TB; ++t) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + t * ldx + k + u);
            acc[t] = fmaf(xv.x, wv[u], acc[t]);
            acc[t] = fmaf(xv.y, wv[u + 1], acc[t]);
            acc[t] = fmaf(xv.z, wv[u + 2], acc[t]);
            acc[t] = fmaf(xv.w, wv[u + 3], acc[t]);
          }
        }
      }
    }
    epi(o, acc);
  }
}

// LayerNorm of the IF_TB rows of xs into hs (and hs + pos into hp when pos != nullptr); warp t handles row t
__device__ __forceinline__ void if_layernorm(const float* xs, int ld, int D, const float* __restrict__ wb, float* hs, float* hp, const float* __restrict__ pos,
                                             int tok0, int n_tok, int T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < IF_TB) {
    const float* x = xs + warp * ld;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)D;
    float v = 0.f;
    for (int c = lane; c < D; c += 32) { const float d = x[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)D + IF_LN_EPS);
    const int tok = min(tok0 + warp, n_tok - 1);
    const float* p = pos ? pos + (size_t)(tok % T) * D : nullptr;
    for (int c = lane; c < D; c += 32) {
      const float h = (x[c] - mean) * rstd * __ldg(wb + c) + __ldg(wb + D + c);
      hs[warp * ld + c] = h;
      if (hp) hp[warp * ld + c] = h + (p ? __ldg(p + c) : 0.f);
    }
  }
}

struct IfTokenArgs {
  // optional input projection (first layer of an encoder): x = in P + b
  const float* in; int in_ld, in_dim; const float* proj;      // proj: [in_dim][D] | bias [D]
  // residual stream
  float* x; int x_ld;                                          // [n_tok][x_ld], D columns
  float* y; int y_ld;                                          // where the finished layer's stream is written (may be x)
  const float* attn;                                           // [n_tok][D] (tail only)
  const float* tail;                                           // pack of the layer to finish, or nullptr
  const float* head;                                           // pack of the layer to start, or nullptr
  const float* final_ln;                                       // [2D] LayerNorm applied to the stream before it is written to y, or nullptr
  const float* pos;                                            // [T][D]
  float* qkv;                                                  // [n_tok][3D]
  int n_tok, T, D, F, Fh, act, heads;                          // F: feed-forward width of the tail layer, Fh: of the head layer
};

__global__ void __launch_bounds__(IF_THREADS) infill_token_kernel(IfTokenArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int D = a.D, ld = if_pad4(a.D);
  float* xs = sm;                       // [TB][ld] residual stream
  float* hs = xs + IF_TB * ld;          // [TB][ld] LayerNorm output / attention rows
  float* hp = hs + IF_TB * ld;          // [TB][ld] LayerNorm output + positional embedding
  float* fs = hp + IF_TB * ld;          // [TB][max(F, in_dim) + 1]
  const int ldf = if_pad4(max(max(a.F, a.in_dim), 1));
  const int tok0 = blockIdx.x * IF_TB;
  const int tid = threadIdx.x;
  for (int i = tid; i < IF_TB * (3 * ld + ldf); i += IF_THREADS) sm[i] = 0.f;      // the pad columns must hold finite values
  __syncthreads();

  if (a.in) {                            // feature projection (mfiller_cond.py:91, 93)
    for (int i = tid; i < IF_TB * a.in_dim; i += IF_THREADS) {
      const int t = i / a.in_dim, c = i % a.in_dim;
      fs[t * ldf + c] = a.in[(size_t)min(tok0 + t, a.n_tok - 1) * a.in_ld + c];
    }
    __syncthreads();
    if_dense(fs, ldf, a.in_dim, a.proj, D, a.proj + (size_t)a.in_dim * D, D, [&](int o, const float (&acc)[IF_TB]) {
#pragma unroll
      for (int t = 0; t < IF_TB; ++t) xs[t * ld + o] = acc[t];
    });
  } else {
    for (int i = tid; i < IF_TB * D; i += IF_THREADS) {
      const int t = i / D, c = i % D;
      xs[t * ld + c] = a.x[(size_t)min(tok0 + t, a.n_tok - 1) * a.x_ld + c];
    }
  }
  if (a.tail) {                          // forward_pre (former_deci.py:78-93) after the attention
    const IfLayer L = if_layer(D, a.F);
    for (int i = tid; i < IF_TB * D; i += IF_THREADS) {
      const int t = i / D, c = i % D;
      hs[t * ld + c] = a.attn[(size_t)min(tok0 + t, a.n_tok - 1) * D + c];
    }
    __syncthreads();
    if_dense(hs, ld, D, a.tail + L.wo, D, a.tail + L.bo, D, [&](int o, const float (&acc)[IF_TB]) {       // src = src + out_proj(attn)
#pragma unroll
      for (int t = 0; t < IF_TB; ++t) xs[t * ld + o] += acc[t];
    });
    __syncthreads();
    if_layernorm(xs, ld, D, a.tail + L.ln2, hs, nullptr, nullptr, tok0, a.n_tok, a.T);                  // src2 = norm2(src)
    __syncthreads();
    if_dense(hs, ld, D, a.tail + L.w1, a.F, a.tail + L.b1, a.F, [&](int o, const float (&acc)[IF_TB]) {      // activation(linear1(src2))
#pragma unroll
      for (int t = 0; t < IF_TB; ++t) fs[t * ldf + o] = if_act(acc[t], a.act);
    });
    __syncthreads();
    if_dense(fs, ldf, a.F, a.tail + L.w2, D, a.tail + L.b2, D, [&](int o, const float (&acc)[IF_TB]) {     // src = src + linear2(.)
#pragma unroll
      for (int t = 0; t < IF_TB; ++t) xs[t * ld + o] += acc[t];
    });
  }
  __syncthreads();
  if (a.final_ln) {                      // TransformerEncoder.norm (former_deci.py:126-127)
    if_layernorm(xs, ld, D, a.final_ln, hs, nullptr, nullptr, tok0, a.n_tok, a.T);
    __syncthreads();
    for (int i = tid; i < IF_TB * D; i += IF_THREADS) { const int t = i / D, c = i % D; xs[t * ld + c] = hs[t * ld + c]; }
    __syncthreads();
  }
  if (a.y && (a.tail || a.in || a.final_ln)) {
    for (int i = tid; i < IF_TB * D; i += IF_THREADS) {
      const int t = i / D, c = i % D;
      if (tok0 + t < a.n_tok) a.y[(size_t)(tok0 + t) * a.y_ld + c] = xs[t * ld + c];
    }
  }
  if (a.head) {                          // forward_pre (former_deci.py:83-89) up to the attention
    const IfLayer L = if_layer(D, a.Fh);
    if_layernorm(xs, ld, D, a.head + L.ln1, hs, hp, a.pos, tok0, a.n_tok, a.T);                         // src2 = norm1(src); q = k = src2 + pos
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)(D / a.heads));
    const float* W = a.head + L.wqkv;
    const float* bq = a.head + L.bqkv;
    if_dense(hp, ld, D, W, 3 * D, bq, 2 * D, [&](int o, const float (&acc)[IF_TB]) {                           // q (scaled), k: columns [0, 2D) of Wqkv^T
#pragma unroll
      for (int t = 0; t < IF_TB; ++t)
        if (tok0 + t < a.n_tok) a.qkv[(size_t)(tok0 + t) * 3 * D + o] = o < D ? acc[t] * scale : acc[t];
    });
    if_dense(hs, ld, D, W + 2 * D, 3 * D, bq + 2 * D, D, [&](int o, const float (&acc)[IF_TB]) {               // v: columns [2D, 3D), no pos
#pragma unroll
      for (int t = 0; t < IF_TB; ++t)
        if (tok0 + t < a.n_tok) a.qkv[(size_t)(tok0 + t) * 3 * D + 2 * D + o] = acc[t];
    });
  }
}

// grid (ceil(T / IF_QB), heads, clips); nn.MultiheadAttention core (former_deci.py:84-88) with key_padding_mask (True = ignored key)
__device__ __forceinline__ void if_stage_head(float* dst, int ldk, const float* __restrict__ src, int T, int dh, int pitch) {
  if ((dh & 3) == 0 && (pitch & 3) == 0 && (reinterpret_cast<size_t>(src) & 15) == 0) {
    const int q4 = dh >> 2;
    for (int i = threadIdx.x; i < T * q4; i += IF_THREADS) {
      const int j = i / q4, c = (i - j * q4) * 4;
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)j * pitch + c));
      float* d = dst + j * ldk + c;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = threadIdx.x; i < T * dh; i += IF_THREADS) {
      const int j = i / dh, c = i - j * dh;
      dst[j * ldk + c] = __ldg(src + (size_t)j * pitch + c);
    }
  }
}

__global__ void __launch_bounds__(IF_THREADS) infill_attn_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ key_mask, int T, int D, int heads,
                                                                 float* __restrict__ attn) {
  extern __shared__ __align__(16) float sm[];
  const int dh = D / heads, ldk = dh | 1;                // odd pitch: lane j reads row j conflict-free
  float* Ks = sm;                                        // [T][ldk]: the head's keys, then its values
  float* qs = Ks + (size_t)T * ldk;                      // [QB][dh]
  float* ps = qs + IF_QB * dh;                           // [QB][T]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * IF_QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* base = qkv + (size_t)b * T * 3 * D;
  if_stage_head(Ks, ldk, base + D + h * dh, T, dh, 3 * D);
  for (int i = threadIdx.x; i < IF_QB * dh; i += IF_THREADS) {
    const int qi = i / dh, c = i % dh;
    qs[i] = base[(size_t)min(q0 + qi, T - 1) * 3 * D + h * dh + c];
  }
  __syncthreads();
  const int q = min(q0 + warp, T - 1);                   // warps past the end repeat the last query and do not store
  float s[IF_MAX_T / 32];
#pragma unroll
  for (int i = 0; i < IF_MAX_T / 32; ++i) s[i] = 0.f;
  for (int c = 0; c < dh; ++c) {
    const float qc = qs[warp * dh + c];
#pragma unroll
    for (int i = 0; i < IF_MAX_T / 32; ++i) {
      const int j = lane + 32 * i;
      if (j < T) s[i] = fmaf(qc, Ks[j * ldk + c], s[i]);
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < IF_MAX_T / 32; ++i) {
    const int j = lane + 32 * i;
    if (j >= T || (key_mask && key_mask[(size_t)b * T + j])) s[i] = -INFINITY;
    m = fmaxf(m, s[i]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < IF_MAX_T / 32; ++i) {
    const int j = lane + 32 * i;
    s[i] = j < T ? expf(s[i] - m) : 0.f;                 // every key masked: exp(-inf + inf) = NaN, as torch's softmax
    sum += s[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < IF_MAX_T / 32; ++i) {
    const int j = lane + 32 * i;
    if (j < T) ps[warp * T + j] = s[i] * inv;
  }
  __syncthreads();                                       // every warp is done with the keys
  if_stage_head(Ks, ldk, base + 2 * D + h * dh, T, dh, 3 * D);
  __syncthreads();
  for (int c0 = 0; c0 < dh; c0 += 128) {
    float o4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < T; ++j) {
      const float p = ps[warp * T + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + lane + 32 * i;
        if (c < dh) o4[i] = fmaf(p, Ks[j * ldk + c], o4[i]);
      }
    }
    if (q0 + warp < T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + lane + 32 * i;
        if (c < dh) attn[((size_t)b * T + q) * D + h * dh + c] = o4[i];
      }
    }
  }
}

// predictor (mfiller_cond.py:57-73): Linear -> LeakyReLU -> ... -> Linear; pack = for each layer W^T [in][out] | bias [out]
struct IfMlpArgs { const float* x; int x_ld; int n_tok; int n_layers; int dims[6]; const float* pack; float* out; int out_ld; };

__global__ void __launch_bounds__(IF_THREADS) infill_mlp_kernel(IfMlpArgs a) {
  extern __shared__ __align__(16) float sm[];
  int wmax = 0;
  for (int i = 0; i <= a.n_layers; ++i) wmax = max(wmax, a.dims[i]);
  const int ld = if_pad4(wmax);
  float* cur = sm;
  float* nxt = sm + IF_TB * ld;
  const int tok0 = blockIdx.x * IF_TB;
  for (int i = threadIdx.x; i < 2 * IF_TB * ld; i += IF_THREADS) sm[i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < IF_TB * a.dims[0]; i += IF_THREADS) {
    const int t = i / a.dims[0], c = i % a.dims[0];
    cur[t * ld + c] = a.x[(size_t)min(tok0 + t, a.n_tok - 1) * a.x_ld + c];
  }
  __syncthreads();
  const float* p = a.pack;
  for (int l = 0; l < a.n_layers; ++l) {
    const int K = a.dims[l], N = a.dims[l + 1];
    const bool last = l == a.n_layers - 1;
    if_dense(cur, ld, K, p, N, p + (size_t)K * N, N, [&](int o, const float (&acc)[IF_TB]) {
#pragma unroll
      for (int t = 0; t < IF_TB; ++t) {
        if (!last) nxt[t * ld + o] = if_act(acc[t], 2);
        else if (tok0 + t < a.n_tok) a.out[(size_t)(tok0 + t) * a.out_ld + o] = acc[t];
      }
    });
    __syncthreads();
    p += (size_t)K * N + N;
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
}

// one clip of the autoregressive loop (test_infill_autoreg.py:93-105, 116-153; test_cinfill_autoreg.py:43-49 with obj_dim 6):
// data_smpl[t] = [rot6d_smpl | trans_smpl][start + t]; data_obj[t] = ctx frames (t < n_ctx) take the running prediction rot6d_out, the
// others the input rotation rot6d_obj; rows whose mask is set are zeroed.
__global__ void infill_pack_clip_kernel(const float* __restrict__ rot6d_smpl, const float* __restrict__ trans_smpl, const float* __restrict__ rot6d_obj,
                                        const float* __restrict__ rot6d_out, const unsigned char* __restrict__ mask, int start, int T, int n_ctx,
                                        float* __restrict__ data_smpl, float* __restrict__ data_obj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * 153) return;
  const int t = i / 153, c = i % 153, f = start + t;
  if (c < 144) data_smpl[t * 147 + c] = rot6d_smpl[(size_t)f * 144 + c];
  else if (c < 147) data_smpl[t * 147 + c] = trans_smpl[(size_t)f * 3 + c - 144];
  else {
    const int k = c - 147;
    const float v = t < n_ctx ? rot6d_out[(size_t)f * 6 + k] : rot6d_obj[(size_t)f * 6 + k];
    data_obj[t * 6 + k] = mask[t] ? v * 0.0f : v;      // data * (1 - mask): keeps NaN / inf of the reference arithmetic
  }
}

__global__ void infill_commit_clip_kernel(const float* __restrict__ pred, int start, int t0, int T, float* __restrict__ rot6d_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (T - t0) * 6) return;
  const int t = t0 + i / 6, k = i % 6;
  rot6d_out[(size_t)(start + t) * 6 + k] = pred[t * 6 + k];
}

static int if_check_dims(const char* who, int n_tok, int T, int D, int F, int heads) {
  VT_CHECK_ARG(n_tok > 0 && T > 0 && n_tok % T == 0, "%s: %d tokens are not whole clips of %d", who, n_tok, T);
  VT_CHECK_ARG(T <= IF_MAX_T, "%s: clips of %d frames (at most %d)", who, T, IF_MAX_T);
  VT_CHECK_ARG(D > 0 && D <= IF_MAX_D && heads > 0 && D % heads == 0, "%s: model width %d with %d heads", who, D, heads);
  VT_CHECK_ARG(F >= 0 && F <= IF_MAX_F, "%s: feed-forward width %d (at most %d)", who, F, IF_MAX_F);
  return 0;
}

static int if_launch_token(const IfTokenArgs& a, const char* who, cudaStream_t st) {
  const size_t smem = (size_t)(3 * IF_TB * if_pad4(a.D) + IF_TB * if_pad4(max(max(a.F, a.in_dim), 1))) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(infill_token_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, who);
  }
  infill_token_kernel<<<ceil_div(a.n_tok, IF_TB), IF_THREADS, smem, st>>>(a);
  VT_CHECK_LAUNCH(who);
  return 0;
}

}  // namespace vt

using namespace vt;

extern "C" {

long long vt_infill_layer_pack_floats(int D, int F) { return if_layer(D, F).total; }

int vt_infill_head(const float* in, int in_ld, int in_dim, const float* proj, float* x, int x_ld, int n_tok, int T, int D, int F, int heads,
                   const float* layer, const float* pos, float* qkv, void* stream) {
  if (if_check_dims("vt_infill_head", n_tok, T, D, F, heads)) return -1;
  VT_CHECK_ARG(layer && pos && qkv && x, "vt_infill_head: null pointer");
  VT_CHECK_ARG(!in || (proj && in_dim > 0 && in_dim <= IF_MAX_F && in_ld >= in_dim), "vt_infill_head: projection input %d wide (pitch %d)", in_dim, in_ld);
  IfTokenArgs a{};
  a.in = in; a.in_ld = in_ld; a.in_dim = in ? in_dim : 0; a.proj = proj;
  a.x = x; a.x_ld = x_ld; a.y = x; a.y_ld = x_ld;
  a.head = layer; a.pos = pos; a.qkv = qkv;
  a.n_tok = n_tok; a.T = T; a.D = D; a.F = 0; a.Fh = F; a.act = 0; a.heads = heads;
  return if_launch_token(a, "vt_infill_head", (cudaStream_t)stream);
}

int vt_infill_attn(const float* qkv, const unsigned char* key_mask, int n_clips, int T, int D, int heads, float* attn, void* stream) {
  if (if_check_dims("vt_infill_attn", n_clips * T, T, D, 0, heads)) return -1;
  const int dh = D / heads;
  const size_t smem = ((size_t)T * (dh | 1) + IF_QB * dh + IF_QB * T) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(infill_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "vt_infill_attn smem attr");
  }
  infill_attn_kernel<<<dim3(ceil_div(T, IF_QB), heads, n_clips), IF_THREADS, smem, (cudaStream_t)stream>>>(qkv, key_mask, T, D, heads, attn);
  VT_CHECK_LAUNCH("vt_infill_attn");
  return 0;
}

int vt_infill_tail(float* x, int x_ld, const float* attn, int n_tok, int T, int D, int F, int heads, int activation, const float* layer,
                   const float* final_ln, float* y, int y_ld, const float* next_layer, int next_F, const float* pos, float* qkv, void* stream) {
  if (if_check_dims("vt_infill_tail", n_tok, T, D, F, heads)) return -1;
  VT_CHECK_ARG(x && attn && layer && y, "vt_infill_tail: null pointer");
  VT_CHECK_ARG(activation >= 0 && activation <= 2, "vt_infill_tail: activation %d (0 gelu, 1 relu, 2 leaky_relu)", activation);
  VT_CHECK_ARG(!next_layer || (pos && qkv && next_F >= 0 && next_F <= IF_MAX_F), "vt_infill_tail: next layer needs pos and qkv");
  IfTokenArgs a{};
  a.x = x; a.x_ld = x_ld; a.y = y; a.y_ld = y_ld; a.attn = attn; a.tail = layer; a.final_ln = final_ln;
  a.head = next_layer; a.pos = pos; a.qkv = qkv;
  a.n_tok = n_tok; a.T = T; a.D = D; a.F = F; a.Fh = next_F; a.act = activation; a.heads = heads;
  return if_launch_token(a, "vt_infill_tail", (cudaStream_t)stream);
}

int vt_infill_mlp(const float* x, int x_ld, int n_tok, int n_layers, const int* dims, const float* pack, float* out, int out_ld, void* stream) {
  VT_CHECK_ARG(n_layers >= 1 && n_layers <= 5, "vt_infill_mlp: %d layers (1..5)", n_layers);
  VT_CHECK_ARG(x && pack && out && dims && n_tok > 0, "vt_infill_mlp: null pointer or no tokens");
  IfMlpArgs a{};
  a.x = x; a.x_ld = x_ld; a.n_tok = n_tok; a.n_layers = n_layers; a.pack = pack; a.out = out; a.out_ld = out_ld;
  int wmax = 0;
  for (int i = 0; i <= n_layers; ++i) {
    VT_CHECK_ARG(dims[i] > 0 && dims[i] <= IF_MAX_F, "vt_infill_mlp: width %d of layer %d", dims[i], i);
    a.dims[i] = dims[i];
    wmax = max(wmax, dims[i]);
  }
  const size_t smem = (size_t)2 * IF_TB * if_pad4(wmax) * sizeof(float);
  infill_mlp_kernel<<<ceil_div(n_tok, IF_TB), IF_THREADS, smem, (cudaStream_t)stream>>>(a);
  VT_CHECK_LAUNCH("vt_infill_mlp");
  return 0;
}

int vt_infill_pack_clip(const float* rot6d_smpl, const float* trans_smpl, const float* rot6d_obj, const float* rot6d_out, const unsigned char* mask,
                        int L, int start, int T, int n_ctx, float* data_smpl, float* data_obj, void* stream) {
  VT_CHECK_ARG(start >= 0 && T > 0 && start + T <= L, "vt_infill_pack_clip: clip [%d, %d) outside the %d frames", start, start + T, L);
  VT_CHECK_ARG(n_ctx >= 0 && n_ctx <= T, "vt_infill_pack_clip: %d context frames in a clip of %d", n_ctx, T);
  infill_pack_clip_kernel<<<ceil_div(T * 153, 256), 256, 0, (cudaStream_t)stream>>>(rot6d_smpl, trans_smpl, rot6d_obj, rot6d_out, mask, start, T, n_ctx,
                                                                                    data_smpl, data_obj);
  VT_CHECK_LAUNCH("vt_infill_pack_clip");
  return 0;
}

int vt_infill_commit_clip(const float* pred, int L, int start, int t0, int T, float* rot6d_out, void* stream) {
  VT_CHECK_ARG(start >= 0 && T > 0 && start + T <= L && t0 >= 0 && t0 <= T, "vt_infill_commit_clip: clip [%d, %d) from %d outside the %d frames", start,
               start + T, t0, L);
  if (t0 == T) return 0;
  infill_commit_clip_kernel<<<ceil_div((T - t0) * 6, 256), 256, 0, (cudaStream_t)stream>>>(pred, start, t0, T, rot6d_out);
  VT_CHECK_LAUNCH("vt_infill_commit_clip");
  return 0;
}

}  // extern "C"
