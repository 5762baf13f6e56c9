// One-time re-packing of a reference checkpoint into the layouts the sm_100a kernels read -- HOST code behind the C ABI
// (`vt_pack_weights_<op>`), so that a consumer without Python can load `CHORETriplaneVisibility.state_dict()` tensors
// (model/chore_tri_vis.py, model/HGFilters.py: NCHW Conv2d weights; model/chore.py:113-126: Conv1d(k=1) decoders) and the SMPL-H model
// buffers (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:30-71).  Inputs and outputs are HOST pointers (the caller uploads the
// packed buffers); nothing is allocated or kept.  vistracker_b200/weights.py and smpl.py are thin callers of these functions.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int PK_KC = 64;                    // conv_mma.cu MM_KC: input channels per K chunk
constexpr int PK_H = 128, PK_NREF = 611, PK_K = 616, PK_KB = 640, PK_NHEAD = 5;
static const int kHeadOut[PK_NHEAD] = {2, 9, 14, 3, 1};   // df | pca | parts | centers | visibility

static inline unsigned short half_bits(float x) {
  const __half h = __float2half_rn(x);
  unsigned short u;
  memcpy(&u, &h, 2);
  return u;
}
static inline float half_value(unsigned short u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}

// internal feature index -> reference feature index (model/chore_triplane.py:139-151: im_feat 256 | x, y, z - 2.2 | tmpx 64 | tri_tmpx r, b, t
// 3 x 32 | tri_feat r, b, t 3 x 64); the CUDA-core kernels (query.cu) move the three scalars to the end
static inline int perm_cc(int i) { return i < 256 ? i : (i < 608 ? i + 3 : 256 + (i - 608)); }
// tensor-core kernels (query_tc.cu, 640 padded): im_feat 256 | tmpx 64 | tri_feat r, b, t 192 | tri_tmpx r, b 64 | tri_tmpx t 32, x, y, z - 2.2, 29 zeros
static inline int perm_tc(int i) {
  if (i < 256) return i;
  if (i < 320) return 259 + (i - 256);
  if (i < 512) return 419 + (i - 320);
  if (i < 576) return 323 + (i - 512);
  if (i < 608) return 387 + (i - 576);
  if (i < 611) return 256 + (i - 608);
  return -1;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_conv_cin_pad(int cin) { return (cin + PK_KC - 1) / PK_KC * PK_KC; }

int vt_pack_weights_conv(const float* w, int cout, int cin, int ks, float* ffma, void* hi, void* lo) {
  VT_CHECK_ARG(w != nullptr && cout > 0 && cin > 0 && ks > 0, "vt_pack_weights_conv: w [%d][%d][%d][%d]", cout, cin, ks, ks);
  VT_CHECK_ARG((hi == nullptr) == (lo == nullptr), "vt_pack_weights_conv: hi and lo planes come together");
  const int taps = ks * ks, cin_pad = vt_conv_cin_pad(cin);
  unsigned short* ph = (unsigned short*)hi;
  unsigned short* pl = (unsigned short*)lo;
  if (ph) {
    memset(ph, 0, (size_t)taps * cout * cin_pad * 2);
    memset(pl, 0, (size_t)taps * cout * cin_pad * 2);
  }
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < taps; ++t) {
        const float x = w[((size_t)o * cin + c) * taps + t];
        if (ffma) ffma[((size_t)t * cin + c) * cout + o] = x;
        if (ph) {
          if (!(fabsf(x) <= 65504.f)) { set_error("vt_pack_weights_conv: weight %g exceeds the fp16 range of the fp16 x 2 tensor-core path", x); return -4; }
          const unsigned short h = half_bits(x);
          const size_t at = ((size_t)t * cout + o) * cin_pad + c;
          ph[at] = h;
          pl[at] = half_bits((x - half_value(h)) * kLoScale);
        }
      }
  return 0;
}

int vt_pack_weights_stem(const float* w, int cout, int cin, float* out) {
  VT_CHECK_ARG(w != nullptr && out != nullptr && cout > 0 && cin > 0, "vt_pack_weights_stem: w [%d][%d][7][7]", cout, cin);
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 49; ++t) out[((size_t)t * cin + c) * cout + o] = w[((size_t)o * cin + c) * 49 + t];
  return 0;
}

// w[h * 4 + l], b[h * 4 + l]: Conv1d(k = 1) weight [out][in] and bias [out] of head h (df, pca, parts, centers, visibility), layer l
// (in = 611, 128, 128, 128; out = 128, 128, 128, n_out(h)).
int vt_pack_weights_decoders(const float* const* w, const float* const* b, float* wpack, float* wpack_bwd) {
  VT_CHECK_ARG(w != nullptr && b != nullptr, "vt_pack_weights_decoders: weight / bias tables are required");
  if (wpack) {
    // per head: W1 [616][128] (k-major, rows in the internal feature order, zero padded) | b1 | W2^T [128][128] | b2 | W3^T | b3 | W4^T [128][16] | b4 [16]
    float* p = wpack;
    for (int h = 0; h < PK_NHEAD; ++h) {
      const float* w1 = w[h * 4];
      memset(p, 0, (size_t)PK_K * PK_H * 4);
      for (int i = 0; i < PK_NREF; ++i)
        for (int u = 0; u < PK_H; ++u) p[(size_t)i * PK_H + u] = w1[(size_t)u * PK_NREF + perm_cc(i)];
      p += (size_t)PK_K * PK_H;
      memcpy(p, b[h * 4], PK_H * 4); p += PK_H;
      for (int l = 1; l < 3; ++l) {
        const float* wl = w[h * 4 + l];
        for (int j = 0; j < PK_H; ++j)
          for (int u = 0; u < PK_H; ++u) p[(size_t)j * PK_H + u] = wl[(size_t)u * PK_H + j];
        p += (size_t)PK_H * PK_H;
        memcpy(p, b[h * 4 + l], PK_H * 4); p += PK_H;
      }
      const int nout = kHeadOut[h];
      memset(p, 0, (size_t)(PK_H * 16 + 16) * 4);
      for (int j = 0; j < PK_H; ++j)
        for (int o = 0; o < nout; ++o) p[j * 16 + o] = w[h * 4 + 3][(size_t)o * PK_H + j];
      p += PK_H * 16;
      memcpy(p, b[h * 4 + 3], nout * 4); p += 16;
    }
    if (p - wpack != vt_query_wpack_floats()) { set_error("vt_pack_weights_decoders: layout drifted from csrc/query.cu"); return -2; }
  }
  if (wpack_bwd) {
    // per head, torch [out][in] layout: W1b [128][640] (internal feature order, zero padded) | W2b | W3b [128][128] | W4b [16][128]
    float* p = wpack_bwd;
    for (int h = 0; h < PK_NHEAD; ++h) {
      memset(p, 0, (size_t)PK_H * PK_KB * 4);
      for (int u = 0; u < PK_H; ++u)
        for (int i = 0; i < PK_NREF; ++i) p[(size_t)u * PK_KB + i] = w[h * 4][(size_t)u * PK_NREF + perm_cc(i)];
      p += (size_t)PK_H * PK_KB;
      for (int l = 1; l < 3; ++l) { memcpy(p, w[h * 4 + l], (size_t)PK_H * PK_H * 4); p += (size_t)PK_H * PK_H; }
      memset(p, 0, 16 * PK_H * 4);
      memcpy(p, w[h * 4 + 3], (size_t)kHeadOut[h] * PK_H * 4);
      p += 16 * PK_H;
    }
    if (p - wpack_bwd != vt_query_wpack_bwd_floats()) { set_error("vt_pack_weights_decoders: backward layout drifted from csrc/query.cu"); return -2; }
  }
  return 0;
}

// fp16 hi / lo planes of the tcgen05 decoder kernels (x ~= hi + lo, lo not rescaled): W1 [5 * 128][640], W2 | W3 [2 * 5 * 128][128] (layer-major,
// then head; torch [out][in] = K-major B operands), their transposes W2^T | W3^T [2 * 5 * 128][128] and W1^T [5 * 640][128].  The four
// backward planes may be null.
int vt_pack_weights_decoders_tc(const float* const* w, void* w1_hi, void* w1_lo, void* w23_hi, void* w23_lo, void* w23t_hi, void* w23t_lo,
                                void* w1t_hi, void* w1t_lo) {
  VT_CHECK_ARG(w != nullptr && w1_hi && w1_lo && w23_hi && w23_lo, "vt_pack_weights_decoders_tc: the forward planes are required");
  const bool bwd = w23t_hi != nullptr;
  VT_CHECK_ARG(bwd == (w23t_lo != nullptr) && bwd == (w1t_hi != nullptr) && bwd == (w1t_lo != nullptr),
               "vt_pack_weights_decoders_tc: the four backward planes come together");
  int bad = 0;
  auto put = [&](void* hi, void* lo, size_t at, float x) {
    if (!(fabsf(x) <= 65504.f)) bad = 1;
    const unsigned short h = half_bits(x);
    ((unsigned short*)hi)[at] = h;
    ((unsigned short*)lo)[at] = half_bits(x - half_value(h));
  };
  for (int h = 0; h < PK_NHEAD; ++h) {
    for (int u = 0; u < PK_H; ++u)
      for (int i = 0; i < PK_KB; ++i) {
        const int r = perm_tc(i);
        const float x = r >= 0 ? w[h * 4][(size_t)u * PK_NREF + r] : 0.f;
        put(w1_hi, w1_lo, ((size_t)h * PK_H + u) * PK_KB + i, x);
        if (bwd) put(w1t_hi, w1t_lo, ((size_t)h * PK_KB + i) * PK_H + u, x);
      }
    for (int l = 0; l < 2; ++l)
      for (int u = 0; u < PK_H; ++u)
        for (int j = 0; j < PK_H; ++j) {
          const float x = w[h * 4 + 1 + l][(size_t)u * PK_H + j];
          put(w23_hi, w23_lo, ((size_t)(l * PK_NHEAD + h) * PK_H + u) * PK_H + j, x);
          if (bwd) put(w23t_hi, w23t_lo, ((size_t)(l * PK_NHEAD + h) * PK_H + j) * PK_H + u, x);
        }
  }
  if (bad) { set_error("vt_pack_weights_decoders_tc: a weight exceeds the fp16 range"); return -4; }
  return 0;
}

// ---- SMPL-H: th_v_template [V][3], th_shapedirs [V][3][n_betas], th_posedirs [V][3][9 (J - 1)], th_J_regressor [J][V] (float64, as the
// reference's pickles hold them), th_weights [V][J] fp32, kintree parents [J] -> the buffers of SmplModel (include/vistracker_b200.h)
static inline int pad4(int x) { return (x + 3) / 4 * 4; }

int vt_smpl_pack_dims(int V, int J, int n_betas, const float* weights, int* kd, int* kdp, int* nv3p, int* nnz) {
  VT_CHECK_ARG(V > 0 && J > 1 && n_betas > 0 && weights != nullptr, "vt_smpl_pack_dims: V %d J %d n_betas %d", V, J, n_betas);
  const int k = 9 * (J - 1) + n_betas;
  int most = 0;
  for (int v = 0; v < V; ++v) {
    int c = 0;
    for (int j = 0; j < J; ++j) c += weights[(size_t)v * J + j] != 0.f;
    most = c > most ? c : most;
  }
  if (kd) *kd = k;
  if (kdp) *kdp = pad4(k);
  if (nv3p) *nv3p = pad4(3 * V);
  if (nnz) *nnz = most;
  return 0;
}

int vt_pack_weights_smpl(const double* v_template, const double* shapedirs, const double* posedirs, const double* J_regressor,
                         const float* weights, const int* parents, int V, int J, int n_betas, float* templ, float* dirs, float* dirsT,
                         float* j_templ, float* j_dirs, int* parents_out, int* skin_idx, float* skin_w) {
  VT_CHECK_ARG(v_template && shapedirs && posedirs && J_regressor && weights && parents, "vt_pack_weights_smpl: all model buffers are required");
  VT_CHECK_ARG(templ && dirs && dirsT && j_templ && j_dirs && parents_out && skin_idx && skin_w, "vt_pack_weights_smpl: all outputs are required");
  int kd, kdp, nv3p, nnz;
  if (int rc = vt_smpl_pack_dims(V, J, n_betas, weights, &kd, &kdp, &nv3p, &nnz)) return rc;
  const int np = 9 * (J - 1);
  for (int i = 0; i < 3 * V; ++i) templ[i] = (float)v_template[i];
  // dirs [kdp][nv3p]: rows 0 .. 9 (J - 1) - 1 pose blend shapes, then the shape blend shapes; dirsT its transpose
  memset(dirs, 0, (size_t)kdp * nv3p * 4);
  memset(dirsT, 0, (size_t)kdp * nv3p * 4);
  for (int i = 0; i < 3 * V; ++i) {
    for (int k = 0; k < np; ++k) { const float x = (float)posedirs[(size_t)i * np + k]; dirs[(size_t)k * nv3p + i] = x; dirsT[(size_t)i * kdp + k] = x; }
    for (int k = 0; k < n_betas; ++k) {
      const float x = (float)shapedirs[(size_t)i * n_betas + k];
      dirs[(size_t)(np + k) * nv3p + i] = x; dirsT[(size_t)i * kdp + np + k] = x;
    }
  }
  // joint template J_regressor @ v_template and its shape derivative einsum('jv,vck->jck'), accumulated in float64
  for (int j = 0; j < J; ++j) {
    double acc[3] = {0, 0, 0};
    for (int v = 0; v < V; ++v) {
      const double r = J_regressor[(size_t)j * V + v];
      if (r == 0.0) continue;
      for (int c = 0; c < 3; ++c) acc[c] += r * v_template[v * 3 + c];
    }
    for (int c = 0; c < 3; ++c) j_templ[j * 3 + c] = (float)acc[c];
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < n_betas; ++k) {
        double a = 0;
        for (int v = 0; v < V; ++v) a += J_regressor[(size_t)j * V + v] * shapedirs[((size_t)v * 3 + c) * n_betas + k];
        j_dirs[((size_t)j * 3 + c) * n_betas + k] = (float)a;
      }
  }
  for (int j = 0; j < J; ++j) parents_out[j] = parents[j] > 0 ? parents[j] : 0;
  // skinning weights compacted to the nnz non-zero entries of a vertex (joint order kept), zero entries filling up
  for (int v = 0; v < V; ++v) {
    int n = 0;
    for (int j = 0; j < J && n < nnz; ++j)
      if (weights[(size_t)v * J + j] != 0.f) { skin_idx[(size_t)v * nnz + n] = j; skin_w[(size_t)v * nnz + n] = weights[(size_t)v * J + j]; ++n; }
    for (int j = 0; j < J && n < nnz; ++j)
      if (weights[(size_t)v * J + j] == 0.f) { skin_idx[(size_t)v * nnz + n] = j; skin_w[(size_t)v * nnz + n] = 0.f; ++n; }
  }
  return 0;
}

// ---- caller-owned workspaces.  Only three operators take one; every other entry point works in the buffers named in its signature.
long long vt_workspace_bytes_raster_cull(int B, int F) { return 4 * vt_raster_cull_floats(B, F); }
long long vt_workspace_bytes_procrustes(int B) { return B > 0 ? (long long)B * 16 * 8 : 0; }

}  // extern "C"
