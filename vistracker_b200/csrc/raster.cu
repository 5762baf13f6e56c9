// Triangle rasteriser with the semantics of `neural_renderer` (the PyTorch port of Kato et al., "Neural 3D Mesh Renderer") as the
// reference drives it:
//   * silhouettes in `projection` camera mode for the occlusion-aware mask loss (recon/obj_pose_roi.py:87-94,183-202):
//     per-frame ROI intrinsics K, R = I, t = 0, orig_size = 1, anti_aliasing = False, fill_back = True, near 0.1, far 100;
//   * orthographic depth / occupancy in `look` mode for the triplane renderings (render/render_triplane_nr.py:25-30,88-110).
// neural_renderer is NOT vendored by the reference and not installable here, so this file restates its published algorithm from
// the upstream sources as recalled (PARITY UNPINNED -- see DESIGN.md): pixel centres at (2i + 1 - S) / S, a pixel is covered when
// it passes the three edge tests of a counter-clockwise face (the reversed copies added by fill_back make every triangle CCW
// once), depth 1 / sum(w_k / z_k) must lie in (near, far), nearest face wins (first one on ties), image rows are flipped on
// output.  The backward pass is NMR's hand-designed pseudo-gradient: for every face edge and both axes, walk the pixel
// rows/columns the edge crosses and, where moving the edge would flip pixels whose intensity change reduces the loss
// (diff_grad > 0), add  -diff_grad / distance  to the two edge vertices.
//
// Forward is tiled: a CTA owns a 16x16 pixel tile, culls the face list against the tile in order-preserving chunks (ballot
// compaction into shared memory) and only then runs the per-pixel tests -- upstream loops every pixel over every face.
// (Measured, profiles/r02l_raster_fwd_ncu.txt: the 0.33 ms for the 96 silhouettes of a batch were NOT the culling -- a 64x64 super-tile variant
// that culled once per 16 tiles took the same 0.33 ms -- but the depth of the covered pixels: ten IEEE divisions per covered (pixel, face) pair,
// executed by warps in which a small face leaves most lanes idle.  The arithmetic is upstream's and stays; the depths are now evaluated
// lane-parallel over different faces, see the per-pixel loop.)
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

constexpr int RT = 16;                 // tile edge
constexpr int RCHUNK = 256;            // faces examined per compaction round
static_assert(RCHUNK <= 256, "the forward kernel packs slot numbers of a chunk into bytes");
constexpr float R_NEAR = 0.1f, R_FAR = 100.f, R_EPS = 1e-3f;

// camera: mode 0 = nr.projection with per-frame (fx, fy, cx, cy), orig_size 1, eps 1e-9, no distortion; mode 1 = (x, y, z) as given
__device__ __forceinline__ void to_ndc(const float* v, int mode, const float* K, float& u, float& w, float& z) {
  if (mode == 0) {
    const float zz = v[2] + 1e-9f;
    const float px = K[0] * (v[0] / zz) + K[2];
    const float py = K[1] * (v[1] / zz) + K[3];
    u = 2.f * (px - 0.5f);
    w = 2.f * ((1.f - py) - 0.5f);
    z = v[2];
  } else { u = v[0]; w = v[1]; z = v[2]; }
}

// faces_ndc[b][f][9]: (x, y, z) of the three vertices; faces f >= F are the reversed copies (fill_back)
__global__ void raster_setup_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int B, int V, int F, int mode,
                                    const float* __restrict__ K, float* __restrict__ faces_ndc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * F) return;
  const int b = i / (2 * F), f2 = i % (2 * F), f = f2 % F;
  const bool rev = f2 >= F;
  for (int k = 0; k < 3; ++k) {
    const int vi = faces[f * 3 + (rev ? 2 - k : k)];
    float u, w, z;
    to_ndc(verts + ((size_t)b * V + vi) * 3, mode, K ? K + (size_t)b * 4 : nullptr, u, w, z);
    float* o = faces_ndc + (size_t)i * 9 + k * 3;
    o[0] = u; o[1] = w; o[2] = z;
  }
}

// Culling records for the tiled forward pass: face_bbox[b][f] = (min x, max x, min y, max y) of a counter-clockwise face (a clockwise one
// gets an empty box), chunk_bbox[b][c] = the union over faces [256 c, 256 c + 256).  A tile skips every chunk whose box misses it -- mesh
// faces are spatially coherent in index order, so a tile of a 1024^2 SMPL rendering scans a few chunks instead of all 27 552 faces.
__global__ void __launch_bounds__(RCHUNK) raster_bbox_kernel(const float* __restrict__ faces_ndc, int nf, int nchunks, float4* __restrict__ face_bbox,
                                                              float4* __restrict__ chunk_bbox) {
  __shared__ float4 s_red[RCHUNK / 32];
  const int b = blockIdx.y, f = blockIdx.x * RCHUNK + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 bb = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  if (f < nf) {
    const float* p = faces_ndc + ((size_t)b * nf + f) * 9;
    const bool front = !((p[7] - p[1]) * (p[3] - p[0]) < (p[4] - p[1]) * (p[6] - p[0]));
    if (front) bb = make_float4(fminf(p[0], fminf(p[3], p[6])), fmaxf(p[0], fmaxf(p[3], p[6])), fminf(p[1], fminf(p[4], p[7])), fmaxf(p[1], fmaxf(p[4], p[7])));
    face_bbox[(size_t)b * nf + f] = bb;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    bb.x = fminf(bb.x, __shfl_xor_sync(0xffffffffu, bb.x, o)); bb.y = fmaxf(bb.y, __shfl_xor_sync(0xffffffffu, bb.y, o));
    bb.z = fminf(bb.z, __shfl_xor_sync(0xffffffffu, bb.z, o)); bb.w = fmaxf(bb.w, __shfl_xor_sync(0xffffffffu, bb.w, o));
  }
  if (lane == 0) s_red[warp] = bb;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RCHUNK / 32; ++w) {
      bb.x = fminf(bb.x, s_red[w].x); bb.y = fmaxf(bb.y, s_red[w].y); bb.z = fminf(bb.z, s_red[w].z); bb.w = fmaxf(bb.w, s_red[w].w);
    }
    chunk_bbox[(size_t)b * nchunks + blockIdx.x] = bb;
  }
}

// One pixel centre against one counter-clockwise face q = (x, y, z) x 3 in NDC, in two parts: the three edge tests (covered or not), and the
// depth 1 / sum(w_k / z_k) of a covered pixel with the clamped, renormalised barycentric weights (false when it is outside (near, far)).
__device__ __forceinline__ bool raster_edges(const float* q, float xp, float yp) {
  return !(((yp - q[1]) * (q[3] - q[0]) < (xp - q[0]) * (q[4] - q[1])) ||
           ((yp - q[4]) * (q[6] - q[3]) < (xp - q[3]) * (q[7] - q[4])) ||
           ((yp - q[7]) * (q[0] - q[6]) < (xp - q[6]) * (q[1] - q[7])));
}
__device__ __forceinline__ bool raster_depth(const float* q, float xp, float yp, float& zp) {
  const float den = q[6] * (q[1] - q[4]) + q[0] * (q[4] - q[7]) + q[3] * (q[7] - q[1]);
  float w0 = ((q[4] - q[7]) * xp + (q[6] - q[3]) * yp + (q[3] * q[7] - q[6] * q[4])) / den;
  float w1 = ((q[7] - q[1]) * xp + (q[0] - q[6]) * yp + (q[6] * q[1] - q[0] * q[7])) / den;
  float w2 = ((q[1] - q[4]) * xp + (q[3] - q[0]) * yp + (q[0] * q[4] - q[3] * q[1])) / den;
  w0 = fminf(fmaxf(w0, 0.f), 1.f); w1 = fminf(fmaxf(w1, 0.f), 1.f); w2 = fminf(fmaxf(w2, 0.f), 1.f);
  const float ws = fmaxf(w0 + w1 + w2, 1e-10f);
  w0 /= ws; w1 /= ws; w2 /= ws;
  zp = 1.f / (w0 / q[2] + w1 / q[5] + w2 / q[8]);
  return !(zp <= R_NEAR || zp >= R_FAR);
}

__global__ void __launch_bounds__(RT * RT) raster_fwd_kernel(const float* __restrict__ faces_ndc, int nf, int is,
                                                             int* __restrict__ face_index /*[B][is][is] internal (y up)*/,
                                                             float* __restrict__ alpha /*[B][is][is] image rows (flipped) or null*/,
                                                             float* __restrict__ depth /*[B][is][is] image rows or null*/,
                                                             const float4* __restrict__ face_bbox, const float4* __restrict__ chunk_bbox) {
  __shared__ float sf[RCHUNK][9];
  __shared__ int s_id[RCHUNK];
  __shared__ int s_warp_cnt[RT * RT / 32];
  __shared__ int s_total;
  const int b = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int xi = blockIdx.x * RT + tid % RT, yi = blockIdx.y * RT + tid / RT;
  const float xp = (2.f * xi + 1.f - is) / is, yp = (2.f * yi + 1.f - is) / is;
  const float tx0 = (2.f * (blockIdx.x * RT) + 1.f - is) / is, tx1 = (2.f * (blockIdx.x * RT + RT - 1) + 1.f - is) / is;
  const float ty0 = (2.f * (blockIdx.y * RT) + 1.f - is) / is, ty1 = (2.f * (blockIdx.y * RT + RT - 1) + 1.f - is) / is;
  float depth_min = R_FAR; int best = -1;
  const float* fb = faces_ndc + (size_t)b * nf * 9;
  const int nchunks = (nf + RCHUNK - 1) / RCHUNK;
  for (int c0 = 0; c0 < nf; c0 += RCHUNK) {
    if (chunk_bbox) {                                  // uniform over the CTA: no barrier has been entered in this iteration yet
      const float4 cb = __ldg(chunk_bbox + (size_t)b * nchunks + c0 / RCHUNK);
      if (!(cb.y >= tx0 && cb.x <= tx1 && cb.w >= ty0 && cb.z <= ty1)) continue;
    }
    // ---- cull: keep counter-clockwise faces whose bounding box touches the tile, preserving face order
    const int f = c0 + tid;
    float p[9];
    bool keep = false;
    if (f < nf && face_bbox) {
      const float4 bb = __ldg(face_bbox + (size_t)b * nf + f);
      keep = bb.y >= tx0 && bb.x <= tx1 && bb.w >= ty0 && bb.z <= ty1;
      if (keep) {
#pragma unroll
        for (int k = 0; k < 9; ++k) p[k] = fb[(size_t)f * 9 + k];
      }
    } else if (f < nf) {
#pragma unroll
      for (int k = 0; k < 9; ++k) p[k] = fb[(size_t)f * 9 + k];
      const bool front = !((p[7] - p[1]) * (p[3] - p[0]) < (p[4] - p[1]) * (p[6] - p[0]));
      const float mnx = fminf(p[0], fminf(p[3], p[6])), mxx = fmaxf(p[0], fmaxf(p[3], p[6]));
      const float mny = fminf(p[1], fminf(p[4], p[7])), mxy = fmaxf(p[1], fmaxf(p[4], p[7]));
      keep = front && mxx >= tx0 && mnx <= tx1 && mxy >= ty0 && mny <= ty1;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
    if (tid == RT * RT - 1) s_total = off + __popc(bal);
    if (keep) {
      const int slot = off + __popc(bal & ((1u << lane) - 1));
#pragma unroll
      for (int k = 0; k < 9; ++k) sf[slot][k] = p[k];
      s_id[slot] = f;
    }
    __syncthreads();
    const int cnt = s_total;
    // ---- per-pixel tests over the compacted faces
    // Edge tests first, depths later: a face covers a few of a warp's 32 pixels, so evaluating the depth (ten IEEE divisions) inside the face loop ran
    // it with most lanes idle -- it was the bulk of this kernel (profiles/r02l_raster_fwd_ncu.txt).  Every lane queues the slots of the faces
    // that cover ITS pixel (up to four, packed into one word) and the warp evaluates the queued depths together, each lane on its own face, when
    // a queue is full and at the end of the chunk: a pixel still sees its faces in face order with the same arithmetic -> same image.
    {
      const bool inside = xi < is && yi < is;
      uint32_t pend = 0;
      int np = 0;
      auto flush = [&]() {
#pragma unroll 1
        for (int t = 0; __any_sync(0xffffffffu, t < np); ++t) {
          if (t < np) {
            const int j = (pend >> (8 * t)) & 0xFF;
            float zp;
            if (raster_depth(sf[j], xp, yp, zp) && zp < depth_min) { depth_min = zp; best = s_id[j]; }
          }
        }
        pend = 0; np = 0;
      };
      for (int j = 0; j < cnt; ++j) {
        if (inside && raster_edges(sf[j], xp, yp)) { pend |= (uint32_t)j << (8 * np); ++np; }
        if (__any_sync(0xffffffffu, np == 4)) flush();
      }
      if (__any_sync(0xffffffffu, np > 0)) flush();
    }
    __syncthreads();
  }
  if (xi < is && yi < is) {
    face_index[((size_t)b * is + yi) * is + xi] = best;
    const size_t o = ((size_t)b * is + (is - 1 - yi)) * is + xi;          // vertical flip on output
    if (alpha) alpha[o] = best >= 0 ? 1.f : 0.f;
    if (depth) depth[o] = depth_min;
  }
}

// Inclusive prefix counts of the "active" background pixels -- alpha == 0 and g_alpha < 0, the only ones the OUT walk of raster_bwd_kernel
// can take a contribution from ((a - alpha_in) g_alpha > 0 with alpha_in = 1) -- along every column (z = 0: P[b][col][row + 1]) and every row
// (z = 1: P[b][row][col + 1]) of the y-up internal maps.  One warp scans one line.  The backward kernel skips a walk whose range holds none:
// in the silhouette loss of the fitters the active pixels are the part of the target mask the render does not cover yet, a thin band,
// while every visible edge pixel column used to walk to the image border.
__global__ void __launch_bounds__(256) raster_active_prefix_kernel(const float* __restrict__ alpha, const float* __restrict__ g_alpha, int is,
                                                                   unsigned short* __restrict__ pcol, unsigned short* __restrict__ prow) {
  const int line = blockIdx.x * 8 + (threadIdx.x >> 5), bn = blockIdx.y, by_row = blockIdx.z, lane = threadIdx.x & 31;
  if (line >= is) return;                                  // (whole warps: no barrier below)
  unsigned short* P = (by_row ? prow : pcol) + ((size_t)bn * is + line) * (is + 1);
  if (lane == 0) P[0] = 0;
  unsigned carry = 0;
  for (int i0 = 0; i0 < is; i0 += 32) {
    const int i = i0 + lane;
    unsigned act = 0;
    if (i < is) {
      const int row = by_row ? line : i, col = by_row ? i : line;
      const size_t o = ((size_t)bn * is + (is - 1 - row)) * is + col;
      act = (alpha[o] == 0.f && g_alpha[o] < 0.f) ? 1u : 0u;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (i < is) P[i + 1] = (unsigned short)(carry + __popc(bal & (0xffffffffu >> (31 - lane))));
    carry += __popc(bal);
  }
}

// NMR pseudo-gradient of the silhouette w.r.t. the (x, y) of every face vertex.  One HALF-WARP per (frame, face): its 16 lanes split the
// pixel columns / rows an edge crosses (the upstream kernel walks them in one thread: up to `is` iterations, each with an inner walk of
// up to `is` pixels -- latency-bound), partial sums are combined with shuffles at the end.  (With a whole warp per face the kernel ran at 12
// active lanes per instruction and 63 % of its issue slots -- the faces of a fitted template are ~16 pixels wide, profiles/r02l_raster_fwd_ncu.txt;
// consecutive faces are mesh neighbours, so the two halves of a warp walk edges of similar length.)
constexpr int RB_LANES = 16;           // lanes per face in raster_bwd_kernel (measured per step at 96 frames: 32 lanes 0.29 ms, 16: 0.24, 8: 0.25)
__global__ void __launch_bounds__(128) raster_bwd_kernel(const float* __restrict__ faces_ndc, const int* __restrict__ face_index,
                                  const float* __restrict__ alpha /*image rows*/, const float* __restrict__ g_alpha /*image rows*/,
                                  int B, int nf, int is, float* __restrict__ g_faces /*[B][nf][9]*/,
                                  const unsigned short* __restrict__ pcol, const unsigned short* __restrict__ prow /*optional: see above*/) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) / RB_LANES, lane = threadIdx.x & (RB_LANES - 1);
  const bool live = i < B * nf;
  const int bn = live ? i / nf : 0, fn = live ? i % nf : 0;
  const float* face = faces_ndc + (size_t)(live ? i : 0) * 9;
  float grad_face[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  float* out = g_faces + (size_t)(live ? i : 0) * 9;
  // back side: a zero gradient (the sums below stay zero); both halves of the warp reach the shuffles at the end
  const bool front = live && !((face[7] - face[1]) * (face[3] - face[0]) < (face[4] - face[1]) * (face[6] - face[0]));
  // internal maps are y-up: internal (row = y index, col = x index) lives at image row is-1-row
  auto A = [&](int row, int col) { return alpha[((size_t)bn * is + (is - 1 - row)) * is + col]; };
  auto GA = [&](int row, int col) { return g_alpha[((size_t)bn * is + (is - 1 - row)) * is + col]; };
  auto FI = [&](int row, int col) { return face_index[((size_t)bn * is + row) * is + col]; };
  for (int edge = 0; front && edge < 3; ++edge) {
    int pi[3];
    float pp[3][2];
    for (int n = 0; n < 3; ++n) pi[n] = (edge + n) % 3;
    for (int n = 0; n < 3; ++n)
      for (int d = 0; d < 2; ++d) pp[n][d] = 0.5f * (face[3 * pi[n] + d] * is + is - 1);
    for (int axis = 0; axis < 2; ++axis) {
      float p[3][2];
      for (int n = 0; n < 3; ++n)
        for (int d = 0; d < 2; ++d) p[n][d] = pp[n][(d + axis) % 2];
      int direction;
      if (axis == 0) direction = (p[0][0] < p[1][0]) ? -1 : 1;
      else direction = (p[0][0] < p[1][0]) ? 1 : -1;
      const int d0_from = (int)fmaxf(ceilf(fminf(p[0][0], p[1][0])), 0.f);
      const int d0_to = (int)fminf(fmaxf(p[0][0], p[1][0]), (float)(is - 1));
      // the two edge vertices' sums of this (edge, axis) in registers: grad_face is indexed by run-time values and lives in local memory --
      // a load / add / store through it per contributing pixel chained the walks on memory latency
      float acc0 = 0.f, acc1 = 0.f;
      for (int d0 = d0_from + lane; d0 <= d0_to; d0 += RB_LANES) {
        const float d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
        const int d1_in = direction > 0 ? (int)floorf(d1_cross) : (int)ceilf(d1_cross);
        const int d1_out = d1_in + direction;
        if (d1_in < 0 || is <= d1_in) continue;
        if (d1_out < 0 || is <= d1_out) continue;
        // axis 0: d0 runs along x, d1 along y;  axis 1: d0 along y, d1 along x
        const float alpha_in = axis == 0 ? A(d1_in, d0) : A(d0, d1_in);
        const float alpha_out = axis == 0 ? A(d1_out, d0) : A(d0, d1_out);
        const bool is_in_fn = (axis == 0 ? FI(d1_in, d0) : FI(d0, d1_in)) == fn;
        if (is_in_fn) {     // "out": pixels beyond the edge that would become covered
          const int d1_limit = direction > 0 ? is - 1 : 0;
          const int d1_from = max(min(d1_out, d1_limit), 0);
          int d1_to = min(max(d1_out, d1_limit), is - 1);
          int d1_first = d1_from;
          if (pcol) {
            // Only ACTIVE pixels (alpha == 0, g_alpha < 0) pass the `diff_grad > 0` test below.  None in the range: nothing to walk.  Otherwise the
            // walk is clipped to [first active, last active] by two binary searches on the (monotone) prefix counts: in the fitters the active
            // pixels are the band of the target mask the render does not cover yet, while the range runs to the image border (a pixel-by-pixel
            // walk of up to `is` dependent loads was 60 % of this kernel).  Skipped pixels contribute nothing: same sums, same order.
            const unsigned short* P = (axis == 0 ? pcol : prow) + ((size_t)bn * is + d0) * (is + 1);
            const unsigned base = P[d1_from], top = P[d1_to + 1];
            if (top == base) {
              d1_to = d1_from - 1;
            } else {
              int lo = d1_from, hi = d1_to;
              while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid + 1] > base) hi = mid; else lo = mid + 1; }
              d1_first = lo;
              hi = d1_to;
              while (lo < hi) { const int mid = (lo + hi) >> 1; if (P[mid + 1] >= top) hi = mid; else lo = mid + 1; }
              d1_to = lo;
            }
          }
          for (int d1 = d1_first; d1 <= d1_to; ++d1) {
            const float a = axis == 0 ? A(d1, d0) : A(d0, d1);
            const float ga = axis == 0 ? GA(d1, d0) : GA(d0, d1);
            const float diff_grad = (a - alpha_in) * ga;
            if (diff_grad <= 0) continue;
            if (p[1][0] != d0) {
              float dist = (p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2.f / is;
              dist = (0 < dist) ? dist + R_EPS : dist - R_EPS;
              acc0 -= diff_grad / dist;
            }
            if (p[0][0] != d0) {
              float dist = (p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2.f / is;
              dist = (0 < dist) ? dist + R_EPS : dist - R_EPS;
              acc1 -= diff_grad / dist;
            }
          }
        }
        // "in": pixels of this face that would become uncovered.  They are covered (alpha = 1: the index map names this face), so with a covered
        // pixel beyond the edge (alpha_out = 1, every interior edge of the mesh) (a - alpha_out) is zero and the walk cannot contribute: skipped
        // in the workspace mode (only silhouette edges walk); same result bit for bit
        if (!(pcol && alpha_out >= 1.f)) {
          float d0_cross2;
          if ((d0 - p[0][0]) * (d0 - p[2][0]) < 0) d0_cross2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1];
          else d0_cross2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1];
          const int d1_limit = direction > 0 ? (int)ceilf(d0_cross2) : (int)floorf(d0_cross2);
          const int d1_from = max(min(d1_in, d1_limit), 0), d1_to = min(max(d1_in, d1_limit), is - 1);
          for (int d1 = d1_from; d1 <= d1_to; ++d1) {
            if ((axis == 0 ? FI(d1, d0) : FI(d0, d1)) != fn) continue;
            const float a = axis == 0 ? A(d1, d0) : A(d0, d1);
            const float ga = axis == 0 ? GA(d1, d0) : GA(d0, d1);
            const float diff_grad = (a - alpha_out) * ga;
            if (diff_grad <= 0) continue;
            if (p[1][0] != d0) {
              float dist = (p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2.f / is;
              dist = (0 < dist) ? dist + R_EPS : dist - R_EPS;
              acc0 -= diff_grad / dist;
            }
            if (p[0][0] != d0) {
              float dist = (p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2.f / is;
              dist = (0 < dist) ? dist + R_EPS : dist - R_EPS;
              acc1 -= diff_grad / dist;
            }
          }
        }
      }
      grad_face[pi[0] * 3 + (1 - axis)] += acc0;
      grad_face[pi[1] * 3 + (1 - axis)] += acc1;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    float v = grad_face[k];
#pragma unroll
    for (int o = RB_LANES / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);      // within the face's lanes
    if (lane == 0 && live) out[k] = v;
  }
}

// g_faces (NDC x, y per face vertex) -> camera-space vertex gradients through vertices_to_faces and the projection
__global__ void raster_bwd_verts_kernel(const float* __restrict__ g_faces, const float* __restrict__ verts, const int* __restrict__ faces,
                                        int B, int V, int F, int mode, const float* __restrict__ K, float* __restrict__ g_verts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 2 * F) return;
  const int b = i / (2 * F), f2 = i % (2 * F), f = f2 % F;
  const bool rev = f2 >= F;
  for (int k = 0; k < 3; ++k) {
    const float gu = g_faces[(size_t)i * 9 + k * 3], gw = g_faces[(size_t)i * 9 + k * 3 + 1];
    if (gu == 0.f && gw == 0.f) continue;
    const int vi = faces[f * 3 + (rev ? 2 - k : k)];
    const float* v = verts + ((size_t)b * V + vi) * 3;
    float* g = g_verts + ((size_t)b * V + vi) * 3;
    if (mode == 0) {
      const float* Kb = K + (size_t)b * 4;
      const float zz = v[2] + 1e-9f;
      atomicAdd(g, gu * 2.f * Kb[0] / zz);
      atomicAdd(g + 1, -gw * 2.f * Kb[1] / zz);
      atomicAdd(g + 2, (-gu * 2.f * Kb[0] * v[0] + gw * 2.f * Kb[1] * v[1]) / (zz * zz));
    } else {
      atomicAdd(g, gu);
      atomicAdd(g + 1, gw);
    }
  }
}

}  // namespace vt

using namespace vt;

extern "C" {

long long vt_raster_cull_floats(int B, int F) { return (B <= 0 || F <= 0) ? 0 : (long long)B * (2 * F + ceil_div(2 * F, RCHUNK)) * 4; }

int vt_raster_fwd(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                  float* faces_ndc, int* face_index, float* alpha, float* depth, float* cull_ws, void* stream) {
  VT_CHECK_ARG(mode == 0 || mode == 1, "vt_raster_fwd: camera mode %d (0 projection, 1 orthographic look)", mode);
  VT_CHECK_ARG(mode == 1 || K4 != nullptr, "vt_raster_fwd: projection mode needs per-frame intrinsics");
  VT_CHECK_ARG(image_size > 0 && image_size <= 4096, "vt_raster_fwd: image size %d", image_size);
  if (B <= 0 || F <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  raster_setup_kernel<<<ceil_div(B * 2 * F, 256), 256, 0, s>>>(verts, faces, B, V, F, mode, K4, faces_ndc);
  VT_CHECK_LAUNCH("vt_raster_fwd(setup)");
  float4* face_bbox = nullptr;
  float4* chunk_bbox = nullptr;
  if (cull_ws) {
    VT_CHECK_ARG((reinterpret_cast<size_t>(cull_ws) & 15) == 0, "vt_raster_fwd: cull workspace must be 16-byte aligned");
    const int nchunks = ceil_div(2 * F, RCHUNK);
    face_bbox = reinterpret_cast<float4*>(cull_ws);
    chunk_bbox = face_bbox + (size_t)B * 2 * F;
    raster_bbox_kernel<<<dim3(nchunks, B), RCHUNK, 0, s>>>(faces_ndc, 2 * F, nchunks, face_bbox, chunk_bbox);
    VT_CHECK_LAUNCH("vt_raster_fwd(bbox)");
  }
  dim3 grid(ceil_div(image_size, RT), ceil_div(image_size, RT), B);
  raster_fwd_kernel<<<grid, RT * RT, 0, s>>>(faces_ndc, 2 * F, image_size, face_index, alpha, depth, face_bbox, chunk_bbox);
  VT_CHECK_LAUNCH("vt_raster_fwd");
  return 0;
}

long long vt_workspace_bytes_raster_bwd(int B, int image_size) {
  return (B <= 0 || image_size <= 0) ? 0 : (long long)B * 2 * image_size * (image_size + 1) * 2;
}

int vt_raster_bwd(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                  const float* faces_ndc, const int* face_index, const float* alpha, const float* g_alpha, float* g_faces,
                  float* g_verts, void* stream) {
  return vt_raster_bwd_ws(verts, faces, B, V, F, mode, K4, image_size, faces_ndc, face_index, alpha, g_alpha, g_faces, g_verts, nullptr, stream);
}

int vt_raster_bwd_ws(const float* verts, const int* faces, int B, int V, int F, int mode, const float* K4, int image_size,
                     const float* faces_ndc, const int* face_index, const float* alpha, const float* g_alpha, float* g_faces,
                     float* g_verts, void* skip_ws, void* stream) {
  VT_CHECK_ARG(mode == 0 || mode == 1, "vt_raster_bwd: camera mode %d", mode);
  VT_CHECK_ARG(image_size > 0 && image_size <= 4096, "vt_raster_bwd: image size %d", image_size);
  if (B <= 0 || F <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned short* pcol = nullptr;
  unsigned short* prow = nullptr;
  if (skip_ws) {
    pcol = reinterpret_cast<unsigned short*>(skip_ws);
    prow = pcol + (size_t)B * image_size * (image_size + 1);
    raster_active_prefix_kernel<<<dim3(ceil_div(image_size, 8), B, 2), 256, 0, s>>>(alpha, g_alpha, image_size, pcol, prow);
    VT_CHECK_LAUNCH("vt_raster_bwd(prefix)");
  }
  cudaError_t e = cudaMemsetAsync(g_verts, 0, (size_t)B * V * 3 * sizeof(float), s);
  if (e != cudaSuccess) return cuda_fail(e, "vt_raster_bwd memset");
  raster_bwd_kernel<<<ceil_div(B * 2 * F, 128 / RB_LANES), 128, 0, s>>>(faces_ndc, face_index, alpha, g_alpha, B, 2 * F, image_size, g_faces, pcol, prow);   // RB_LANES lanes per face
  VT_CHECK_LAUNCH("vt_raster_bwd");
  raster_bwd_verts_kernel<<<ceil_div(B * 2 * F, 256), 256, 0, s>>>(g_faces, verts, faces, B, V, F, mode, K4, g_verts);
  VT_CHECK_LAUNCH("vt_raster_bwd(verts)");
  return 0;
}

}  // extern "C"
