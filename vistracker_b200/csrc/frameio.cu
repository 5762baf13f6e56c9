// Test-time frame preparation on the device: what `TestDataTriplane.get_item` (data/testdata_triplane.py:42-74)
// does per frame on DataLoader workers -- crop centre from the two masks (data/base_data.py:139-171), 1200^2 crop with zero padding (:204-232),
// cv2.resize to 512^2 (:234-246), / 255, background masking and channel stacking (:252-265, data/train_data.py:143-162), triplane png / 255
// (testdata_triplane.py:52-56, 77-80) -- as two launches over a batch of decoded uint8 frames that already sit in HBM:
//
//   vt_mask_bbox            bounding box of (person + object, uint8 wrap-around) > 127 per frame, and (min + max) // 2
//   vt_prepare_image_crop   one thread per output pixel: 4 source pixels x (3 + 1 + 1) channels -> OpenCV's 8-bit fixed-point bilinear
//                           (11-bit weights from host-built tables, int32 horizontal pass, `>> 4`, `>> 16`, `+ 2 >> 2` vertical pass),
//                           (double) v / 255 -> float, RGB zeroed where neither mask exceeds 0.5, [B][8][S][S] written once
//
// cv2.resize / cv2.findContours themselves are not available offline: the arithmetic restates OpenCV's portable 8-bit path (PARITY UNPINNED for
// those two steps, pinned for the crop / compose / layout).  HBM-bound: 5 B read per source pixel touched, 32 B written per
// output pixel.
#include "common.cuh"
#include "vt_internal.h"

namespace vt {

__global__ void bbox_init_kernel(int* __restrict__ bbox, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) { bbox[b * 4 + 0] = 50000; bbox[b * 4 + 1] = 50000; bbox[b * 4 + 2] = -100; bbox[b * 4 + 3] = -100; }   // base_data.py:152
}

// grid (rows / rows_per_block, B)
__global__ void __launch_bounds__(256) mask_bbox_kernel(const unsigned char* __restrict__ person, const unsigned char* __restrict__ obj, int H, int W, int thres,
                                                        int* __restrict__ bbox) {
  const int b = blockIdx.y;
  const unsigned char* p = person + (size_t)b * H * W;
  const unsigned char* o = obj + (size_t)b * H * W;
  int xmin = 50000, ymin = 50000, xmax = -100, ymax = -100;
  const int y0 = blockIdx.x * 8;
  for (int y = y0; y < min(y0 + 8, H); ++y) {
    for (int x = threadIdx.x; x < W; x += 256) {
      const unsigned char s = (unsigned char)(p[(size_t)y * W + x] + o[(size_t)y * W + x]);      // numpy uint8 `+=` wraps
      if ((int)s > thres) { xmin = min(xmin, x); xmax = max(xmax, x + 1); ymin = min(ymin, y); ymax = max(ymax, y + 1); }
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, d)); ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, d));
    xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, d)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, d));
  }
  if ((threadIdx.x & 31) == 0 && xmax > -100) {
    atomicMin(bbox + b * 4 + 0, xmin); atomicMin(bbox + b * 4 + 1, ymin); atomicMax(bbox + b * 4 + 2, xmax); atomicMax(bbox + b * 4 + 3, ymax);
  }
}

__global__ void bbox_center_kernel(const int* __restrict__ bbox, int B, float* __restrict__ center) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // (bmin + bmax) // 2 with numpy's floor division
  const int sx = bbox[b * 4 + 0] + bbox[b * 4 + 2], sy = bbox[b * 4 + 1] + bbox[b * 4 + 3];
  center[b * 2 + 0] = (float)(sx >= 0 ? sx / 2 : -((-sx + 1) / 2));
  center[b * 2 + 1] = (float)(sy >= 0 ? sy / 2 : -((-sy + 1) / 2));
}

struct CropArgs {
  const unsigned char* rgb; const unsigned char* person; const unsigned char* obj; const unsigned char* tri;
  const float* center; const int* tab;      // tab: [3][S] = source index, weight of it, weight of the next pixel (11-bit fixed point)
  float* out; int B, H, W, crop, S, C;
};

__device__ __forceinline__ int cv_vresize(int h0, int h1, int b0, int b1) {
  const int v = ((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
  return min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256) prepare_image_crop_kernel(CropArgs a) {
  const int S = a.S;
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (dx >= S || dy >= S) return;
  const int cx = (int)a.center[b * 2 + 0], cy = (int)a.center[b * 2 + 1];
  const int tlx = cx - a.crop / 2, tly = cy - a.crop / 2, brx = cx + a.crop / 2, bry = cy + a.crop / 2;    // crop is even: np.round is exact
  const int X1 = max(0, tlx), Y1 = max(0, tly), X2 = min(a.W - 1, brx), Y2 = min(a.H - 1, bry);              // base_data.py:218-221
  const int x0 = a.tab[dx], x1 = min(x0 + 1, a.crop - 1), ax0 = a.tab[S + dx], ax1 = a.tab[2 * S + dx];
  const int y0 = a.tab[dy], y1 = min(y0 + 1, a.crop - 1), by0 = a.tab[S + dy], by1 = a.tab[2 * S + dy];
  const int xs[2] = {tlx + x0, tlx + x1}, ys[2] = {tly + y0, tly + y1};
  int px[2][2][5];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool in = xs[i] >= X1 && xs[i] < X2 && ys[j] >= Y1 && ys[j] < Y2;
      const size_t o = ((size_t)b * a.H + (in ? ys[j] : 0)) * a.W + (in ? xs[i] : 0);
      px[j][i][0] = in ? a.rgb[o * 3 + 0] : 0; px[j][i][1] = in ? a.rgb[o * 3 + 1] : 0; px[j][i][2] = in ? a.rgb[o * 3 + 2] : 0;
      px[j][i][3] = in ? a.person[o] : 0; px[j][i][4] = in ? a.obj[o] : 0;
    }
  int v[5];
#pragma unroll
  for (int c = 0; c < 5; ++c)
    v[c] = cv_vresize(px[0][0][c] * ax0 + px[0][1][c] * ax1, px[1][0][c] * ax0 + px[1][1][c] * ax1, by0, by1);
  const bool fg = v[3] > 127 || v[4] > 127;                     // (v / 255.) > 0.5
  const size_t plane = (size_t)S * S, o = (size_t)b * a.C * plane + (size_t)dy * S + dx;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const float f = (float)((double)v[c] / 255.0);
    a.out[o + c * plane] = (c < 3 && !fg) ? 0.f : f;
  }
  if (a.tri)
#pragma unroll
    for (int k = 0; k < 3; ++k) a.out[o + (5 + k) * plane] = (float)((double)a.tri[(((size_t)b * S + dy) * S + dx) * 3 + k] / 255.0);
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_mask_bbox(const unsigned char* person, const unsigned char* obj, int B, int H, int W, int thres, int* bbox, float* center, void* stream) {
  VT_CHECK_ARG(person && obj && bbox, "vt_mask_bbox: null pointer");
  VT_CHECK_ARG(H > 0 && W > 0 && W < 50000 && H < 50000, "vt_mask_bbox: %d x %d masks", H, W);
  if (B <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  bbox_init_kernel<<<ceil_div(B, 128), 128, 0, s>>>(bbox, B);
  mask_bbox_kernel<<<dim3(ceil_div(H, 8), B), 256, 0, s>>>(person, obj, H, W, thres, bbox);
  if (center) bbox_center_kernel<<<ceil_div(B, 128), 128, 0, s>>>(bbox, B, center);
  VT_CHECK_LAUNCH("vt_mask_bbox");
  return 0;
}

int vt_prepare_image_crop(const unsigned char* rgb, const unsigned char* person, const unsigned char* obj, const unsigned char* triplane, int B, int H, int W,
                          const float* crop_center, int crop_size, int net_size, const int* resize_tab, float* images, int channels, void* stream) {
  VT_CHECK_ARG(rgb && person && obj && crop_center && resize_tab && images, "vt_prepare_image_crop: null pointer");
  VT_CHECK_ARG(crop_size > 0 && crop_size % 2 == 0 && net_size > 0, "vt_prepare_image_crop: crop %d (even) -> %d", crop_size, net_size);
  VT_CHECK_ARG(channels >= (triplane ? 8 : 5), "vt_prepare_image_crop: %d output channels", channels);
  if (B <= 0) return 0;
  CropArgs a{rgb, person, obj, triplane, crop_center, resize_tab, images, B, H, W, crop_size, net_size, channels};
  prepare_image_crop_kernel<<<dim3(ceil_div(net_size, 32), ceil_div(net_size, 8), B), 256, 0, (cudaStream_t)stream>>>(a);
  VT_CHECK_LAUNCH("vt_prepare_image_crop");
  return 0;
}

}  // extern "C"
