// Post-raster composite of the reference's frame loop, on the device:
//   voicepuppet/pixrefer/infer_bfmvid.py:111     channel swap (cv2.cvtColor BGR2RGB)
//   :112-113  cv2.resize to S x S, S = int(round(res / ratio))          (8-bit bilinear, fixed point)
//   :115-121  paste into a zero canvas at (center - S // 2 - t)
//   :234-236  swap back, float32 / 255 -> inputs[0, ..., 3:6] of PixReferNet
// so that rasterized frames never leave the GPU between the rasterizer and the network input tensor.
// cv2.resize's arithmetic (OpenCV imgproc/src/resize.cpp, INTER_LINEAR on 8-bit data: 11-bit coefficients,
// HResizeLinear / VResizeLinear; exact 2x downscale = 2x2 area mean) is restated integer for integer, see
// oracle/composite.py for the derivation and tests/test_oracle_composite.py for the pin against cv2.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "common.h"
#include "launch.h"

namespace vp {

namespace {

struct AxisEntry {  // per destination index
  int s0, s1;       // source indices (already clipped)
  int c0, c1;       // 11-bit coefficients
};

std::vector<AxisEntry> axis_table(int ssize, int dsize, bool is_y) {
  std::vector<AxisEntry> t((size_t)dsize);
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    if (!is_y) {  // resize.cpp clamps the x taps and their weight; rows are only clipped when fetched
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    AxisEntry e;
    e.c0 = (int)std::nearbyintf((1.f - f) * 2048.f);  // cvRound: round half to even
    e.c1 = (int)std::nearbyintf(f * 2048.f);
    e.s0 = std::min(std::max(s, 0), ssize - 1);
    e.s1 = std::min(std::max(s + 1, 0), ssize - 1);
    t[(size_t)d] = e;
  }
  return t;
}

struct TableKey {
  int device, res, size;
  bool operator<(const TableKey& o) const {
    return device != o.device ? device < o.device : (res != o.res ? res < o.res : size < o.size);
  }
};
std::mutex g_table_mutex;
std::map<TableKey, AxisEntry*> g_tables;  // [2 * size]: x table, then y table; never freed (a few KB per size)

int tables_for(int device, int res, int size, const AxisEntry** out) {
  std::lock_guard<std::mutex> lock(g_table_mutex);
  const TableKey key{device, res, size};
  auto it = g_tables.find(key);
  if (it == g_tables.end()) {
    std::vector<AxisEntry> host = axis_table(res, size, false);
    const std::vector<AxisEntry> ty = axis_table(res, size, true);
    host.insert(host.end(), ty.begin(), ty.end());
    AxisEntry* dev = nullptr;
    VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&dev), host.size() * sizeof(AxisEntry)));
    VP_CUDA(cudaMemcpy(dev, host.data(), host.size() * sizeof(AxisEntry), cudaMemcpyHostToDevice));
    it = g_tables.emplace(key, dev).first;
  }
  *out = it->second;
  return VP_OK;
}

struct CompositeArgs {
  const unsigned char* frames;  // [T][res][res][3]
  const AxisEntry* xtab;
  const AxisEntry* ytab;
  unsigned char* canvas;        // [T][H][W][3] or NULL
  float* inputs;                // [T][H][W][in_c] or NULL; channels ch0..ch0+2 are written
  int res, size, x0, y0, H, W, in_c, ch0, swap_rb, mode;  // mode 0: bilinear, 1: 2x2 area, 2: copy
};

__global__ void __launch_bounds__(256) composite_kernel(const CompositeArgs a) {
  const int frame = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.H * a.W) return;
  const int y = p / a.W, x = p - y * a.W;
  const int dx = x - a.x0, dy = y - a.y0;
  int v[3] = {0, 0, 0};  // the resized raster's channels, in the raster's own order
  if (dx >= 0 && dx < a.size && dy >= 0 && dy < a.size) {
    const unsigned char* src = a.frames + (size_t)frame * a.res * a.res * 3;
    if (a.mode == 2) {
      const unsigned char* s = src + ((size_t)dy * a.res + dx) * 3;
      v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
    } else if (a.mode == 1) {
      const unsigned char* s0 = src + ((size_t)(2 * dy) * a.res + 2 * dx) * 3;
      const unsigned char* s1 = s0 + (size_t)a.res * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) v[k] = (s0[k] + s0[3 + k] + s1[k] + s1[3 + k] + 2) >> 2;
    } else {
      const AxisEntry ex = a.xtab[dx], ey = a.ytab[dy];
      const unsigned char* r0 = src + (size_t)ey.s0 * a.res * 3;
      const unsigned char* r1 = src + (size_t)ey.s1 * a.res * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int h0 = r0[3 * ex.s0 + k] * ex.c0 + r0[3 * ex.s1 + k] * ex.c1;   // HResizeLinear
        const int h1 = r1[3 * ex.s0 + k] * ex.c0 + r1[3 * ex.s1 + k] * ex.c1;
        v[k] = (((ey.c0 * (h0 >> 4)) >> 16) + ((ey.c1 * (h1 >> 4)) >> 16) + 2) >> 2;   // VResizeLinear (8u)
      }
    }
  }
  const size_t q = (size_t)frame * a.H * a.W + p;
  if (a.canvas) {  // render_face's return value: channels swapped (:111), uint8
    unsigned char* c = a.canvas + q * 3;
    c[0] = (unsigned char)(a.swap_rb ? v[2] : v[0]);
    c[1] = (unsigned char)v[1];
    c[2] = (unsigned char)(a.swap_rb ? v[0] : v[2]);
  }
  if (a.inputs) {  // :234-236: swapped back (the raster's own order), float32 / 255
    float* o = a.inputs + q * a.in_c + a.ch0;
    o[0] = __fdiv_rn((float)v[0], 255.0f);
    o[1] = __fdiv_rn((float)v[1], 255.0f);
    o[2] = __fdiv_rn((float)v[2], 255.0f);
  }
}

}  // namespace
}  // namespace vp

using namespace vp;

// Host only: the resize coefficient table of one axis, [dsize][4] = (s0, s1, c0, c1) per destination index, as the
// kernel uses it (lets the CPU test-suite compare the table builder with the cv2-pinned oracle over many sizes).
extern "C" int vp_composite_axis_table(int ssize, int dsize, int is_y, int* out4) {
  VP_REQUIRE(ssize > 0 && dsize > 0 && out4 != nullptr, "bad argument");
  const std::vector<AxisEntry> t = axis_table(ssize, dsize, is_y != 0);
  for (int d = 0; d < dsize; ++d) {
    out4[4 * d] = t[(size_t)d].s0;
    out4[4 * d + 1] = t[(size_t)d].s1;
    out4[4 * d + 2] = t[(size_t)d].c0;
    out4[4 * d + 3] = t[(size_t)d].c1;
  }
  return VP_OK;
}

// infer_bfmvid.py:80-82,112-121: size and top-left corner of the pasted face.
extern "C" int vp_composite_placement(int res, int center_x, int center_y, double ratio, const double* transform_params5,
                                      int* size, int* x0, int* y0) {
  VP_REQUIRE(transform_params5 && size && x0 && y0, "null argument");
  VP_REQUIRE(res > 0, "res > 0");
  ratio *= transform_params5[2];
  VP_REQUIRE(ratio > 0.0 && std::isfinite(ratio), "ratio * transform_params[2] must be positive");
  const int tx = -(int)(transform_params5[3] / ratio);
  const int ty = -(int)(transform_params5[4] / ratio);
  const int s = (int)std::nearbyint((double)res / ratio);   // Python's round(): half to even
  *size = s;
  *x0 = center_x - s / 2 - tx;
  *y0 = center_y - s / 2 - ty;
  return VP_OK;
}

extern "C" int vp_composite_dev(const unsigned char* frames_dev, int nframes, int res, int size, int x0, int y0,
                                int canvas_h, int canvas_w, unsigned char* canvas_dev, int swap_rb, float* inputs_dev,
                                int in_channels, int channel_offset, int device, void* stream) {
  VP_REQUIRE(nframes >= 0 && res > 0 && size > 0 && canvas_h > 0 && canvas_w > 0, "bad size");
  VP_REQUIRE(nframes == 0 || frames_dev != nullptr, "null frames");
  VP_REQUIRE(canvas_dev != nullptr || inputs_dev != nullptr, "no output requested");
  VP_REQUIRE(inputs_dev == nullptr || (channel_offset >= 0 && channel_offset + 3 <= in_channels), "bad channel range");
  // numpy's slice assignment raises when the face does not fit the canvas (infer_bfmvid.py:121)
  VP_REQUIRE(x0 >= 0 && y0 >= 0 && x0 + size <= canvas_w && y0 + size <= canvas_h, "the resized face does not fit the canvas");
  VP_REQUIRE((long long)canvas_h * canvas_w < (1ll << 31), "canvas too large");
  if (nframes == 0) return VP_OK;
  VP_CUDA(cudaSetDevice(device));
  CompositeArgs a;
  a.frames = frames_dev;
  a.xtab = a.ytab = nullptr;
  a.mode = (size == res) ? 2 : ((res == 2 * size) ? 1 : 0);
  if (a.mode == 0) {
    const AxisEntry* tab = nullptr;
    VP_TRY(tables_for(device, res, size, &tab));
    a.xtab = tab;
    a.ytab = tab + size;
  }
  a.canvas = canvas_dev;
  a.inputs = inputs_dev;
  a.res = res;
  a.size = size;
  a.x0 = x0;
  a.y0 = y0;
  a.H = canvas_h;
  a.W = canvas_w;
  a.in_c = in_channels;
  a.ch0 = channel_offset;
  a.swap_rb = swap_rb;
  dim3 grid((unsigned)(((size_t)canvas_h * canvas_w + 255) / 256), (unsigned)nframes);
  composite_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  VP_LAUNCH_CHECK();
  return VP_OK;
}
