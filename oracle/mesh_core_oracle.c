/*
 * TEST INFRASTRUCTURE -- CPU oracle for the rasterizer half of the hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 *
 * What it restates: the reference's sequential z-buffer loops
 *   _render_colors_core        /root/reference/utils/cython/mesh_core.cpp:169-231
 *   _rasterize_triangles_core  /root/reference/utils/cython/mesh_core.cpp:108-166
 *   isPointInTri / get_point_weight                       mesh_core.cpp:23-82
 *   point::dot, operator-                                 mesh_core.h:19-30
 * in the ORDER-INDEPENDENT form the GPU uses (SURVEY.md section 8c identities 1 and 2):
 * a pixel's winner is the candidate with the largest depth that is strictly greater
 * than the caller's initial depth, ties going to the lowest triangle index.  That is
 * what "for i in order: if (d > depth[p]) write" converges to, so image / mask / depth /
 * triangle / weights must equal the reference bit for bit -- tests/test_oracle_raster.py
 * pins this against oracle/_ref (the compiled reference) and tests/golden/.
 * `reverse` walks the triangles backwards to demonstrate the order independence.
 *
 * Arithmetic contract (why this is bit-exact): every float32 +,-,*,/ is individually
 * rounded, in the reference's association order; build with -ffp-contract=off and no
 * -ffast-math / -mfma.  float->int follows x86 cvttss2si (out of range / NaN -> INT_MIN),
 * which is what the compiled reference does for its (int) casts.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float ax, ay;           /* corner 0 */
  float e0x, e0y;         /* corner2 - corner0   (mesh_core.cpp:27) */
  float e1x, e1y;         /* corner1 - corner0   (mesh_core.cpp:28) */
  float d00, d01, d11;    /* mesh_core.cpp:32-35 */
  float inv;              /* mesh_core.cpp:39-43 */
  float z0, z1, z2;
  int x_lo, x_hi, y_lo, y_hi;
  int live;
} tri_setup;

static int trunc_x86(float f) {
  if (f >= -2147483648.0f && f < 2147483648.0f) return (int)f;
  return INT_MIN;
}

/* std::min / std::max as libstdc++ defines them: (b < a) ? b : a  /  (a < b) ? b : a */
static float lo2(float a, float b) { return (b < a) ? b : a; }
static float hi2(float a, float b) { return (a < b) ? b : a; }
static int ilo2(int a, int b) { return (b < a) ? b : a; }
static int ihi2(int a, int b) { return (a < b) ? b : a; }

static void setup_triangle(tri_setup* s, const float* vertices, const int* tri, int h, int w) {
  const float* q0 = vertices + 3 * (size_t)tri[0];
  const float* q1 = vertices + 3 * (size_t)tri[1];
  const float* q2 = vertices + 3 * (size_t)tri[2];
  /* bounding box, mesh_core.cpp:132-136 / 194-198 */
  s->x_lo = ihi2(trunc_x86(ceilf(lo2(q0[0], lo2(q1[0], q2[0])))), 0);
  s->x_hi = ilo2(trunc_x86(floorf(hi2(q0[0], hi2(q1[0], q2[0])))), w - 1);
  s->y_lo = ihi2(trunc_x86(ceilf(lo2(q0[1], lo2(q1[1], q2[1])))), 0);
  s->y_hi = ilo2(trunc_x86(floorf(hi2(q0[1], hi2(q1[1], q2[1])))), h - 1);
  s->live = !(s->x_hi < s->x_lo || s->y_hi < s->y_lo);
  s->ax = q0[0]; s->ay = q0[1];
  s->e0x = q2[0] - q0[0]; s->e0y = q2[1] - q0[1];
  s->e1x = q1[0] - q0[0]; s->e1y = q1[1] - q0[1];
  s->d00 = s->e0x * s->e0x + s->e0y * s->e0y;
  s->d01 = s->e0x * s->e1x + s->e0y * s->e1y;
  s->d11 = s->e1x * s->e1x + s->e1y * s->e1y;
  {
    float den = s->d00 * s->d11 - s->d01 * s->d01;
    s->inv = (den == 0) ? 0.0f : 1 / den;
  }
  s->z0 = q0[2]; s->z1 = q1[2]; s->z2 = q2[2];
}

/* barycentric (u, v) of integer pixel (x, y); mesh_core.cpp:29,34,36,45-46 */
static void pixel_uv(const tri_setup* s, int x, int y, float* u, float* v) {
  float px = (float)x - s->ax, py = (float)y - s->ay;
  float d02 = s->e0x * px + s->e0y * py;
  float d12 = s->e1x * px + s->e1y * py;
  *u = (s->d11 * d02 - s->d01 * d12) * s->inv;
  *v = (s->d00 * d12 - s->d01 * d02) * s->inv;
}

static int uv_inside(float u, float v) { return (u >= 0) && (v >= 0) && (u + v < 1); }

/* order-preserving map float -> uint32 with -0 == +0; caller filters NaN */
static uint32_t depth_bits(float d) {
  uint32_t b;
  if (d == 0) d = 0.0f;
  memcpy(&b, &d, 4);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

#define NO_TRI 0xFFFFFFFFu

static uint64_t* keys_from_depth(const float* depth, size_t n) {
  uint64_t* keys = (uint64_t*)malloc(n * sizeof(uint64_t));
  size_t p;
  if (!keys) return NULL;
  for (p = 0; p < n; p++) {
    float d = depth[p];
    /* a NaN in the caller's buffer makes every "d > depth" test false: nothing may win */
    keys[p] = (d != d) ? ~(uint64_t)0 : (((uint64_t)depth_bits(d) << 32) | NO_TRI);
  }
  return keys;
}

static void offer(uint64_t* keys, size_t p, float d, int tri) {
  uint64_t k;
  if (d != d) return;
  k = ((uint64_t)depth_bits(d) << 32) | (uint64_t)(NO_TRI - 1u - (uint32_t)tri);
  if (k > keys[p]) keys[p] = k;
}

static int key_winner(uint64_t k) {
  uint32_t low = (uint32_t)k;
  if (low == NO_TRI) return -1;
  return (int)(NO_TRI - 1u - low);
}

/* flat depth of mesh_core.cpp:204: (z0+z1+z2)/3 evaluated left to right in float */
static float flat_depth(const tri_setup* s) { return ((s->z0 + s->z1) + s->z2) / 3; }

/*
 * render_colors, order independent.  triangle_out (may be NULL) receives the implied
 * winner per pixel (-1 = untouched), which the reference never materialises.
 */
int vpo_render_colors(unsigned char* image, unsigned char* face_mask, const float* vertices,
                      const int* triangles, const float* colors, float* depth_buffer,
                      int* triangle_out, int ntri, int h, int w, int c, int reverse) {
  size_t npix = (size_t)h * (size_t)w, p;
  uint64_t* keys = keys_from_depth(depth_buffer, npix);
  int n, x, y, k;
  if (!keys) return -1;
  for (n = 0; n < ntri; n++) {
    int i = reverse ? (ntri - 1 - n) : n;
    tri_setup s;
    float d;
    setup_triangle(&s, vertices, triangles + 3 * (size_t)i, h, w);
    if (!s.live) continue;
    d = flat_depth(&s);
    for (y = s.y_lo; y <= s.y_hi; y++)
      for (x = s.x_lo; x <= s.x_hi; x++) {
        float u, v;
        pixel_uv(&s, x, y, &u, &v);
        if (uv_inside(u, v)) offer(keys, (size_t)y * w + x, d, i);
      }
  }
  for (p = 0; p < npix; p++) {
    int i = key_winner(keys[p]);
    if (triangle_out) triangle_out[p] = i;
    if (i < 0) continue;
    {
      const int* t = triangles + 3 * (size_t)i;
      tri_setup s;
      setup_triangle(&s, vertices, t, h, w);
      for (k = 0; k < c; k++) {
        /* mesh_core.cpp:215-221: float sum -> (int) -> integer /3 -> float -> unsigned char */
        float sum = (colors[(size_t)c * t[0] + k] + colors[(size_t)c * t[1] + k]) + colors[(size_t)c * t[2] + k];
        float pc = (float)(trunc_x86(sum) / 3);
        image[p * c + k] = (unsigned char)trunc_x86(pc);
      }
      face_mask[p] = 255;
      depth_buffer[p] = flat_depth(&s);
    }
  }
  free(keys);
  return 0;
}

/* weights of mesh_core.cpp:79-81 and the interpolated depth of :151 */
static float weights_and_depth(const tri_setup* s, float u, float v, float* wgt) {
  wgt[0] = 1 - u - v;
  wgt[1] = v;
  wgt[2] = u;
  return wgt[0] * s->z0 + wgt[1] * s->z1 + wgt[2] * s->z2;
}

int vpo_rasterize_triangles(const float* vertices, const int* triangles, float* depth_buffer,
                            int* triangle_buffer, float* barycentric_weight,
                            int nver, int ntri, int h, int w, int reverse) {
  size_t npix = (size_t)h * (size_t)w, p;
  uint64_t* keys = keys_from_depth(depth_buffer, npix);
  int n, x, y;
  (void)nver;
  if (!keys) return -1;
  for (n = 0; n < ntri; n++) {
    int i = reverse ? (ntri - 1 - n) : n;
    tri_setup s;
    setup_triangle(&s, vertices, triangles + 3 * (size_t)i, h, w);
    if (!s.live) continue;
    for (y = s.y_lo; y <= s.y_hi; y++)
      for (x = s.x_lo; x <= s.x_hi; x++) {
        float u, v, wgt[3];
        float fx = (float)x, fy = (float)y;
        pixel_uv(&s, x, y, &u, &v);
        /* mesh_core.cpp:148: the 2-pixel frame always qualifies */
        if (fx < 2 || fx > w - 3 || fy < 2 || fy > h - 3 || uv_inside(u, v))
          offer(keys, (size_t)y * w + x, weights_and_depth(&s, u, v, wgt), i);
      }
  }
  for (p = 0; p < npix; p++) {
    int i = key_winner(keys[p]);
    if (i < 0) continue;
    {
      tri_setup s;
      float u, v, wgt[3];
      setup_triangle(&s, vertices, triangles + 3 * (size_t)i, h, w);
      pixel_uv(&s, (int)(p % (size_t)w), (int)(p / (size_t)w), &u, &v);
      depth_buffer[p] = weights_and_depth(&s, u, v, wgt);
      triangle_buffer[p] = i;
      barycentric_weight[3 * p + 0] = wgt[0];
      barycentric_weight[3 * p + 1] = wgt[1];
      barycentric_weight[3 * p + 2] = wgt[2];
    }
  }
  free(keys);
  return 0;
}

/*
 * Per pixel: how many pixels have their two largest candidate depths within `ulps`
 * float32 steps of each other (render_colors flat depth).  north_star lets the
 * end-to-end triangle-id comparison exclude exactly these near-tie pixels, so the
 * oracle has to be able to name them.  near_tie[p] is set to 1 for such pixels.
 */
int vpo_render_colors_near_ties(const float* vertices, const int* triangles, int ntri, int h, int w,
                                int ulps, unsigned char* near_tie) {
  size_t npix = (size_t)h * (size_t)w, p;
  uint32_t* best = (uint32_t*)calloc(npix, sizeof(uint32_t));
  uint32_t* second = (uint32_t*)calloc(npix, sizeof(uint32_t));
  int i, x, y, count = 0;
  if (!best || !second) { free(best); free(second); return -1; }
  for (i = 0; i < ntri; i++) {
    tri_setup s;
    float d;
    uint32_t b;
    setup_triangle(&s, vertices, triangles + 3 * (size_t)i, h, w);
    if (!s.live) continue;
    d = flat_depth(&s);
    if (d != d) continue;
    b = depth_bits(d);
    for (y = s.y_lo; y <= s.y_hi; y++)
      for (x = s.x_lo; x <= s.x_hi; x++) {
        float u, v;
        size_t q = (size_t)y * w + x;
        pixel_uv(&s, x, y, &u, &v);
        if (!uv_inside(u, v)) continue;
        if (b > best[q]) { second[q] = best[q]; best[q] = b; }
        else if (b > second[q]) second[q] = b;
      }
  }
  for (p = 0; p < npix; p++) {
    int tie = second[p] != 0 && (best[p] - second[p]) <= (uint32_t)ulps;
    near_tie[p] = (unsigned char)tie;
    count += tie;
  }
  free(best);
  free(second);
  return count;
}

/*
 * render_texture, order independent: _render_texture_core, mesh_core.cpp:234-333.
 * The z-buffer decision is rasterize_triangles' (border rule :290, interpolated depth :293, strict '>'
 * :295), so the winner per pixel is found the same way; the winner's texel is then recomputed with the
 * reference's float32 expressions:
 *   tex_p = tex_p0*w0 + tex_p1*w1 + tex_p2*w2  (point::operator*, operator+, mesh_core.h:32-46; left to right)
 *   clamp to [0, tex_w-1] x [0, tex_h-1] with std::min / std::max (:302-303)
 *   nearest: texture[round(y)][round(x)] (:311); bilinear: ul*(1-xd)*(1-yd) + ur*xd*(1-yd) + dl*(1-xd)*yd +
 *   dr*xd*yd, left to right (:315-320).
 * Reference quirk kept: the texture y coordinate is read with the MESH vertex index, stride 3 (:270-272).
 * The image is written only where a triangle wins; elsewhere the caller's values stay.
 */
int vpo_render_texture(float* image, const float* vertices, const int* triangles, const float* texture,
                       const float* tex_coords, const int* tex_triangles, float* depth_buffer,
                       int nver, int tex_nver, int ntri, int h, int w, int c, int tex_h, int tex_w, int tex_c,
                       int mapping_type, int reverse) {
  size_t npix = (size_t)h * (size_t)w, p;
  uint64_t* keys = keys_from_depth(depth_buffer, npix);
  int n, x, y, k;
  (void)nver; (void)tex_nver;
  if (!keys) return -1;
  for (n = 0; n < ntri; n++) {
    int i = reverse ? (ntri - 1 - n) : n;
    tri_setup s;
    setup_triangle(&s, vertices, triangles + 3 * (size_t)i, h, w);
    if (!s.live) continue;
    for (y = s.y_lo; y <= s.y_hi; y++)
      for (x = s.x_lo; x <= s.x_hi; x++) {
        float u, v, wgt[3];
        float fx = (float)x, fy = (float)y;
        pixel_uv(&s, x, y, &u, &v);
        if (fx < 2 || fx > w - 3 || fy < 2 || fy > h - 3 || uv_inside(u, v))
          offer(keys, (size_t)y * w + x, weights_and_depth(&s, u, v, wgt), i);
      }
  }
  for (p = 0; p < npix; p++) {
    int i = key_winner(keys[p]);
    if (i < 0) continue;
    {
      const int* t = triangles + 3 * (size_t)i;
      const int* tt = tex_triangles + 3 * (size_t)i;
      tri_setup s;
      float u, v, wgt[3], tx, ty, xd, yd;
      setup_triangle(&s, vertices, t, h, w);
      pixel_uv(&s, (int)(p % (size_t)w), (int)(p / (size_t)w), &u, &v);
      depth_buffer[p] = weights_and_depth(&s, u, v, wgt);
      /* x from the texture triangle's vertex, y from the MESH triangle's vertex (:270-272) */
      tx = (wgt[0] * tex_coords[3 * (size_t)tt[0]] + wgt[1] * tex_coords[3 * (size_t)tt[1]]) +
           wgt[2] * tex_coords[3 * (size_t)tt[2]];
      ty = (wgt[0] * tex_coords[3 * (size_t)t[0] + 1] + wgt[1] * tex_coords[3 * (size_t)t[1] + 1]) +
           wgt[2] * tex_coords[3 * (size_t)t[2] + 1];
      tx = hi2(lo2(tx, (float)(tex_w - 1)), 0.0f);
      ty = hi2(lo2(ty, (float)(tex_h - 1)), 0.0f);
      yd = ty - floorf(ty);
      xd = tx - floorf(tx);
      for (k = 0; k < c; k++) {
        if (mapping_type == 0) {
          image[p * c + k] = texture[(size_t)trunc_x86(roundf(ty)) * tex_w * tex_c + (size_t)trunc_x86(roundf(tx)) * tex_c + k];
        } else {
          size_t y0 = (size_t)trunc_x86(floorf(ty)), y1 = (size_t)trunc_x86(ceilf(ty));
          size_t x0 = (size_t)trunc_x86(floorf(tx)), x1 = (size_t)trunc_x86(ceilf(tx));
          float ul = texture[y0 * tex_w * tex_c + x0 * tex_c + k];
          float ur = texture[y0 * tex_w * tex_c + x1 * tex_c + k];
          float dl = texture[y1 * tex_w * tex_c + x0 * tex_c + k];
          float dr = texture[y1 * tex_w * tex_c + x1 * tex_c + k];
          image[p * c + k] = ((ul * (1 - xd) * (1 - yd) + ur * xd * (1 - yd)) + dl * (1 - xd) * yd) + dr * xd * yd;
        }
      }
    }
  }
  free(keys);
  return 0;
}

/*
 * get_normal, gather form: _get_normal_core, mesh_core.cpp:85-105 adds tri_normal[i] to the three corner
 * vertices of triangle i, walking i upwards; per vertex that is a left-to-right float32 sum over its
 * (triangle, corner) incidences in ascending order, starting from the caller's value.  The restatement
 * builds the incidence lists with a stable counting sort and sums each vertex on its own (what the GPU does).
 * Returns -1 on allocation failure, -2 when a triangle names a vertex outside [0, nver).
 */
int vpo_get_normal(float* normal, const float* tri_normal, const int* triangles, int nver, int ntri) {
  size_t ninc = 3 * (size_t)ntri, q;
  int* start = (int*)calloc((size_t)nver + 1, sizeof(int));
  int* fill;
  int* inc;
  int v;
  if (!start) return -1;
  for (q = 0; q < ninc; q++) {
    if (triangles[q] < 0 || triangles[q] >= nver) { free(start); return -2; }
    start[triangles[q] + 1]++;
  }
  for (v = 0; v < nver; v++) start[v + 1] += start[v];
  fill = (int*)malloc(((size_t)nver + 1) * sizeof(int));
  inc = (int*)malloc((ninc ? ninc : 1) * sizeof(int));
  if (!fill || !inc) { free(start); free(fill); free(inc); return -1; }
  memcpy(fill, start, ((size_t)nver + 1) * sizeof(int));
  for (q = 0; q < ninc; q++) inc[fill[triangles[q]]++] = (int)(q / 3);   /* stable: ascending (triangle, corner) */
  for (v = 0; v < nver; v++) {
    int j, k;
    for (j = start[v]; j < start[v + 1]; j++)
      for (k = 0; k < 3; k++) normal[3 * (size_t)v + k] = normal[3 * (size_t)v + k] + tri_normal[3 * (size_t)inc[j] + k];
  }
  free(start); free(fill); free(inc);
  return 0;
}
