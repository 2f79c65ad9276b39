// TEST INFRASTRUCTURE -- not part of the product path.
//
// C-ABI doorway onto the *unmodified* reference rasterizer.  This file holds no
// algorithm: it is compiled together with /root/reference/utils/cython/mesh_core.cpp
// (sources stay where they are; see oracle/build_ref.py) and only forwards to the
// C++-linkage functions declared in the reference's mesh_core.h:53-77, so that
// ctypes can call them without the GIL (needed for the multi-threaded CPU baseline).
#include "mesh_core.h"

extern "C" {

void ref_render_colors_core(unsigned char* image, unsigned char* face_mask, float* vertices,
                            int* triangles, float* colors, float* depth_buffer,
                            int ntri, int h, int w, int c) {
  _render_colors_core(image, face_mask, vertices, triangles, colors, depth_buffer, ntri, h, w, c);
}

void ref_rasterize_triangles_core(float* vertices, int* triangles, float* depth_buffer,
                                  int* triangle_buffer, float* barycentric_weight,
                                  int nver, int ntri, int h, int w) {
  _rasterize_triangles_core(vertices, triangles, depth_buffer, triangle_buffer,
                            barycentric_weight, nver, ntri, h, w);
}

void ref_render_texture_core(float* image, float* vertices, int* triangles, float* texture,
                             float* tex_coords, int* tex_triangles, float* depth_buffer,
                             int nver, int tex_nver, int ntri, int h, int w, int c,
                             int tex_h, int tex_w, int tex_c, int mapping_type) {
  _render_texture_core(image, vertices, triangles, texture, tex_coords, tex_triangles,
                       depth_buffer, nver, tex_nver, ntri, h, w, c, tex_h, tex_w, tex_c,
                       mapping_type);
}

void ref_get_normal_core(float* normal, float* tri_normal, int* triangles, int ntri) {
  _get_normal_core(normal, tri_normal, triangles, ntri);
}

}  // extern "C"
