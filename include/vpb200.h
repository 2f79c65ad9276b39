/*
 * libvpb200 -- C ABI of the B200-native BFM reconstruction + rasterization path.
 *
 * Drop-in scope (reference taylorlu/voicepuppet, paths relative to its root):
 *   utils/cython/mesh_core_cython.pyx:40-99   render_colors_core, rasterize_triangles_core,
 *                                             render_texture_core, get_normal_core
 *   utils/reconstruct_mesh.py:5-223           Shape/Texture formation, Compute_norm,
 *                                             Projection_layer, Illumination_layer,
 *                                             Reconstruction, Reconstruction_rotation
 *   voicepuppet/pixrefer/infer_bfmvid.py:76-122,231-243   the per-frame render loop
 *
 * Conventions
 *   - every function returns 0 on success; on failure a non-zero code, and
 *     vp_last_error() (thread local) describes it.  There is no CPU fallback: without a
 *     CUDA device every compute entry point fails with VP_ERR_CUDA.
 *   - plain pointers + sizes only.  Pointers are HOST pointers unless the name ends in
 *     _dev or the parameter says "device".  The callee never keeps or frees caller memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Host-pointer
 *     entry points synchronise before returning; _dev entry points only enqueue.
 *   - indices are 0-based int32 (the Python shim converts the model's 1-based MATLAB
 *     doubles exactly as the reference does with (x - 1).astype(np.int32)).
 */
#ifndef VPB200_H_
#define VPB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_OK 0
#define VP_ERR_ARG 1
#define VP_ERR_CUDA 2
#define VP_ERR_STATE 3

#define VP_N_ID 80
#define VP_N_EX 64
#define VP_N_TEX 80
#define VP_N_GAMMA 27
#define VP_RING 8 /* point_buf slots per vertex */

/* bit flags for vp_model_create(float64_mask): which float arrays are float64 */
#define VP_F64_MEANSHAPE 1
#define VP_F64_IDBASE 2
#define VP_F64_EXBASE 4
#define VP_F64_MEANTEX 8
#define VP_F64_TEXBASE 16

typedef struct vp_model vp_model; /* opaque: device-resident BFM model + workspaces */

const char* vp_last_error(void);
int vp_version(void);
/* number of visible CUDA devices, or a negative VP_ERR code */
int vp_device_count(void);

/* Page-locked host memory for the host-pointer entry points: copies from/to such buffers are
 * truly asynchronous, so vp_render_sequence can overlap the device->host drain of one chunk
 * with the rendering of the next.  Pageable memory works too (slower). */
int vp_host_alloc(void** out, size_t bytes);
int vp_host_free(void* p);

/* Peer memory (CUDA IPC over NVLink), for the multi-GPU gather without a copy: rank 0 exports the
 * buffer that receives every rank's frames, the other ranks open it and pass `base + offset + their
 * slice` as image_dev to vp_render_sequence_dev: the resolve kernel's stores land in rank 0's HBM.
 *   vp_ipc_export: handle64 = cudaIpcMemHandle_t of the allocation containing dev_ptr, *offset = dev_ptr - base
 *   vp_ipc_open:   maps the allocation into this process on `device` (peer access enabled lazily) */
int vp_ipc_export(const void* dev_ptr, unsigned char* handle64, unsigned long long* offset);
int vp_ipc_open(const unsigned char* handle64, int device, void** base_out);
int vp_ipc_close(void* base);
/* Completion flags in (peer-mapped) device memory: vp_peer_signal publishes `value` in *flag_dev with
 * system-scope release semantics after everything enqueued on `stream` before it; vp_peer_wait makes
 * `stream` wait until flags_dev[0..n) have all reached `value` (n <= 32; wrap-around safe compare;
 * traps after a bounded spin instead of hanging).  Use an increasing step counter as `value`. */
int vp_peer_signal(unsigned int* flag_dev, unsigned int value, void* stream);
int vp_peer_wait(const unsigned int* flags_dev, int n, unsigned int value, void* stream);
/* A wait gives up after VPB200_PEER_TIMEOUT_S seconds (default 60) instead of hanging the GPU; this returns how many
 * waits on the current device gave up since the library was loaded (check it after synchronising), < 0 on error. */
int vp_peer_timeouts(void);
/* Asynchronous device-to-device copy (copy engine), e.g. finished frames -> the peer-mapped buffer. */
int vp_copy_async(void* dst_dev, const void* src_dev, size_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * mesh_core_cython replacements: caller-initialised buffers, mutated in place.
 * --------------------------------------------------------------------------------------- */

/* render_colors_core (mesh_core_cython.pyx:64-78 -> mesh_core.cpp:169-231).
 * image[h*w*c] u8, face_mask[h*w] u8, vertices[3*nver] f32, triangles[3*ntri] i32,
 * colors[c*nver] f32, depth_buffer[h*w] f32.  `nver` is new (the reference trusts the
 * indices); it sizes the host->device copy.  triangle_id (may be NULL) receives the winning
 * triangle per pixel, -1 where nothing was drawn (the reference never materialises it). */
int vp_render_colors_core(unsigned char* image, unsigned char* face_mask, const float* vertices,
                          const int* triangles, const float* colors, float* depth_buffer,
                          int* triangle_id, int nver, int ntri, int h, int w, int c);

/* rasterize_triangles_core (mesh_core_cython.pyx:49-62 -> mesh_core.cpp:108-166).
 * vertices[nver*3], triangles[ntri*3], depth_buffer[h*w], triangle_buffer[h*w],
 * barycentric_weight[h*w*3]. */
int vp_rasterize_triangles_core(const float* vertices, const int* triangles, float* depth_buffer,
                                int* triangle_buffer, float* barycentric_weight,
                                int nver, int ntri, int h, int w);

/* render_texture_core (mesh_core_cython.pyx:80-99 -> mesh_core.cpp:234-333): z-buffer decision of
 * rasterize_triangles_core, the winner's texel sampled nearest (mapping_type 0) or bilinearly.
 * image[h*w*c] f32 (written only where a triangle wins), vertices[nver*3], triangles[ntri*3],
 * texture[tex_h*tex_w*tex_c] f32, tex_coords[tex_nver*3] (x, y, unused), tex_triangles[ntri*3],
 * depth_buffer[h*w].  Reference quirk kept: the texture y coordinate is read with the MESH vertex index
 * (mesh_core.cpp:270-272), so tex_coords needs a row for every mesh index used.  Unlike the reference,
 * out-of-range indices are an error (VP_ERR_ARG) instead of an out-of-bounds read. */
int vp_render_texture_core(float* image, const float* vertices, const int* triangles, const float* texture,
                           const float* tex_coords, const int* tex_triangles, float* depth_buffer,
                           int nver, int tex_nver, int ntri, int h, int w, int c,
                           int tex_h, int tex_w, int tex_c, int mapping_type);

/* get_normal_core (mesh_core_cython.pyx:40-47 -> mesh_core.cpp:85-105): normal[v] += tri_normal[i] for the
 * three corners of every triangle i, accumulated in ascending triangle order in float32 -- reproduced
 * bit-exactly by an ordered per-vertex gather.  normal[nver*3] in/out, tri_normal[ntri*3],
 * triangles[ntri*3].  `nver` is new (the reference trusts the indices). */
int vp_get_normal_core(float* normal, const float* tri_normal, const int* triangles, int nver, int ntri);

/* Batched device-pointer form of render_colors_core: `nframes` meshes sharing one triangle
 * list.  vertices[nframes][3*nver], colors[nframes][c*nver], image[nframes][h*w*c],
 * face_mask[nframes][h*w], depth_buffer[nframes][h*w], triangle_id (NULL ok)[nframes][h*w]. */
int vp_render_colors_batch_dev(unsigned char* image, unsigned char* face_mask, const float* vertices,
                               const int* triangles, const float* colors, float* depth_buffer,
                               int* triangle_id, int nframes, int nver, int ntri, int h, int w, int c,
                               int device, void* stream);

/* ---------------------------------------------------------------------------------------
 * Model object (utils/bfm_load_data.py:9-21) and reconstruction (utils/reconstruct_mesh.py)
 * --------------------------------------------------------------------------------------- */

/* meanshape[3N], idBase[3N][80], exBase[3N][64], meantex[3N], texBase[3N][80]: float32, or
 * float64 where the matching VP_F64_* bit is set.  tri[ntri][3], point_buf[nver][8] with pad
 * entries == ntri (i.e. the reference's F+1 after its "- 1").
 * center: the 3 per-axis means of meanshape that reconstruct_mesh.py:27 subtracts; pass the
 * host framework's value to reproduce its rounding, or NULL to have it computed here. */
int vp_model_create(vp_model** out, int device, int nver, int ntri, const void* meanshape,
                    const void* idBase, const void* exBase, const void* meantex,
                    const void* texBase, int float64_mask, const int* tri, const int* point_buf,
                    const double* center);
void vp_model_destroy(vp_model* m);
int vp_model_nver(const vp_model* m);
int vp_model_ntri(const vp_model* m);
int vp_model_ntiles(const vp_model* m);

/* Host-only introspection of the one-off mesh analysis vp_model_create performs (no CUDA
 * device needed): vertices renumbered along a Morton curve, triangles renumbered for the
 * rasterizer, vertex tiles with tile-local adjacency.  Used by the CPU test-suite to check the
 * tables against Compute_norm (reconstruct_mesh.py:35-52).
 *   tri[ntri][3], point_buf[nver][8] 0-based (pad = anything outside [0, ntri)), xyz[nver][3].
 * vp_topology_copy: v_int2orig[nver], tri_int[ntri][4] (internal a,b,c + original index),
 *   tiles[ntiles][7] (v_begin, nv, nlv, nlt, halo_off, ltri_off, fan), ltri[nltri] (3 x 10-bit local
 *   vertex ids), halo[nhalo] (internal vertex ids), ring[nver][8] (local triangle id, 0xFFFF pad),
 *   fan[nver][5] (tiles with fan = 1: byte offsets (local vertex * 16) of the ring vertices u0..u8, two per
 *   word, and in the top half of word 4 the mask of the pairs (u_i, u_i+1) that are faces of the vertex). */
typedef struct vp_topology vp_topology;
int vp_topology_build(vp_topology** out, int nver, int ntri, const int* tri, const int* point_buf,
                      const double* xyz);
void vp_topology_destroy(vp_topology* t);
int vp_topology_sizes(const vp_topology* t, int* ntiles, int* nltri, int* nhalo);
int vp_topology_copy(const vp_topology* t, int* v_int2orig, int* tri_int, int* tiles, uint32_t* ltri,
                     int* halo, uint16_t* ring, uint32_t* fan);
/* Optional bank-conflict-aware shared-memory slots of the fan tiles (used by the vertex kernel when the model was
 * always built): slot_off[ntiles] (offset into slot_tab, -1 for generic tiles),
 * slot_tab[vp_topology_slot_count()] (slot of local vertex i), fan_slot[nver][5] (fan records in slot space). */
/* Triangle ownership of the fused vertex + raster kernel (csrc/fused.cu): tile i owns rows
 * own_tri_off[i] .. own_tri_off[i + 1] of tri_int (the triangles whose smallest internal vertex it holds);
 * own_ltri[j] = the corners of row j as 3 x 10-bit vertex indices local to the owner tile; tri_by_orig[t][4] =
 * internal vertex ids of ORIGINAL triangle t (+ pad).  Returns 1 when every tile has fan records and every
 * owned triangle's corners are local to its owner (the fused kernel applies), 0 otherwise, < 0 on error. */
int vp_topology_copy_owned(const vp_topology* t, int* own_tri_off, uint32_t* own_ltri, int* tri_by_orig);
int vp_topology_slot_count(const vp_topology* t);
int vp_topology_copy_slots(const vp_topology* t, int* slot_off, uint16_t* slot_tab, uint32_t* fan_slot);

/* Post-raster composite of the frame loop, on the device (voicepuppet/pixrefer/infer_bfmvid.py:111-121 and
 * :234-236): the rasterized frames are resized to size x size exactly like cv2.resize does for 8-bit images
 * (fixed-point bilinear; 2x2 area mean for an exact 2x downscale), pasted at (x0, y0) into a zero canvas and
 * written as  canvas[T][H][W][3] uint8 (render_face's return value; channels swapped when swap_rb, :111)  and/or
 * inputs[T][H][W][in_channels] float32, channels channel_offset..+2 = raster / 255 in the raster's own channel
 * order (what lands in PixReferNet's inputs[..., 3:6]).  All pointers are device pointers; asynchronous on
 * `stream`.  A face that does not fit the canvas is VP_ERR_ARG (numpy raises at :121).
 * vp_composite_placement evaluates :80-82 and :112-121: size = round(res / (ratio * tp[2])), x0 / y0. */
/* Host only (no device needed): the per-axis coefficient table [dsize][4] = (s0, s1, c0, c1) vp_composite_dev uses. */
int vp_composite_axis_table(int ssize, int dsize, int is_y, int* out4);
int vp_composite_placement(int res, int center_x, int center_y, double ratio, const double* transform_params5,
                           int* size, int* x0, int* y0);
int vp_composite_dev(const unsigned char* frames_dev, int nframes, int res, int size, int x0, int y0,
                     int canvas_h, int canvas_w, unsigned char* canvas_dev, int swap_rb, float* inputs_dev,
                     int in_channels, int channel_offset, int device, void* stream);

/* Training-time twin (voicepuppet/bfmnet/bfmnet.py:215-268, BFMNet's vertex loss): with Delta = ex_label - ex_pred
 * and D = exBase . Delta,
 *   loss = 1/B sum_{b,t<len_b} sum_r M_r |D[b,t,r]|  +  1/B sum_{b,t<len_b-1} sum_r M_r |D[b,t,r] - D[b,t+1,r]|
 * (identity and mean cancel in the reference's face_shape - output_face_shape), and its gradient with respect to
 * Delta (= minus the gradient with respect to the predicted coefficients).
 * vp_loss_mask_create: mask[nver*3] f32 in the model's vertex order (the reference's mouth_mask, :134-137) -> device
 * array in the library's row order.  vp_expression_loss_dev: delta_ex_dev[batch*frames][64], seq_len_dev[batch] i32,
 * loss_dev one double, grad_delta_dev[batch*frames][64] (NULL = forward only); device pointers, asynchronous. */
int vp_loss_mask_create(vp_model* m, const float* mask_host, float** mask_dev);
void vp_loss_mask_destroy(float* mask_dev);
int vp_expression_loss_dev(vp_model* m, const float* delta_ex_dev, const int* seq_len_dev, const float* mask_dev,
                           int batch, int frames, double* loss_dev, float* grad_delta_dev, void* stream);

/* Expression-basis kernel selection: 0 = automatic (FP32 streamed kernel below 16 frames per launch,
 * tcgen05 3xTF32 GEMM from 16 frames up), 1 = always FP32 SIMT, 2 = always tcgen05 3xTF32. */
int vp_set_basis_mode(vp_model* m, int mode);

/* Vertex-normal path selection: 0 = automatic (fan records where the mesh chains into fans, positions staged at
 * bank-conflict-aware shared-memory slots; the generic ring-of-faces kernel elsewhere), 1 = always the generic
 * kernel, 2 = fan records with the identity slot placement (tests compare the three). */
int vp_set_vertex_mode(vp_model* m, int mode);
int vp_model_fan_tiles(const vp_model* m); /* tiles that take the fan path (of vp_model_ntiles) */

/* Chunk pipeline of vp_render_sequence*: 0 = automatic (the measured best: today the separate kernels, vertex
 * records -> scatter -> resolve, at every size), 1 = always the separate kernels, 2 = the fused vertex + z-buffer
 * kernel (csrc/fused.cu: one CTA per vertex tile projects its own and halo vertices and rasterizes the triangles it
 * owns out of shared memory; no vertex records, colours resolved from per-vertex colours).  Mode 2 needs every tile
 * to have fan records and every triangle's corners to be local to its owner tile (any manifold mesh with a
 * consistent point_buf: vp_model_fused_available() == 1) and is bit-identical to mode 1 (tests). */
int vp_set_raster_path(vp_model* m, int mode);
int vp_model_fused_available(const vp_model* m);

/* Per-clip constants ("identity mean precomputed once"): base shape = meanshape + idBase.id
 * - center, texture = meantex + texBase.tex.  Either pointer may be NULL to keep the old one. */
int vp_set_identity(vp_model* m, const float* id_coeff80, const float* tex_coeff80);
/* Replace the base shape by an explicit [nver][3] float64 array (stage-function callers). */
int vp_set_base_shape(vp_model* m, const double* shape);
/* Replace the texture by an explicit [nver][3] float32 array. */
int vp_set_texture(vp_model* m, const float* texture);
/* Copy out the current texture [nver][3] float32 / base shape [nver][3] float64. */
int vp_get_texture(vp_model* m, float* texture);
int vp_get_base_shape(vp_model* m, double* shape);

/* Per-frame inputs of the batched entry points (all host, row-major):
 *   ex[nframes][64]        expression coefficients; NULL = no expression displacement
 *   rotation[nframes][9]   float64 row-major 3x3, the matrix of Compute_rotation_matrix
 *   translation[nframes][3], gamma[nframes][27] float32
 * rotate_shape_first: 0 = Reconstruction (reconstruct_mesh.py:172-194),
 *                     1 = Reconstruction_rotation (:198-223: shape rotated, then projected
 *                         with the same rotation again; normals rotated once). */
typedef struct vp_frames {
  int nframes;
  const float* ex;
  const double* rotation;
  const float* translation;
  const float* gamma;
  int rotate_shape_first;
  double focal;  /* 1015.0 */
  double center; /* 112.0 */
} vp_frames;

/* Reconstruction outputs, any may be NULL; all [nframes][nver][k] in the model's vertex order.
 *   face_shape f64[3] (rotated once when rotate_shape_first), face_norm f32[3] (unit normal
 *   BEFORE rotation, as Compute_norm returns), face_color f32[3] (unclamped),
 *   projection f64[2] = (x, image_size - y) when flip_y else (x, y), z_buffer f64[1]. */
typedef struct vp_recon_out {
  double* face_shape;
  float* face_norm;
  float* face_color;
  double* projection;
  double* z_buffer;
  int flip_y;
  double image_size; /* 224.0 */
} vp_recon_out;

int vp_reconstruct(vp_model* m, const vp_frames* frames, const vp_recon_out* out);

/* Illumination_layer on caller-supplied texture/normals (reconstruct_mesh.py:129-168).
 * texture[n][3], norm[n][3] float64, gamma[27] float32 -> color[n][3], lighting[n][3] float64. */
int vp_illumination(int device, int n, const double* texture, const double* norm, const float* gamma,
                    double* color, double* lighting);

/* Projection_layer on a caller-supplied shape (reconstruct_mesh.py:100-120), float64:
 * shape[n][3], rotation9 row-major 3x3, translation3 -> projection[n][2] (not flipped), z_buffer[n]. */
int vp_projection(int device, int n, const double* shape, const double* rotation9,
                  const float* translation3, double focal, double center, double* projection,
                  double* z_buffer);

/* The whole hot path, coefficients -> rendered frames (infer_bfmvid.py:91-109 per frame):
 * reconstruction, colours clipped to [0,255] and truncated, vertices (x, S - y, -z) scaled by
 * res/224, flat-shaded z-buffer rasterization at h = w = res.
 *   image[nframes][res][res][3] u8, face_mask[nframes][res][res] u8 (NULL ok).
 * outputs_on_device: 0 = host pointers (copied back, pipelined per chunk; synchronous),
 *                    1 = device pointers (asynchronous: ordered on `stream`, nothing is synchronised). */
int vp_render_sequence(vp_model* m, const vp_frames* frames, int res, unsigned char* image,
                       unsigned char* face_mask, int outputs_on_device, void* stream);

/* Device-resident variant: per-frame inputs already on the device, packed as
 * ex_dev[nframes][64] f32 and params_dev[nframes] of vp_frame_params; nothing is copied or
 * synchronised.  This is what bench.py's device-resident `value` measures. */
typedef struct vp_frame_params {
  double rotation[9];
  float translation[3];
  float gamma[27];
} vp_frame_params;

int vp_render_sequence_dev(vp_model* m, int nframes, const float* ex_dev,
                           const vp_frame_params* params_dev, int rotate_shape_first, int res,
                           unsigned char* image_dev, unsigned char* face_mask_dev, void* stream);

/* vp_render_sequence_dev in chunks of `notify_frames` frames, recording events[i] (cudaEvent_t handles
 * owned by the caller) on `stream` as soon as chunk i is complete, so that finished frames can be moved
 * (NCCL gather to rank 0, device->host copy) while the rest is still rendering.  The expression
 * contraction still runs once for the whole call.  nevents >= ceil(nframes / notify_frames). */
int vp_render_sequence_dev_notify(vp_model* m, int nframes, const float* ex_dev,
                                  const vp_frame_params* params_dev, int rotate_shape_first, int res,
                                  unsigned char* image_dev, unsigned char* face_mask_dev, void* stream,
                                  int notify_frames, void** events, int nevents);

/* Same with an explicit plan: chunk i renders chunk_frames[i] frames (they add up to nframes) and
 * events[i] is recorded when it is complete. */
int vp_render_sequence_dev_chunks(vp_model* m, int nframes, const float* ex_dev,
                                  const vp_frame_params* params_dev, int rotate_shape_first, int res,
                                  unsigned char* image_dev, unsigned char* face_mask_dev, void* stream,
                                  const int* chunk_frames, int nchunks, void** events);

/* The expression contraction alone (device pointers, asynchronous on `stream`):
 * disp_dev[t][r] = sum_k exBase[r][k] * ex_dev[t][k], r in the library's internal row order
 * (3 * internal vertex + axis), vp_model_rows_pad() floats per frame. */
int vp_basis_dev(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, void* stream);
int vp_model_rows_pad(const vp_model* m);
/* Diagnostics: one tcgen05 basis launch (nframes <= 128) that also writes a clock64() timeline of
 * CTA 0 to trace_dev[0..256) (4 roles x 16 tiles x 4 marks) and the %globaltimer entry / exit time (ns) of every CTA c
 * to trace_dev[256 + 2 c], trace_dev[257 + 2 c]; trace_dev holds 1024 int64 entries. */
int vp_debug_basis_trace(vp_model* m, const float* ex_dev, float* disp_dev, int nframes,
                         long long* trace_dev, void* stream);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
unsigned long long vp_launch_count(void);

/* Per-kernel device time of the last vp_render_sequence(_dev) call when profiling is enabled
 * (CUDA events around each kernel, accumulated over chunks; enabling it makes the _dev variant
 * synchronise).  Slots: names "basis;vertex;scatter;resolve;fused" (semicolon separated; "fused" = the fused vertex + scatter kernel, which replaces "vertex" and "scatter" when it applies). */
int vp_set_profiling(vp_model* m, int enabled);
int vp_get_profile(vp_model* m, char* names, int names_cap, float* ms, int ms_cap);
/* launches[k] = number of kernel launches summed into ms[k] of vp_get_profile by the last profiled call. */
int vp_get_profile_launches(vp_model* m, int* launches, int cap);

#ifdef __cplusplus
}
#endif
#endif /* VPB200_H_ */
