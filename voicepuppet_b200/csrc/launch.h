// Internal launch API between the translation units of libvpb200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <vector>

#include "common.h"
#include "vp_math.cuh"

namespace vp {

// ---- rasterizer (raster.cu) -----------------------------------------------------------
int launch_keys_from_depth(const float* depth_dev, unsigned long long* keys_dev, size_t n, cudaStream_t st);
int launch_scatter_generic(int mode, const float* vertices, size_t frame_stride, const int* triangles,
                           unsigned long long* keys, int nframes, int ntri, int h, int w, cudaStream_t st);
// Fused path: keys carry the chunk's epoch in their top bits (see EpochKey in raster.cuh), so the
// z-buffer is cleared only when the epoch counter wraps (epoch_limit) or the buffer is reallocated.
uint32_t epoch_limit(int ntri);
int inline_box_pixels();
int walk_group_lanes();
int walk_group_min_pixels();
int launch_scatter_packed(const float4* vrec, size_t frame_stride, const int4* triangles,
                          unsigned long long* keys, uint32_t* tri_color, uint32_t epoch, int nframes, int ntri, int h,
                          int w, cudaStream_t st);
int launch_resolve_packed(const unsigned long long* keys, const uint32_t* tri_color, uint32_t epoch,
                          unsigned char* image, unsigned char* mask, int nframes, int ntri, int h, int w, cudaStream_t st);

// ---- vertex-tile topology (topology.cpp builds it, reconstruct.cu consumes it) ----------
constexpr int kTileV = 128;    // max own vertices per tile == threads per CTA of the vertex kernel
constexpr int kTileLV = 256;   // max own + halo vertices (two staging slots per thread)
constexpr int kTileLT = 512;   // max triangles touching the tile's own vertices
constexpr uint16_t kRingPad = 0xFFFF;

struct TileDesc {
  int v_begin;   // first own vertex (internal order); own vertices are contiguous
  int nv;        // own vertices
  int nlv;       // local vertices (own first, then halo)
  int nlt;       // local triangles
  int halo_off;  // into halo[] (internal vertex ids of local vertices nv..nlv-1)
  int ltri_off;  // into ltri[] (3 x 10-bit local vertex indices)
  int fan;       // 1: every own vertex has a fan record (see Topology::fan), 0: generic ring-of-faces path
};

// Fan record of one vertex v: up to 9 tile-local vertices u0..u8 such that every face point_buf lists for
// v is (v, u_i, u_i+1) in the triangle's cyclic order for the i whose bit is set in `mask`; then
// sum of face normals = sum over set bits of (u_i - v) x (u_i+1 - v), which needs 9 position gathers
// instead of a staged pass over the tile's triangles plus 8 normal gathers.  Five words: byte offsets
// (local index * 16) of u_2k | u_2k+1 << 16 for k = 0..3, then u_8 | mask << 16.  Unused entries name v.
constexpr int kFanWords = 5;
constexpr int kFanEntries = 9;

// Host-side result of the one-off mesh analysis (topology.cpp).
struct Topology {
  int nver = 0, ntri = 0;
  std::vector<int> v_int2orig, v_orig2int;
  std::vector<int> tri_int;        // [ntri][4]: internal vertex ids a,b,c + ORIGINAL triangle index
  std::vector<TileDesc> tiles;
  std::vector<uint32_t> ltri;
  std::vector<int> halo;
  std::vector<uint16_t> ring;      // [nver][8] in internal vertex order
  std::vector<uint32_t> fan;       // [nver][kFanWords] in internal vertex order (zeros where the tile is generic)
  // Optional (build_topology(..., with_slots)): shared-memory slots decoupled from the local numbering.  The
  // distinct vertices a quarter-warp reads at one fan step should sit in different 16-byte bank groups, so the
  // local vertices of a fan tile are 8-coloured (bank group = slot % 8) and slot = colour + 8 * rank in colour.
  std::vector<int> slot_off;       // [ntiles] offset of the tile's slots in slot_tab, -1 for generic tiles
  std::vector<uint16_t> slot_tab;  // per fan tile nlv entries: shared-memory slot of local vertex i (< nlv + 8)
  std::vector<uint32_t> fan_slot;  // [nver][kFanWords]: the fan record with slot byte offsets
  // Triangle ownership for the fused vertex + raster kernel (fused.cu): a triangle belongs to the tile of its
  // smallest internal vertex; tri_int is sorted by that vertex, so a tile owns a contiguous run of it.
  std::vector<int> own_tri_off;    // [ntiles + 1] run of tri_int rows owned by tile i
  std::vector<uint32_t> own_ltri;  // [ntri] 3 x 10-bit local vertex indices of tri_int row j in its OWNER's numbering
  std::vector<int> tri_by_orig;    // [ntri][4] internal vertex ids of ORIGINAL triangle t (+ pad): the resolve pass's gather
  bool fused_ok = false;           // every tile has fan records and every owned triangle's corners are local to its owner
};

// tri: [ntri][3] 0-based original vertex ids; point_buf: [nver][8] 0-based original triangle ids
// (anything outside [0, ntri) is a pad slot); xyz: [nver][3] mean shape used for the spatial order.
int build_topology(Topology& out, int nver, int ntri, const int* tri, const int* point_buf, const double* xyz,
                   bool with_slots = false);

// Optional per-vertex outputs of the vertex kernel, in the MODEL's (original) vertex order.
struct ReconOut {
  double* shape = nullptr;   // [T][nver][3]
  float* norm = nullptr;     // [T][nver][3]
  float* color = nullptr;    // [T][nver][3]
  double* proj = nullptr;    // [T][nver][2]
  double* zbuf = nullptr;    // [T][nver]
  int flip_y = 1;
};

}  // namespace vp

// The model object behind the opaque C handle.
struct vp_model {
  int device = 0;
  int nver = 0, ntri = 0;
  int rows = 0;         // 3 * nver
  int rows_pad = 0;     // rows rounded up to 128: row count of exb and floats per frame in disp
  int vrec_stride = 0;  // float4 per frame in the vertex-record buffer
  double center[3] = {0, 0, 0};

  vp::Topology topo;    // host copy (permutations are needed by the setters/getters)

  // device: model (internal vertex order; row = 3 * internal vertex + axis)
  float* exb = nullptr;         // [rows_pad][64] float32
  void* idb = nullptr;          // [rows][80] float32/float64
  void* texb = nullptr;         // [rows][80]
  double* meanshape = nullptr;  // [rows]
  double* meantex = nullptr;    // [rows]
  bool idb64 = false, texb64 = false;
  int4* tri = nullptr;          // [ntri] internal vertex ids + original triangle index
  int* v_int2orig_dev = nullptr;
  int* t_orig2int_dev = nullptr; // [ntri] original triangle index -> internal (rasterizer) order
  // device: vertex tiles
  int ntiles = 0;
  vp::TileDesc* tiles = nullptr;
  uint32_t* ltri = nullptr;
  int* halo = nullptr;
  uint16_t* ring = nullptr;     // [nver][8] local triangle index per point_buf slot
  uint32_t* fan = nullptr;      // [nver][5] fan records (tiles with TileDesc::fan)
  // bank-conflict-aware slot tables of the fan kernel (see Topology)
  int* slot_off = nullptr;
  uint16_t* slot_tab = nullptr;
  uint32_t* fan_slot = nullptr;
  bool have_slots = false;
  // fused vertex + raster kernel (fused.cu): triangle ownership tables, see Topology
  int* own_tri_off = nullptr;
  uint32_t* own_ltri = nullptr;
  int4* tri_by_orig = nullptr;  // [ntri] internal vertex ids of ORIGINAL triangle t
  bool fused_ok = false;
  int fused_mode = 0;           // vp_set_raster_path: 0 = automatic (separate kernels), 1 = separate, 2 = fused kernel
  int* tile_list = nullptr;     // tile ids: the n_fan_tiles fan tiles first, then the generic ones
  int n_fan_tiles = 0;
  // TMA descriptor of exb for the tcgen05 basis kernel (a CUtensorMap, kept opaque here)
  alignas(64) unsigned char tmap_exb[128] = {0};
  bool have_tmap = false;
  // device: per-clip state
  double* base = nullptr;       // [nver][3]
  float* tex = nullptr;         // [nver][3]
  float* coeff_tmp = nullptr;   // [160] staging for id/tex coefficients
  bool have_base = false, have_tex = false;

  // workspaces (grow only)
  vp::DevBuf ws_fshared, ws_ex, ws_params, ws_disp, ws_vrec, ws_vcol, ws_keys, ws_tricol, ws_img[2], ws_mask[2], ws_out;
  uint32_t key_epoch = 0;       // epoch of the last chunk rendered into ws_keys (0 = buffer must be cleared)
  // page-locked staging of the per-frame inputs (so their upload is a real async copy)
  void* h_stage = nullptr;
  size_t h_stage_cap = 0;
  cudaEvent_t ev_stage = nullptr;  // recorded after the uploads that read h_stage
  // second compute stream: consecutive chunks alternate between the caller's stream and this one (and between
  // the two halves of the chunk workspaces), so the ramp-up of one chunk's kernels fills the tails of the other's
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_basis = nullptr, ev_aux_done = nullptr, ev_main_done = nullptr;
  // recorded on the caller's stream at the end of every asynchronous render; the entry points that rewrite model
  // state (identity, base shape, texture, workspaces) on another stream wait for it first (wait_for_renders)
  cudaEvent_t ev_render_done = nullptr;
  bool render_pending = false;
  cudaStream_t last_render_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_render[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};

  int basis_mode = 0;           // vp::BasisMode
  int vertex_mode = 0;          // 0 = fan records where available (bank-aware slots), 1 = generic kernel everywhere,
                                // 2 = fan records with local vertex i staged at slot i (tests compare the placements)
  // profiling
  bool profiling = false;
  float prof_ms[8] = {0};
  int prof_launches[8] = {0};   // launches summed into prof_ms, per slot

  std::mutex mu;
};

namespace vp {

// Call (under m->mu) before touching state an asynchronous render still in flight may read: the renders are
// ordered on their caller's stream, which need not be the stream (often the legacy one) the setters use.
inline int wait_for_renders(vp_model* m) {
  if (m->render_pending && m->ev_render_done) {
    VP_CUDA(cudaEventSynchronize(m->ev_render_done));
    m->render_pending = false;
  }
  return VP_OK;
}

// A render on `st` reuses the workspaces of the previous one: order it behind that one when it ran on another stream.
inline int order_after_renders(vp_model* m, cudaStream_t st) {
  if (m->render_pending && m->ev_render_done && m->last_render_stream != st)
    VP_CUDA(cudaStreamWaitEvent(st, m->ev_render_done, 0));
  return VP_OK;
}
inline void note_render(vp_model* m, cudaStream_t st) {
  if (m->ev_render_done && cudaEventRecord(m->ev_render_done, st) == cudaSuccess) {
    m->render_pending = true;
    m->last_render_stream = st;
  }
}

enum ProfSlot { kProfBasis = 0, kProfVertex = 1, kProfScatter = 2, kProfResolve = 3, kProfFused = 4, kProfSlots = 5 };

// ---- reconstruction (reconstruct.cu) ----------------------------------------------------
int launch_identity(vp_model* m, const float* id_dev, const float* tex_dev, cudaStream_t st);
int launch_basis(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st);
int launch_basis_simt(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st);
// tcgen05 3xTF32 flavour (basis_tc.cu); basis_tc_prepare builds the TMA descriptor once per model
int basis_tc_prepare(vp_model* m);
int launch_basis_tc(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, cudaStream_t st,
                    long long* trace_dev = nullptr);
enum BasisMode { kBasisAuto = 0, kBasisSimt = 1, kBasisTensor = 2 };
constexpr int kBasisTensorMinFrames = 16;  // below this the frame batch is a GEMV, not a dense contraction
// disp_dev may be NULL (no expression displacement).  vrec_dev may be NULL (no raster records).
// frame_constants: NULL (prepared here from params_dev) or this chunk's slice of prepare_frame_constants().
int launch_vertex(vp_model* m, const float* disp_dev, const FrameParams* params_dev, int nframes,
                  int rotate_first, double focal, double center, double image_size, double raster_scale,
                  float4* vrec_dev, const ReconOut& out, cudaStream_t st, const void* frame_constants = nullptr);
int prepare_frame_constants(vp_model* m, const FrameParams* params_dev, int nframes, int rotate_first, double focal,
                            double center, double image_size, double raster_scale, cudaStream_t st, const void** out);
size_t frame_constants_stride();

// ---- fused vertex + raster path (fused.cu) ----------------------------------------------
bool fused_available(const vp_model* m);
int launch_fused(vp_model* m, const float* disp_dev, int nframes, const void* frame_constants, uint32_t* vcol,
                 unsigned long long* keys, uint32_t epoch, int res, cudaStream_t st);
int launch_resolve_vcol(const vp_model* m, const unsigned long long* keys, const uint32_t* vcol, uint32_t epoch,
                        unsigned char* image, unsigned char* mask, int nframes, int h, int w, cudaStream_t st);

}  // namespace vp
