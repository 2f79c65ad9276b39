// The device-resident BFM model object (reference utils/bfm_load_data.py:9-21, class BFM) and
// its per-clip state (identity shape and texture, reconstruct_mesh.py:20-29,58-62).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "launch.h"

using namespace vp;

namespace {

template <typename T>
int upload(T** dev, const std::vector<T>& host) {
  *dev = nullptr;
  const size_t bytes = std::max<size_t>(host.size(), 1) * sizeof(T);
  VP_CUDA(cudaMalloc(reinterpret_cast<void**>(dev), bytes));
  if (!host.empty()) VP_CUDA(cudaMemcpy(*dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  return VP_OK;
}

inline double load_as_double(const void* p, bool is64, size_t i) {
  return is64 ? static_cast<const double*>(p)[i] : (double)static_cast<const float*>(p)[i];
}

// Row-permuted copy of a [3*nver][k] basis: internal row 3*i+a <- original row 3*int2orig[i]+a.
template <typename Src, typename Dst>
void permute_rows(std::vector<Dst>& dst, const Src* src, const std::vector<int>& int2orig, int k, size_t rows_out) {
  dst.assign(rows_out * (size_t)k, Dst(0));
  for (size_t i = 0; i < int2orig.size(); ++i)
    for (int a = 0; a < 3; ++a) {
      const Src* s = src + (3 * (size_t)int2orig[i] + a) * k;
      Dst* d = dst.data() + (3 * i + a) * k;
      for (int j = 0; j < k; ++j) d[j] = static_cast<Dst>(s[j]);
    }
}

template <typename Src>
int upload_basis(void** dev, const void* src, const std::vector<int>& int2orig, int k, size_t rows_out) {
  std::vector<Src> tmp;
  permute_rows<Src, Src>(tmp, static_cast<const Src*>(src), int2orig, k, rows_out);
  Src* d = nullptr;
  VP_TRY(upload(&d, tmp));
  *dev = d;
  return VP_OK;
}

void free_model(vp_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->exb);
  cudaFree(m->idb);
  cudaFree(m->texb);
  cudaFree(m->meanshape);
  cudaFree(m->meantex);
  cudaFree(m->tri);
  cudaFree(m->v_int2orig_dev);
  cudaFree(m->t_orig2int_dev);
  cudaFree(m->tiles);
  cudaFree(m->ltri);
  cudaFree(m->halo);
  cudaFree(m->ring);
  cudaFree(m->fan);
  cudaFree(m->own_tri_off);
  cudaFree(m->own_ltri);
  cudaFree(m->tri_by_orig);
  cudaFree(m->slot_off);
  cudaFree(m->slot_tab);
  cudaFree(m->fan_slot);
  cudaFree(m->tile_list);
  cudaFree(m->base);
  cudaFree(m->tex);
  cudaFree(m->coeff_tmp);
  m->ws_fshared.release();
  m->ws_ex.release();
  m->ws_params.release();
  m->ws_disp.release();
  m->ws_vrec.release();
  m->ws_keys.release();
  m->ws_tricol.release();
  m->ws_out.release();
  for (int i = 0; i < 2; ++i) {
    m->ws_img[i].release();
    m->ws_mask[i].release();
    if (m->ev_render[i]) cudaEventDestroy(m->ev_render[i]);
    if (m->ev_copy[i]) cudaEventDestroy(m->ev_copy[i]);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->ev_stage) cudaEventDestroy(m->ev_stage);
  for (cudaEvent_t e : {m->ev_fork, m->ev_basis, m->ev_aux_done, m->ev_main_done, m->ev_render_done})
    if (e) cudaEventDestroy(e);
  if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  (void)cudaGetLastError();
  delete m;
}

int create_model(vp_model* m, int nver, int ntri, const void* meanshape, const void* idBase, const void* exBase,
                 const void* meantex, const void* texBase, int f64mask, const int* tri, const int* point_buf,
                 const double* center) {
  m->nver = nver;
  m->ntri = ntri;
  m->rows = 3 * nver;
  m->rows_pad = (m->rows + 127) / 128 * 128;
  m->vrec_stride = (nver + 31) / 32 * 32;

  const bool ms64 = f64mask & VP_F64_MEANSHAPE, mt64 = f64mask & VP_F64_MEANTEX, ex64 = f64mask & VP_F64_EXBASE;
  m->idb64 = f64mask & VP_F64_IDBASE;
  m->texb64 = f64mask & VP_F64_TEXBASE;

  std::vector<double> xyz((size_t)m->rows);
  for (size_t i = 0; i < xyz.size(); ++i) xyz[i] = load_as_double(meanshape, ms64, i);
  // bank-conflict-aware shared-memory slots of the fan vertex kernel: measured on B200 (profiles/r02a_optin_flavours.jsonl:
  // 56.3 -> 51.8 us per 75 frames at 256x256), always built
  const bool with_slots = true;
  VP_TRY(build_topology(m->topo, nver, ntri, tri, point_buf, xyz.data(), with_slots));
  const std::vector<int>& i2o = m->topo.v_int2orig;

  if (center) {
    for (int a = 0; a < 3; ++a) m->center[a] = center[a];
  } else {
    for (int a = 0; a < 3; ++a) {
      double s = 0;
      for (int v = 0; v < nver; ++v) s += xyz[3 * (size_t)v + a];
      m->center[a] = nver ? s / nver : 0.0;
    }
  }

  // means, permuted, as float64
  std::vector<double> tmp((size_t)m->rows);
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) tmp[3 * i + a] = xyz[3 * (size_t)i2o[i] + a];
  VP_TRY(upload(&m->meanshape, tmp));
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) tmp[3 * i + a] = load_as_double(meantex, mt64, 3 * (size_t)i2o[i] + a);
  VP_TRY(upload(&m->meantex, tmp));

  // expression basis: float32 whatever the source dtype (the contraction runs in FP32 / 3xTF32)
  {
    std::vector<float> e;
    if (ex64)
      permute_rows<double, float>(e, static_cast<const double*>(exBase), i2o, VP_N_EX, (size_t)m->rows_pad);
    else
      permute_rows<float, float>(e, static_cast<const float*>(exBase), i2o, VP_N_EX, (size_t)m->rows_pad);
    VP_TRY(upload(&m->exb, e));
  }
  // identity and texture bases keep their dtype; they are contracted once per clip in float64
  if (m->idb64)
    VP_TRY(upload_basis<double>(&m->idb, idBase, i2o, VP_N_ID, (size_t)m->rows));
  else
    VP_TRY(upload_basis<float>(&m->idb, idBase, i2o, VP_N_ID, (size_t)m->rows));
  if (m->texb64)
    VP_TRY(upload_basis<double>(&m->texb, texBase, i2o, VP_N_TEX, (size_t)m->rows));
  else
    VP_TRY(upload_basis<float>(&m->texb, texBase, i2o, VP_N_TEX, (size_t)m->rows));

  // topology
  {
    int* t4 = nullptr;
    VP_TRY(upload(&t4, m->topo.tri_int));
    m->tri = reinterpret_cast<int4*>(t4);
  }
  VP_TRY(upload(&m->v_int2orig_dev, m->topo.v_int2orig));
  {
    std::vector<int> o2i((size_t)ntri);
    for (int i = 0; i < ntri; ++i) o2i[m->topo.tri_int[4 * (size_t)i + 3]] = i;
    VP_TRY(upload(&m->t_orig2int_dev, o2i));
  }
  m->ntiles = (int)m->topo.tiles.size();
  VP_TRY(upload(&m->tiles, m->topo.tiles));
  VP_TRY(upload(&m->ltri, m->topo.ltri));
  VP_TRY(upload(&m->halo, m->topo.halo));
  VP_TRY(upload(&m->ring, m->topo.ring));
  VP_TRY(upload(&m->fan, m->topo.fan));
  VP_TRY(upload(&m->own_tri_off, m->topo.own_tri_off));
  VP_TRY(upload(&m->own_ltri, m->topo.own_ltri));
  {
    int* t4 = nullptr;
    VP_TRY(upload(&t4, m->topo.tri_by_orig));
    m->tri_by_orig = reinterpret_cast<int4*>(t4);
  }
  m->fused_ok = m->topo.fused_ok && nver > 0 && ntri > 0;
  if (with_slots) {
    VP_TRY(upload(&m->slot_off, m->topo.slot_off));
    VP_TRY(upload(&m->slot_tab, m->topo.slot_tab));
    VP_TRY(upload(&m->fan_slot, m->topo.fan_slot));
    m->have_slots = true;
  }
  {
    std::vector<int> order;
    for (int pass = 1; pass >= 0; --pass)
      for (int i = 0; i < m->ntiles; ++i)
        if (m->topo.tiles[i].fan == pass) order.push_back(i);
    m->n_fan_tiles = 0;
    for (const vp::TileDesc& td : m->topo.tiles) m->n_fan_tiles += td.fan;
    VP_TRY(upload(&m->tile_list, order));
  }

  // per-clip state
  VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->base), std::max<size_t>(m->rows, 1) * sizeof(double)));
  VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tex), std::max<size_t>(m->rows, 1) * sizeof(float)));
  VP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->coeff_tmp), 160 * sizeof(float)));
  VP_CUDA(cudaMemset(m->coeff_tmp, 0, 160 * sizeof(float)));

  if (basis_tc_prepare(m) != VP_OK) m->have_tmap = false;  // the SIMT kernel still works; mode 2 reports it

  VP_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  VP_CUDA(cudaEventCreateWithFlags(&m->ev_stage, cudaEventDisableTiming));
  VP_CUDA(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
  for (cudaEvent_t* e : {&m->ev_fork, &m->ev_basis, &m->ev_aux_done, &m->ev_main_done, &m->ev_render_done})
    VP_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    VP_CUDA(cudaEventCreateWithFlags(&m->ev_render[i], cudaEventDisableTiming));
    VP_CUDA(cudaEventCreateWithFlags(&m->ev_copy[i], cudaEventDisableTiming));
  }
  return VP_OK;
}

}  // namespace

extern "C" int vp_model_create(vp_model** out, int device, int nver, int ntri, const void* meanshape,
                               const void* idBase, const void* exBase, const void* meantex, const void* texBase,
                               int float64_mask, const int* tri, const int* point_buf, const double* center) {
  VP_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  VP_REQUIRE(nver > 0 && ntri >= 0, "nver > 0 and ntri >= 0");
  VP_REQUIRE(meanshape && idBase && exBase && meantex && texBase && point_buf, "null model array");
  VP_REQUIRE(ntri == 0 || tri, "null triangle array");
  VP_CUDA(cudaSetDevice(device));
  vp_model* m = new vp_model();
  m->device = device;
  const int rc = create_model(m, nver, ntri, meanshape, idBase, exBase, meantex, texBase, float64_mask, tri,
                              point_buf, center);
  if (rc != VP_OK) {
    free_model(m);
    return rc;
  }
  *out = m;
  return VP_OK;
}

extern "C" void vp_model_destroy(vp_model* m) { free_model(m); }

extern "C" int vp_model_nver(const vp_model* m) { return m ? m->nver : -1; }
extern "C" int vp_model_ntri(const vp_model* m) { return m ? m->ntri : -1; }
extern "C" int vp_model_ntiles(const vp_model* m) { return m ? m->ntiles : -1; }

extern "C" int vp_set_basis_mode(vp_model* m, int mode) {
  VP_REQUIRE(m != nullptr, "null model");
  VP_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (auto), 1 (FP32 SIMT) or 2 (tcgen05 3xTF32)");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(mode != kBasisTensor || m->have_tmap, "tensor-core basis kernel unavailable on this device/driver");
  m->basis_mode = mode;
  return VP_OK;
}

extern "C" int vp_set_raster_path(vp_model* m, int mode) {
  VP_REQUIRE(m != nullptr, "null model");
  VP_REQUIRE(mode == 0 || mode == 1 || mode == 2, "raster path must be 0 (automatic), 1 (separate kernels) or 2 (fused kernel)");
  VP_REQUIRE(mode != 2 || m->fused_ok, "the fused kernel cannot take this mesh (vp_model_fused_available() == 0)");
  std::lock_guard<std::mutex> lock(m->mu);
  m->fused_mode = mode;
  return VP_OK;
}

extern "C" int vp_model_fused_available(const vp_model* m) { return (m && m->fused_ok) ? 1 : 0; }

extern "C" int vp_set_vertex_mode(vp_model* m, int mode) {
  VP_REQUIRE(m != nullptr, "null model");
  VP_REQUIRE(mode == 0 || mode == 1 || mode == 2, "mode must be 0 (auto), 1 (generic kernel) or 2 (fan records, identity slots)");
  std::lock_guard<std::mutex> lock(m->mu);
  m->vertex_mode = mode;
  return VP_OK;
}

extern "C" int vp_model_fan_tiles(const vp_model* m) { return m ? m->n_fan_tiles : -1; }

extern "C" int vp_set_identity(vp_model* m, const float* id_coeff80, const float* tex_coeff80) {
  VP_REQUIRE(m != nullptr, "null model");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));
  cudaStream_t st = nullptr;
  if (id_coeff80)
    VP_CUDA(cudaMemcpyAsync(m->coeff_tmp, id_coeff80, VP_N_ID * sizeof(float), cudaMemcpyHostToDevice, st));
  if (tex_coeff80)
    VP_CUDA(cudaMemcpyAsync(m->coeff_tmp + VP_N_ID, tex_coeff80, VP_N_TEX * sizeof(float), cudaMemcpyHostToDevice, st));
  VP_TRY(launch_identity(m, id_coeff80 ? m->coeff_tmp : nullptr, tex_coeff80 ? m->coeff_tmp + VP_N_ID : nullptr, st));
  VP_CUDA(cudaStreamSynchronize(st));
  if (id_coeff80) m->have_base = true;
  if (tex_coeff80) m->have_tex = true;
  return VP_OK;
}

extern "C" int vp_set_base_shape(vp_model* m, const double* shape) {
  VP_REQUIRE(m != nullptr && shape != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));
  std::vector<double> tmp((size_t)m->rows);
  const std::vector<int>& i2o = m->topo.v_int2orig;
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) tmp[3 * i + a] = shape[3 * (size_t)i2o[i] + a];
  VP_CUDA(cudaMemcpy(m->base, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
  m->have_base = true;
  return VP_OK;
}

extern "C" int vp_set_texture(vp_model* m, const float* texture) {
  VP_REQUIRE(m != nullptr && texture != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));
  std::vector<float> tmp((size_t)m->rows);
  const std::vector<int>& i2o = m->topo.v_int2orig;
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) tmp[3 * i + a] = texture[3 * (size_t)i2o[i] + a];
  VP_CUDA(cudaMemcpy(m->tex, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice));
  m->have_tex = true;
  return VP_OK;
}

extern "C" int vp_get_texture(vp_model* m, float* texture) {
  VP_REQUIRE(m != nullptr && texture != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(m->have_tex, "no texture set (call vp_set_identity first)");
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));
  std::vector<float> tmp((size_t)m->rows);
  VP_CUDA(cudaMemcpy(tmp.data(), m->tex, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
  const std::vector<int>& i2o = m->topo.v_int2orig;
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) texture[3 * (size_t)i2o[i] + a] = tmp[3 * i + a];
  return VP_OK;
}

extern "C" int vp_get_base_shape(vp_model* m, double* shape) {
  VP_REQUIRE(m != nullptr && shape != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(m->have_base, "no base shape set (call vp_set_identity first)");
  VP_CUDA(cudaSetDevice(m->device));
  VP_TRY(wait_for_renders(m));
  std::vector<double> tmp((size_t)m->rows);
  VP_CUDA(cudaMemcpy(tmp.data(), m->base, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  const std::vector<int>& i2o = m->topo.v_int2orig;
  for (size_t i = 0; i < i2o.size(); ++i)
    for (int a = 0; a < 3; ++a) shape[3 * (size_t)i2o[i] + a] = tmp[3 * i + a];
  return VP_OK;
}
