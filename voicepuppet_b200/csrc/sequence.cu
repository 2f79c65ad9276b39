// The whole hot path, coefficients -> rendered frames: the batched replacement of the frame
// loop in voicepuppet/pixrefer/infer_bfmvid.py:231-243 (render_face, :79-109, per frame).
// Frames are processed in chunks sized so that every intermediate (displacements, vertex
// records, z-buffer keys, per-triangle colours) stays resident in the 126 MB L2; per chunk the
// launches are K1 basis -> K2 vertex -> K3 scatter -> K4 resolve.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "launch.h"

namespace vp {

namespace {

size_t env_size(const char* name, size_t dflt) {
  const char* s = std::getenv(name);
  if (!s || !*s) return dflt;
  const long long v = std::atoll(s);
  return v > 0 ? (size_t)v : dflt;
}

// Frames per raster chunk.  Measured on B200 (profiles/r01_chunk_sweep.txt): large launches beat
// L2 residency of the intermediates -- one 75-frame chunk at 256x256 (134 MB of intermediates) runs 10 %
// faster than two 38-frame chunks -- so the device-output path takes the largest chunk within
// VPB200_CHUNK_MB (default 320 MB of vertex records + z-buffer keys + per-triangle colours; the round-2 sweep at
// 512x512 and 1024x1024, profiles/r02f_pairs_chunks.txt, confirms it: chunks small enough to keep the z-buffer
// keys in the 126 MB L2 lose more on launch efficiency than they save on HBM traffic).  The
// host-output path cuts the sequence into at least four chunks so that the device->host drain of one
// chunk overlaps the rendering of the next.  Chunks are equal-sized so that no launch runs nearly empty.
int chunk_frames(const vp_model* m, int res, int nframes, bool host_outputs) {
  const size_t per_frame = fused_available(m) ? (size_t)m->vrec_stride * 4 + (size_t)res * res * 8
                                              : (size_t)m->vrec_stride * 16 + (size_t)res * res * 8 + (size_t)m->ntri * 4;
  const size_t forced = env_size("VPB200_CHUNK_FRAMES", 0);
  size_t c = forced ? forced : (env_size("VPB200_CHUNK_MB", 320) << 20) / per_frame;
  c = std::max<size_t>(c, 4);
  c = std::min<size_t>(c, 1024);
  const size_t t = (size_t)std::max(nframes, 1);
  if (forced) return (int)std::min(c, t);
  if (host_outputs) c = std::min(c, std::max<size_t>(8, (t + 3) / 4));
  size_t nchunks = std::max<size_t>(1, (t * 10 + c * 11 - 1) / (c * 11));  // ceil(t / (1.1 c))
  // two chunks on the two streams of ChunkRunner beat one large launch sequence (75 frames at 256x256: 149 vs
  // 156 us): the second chunk's kernels fill the tails of the first's
  if (!host_outputs && nchunks == 1 && t >= 48) nchunks = 2;  // (profiling keeps the same chunks, on one stream)
  return (int)((t + nchunks - 1) / nchunks);
}

// CUDA-event pairs around individual launches, summed per slot when profiling is enabled.
struct Profiler {
  vp_model* m;
  cudaStream_t st;
  struct Span { int slot; cudaEvent_t a, b; };
  std::vector<Span> spans;
  explicit Profiler(vp_model* m_, cudaStream_t st_) : m(m_), st(st_) {}
  void begin(int slot) {
    if (!m->profiling) return;
    Span s{slot, nullptr, nullptr};
    if (cudaEventCreate(&s.a) != cudaSuccess || cudaEventCreate(&s.b) != cudaSuccess) return;
    cudaEventRecord(s.a, st);
    spans.push_back(s);
  }
  void end() {
    if (!m->profiling || spans.empty()) return;
    cudaEventRecord(spans.back().b, st);
  }
  void finish() {
    if (!m->profiling) return;
    cudaStreamSynchronize(st);
    for (int k = 0; k < kProfSlots; ++k) {
      m->prof_ms[k] = 0.f;
      m->prof_launches[k] = 0;
    }
    for (const Span& s : spans) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
        m->prof_ms[s.slot] += ms;
        ++m->prof_launches[s.slot];
      }
      cudaEventDestroy(s.a);
      cudaEventDestroy(s.b);
    }
    spans.clear();
  }
};

// The basis contraction runs over a larger group of frames than the raster chunk: ONE launch per group (the tcgen05
// kernel walks the group in 128-frame blocks, reading the 27 MB basis from HBM once and from L2 afterwards).  Round 2
// measured one launch per 96..128 frames at 0.49..0.53 of the HBM peak by CUDA events, 10 us of each launch being ramp
// and drain of the 148 x 576-thread grid; groups of about 1024 frames amortise that: 0.61 (4096 frames at 1024x1024) /
// 0.62 (12000 at 256x256) per launch, and the step gains 4..8 % (profiles/r02m_basis_blocks.txt).  The displacements of
// such a group, 0.43 MB per frame, no longer stay in L2 for the vertex kernel -- +4 % of the path's HBM bytes at 256x256,
// +2 % at 1024x1024 -- which the LSU-bound vertex kernel does not notice (7.02 -> 6.99 ms per 12000 frames).
int basis_group_frames(int chunk, int nframes) {
  const int cap = (int)env_size("VPB200_BASIS_FRAMES", 1024);
  const int t = std::max(nframes, 1);
  const int ngroups = (t + cap - 1) / cap;
  const int per_group = (t + ngroups - 1) / ngroups;
  return (per_group + chunk - 1) / chunk * chunk;  // a whole number of raster chunks
}

// Issues the chunks of one sequence.  With `dual`, consecutive chunks alternate between the caller's stream
// and the model's auxiliary stream, and between the two halves ("slots") of the chunk workspaces: a chunk is
// three dependent kernels of a few tens of microseconds, each with a ramp-up and a tail during which most
// SMs idle; with two chunks in flight the other chunk's kernels fill those gaps (measured on B200: 25-frame
// chunks cost 66 us each on one stream against 52 us of work).  Everything is joined back onto the caller's
// stream in finish(), so the call stays stream-ordered for the caller.
struct ChunkRunner {
  vp_model* m;
  cudaStream_t st;
  Profiler prof;
  int res = 0, chunk_cap = 0, group = 0, nframes = 0, rotate_first = 0;
  const float* ex_dev = nullptr;
  const FrameParams* params_dev = nullptr;
  const char* fconst = nullptr;
  bool dual = false, aux_used = false, aux_needs_basis = false, fused = false;
  bool planned = false;  // the caller gave a chunk plan and waits on per-chunk events
  int issued = 0;
  int basis_upto = 0;   // frames [group start, basis_upto) of the current basis group are contracted

  ChunkRunner(vp_model* m_, cudaStream_t st_) : m(m_), st(st_), prof(m_, st_) {}

  int init(int nframes_, int res_, int chunk_cap_, int nchunks, const float* ex, const FrameParams* params,
           int rotate_first_, bool allow_dual = true) {
    VP_TRY(order_after_renders(m, st));
    nframes = nframes_;
    res = res_;
    chunk_cap = chunk_cap_;
    ex_dev = ex;
    params_dev = params;
    rotate_first = rotate_first_;
    group = basis_group_frames(chunk_cap, nframes);
    static const int dual_env = [] { const char* e = std::getenv("VPB200_DUAL"); return e ? std::atoi(e) : 1; }();
    // one z-buffer epoch per run() call: a planned chunk that crosses a basis group or exceeds 1024 frames takes
    // several, so budget for the worst case (and run() refuses to go past the limit)
    const uint32_t launches_max = (uint32_t)nchunks + (uint32_t)(nframes / std::max(group, 1)) + (uint32_t)(nframes / 1024) + 4u;
    dual = allow_dual && dual_env != 0 && nchunks >= 2 && !m->profiling && launches_max + 2u < epoch_limit(m->ntri);
    fused = fused_available(m);
    const int slots = dual ? 2 : 1;
    const size_t npix = (size_t)res * res;
    VP_CUDA(m->ws_disp.reserve((size_t)group * m->rows_pad * sizeof(float), m->device));  // >= any group
    if (fused) {
      VP_CUDA(m->ws_vcol.reserve((size_t)slots * chunk_cap * m->vrec_stride * sizeof(uint32_t), m->device));
    } else {
      VP_CUDA(m->ws_vrec.reserve((size_t)slots * chunk_cap * m->vrec_stride * sizeof(float4), m->device));
      VP_CUDA(m->ws_tricol.reserve((size_t)slots * chunk_cap * std::max(m->ntri, 1) * sizeof(uint32_t), m->device));
    }
    void* before = m->ws_keys.ptr;
    VP_CUDA(m->ws_keys.reserve((size_t)slots * chunk_cap * npix * sizeof(unsigned long long), m->device));
    if (m->ws_keys.ptr != before) m->key_epoch = 0;
    // every chunk gets a fresh epoch; stale keys lose every atomicMax and read as background.  The z-buffer is
    // cleared only when the epoch counter would wrap during this call (or the buffer is new)
    if (m->key_epoch == 0 || m->key_epoch + launches_max + 2u >= epoch_limit(m->ntri)) {
      VP_CUDA(cudaMemsetAsync(m->ws_keys.ptr, 0, m->ws_keys.cap, st));
      m->key_epoch = 0;
    }
    const void* fc = nullptr;
    VP_TRY(prepare_frame_constants(m, params_dev, nframes, rotate_first, 1015.0, 112.0, 224.0, (double)res / 224.0, st, &fc));
    fconst = static_cast<const char*>(fc);
    if (dual) VP_CUDA(cudaEventRecord(m->ev_fork, st));
    return VP_OK;
  }

  // frames [t0, t0 + n): n <= chunk_cap and inside one basis group.  *used = the stream the chunk runs on.
  int run(int t0, int n, unsigned char* image_dev, unsigned char* mask_dev, cudaStream_t* used) {
    if (ex_dev) {
      // The expression coefficients of a basis group are contracted on `st` when its first chunk is issued: the whole
      // group in one launch by default, or lazily in pieces of VPB200_BASIS_PIECE frames (only as far as this chunk
      // needs: a short first chunk of the push gather's plan then starts sooner).
      const int g0 = t0 - t0 % group, gend = std::min(g0 + group, nframes);
      if (t0 == g0) {
        if (aux_used) {                 // chunks on the auxiliary stream may still read the previous group
          VP_CUDA(cudaEventRecord(m->ev_aux_done, m->aux_stream));
          VP_CUDA(cudaStreamWaitEvent(st, m->ev_aux_done, 0));
        }
        basis_upto = g0;
      }
      if (basis_upto < t0 + n) {
        static const int piece = (int)env_size("VPB200_BASIS_PIECE", 1 << 20);
        int upto = (int)std::min<long long>(gend, g0 + (long long)(t0 + n - g0 + piece - 1) / piece * piece);
        // a caller with a chunk plan (push gather, notify) wants its first, short chunk out early: contract just the
        // frames that chunk needs first (one 128-frame block), the rest of the first group with the next chunk
        if (planned && g0 == 0 && basis_upto == 0) upto = std::min(upto, (t0 + n + 127) / 128 * 128);
        prof.begin(kProfBasis);
        VP_TRY(launch_basis(m, ex_dev + (size_t)basis_upto * VP_N_EX,
                            m->ws_disp.as<float>() + (size_t)(basis_upto - g0) * m->rows_pad, upto - basis_upto, st));
        prof.end();
        basis_upto = upto;
        if (dual) {
          VP_CUDA(cudaEventRecord(m->ev_basis, st));
          aux_needs_basis = true;
        }
      }
    }
    const int slot = dual ? (issued & 1) : 0;
    cudaStream_t cs = slot ? m->aux_stream : st;
    if (slot) {
      if (!aux_used) VP_CUDA(cudaStreamWaitEvent(cs, m->ev_fork, 0));
      if (aux_needs_basis) VP_CUDA(cudaStreamWaitEvent(cs, m->ev_basis, 0));
      aux_used = true;
      aux_needs_basis = false;
    }
    ++issued;
    VP_REQUIRE(m->key_epoch + 1u <= epoch_limit(m->ntri), "z-buffer epoch counter exhausted (more launches than budgeted)");
    const uint32_t epoch = ++m->key_epoch;
    const size_t npix = (size_t)res * res;
    if (fused) {
      uint32_t* vcol = m->ws_vcol.as<uint32_t>() + (size_t)slot * chunk_cap * m->vrec_stride;
      unsigned long long* fkeys = m->ws_keys.as<unsigned long long>() + (size_t)slot * chunk_cap * npix;
      const float* fdisp = ex_dev ? m->ws_disp.as<float>() + (size_t)(t0 % group) * m->rows_pad : nullptr;
      prof.begin(kProfFused);
      VP_TRY(launch_fused(m, fdisp, n, fconst + (size_t)t0 * frame_constants_stride(), vcol, fkeys, epoch, res, cs));
      prof.end();
      prof.begin(kProfResolve);
      VP_TRY(launch_resolve_vcol(m, fkeys, vcol, epoch, image_dev, mask_dev, n, res, res, cs));
      prof.end();
      if (used) *used = cs;
      return VP_OK;
    }
    float4* vrec = m->ws_vrec.as<float4>() + (size_t)slot * chunk_cap * m->vrec_stride;
    unsigned long long* keys = m->ws_keys.as<unsigned long long>() + (size_t)slot * chunk_cap * npix;
    uint32_t* tricol = m->ws_tricol.as<uint32_t>() + (size_t)slot * chunk_cap * std::max(m->ntri, 1);
    const float* disp = ex_dev ? m->ws_disp.as<float>() + (size_t)(t0 % group) * m->rows_pad : nullptr;
    ReconOut none;
    prof.begin(kProfVertex);
    VP_TRY(launch_vertex(m, disp, params_dev + t0, n, rotate_first, 1015.0, 112.0, 224.0, (double)res / 224.0, vrec,
                         none, cs, fconst + (size_t)t0 * frame_constants_stride()));
    prof.end();
    prof.begin(kProfScatter);
    VP_TRY(launch_scatter_packed(vrec, (size_t)m->vrec_stride, m->tri, keys, tricol, epoch, n, m->ntri, res, res, cs));
    prof.end();
    prof.begin(kProfResolve);
    VP_TRY(launch_resolve_packed(keys, tricol, epoch, image_dev, mask_dev, n, m->ntri, res, res, cs));
    prof.end();
    if (used) *used = cs;
    return VP_OK;
  }

  // the widest chunk that starts at t0, is at most `want` frames and does not cross a basis group
  int clip(int t0, int want) const { return std::min(std::min(want, chunk_cap), group - t0 % group); }

  int finish(int rc) {
    if (aux_used) {  // join: work the caller enqueues on `st` after this call sees every frame
      cudaError_t e = cudaEventRecord(m->ev_aux_done, m->aux_stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, m->ev_aux_done, 0);
      if (e != cudaSuccess && rc == VP_OK) {
        set_error("joining the auxiliary stream failed: %s", cudaGetErrorString(e));
        rc = VP_ERR_CUDA;
      }
    }
    if (rc != VP_OK) m->key_epoch = 0;
    note_render(m, st);
    prof.finish();
    return rc;
  }
};

int check_sequence_args(const vp_model* m, int nframes, int res, const void* image) {
  VP_REQUIRE(m != nullptr, "null model");
  VP_REQUIRE(nframes >= 0, "nframes >= 0");
  VP_REQUIRE(res >= 2 && res <= 8192 && res % 2 == 0, "res must be even and in 2..8192");
  VP_REQUIRE(nframes == 0 || image != nullptr, "null image buffer");
  return VP_OK;
}

}  // namespace
}  // namespace vp

using namespace vp;

// plan == nullptr: uniform chunks chosen by chunk_frames().  Otherwise plan[0..nplan) are the chunk sizes
// (sum == nframes) and events[i] is recorded when chunk i is complete; a planned chunk that crosses a basis
// group boundary is rendered in two launches.
static int render_sequence_dev_impl(vp_model* m, int nframes, const float* ex_dev, const vp_frame_params* params_dev,
                                    int rotate_shape_first, int res, unsigned char* image_dev,
                                    unsigned char* face_mask_dev, void* stream, const int* plan, int nplan,
                                    void** events) {
  VP_TRY(check_sequence_args(m, nframes, res, image_dev));
  VP_REQUIRE(nframes == 0 || params_dev != nullptr, "null params");
  if (nframes == 0) return VP_OK;
  int plan_max = 0;
  if (plan) {
    VP_REQUIRE(nplan > 0 && events != nullptr, "bad chunk plan");
    long long sum = 0;
    for (int i = 0; i < nplan; ++i) {
      VP_REQUIRE(plan[i] > 0, "chunk sizes must be positive");
      sum += plan[i];
      plan_max = std::max(plan_max, plan[i]);
    }
    VP_REQUIRE(sum == nframes, "chunk sizes must add up to nframes");
  }
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(m->have_base && m->have_tex, "no identity set (call vp_set_identity first)");
  VP_CUDA(cudaSetDevice(m->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunk = plan ? std::min(plan_max, 1024) : chunk_frames(m, res, nframes, false);
  const int nchunks = plan ? nplan : (nframes + chunk - 1) / chunk;
  const size_t npix = (size_t)res * res;
  ChunkRunner run(m, st);
  run.planned = plan != nullptr;
  int rc = run.init(nframes, res, chunk, nchunks, ex_dev, reinterpret_cast<const FrameParams*>(params_dev),
                    rotate_shape_first);
  int t0 = 0;
  for (int ci = 0; ci < nchunks && rc == VP_OK; ++ci) {
    int left = plan ? plan[ci] : std::min(chunk, nframes - t0);
    bool on_main = false, on_aux = false;
    while (left > 0 && rc == VP_OK) {  // more than one launch only when the chunk crosses a basis group
      const int n = run.clip(t0, left);
      cudaStream_t used = st;
      rc = run.run(t0, n, image_dev + (size_t)t0 * npix * 3, face_mask_dev ? face_mask_dev + (size_t)t0 * npix : nullptr,
                   &used);
      (used == st ? on_main : on_aux) = true;
      t0 += n;
      left -= n;
    }
    if (rc == VP_OK && plan) {  // events[ci]: recorded where it covers every part of the chunk
      if (on_aux && on_main) {
        VP_CUDA(cudaEventRecord(m->ev_main_done, m->aux_stream));
        VP_CUDA(cudaStreamWaitEvent(st, m->ev_main_done, 0));
      }
      VP_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(events[ci]), (on_aux && !on_main) ? m->aux_stream : st));
    }
  }
  return run.finish(rc);
}

extern "C" int vp_render_sequence_dev(vp_model* m, int nframes, const float* ex_dev, const vp_frame_params* params_dev,
                                      int rotate_shape_first, int res, unsigned char* image_dev,
                                      unsigned char* face_mask_dev, void* stream) {
  return render_sequence_dev_impl(m, nframes, ex_dev, params_dev, rotate_shape_first, res, image_dev, face_mask_dev,
                                  stream, nullptr, 0, nullptr);
}

// Same, rendered in chunks of `notify_frames` frames with events[i] (cudaEvent_t) recorded on `stream`
// as soon as chunk i is complete: lets the caller start moving finished frames (NCCL gather, D2H copy)
// while the rest of the sequence is still rendering.  The basis contraction still runs once per group.
extern "C" int vp_render_sequence_dev_notify(vp_model* m, int nframes, const float* ex_dev,
                                             const vp_frame_params* params_dev, int rotate_shape_first, int res,
                                             unsigned char* image_dev, unsigned char* face_mask_dev, void* stream,
                                             int notify_frames, void** events, int nevents) {
  VP_REQUIRE(notify_frames > 0, "notify_frames must be positive");
  VP_REQUIRE(nframes >= 0 && (nframes + notify_frames - 1) / notify_frames <= nevents, "not enough events for the chunks");
  std::vector<int> plan;
  for (int t0 = 0; t0 < nframes; t0 += notify_frames) plan.push_back(std::min(notify_frames, nframes - t0));
  if (plan.empty()) return VP_OK;
  return render_sequence_dev_impl(m, nframes, ex_dev, params_dev, rotate_shape_first, res, image_dev, face_mask_dev,
                                  stream, plan.data(), (int)plan.size(), events);
}

// Same with an explicit chunk plan: chunk_frames[i] frames in chunk i (they add up to nframes), events[i]
// recorded when chunk i is complete.  Lets a caller whose consumer is the slow side (the NVLink ingest of
// rank 0 in a multi-GPU gather) start it early with a short first chunk and end it with a short last one.
extern "C" int vp_render_sequence_dev_chunks(vp_model* m, int nframes, const float* ex_dev,
                                             const vp_frame_params* params_dev, int rotate_shape_first, int res,
                                             unsigned char* image_dev, unsigned char* face_mask_dev, void* stream,
                                             const int* chunk_frames, int nchunks, void** events) {
  VP_REQUIRE(chunk_frames != nullptr && nchunks > 0 && events != nullptr, "bad chunk plan");
  return render_sequence_dev_impl(m, nframes, ex_dev, params_dev, rotate_shape_first, res, image_dev, face_mask_dev,
                                  stream, chunk_frames, nchunks, events);
}

extern "C" int vp_render_sequence(vp_model* m, const vp_frames* fr, int res, unsigned char* image,
                                  unsigned char* face_mask, int outputs_on_device, void* stream) {
  VP_REQUIRE(fr != nullptr, "null frames");
  VP_TRY(check_sequence_args(m, fr->nframes, res, image));
  const int T = fr->nframes;
  if (T == 0) return VP_OK;
  VP_REQUIRE(fr->rotation && fr->translation && fr->gamma, "null per-frame array");
  VP_REQUIRE(fr->focal == 1015.0 && fr->center == 112.0, "the fused path renders with focal 1015 / center 112");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_REQUIRE(m->have_base && m->have_tex, "no identity set (call vp_set_identity first)");
  VP_CUDA(cudaSetDevice(m->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VP_TRY(order_after_renders(m, st));   // the uploads below rewrite workspaces a render on another stream may still read

  // per-frame inputs: one upload for the whole sequence (T * 448 bytes) out of page-locked staging, so
  // that the copies are asynchronous; the staging is reused only after the previous call's uploads ran
  const size_t params_bytes = (size_t)T * sizeof(FrameParams), ex_bytes = fr->ex ? (size_t)T * VP_N_EX * sizeof(float) : 0;
  if (m->h_stage_cap < params_bytes + ex_bytes) {
    if (m->h_stage) {
      VP_CUDA(cudaEventSynchronize(m->ev_stage));
      cudaFreeHost(m->h_stage);
      m->h_stage = nullptr;
      m->h_stage_cap = 0;
    }
    const size_t want = std::max<size_t>(2 * (params_bytes + ex_bytes), 1 << 16);
    VP_CUDA(cudaHostAlloc(&m->h_stage, want, cudaHostAllocDefault));
    m->h_stage_cap = want;
  } else {
    VP_CUDA(cudaEventSynchronize(m->ev_stage));
  }
  FrameParams* hp = static_cast<FrameParams*>(m->h_stage);
  for (int t = 0; t < T; ++t) {
    for (int k = 0; k < 9; ++k) hp[t].rot[k] = fr->rotation[9 * (size_t)t + k];
    for (int k = 0; k < 3; ++k) hp[t].trans[k] = fr->translation[3 * (size_t)t + k];
    for (int k = 0; k < VP_N_GAMMA; ++k) hp[t].gamma[k] = fr->gamma[VP_N_GAMMA * (size_t)t + k];
  }
  VP_CUDA(m->ws_params.reserve(params_bytes, m->device));
  VP_CUDA(cudaMemcpyAsync(m->ws_params.ptr, hp, params_bytes, cudaMemcpyHostToDevice, st));
  if (fr->ex) {
    float* hex = reinterpret_cast<float*>(static_cast<char*>(m->h_stage) + params_bytes);
    std::memcpy(hex, fr->ex, ex_bytes);
    VP_CUDA(m->ws_ex.reserve(ex_bytes, m->device));
    VP_CUDA(cudaMemcpyAsync(m->ws_ex.ptr, hex, ex_bytes, cudaMemcpyHostToDevice, st));
  }
  VP_CUDA(cudaEventRecord(m->ev_stage, st));
  const float* ex_dev = fr->ex ? m->ws_ex.as<float>() : nullptr;
  const FrameParams* params_dev = m->ws_params.as<FrameParams>();

  const int chunk = chunk_frames(m, res, T, outputs_on_device == 0);
  const bool forced_chunk = env_size("VPB200_CHUNK_FRAMES", 0) != 0;
  const size_t npix = (size_t)res * res;
  // The drain over PCIe (about 3.5 us per 256x256 frame) is the slow side of the host-output pipeline, so what
  // matters is how soon it starts: the first chunk is short, the following ones full-sized.
  const int first = (outputs_on_device || forced_chunk || T < 4 * 8) ? chunk : std::max(6, chunk / 3);
  const int nchunks_est = 2 + (T + chunk - 1) / chunk + T / 96;
  ChunkRunner run(m, st);
  // host outputs: the PCIe drain is the slow side and wants the FIRST chunk as early as possible, which two
  // chunks in flight delay (measured: 454 vs 417 us per 75-frame call), so that path stays on one stream
  int rc = run.init(T, res, chunk, nchunks_est, ex_dev, params_dev, fr->rotate_shape_first, outputs_on_device != 0);
  if (rc != VP_OK) return run.finish(rc);
  if (outputs_on_device) {
    for (int t0 = 0, n = 0; t0 < T && rc == VP_OK; t0 += n) {
      n = run.clip(t0, T - t0);
      rc = run.run(t0, n, image + (size_t)t0 * npix * 3, face_mask ? face_mask + (size_t)t0 * npix : nullptr, nullptr);
    }
    // device outputs: asynchronous, ordered on `stream` like any other work the caller enqueues there
    return run.finish(rc);
  }
  // host outputs, double-buffered: chunk i renders while chunk i-1 drains to the host on the copy stream
  for (int b = 0; b < 2; ++b) {
    VP_CUDA(m->ws_img[b].reserve((size_t)chunk * npix * 3, m->device));
    if (face_mask) VP_CUDA(m->ws_mask[b].reserve((size_t)chunk * npix, m->device));
  }
  int ci = 0;
  for (int t0 = 0, n = 0; t0 < T && rc == VP_OK; t0 += n, ++ci) {
    n = run.clip(t0, std::min(t0 == 0 ? first : chunk, T - t0));
    const int b = ci & 1;
    // the staging image of chunk ci - 2 must have drained; the wait goes on the stream chunk ci will use, which
    // is the one chunk ci - 2 used (chunks alternate), or the only one
    cudaStream_t cs = (run.dual && (run.issued & 1)) ? m->aux_stream : st;
    if (ci >= 2) VP_CUDA(cudaStreamWaitEvent(cs, m->ev_copy[b], 0));
    cudaStream_t used = st;
    rc = run.run(t0, n, m->ws_img[b].as<unsigned char>(), face_mask ? m->ws_mask[b].as<unsigned char>() : nullptr, &used);
    if (rc != VP_OK) break;
    VP_CUDA(cudaEventRecord(m->ev_render[b], used));
    VP_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_render[b], 0));
    VP_CUDA(cudaMemcpyAsync(image + (size_t)t0 * npix * 3, m->ws_img[b].ptr, (size_t)n * npix * 3, cudaMemcpyDeviceToHost,
                            m->copy_stream));
    if (face_mask)
      VP_CUDA(cudaMemcpyAsync(face_mask + (size_t)t0 * npix, m->ws_mask[b].ptr, (size_t)n * npix, cudaMemcpyDeviceToHost,
                              m->copy_stream));
    VP_CUDA(cudaEventRecord(m->ev_copy[b], m->copy_stream));
  }
  rc = run.finish(rc);
  cudaError_t e1 = cudaStreamSynchronize(st);
  cudaError_t e2 = cudaStreamSynchronize(m->copy_stream);
  if (rc == VP_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
    set_error("stream synchronize failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    rc = VP_ERR_CUDA;
  }
  return rc;
}

// The expression contraction on its own, device pointers, nothing synchronised:
// disp[t][r] = sum_k exBase[r][k] * ex[t][k] in the library's internal row order (r = 3 * internal
// vertex + axis, vp_model_rows_pad() floats per frame).  Used by the micro-benchmarks and tests.
extern "C" int vp_basis_dev(vp_model* m, const float* ex_dev, float* disp_dev, int nframes, void* stream) {
  VP_REQUIRE(m != nullptr && ex_dev != nullptr && disp_dev != nullptr && nframes >= 0, "bad argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  return launch_basis(m, ex_dev, disp_dev, nframes, static_cast<cudaStream_t>(stream));
}

// Diagnostics: one tcgen05 basis launch with a clock64() timeline of CTA 0 written to trace_dev[0..256)
// (4 roles x 16 tiles x 4 marks; see VP_TRACE in basis_tc.cu) and the %globaltimer entry / exit time (ns) of CTA c
// in trace_dev[256 + 2 c], trace_dev[257 + 2 c]; trace_dev holds 1024 entries.
extern "C" int vp_debug_basis_trace(vp_model* m, const float* ex_dev, float* disp_dev, int nframes,
                                    long long* trace_dev, void* stream) {
  VP_REQUIRE(m != nullptr && ex_dev && disp_dev && trace_dev && nframes > 0 && nframes <= 128, "bad argument");
  std::lock_guard<std::mutex> lock(m->mu);
  VP_CUDA(cudaSetDevice(m->device));
  return launch_basis_tc(m, ex_dev, disp_dev, nframes, static_cast<cudaStream_t>(stream), trace_dev);
}

extern "C" int vp_model_rows_pad(const vp_model* m) { return m ? m->rows_pad : -1; }

extern "C" int vp_set_profiling(vp_model* m, int enabled) {
  VP_REQUIRE(m != nullptr, "null model");
  std::lock_guard<std::mutex> lock(m->mu);
  m->profiling = enabled != 0;
  return VP_OK;
}

extern "C" int vp_get_profile_launches(vp_model* m, int* launches, int cap) {
  VP_REQUIRE(m != nullptr && launches != nullptr, "null argument");
  std::lock_guard<std::mutex> lock(m->mu);
  for (int k = 0; k < kProfSlots && k < cap; ++k) launches[k] = m->prof_launches[k];
  return VP_OK;
}

extern "C" int vp_get_profile(vp_model* m, char* names, int names_cap, float* ms, int ms_cap) {
  VP_REQUIRE(m != nullptr, "null model");
  std::lock_guard<std::mutex> lock(m->mu);
  static const char kNames[] = "basis;vertex;scatter;resolve;fused";
  if (names && names_cap > 0) {
    std::snprintf(names, (size_t)names_cap, "%s", kNames);
  }
  for (int k = 0; k < kProfSlots && k < ms_cap; ++k)
    if (ms) ms[k] = m->prof_ms[k];
  return VP_OK;
}
