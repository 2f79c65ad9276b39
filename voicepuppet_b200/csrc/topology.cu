// One-off mesh analysis behind vp_model_create (host only, no CUDA calls).
//
// The reference recomputes vertex normals per frame with two fancy-index gathers over the
// whole mesh (utils/reconstruct_mesh.py:35-52: shape[tri] and face_norm[point_buf]).  Here the
// adjacency is compiled once into vertex TILES: up to 128 spatially adjacent vertices, the
// triangles their point_buf rows name and the halo vertices those triangles touch, all with
// tile-local indices, so that the per-frame vertex kernel works out of shared memory.
//   * vertices are renumbered along a Morton curve of the mean shape's (x, y);
//   * a tile is a run of consecutive renumbered vertices, grown greedily until it would
//     exceed 128 own vertices, kTileLV local vertices or kTileLT local triangles;
//   * point_buf slot order is preserved (the reference sums the ring in column order);
//   * triangles are renumbered by their smallest renumbered vertex for the rasterizer, and
//     keep their ORIGINAL index for the z-buffer tie-break;
//   * where every face a vertex's point_buf row lists really contains the vertex and the faces chain
//     into at most 9 ring vertices (any manifold mesh: closed or open fans up to valence 8), the tile
//     also gets FAN records (launch.h), which let the vertex kernel skip the per-triangle pass.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <numeric>

#include "launch.h"

namespace vp {

namespace {

uint32_t spread16(uint32_t v) {
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

uint32_t quantize(double v, double lo, double hi) {
  if (!(hi > lo) || !(v == v)) return 0;
  double q = (v - lo) / (hi - lo) * 65535.0;
  if (q < 0) q = 0;
  if (q > 65535.0) q = 65535.0;
  return (uint32_t)q;
}


// ---- bank-conflict-aware shared-memory slots (optional) ---------------------------------------
// For every fan tile: the "access groups" are the sets of distinct local vertices the 8 lanes of a quarter-warp
// read at one fan step.  An LDS.128 serialises distinct slots of one bank group (slot % 8), so the local
// vertices are 8-coloured greedily (most constrained first, balanced colours), refined by swaps that do not
// increase the number of excess wavefronts, and slot = colour + 8 * rank within the colour.
void assign_slots(Topology& t) {
  const int ntiles = (int)t.tiles.size();
  t.slot_off.assign(ntiles, -1);
  t.slot_tab.clear();
  t.fan_slot.assign(t.fan.size(), 0u);
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  auto next = [&rng]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
  for (int ti = 0; ti < ntiles; ++ti) {
    const TileDesc& td = t.tiles[ti];
    if (!td.fan) continue;
    const int nlv = td.nlv, nv = td.nv;
    auto entry = [&](int v, int i) {  // local index of fan entry i of own vertex v
      const uint32_t* w = &t.fan[(size_t)(td.v_begin + v) * kFanWords];
      const uint32_t word = w[i >> 1];
      return (int)((((i & 1) ? (word >> 16) : word) & 0xFFFFu) >> 4);
    };
    // access groups
    std::vector<std::vector<int>> groups;
    for (int i = 0; i < kFanEntries; ++i)
      for (int q = 0; q * 8 < nv; ++q) {
        std::vector<int> g;
        for (int v = q * 8; v < std::min(nv, q * 8 + 8); ++v) {
          const int x = entry(v, i);
          if (std::find(g.begin(), g.end(), x) == g.end()) g.push_back(x);
        }
        if (g.size() > 1) groups.push_back(g);
      }
    std::vector<std::vector<int>> member(nlv);
    for (int gi = 0; gi < (int)groups.size(); ++gi)
      for (int x : groups[gi]) member[x].push_back(gi);
    const int cap = (nlv + 7) / 8;
    std::vector<int> cls(nlv, -1), cnt(8, 0), order(nlv);
    std::vector<std::array<int, 8>> gc(groups.size());
    for (auto& a : gc) a.fill(0);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return member[a].size() > member[b].size(); });
    auto gmax = [&](int gi) { return *std::max_element(gc[gi].begin(), gc[gi].end()); };
    for (int x : order) {
      int best_c = -1, best_add = 1 << 30, best_cnt = 1 << 30;
      for (int c = 0; c < 8; ++c) {
        if (cnt[c] >= cap) continue;
        int add = 0;
        for (int gi : member[x]) add += (gc[gi][c] + 1 > std::max(1, gmax(gi)));
        if (add < best_add || (add == best_add && cnt[c] < best_cnt)) {
          best_c = c;
          best_add = add;
          best_cnt = cnt[c];
        }
      }
      cls[x] = best_c;
      ++cnt[best_c];
      for (int gi : member[x]) ++gc[gi][best_c];
    }
    // refinement: swap the colours of two vertices when the affected groups do not get worse
    auto excess_of = [&](const std::vector<int>& gis) {
      int e = 0;
      for (int gi : gis) e += gmax(gi) - 1;
      return e;
    };
    for (int it = 0; it < 1500 && nlv > 1; ++it) {
      const int a = (int)(next() % (uint32_t)nlv), b = (int)(next() % (uint32_t)nlv);
      if (cls[a] == cls[b]) continue;
      std::vector<int> aff = member[a];
      for (int gi : member[b])
        if (std::find(aff.begin(), aff.end(), gi) == aff.end()) aff.push_back(gi);
      const int before = excess_of(aff);
      auto move = [&](int x, int from, int to) {
        for (int gi : member[x]) {
          --gc[gi][from];
          ++gc[gi][to];
        }
      };
      const int ca = cls[a], cb = cls[b];
      move(a, ca, cb);
      move(b, cb, ca);
      if (excess_of(aff) <= before) {
        cls[a] = cb;
        cls[b] = ca;
      } else {
        move(a, cb, ca);
        move(b, ca, cb);
      }
    }
    // slots and the fan records in slot space
    t.slot_off[ti] = (int)t.slot_tab.size();
    std::vector<int> rank(8, 0), slot(nlv);
    for (int x = 0; x < nlv; ++x) {
      slot[x] = cls[x] + 8 * rank[cls[x]]++;
      t.slot_tab.push_back((uint16_t)slot[x]);
    }
    for (int v = 0; v < nv; ++v) {
      const uint32_t* w = &t.fan[(size_t)(td.v_begin + v) * kFanWords];
      uint32_t* o = &t.fan_slot[(size_t)(td.v_begin + v) * kFanWords];
      for (int k = 0; k < 4; ++k) o[k] = ((uint32_t)slot[entry(v, 2 * k)] << 4) | ((uint32_t)slot[entry(v, 2 * k + 1)] << 20);
      o[4] = ((uint32_t)slot[entry(v, 8)] << 4) | (w[4] & 0xFFFF0000u);
    }
  }
}

}  // namespace

int build_topology(Topology& out, int nver, int ntri, const int* tri, const int* point_buf, const double* xyz,
                   bool with_slots) {
  VP_REQUIRE(nver >= 0 && ntri >= 0, "negative element count");
  VP_REQUIRE(nver == 0 || (point_buf && xyz), "null model array");
  VP_REQUIRE(ntri == 0 || tri, "null triangle array");
  for (size_t i = 0; i < (size_t)ntri * 3; ++i)
    VP_REQUIRE(tri[i] >= 0 && tri[i] < nver, "triangle index out of range");

  out = Topology();
  out.nver = nver;
  out.ntri = ntri;

  // ---- Morton order of the vertices ------------------------------------------------------
  double lo[2] = {INFINITY, INFINITY}, hi[2] = {-INFINITY, -INFINITY};
  for (int v = 0; v < nver; ++v)
    for (int a = 0; a < 2; ++a) {
      const double c = xyz[3 * (size_t)v + a];
      if (c == c && std::fabs(c) != INFINITY) {
        lo[a] = std::min(lo[a], c);
        hi[a] = std::max(hi[a], c);
      }
    }
  std::vector<uint32_t> code(nver);
  for (int v = 0; v < nver; ++v)
    code[v] = spread16(quantize(xyz[3 * (size_t)v], lo[0], hi[0])) |
              (spread16(quantize(xyz[3 * (size_t)v + 1], lo[1], hi[1])) << 1);
  out.v_int2orig.resize(nver);
  std::iota(out.v_int2orig.begin(), out.v_int2orig.end(), 0);
  std::stable_sort(out.v_int2orig.begin(), out.v_int2orig.end(),
                   [&](int a, int b) { return code[a] < code[b]; });
  out.v_orig2int.resize(nver);
  for (int i = 0; i < nver; ++i) out.v_orig2int[out.v_int2orig[i]] = i;

  // ---- triangles for the rasterizer ------------------------------------------------------
  std::vector<int> t_order(ntri);
  std::iota(t_order.begin(), t_order.end(), 0);
  std::vector<int> t_key(ntri);
  for (int f = 0; f < ntri; ++f)
    t_key[f] = std::min(out.v_orig2int[tri[3 * (size_t)f]],
                        std::min(out.v_orig2int[tri[3 * (size_t)f + 1]], out.v_orig2int[tri[3 * (size_t)f + 2]]));
  std::stable_sort(t_order.begin(), t_order.end(), [&](int a, int b) { return t_key[a] < t_key[b]; });
  out.tri_int.resize((size_t)ntri * 4);
  for (int i = 0; i < ntri; ++i) {
    const int f = t_order[i];
    out.tri_int[4 * (size_t)i + 0] = out.v_orig2int[tri[3 * (size_t)f]];
    out.tri_int[4 * (size_t)i + 1] = out.v_orig2int[tri[3 * (size_t)f + 1]];
    out.tri_int[4 * (size_t)i + 2] = out.v_orig2int[tri[3 * (size_t)f + 2]];
    out.tri_int[4 * (size_t)i + 3] = f;
  }

  // first tri_int row whose smallest vertex is >= v (rows are sorted by it)
  std::vector<int> first_row(nver + 1, ntri);
  for (int i = ntri - 1; i >= 0; --i) first_row[t_key[t_order[i]]] = i;
  for (int iv = nver - 1; iv >= 0; --iv) first_row[iv] = std::min(first_row[iv], first_row[iv + 1]);
  out.own_ltri.assign((size_t)ntri, 0u);
  out.own_tri_off.clear();
  out.fused_ok = true;
  out.tri_by_orig.assign((size_t)ntri * 4, 0);
  for (int f = 0; f < ntri; ++f)
    for (int c = 0; c < 3; ++c) out.tri_by_orig[4 * (size_t)f + c] = out.v_orig2int[tri[3 * (size_t)f + c]];

  // ---- vertex tiles ----------------------------------------------------------------------
  out.ring.assign((size_t)nver * VP_RING, kRingPad);
  out.fan.assign((size_t)nver * kFanWords, 0u);
  std::vector<int> tri_stamp(ntri, -1), tri_local(ntri, 0);   // tile id that last saw the triangle
  std::vector<int> ver_stamp(nver, -1), ver_local(nver, 0);   // ... the (internal) vertex
  std::vector<int> cur_tris, cur_touched;                     // in first-seen order
  int v = 0;
  while (v < nver) {
    const int tile_id = (int)out.tiles.size();
    TileDesc td;
    td.v_begin = v;
    td.nv = 0;
    cur_tris.clear();
    cur_touched.clear();
    int n_touched = 0;  // distinct internal vertices that are own or touched
    while (v < nver && td.nv < kTileV) {
      // what would adding internal vertex v cost?
      const int ov = out.v_int2orig[v];
      int new_t[VP_RING], n_new_t = 0;
      int new_v[3 * VP_RING + 1], n_new_v = 0;
      auto want_vertex = [&](int iv) {
        if (ver_stamp[iv] == tile_id) return;
        for (int k = 0; k < n_new_v; ++k)
          if (new_v[k] == iv) return;
        new_v[n_new_v++] = iv;
      };
      want_vertex(v);
      for (int s = 0; s < VP_RING; ++s) {
        const int f = point_buf[(size_t)ov * VP_RING + s];
        if (f < 0 || f >= ntri || tri_stamp[f] == tile_id) continue;
        bool dup = false;
        for (int k = 0; k < n_new_t; ++k) dup |= (new_t[k] == f);
        if (dup) continue;
        new_t[n_new_t++] = f;
        for (int c = 0; c < 3; ++c) want_vertex(out.v_orig2int[tri[3 * (size_t)f + c]]);
      }
      if (td.nv > 0 && ((int)cur_tris.size() + n_new_t > kTileLT || n_touched + n_new_v > kTileLV)) break;
      for (int k = 0; k < n_new_t; ++k) {
        tri_stamp[new_t[k]] = tile_id;
        tri_local[new_t[k]] = (int)cur_tris.size();
        cur_tris.push_back(new_t[k]);
      }
      for (int k = 0; k < n_new_v; ++k) {
        ver_stamp[new_v[k]] = tile_id;
        cur_touched.push_back(new_v[k]);
      }
      n_touched += n_new_v;
      for (int s = 0; s < VP_RING; ++s) {
        const int f = point_buf[(size_t)ov * VP_RING + s];
        if (f >= 0 && f < ntri) out.ring[(size_t)v * VP_RING + s] = (uint16_t)tri_local[f];
      }
      ++td.nv;
      ++v;
    }
    // local numbering: own vertices 0..nv-1 in internal order, then the halo in first-seen order
    td.halo_off = (int)out.halo.size();
    int next = td.nv;
    for (int iv : cur_touched) {
      if (iv >= td.v_begin && iv < td.v_begin + td.nv) {
        ver_local[iv] = iv - td.v_begin;
      } else {
        ver_local[iv] = next++;
        out.halo.push_back(iv);
      }
    }
    td.nlv = next;
    td.nlt = (int)cur_tris.size();
    // ---- fan records ---------------------------------------------------------------------
    td.fan = 1;
    for (int iv = td.v_begin; iv < td.v_begin + td.nv && td.fan; ++iv) {
      const int ov = out.v_int2orig[iv];
      int fp[VP_RING], fq[VP_RING], nf = 0;  // the other two corners (tile-local) of every listed face
      for (int s = 0; s < VP_RING && td.fan; ++s) {
        const int f = point_buf[(size_t)ov * VP_RING + s];
        if (f < 0 || f >= ntri) continue;
        int k = -1;
        for (int c = 2; c >= 0; --c)
          if (tri[3 * (size_t)f + c] == ov) k = c;
        if (k < 0) {  // point_buf names a face that does not contain the vertex: generic path
          td.fan = 0;
          break;
        }
        fp[nf] = ver_local[out.v_orig2int[tri[3 * (size_t)f + (k + 1) % 3]]];
        fq[nf] = ver_local[out.v_orig2int[tri[3 * (size_t)f + (k + 2) % 3]]];
        ++nf;
      }
      if (!td.fan) break;
      int entries[2 * VP_RING + 2], ne = 0;
      uint32_t mask = 0;
      bool used[VP_RING] = {false, false, false, false, false, false, false, false};
      for (int done = 0; done < nf && ne <= kFanEntries;) {
        int start = -1;  // a face no other unused face leads into, else the first unused one
        for (int i = 0; i < nf && start < 0; ++i) {
          if (used[i]) continue;
          bool led = false;
          for (int j = 0; j < nf; ++j) led |= (!used[j] && j != i && fq[j] == fp[i]);
          if (!led) start = i;
        }
        for (int i = 0; i < nf && start < 0; ++i)
          if (!used[i]) start = i;
        entries[ne++] = fp[start];
        for (int cur = start; cur >= 0 && ne <= kFanEntries;) {
          mask |= 1u << (ne - 1);
          entries[ne++] = fq[cur];
          used[cur] = true;
          ++done;
          int nxt = -1;
          for (int j = 0; j < nf && nxt < 0; ++j)
            if (!used[j] && fp[j] == fq[cur]) nxt = j;
          cur = nxt;
        }
      }
      if (ne > kFanEntries) {
        td.fan = 0;
        break;
      }
      const int self = iv - td.v_begin;
      while (ne < kFanEntries) entries[ne++] = self;
      uint32_t* w = &out.fan[(size_t)iv * kFanWords];
      for (int k = 0; k < 4; ++k) w[k] = ((uint32_t)entries[2 * k] << 4) | ((uint32_t)entries[2 * k + 1] << 20);
      w[4] = ((uint32_t)entries[8] << 4) | (mask << 16);
    }
    if (!td.fan)
      for (int iv = td.v_begin; iv < td.v_begin + td.nv; ++iv)
        for (int k = 0; k < kFanWords; ++k) out.fan[(size_t)iv * kFanWords + k] = 0;
    td.ltri_off = (int)out.ltri.size();
    for (int f : cur_tris) {
      const uint32_t a = (uint32_t)ver_local[out.v_orig2int[tri[3 * (size_t)f]]];
      const uint32_t b = (uint32_t)ver_local[out.v_orig2int[tri[3 * (size_t)f + 1]]];
      const uint32_t c = (uint32_t)ver_local[out.v_orig2int[tri[3 * (size_t)f + 2]]];
      out.ltri.push_back(a | (b << 10) | (c << 20));
    }
    // ---- owned triangles (fused kernel): corners in this tile's local numbering --------------
    out.own_tri_off.push_back(first_row[td.v_begin]);
    for (int i = first_row[td.v_begin]; i < first_row[td.v_begin + td.nv]; ++i) {
      uint32_t packed = 0;
      for (int c = 0; c < 3; ++c) {
        const int iv = out.tri_int[4 * (size_t)i + c];
        if (ver_stamp[iv] != tile_id) out.fused_ok = false;  // a corner the tile never staged (inconsistent point_buf)
        else packed |= (uint32_t)ver_local[iv] << (10 * c);
      }
      out.own_ltri[i] = packed;
    }
    if (!td.fan || first_row[td.v_begin + td.nv] - first_row[td.v_begin] > kTileLT) out.fused_ok = false;
    out.tiles.push_back(td);
  }
  out.own_tri_off.push_back(ntri);
  if (with_slots) assign_slots(out);
  return VP_OK;
}

}  // namespace vp

// ---- introspection entry points (host only; used by the CPU test-suite) --------------------
struct vp_topology {
  vp::Topology t;
};

extern "C" int vp_topology_build(vp_topology** out, int nver, int ntri, const int* tri, const int* point_buf,
                                 const double* xyz) {
  VP_REQUIRE(out != nullptr, "null out pointer");
  *out = nullptr;
  vp_topology* h = new vp_topology();
  const int rc = vp::build_topology(h->t, nver, ntri, tri, point_buf, xyz, true);
  if (rc != VP_OK) {
    delete h;
    return rc;
  }
  *out = h;
  return VP_OK;
}

extern "C" void vp_topology_destroy(vp_topology* h) { delete h; }

extern "C" int vp_topology_sizes(const vp_topology* h, int* ntiles, int* nltri, int* nhalo) {
  VP_REQUIRE(h != nullptr, "null handle");
  if (ntiles) *ntiles = (int)h->t.tiles.size();
  if (nltri) *nltri = (int)h->t.ltri.size();
  if (nhalo) *nhalo = (int)h->t.halo.size();
  return VP_OK;
}

extern "C" int vp_topology_copy(const vp_topology* h, int* v_int2orig, int* tri_int, int* tiles, uint32_t* ltri,
                                int* halo, uint16_t* ring, uint32_t* fan) {
  VP_REQUIRE(h != nullptr, "null handle");
  const vp::Topology& t = h->t;
  if (v_int2orig) std::memcpy(v_int2orig, t.v_int2orig.data(), t.v_int2orig.size() * sizeof(int));
  if (tri_int) std::memcpy(tri_int, t.tri_int.data(), t.tri_int.size() * sizeof(int));
  if (tiles) std::memcpy(tiles, t.tiles.data(), t.tiles.size() * sizeof(vp::TileDesc));
  if (ltri) std::memcpy(ltri, t.ltri.data(), t.ltri.size() * sizeof(uint32_t));
  if (halo) std::memcpy(halo, t.halo.data(), t.halo.size() * sizeof(int));
  if (ring) std::memcpy(ring, t.ring.data(), t.ring.size() * sizeof(uint16_t));
  if (fan) std::memcpy(fan, t.fan.data(), t.fan.size() * sizeof(uint32_t));
  return VP_OK;
}

/* Triangle ownership of the fused vertex + raster kernel: own_tri_off[ntiles + 1], own_ltri[ntri],
 * tri_by_orig[ntri][4]; returns 1 when the fused kernel can take the mesh, 0 when it cannot, < 0 on error. */
extern "C" int vp_topology_copy_owned(const vp_topology* h, int* own_tri_off, uint32_t* own_ltri, int* tri_by_orig) {
  if (h == nullptr) return VP_ERR_ARG;
  const vp::Topology& t = h->t;
  if (own_tri_off) std::memcpy(own_tri_off, t.own_tri_off.data(), t.own_tri_off.size() * sizeof(int));
  if (own_ltri) std::memcpy(own_ltri, t.own_ltri.data(), t.own_ltri.size() * sizeof(uint32_t));
  if (tri_by_orig) std::memcpy(tri_by_orig, t.tri_by_orig.data(), t.tri_by_orig.size() * sizeof(int));
  return t.fused_ok ? 1 : 0;
}

/* The optional slot tables (always built by vp_topology_build): slot_off[ntiles] (-1 = generic tile),
 * slot_tab[vp_topology_slot_count()], fan_slot[nver][5]. */
extern "C" int vp_topology_slot_count(const vp_topology* h) { return h ? (int)h->t.slot_tab.size() : -1; }

extern "C" int vp_topology_copy_slots(const vp_topology* h, int* slot_off, uint16_t* slot_tab, uint32_t* fan_slot) {
  VP_REQUIRE(h != nullptr, "null handle");
  const vp::Topology& t = h->t;
  if (slot_off) std::memcpy(slot_off, t.slot_off.data(), t.slot_off.size() * sizeof(int));
  if (slot_tab) std::memcpy(slot_tab, t.slot_tab.data(), t.slot_tab.size() * sizeof(uint16_t));
  if (fan_slot) std::memcpy(fan_slot, t.fan_slot.data(), t.fan_slot.size() * sizeof(uint32_t));
  return VP_OK;
}
