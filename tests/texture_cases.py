"""Helpers shared by the CPU (oracle) and GPU tests of render_texture_core / get_normal_core."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'texture_normals.npz')
TEXTURE_CASES = ('tex_a', 'tex_b', 'tex_c')
NORMAL_CASES = ('nrm_a', 'nrm_b')
CASE_KEYS = ('vertices', 'triangles', 'tex_coords', 'tex_triangles', 'texture', 'h', 'w', 'c', 'tex_h', 'tex_w', 'tex_c')


def load():
  with np.load(GOLDEN) as z:
    return {k: z[k] for k in z.files}


def case_of(g, name):
  c = {k: g['%s_%s' % (name, k)] for k in CASE_KEYS}
  for k in ('h', 'w', 'c', 'tex_h', 'tex_w', 'tex_c'):
    c[k] = int(c[k])
  return c


def run_texture(fn, case, mapping, init_depth=None, **kw):
  """fn has the signature of mesh_core_cython.render_texture_core (utils/cython/mesh_core_cython.pyx:80-99)."""
  h, w, c = case['h'], case['w'], case['c']
  image = np.full((h, w, c), -1.0, dtype=np.float32)
  depth = np.full((h, w), -99999.0, dtype=np.float32) if init_depth is None else init_depth.copy()
  fn(image, case['vertices'], case['triangles'], case['texture'], case['tex_coords'], case['tex_triangles'], depth,
     case['vertices'].shape[0], case['tex_coords'].shape[0], case['triangles'].shape[0], h, w, c,
     case['tex_h'], case['tex_w'], case['tex_c'], mapping, **kw)
  return image, depth


def random_texture_case(seed):
  rng = np.random.Generator(np.random.PCG64(seed))
  h, w = int(rng.integers(8, 60)), int(rng.integers(8, 60))
  tex_h, tex_w = int(rng.integers(2, 30)), int(rng.integers(2, 30))
  tex_c = int(rng.integers(1, 5))
  c = int(rng.integers(1, tex_c + 1))
  nt = int(rng.integers(1, 300))
  nv = 3 * nt
  extent = [2.0, 8.0, 50.0][seed % 3]
  centre = rng.random((nt, 1, 3)) * np.array([w + 8, h + 8, 4]) - np.array([4, 4, 0])
  verts = (centre + (rng.random((nt, 3, 3)) - 0.5) * np.array([extent, extent, 1.0])).reshape(nv, 3).astype(np.float32)
  if seed % 2:
    verts[:, :2] = np.round(verts[:, :2] * 2) / 2
    verts[:, 2] = np.round(verts[:, 2])
  tris = np.arange(nv, dtype=np.int32).reshape(nt, 3)
  share = rng.random(nt) < 0.3
  tris[share] = rng.integers(0, nv, (int(share.sum()), 3))
  tex_nver = nv + int(rng.integers(0, 9))
  tex_coords = np.zeros((tex_nver, 3), dtype=np.float32)
  tex_coords[:, 0] = rng.random(tex_nver) * (tex_w + 4) - 2
  tex_coords[:, 1] = rng.random(tex_nver) * (tex_h + 4) - 2
  if seed % 2:
    tex_coords[:, :2] = np.round(tex_coords[:, :2] * 2) / 2      # x.5 texels: round() half away from zero
  tex_tris = rng.integers(0, tex_nver, (nt, 3)).astype(np.int32)
  texture = (rng.standard_normal((tex_h, tex_w, tex_c)) * 100).astype(np.float32)
  return dict(vertices=verts, triangles=tris, tex_coords=tex_coords, tex_triangles=tex_tris, texture=texture,
              h=h, w=w, c=c, tex_h=tex_h, tex_w=tex_w, tex_c=tex_c)


def random_normal_case(seed):
  rng = np.random.Generator(np.random.PCG64(seed))
  nver = int(rng.integers(3, 900))
  ntri = int(rng.integers(1, 3000))
  tris = rng.integers(0, nver, (ntri, 3)).astype(np.int32)
  tris[::7, 2] = tris[::7, 0]
  tri_normal = (rng.standard_normal((ntri, 3)) * np.exp(rng.standard_normal((ntri, 1)) * 4)).astype(np.float32)
  init = rng.standard_normal((nver, 3)).astype(np.float32)
  return tris, tri_normal, init
