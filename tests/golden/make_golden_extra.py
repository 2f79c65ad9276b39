"""Generates tests/golden/texture_normals.npz by running the UNMODIFIED reference Cython module
(utils/cython/mesh_core_cython.pyx: render_texture_core :80-99, get_normal_core :40-47) in this container.

Run from the repo root:  python tests/golden/make_golden_extra.py
Inputs and outputs are both stored (they are small); /root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))

from oracle import build_ref  # noqa: E402

build_ref.build()
import mesh_core_cython as ref_raster  # noqa: E402  (reference Cython module)

OUT = os.path.dirname(os.path.abspath(__file__))


def texture_case(seed, h, w, c, tex_h, tex_w, tex_c, lattice):
  """A grid mesh with a bump (self-occlusion), a few stray triangles, UVs partly outside the texture."""
  rng = np.random.Generator(np.random.PCG64(seed))
  gx, gy = 9, 8
  xs, ys = np.meshgrid(np.linspace(-3, w + 2, gx), np.linspace(-2, h + 3, gy))
  z = 3.0 * np.exp(-((xs - w / 2) ** 2 + (ys - h / 2) ** 2) / (0.1 * w * h)) + rng.random((gy, gx))
  verts = np.stack([xs + rng.random((gy, gx)) * 2, ys + rng.random((gy, gx)) * 2, z], axis=2).reshape(-1, 3)
  if lattice:
    verts[:, :2] = np.round(verts[:, :2] * 2) / 2       # pixel centres on edges, exact ties
    verts[:, 2] = np.round(verts[:, 2])
  tris = []
  for j in range(gy - 1):
    for i in range(gx - 1):
      a = j * gx + i
      tris += [[a, a + 1, a + gx], [a + 1, a + gx + 1, a + gx]]
  nver = verts.shape[0]
  tris = np.array(tris, dtype=np.int32)
  extra = rng.integers(0, nver, (12, 3)).astype(np.int32)   # overlapping strays, some degenerate
  tris = np.concatenate([tris, extra])
  ntri = tris.shape[0]
  tex_nver = nver + 5
  tex_coords = np.zeros((tex_nver, 3), dtype=np.float32)
  tex_coords[:, 0] = rng.random(tex_nver) * (tex_w + 6) - 3      # some outside: exercises the clamps
  tex_coords[:, 1] = rng.random(tex_nver) * (tex_h + 6) - 3
  tex_coords[::7, :2] = np.round(tex_coords[::7, :2])            # integral: xd = yd = 0, ceil == floor
  tex_tris = rng.integers(0, tex_nver, (ntri, 3)).astype(np.int32)
  texture = (rng.random((tex_h, tex_w, tex_c)) * 255).astype(np.float32)
  return dict(vertices=verts.astype(np.float32), triangles=tris, tex_coords=tex_coords, tex_triangles=tex_tris,
              texture=texture, h=h, w=w, c=c, tex_h=tex_h, tex_w=tex_w, tex_c=tex_c)


def run_texture(case, mapping, init_depth=None):
  h, w, c = case['h'], case['w'], case['c']
  image = np.full((h, w, c), -1.0, dtype=np.float32)     # untouched pixels keep the caller's value
  depth = np.full((h, w), -99999.0, dtype=np.float32) if init_depth is None else init_depth.copy()
  ref_raster.render_texture_core(image, case['vertices'], case['triangles'], case['texture'], case['tex_coords'],
                                 case['tex_triangles'], depth, case['vertices'].shape[0], case['tex_coords'].shape[0],
                                 case['triangles'].shape[0], h, w, c, case['tex_h'], case['tex_w'], case['tex_c'],
                                 mapping)
  return image, depth


def normal_case(seed, nver, ntri):
  rng = np.random.Generator(np.random.PCG64(seed))
  tris = rng.integers(0, nver, (ntri, 3)).astype(np.int32)
  tris[::9, 1] = tris[::9, 0]                                   # a vertex twice in one triangle: added twice
  tris[5] = [3, 3, 3]
  # wide dynamic range so that the float32 summation ORDER is visible in the result
  tri_normal = (rng.standard_normal((ntri, 3)) * np.exp(rng.standard_normal((ntri, 1)) * 4)).astype(np.float32)
  init = rng.standard_normal((nver, 3)).astype(np.float32)
  init[::4] = 0
  return tris, tri_normal, init


def main():
  data = {}
  specs = [('tex_a', 21, 40, 48, 3, 24, 32, 3, False), ('tex_b', 22, 33, 29, 1, 16, 20, 2, True),
           ('tex_c', 23, 64, 64, 4, 9, 7, 4, True)]
  for name, seed, h, w, c, th, tw, tc, lattice in specs:
    case = texture_case(seed, h, w, c, th, tw, tc, lattice)
    for k, v in case.items():
      data['%s_%s' % (name, k)] = v
    for mapping in (0, 1):
      image, depth = run_texture(case, mapping)
      data['%s_m%d_image' % (name, mapping)] = image
      data['%s_m%d_depth' % (name, mapping)] = depth
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    init = (rng.random((h, w)) * 4).astype(np.float32)          # pre-filled depth: only nearer surfaces win
    init[::5, ::3] = np.nan
    init[1::5, ::4] = np.inf
    image, depth = run_texture(case, 1, init)
    data[name + '_init_depth'] = init
    data[name + '_pre_image'] = image
    data[name + '_pre_depth'] = depth
  for name, seed, nver, ntri in [('nrm_a', 31, 50, 400), ('nrm_b', 32, 700, 1500)]:
    tris, tri_normal, init = normal_case(seed, nver, ntri)
    normal = init.copy()
    ref_raster.get_normal_core(normal, tri_normal, tris, ntri)
    data[name + '_triangles'] = tris
    data[name + '_tri_normal'] = tri_normal
    data[name + '_init'] = init
    data[name + '_normal'] = normal
  path = os.path.join(OUT, 'texture_normals.npz')
  np.savez_compressed(path, **data)
  print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
  main()
