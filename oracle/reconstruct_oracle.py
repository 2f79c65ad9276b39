"""TEST INFRASTRUCTURE -- numpy oracle for the reconstruction half of the hot path.

Restates /root/reference/utils/reconstruct_mesh.py (numpy is the reference's own
arithmetic here, so rounding-relevant steps use the same numpy primitive on the same
operand layout and dtype; structure, naming and batching are ours).  dtype promotion is
part of the contract: float32 coefficients, float32 or float64 model arrays, and a
float64 zero row appended to the face normals (reconstruct_mesh.py:47) make everything
from the vertex normals on float64.

Pinned by tests/test_oracle_reconstruct.py against the live reference (imported from
/root/reference when present) and against tests/golden/*.npz everywhere else.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import numpy as np

FOCAL = 1015.0
CENTER = 112.0
CAMERA_Z = 10.0
IMG = 224


def split_coeff(coeff):
  """reconstruct_mesh.py:5-13 -- 80 id | 64 ex | 80 tex | 3 angles | 27 gamma | 3 translation."""
  bounds = (0, 80, 144, 224, 227, 254, 257)
  return tuple(coeff[:, a:b] for a, b in zip(bounds[:-1], bounds[1:]))


def shape_formation(id_coeff, ex_coeff, model):
  """reconstruct_mesh.py:20-29."""
  s = np.einsum('ij,aj->ai', model.idBase, id_coeff) + np.einsum('ij,aj->ai', model.exBase, ex_coeff) \
      + model.meanshape
  s = s.reshape(1, -1, 3)
  centre = np.mean(np.reshape(model.meanshape, [1, -1, 3]), axis=1, keepdims=True)   # of the MEAN shape
  return s - centre


def texture_formation(tex_coeff, model):
  """reconstruct_mesh.py:58-62."""
  return (np.einsum('ij,aj->ai', model.texBase, tex_coeff) + model.meantex).reshape(1, -1, 3)


def compute_norm(face_shape, model):
  """reconstruct_mesh.py:35-52: area-weighted one-ring normals, pad slot -> zero row."""
  tri = (model.tri - 1).astype(np.int32)
  ring = (model.point_buf - 1).astype(np.int32)
  a = face_shape[:, tri[:, 0], :]
  b = face_shape[:, tri[:, 1], :]
  c = face_shape[:, tri[:, 2], :]
  fn = np.cross(a - b, b - c)
  fn = np.concatenate([fn, np.zeros([1, 1, 3])], axis=1)
  vn = np.sum(fn[:, ring, :], axis=2)
  return vn / np.expand_dims(np.linalg.norm(vn, axis=2), 2)


def rotation_matrix(angles):
  """reconstruct_mesh.py:68-91: (Rz Ry Rx)^T; cos/sin are taken of float32 scalars."""
  ax, ay, az = angles[:, 0][0], angles[:, 1][0], angles[:, 2][0]
  cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
  rx = np.array([1.0, 0, 0, 0, cx, -sx, 0, sx, cx]).reshape(1, 3, 3)
  ry = np.array([cy, 0, sy, 0, 1, 0, -sy, 0, cy]).reshape(1, 3, 3)
  rz = np.array([cz, -sz, 0, sz, cz, 0, 0, 0, 1]).reshape(1, 3, 3)
  return np.transpose(np.matmul(np.matmul(rz, ry), rx), axes=[0, 2, 1])


def projection_layer(face_shape, rotation, translation, focal=FOCAL, center=CENTER):
  """reconstruct_mesh.py:100-120."""
  cam = np.matmul(face_shape, rotation) + np.reshape(translation, [1, 1, 3])
  flip_z = np.array([1.0, 0, 0, 0, 1, 0, 0, 0, -1.0]).reshape(1, 3, 3)
  cam = np.matmul(cam, flip_z) + np.array([0.0, 0.0, CAMERA_Z]).reshape(1, 1, 3)
  k = np.array([focal, 0.0, center, 0.0, focal, center, 0.0, 0.0, 1.0]).reshape(1, 3, 3)
  aug = np.matmul(cam, np.transpose(k, [0, 2, 1]))
  proj = aug[:, :, 0:2] / np.reshape(aug[:, :, 2], [1, aug.shape[1], 1])
  return proj, -np.reshape(aug[:, :, 2], [1, -1, 1])


def illumination_layer(face_texture, norm, gamma):
  """reconstruct_mesh.py:129-168: 9-band SH, +0.8 ambient on band 0."""
  n = face_texture.shape[1]
  g = np.reshape(gamma, [-1, 3, 9]) + np.array([0.8, 0, 0, 0, 0, 0, 0, 0, 0]).reshape(1, 1, 9)
  a0 = np.pi
  a1 = 2 * np.pi / np.sqrt(3.0)
  a2 = 2 * np.pi / np.sqrt(8.0)
  c0 = 1 / np.sqrt(4 * np.pi)
  c1 = np.sqrt(3.0) / np.sqrt(4 * np.pi)
  c2 = 3 * np.sqrt(5.0) / np.sqrt(12 * np.pi)
  nx, ny, nz = norm[:, :, 0], norm[:, :, 1], norm[:, :, 2]
  bands = [
      np.tile(np.reshape(a0 * c0, [1, 1, 1]), [1, n, 1]),
      np.reshape(-a1 * c1 * ny, [1, n, 1]),
      np.reshape(a1 * c1 * nz, [1, n, 1]),
      np.reshape(-a1 * c1 * nx, [1, n, 1]),
      np.reshape(a2 * c2 * nx * ny, [1, n, 1]),
      np.reshape(-a2 * c2 * ny * nz, [1, n, 1]),
      np.reshape(a2 * c2 * 0.5 / np.sqrt(3.0) * (3 * np.square(nz) - 1), [1, n, 1]),
      np.reshape(-a2 * c2 * nx * nz, [1, n, 1]),
      np.reshape(a2 * c2 * 0.5 * (np.square(nx) - np.square(ny)), [1, n, 1]),
  ]
  y = np.concatenate(bands, axis=2)
  lit = [np.squeeze(np.matmul(y, np.expand_dims(g[:, ch, :], 2)), 2) for ch in range(3)]
  color = np.stack([lit[ch] * face_texture[:, :, ch] for ch in range(3)], axis=2)
  return color, np.stack(lit, axis=2) * 128


def _finish(face_shape_for_projection, face_texture, normal_r, rotation, translation, gamma, keypoints):
  proj, zbuf = projection_layer(face_shape_for_projection, rotation, translation)
  proj = np.stack([proj[:, :, 0], IMG - proj[:, :, 1]], axis=2)
  lms = proj[:, keypoints, :]
  color, _ = illumination_layer(face_texture, normal_r, gamma)
  return color, proj, zbuf, lms


def reconstruction(coeff, model):
  """reconstruct_mesh.py:172-194 -> 7-tuple."""
  idc, exc, texc, ang, gam, trans = split_coeff(coeff)
  shape = shape_formation(idc, exc, model)
  tex = texture_formation(texc, model)
  rot = rotation_matrix(ang)
  nrm_r = np.matmul(compute_norm(shape, model), rot)
  color, proj, zbuf, lms = _finish(shape, tex, nrm_r, rot, trans, gam, model.keypoints)
  return shape, tex, color, proj, zbuf, lms, trans


def reconstruction_rotation(coeff, model, angles):
  """reconstruct_mesh.py:198-223 -> 6-tuple.  Quirks kept: coefficient angles ignored, normals
  rotated once, the shape rotated before Projection_layer rotates it again (:211 then :111)."""
  idc, exc, texc, _, gam, trans = split_coeff(coeff)
  shape = shape_formation(idc, exc, model)
  tex = texture_formation(texc, model)
  rot = rotation_matrix(angles)
  nrm_r = np.matmul(compute_norm(shape, model), rot)
  shape = np.matmul(shape, rot)
  color, proj, zbuf, lms = _finish(shape, tex, nrm_r, rot, trans, gam, model.keypoints)
  return shape, tex, color, proj, zbuf, lms


# ---------------------------------------------------------------------------------------
# frame loop: voicepuppet/pixrefer/infer_bfmvid.py:76-110 (everything before cv2)
# ---------------------------------------------------------------------------------------

def jitter_angle_sequence(n_frames):
  """The module-global triangle wave of infer_bfmvid.py:76-77,85-89, as [T,1,3] float32
  (the value `angles` holds when frame t calls Reconstruction_rotation)."""
  angles = np.array([[0, 0, 0]], dtype=np.float32)
  shift = 0.005
  out = np.zeros((n_frames, 1, 3), dtype=np.float32)
  for t in range(n_frames):
    angles[0][0] += shift
    angles[0][1] += shift
    angles[0][2] += shift
    if angles[0][1] > 0.03 or angles[0][1] < -0.03:
      shift = -shift
    out[t] = angles
  return out


def raster_inputs(face_projection, z_buffer, face_color, res=IMG):
  """infer_bfmvid.py:93-105 plus the resolution convention of SURVEY.md section 7:
  xy scaled by res/224 in float64 before the float32 cast, z untouched."""
  pos = np.squeeze(np.concatenate([face_projection, z_buffer], axis=2), 0)
  if res != IMG:
    pos = pos * np.array([res / 224.0, res / 224.0, 1.0])
  vertices = pos.reshape(-1).astype(np.float32).copy()
  col = np.clip(np.squeeze(face_color, 0), 0, 255).astype(np.int32)
  colors = col.reshape(-1).astype(np.float32).copy()
  return vertices, colors


def triangles_flat(model):
  """infer_bfmvid.py:104."""
  return (model.tri - 1).reshape(-1).astype(np.int32).copy()
