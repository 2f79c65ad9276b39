import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE_ROOT = '/root/reference'


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def have_reference_sources():
  return os.path.exists(os.path.join(REFERENCE_ROOT, 'utils', 'reconstruct_mesh.py'))


@pytest.fixture(scope='session')
def golden_small():
  with np.load(os.path.join(GOLDEN, 'recon_small.npz')) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def golden_full():
  with np.load(os.path.join(GOLDEN, 'raster_full.npz')) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def golden_edges():
  with np.load(os.path.join(GOLDEN, 'raster_edge_cases.npz')) as z:
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def small_model(golden_small):
  from voicepuppet_b200 import synthetic
  m = synthetic.make_model(int(golden_small['n_vertices']), int(golden_small['n_boundary']))
  assert m.checksum() == str(golden_small['model_sha256']), 'synthetic model is not bit-reproducible on this host'
  return m


@pytest.fixture(scope='session')
def full_model(golden_full):
  from voicepuppet_b200 import synthetic
  m = synthetic.cached_model()
  assert m.checksum() == str(golden_full['model_sha256']), 'synthetic model is not bit-reproducible on this host'
  return m


@pytest.fixture(scope='session')
def live_reference():
  """The unmodified reference reconstruct_mesh module (only where /root/reference exists)."""
  if not have_reference_sources():
    pytest.skip('/root/reference not present on this host')
  sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'utils'))
  import reconstruct_mesh
  return reconstruct_mesh
