"""TEST INFRASTRUCTURE -- builds the reference's own native rasterizer into oracle/_ref/.

Recipe (no reference build system is run, no reference source is copied):
  1. libmesh_core_ref.so : g++ on /root/reference/utils/cython/mesh_core.cpp (in place)
     + oracle/ref_shim.cpp (extern "C" forwarders).
  2. mesh_core_cython*.so : `cython --cplus` on the reference .pyx (in place, generated
     C++ written under oracle/_ref/gen/) + g++ with mesh_core.cpp -> the very Python
     module the reference scripts import (utils/cython/mesh_core_cython.pyx:40-99).
Flags follow what the reference's distutils build gives on this interpreter
(-O2 -fno-strict-overflow, baseline x86-64: SSE2 scalar float math, no FMA).

oracle/_ref/ is git-ignored but travels to the GPU box with gpurun snapshots;
/root/reference does not exist there, so `build()` is a no-op when the sources are absent
and the prebuilt files are used.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('VP_REFERENCE_ROOT', '/root/reference')
CY_DIR = os.path.join(REF_SRC, 'utils', 'cython')
OUT = os.path.join(HERE, '_ref')
CXXFLAGS = ['-O2', '-fno-strict-overflow', '-fPIC', '-fwrapv', '-DNDEBUG']


def ref_available():
  return os.path.exists(os.path.join(CY_DIR, 'mesh_core.cpp'))


def lib_path():
  return os.path.join(OUT, 'libmesh_core_ref.so')


def cython_module_path():
  suffix = sysconfig.get_config_var('EXT_SUFFIX')
  return os.path.join(OUT, 'mesh_core_cython' + suffix)


def _run(cmd):
  subprocess.run(cmd, check=True)


def _stale(target, sources):
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build(verbose=False):
  """Build both artefacts if the reference sources are present. Returns True if usable."""
  if not ref_available():
    return os.path.exists(lib_path())
  os.makedirs(os.path.join(OUT, 'gen'), exist_ok=True)
  core = os.path.join(CY_DIR, 'mesh_core.cpp')
  shim = os.path.join(HERE, 'ref_shim.cpp')
  if _stale(lib_path(), [core, shim]):
    _run(['g++'] + CXXFLAGS + ['-shared', '-I', CY_DIR, shim, core, '-o', lib_path()])
    if verbose:
      print('built', lib_path())
  mod = cython_module_path()
  pyx = os.path.join(CY_DIR, 'mesh_core_cython.pyx')
  if _stale(mod, [core, pyx]):
    try:
      import numpy
      gen = os.path.join(OUT, 'gen', 'mesh_core_cython.cpp')
      _run([sys.executable, '-m', 'cython', '--cplus', '-3', pyx, '-o', gen])
      _run(['g++'] + CXXFLAGS + ['-shared', '-w', '-I', CY_DIR, '-I', sysconfig.get_paths()['include'],
                                 '-I', numpy.get_include(), gen, core, '-o', mod])
      if verbose:
        print('built', mod)
    except (subprocess.CalledProcessError, ImportError) as e:  # the ctypes lib is enough
      print('warning: could not build the Cython module: %s' % e, file=sys.stderr)
  return os.path.exists(lib_path())


if __name__ == '__main__':
  ok = build(verbose=True)
  print('reference rasterizer available:', ok)
