"""TEST INFRASTRUCTURE -- compiles oracle/mesh_core_oracle.c into oracle/_build/libvp_oracle.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'mesh_core_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libvp_oracle.so')
# no FMA contraction, no fast-math: every float op individually rounded
CFLAGS = ['-O2', '-std=c99', '-ffp-contract=off', '-fno-fast-math', '-fPIC', '-shared', '-Wall']


def build(force=False):
  if (not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
    return LIB
  os.makedirs(OUT_DIR, exist_ok=True)
  subprocess.run(['gcc'] + CFLAGS + [SRC, '-o', LIB, '-lm'], check=True)
  return LIB


if __name__ == '__main__':
  print(build(force=True))
