"""nvcc build of the in-tree C-ABI library (importable without the library being present)."""
from __future__ import annotations

import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, 'libmvx_b200.so')
CSRC = os.path.join(_PKG, 'csrc')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def nvcc_command(out: str = LIB_PATH):
    return ['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
            '-Xcompiler', '-fPIC', '-shared', '-o', out] + sources()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libmvx_b200.so (cross-compiles without a GPU)."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_ROOT, 'include', 'mvx_b200.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = nvcc_command()
    if verbose:
        print(' '.join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == '__main__':
    build(force='-f' in sys.argv, verbose=True)
