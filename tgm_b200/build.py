"""Compile the CUDA library in-tree for sm_100a (B200).  No JIT cache, no torch extension:
`tgm_b200/csrc/libtgm_b200.so` is a plain C-ABI shared object (include/tgm_b200.h)."""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB_PATH = os.path.join(CSRC, 'libtgm_b200.so')
SOURCES = ['store.cu', 'recency_ring.cu', 'csr.cu', 'frontier.cu', 'dedup.cu', 'aggregate.cu', 'attention.cu', 'attention_bwd.cu',
           'tgn_memory.cu', 'graph_attn.cu', 'dygformer.cu']
LINK_FLAGS = ['-lcublas', '-Xlinker', '-rpath=/usr/local/cuda/lib64']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build tgm_b200/csrc/libtgm_b200.so')
    return exe


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(os.path.dirname(CSRC), '..', 'include', 'tgm_b200.h'))
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc all sources into one shared object; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, '-o', LIB_PATH, *SOURCES, *LINK_FLAGS]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed:\n{proc.stdout}\n{proc.stderr}')
    if verbose:
        print(proc.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
