"""Compile the CUDA library in-tree for sm_100a (B200).  No JIT cache, no torch extension:
`tgm_b200/csrc/libtgm_b200.so` is a plain C-ABI shared object (include/tgm_b200.h).

Every .cu is compiled to an object file (in parallel), then linked.  `gemm_fastf32.cu` instantiates
a CUTLASS/CuTe sm_100 collective; the header tree is the one vendored with the flashinfer package
of this image.  Without it the file compiles to a stub and those GEMMs stay on cuBLAS."""
from __future__ import annotations

import importlib.util
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc')
LIB_PATH = os.path.join(CSRC, 'libtgm_b200.so')
OBJ_DIR = os.path.join(CSRC, '_obj')
SOURCES = ['store.cu', 'recency_ring.cu', 'csr.cu', 'frontier.cu', 'dedup.cu', 'aggregate.cu', 'attention.cu', 'attn_fold.cu', 'tgat.cu', 'attention_bwd.cu',
           'tgn_memory.cu', 'graph_attn.cu', 'dygformer.cu', 'gemm_fastf32.cu', 'uniform_exact.cu', 'negatives.cu', 'memory_join.cu', 'tc_linear.cu', 'small_gemm.cu']
LINK_FLAGS = ['-lcublas', '-Xlinker', '-rpath=/usr/local/cuda/lib64']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']


def _nvcc() -> str:
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: cannot build tgm_b200/csrc/libtgm_b200.so')
    return exe


def cutlass_include_dirs() -> list:
    """CUTLASS/CuTe header tree (include/ and tools/util/include), or [] when absent."""
    roots = [os.environ.get('TGM_CUTLASS_DIR', '')]
    for pkg in ('flashinfer', 'tilelang'):
        try:
            spec = importlib.util.find_spec(pkg)
        except (ImportError, ValueError):
            spec = None
        if spec and spec.submodule_search_locations:
            base = list(spec.submodule_search_locations)[0]
            roots += [os.path.join(base, 'data', 'cutlass'), os.path.join(base, '3rdparty', 'cutlass')]
    for root in roots:
        inc = os.path.join(root, 'include')
        if root and os.path.exists(os.path.join(inc, 'cutlass', 'gemm', 'collective', 'builders',
                                                'sm100_9xBF16_umma_builder.inl')):
            util = os.path.join(root, 'tools', 'util', 'include')
            return [inc] + ([util] if os.path.isdir(util) else [])
    return []


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh'))]
    deps.append(os.path.join(os.path.dirname(CSRC), '..', 'include', 'tgm_b200.h'))
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src[:-3] + '.o')
    cmd = [_nvcc(), *NVCC_FLAGS, '-c', src, '-o', obj]
    if src == 'gemm_fastf32.cu':
        inc = cutlass_include_dirs()
        if inc:
            cmd += ['-DTGM_HAVE_CUTLASS', '--expt-relaxed-constexpr'] + [f'-I{d}' for d in inc]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{proc.stdout}\n{proc.stderr}')
    if verbose:
        print(proc.stderr)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc every source into an object (in parallel), link one shared object; returns its path."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(lambda s: _compile(s, verbose), SOURCES))
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB_PATH, *objs,
           *LINK_FLAGS]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'link failed:\n{proc.stdout}\n{proc.stderr}')
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force=True, verbose=True))
