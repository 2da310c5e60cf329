"""Builds libtorecsys_b200.so (the C-ABI library, include/torecsys_b200.h) with nvcc for sm_100a, in-tree.

    python -m torecsys_b200.build [--force] [--verbose]

No torch headers are involved: the library is plain CUDA C++ behind an `extern "C"` surface, so a build is one
nvcc call per translation unit (a few seconds each, cross-compiled without a GPU) plus a link.
"""
import argparse
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, 'csrc')
OBJ_DIR = os.path.join(ROOT, 'build', 'obj')
LIB_PATH = os.path.join(PKG_DIR, 'libtorecsys_b200.so')
STAMP = os.path.join(ROOT, 'build', 'sources.sha1')

SOURCES = ['api.cu', 'embedding.cu', 'pairwise.cu', 'dense.cu', 'cin.cu', 'cin_tc.cu', 'fused_models.cu',
           'deepfm_fast.cu', 'deepfm_packed.cu', 'deepfm_tc5.cu', 'dcn_tc.cu', 'dcn_tc5.cu', 'afm_tc.cu', 'afm_tc5.cu', 'ipn_tc.cu', 'bilinear_tc.cu', 'pnn_senet.cu', 'cross_tc5.cu', 'backward.cu', 'mlp_bwd.cu', 'bilinear_bwd.cu', 'afm_bwd.cu', 'fused_more.cu', 'ffm_interleaved.cu', 'ffm_blocks.cu', 'session.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest():
    h = hashlib.sha1()
    names = sorted(os.listdir(CSRC)) + ['../../include/torecsys_b200.h']
    for name in names:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p):
            h.update(name.encode())
            with open(p, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(STAMP):
        return False
    with open(STAMP) as f:
        return f.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + extra + ['-c', os.path.join(CSRC, src), '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{res.stdout}\n{res.stderr}')
        if verbose:
            sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, _sources()))
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    with open(STAMP, 'w') as f:
        f.write(_digest())
    return LIB_PATH


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose))
