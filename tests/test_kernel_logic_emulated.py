"""Kernel LOGIC checks without a GPU: the device code of csrc/bilinear_bwd.cu, csrc/afm_bwd.cu and of opn_vec_pairs_kernel (csrc/pnn_senet.cu) is compiled as plain C++ against
tests/emu/cuda_emu.h (one OS thread per CUDA thread, std::barrier for __syncthreads, CTA-uniform shuffles) and compared
with a float64 restatement of the layer's gradient formulas (bilinear_interaction.py:72-76 / :144-149 differentiated).
This is test infrastructure: it proves index arithmetic, accumulator ownership, the prefetch ring and the reductions,
not the CUDA build -- the GPU parity tests (tests/test_gpu_training.py::test_bilinear_backward_kernel) do that."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, 'tests', 'emu')


def _emulated(tmp_path_factory, source, end_marker, smem_decl, smem_name, driver, start_marker=None):
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    src = open(os.path.join(ROOT, 'torecsys_b200', 'csrc', source)).read()
    if start_marker is None:
        body = src.split('#include "common.cuh"', 1)[1].split(end_marker, 1)[0]
    else:   # one kernel out of a larger file: re-open the namespaces it lives in
        body = 'namespace trs {\nnamespace {\n' + start_marker + src.split(start_marker, 1)[1].split(end_marker, 1)[0]
    assert smem_decl in body
    body = body.replace(smem_decl, f'float* {smem_name} = emu::dyn_smem;')
    out = tmp_path_factory.mktemp('emu')
    cpp = out / (source.replace('.cu', '_emu.cpp'))
    cpp.write_text('#include "cuda_emu.h"\n' + body + '}  // namespace\n}  // namespace trs\nusing namespace trs;\n'
                   + open(os.path.join(EMU, driver)).read())
    exe = out / source.replace('.cu', '_emu')
    res = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-I', EMU, '-Wno-unknown-pragmas', str(cpp), '-o',
                          str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return str(exe)


@pytest.fixture(scope='module')
def emulated_binary(tmp_path_factory):
    return _emulated(tmp_path_factory, 'bilinear_bwd.cu', 'template <int E>\nint bilinear_backward_run',
                     'extern __shared__ __align__(16) float bx_smem[];', 'bx_smem', 'bilinear_bwd_main.inc')


@pytest.fixture(scope='module')
def emulated_afm(tmp_path_factory):
    return _emulated(tmp_path_factory, 'afm_bwd.cu', 'template <int E, int A>\nint afm_backward_run',
                     'extern __shared__ __align__(16) float ab_smem[];', 'ab_smem', 'afm_bwd_main.inc')


# embed, batch, fields, each_type, CTAs of the sample-major kernel, samples per slice of the pair-major kernel
@pytest.mark.parametrize('args', [(16, 37, 6, 1, 3, 64), (8, 65, 5, 1, 2, 128), (32, 130, 3, 0, 5, 64),
                                  (16, 2, 39, 1, 1, 64), (8, 1, 2, 0, 1, 64)])
def test_bilinear_backward_kernel_logic(emulated_binary, args):
    res = subprocess.run([emulated_binary] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr


# embed, attn, batch, fields, with grad_scores, CTAs
@pytest.mark.parametrize('args', [(16, 16, 19, 4, 1, 2), (8, 32, 5, 4, 1, 1), (32, 8, 3, 3, 0, 1)])
def test_afm_backward_kernel_logic(emulated_afm, args):
    res = subprocess.run([emulated_afm] + [str(a) for a in args], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr


@pytest.fixture(scope='module')
def emulated_opn_vec(tmp_path_factory):
    return _emulated(tmp_path_factory, 'pnn_senet.cu', 'template <int E, int PT, bool kVec>\nint launch_opn_vec_pairs',
                     'extern __shared__ __align__(16) float vsm[];', 'vsm', 'opn_vec_main.inc',
                     start_marker='constexpr int kVecThreads = 256;')


# embed, vec (1) / num (0), batch, fields, samples per tile, CTAs
@pytest.mark.parametrize('args', [(16, 1, 11, 6, 4, 2), (8, 1, 5, 24, 2, 1), (32, 0, 7, 3, 3, 3)])
def test_opn_vec_pairs_kernel_logic(emulated_opn_vec, args):
    """Bit-identical to the reference's operation order ((x_i * x_j) first, the kernel second, summed over e in order):
    ragged last tiles, two pairs per thread (276 pairs), both kernel types."""
    res = subprocess.run([emulated_opn_vec] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
