"""Kernel LOGIC checks without a GPU: the device code of csrc/bilinear_bwd.cu, csrc/afm_bwd.cu, of opn_vec_pairs_kernel (csrc/pnn_senet.cu) and of tall_dense_tc_kernel
(csrc/dense.cu, mma.sync through a fragment-layout emulation) is compiled as plain C++ against
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


@pytest.fixture(scope='module')
def emulated_tall_dense(tmp_path_factory):
    """tall_dense_tc_kernel with mma.sync replaced by the fragment-layout emulation of cuda_emu.h; the hi/lo split is the
    product's own text."""
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    src = open(os.path.join(ROOT, 'torecsys_b200', 'csrc', 'dense.cu')).read()
    split = '__device__ __forceinline__ void tall_split_tf32' + src.split(
        '__device__ __forceinline__ void tall_split_tf32', 1)[1].split('\n}\n', 1)[0] + '\n}\n'
    start = 'constexpr int kTallTcWarps = 4, kTallTcFlush = 8;'
    kernel = start + src.split(start, 1)[1].split('static bool tall_dense_tc_ok', 1)[0]
    mma = ('static inline void tall_mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, '
           'uint32_t b0, uint32_t b1) { emu_mma_m16n8k8_tf32(d, a0, a1, a2, a3, b0, b1); }\n')
    out = tmp_path_factory.mktemp('emu')
    cpp = out / 'tall_dense_emu.cpp'
    cpp.write_text('#include "cuda_emu.h"\nnamespace trs {\nnamespace {\n' + mma + split + kernel
                   + '}  // namespace\n}  // namespace trs\nusing namespace trs;\n'
                   + open(os.path.join(EMU, 'tall_dense_main.inc')).read())
    exe = out / 'tall_dense_emu'
    res = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-I', EMU, '-Wno-unknown-pragmas', str(cpp), '-o',
                          str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return str(exe)


# rows, K, outputs, activation id, CTAs
@pytest.mark.parametrize('args', [(37, 1000, 16, 1, 2), (16, 268, 5, 0, 1), (20, 2052, 32, 1, 3), (3, 64, 24, 0, 1)])
def test_tall_dense_tensor_core_kernel_logic(emulated_tall_dense, args):
    """Fragment mapping of the k-permuted 16-byte loads, the four k-quarters of a CTA, flush into the FP32 master sums,
    ragged rows / K tails (K a multiple of 4 but not of 16 or 64) / output counts that are not a multiple of 8."""
    res = subprocess.run([emulated_tall_dense] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr


@pytest.fixture(scope='module')
def emulated_pool_rows(tmp_path_factory):
    """pool_rows<E> of csrc/cin_tc.cu (the CIN epilogue's butterfly reduce-scatter over the rows of a sample) on its own."""
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    src = open(os.path.join(ROOT, 'torecsys_b200', 'csrc', 'cin_tc.cu')).read()
    start = 'template <int E>\n__device__ __forceinline__ void pool_rows'
    body = start + src.split(start, 1)[1].split('__host__ __device__ inline int ss_pitch', 1)[0]
    out = tmp_path_factory.mktemp('emu')
    cpp = out / 'pool_rows_emu.cpp'
    cpp.write_text('#include "cuda_emu.h"\n#include <cstring>\nnamespace trs {\n' + body + '}  // namespace trs\n'
                   + open(os.path.join(EMU, 'pool_rows_main.inc')).read())
    exe = out / 'pool_rows_emu'
    res = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-I', EMU, '-Wno-unknown-pragmas', str(cpp), '-o',
                          str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return str(exe)


# embed, warps, channels that exist in the group, first row beyond the batch (a multiple of embed)
@pytest.mark.parametrize('args', [(16, 8, 32, 256), (16, 2, 32, 48), (8, 4, 20, 128), (32, 3, 7, 64), (16, 1, 1, 16),
                                  (8, 1, 32, 0)])
def test_cin_pool_rows_butterfly_is_bit_identical_to_the_all_reduce(emulated_pool_rows, args):
    """Every (sample, channel) sum of the reduce-scatter equals the all-reduce it replaced bit for bit, is written by
    exactly one lane, respects the channel limit of a partial group and skips samples beyond the batch."""
    res = subprocess.run([emulated_pool_rows] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
