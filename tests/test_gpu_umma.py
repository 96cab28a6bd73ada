"""tcgen05 building blocks: the single-CTA UMMA self-test behind
vadc_debug_umma pins the shared-memory / instruction descriptor encodings,
the SWIZZLE_128B operand layouts (K-major and MN-major), the TMEM accumulator
layout read by tcgen05.ld and the tf32 input semantics (the tensor core
truncates fp32 operands to tf32), for kind::tf32 and kind::f16(bf16)."""
import ctypes

import numpy as np
import pytest
import torch

from videoad_b200 import _lib
from gpu_util import T, N as to_np, dev

pytestmark = pytest.mark.gpu


def tf32_trunc(a):
    return (a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_rne(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def bf16_rne(a):
    return torch.tensor(a).to(torch.bfloat16).float().numpy()


def run(A, B, Nn, Kd, mode):
    out = torch.empty((128, Nn), device=dev(), dtype=torch.float32)
    At, Bt = T(A), T(B)          # keep the device tensors alive across the launch
    rc = _lib.lib().vadc_debug_umma(_lib.ptr(At), _lib.ptr(Bt), _lib.ptr(out), Nn, Kd, mode, _lib.stream())
    _lib.check(rc, "vadc_debug_umma")
    torch.cuda.synchronize()
    return to_np(out)


@pytest.mark.parametrize("mode", [0, 2, 3, 6, 7, 8, 10, 11])
@pytest.mark.parametrize("Nn,Kd", [(32, 64), (64, 192), (192, 64), (256, 128)])
def test_umma_modes(mode, Nn, Kd):
    if (mode & 3) == 3 and Nn % 64:
        pytest.skip("bf16 MN-major B needs N % 64 == 0 (one 128-byte swizzle row = 64 bf16)")
    rng = np.random.default_rng(mode * 100 + Nn + Kd)
    A = rng.standard_normal((128, Kd)).astype(np.float32)      # logical [M, K]
    B = rng.standard_normal((Nn, Kd)).astype(np.float32)       # logical [N, K]
    Ain = np.ascontiguousarray(A.T) if mode & 4 else A
    Bin = np.ascontiguousarray(B.T) if mode & 1 else B
    got = run(Ain, Bin, Nn, Kd, mode)
    if mode & 2:
        ref = bf16_rne(A).astype(np.float64) @ bf16_rne(B).astype(np.float64).T
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err < 2e-6, err
    else:
        ref_t = tf32_trunc(A).astype(np.float64) @ tf32_trunc(B).astype(np.float64).T
        ref_r = tf32_rne(A).astype(np.float64) @ tf32_rne(B).astype(np.float64).T
        et = np.abs(got - ref_t).max() / np.abs(ref_t).max()
        er = np.abs(got - ref_r).max() / np.abs(ref_r).max()
        print(f"mode {mode} N {Nn} K {Kd}: err vs trunc {et:.2e}, vs rne {er:.2e}")
        assert min(et, er) < 2e-6, (et, er)
        assert et < er, "kind::tf32 truncates (does not round) fp32 operands to tf32"


def f16_rne(a):
    return a.astype(np.float16).astype(np.float32)


# NOTE: one kind::f16 instruction cannot take an fp16 A with a bf16 B (or vice versa): setting different
# a_format / b_format fields faults the kernel on B200 (tried with vadc_debug_umma mode bits 4 / 5, which remain
# in the self-test kernel for reference).  This is why the fused backward is bf16 on both sides.


@pytest.mark.parametrize("mode", [3, 7])
@pytest.mark.parametrize("Kd", [64, 128])
def test_umma_bf16_mn_major_b_half_swizzle_row(mode, Kd):
    """N = 32 bf16 columns of an MN-major B fill only half of the 128-byte swizzle row"""
    rng = np.random.default_rng(mode + Kd)
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((32, Kd)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if mode & 4 else A
    got = run(Ain, np.ascontiguousarray(B.T), 32, Kd, mode)
    ref = bf16_rne(A).astype(np.float64) @ bf16_rne(B).astype(np.float64).T
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


@pytest.mark.parametrize("mode", [3 | 64, 7 | 64])
def test_umma_bf16_mn_major_b_second_half_of_the_row(mode):
    """B = columns [32, 64) of a [Kd, 64] MN-major tile, addressed by a +64-byte start offset"""
    Kd = 64
    rng = np.random.default_rng(mode)
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((64, Kd)).astype(np.float32)
    Ain = np.ascontiguousarray(A.T) if mode & 4 else A
    got = run(Ain, np.ascontiguousarray(B.T), 32, Kd, mode)
    ref = bf16_rne(A).astype(np.float64) @ bf16_rne(B[32:]).astype(np.float64).T
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err


def test_tf32_mn_major_needs_the_32B_base_swizzle():
    """tf32 MN-major operands only exist in the SWIZZLE_128B_BASE32B layout, which differs from the
    K-major SWIZZLE_128B tile: this is why the fused kernels use a 3-term bf16 split (kind::f16), whose
    K-major and MN-major SWIZZLE_128B tiles are byte-identical, so GEMM1 (centroids K-major) and GEMM2
    (centroids MN-major) share one shared-memory copy of the centroid operand.  Documented, not asserted
    on hardware behaviour beyond 'the plain SWIZZLE_128B tf32 MN-major tile does not give A*B'."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = rng.standard_normal((64, 64)).astype(np.float32)
    got = run(A, np.ascontiguousarray(B.T), 64, 64, 1)
    ref = tf32_trunc(A).astype(np.float64) @ tf32_trunc(B).astype(np.float64).T
    assert np.abs(got - ref).max() / np.abs(ref).max() > 1e-3
