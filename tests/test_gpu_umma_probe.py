"""GPU: pins the tcgen05 descriptor / 128B-swizzle / TMEM-lane conventions of csrc/tc_common.cuh.
D[128,N] = A[128,K] @ B[N,K]^T in bf16 with fp32 accumulation, every operand major-ness the DCN
kernels use, B optionally loaded by a bulk async copy of a pre-swizzled image."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from slenderobjdet_b200 import _lib

pytestmark = pytest.mark.gpu
_HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe():
    """The probe is test infrastructure: built on demand next to its source, not part of the product library."""
    src, so = os.path.join(_HERE, "probe", "debug_umma.cu"), os.path.join(_HERE, "probe", "libsdb_probe.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                        "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-cudart", "static", "-shared", "-o", so, src],
                       check=True)
    lib = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    lib.sdb_debug_umma_gemm.argtypes = [vp, vp, vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, vp]
    return lib


def _sw128(row, chunk):
    return row * 128 + ((chunk ^ (row & 7)) << 4)


def _b_image(B, b_mn):
    """host-side pre-swizzled image of B [N,K] (bf16 as int16 view) exactly as the kernel lays it out"""
    N, K = B.shape
    raw = B.view(torch.int16).numpy()
    img = np.zeros(N * K, np.int16)
    n, k = np.meshgrid(np.arange(N), np.arange(K), indexing="ij")
    if not b_mn:
        off = (k >> 6) * (N * 128) + _sw128(n, (k & 63) >> 3) + (k & 7) * 2
    else:
        off = (n >> 6) * (K * 128) + _sw128(k, (n & 63) >> 3) + (n & 7) * 2
    img[off // 2] = raw
    return torch.from_numpy(img)


@pytest.mark.parametrize("a_mn", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("N,K", [(256, 128), (64, 64), (128, 256)])
@pytest.mark.parametrize("bulk", [0, 1])
def test_umma_probe(probe, a_mn, b_mn, N, K, bulk):
    g = torch.Generator().manual_seed(N + K + a_mn * 2 + b_mn)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)
    ref = A.float() @ B.float().T
    Ad, Bd = A.cuda(), B.cuda()
    img = _b_image(B, b_mn).cuda()
    D = torch.zeros(128, N, device="cuda")
    assert 0 == (probe.sdb_debug_umma_gemm(_lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(img), _lib.ptr(D), N, K, a_mn, b_mn,
                                       bulk, _lib.stream_ptr()))
    torch.cuda.synchronize()
    err = (D.cpu() - ref).abs().max().item()
    assert err < 1e-3 * K ** 0.5, (a_mn, b_mn, N, K, bulk, err)
