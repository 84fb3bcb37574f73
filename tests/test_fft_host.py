"""The shared-memory inverse real FFT of the mask_istft kernel (lass_b200/csrc/fft.cuh) is written as
__host__ __device__ index math; this test compiles it for the host and checks it against numpy.irfft."""
import ctypes
import os
import subprocess

import numpy as np
import pytest


@pytest.fixture(scope="module")
def shim(tmp_path_factory, repo_root):
    out = tmp_path_factory.mktemp("shim") / "libfftshim.so"
    src = os.path.join(repo_root, "tests", "host_shims", "fft_shim.cpp")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", str(out)])
    lib = ctypes.CDLL(str(out))
    lib.lass_host_irfft.restype = ctypes.c_int
    return lib


@pytest.mark.parametrize("n_fft", [16, 32, 64, 256, 512, 1024, 2048])
def test_irfft_matches_numpy(shim, n_fft):
    rng = np.random.default_rng(n_fft)
    X = (rng.standard_normal(n_fft // 2 + 1) + 1j * rng.standard_normal(n_fft // 2 + 1)).astype(np.complex64)
    j = np.arange(n_fft)
    tw = np.exp(2j * np.pi * j / n_fft).astype(np.complex64)
    out = np.zeros(n_fft, dtype=np.float32)
    shim.lass_host_irfft(X.view(np.float32).ctypes.data_as(ctypes.c_void_p),
                         tw.view(np.float32).ctypes.data_as(ctypes.c_void_p),
                         ctypes.c_int(n_fft), out.ctypes.data_as(ctypes.c_void_p))
    # numpy.irfft ignores Im X[0], Im X[N/2] exactly like the reference's conv basis (sin(0) = sin(pi n) = 0)
    ref = np.fft.irfft(X.astype(np.complex128), n=n_fft)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 2e-6, err
