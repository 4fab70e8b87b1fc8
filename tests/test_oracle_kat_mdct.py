"""Pins the shared IMDCT / FFT oracle against the reference's own test vectors, at its tolerance:

  * verify_fft / verify_fft_inplace   symphonia-core/src/dsp/fft/mod.rs:155-185 (64 complex points vs a
                                      naive f64 DFT, 1e-5)
  * verify_imdct                      symphonia-core/src/dsp/mdct.rs:177-201 (N=32 ramp, scale sqrt(2/64),
                                      vs the analytical IMDCT in f64, 1e-5)
and, beyond the reference, every power-of-two size the decoders use against the analytical forms."""
import ctypes
import json
import os

import numpy as np
import pytest

from tests._oracle import ptr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _imdct_analytical(x, scale):
    n_in = len(x)
    n_out = 2 * n_in
    i = np.arange(n_out)[:, None]
    j = np.arange(n_in)[None, :]
    c = np.cos(np.pi / (2 * n_out) * ((2 * i + 1 + n_in) * (2 * j + 1)))
    return scale * (c * x.astype(np.float64)[None, :]).sum(axis=1)


def _imdct(oracle, x, n, scale):
    out = np.zeros(2 * n, dtype=np.float32)
    oracle.oracle_imdct(ptr(x), ptr(out), ctypes.c_int(n), ctypes.c_double(scale))
    return out


def test_fft64_reference_vector(oracle):
    v = json.load(open(os.path.join(GOLD, "fft64_test_vector.json")))
    x = np.array(v["re"], dtype=np.float32) + 1j * np.array(v["im"], dtype=np.float32)
    buf = np.empty(128, dtype=np.float32)
    buf[0::2], buf[1::2] = x.real, x.imag
    oracle.oracle_fft_inplace(ptr(buf), 64)
    expect = np.fft.fft(x.astype(np.complex128))
    assert np.abs(buf[0::2] - expect.real).max() < 1e-5
    assert np.abs(buf[1::2] - expect.imag).max() < 1e-5


def test_imdct32_reference_vector(oracle):
    v = json.load(open(os.path.join(GOLD, "imdct32_test_vector.json")))
    x = np.array(v["x"], dtype=np.float32)
    scale = np.sqrt(2.0 / 64.0)
    out = _imdct(oracle, x, 32, scale)
    assert np.abs(out.astype(np.float64) - _imdct_analytical(x, scale)).max() < 1e-5


@pytest.mark.parametrize("n", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048])
def test_fft_all_sizes(oracle, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    buf = np.empty(2 * n, dtype=np.float32)
    buf[0::2], buf[1::2] = x.real, x.imag
    oracle.oracle_fft_inplace(ptr(buf), n)
    expect = np.fft.fft(x.astype(np.complex128))
    tol = 1e-5 * max(1.0, np.abs(expect).max())
    assert np.abs(buf[0::2] - expect.real).max() < tol
    assert np.abs(buf[1::2] - expect.imag).max() < tol


@pytest.mark.parametrize("n,scale", [(32, 1.0), (128, 1.0 / 256.0), (1024, 1.0 / 2048.0), (128, 1.0), (1024, 1.0),
                                     (4096, 1.0), (64, -0.5)])
def test_imdct_sizes_and_scales(oracle, n, scale):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    out = _imdct(oracle, x, n, scale)
    # a negative scale adds n/2 to alpha (mdct.rs:45), i.e. negates the transform
    expect = _imdct_analytical(x, abs(scale)) * (1.0 if scale > 0 else -1.0)
    assert np.abs(out - expect).max() < 1e-5 * max(1.0, np.abs(expect).max())
