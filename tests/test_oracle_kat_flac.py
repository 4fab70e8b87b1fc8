"""Pins oracle/oracle_flac.cpp (fixed_predict, lpc_predict, samples_shl, decorrelate_*, output scaling of
symphonia-bundle-flac/src/decoder.rs) with the property FLAC exists for: an encoder written from the format
definition (exact integer residuals, workloads.flac_batch) followed by the restoration gives back the PCM, bit for bit."""
import ctypes

import numpy as np
import pytest

from symphonia_b200 import workloads
from symphonia_b200._native import FLAC_FIXED, FLAC_FRAME_DTYPE, FLAC_INDEPENDENT, FLAC_SUBFRAME_DTYPE
from tests import _oracle


def _restore(oracle, frames, subs, samples):
    out = samples.copy()
    oracle.oracle_flac_restore.restype = ctypes.c_int
    oracle.oracle_flac_restore.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p,
                                           ctypes.c_size_t]
    rc = oracle.oracle_flac_restore(_oracle.ptr(frames), len(frames), _oracle.ptr(subs), len(subs), _oracle.ptr(out), out.size)
    return rc, out


@pytest.mark.parametrize("bps,channels", [(16, 2), (24, 2), (16, 1), (8, 2), (32, 2)])
def test_encode_then_restore_is_lossless(oracle, bps, channels):
    frames, subs, samples, expect = workloads.flac_batch(24, 512, seed=40 + bps + channels, bps=bps, channels=channels,
                                                         return_pcm=True)
    rc, got = _restore(oracle, frames, subs, samples)
    assert rc == 0
    for k, sf in enumerate(subs):
        a, n = int(sf["offset"]), int(sf["n"])
        assert (got[a:a + n] == expect[a:a + n]).all(), f"sub-frame {k} type {sf['type']} order {sf['order']}"
    assert len(set(subs["type"])) >= 3 and len(set(frames["assignment"])) >= (3 if channels == 2 else 1)


def test_fixed_predictors_are_finite_differences(oracle):
    # order k restores a k-th order polynomial from zero residuals (decoder.rs:663-707)
    n = 64
    for order in range(5):
        x = np.polyval(np.arange(1, order + 1)[::-1] if order else [7], np.arange(n)).astype(np.int64) if order else np.full(n, 7)
        frames = np.zeros(1, dtype=FLAC_FRAME_DTYPE)
        frames[0] = (0, 1, FLAC_INDEPENDENT, 32, 0, (0, 0))
        subs = np.zeros(1, dtype=FLAC_SUBFRAME_DTYPE)
        subs[0]["n"], subs[0]["type"], subs[0]["order"] = n, FLAC_FIXED, order
        res = np.zeros(n, dtype=np.int32)
        res[:max(order, 1)] = x[:max(order, 1)]
        if order == 0:
            res[:] = x  # order 0 predicts nothing: the residuals are the samples
        rc, got = _restore(oracle, frames, subs, res)
        assert rc == 0 and (got == x).all(), order


def test_malformed_descriptors_are_rejected(oracle):
    frames, subs, samples = workloads.flac_batch(2, 64, seed=5)
    bad = subs.copy()
    bad[0]["type"], bad[0]["order"] = 3, 33
    assert _restore(oracle, frames, bad, samples)[0] != 0
    bad = subs.copy()
    bad[1]["offset"] = samples.size
    assert _restore(oracle, frames, bad, samples)[0] != 0
