"""Host logic: the descriptor checks the host entry points run before anything reaches the device (no GPU needed).
Every workload the GPU parity tests use must pass them; malformed units must be refused with SYMGPU_ERR_DECODE."""
import ctypes

import numpy as np
import pytest

import symphonia_b200 as sb
from symphonia_b200 import workloads
from symphonia_b200._native import AAC_TNS_DTYPE

OK, DECODE, ARG = 0, 1, 6


def _mp3(units, runs, n_frames):
    units = np.ascontiguousarray(units)
    runs = np.ascontiguousarray(runs)
    return sb.lib().symgpu_mp3_units_check(units.ctypes.data, runs.ctypes.data, len(runs), n_frames)


def _aac(units, tns, n_frames):
    units = np.ascontiguousarray(units)
    tns = np.ascontiguousarray(tns)
    return sb.lib().symgpu_aac_units_check(units.ctypes.data, tns.ctypes.data if len(tns) else None, len(tns), n_frames)


# the parameter sets of tests/test_mp3_parity_gpu.py, tests/test_output_stage_gpu.py, bench.py and __graft_entry__.smoke()
MP3_CASES = [dict(n_streams=8, frames_per_stream=24, seed=11), dict(n_streams=4, frames_per_stream=10, seed=12, joint=False, block_switching=False),
             dict(n_streams=96, frames_per_stream=1, seed=13), dict(n_streams=3, frames_per_stream=29, seed=14),
             dict(n_streams=3, frames_per_stream=9, seed=15, channels=1), dict(n_streams=2, frames_per_stream=8, seed=20, sample_rate_idx=4),
             dict(n_streams=2, frames_per_stream=8, seed=21, sample_rate_idx=8, channels=1), dict(n_streams=4, frames_per_stream=10, seed=123),
             dict(n_streams=64, frames_per_stream=128, seed=workloads.SEED_BASE + 1)]


@pytest.mark.parametrize("kw", MP3_CASES)
def test_mp3_workloads_pass(kw):
    units, spectra, runs = workloads.mp3_batch(**kw)
    assert _mp3(units, runs, len(spectra)) == OK


def test_mp3_malformed_units_are_refused():
    units, spectra, runs = workloads.mp3_batch(4, 6, seed=31)
    n = len(spectra)

    def bad(mutate, expect=DECODE):
        u = units.copy()
        mutate(u)
        assert _mp3(u, runs, n) == expect

    bad(lambda u: u["subblock_gain"].__setitem__((3, 1, 0, 2), 8))       # a 3-bit field
    bad(lambda u: u["block_type"].__setitem__((0, 0, 1), 4))
    bad(lambda u: u["rzero"].__setitem__((5, 1, 1), 577))
    bad(lambda u: u["sample_rate_idx"].__setitem__((2, 0, 0), 9))

    def mismatched_pair(u):  # joint stereo with different block types (stereo.rs:503-505)
        u["flags"][7, 0, 0] |= sb._native.F_MID_SIDE
        u["block_type"][7, 0, 0], u["block_type"][7, 0, 1] = 0, 2
    bad(mismatched_pair)
    r = runs.copy()
    r["n_frames"][-1] += 1
    assert _mp3(units, r, n) == ARG
    # units of a muted channel / granule are not looked at
    u1, s1, r1 = workloads.mp3_batch(2, 3, seed=32, channels=1)
    u1["block_type"][:, :, 1] = 9
    assert _mp3(u1, r1, len(s1)) == OK


@pytest.mark.parametrize("kw", [dict(n_streams=5, frames_per_stream=20, seed=101), dict(n_streams=3, frames_per_stream=16, seed=102, tns_prob=0.0),
                                dict(n_streams=4, frames_per_stream=12, seed=103, tns_prob=0.9), dict(n_streams=3, frames_per_stream=10, seed=105, channels=1),
                                dict(n_streams=2, frames_per_stream=9, seed=124), dict(n_streams=64, frames_per_stream=128, seed=workloads.SEED_BASE + 2)])
def test_aac_workloads_pass(kw):
    units, tns, coeffs, runs = workloads.aac_batch(**kw)
    assert _aac(units, tns, len(coeffs)) == OK


def test_aac_malformed_units_are_refused():
    units, tns, coeffs, runs = workloads.aac_batch(3, 8, seed=111, tns_prob=0.8)
    n = len(coeffs)
    u = units.copy(); u["window_sequence"][3, 1] = 4
    assert _aac(u, tns, n) == DECODE
    u = units.copy(); u["window_shape"][0, 0] = 2
    assert _aac(u, tns, n) == DECODE
    k = int(np.argmax(units["n_tns"].reshape(-1) > 0))
    u = units.copy(); u.reshape(-1)["tns_first"][k] = len(tns)
    assert _aac(u, tns, n) == DECODE
    t = tns.copy(); t["order"][0] = 21
    assert _aac(units, t, n) == DECODE
    t = tns.copy(); t["end"][0] = 1025
    assert _aac(units, t, n) == DECODE
    t = tns.copy(); t["start"][0], t["end"][0] = 40, 30
    assert _aac(units, t, n) == DECODE
    assert _aac(units, np.zeros(0, dtype=AAC_TNS_DTYPE), n) == DECODE  # filters referenced but none given
