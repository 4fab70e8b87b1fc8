"""GPU parity: the fused CUDA kernel (through the C ABI) vs the oracle, bit for bit.

Bar: every PCM word identical (uint32 view), including the sign of zero.  No tolerance.
"""
import numpy as np
import pytest

from tests import _oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import symphonia_b200 as sb
    eng = sb.Engine(0)
    yield eng
    eng.close()


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _compare(got, want, what):
    g, w = _bits(got), _bits(want)
    bad = np.nonzero(g != w)
    n_bad = len(bad[0])
    if n_bad:
        f, c, i = bad[0][0], bad[1][0], bad[2][0]
        raise AssertionError(f"{what}: {n_bad} of {g.size} PCM words differ; first at frame {f} ch {c} sample {i}: "
                             f"gpu {got[f, c, i]!r} oracle {want[f, c, i]!r}")


def _run_case(engine, oracle, n_streams, frames, **kw):
    from symphonia_b200 import workloads
    units, spectra, runs = workloads.mp3_batch(n_streams, frames, **kw)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, n_streams)
    assert rc == 0
    engine.mp3_streams_alloc(n_streams)
    got = engine.mp3_synth_host(units, spectra, runs)
    ch = kw.get("channels", 2)
    per = 1152 if kw.get("sample_rate_idx", 0) < 3 else 576
    _compare(got[:, :ch, :per], want[:, :ch, :per], f"S={n_streams} F={frames} {kw}")
    assert np.abs(want).max() > 1e-3  # the case is not trivially silent
    return units, spectra, runs, want


def test_mixed_workload(engine, oracle):
    _run_case(engine, oracle, 8, 24, seed=11)


def test_long_blocks_plain_stereo(engine, oracle):
    _run_case(engine, oracle, 4, 10, seed=12, joint=False, block_switching=False)


def test_many_short_streams(engine, oracle):
    # F=1: every tile both loads and stores stream state
    _run_case(engine, oracle, 96, 1, seed=13)


def test_tile_boundaries(engine, oracle):
    # 7.5, 8 and 8.5 frames per tile of 15 granules: runs of 7, 8, 15, 16 frames
    for frames in (7, 8, 15, 16, 23):
        _run_case(engine, oracle, 3, frames, seed=20 + frames)


def test_mono(engine, oracle):
    _run_case(engine, oracle, 5, 9, seed=14, channels=1)


@pytest.mark.parametrize("sr", [1, 2, 3, 4, 5, 6, 7, 8])
def test_other_sample_rates(engine, oracle, sr):
    _run_case(engine, oracle, 3, 9, seed=30 + sr, sample_rate_idx=sr)


def test_mpeg2_mono(engine, oracle):
    _run_case(engine, oracle, 3, 20, seed=41, sample_rate_idx=4, channels=1)


def test_state_carries_across_batches(engine, oracle):
    from symphonia_b200 import workloads
    S, F = 6, 20
    units, spectra, runs = workloads.mp3_batch(S, F, seed=15)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
    engine.mp3_streams_alloc(S)
    # feed the same streams in three uneven slices (7 + 1 + 12 frames); stream state must carry
    u4 = units.reshape(S, F, 2, 2)
    s4 = spectra.reshape(S, F, 2, 2, 576)
    got = np.zeros((S, F, 2, 1152), dtype=np.float32)
    lo = 0
    for part in (7, 1, 12):
        hi = lo + part
        r = runs.copy()
        r["first_frame"] = np.arange(S) * part
        r["n_frames"] = part
        out = engine.mp3_synth_host(np.ascontiguousarray(u4[:, lo:hi]).reshape(-1, 2, 2),
                                    np.ascontiguousarray(s4[:, lo:hi]).reshape(-1, 2, 2, 576), r)
        got[:, lo:hi] = out.reshape(S, part, 2, 1152)
        lo = hi
    _compare(got.reshape(S * F, 2, 1152), want, "state carry")


def test_stream_reset(engine, oracle):
    from symphonia_b200 import workloads
    units, spectra, runs = workloads.mp3_batch(2, 5, seed=16)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, 2)
    engine.mp3_streams_alloc(2)
    first = engine.mp3_synth_host(units, spectra, runs)
    dirty = engine.mp3_synth_host(units, spectra, runs)   # state now non-zero: output must differ
    assert (_bits(dirty) != _bits(first)).any()
    engine.mp3_stream_reset(0)
    engine.mp3_stream_reset(1)
    again = engine.mp3_synth_host(units, spectra, runs)
    _compare(again, want, "after reset")


def test_device_resident_entry_point(engine, oracle):
    import torch
    from symphonia_b200 import workloads
    S, F = 4, 12
    units, spectra, runs = workloads.mp3_batch(S, F, seed=17)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
    engine.mp3_streams_alloc(S)
    dev = torch.device("cuda", 0)
    u_t = torch.from_numpy(units.view(np.uint8).reshape(-1)).to(dev)
    s_t = torch.from_numpy(spectra).to(dev)
    p_t = torch.zeros((S * F, 2, 1152), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    engine.mp3_synth_dev(u_t, s_t, runs, p_t)
    engine.sync()
    _compare(p_t.cpu().numpy(), want, "device entry point")


def test_full_size_properties(engine, oracle):
    """BASELINE config 2 size (8192 frames): linearity-free size-independent checks.
    (a) batch-split invariance: one launch == the same streams cut into two launches, bit for bit;
    (b) EVERY stream of the batch matches the oracle (multi-threaded oracle pass)."""
    from symphonia_b200 import workloads
    S, F = 64, 128
    units, spectra, runs = workloads.mp3_batch(S, F, seed=workloads.SEED_BASE + 1)
    engine.mp3_streams_alloc(S)
    whole = engine.mp3_synth_host(units, spectra, runs)
    engine.mp3_streams_alloc(S)
    u4 = units.reshape(S, F, 2, 2)
    s4 = spectra.reshape(S, F, 2, 2, 576)
    halves = []
    for lo, hi in ((0, 50), (50, 128)):
        r = runs.copy()
        r["first_frame"] = np.arange(S) * (hi - lo)
        r["n_frames"] = hi - lo
        out = engine.mp3_synth_host(np.ascontiguousarray(u4[:, lo:hi]).reshape(-1, 2, 2),
                                    np.ascontiguousarray(s4[:, lo:hi]).reshape(-1, 2, 2, 576), r)
        halves.append(out.reshape(S, hi - lo, 2, 1152))
    split = np.concatenate(halves, axis=1).reshape(S * F, 2, 1152)
    _compare(split, whole, "batch-split invariance")
    import ctypes
    import os
    states = (_oracle.Mp3State * S)()
    want = np.zeros((S * F, 2, 1152), dtype=np.float32)
    rc = oracle.oracle_mp3_batch_mt(ctypes.byref(states), _oracle.ptr(units), _oracle.ptr(spectra), _oracle.ptr(runs),
                                    ctypes.c_uint32(len(runs)), _oracle.ptr(want), ctypes.c_int(min(os.cpu_count() or 1, S)))
    assert rc == 0
    _compare(whole, want, "all 64 streams of the full batch")


@pytest.mark.parametrize("mode", ["2", "1", "0", "s"])
def test_pinned_host_buffers_and_the_zero_copy_paths(oracle, mode, monkeypatch):
    """SYMGPU_ZERO_COPY=2: pinned (device-mapped) host buffers go to the kernel as they are -- its TMA copies read the spectra
    across PCIe, its PCM stores land in host memory; =1: only the output; =0: staged copies (the default).  Same bits as the
    oracle in every mode; long runs and the serving shape."""
    import torch
    import symphonia_b200 as sb
    from symphonia_b200 import workloads
    monkeypatch.setenv("SYMGPU_ZERO_COPY", mode)
    engine = sb.Engine(0)
    for S, F, seed in ((6, 40, 31), (96, 1, 32)):
        units, spectra, runs = workloads.mp3_batch(S, F, seed=seed)
        rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
        assert rc == 0
        u_pin = torch.from_numpy(units.view(np.uint8).reshape(-1).copy()).pin_memory()
        s_pin = torch.from_numpy(spectra.copy()).pin_memory()
        p_pin = torch.zeros((S * F, 2, 1152), dtype=torch.float32).pin_memory()
        engine.mp3_streams_alloc(S)
        launches = engine.launch_count
        got = engine.mp3_synth_host(u_pin.numpy().view(sb._native.MP3_GC_DTYPE).reshape(S * F, 2, 2), s_pin.numpy(), runs, out=p_pin.numpy())
        assert engine.launch_count == launches + 1, "one launch, no staging kernels"
        _compare(got, want, f"zero-copy mode {mode}, host entry point S={S} F={F}")
    engine.close()


@pytest.mark.parametrize("pinned", [True, False])
def test_one_packet_per_call_with_and_without_pinned_buffers(oracle, pinned, monkeypatch):
    """Config 1's call pattern: one host call per packet, the stream state carried between the calls.  With pinned (device-mapped)
    buffers a batch below the pipeline threshold is one launch on the caller's memory (the default); with pageable buffers it is
    staged through device memory.  Same bits either way."""
    import torch
    import symphonia_b200 as sb
    from symphonia_b200 import workloads
    monkeypatch.delenv("SYMGPU_ZERO_COPY", raising=False)
    engine = sb.Engine(0)
    F = 24
    units, spectra, runs = workloads.mp3_batch(1, F, seed=77, joint=False)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, 1)
    assert rc == 0
    if pinned:
        u_buf = torch.from_numpy(units.view(np.uint8).reshape(-1).copy()).pin_memory().numpy().view(sb._native.MP3_GC_DTYPE).reshape(F, 2, 2)
        s_buf = torch.from_numpy(spectra.copy()).pin_memory().numpy()
        p_keep = torch.zeros((1, 2, 1152), dtype=torch.float32).pin_memory()
        p_buf = p_keep.numpy()
    else:
        u_buf, s_buf, p_buf = units, spectra, np.zeros((1, 2, 1152), dtype=np.float32)
    one = runs.copy()
    one["n_frames"] = 1
    engine.mp3_streams_alloc(1)
    got = np.zeros((F, 2, 1152), dtype=np.float32)
    for f in range(F):
        launches = engine.launch_count
        engine.mp3_synth_host(u_buf[f:f + 1], s_buf[f:f + 1], one, out=p_buf)
        assert engine.launch_count == launches + 1
        got[f] = p_buf[0]
    _compare(got, want, f"one packet per call, pinned={pinned}")
    engine.close()
