"""GPU parity for the AAC-LC and Vorbis kernels (through the C ABI) vs the oracle: bit-exact PCM."""
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


def _cmp(got, want, what):
    g = np.ascontiguousarray(got, dtype=np.float32).view(np.uint32)
    w = np.ascontiguousarray(want, dtype=np.float32).view(np.uint32)
    bad = g != w
    if bad.any():
        idx = tuple(int(x[0]) for x in np.nonzero(bad))
        raise AssertionError(f"{what}: {int(bad.sum())} of {g.size} PCM words differ; first at {idx}: gpu {got[idx]!r} "
                             f"oracle {want[idx]!r}")
    assert np.abs(want).max() > 1e-6


# ---- AAC -------------------------------------------------------------------------------------------

def _aac_case(engine, oracle, S, F, **kw):
    from symphonia_b200 import workloads
    units, tns, coeffs, runs = workloads.aac_batch(S, F, **kw)
    rc, want = _oracle.aac_batch(oracle, units, tns, coeffs, runs, S)
    engine.aac_streams_alloc(S)
    got = engine.aac_synth_host(units, tns, coeffs, runs)
    ch = kw.get("channels", 2)
    _cmp(got[:, :ch], want[:, :ch], f"aac S={S} F={F} {kw}")
    return units, tns, coeffs, runs, want


def test_aac_mixed(engine, oracle):
    _aac_case(engine, oracle, 6, 40, seed=101)


def test_aac_long_only_no_tns(engine, oracle):
    _aac_case(engine, oracle, 3, 12, seed=102, tns_prob=0.0, block_switching=False)


def test_aac_heavy_tns(engine, oracle):
    _aac_case(engine, oracle, 3, 20, seed=103, tns_prob=0.9)


def test_aac_chunk_boundaries_and_single_frames(engine, oracle):
    for F in (1, 7, 8, 9, 17):
        _aac_case(engine, oracle, 5, F, seed=110 + F)


def test_aac_mono(engine, oracle):
    _aac_case(engine, oracle, 3, 10, seed=104, channels=1)


def test_aac_state_carry_and_reset(engine, oracle):
    from symphonia_b200 import workloads
    S, F = 4, 21
    units, tns, coeffs, runs = workloads.aac_batch(S, F, seed=105)
    rc, want = _oracle.aac_batch(oracle, units, tns, coeffs, runs, S)
    engine.aac_streams_alloc(S)
    u3, c3 = units.reshape(S, F, 2), coeffs.reshape(S, F, 2, 1024)
    got = np.zeros((S, F, 2, 1024), dtype=np.float32)
    lo = 0
    for part in (5, 1, 15):
        hi = lo + part
        r = runs.copy()
        r["first_frame"] = np.arange(S) * part
        r["n_frames"] = part
        out = engine.aac_synth_host(np.ascontiguousarray(u3[:, lo:hi]).reshape(-1, 2), tns,
                                    np.ascontiguousarray(c3[:, lo:hi]).reshape(-1, 2, 1024), r)
        got[:, lo:hi] = out.reshape(S, part, 2, 1024)
        lo = hi
    _cmp(got.reshape(S * F, 2, 1024), want, "aac state carry")
    for s in range(S):
        engine.aac_stream_reset(s)
    again = engine.aac_synth_host(units, tns, coeffs, runs)
    _cmp(again, want, "aac after reset")


def test_aac_device_entry_point(engine, oracle):
    import torch
    from symphonia_b200 import workloads
    S, F = 3, 10
    units, tns, coeffs, runs = workloads.aac_batch(S, F, seed=106)
    rc, want = _oracle.aac_batch(oracle, units, tns, coeffs, runs, S)
    engine.aac_streams_alloc(S)
    dev = torch.device("cuda", 0)
    u_t = torch.from_numpy(units.view(np.uint8).reshape(-1).copy()).to(dev)
    t_t = torch.from_numpy(tns.view(np.uint8).reshape(-1).copy()).to(dev)
    c_t = torch.from_numpy(coeffs).to(dev)
    p_t = torch.zeros((S * F, 2, 1024), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    engine.aac_synth_dev(u_t, t_t, len(tns), c_t, runs, p_t)
    engine.sync()
    _cmp(p_t.cpu().numpy(), want, "aac device entry point")


# ---- Vorbis ----------------------------------------------------------------------------------------

def _vorbis_case(engine, oracle, S, F, **kw):
    from symphonia_b200 import workloads
    wl = workloads.vorbis_batch(S, F, **kw)
    rc, want = _oracle.vorbis_batch(oracle, wl)
    engine.vorbis_streams_set(wl["streams"])
    engine.vorbis_floors_set(wl["floors"])
    got = engine.vorbis_synth_host(wl["units"], wl["floor_y"], wl["residue"], wl["runs"], wl["slot"])
    ch = kw.get("channels", 2)
    mask = np.arange(wl["slot"])[None, None, :] < wl["out_len"][:, None, None]
    mask = np.broadcast_to(mask, got.shape).copy()
    mask[:, ch:, :] = False
    _cmp(np.where(mask, got, 0), np.where(mask, want, 0), f"vorbis S={S} F={F} {kw}")
    return wl, want


def test_vorbis_mixed(engine, oracle):
    _vorbis_case(engine, oracle, 6, 40, seed=201)


def test_vorbis_floor_with_65_posts(engine, oracle):
    """The largest floor-1 setup (floor.rs:455-560): post 64 has a step-2 flag of its own in the kernel (a 64-bit mask cannot hold
    it); most of its Y values are non-zero here so that the flag matters."""
    wl, _ = _vorbis_case(engine, oracle, 4, 24, seed=207, posts=65, unused_prob=0.0)
    assert wl["floors"]["n_posts"].max() == 65            # the long-block setups (a 256-sample block has room for 64 posts)
    last = wl["floor_y"].reshape(-1, 65)[:, 64]
    assert (last != 0).mean() > 0.25


@pytest.mark.parametrize("bs", [(6, 6), (6, 9), (7, 10), (8, 11), (9, 12), (8, 13), (11, 11)])
def test_vorbis_block_sizes(engine, oracle, bs):
    _vorbis_case(engine, oracle, 2, 12, seed=210 + bs[0] + 16 * bs[1], bs_exp=bs)


def test_vorbis_uncoupled_and_mono(engine, oracle):
    _vorbis_case(engine, oracle, 3, 14, seed=202, coupled=False)
    _vorbis_case(engine, oracle, 3, 14, seed=203, channels=1)


def test_vorbis_many_unused_floors(engine, oracle):
    _vorbis_case(engine, oracle, 3, 20, seed=204, unused_prob=0.5)


def test_vorbis_chunk_boundaries(engine, oracle):
    for F in (1, 7, 8, 9, 17):
        _vorbis_case(engine, oracle, 4, F, seed=220 + F)


def test_vorbis_state_carry(engine, oracle):
    from symphonia_b200 import workloads
    S, F = 4, 19
    wl = workloads.vorbis_batch(S, F, seed=205)
    rc, want = _oracle.vorbis_batch(oracle, wl)
    engine.vorbis_streams_set(wl["streams"])
    engine.vorbis_floors_set(wl["floors"])
    slot = wl["slot"]
    u2 = wl["units"].reshape(S, F)
    fy = wl["floor_y"].reshape(S, F, 2, 65)
    rs = wl["residue"].reshape(S, F, 2, slot)
    got = np.zeros((S, F, 2, slot), dtype=np.float32)
    lo = 0
    for part in (6, 1, 12):
        hi = lo + part
        r = wl["runs"].copy()
        r["first_packet"] = np.arange(S) * part
        r["n_packets"] = part
        out = engine.vorbis_synth_host(np.ascontiguousarray(u2[:, lo:hi]).reshape(-1),
                                       np.ascontiguousarray(fy[:, lo:hi]).reshape(-1, 2, 65),
                                       np.ascontiguousarray(rs[:, lo:hi]).reshape(-1, 2, slot), r, slot)
        got[:, lo:hi] = out.reshape(S, part, 2, slot)
        lo = hi
    mask = np.arange(slot)[None, None, :] < wl["out_len"][:, None, None]
    got = got.reshape(S * F, 2, slot)
    _cmp(np.where(mask, got, 0), np.where(mask, want, 0), "vorbis state carry")


# ---- BASELINE sizes: every frame / packet of the 8192-unit batches against the (multi-threaded) oracle --------------------

def test_aac_full_size_matches_oracle(engine, oracle):
    import os
    from symphonia_b200 import workloads
    S, F = 64, 128
    units, tns, coeffs, runs = workloads.aac_batch(S, F, seed=workloads.SEED_BASE + 2)
    rc, want = _oracle.aac_batch(oracle, units, tns, coeffs, runs, S, n_threads=min(os.cpu_count() or 1, S))
    assert rc == 0
    engine.aac_streams_alloc(S)
    got = engine.aac_synth_host(units, tns, coeffs, runs)
    _cmp(got, want, "aac 8192 frames")


def test_vorbis_full_size_matches_oracle(engine, oracle):
    import os
    from symphonia_b200 import workloads
    S, F = 64, 128
    wl = workloads.vorbis_batch(S, F, seed=workloads.SEED_BASE + 3)
    rc, want = _oracle.vorbis_batch(oracle, wl, n_threads=min(os.cpu_count() or 1, S))
    assert rc == 0
    engine.vorbis_streams_set(wl["streams"])
    engine.vorbis_floors_set(wl["floors"])
    got = engine.vorbis_synth_host(wl["units"], wl["floor_y"], wl["residue"], wl["runs"], wl["slot"])
    mask = np.broadcast_to(np.arange(wl["slot"])[None, None, :] < wl["out_len"][:, None, None], got.shape)
    _cmp(np.where(mask, got, 0), np.where(mask, want, 0), "vorbis 8192 packets")


# ---- Vorbis with more than two channels and several coupling steps (symgpu_vorbis_mc_*) ---------------------------------

def _vorbis_mc_case(engine, oracle, **kw):
    from symphonia_b200 import workloads
    wl = workloads.vorbis_mc_batch(**kw)
    rc, want = _oracle.vorbis_mc_batch(oracle, wl)
    assert rc == 0
    engine.vorbis_mc_streams_set(wl["streams"])
    engine.vorbis_floors_set(wl["floors"])
    got = engine.vorbis_mc_synth_host(wl["units"], wl["floor_y"], wl["residue"], wl["runs"], wl["channels"], wl["slot"])
    mask = np.broadcast_to(np.arange(wl["slot"])[None, None, :] < wl["out_len"][:, None, None], got.shape)
    _cmp(np.where(mask, got, 0), np.where(mask, want, 0), f"vorbis multichannel {kw}")


def test_vorbis_5_1_with_three_coupling_steps(engine, oracle):
    """Six channels, three coupling steps in which channel 0 takes part twice (the order of the steps matters: the reference walks
    the mapping's list front to back, lib.rs:252-278), four streams with long / short block switching."""
    _vorbis_mc_case(engine, oracle, n_streams=4, packets_per_stream=24, seed=301)


@pytest.mark.parametrize("channels,couplings", [(1, ()), (2, ((0, 1),)), (3, ((0, 1), (2, 0))), (5, ((0, 1), (2, 3))), (8, ((0, 1), (2, 3), (4, 5), (7, 6), (0, 7)))])
def test_vorbis_channel_counts(engine, oracle, channels, couplings):
    _vorbis_mc_case(engine, oracle, n_streams=3, packets_per_stream=10, seed=310 + channels, channels=channels, couplings=couplings)


@pytest.mark.parametrize("variant", ["pair", "warp"])
def test_aac_older_kernel_shapes_match_too(variant):
    """`SYMGPU_AAC_KERNEL=pair` (two warps per frame, named barriers: round 1's shape) and `=warp` (one warp per frame, array
    layout; measured slower) stay selectable next to the default Z layout.  The variable is read once per process, so
    the AAC cases above are re-run in a child process with it set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SYMGPU_AAC_KERNEL=variant)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_aac_mixed or test_aac_chunk_boundaries or test_aac_state_carry or test_aac_heavy_tns or test_aac_mono"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("mode", ["inline", "sorted"])
def test_aac_other_tns_arrangements_match_too(mode):
    """`SYMGPU_AAC_TNS=inline`: the Z kernel runs a frame's TNS filters on the lanes of the frame's warp instead of a pre-pass
    (faster when few frames carry filters); `=sorted`: round 1's three-kernel pre-pass.  The default is the one-kernel
    pre-pass (a warp per filtered channel-frame).  Read once per process, so the TNS cases are re-run in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SYMGPU_AAC_TNS=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_aac_mixed or test_aac_heavy_tns or test_aac_chunk_boundaries or test_aac_device_entry_point"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_vorbis_array_layout_kernel_matches_too():
    """`SYMGPU_VORBIS_KERNEL=pair`: round 1's kernel (64 threads per packet, array layout) stays selectable next to the default Z
    layout; the Vorbis cases above are re-run in a child process with it set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, SYMGPU_VORBIS_KERNEL="pair")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_vorbis_mixed or test_vorbis_block_sizes or test_vorbis_state_carry or test_vorbis_5_1 or test_vorbis_floor_with_65"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_vorbis_small_ctas_in_a_fresh_process():
    """Batches of very short runs get small CTAs (two packet slots = 128 threads for one-packet runs).  Shared memory keeps what an
    earlier, larger launch of the same process left there, so a table that a small CTA fails to fill completely goes unnoticed
    unless the small launch comes first: the one-packet cases run here in a process of their own."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", "test_vorbis_chunk_boundaries"],
                       cwd=root, env=dict(os.environ), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
