"""GPU parity of the MPEG Layer I / II polyphase synthesis against the oracle's restatement of
synthesis::synthesis (synthesis.rs:158-344), bit for bit."""
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


def _same(got, want, n_slots, channels, what):
    g = got[:, :channels, :32 * n_slots].view(np.uint32)
    w = want[:, :channels, :32 * n_slots].view(np.uint32)
    bad = g != w
    assert not bad.any(), f"{what}: {int(bad.sum())} of {g.size} PCM words differ, first at {np.argwhere(bad)[0]}"
    assert np.abs(want).max() > 1e-3


@pytest.mark.parametrize("layer", [1, 2])
@pytest.mark.parametrize("shape", [(5, 30), (40, 1), (3, 2), (2, 100)])  # chains with carries, single frames, short runs, long runs
def test_layers_match_the_oracle(engine, oracle, layer, shape):
    from symphonia_b200 import workloads
    S, F = shape
    x, runs = workloads.mpa12_batch(S, F, layer=layer, seed=700 + 10 * layer + S)
    rc, want, _ = _oracle.mpa12_batch(oracle, x, runs, S)
    assert rc == 0
    engine.mp3_streams_alloc(S)
    got = engine.mpa12_synth_host(x, runs)
    _same(got, want, x.shape[-1], 2, f"layer {layer} S={S} F={F}")


def test_mono_and_state_across_batches(engine, oracle):
    from symphonia_b200 import workloads
    S, F = 4, 24
    x, runs = workloads.mpa12_batch(S, F, layer=2, seed=731, channels=1)
    rc, want, _ = _oracle.mpa12_batch(oracle, x, runs, S)
    engine.mp3_streams_alloc(S)
    _same(engine.mpa12_synth_host(x, runs), want, 36, 1, "mono")
    # two batches of 10 + 14 frames per stream continue where one batch of 24 would be
    x, runs = workloads.mpa12_batch(S, F, layer=1, seed=732)
    rc, want, _ = _oracle.mpa12_batch(oracle, x, runs, S)
    engine.mp3_streams_alloc(S)
    xs = x.reshape(S, F, 2, 32, 12)
    outs = []
    for lo, hi in ((0, 10), (10, 24)):
        r = runs.copy()
        r["first_frame"] = np.arange(S) * (hi - lo)
        r["n_frames"] = hi - lo
        outs.append(engine.mpa12_synth_host(np.ascontiguousarray(xs[:, lo:hi]).reshape(-1, 2, 32, 12), r).reshape(S, hi - lo, 2, 1152))
    got = np.concatenate(outs, axis=1).reshape(S * F, 2, 1152)
    _same(got, want, 12, 2, "state carried across batches")
    # reset: the stream starts from silence again
    engine.mp3_stream_reset(2)
    r1 = runs[2:3].copy()
    r1["first_frame"] = 0
    got_reset = engine.mpa12_synth_host(np.ascontiguousarray(xs[2]), r1)
    _same(got_reset, want.reshape(S, F, 2, 1152)[2], 12, 2, "after reset")


def test_bad_slot_count_is_rejected(engine):
    import symphonia_b200 as sb
    from symphonia_b200 import workloads
    x, runs = workloads.mpa12_batch(1, 2, layer=2, seed=733)
    with pytest.raises(sb.SymgpuError) as e:
        engine.mpa12_synth_host(np.zeros((2, 2, 32, 18), np.float32), runs)
    assert e.value.status == 6
