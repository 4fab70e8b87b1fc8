"""Error behaviour of the C ABI on a real device: malformed batches are rejected with the documented
status codes (never a crash, never silent garbage), and empty batches are no-ops."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import symphonia_b200 as sb
    eng = sb.Engine(0)
    yield eng
    eng.close()


def test_mp3_argument_and_limit_errors(engine):
    import symphonia_b200 as sb
    from symphonia_b200 import workloads
    units, spectra, runs = workloads.mp3_batch(3, 4, seed=5)
    engine.mp3_streams_alloc(2)                       # only 2 state slots, the batch names stream 2
    with pytest.raises(sb.SymgpuError) as e:
        engine.mp3_synth_host(units, spectra, runs)
    assert e.value.status == 3                        # SYMGPU_ERR_LIMIT
    engine.mp3_streams_alloc(3)
    bad = runs.copy()
    bad["n_frames"][0] = 3                            # runs no longer tile the batch
    with pytest.raises(sb.SymgpuError) as e:
        engine.mp3_synth_host(units, spectra, bad)
    assert e.value.status == 6                        # SYMGPU_ERR_ARG
    bad = runs.copy()
    bad["granules_per_frame"][1] = 3
    with pytest.raises(sb.SymgpuError) as e:
        engine.mp3_synth_host(units, spectra, bad)
    assert e.value.status == 6
    # still healthy afterwards
    out = engine.mp3_synth_host(units, spectra, runs)
    assert np.isfinite(out).all() and np.abs(out).max() > 0
    lib = sb.lib()
    assert lib.symgpu_mp3_synth_host(engine._ctx, None, None, None, 0, 0, None) == 6
    assert lib.symgpu_strerror(1) == b"symgpu: malformed synthesis unit"


def test_empty_batches_are_noops(engine):
    import symphonia_b200 as sb
    lib = sb.lib()
    engine.mp3_streams_alloc(1)
    one = np.zeros(4, dtype=np.float32)
    p = one.ctypes.data_as(ctypes.c_void_p)
    before = engine.launch_count
    assert lib.symgpu_mp3_synth_host(engine._ctx, p, p, p, 0, 0, p) == 0
    assert engine.launch_count == before


def test_vorbis_rejects_unsupported_configurations(engine):
    import symphonia_b200 as sb
    from symphonia_b200._native import VORBIS_STREAM_DTYPE
    s = np.zeros(1, dtype=VORBIS_STREAM_DTYPE)
    s["bs0_exp"], s["bs1_exp"], s["channels"] = 8, 11, 6     # 5.1: not supported in this version
    with pytest.raises(sb.SymgpuError) as e:
        engine.vorbis_streams_set(s)
    assert e.value.status == 2                        # SYMGPU_ERR_UNSUPPORTED
    s["channels"], s["bs0_exp"] = 2, 12                      # blocksize_0 > blocksize_1
    with pytest.raises(sb.SymgpuError) as e:
        engine.vorbis_streams_set(s)
    assert e.value.status == 6


def test_vorbis_floor_setups_are_validated(engine):
    """The kernel divides by x differences and sweeps posts by dependency level: a setup that the reference's
    own parser would have refused (floor.rs:300-420) must be refused here, not hang or corrupt a launch."""
    import symphonia_b200 as sb
    from symphonia_b200 import workloads
    from symphonia_b200._native import VORBIS_FLOOR1_DTYPE
    good = np.array([workloads.make_floor1_setup([0, 128, 64, 32, 96, 16], 2)], dtype=VORBIS_FLOOR1_DTYPE)
    engine.vorbis_floors_set(good)

    def rejected(mutate):
        bad = good.copy()
        mutate(bad[0])
        with pytest.raises(sb.SymgpuError) as e:
            engine.vorbis_floors_set(bad)
        assert e.value.status == 6

    def dup_x(f):
        f["x_list"][3] = f["x_list"][2]                 # two posts at the same x: a zero-length segment

    def self_neighbour(f):
        f["low"][4] = 4                                  # neighbour that is not an earlier post

    def wrong_side(f):
        f["low"][2], f["high"][2] = f["high"][2], f["low"][2]

    def unsorted(f):
        f["sort_order"][1], f["sort_order"][2] = f["sort_order"][2], f["sort_order"][1]

    def not_a_permutation(f):
        f["sort_order"][2] = f["sort_order"][1]

    def no_origin(f):
        f["x_list"][0] = 5

    for m in (dup_x, self_neighbour, wrong_side, unsorted, not_a_permutation, no_origin):
        rejected(m)
    engine.vorbis_floors_set(good)  # the context still accepts a valid set afterwards
