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
