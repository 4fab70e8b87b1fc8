"""GPU parity of the output stage (trim + interleave + sample conversion) and of a mixed
MP3 + AAC + Vorbis corpus on one context.  Integer samples must be identical to the oracle's; f32 samples
bit-identical (uint32 view)."""
import numpy as np
import pytest

from symphonia_b200._native import FMT_F32, FMT_NUMPY, FMT_S16, FMT_S24, FMT_S32, FMT_U8, PCM_SPAN_DTYPE
from tests import _oracle

pytestmark = pytest.mark.gpu
ALL_FORMATS = (FMT_F32, FMT_S16, FMT_S24, FMT_S32, FMT_U8)


@pytest.fixture(scope="module")
def engine():
    import symphonia_b200 as sb
    eng = sb.Engine(0)
    yield eng
    eng.close()


def _same(got, want, what):
    assert got.dtype == want.dtype and got.shape == want.shape
    g = got.view(np.uint32) if got.dtype == np.float32 else got
    w = want.view(np.uint32) if want.dtype == np.float32 else want
    bad = g != w
    assert not bad.any(), f"{what}: {int(bad.sum())} of {g.size} samples differ, first at {np.argwhere(bad)[0]}"


def _nasty_pcm(rng, shape):
    """PCM with everything the conversion rules distinguish: in range, beyond +-1, exact +-1 and 0 and -0,
    values a hair either side of an integer step, subnormals, infinities and NaNs."""
    x = rng.normal(0.0, 0.7, size=shape).astype(np.float32)
    flat = x.reshape(-1)
    n = flat.size
    special = np.array([1.0, -1.0, 0.0, -0.0, np.inf, -np.inf, np.nan, 1e-42, -1e-42, 0.99999994, -0.99999994,
                        1.0000001, -1.0000001, 3.0517578e-05, -3.0517578e-05, 0.5, -0.5], dtype=np.float32)
    idx = rng.choice(n, size=min(n // 4, 4096), replace=False)
    flat[idx] = special[rng.integers(0, len(special), size=len(idx))]
    k = rng.integers(-32768, 32768, size=min(n // 4, 4096)).astype(np.float32)
    idx2 = rng.choice(n, size=len(k), replace=False)
    flat[idx2] = np.nextafter(k / np.float32(32768.0), rng.choice(np.array([-2.0, 2.0], dtype=np.float32), size=len(k)))
    return x


@pytest.mark.parametrize("fmt", ALL_FORMATS)
def test_uniform_stereo_packets(engine, oracle, fmt):
    rng = np.random.default_rng(100 + fmt)
    pcm = _nasty_pcm(rng, (37, 2, 1152))
    got = engine.pcm_pack_host(pcm, None, 2, fmt, 37 * 1152, plane_stride=1152, frames=1152, n_spans=37)
    want = _oracle.pcm_pack(oracle, pcm, None, 2, fmt, 37 * 1152, plane_stride=1152, frames=1152, n_spans=37)
    _same(got, want, f"uniform stereo fmt={fmt}")
    assert np.count_nonzero(want) > want.size // 2


@pytest.mark.parametrize("channels", [1, 2, 3, 6, 8])
def test_trimmed_spans_any_channel_count(engine, oracle, channels):
    rng = np.random.default_rng(200 + channels)
    n_packets, slot = 23, 1024
    pcm = _nasty_pcm(rng, (n_packets, channels, slot))
    spans = np.zeros(n_packets, dtype=PCM_SPAN_DTYPE)
    dst = 5  # the first five output frames belong to nobody: they must keep the caller's bytes
    for p in range(n_packets):
        frames = int(rng.integers(1, slot + 1))
        ts = int(rng.choice([0, 0, 1, 3, 4, 7, frames // 2, frames, frames + 9]))
        te = int(rng.choice([0, 0, 2, 4, 5, frames // 3, frames + 1]))
        spans[p] = (p * channels * slot, slot, frames, ts, te, dst)
        n = max(frames - te, 0)
        dst += 0 if ts >= n else n - ts
    for fmt in ALL_FORMATS:
        sentinel = np.full((dst + 3, channels), 77, dtype=FMT_NUMPY[fmt])
        got = engine.pcm_pack_host(pcm, spans, channels, fmt, dst + 3, out=sentinel.copy())
        want = sentinel.copy()
        ref = _oracle.pcm_pack(oracle, pcm, spans, channels, fmt, dst + 3)
        want[5:dst] = ref[5:dst]
        _same(got, want, f"trimmed spans ch={channels} fmt={fmt}")


def test_argument_errors(engine):
    import symphonia_b200 as sb
    pcm = np.zeros((2, 2, 64), dtype=np.float32)
    with pytest.raises(sb.SymgpuError):  # unknown format
        engine._check(engine._lib.symgpu_pcm_pack_host(engine._ctx, pcm.ctypes.data, pcm.size, None, 2, 2, 64, 64, 9,
                                                       pcm.ctypes.data, pcm.nbytes))
    spans = np.zeros(1, dtype=PCM_SPAN_DTYPE)
    spans[0] = (0, 64, 300, 0, 0, 0)  # reads past the PCM the caller described
    with pytest.raises(sb.SymgpuError):
        engine.pcm_pack_host(pcm, spans, 2, FMT_S16, 300)
    spans[0] = (0, 64, 64, 0, 0, 10)  # writes past `out`
    with pytest.raises(sb.SymgpuError):
        engine.pcm_pack_host(pcm, spans, 2, FMT_S16, 64)
    with pytest.raises(sb.SymgpuError):  # 9 channels
        engine.pcm_pack_host(np.zeros((1, 9, 8), np.float32), None, 9, FMT_S16, 8, plane_stride=8, frames=8, n_spans=1)


@pytest.mark.parametrize("fmt", [FMT_S16, FMT_F32, FMT_S24])
@pytest.mark.parametrize("shape", [(3, 20), (16, 40)])  # single-shot path; pipelined path (>= 512 frames)
def test_mp3_host_packed(engine, oracle, fmt, shape):
    from symphonia_b200 import workloads
    S, F = shape
    units, spectra, runs = workloads.mp3_batch(S, F, seed=300 + S)
    rc, pcm, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
    assert rc == 0
    want = _oracle.pcm_pack(oracle, pcm, None, 2, fmt, S * F * 1152, plane_stride=1152, frames=1152, n_spans=S * F)
    engine.mp3_streams_alloc(S)
    got = engine.mp3_synth_host_packed(units, spectra, runs, fmt)
    _same(got, want, f"mp3 packed fmt={fmt} S={S} F={F}")
    if fmt == FMT_S16:  # the synthetic spectra are loud: plenty of samples clip, plenty do not
        mag = np.abs(got.astype(np.int32))
        assert mag.max() >= 32767 and (mag < 32767).mean() > 0.05 and (mag > 0).mean() > 0.5


@pytest.mark.parametrize("shape", [(3, 20), (16, 40)])
def test_mp3_host_quantized_input(engine, oracle, shape):
    """i16 quantised spectra in (POW43 lookup on the device), f32 planar or i16 interleaved out."""
    from symphonia_b200 import workloads
    S, F = shape
    units, spectra, runs = workloads.mp3_batch(S, F, seed=500 + S)
    quant = workloads.mp3_quantize(spectra)
    # the device table is the oracle's table (f32 powf of an f32 exponent, requantize.rs:23-32)
    ref43 = np.array([oracle.oracle_mp3_pow43(int(i)) for i in np.unique(np.abs(quant))], dtype=np.float32)
    assert (ref43 == workloads._native.mp3_pow43()[np.unique(np.abs(quant))]).all()
    rc, pcm, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
    assert rc == 0 and np.abs(quant).max() > 100
    engine.mp3_streams_alloc(S)
    _same(engine.mp3_synth_host_quantized(units, quant, runs), pcm, "quantised in, f32 out")
    want16 = _oracle.pcm_pack(oracle, pcm, None, 2, FMT_S16, S * F * 1152, plane_stride=1152, frames=1152, n_spans=S * F)
    engine.mp3_streams_alloc(S)
    _same(engine.mp3_synth_host_quantized(units, quant, runs, FMT_S16), want16, "quantised in, i16 out")


def test_mixed_corpus_on_one_context(engine, oracle):
    """SURVEY §8d config 5 in miniature: MP3, AAC and Vorbis streams served by one context back to back, twice
    (state carried across calls), and the Vorbis PCM -- variable frames per packet -- packed to i16 with a
    gapless trim on each stream's first and last packet."""
    from symphonia_b200 import workloads
    S = 5
    mu, ms, mr = workloads.mp3_batch(S, 12, seed=401)
    au, at, ac, ar = workloads.aac_batch(S, 12, seed=402)
    wl = workloads.vorbis_batch(S, 12, seed=403)
    _, want_mp3, _ = _oracle.mp3_batch(oracle, mu, ms, mr, S)
    _, want_aac = _oracle.aac_batch(oracle, au, at, ac, ar, S)
    _, want_vor = _oracle.vorbis_batch(oracle, wl)
    engine.mp3_streams_alloc(S)
    engine.aac_streams_alloc(S)
    engine.vorbis_streams_set(wl["streams"])
    engine.vorbis_floors_set(wl["floors"])

    def halves(arr, per_stream, lo, hi):
        a = arr.reshape((S, per_stream) + arr.shape[1:])
        return np.ascontiguousarray(a[:, lo:hi]).reshape((-1,) + arr.shape[1:])

    got_mp3, got_aac, got_vor = [], [], []
    for lo, hi in ((0, 5), (5, 12)):  # interleave the three codecs call by call
        n = hi - lo
        r = mr.copy(); r["first_frame"] = np.arange(S) * n; r["n_frames"] = n
        got_mp3.append(engine.mp3_synth_host(halves(mu, 12, lo, hi), halves(ms, 12, lo, hi), r).reshape(S, n, 2, 1152))
        # AAC: TNS filter indices are per batch; rebuild them for the slice
        u = au.reshape(S, 12, 2)[:, lo:hi].copy()
        tns_rows = []
        for rec in u.reshape(-1):
            first = len(tns_rows)
            tns_rows.extend(at[rec["tns_first"]:rec["tns_first"] + rec["n_tns"]])
            rec["tns_first"] = first
        ra = ar.copy(); ra["first_frame"] = np.arange(S) * n; ra["n_frames"] = n
        t = np.array(tns_rows, dtype=at.dtype) if tns_rows else at[:0]
        got_aac.append(engine.aac_synth_host(u.reshape(-1, 2), t, halves(ac, 12, lo, hi), ra).reshape(S, n, 2, 1024))
        rv = wl["runs"].copy(); rv["first_packet"] = np.arange(S) * n; rv["n_packets"] = n
        got_vor.append(engine.vorbis_synth_host(halves(wl["units"], 12, lo, hi), halves(wl["floor_y"], 12, lo, hi),
                                                halves(wl["residue"], 12, lo, hi), rv, wl["slot"]).reshape(S, n, 2, wl["slot"]))
    got_mp3 = np.concatenate(got_mp3, axis=1).reshape(want_mp3.shape)
    got_aac = np.concatenate(got_aac, axis=1).reshape(want_aac.shape)
    got_vor = np.concatenate(got_vor, axis=1).reshape(want_vor.shape)
    _same(got_mp3, want_mp3, "mixed corpus: mp3")
    _same(got_aac, want_aac, "mixed corpus: aac")
    mask = np.broadcast_to(np.arange(wl["slot"])[None, None, :] < wl["out_len"][:, None, None], got_vor.shape)
    _same(np.where(mask, got_vor, 0).astype(np.float32), np.where(mask, want_vor, 0).astype(np.float32), "mixed corpus: vorbis")

    # Output stage over the Vorbis packets: frames vary per packet; trim the encoder delay off each stream.
    P = len(wl["units"])
    spans = np.zeros(P, dtype=PCM_SPAN_DTYPE)
    dst = 0
    for p in range(P):
        first, last = p % 12 == 0, p % 12 == 11
        frames = int(wl["out_len"][p])
        ts, te = (min(100, frames) if first else 0), (min(37, frames) if last else 0)
        spans[p] = (p * 2 * wl["slot"], wl["slot"], frames, ts, te, dst)
        dst += max(max(frames - te, 0) - ts, 0)
    got = engine.pcm_pack_host(got_vor, spans, 2, FMT_S16, dst)
    want = _oracle.pcm_pack(oracle, want_vor, spans, 2, FMT_S16, dst)
    _same(got, want, "mixed corpus: vorbis i16 with gapless trim")
