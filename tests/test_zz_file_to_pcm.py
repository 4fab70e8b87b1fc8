"""The whole chain SURVEY §8 draws: file bytes -> packetiser (N2) -> entropy front-end (N1) -> fused synthesis kernel
-> PCM, against the oracle's synthesis of the same units.  The CPU test checks everything up to the launch (the
tables are well-formed, the oracle synthesises them); the GPU test adds the device and compares every PCM word.
(Named to run last: it spans every layer below it.)"""
import numpy as np
import pytest

from symphonia_b200 import _native as nat
from symphonia_b200 import frontend, packetizer
from tests import _mp3_bitstream as bw
from tests import _oracle
from tests import _streams as st


def _corpus():
    """Three files: MPEG-1 joint stereo with a LAME tag and junk, MPEG-2 mono, MPEG-1 dual channel with CRCs."""
    rng = np.random.default_rng(77)
    files = []
    frames, _ = bw.gen_stream(rng, 30, version="1", mode=1, bitrate_idx=9, pair_blocks=True)
    tag = st.mpa_tag_frame(rng, dict(version="1", layer=3, bitrate_idx=9, rate_idx=0, mode=1), num_frames=30)
    noise = rng.integers(0, 255, 200, dtype=np.uint8).tobytes()
    files.append(noise + tag + b"".join(frames[:17]) + noise[:33] + b"".join(frames[17:]))
    frames, _ = bw.gen_stream(rng, 24, version="2", mode=3, bitrate_idx=8, rate_idx=1)
    files.append(b"".join(frames))
    frames, _ = bw.gen_stream(rng, 20, version="1", mode=2, bitrate_idx=11, rate_idx=1, protected=True)
    files.append(b"".join(frames))
    return files


def _config0_file():
    """BASELINE.json configs[0]: MP3 CBR 320 kbit/s, 44.1 kHz, stereo, one stream."""
    rng = np.random.default_rng(79)
    frames, _ = bw.gen_stream(rng, 40, version="1", mode=0, bitrate_idx=14, rate_idx=0, padding=0, fill=(0.7, 1.0))
    assert all(len(f) == 1044 for f in frames)
    return b"".join(frames)


def _batch(files):
    units, quant, runs, spans = [], [], [], []
    at = 0
    for s, data in enumerate(files):
        track, packets = packetizer.mpa_index(data)
        fe = frontend.Mp3Frontend()
        u, q, frame_of, info = fe.decode_packets(data, packets)
        assert len(frame_of) == len(packets)  # every packet of these files decodes
        units.append(u), quant.append(q)
        runs.append((s, at, len(u), int(info["granules"]), int(info["channels"]), 0))
        kept = packets[frame_of]
        spans.append((kept["dur"].astype(np.int64), kept["trim_start"].astype(np.int64), np.minimum(kept["trim_end"], kept["dur"]).astype(np.int64)))
        at += len(u)
    return np.concatenate(units), np.concatenate(quant), np.array(runs, dtype=nat.MP3_RUN_DTYPE), spans


def _spectra(quant):
    pow43 = nat.mp3_pow43()
    mag = pow43[np.abs(quant.astype(np.int32))]
    return np.where(quant < 0, -mag, mag).astype(np.float32)  # read_huffman_samples: (1 - 2 * sign) * POW43[x], +0.0 for x = 0


def test_chain_up_to_the_launch(oracle):
    files = _corpus()
    units, quant, runs, spans = _batch(files)
    assert len(units) == 74 and runs["granules_per_frame"].tolist() == [2, 1, 2] and runs["channels"].tolist() == [2, 1, 2]
    assert nat.lib().symgpu_mp3_units_check(units.ctypes.data, runs.ctypes.data, len(runs), len(units)) == 0
    rc, pcm, _ = _oracle.mp3_batch(oracle, units.reshape(-1), _spectra(quant), runs, len(files))
    assert rc == 0 and np.isfinite(pcm).all() and np.abs(pcm).max() > 0
    # gapless: the LAME tag's delay / padding arrive as per-packet trims that leave exactly the tagged length
    dur, t0, t1 = spans[0]
    assert int((dur - t0 - t1).sum()) == 30 * 1152 - 1105 - 471 and int(t0[0]) == 1105


def _device_vs_oracle(oracle, files):
    import symphonia_b200 as sb
    units, quant, runs, _ = _batch(files)
    rc, want, _ = _oracle.mp3_batch(oracle, units.reshape(-1), _spectra(quant), runs, len(files))
    assert rc == 0
    with sb.Engine(0) as eng:
        eng.mp3_streams_alloc(len(files))
        got = eng.mp3_synth_host_quantized(units.reshape(-1), quant, runs)
        assert eng.launch_count >= 2  # dequantise + synthesis
    same = got.view(np.uint32) == want.view(np.uint32)
    assert same.all(), f"{int((~same).sum())} PCM words differ, first at {np.argwhere(~same)[0]}"


@pytest.mark.gpu
def test_file_bytes_to_pcm_on_the_device(oracle):
    """Joint stereo (both channels share their window sequence, as the format requires) and MPEG-2 mono."""
    _device_vs_oracle(oracle, _corpus()[:2])


@pytest.mark.gpu
def test_file_bytes_to_pcm_independent_channels(oracle):
    """Dual channel: the two channels of a granule choose block types independently.  The synthetic workloads of the
    kernel parity tests never do that (they draw one window sequence per stream), so this is the first time the
    kernel's per-channel handling meets the oracle; kept apart so that a difference here is not mistaken for one above."""
    _device_vs_oracle(oracle, _corpus()[2:])


# ------------------------------------------------------------------------------------------- Layer I / II files

def _mpa12_corpus():
    from tests import _mpa12_bitstream as b12
    rng = np.random.default_rng(78)
    l2 = [b12.gen_layer2_frame(rng, "1", 12, 0, 1, mode_ext=k % 4)[0] for k in range(20)]   # 256 kbit/s joint stereo, table b
    l2m = [b12.gen_layer2_frame(rng, "2", 6, 1, 3)[0] for _ in range(16)]                    # MPEG-2 mono
    l1 = [b12.gen_layer1_frame(rng, "1", 9, 1, 0)[0] for _ in range(24)]
    noise = rng.integers(0, 255, 90, dtype=np.uint8).tobytes()
    return [(2, noise + b"".join(l2)), (2, b"".join(l2m)), (1, b"".join(l1[:10]) + noise + b"".join(l1[10:]))]


def _mpa12_batches():
    out = []
    for layer, data in _mpa12_corpus():
        track, packets = packetizer.mpa_index(data)
        assert int(track["layer"]) == layer
        sub, frame_of, info = frontend.mpa12_decode_packets(data, packets, layer)
        assert len(frame_of) == len(packets) > 0
        runs = np.zeros(1, dtype=nat.MPA12_RUN_DTYPE)
        runs[0] = (0, 0, len(sub), int(info["channels"]), (0, 0, 0))
        out.append((layer, sub, runs))
    return out


def test_layer12_chain_up_to_the_launch(oracle):
    for layer, sub, runs in _mpa12_batches():
        rc, pcm, _ = _oracle.mpa12_batch(oracle, sub, runs, 1)
        assert rc == 0 and np.isfinite(pcm).all() and np.abs(pcm).max() > 0.01
        assert not pcm[:, :, 32 * sub.shape[-1]:].any()  # a Layer I frame fills 384 of the 1152 samples of its slot


@pytest.mark.gpu
def test_layer12_file_bytes_to_pcm_on_the_device(oracle):
    import symphonia_b200 as sb
    with sb.Engine(0) as eng:
        for layer, sub, runs in _mpa12_batches():
            rc, want, _ = _oracle.mpa12_batch(oracle, sub, runs, 1)
            assert rc == 0
            eng.mp3_streams_alloc(1)
            got = eng.mpa12_synth_host(sub, runs)
            n = 32 * sub.shape[-1]
            ch = int(runs[0]["channels"])
            same = got[:, :ch, :n].view(np.uint32) == want[:, :ch, :n].view(np.uint32)
            assert same.all(), (layer, int((~same).sum()))


# ------------------------------------------------------------------------------------------- the C++ plug-in interface on files

@pytest.mark.gpu
def test_cpp_decoder_on_real_files(tmp_path, oracle):
    """include/symgpu/decoder.hpp end to end in C++: MpaIndexer cuts the file, the registry hands out GpuMpaDecoder at
    Tier::Preferred, decode() takes one real frame at a time (front-end on the CPU, synthesis on the GPU) and applies the
    gapless trims the packetiser derived from the LAME tag -- against the oracle's PCM with the same trims."""
    import subprocess

    from tests import test_cpp_host
    exe = test_cpp_host._build()
    inp, outp = tmp_path / "in.mp3", tmp_path / "out.bin"
    # Layer III: the tagged joint-stereo file, and BASELINE config 0 -- MP3 CBR 320 kbit/s 44.1 kHz stereo, one stream, one
    # Decoder::decode call per packet ("plumbing") -- with real frames
    for data, note in ((_corpus()[0], "delay 1105 padding 471"), (_config0_file(), "decoded 40 of 40 packets")):
        units, quant, runs, spans = _batch([data])
        rc, want, _ = _oracle.mp3_batch(oracle, units.reshape(-1), _spectra(quant), runs, 1)
        assert rc == 0
        dur, t0, t1 = spans[0]
        expect = b"".join(want[k, ch, int(t0[k]):int(dur[k] - t1[k])].tobytes() for k in range(len(want)) for ch in range(2))
        inp.write_bytes(data)
        res = subprocess.run([exe, "file", "3", str(inp), str(outp)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert note in res.stdout
        assert outp.read_bytes() == expect
    # Layer II and Layer I
    for (layer, blob), (_, sub, runs12) in zip(_mpa12_corpus(), _mpa12_batches()):
        rc, want, _ = _oracle.mpa12_batch(oracle, sub, runs12, 1)
        assert rc == 0
        ch, n = int(runs12[0]["channels"]), 32 * sub.shape[-1]
        _, pk = packetizer.mpa_index(blob)  # an untagged file still gets an (extrapolated) length, hence possibly an end trim
        assert len(pk) == len(want)
        cut = [(int(p["trim_start"]), n - min(int(p["trim_end"]), n - int(p["trim_start"]))) for p in pk]
        expect = b"".join(want[k, c, a:b].tobytes() for k, (a, b) in enumerate(cut) for c in range(ch))
        inp.write_bytes(blob)
        res = subprocess.run([exe, "file", str(layer), str(inp), str(outp)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert outp.read_bytes() == expect, layer


# ------------------------------------------------------------------------------------------- the one-call file decoder

def _decode_expect(oracle, data, fmt):
    from symphonia_b200 import decode
    layer, payload, runs, spans, rate, channels, total = decode.mpeg_audio_plan(data)
    if layer == 3:
        rc, pcm, _ = _oracle.mp3_batch(oracle, payload[0], _spectra(payload[1]), runs, 1)
    else:
        rc, pcm, _ = _oracle.mpa12_batch(oracle, payload, runs, 1)
    assert rc == 0
    return _oracle.pcm_pack(oracle, pcm, spans, channels, fmt, total), rate, channels, total


def test_one_call_decoder_plan(oracle):
    """CPU half of symphonia_b200.decode: spans tile the output exactly and the oracle renders them."""
    for data in [_corpus()[0], _corpus()[1]] + [blob for _, blob in _mpa12_corpus()]:
        want, rate, channels, total = _decode_expect(oracle, data, nat.FMT_S16)
        assert want.shape == (total, channels) and rate in (44100, 48000, 24000) and np.abs(want).max() > 0
    want, rate, channels, total = _decode_expect(oracle, _corpus()[0], nat.FMT_S16)
    assert total == 30 * 1152 - 1105 - 471  # the LAME tag's delay and padding are gone


@pytest.mark.gpu
def test_one_call_decoder_on_the_device(oracle):
    import symphonia_b200 as sb
    from symphonia_b200 import decode
    with sb.Engine(0) as eng:
        eng.mp3_streams_alloc(2)
        for data in [_corpus()[0], _corpus()[1]] + [blob for _, blob in _mpa12_corpus()]:
            for fmt in (nat.FMT_S16, nat.FMT_F32):
                want, rate, channels, total = _decode_expect(oracle, data, fmt)
                got, got_rate = decode.decode_mpeg_audio(eng, data, fmt, stream=1)
                assert got_rate == rate and got.shape == want.shape
                assert (got.view(np.uint8) == want.view(np.uint8)).all()
