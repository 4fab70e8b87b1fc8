"""FLAC front-end (`symgpu_flac_fe_decode_packets`, SURVEY §8f N1 for the FLAC row) against
oracle/flac_frontend_oracle.py and against the descriptors an exact-integer encoder started from
(`workloads.flac_batch`), serialised by an independent frame writer (tests/_flac_bitstream.py).  Integers throughout.
The strongest property is the format's own: restoring what the front-end read gives back the PCM that was encoded."""
import numpy as np
import pytest

from oracle import flac_frontend_oracle as fo
from oracle.mp3_frontend_oracle import DecodeError
from symphonia_b200 import _native as nat
from symphonia_b200 import frontend, workloads
from tests import _flac_bitstream as fw
from tests import test_oracle_kat_flac as kat


def test_oracle_known_answers_of_the_reference():
    # frame.rs:335-355
    data = bytes([0x24, 0xC2, 0xA2, 0xE0, 0xA4, 0xB9, 0xE2, 0x82, 0xAC, 0xF0, 0x90, 0x8D, 0x88, 0xFF, 0x80, 0xBF])
    at, out = 0, []
    for _ in range(7):
        v, at = fo.utf8_decode(data, at)
        out.append(v)
    assert out == [36, 162, 2361, 8364, 66376, None, None]
    # decoder.rs:642-658
    assert [fo.rice_signed(i) for i in range(11)] == [0, -1, 1, -2, 2, -3, 3, -4, 4, -5, 5] and fo.rice_signed(0xFFFFFFFF) == -2147483648
    assert fo.crc8(b"123456789") == 0xF4  # CRC-8 (polynomial 0x07) catalogue check value
    for v in (0, 127, 128, 2047, 2048, 65535, 65536, (1 << 21) - 1, 1 << 21, (1 << 26) - 1, 1 << 26, (1 << 31) - 1, (1 << 36) - 1):
        assert fo.utf8_decode(fw.utf8_encode(v), 0)[0] == v


def _stream(rng, n_frames, block, bps, channels, seed):
    frames, subs, samples, expect = workloads.flac_batch(n_frames, block, seed=seed, bps=bps, channels=channels, return_pcm=True)
    packets = []
    for f in range(n_frames):
        fr = frames[f]
        ss = subs[int(fr["first_subframe"]):int(fr["first_subframe"]) + int(fr["channels"])]
        packets.append(fw.write_frame(rng, fr, ss, samples, f, stream_bps=bps))
    return packets, frames, subs, samples, expect


def _table(packets):
    t = np.zeros(len(packets), dtype=nat.PIECE_DTYPE)
    at = 0
    for k, p in enumerate(packets):
        t[k]["offset"], t[k]["len"] = at, len(p)
        at += len(p)
    return b"".join(packets), t


def _check_against_oracle(packets, stream_bps=0, stream_channels=0, max_block=0):
    """Same accept / refuse decision per packet; same descriptors and samples for the accepted ones."""
    data, table = _table(packets)
    frames, infos, frame_of, subs, samples = frontend.flac_decode_packets(data, table, stream_bps, stream_channels, max_block)
    accepted, k = [], 0
    for i, p in enumerate(packets):
        try:
            h, want = fo.decode_packet(p, stream_bps, stream_channels, max_block)
        except (DecodeError, fo.Unsupported):
            continue
        accepted.append(i)
        assert k < len(frames) and int(frame_of[k]) == i, (i, frame_of.tolist())
        fr = frames[k]
        assert (int(fr["channels"]), int(fr["assignment"]), int(fr["bits_per_sample"])) == (h["channels"], h["assignment"], h["bps"])
        assert (int(infos[k]["sequence"]), int(infos[k]["block_size"]), int(infos[k]["sample_rate"]), bool(infos[k]["by_sample"])) == \
            (h["sequence"], h["block"], h["rate"] or 0, h["by_sample"])
        for c, w in enumerate(want):
            sf = subs[int(fr["first_subframe"]) + c]
            assert (int(sf["type"]), int(sf["order"]), int(sf["shift"]), int(sf["wasted"]), int(sf["n"])) == (w["type"], w["order"], w["shift"], w["wasted"], h["block"])
            assert [int(v) for v in sf["coeffs"]] == w["coeffs"]
            off = int(sf["offset"])
            assert samples[off:off + h["block"]].tolist() == w["samples"], (i, c)
        k += 1
    assert k == len(frames)
    return accepted, (frames, infos, frame_of, subs, samples)


@pytest.mark.parametrize("bps,channels,block", [(16, 2, 512), (24, 2, 1152), (8, 1, 192), (16, 2, 4096), (20, 3, 300), (12, 2, 97), (32, 1, 256), (16, 8, 64)])
def test_round_trip_is_lossless(oracle, bps, channels, block):
    rng = np.random.default_rng(bps * 100 + block)
    n_frames = 30 if block < 2000 else 8
    packets, frames, subs, samples, expect = _stream(rng, n_frames, block, bps, channels, seed=600 + bps + block)
    accepted, (gf, gi, gof, gs, gsm) = _check_against_oracle(packets, stream_bps=bps)
    assert accepted == list(range(n_frames))
    assert [int(v) for v in gi["sequence"]] == list(range(n_frames))
    # descriptors equal the encoder's (their sample offsets differ: the front-end packs blocks densely)
    for f in range(n_frames):
        for c in range(channels):
            a, b = subs[f * channels + c], gs[int(gf[f]["first_subframe"]) + c]
            assert (int(a["type"]), int(a["wasted"]), int(a["n"])) == (int(b["type"]), int(b["wasted"]), int(b["n"]))
            if int(a["type"]) >= 2:
                assert int(a["order"]) == int(b["order"])
            if int(a["type"]) == 3:
                assert int(a["shift"]) == int(b["shift"]) and (a["coeffs"] == b["coeffs"]).all()
    # restoring what was read gives back the PCM the encoder started from
    rc, pcm = kat._restore(oracle, gf, gs, gsm.copy())
    assert rc == 0
    for f in range(n_frames):
        for c in range(channels):
            a, b = subs[f * channels + c], gs[int(gf[f]["first_subframe"]) + c]
            n = int(a["n"])
            assert (pcm[int(b["offset"]):int(b["offset"]) + n] == expect[int(a["offset"]):int(a["offset"]) + n]).all(), (f, c)


def test_damage_and_stream_limits():
    rng = np.random.default_rng(77)
    packets, frames, subs, samples, _ = _stream(rng, 60, 576, 16, 2, seed=901)
    hit = []
    for k, p in enumerate(packets):
        b = bytearray(p)
        kind = k % 6
        if kind == 1:
            b[int(rng.integers(2, 6))] ^= 1 << int(rng.integers(8))        # header byte: the CRC-8 catches it
        elif kind == 2:
            b[int(rng.integers(8, len(b)))] ^= 1 << int(rng.integers(8))   # sub-frame data: decodes to something, or runs out of bits
        elif kind == 3:
            b = b[:int(rng.integers(6, len(b)))]                            # cut
        elif kind == 4:
            b = bytearray(rng.integers(0, 255, 5, dtype=np.uint8).tobytes()) + b  # junk before the sync code
        hit.append(bytes(b))
    accepted, _ = _check_against_oracle(hit, stream_bps=16)
    assert 25 < len(accepted) < 60
    # what the stream information block allows
    assert _check_against_oracle(packets[:6], stream_bps=16, stream_channels=1)[0] == []
    assert _check_against_oracle(packets[:6], stream_bps=16, stream_channels=2, max_block=100)[0] == []  # (the stream's first block is a short one: 193 samples)
    assert len(_check_against_oracle(packets[:6], stream_bps=16, stream_channels=2, max_block=576)[0]) == 6
    no_bps = [p for p in packets if (p[3] >> 1) & 7 == 0]
    assert no_bps and _check_against_oracle(no_bps, stream_bps=0)[0] == []  # bits per sample neither in the frame nor in the stream
    # reserved header values, each with a valid CRC-8
    base = bytearray(packets[0])

    hdr_len = next(e for e in range(5, 17) if fw.crc8(bytes(base[:e])) == base[e])
    cases = []
    for i, value in ((2, base[2] & 0x0F), (2, base[2] | 0x0F), (3, (base[3] & 0x0F) | 0xB0), (3, (base[3] & 0xF1) | 0x06), (3, base[3] | 0x01), (4, 0xFF), (4, 0x80)):
        b = bytearray(base)
        b[i] = value
        b[hdr_len] = fw.crc8(bytes(b[:hdr_len]))
        cases.append(bytes(b))
    assert _check_against_oracle(cases, stream_bps=16)[0] == []


# ------------------------------------------------------------------------------------------- native container -> PCM

def test_native_file_to_pcm(oracle):
    """A .flac file end to end on the CPU side of the boundary: marker + STREAMINFO + metadata, frames split by checksum,
    front-end, restoration oracle -- back to the PCM the encoder started from."""
    import symphonia_b200 as sb
    from symphonia_b200 import packetizer
    rng = np.random.default_rng(88)
    for bps, channels, block, fixed in ((16, 2, 1152, True), (24, 2, 4096, True), (16, 1, 576, True)):
        n_frames = 14
        frames, subs, samples, expect = workloads.flac_batch(n_frames, block, seed=300 + bps + block, bps=bps, channels=channels, return_pcm=True)
        # fixed-block-size stream: every block but the last has the stream's size (flac_batch makes every 7th short; keep the shape
        # a real encoder produces by re-encoding those positions at full size would change the data -- use a stream whose only short
        # block is the last one instead)
        order = [f for f in range(n_frames) if f % 7] + [0]
        pk = []
        for number, f in enumerate(order):
            fr = frames[f]
            ss = subs[int(fr["first_subframe"]):int(fr["first_subframe"]) + channels]
            pk.append(fw.write_frame(rng, fr, ss, samples, number, stream_bps=bps))
        total = sum(int(subs[int(frames[f]["first_subframe"])]["n"]) for f in order)
        info_block = fw.stream_info_block(block, block, 44100, channels, bps, total, min(map(len, pk)), max(map(len, pk)), md5=bytes(range(1, 17)))
        data = fw.native_file(pk, info_block, extra_blocks=[(4, b"\x07\x00\x00\x00example\x00\x00\x00\x00"), (1, bytes(300))])
        info, packets = packetizer.flac_index(data)
        assert (int(info["sample_rate"]), int(info["channels"]), int(info["bits_per_sample"]), int(info["block_max"]), int(info["n_samples"])) == \
            (44100, channels, bps, block, total) and bool(info["has_md5"])
        at = int(info["first_frame_pos"])
        assert data[at:at + 2] in (b"\xff\xf8", b"\xff\xf9")
        # the packets are the frames as written, time stamps count samples
        assert [int(p["size"]) for p in packets] == [len(x) for x in pk]
        assert [int(p["ts"]) for p in packets] == [k * block for k in range(len(pk))]
        assert int(packets[-1]["ts"]) + int(packets[-1]["dur"]) == total
        table = np.zeros(len(packets), dtype=nat.PIECE_DTYPE)
        table["offset"], table["len"] = packets["offset"], packets["size"]
        gf, gi, gof, gs, gsm = frontend.flac_decode_packets(data, table, int(info["bits_per_sample"]), int(info["channels"]), int(info["block_max"]))
        assert len(gf) == len(pk)
        rc, pcm = kat._restore(oracle, gf, gs, gsm.copy())
        assert rc == 0
        for k, f in enumerate(order):
            for c in range(channels):
                a, b = subs[f * channels + c], gs[int(gf[k]["first_subframe"]) + c]
                n = int(a["n"])
                assert (pcm[int(b["offset"]):int(b["offset"]) + n] == expect[int(a["offset"]):int(a["offset"]) + n]).all(), (k, c)
        # damage: a flipped payload bit fails that frame's CRC-16 and only that frame goes
        hurt = bytearray(data)
        victim = packets[5]
        hurt[int(victim["offset"]) + int(victim["size"]) // 2] ^= 0x20
        _, p2 = packetizer.flac_index(bytes(hurt))
        assert [int(p["ts"]) for p in p2] == [int(p["ts"]) for k, p in enumerate(packets) if k != 5]
        cut = int(packets[8]["offset"])
        _, p3 = packetizer.flac_index(data[:cut] + rng.integers(0, 256, 333, dtype=np.uint8).tobytes() + data[cut:])
        # (the frame in FRONT of the junk goes with it: nothing behind its checksum looks like a frame, so nothing vouches for its end --
        #  the reference's fragment parser cuts at sync codes too and loses it the same way)
        assert [int(p["ts"]) for p in p3] == [int(p["ts"]) for k, p in enumerate(packets) if k != 7]
        _, p4 = packetizer.flac_index(data[:-7])  # a cut last frame has no checksum to vouch for it
        assert len(p4) == len(packets) - 1
    # container-level refusals
    with pytest.raises(sb.SymgpuError) as e:
        packetizer.flac_index(b"OggS" + bytes(100))
    assert e.value.status == 2
    for bad_block in (fw.stream_info_block(8, 4096, 44100, 2, 16, 0), fw.stream_info_block(4096, 1024, 44100, 2, 16, 0),
                      fw.stream_info_block(4096, 4096, 0, 2, 16, 0), fw.stream_info_block(4096, 4096, 44100, 2, 3, 0),
                      fw.stream_info_block(4096, 4096, 44100, 2, 16, 0, frame_min=900, frame_max=100)):
        with pytest.raises(sb.SymgpuError) as e:
            packetizer.flac_index(fw.native_file([], bad_block))
        assert e.value.status == 1
    with pytest.raises(sb.SymgpuError):
        packetizer.flac_index(b"fLaC" + bytes([0x84]) + (10).to_bytes(3, "big") + bytes(10))  # first block is not STREAMINFO
    with pytest.raises(sb.SymgpuError):
        packetizer.flac_index(b"fLaC" + bytes([0x00]) + (34).to_bytes(3, "big") + fw.stream_info_block(4096, 4096, 44100, 2, 16, 0))  # metadata never ends


def test_lpc_precision_code_1111_is_refused():
    """decoder.rs:478-482: a quantised-coefficient precision of 16 bits (code 1111) is a reserved value."""
    rng = np.random.default_rng(99)
    frames, subs, samples = workloads.flac_batch(40, 576, seed=41, bps=16, channels=2)
    lpc = [f for f in range(40) if any(int(subs[f * 2 + c]["type"]) == 3 for c in range(2))]
    assert len(lpc) > 10
    ok = [fw.write_frame(rng, frames[f], subs[f * 2:f * 2 + 2], samples, k, stream_bps=16, lpc_precision=15) for k, f in enumerate(lpc)]
    bad = [fw.write_frame(rng, frames[f], subs[f * 2:f * 2 + 2], samples, k, stream_bps=16, lpc_precision=16) for k, f in enumerate(lpc)]
    assert _check_against_oracle(ok, stream_bps=16)[0] == list(range(len(lpc)))
    assert _check_against_oracle(bad, stream_bps=16)[0] == []


def test_partition_smaller_than_the_predictor_order_is_refused():
    """decoder.rs:548-557: the first partition holds (block >> order) - predictor order residuals; a negative count is an error,
    as is a partition order that does not divide the block."""
    rng = np.random.default_rng(98)
    frames, subs, samples = workloads.flac_batch(30, 64, seed=43, bps=16, channels=1)
    high = [f for f in range(30) if int(subs[f]["type"]) == 3 and int(subs[f]["order"]) > 4 and int(subs[f]["n"]) == 64]
    assert len(high) >= 3
    ok = [fw.write_frame(rng, frames[f], subs[f:f + 1], samples, k, stream_bps=16, force_po=2) for k, f in enumerate(high)]   # 16 per partition
    bad = [fw.write_frame(rng, frames[f], subs[f:f + 1], samples, k, stream_bps=16, force_po=4) for k, f in enumerate(high)]  # 4 per partition < order
    odd = [fw.write_frame(rng, frames[f], subs[f:f + 1], samples, k, stream_bps=16, force_po=7) for k, f in enumerate(high)]  # 64 >> 7 = 0
    good_orders = [f for f in high if int(subs[f]["order"]) <= 16]
    assert _check_against_oracle(ok, stream_bps=16)[0] == [k for k, f in enumerate(high) if f in good_orders]
    assert _check_against_oracle(bad, stream_bps=16)[0] == []
    assert _check_against_oracle(odd, stream_bps=16)[0] == []


def _flac_file(seed, bps, channels, block, n_frames=10):
    rng = np.random.default_rng(seed)
    frames, subs, samples, expect = workloads.flac_batch(n_frames, block, seed=seed, bps=bps, channels=channels, return_pcm=True)
    order = [f for f in range(n_frames) if f % 7] + [0]
    pk = [fw.write_frame(rng, frames[f], subs[int(frames[f]["first_subframe"]):int(frames[f]["first_subframe"]) + channels], samples, k, stream_bps=bps)
          for k, f in enumerate(order)]
    total = sum(int(subs[int(frames[f]["first_subframe"])]["n"]) for f in order)
    data = fw.native_file(pk, fw.stream_info_block(block, block, 44100, channels, bps, total, min(map(len, pk)), max(map(len, pk))))
    want = np.zeros((total, channels), dtype=np.int32)
    at = 0
    for f in order:
        n = int(subs[f * channels]["n"])
        for c in range(channels):
            a = subs[f * channels + c]
            want[at:at + n, c] = expect[int(a["offset"]):int(a["offset"]) + n]
        at += n
    return data, want


def test_one_call_flac_plan(oracle):
    """CPU half of decode.decode_flac: plan + restoration oracle + interleave = the PCM the encoder started from (scaled to 32 bits)."""
    from symphonia_b200 import decode
    for seed, (bps, channels, block) in enumerate(((16, 2, 1152), (24, 2, 4096), (16, 1, 576), (20, 2, 256))):
        data, want = _flac_file(700 + seed, bps, channels, block)
        plan = decode.flac_plan(data)
        assert (plan["channels"], plan["bits_per_sample"], plan["sample_rate"], plan["total_frames"]) == (channels, bps, 44100, len(want))
        rc, restored = kat._restore(oracle, plan["frames"], plan["subframes"], plan["samples"].copy())
        assert rc == 0
        got = decode.flac_interleave(plan, restored)
        assert got.dtype == np.int32 and (got == want).all()


@pytest.mark.gpu
def test_one_call_flac_on_the_device(oracle):
    import symphonia_b200 as sb
    from symphonia_b200 import decode
    with sb.Engine(0) as eng:
        for seed, (bps, channels, block) in enumerate(((16, 2, 1152), (24, 2, 4096), (16, 1, 576), (20, 2, 256))):
            data, want = _flac_file(700 + seed, bps, channels, block)
            got, rate = decode.decode_flac(eng, data)
            assert rate == 44100 and (got == want).all()
