"""MP3 entropy front-end (`symgpu_mp3_fe_*`) against oracle/mp3_frontend_oracle.py and against the
ground truth of an independent bitstream writer (tests/_mp3_bitstream.py).  CPU only: the front-end's output is
the synthesis kernels' input (`symgpu_mp3_gc` + int16 quantised spectra), integers throughout, so the bar is equality."""
import numpy as np
import pytest

import symphonia_b200 as sb
from oracle import mp3_frontend_oracle as fo
from oracle import packetizer_oracle as po
from symphonia_b200 import _native as nat
from symphonia_b200 import frontend, packetizer
from tests import _mp3_bitstream as bw
from tests import _streams as st

BLOCK = {"long": 0, "start": 1, "short": 2, "end": 3}


def _expected_unit(h, c):
    """GranuleChannel -> symgpu_mp3_gc, written from the header file's field list (include/symgpu.h)."""
    flags = (nat.F_MIXED if c["mixed"] else 0) | (nat.F_SCALEFAC_SCALE if c["scalefac_scale"] else 0) | (nat.F_PREFLAG if c["preflag"] else 0) | \
        (nat.F_SFC_LSB if c["scalefac_compress"] & 1 else 0) | (nat.F_MPEG1 if h["version"] == "1" else 0) | \
        (nat.F_MID_SIDE if h["mode"] == "joint" and h["mid_side"] else 0) | (nat.F_INTENSITY if h["mode"] == "joint" and h["intensity"] else 0)
    return dict(rzero=c["rzero"], global_gain=c["global_gain"], block_type=BLOCK[c["block_type"]], flags=flags, sample_rate_idx=h["sample_rate_idx"],
                subblock_gain=list(c["subblock_gain"]), scalefacs=list(c["scalefacs"]))


def _compare(units, quant, info, h, granules, oquant, under, used, where):
    n_gr = 2 if h["version"] == "1" else 1
    assert (int(info["channels"]), int(info["granules"]), int(info["sample_rate"]), int(info["sample_rate_idx"])) == \
        (h["n_channels"], n_gr, h["sample_rate"], h["sample_rate_idx"]), where
    assert (int(info["underflow_bytes"]), int(info["main_data_bytes"])) == (under, used), where
    for gr in range(2):
        for ch in range(2):
            u = units[gr, ch]
            if gr >= n_gr or ch >= h["n_channels"]:
                assert int(u["flags"]) & nat.F_MUTE and not quant[gr, ch].any() and int(u["rzero"]) == 0, where
                continue
            want = _expected_unit(h, granules[gr][ch])
            got = dict(rzero=int(u["rzero"]), global_gain=int(u["global_gain"]), block_type=int(u["block_type"]), flags=int(u["flags"]),
                       sample_rate_idx=int(u["sample_rate_idx"]), subblock_gain=[int(x) for x in u["subblock_gain"]], scalefacs=[int(x) for x in u["scalefacs"]])
            assert got == want, (where, gr, ch)
            assert not u["reserved"].any()
            assert quant[gr, ch].tolist() == oquant[gr][ch], (where, gr, ch)
            assert not quant[gr, ch, want["rzero"]:].any()


def _run_both(frames, where=""):
    """Feeds the same packets to the C++ front-end and the oracle; statuses and outputs must agree frame by frame."""
    ofe, cfe = fo.Mp3Frontend(), frontend.Mp3Frontend()
    outcomes = []
    for k, f in enumerate(frames):
        try:
            h, granules, oquant, under, used = ofe.decode(f)
            ok = True
        except fo.DecodeError:
            ok = False
        try:
            units, quant, info = cfe.decode(f)
            assert ok, f"{where} frame {k}: accepted, the oracle refuses it"
            _compare(units, quant, info, h, granules, oquant, under, used, f"{where} frame {k}")
        except sb.SymgpuError as e:
            assert not ok, f"{where} frame {k}: refused ({e}), the oracle accepts it"
            assert e.status in (1, 2)
        outcomes.append(ok)
    cfe.close()
    return outcomes


# ------------------------------------------------------------------------------------------- pins

def test_oracle_decodes_a_frame_assembled_by_hand():
    # MPEG-1 Layer III, 44.1 kHz, 32 kbit/s (104-byte frame), mono, no CRC: 17 bytes of side information.
    # Granule 0: scalefac_compress 1 (slen 0 / 1: ten 1-bit scale factors for bands 11..20), three regions all using
    # Huffman table 1 -- ISO 11172-3 Table B.7: (0,0) '1', (0,1) '001', (1,0) '01', (1,1) '000' -- big_values = 3
    # coding (1,0) (0,1) (1,1) with signs + - - +, then two count1 quads from table A ('1' = 0000, '0101' = 0001):
    # 0001 with sign -, 0000.  Granule 1 is empty (part2_3_length 0).
    w = bw.BitWriterMsb()
    scalefac_bits = [1, 0, 1, 1, 0, 0, 1, 0, 1, 1]
    for b in scalefac_bits:
        w.put(b, 1)
    for code, width in (("01", 2), ("0", 1), ("001", 3), ("1", 1), ("000", 3), ("1", 1), ("0", 1), ("0101", 4), ("1", 1), ("1", 1)):
        w.put(int(code, 2), width)
    part2_3 = w.n
    assert part2_3 == 10 + 3 + 4 + 5 + 5 + 1
    s = bw.BitWriterMsb()
    s.put(0, 9), s.put(0, 5), s.put(0, 4)                       # main_data_begin, private, scfsi
    s.put(part2_3, 12), s.put(3, 9), s.put(150, 8), s.put(1, 4), s.put(0, 1)  # granule 0: lengths, gain, compress, no switching
    s.put(1, 5), s.put(1, 5), s.put(1, 5), s.put(0, 4), s.put(0, 3)          # three table selects, region counts
    s.put(1, 1), s.put(0, 1), s.put(0, 1)                                    # preflag, scalefac_scale, count1 table A
    s.put(0, 12), s.put(0, 9), s.put(0, 8), s.put(0, 4), s.put(0, 1), s.put(0, 15), s.put(0, 4), s.put(0, 3), s.put(0, 3)  # granule 1
    side = s.bytes()
    assert len(side) == 17
    frame = (0xFFFB10C4).to_bytes(4, "big") + side + w.bytes()
    frame += bytes(104 - len(frame))
    h, granules, quant, under, used = fo.Mp3Frontend().decode(frame)
    assert (h["bitrate"], h["sample_rate"], h["n_channels"], 4 + h["frame_size"]) == (32000, 44100, 1, 104)
    c = granules[0][0]
    assert c["scalefacs"] == [0] * 11 + scalefac_bits + [0] * 18 and c["preflag"] and c["global_gain"] == 150
    assert (c["region1_start"], c["region2_start"]) == (4, 8)  # one band each at 44.1 kHz
    assert quant[0][0][:14] == [1, 0, 0, -1, -1, 1, 0, 0, 0, -1, 0, 0, 0, 0] and c["rzero"] == 14 and not any(quant[0][0][14:])
    assert granules[1][0]["rzero"] == 0 and not any(quant[1][0]) and (under, used) == (0, (part2_3 + 7) // 8)
    # and the C++ front-end on the same frame
    units, q, info = frontend.Mp3Frontend().decode(frame)
    _compare(units, q, info, h, granules, quant, under, used, "hand frame")


def test_huffman_tables_are_the_standards():
    # spot values of ISO 11172-3 Table B.7 that any transcription error would break, plus structural facts
    H = bw.HUFF
    assert [format(c, f"0{l}b") for c, l in zip(H["1"]["codes"], H["1"]["lens"])] == ["1", "001", "01", "000"]
    assert [format(c, f"0{l}b") for c, l in zip(H["quadA"]["codes"], H["quadA"]["lens"])][:4] == ["1", "0101", "0100", "00101"]
    assert all(l == 4 for l in H["quadB"]["lens"]) and H["quadB"]["codes"] == list(range(15, -1, -1))
    assert {k: len(v["codes"]) for k, v in H.items()} == {"1": 4, "2": 9, "3": 9, "5": 16, "6": 16, "7": 36, "8": 36, "9": 36, "10": 64, "11": 64,
                                                         "12": 64, "13": 256, "15": 256, "16": 256, "24": 256, "quadA": 16, "quadB": 16}
    assert max(H["13"]["lens"]) == 19 and max(H["16"]["lens"]) == 17 and max(H["24"]["lens"]) == 12 and max(H["15"]["lens"]) == 13
    for t in H.values():  # complete prefix codes
        assert sum(2.0 ** -l for l in t["lens"]) == 1.0


# ------------------------------------------------------------------------------------------- round trips

@pytest.mark.parametrize("version,mode,bitrate_idx,rate_idx,protected", [
    ("1", 1, 9, 0, False), ("1", 0, 14, 1, True), ("1", 3, 5, 2, False), ("1", 2, 11, 0, False), ("2", 1, 8, 0, False), ("2", 3, 3, 1, True),
    ("2", 0, 14, 2, False), ("2.5", 1, 6, 0, False), ("2.5", 3, 1, 2, False), ("2.5", 0, 10, 1, True), ("1", 1, 1, 0, False)])
def test_written_streams_come_back(version, mode, bitrate_idx, rate_idx, protected):
    rng = np.random.default_rng(100 + bitrate_idx + 7 * rate_idx + (mode << 4))
    for rich in (True, False):
        frames, truth = bw.gen_stream(rng, 60, version=version, mode=mode, rate_idx=rate_idx, bitrate_idx=bitrate_idx, protected=protected, rich=rich)
        assert all(_run_both(frames, f"{version}/{mode}/{bitrate_idx}"))
        # what the writer put in is what comes out (the oracle agreeing with the C++ is not enough: both could misread)
        cfe = frontend.Mp3Frontend()
        seen_types, reservoir_used = set(), 0
        for k, (f, t) in enumerate(zip(frames, truth)):
            units, quant, info = cfe.decode(f)
            reservoir_used += t["main_data_begin"] > 0
            for gr in range(t["n_gr"]):
                for ch in range(t["n_ch"]):
                    g, u = t["granules"][gr][ch], units[gr, ch]
                    assert quant[gr, ch].tolist() == g["quant"] and int(u["rzero"]) == g["rzero"], (k, gr, ch)
                    assert [int(x) for x in u["scalefacs"]] == g["scalefacs"] and int(u["global_gain"]) == g["global_gain"]
                    assert int(u["block_type"]) == g["block_type"] and bool(int(u["flags"]) & nat.F_MIXED) == g["mixed"]
                    assert bool(int(u["flags"]) & nat.F_PREFLAG) == bool(g["preflag"]) and [int(x) for x in u["subblock_gain"]] == g["subblock_gain"]
                    seen_types.add((g["block_type"], g["mixed"]))
        assert reservoir_used > 20 and len(seen_types) >= 4
        if rich:
            assert max(abs(v) for t in truth for row in t["granules"] for g in row for v in g["quant"]) > 1000  # linbits in play


def test_values_at_the_limits():
    # the largest codable magnitude is 15 + 2^13 - 1 = 8206 (tables 23 / 31): every pair saturated, both signs
    rng = np.random.default_rng(7)

    class Fixed:  # a stand-in generator: maximal draws
        def __init__(self, r):
            self.r = r

        def integers(self, *a, **k):
            return self.r.integers(*a, **k)

        def __getattr__(self, name):
            return getattr(self.r, name)
    frames, truth = bw.gen_stream(Fixed(rng), 20, version="1", mode=0, bitrate_idx=14, rate_idx=1)
    peak = max(abs(v) for t in truth for row in t["granules"] for g in row for v in g["quant"])
    assert peak == 8206 or peak > 4000
    assert all(_run_both(frames, "limits"))


# ------------------------------------------------------------------------------------------- reservoir and damage

def test_joining_a_stream_in_the_middle():
    rng = np.random.default_rng(21)
    frames, truth = bw.gen_stream(rng, 40, version="1", mode=1, bitrate_idx=7, fill=(0.6, 1.0))
    for start in (1, 2, 5, 11):
        assert truth[start]["main_data_begin"] > 0
        outcomes = _run_both(frames[start:], f"join at {start}")
        assert all(outcomes)  # underflow is not an error: the granules whose bits are missing are silent
    cfe = frontend.Mp3Frontend()
    _, quant, info = cfe.decode(frames[5])
    assert int(info["underflow_bytes"]) == truth[5]["main_data_begin"] and not quant[0].any()
    # a frame lost in the middle: the next one reaches back into bytes that were never seen
    outcomes = _run_both(frames[:10] + frames[11:], "frame 10 lost")
    assert all(outcomes)
    # reset forgets the reservoir
    cfe = frontend.Mp3Frontend()
    for f in frames[:6]:
        cfe.decode(f)
    cfe.reset()
    _, _, info = cfe.decode(frames[6])
    assert int(info["underflow_bytes"]) == truth[6]["main_data_begin"] > 0


def test_damaged_frames_agree_with_the_oracle():
    rng = np.random.default_rng(22)
    refused = accepted = 0
    for version, mode, br in (("1", 1, 9), ("2", 1, 8), ("2.5", 3, 5), ("1", 3, 6)):
        frames, _ = bw.gen_stream(rng, 150, version=version, mode=mode, bitrate_idx=br)
        hit = []
        for k, f in enumerate(frames):
            b = bytearray(f)
            kind = int(rng.integers(6))
            if kind == 0:    # side information: lengths, table selects, block types
                at = 4 + int(rng.integers(0, 17 if (mode == 3 and version == "1") or (mode != 3 and version != "1") else 9 if mode == 3 else 32))
                b[at] ^= 1 << int(rng.integers(8))
            elif kind == 1:  # main data
                for _ in range(int(rng.integers(1, 6))):
                    b[int(rng.integers(40, len(b)))] ^= 1 << int(rng.integers(8))
            elif kind == 2:  # a burst of zeros / ones
                at = int(rng.integers(4, len(b) - 8))
                b[at:at + 8] = bytes([int(rng.choice([0, 255]))]) * 8
            elif kind == 3:  # side information fields forced high: big_values > 288, part2_3_length past the reservoir's end
                for _ in range(2):
                    b[4 + int(rng.integers(1, 9))] |= int(rng.choice([0xFF, 0xF0, 0x3F]))
            hit.append(bytes(b))
        outcomes = _run_both(hit, f"damaged {version}")
        refused += outcomes.count(False)
        accepted += outcomes.count(True)
    assert refused > 20 and accepted > 200  # both paths well exercised


def test_named_malformations():
    rng = np.random.default_rng(23)
    frames, truth = bw.gen_stream(rng, 12, version="1", mode=0, bitrate_idx=9, rich=False, fill=(0.5, 0.8))

    def rebuilt(k, edit):
        t = truth[k]
        granules = [[dict(g) for g in row] for row in t["granules"]]
        edit(granules)
        side = bw.side_info_bytes("1", 2, t["main_data_begin"], t["scfsi"], granules)
        return frames[k][:4] + side + frames[k][4 + 32:]

    def case(k, edit, expect_ok):
        seq = frames[:k] + [rebuilt(k, edit)] + frames[k + 1:]
        outcomes = _run_both(seq, "malformed")
        assert outcomes[k] == expect_ok
        return outcomes

    case(3, lambda g: g[0][0].update(big_values=289), False)
    case(3, lambda g: g[1][1].update(window_switching=1, block_type=0, mixed_bit=0, table_select=[1, 1, 0], subblock_gain=[0, 0, 0]), False)
    case(3, lambda g: g[0][0].update(part2_3_length=1, scalefac_compress=15), False)     # part 2 alone is longer
    # a part2_3_length that stops inside the last count1 quad: the quad is undone, nothing else changes
    k = next(i for i, t in enumerate(truth) if t["granules"][1][1]["rzero"] > 2 * t["granules"][1][1]["big_values"] and not t["granules"][1][1]["stuffing"])
    outcomes = case(k, lambda g: g[1][1].update(part2_3_length=g[1][1]["part2_3_length"] - 1), True)
    assert all(outcomes)
    cfe = frontend.Mp3Frontend()
    for f in frames[:k]:
        cfe.decode(f)
    units, _, _ = cfe.decode(rebuilt(k, lambda g: g[1][1].update(part2_3_length=g[1][1]["part2_3_length"] - 1)))
    assert int(units[1, 1]["rzero"]) == truth[k]["granules"][1][1]["rzero"] - 4
    # a refused frame empties the reservoir: the next frame underflows instead of reading stale bytes
    seq = frames[:4] + [rebuilt(4, lambda g: g[0][0].update(big_values=300))] + frames[5:]
    cfe = frontend.Mp3Frontend()
    for f in seq[:4]:
        cfe.decode(f)
    with pytest.raises(sb.SymgpuError):
        cfe.decode(seq[4])
    _, _, info = cfe.decode(seq[5])
    assert int(info["underflow_bytes"]) == truth[5]["main_data_begin"]


def test_packet_level_checks():
    rng = np.random.default_rng(24)
    frames, _ = bw.gen_stream(rng, 6, version="1", mode=1, bitrate_idx=9)
    other_rate, _ = bw.gen_stream(rng, 2, version="1", mode=1, bitrate_idx=9, rate_idx=1)
    mono, _ = bw.gen_stream(rng, 2, version="1", mode=3, bitrate_idx=9)
    layer2 = st.mpa_frame(rng, dict(version="1", layer=2, bitrate_idx=9, rate_idx=0, mode=0))
    seq = [frames[0], frames[1][:-1], frames[1] + b"\0", b"\xff\xfb", b"", frames[1], other_rate[0], mono[0], layer2, frames[2],
           (0xFFFB0044).to_bytes(4, "big") + bytes(400), b"\x00\x01" + frames[3], frames[3]]
    outcomes = _run_both(seq, "packets")
    # (leading junk inside a packet is skipped by the sync search, decoder.rs:87: the frame behind it is taken)
    assert outcomes == [True, False, False, False, False, True, False, False, False, True, False, True, True]
    cfe = frontend.Mp3Frontend()
    with pytest.raises(sb.SymgpuError) as e:
        cfe.decode((0xFFFB0044).to_bytes(4, "big") + bytes(400))  # free format
    assert e.value.status == 2
    # a decoder that has seen nothing yet takes its signal specification from the first packet it is shown, good or bad
    cfe = frontend.Mp3Frontend()
    with pytest.raises(sb.SymgpuError):
        cfe.decode(layer2)                      # 44.1 kHz stereo is now the spec, although the packet was refused
    cfe.decode(frames[0])
    with pytest.raises(sb.SymgpuError):
        cfe.decode(mono[0])


# ------------------------------------------------------------------------------------------- file -> batch

def test_file_to_batch_through_packetiser_and_front_end():
    rng = np.random.default_rng(25)
    params = dict(version="1", layer=3, bitrate_idx=9, rate_idx=0, mode=1)
    frames, truth = bw.gen_stream(rng, 50, version="1", mode=1, bitrate_idx=9)
    broken = bytearray(frames[20])
    broken[8] = 0xFF  # big_values of granule 0 channel 0 (side-information bits 32..40) = 511: out of range
    broken[9] |= 0x80
    tag = st.mpa_tag_frame(rng, params, num_frames=50)
    noise = rng.integers(0, 255, 300, dtype=np.uint8).tobytes()
    data = noise + tag + b"".join(frames[:20]) + bytes(broken) + b"".join(frames[21:35]) + noise[:60] + b"".join(frames[35:])
    track, packets = packetizer.mpa_index(data)
    assert len(packets) == 50 and int(track["delay"]) == 1105
    fe = frontend.Mp3Frontend()
    units, quant, frame_of, info = fe.decode_packets(data, packets)
    assert frame_of.tolist() == [k for k in range(50) if k != 20] and (int(info["channels"]), int(info["granules"])) == (2, 2)
    # every good frame equals the writer's truth, except the few after the refused frame whose main data reached back
    # into the reservoir that was emptied (at most 511 bytes, i.e. two frames at this rate)
    for row, k in enumerate(frame_of.tolist()):
        t = truth[k]
        if 20 < k < 24:
            continue
        for gr in range(2):
            for ch in range(2):
                assert quant[row, gr, ch].tolist() == t["granules"][gr][ch]["quant"], (k, gr, ch)
    # the batch is what the synthesis entry points take: a device-free check of every unit (when block types allow)
    runs = np.zeros(1, dtype=nat.MP3_RUN_DTYPE)
    runs[0] = (0, 0, len(frame_of), 2, 2, 0)
    rc = nat.lib().symgpu_mp3_units_check(units.ctypes.data, runs.ctypes.data, 1, len(frame_of))
    assert rc in (0, 1)  # 1: a joint-stereo pair with unequal block types, which stereo.rs:503-505 refuses as well
    assert po.mpa_index(data)[1][0][0] == int(packets[0]["offset"])


# ------------------------------------------------------------------------------------------- plan + independent jobs

def _packets_of(frames):
    """A packet table over the concatenated frames (what the packetiser would return for a clean file)."""
    p = np.zeros(len(frames), dtype=nat.MPA_PACKET_DTYPE)
    at = 0
    for k, f in enumerate(frames):
        p[k]["offset"], p[k]["size"] = at, len(f)
        at += len(f)
    return b"".join(frames), p


def _same_as_serial(frames, where):
    data, packets = _packets_of(frames)
    su, sq, sf, sinfo = frontend.Mp3Frontend().decode_packets(data, packets)
    pu, pq, pf, pinfo, rounds = frontend.entropy_decode_cpu(data, packets)
    assert pf.tolist() == sf.tolist(), where
    assert pu.tobytes() == su.tobytes() and (pq == sq).all(), where
    if len(sf):
        assert pinfo.tobytes() == sinfo.tobytes(), where
    return rounds, len(sf)


def test_planned_jobs_equal_the_serial_front_end():
    rng = np.random.default_rng(31)
    for version, mode, br, prot in (("1", 1, 9, False), ("1", 3, 5, True), ("1", 0, 14, False), ("2", 1, 8, False), ("2", 3, 3, False),
                                    ("2.5", 1, 6, True), ("2.5", 3, 1, False)):
        frames, truth = bw.gen_stream(rng, 80, version=version, mode=mode, bitrate_idx=br, protected=prot)
        rounds, good = _same_as_serial(frames, f"{version}/{mode}")
        assert (rounds, good) == (1, 80)
        _same_as_serial(frames[7:], "joined late")            # underflow: silent granules, partial first granule
        _same_as_serial(frames[:30] + frames[31:], "a frame lost")
    # the jobs really are independent: run them in a shuffled order, one at a time
    frames, _ = bw.gen_stream(rng, 40, version="1", mode=1, bitrate_idx=9)
    data, packets = _packets_of(frames)
    md, jobs, frame_of, _ = frontend.entropy_plan(data, packets)
    assert len(jobs) == 160 and md.size == sum(len(f) - 4 - 32 for f in frames)   # main data only: headers and side info stay behind
    want_u, want_q, failed = frontend.entropy_run_cpu(md, jobs)
    assert not failed.any()
    got_u, got_q = np.zeros_like(want_u), np.zeros_like(want_q)
    for k in rng.permutation(len(jobs)):
        frame = k // 4
        u, q, f = frontend.entropy_run_cpu(md, jobs[frame * 4:frame * 4 + 4][[k % 4]])  # a single job; its slot is relative to its own frame
        got_u[frame].reshape(-1)[k % 4] = u[0].reshape(-1)[k % 4]
        got_q[frame].reshape(4, 576)[k % 4] = q[0].reshape(4, 576)[k % 4]
    assert got_u.tobytes() == want_u.tobytes() and (got_q == want_q).all()


def test_planned_jobs_on_damaged_streams():
    rng = np.random.default_rng(32)
    replanned = 0
    for version, mode, br in (("1", 1, 9), ("2", 1, 8), ("1", 3, 6), ("2.5", 3, 5)):
        frames, _ = bw.gen_stream(rng, 120, version=version, mode=mode, bitrate_idx=br)
        hit = []
        for f in frames:
            b = bytearray(f)
            kind = int(rng.integers(8))
            if kind == 0:
                b[4 + int(rng.integers(1, 9))] |= int(rng.choice([0xFF, 0xF0, 0x3F]))   # side information forced high
            elif kind == 1:
                b[4 + int(rng.integers(0, 9))] ^= 1 << int(rng.integers(8))
            elif kind == 2:
                for _ in range(3):
                    b[int(rng.integers(40, len(b)))] ^= 1 << int(rng.integers(8))
            hit.append(bytes(b))
        rounds, good = _same_as_serial(hit, f"damaged {version}")
        replanned += rounds > 1
        assert 60 < good <= 120
    assert replanned >= 2  # decode-time failures occurred and the re-plan reproduced the emptied reservoir


def test_the_device_bit_window_on_the_host(tmp_path):
    """The kernel composes its 32-bit window from five byte loads; the host normally takes one 8-byte load instead.  Build the
    front-end once more with the device's path forced (SYMGPU_MP3E_DEVICE_WINDOW) and require identical output: every
    arithmetic step the device thread performs has then run on the CPU."""
    import ctypes
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libfe_devwin.so")
    csrc = os.path.join(root, "symphonia_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-DSYMGPU_MP3E_DEVICE_WINDOW", "-I/usr/local/cuda/include",
                           "-o", so, os.path.join(csrc, "mp3_frontend.cpp"), os.path.join(csrc, "packetizer.cpp"), os.path.join(csrc, "tables.cpp")])
    L = ctypes.CDLL(so)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    L.symgpu_mp3_entropy_decode_cpu.restype = ctypes.c_int
    L.symgpu_mp3_entropy_decode_cpu.argtypes = [vp, sz, vp, sz, vp, vp, vp, ctypes.POINTER(sz), vp, ctypes.POINTER(ctypes.c_uint32)]
    rng = np.random.default_rng(41)
    for version, mode, br in (("1", 1, 9), ("2", 3, 4), ("2.5", 0, 8), ("1", 0, 14)):
        frames, _ = bw.gen_stream(rng, 60, version=version, mode=mode, bitrate_idx=br)
        hit = [bytes(bytearray(f[:9]) + bytearray([f[9] ^ 0x10]) + bytearray(f[10:])) if k % 9 == 4 else f for k, f in enumerate(frames)]
        data, packets = _packets_of(hit)
        want_u, want_q, want_f, _, _ = frontend.entropy_decode_cpu(data, packets)
        a = np.frombuffer(data, dtype=np.uint8)
        n = len(packets)
        units, quant = np.zeros((n, 2, 2), dtype=nat.MP3_GC_DTYPE), np.zeros((n, 2, 2, 576), dtype=np.int16)
        frame_of, info = np.zeros(n, dtype=np.uint32), np.zeros(1, dtype=nat.MP3_FRAME_INFO_DTYPE)
        good, rounds = sz(0), ctypes.c_uint32(0)
        assert L.symgpu_mp3_entropy_decode_cpu(a.ctypes.data, a.size, packets.ctypes.data, n, units.ctypes.data, quant.ctypes.data, frame_of.ctypes.data,
                                               ctypes.byref(good), info.ctypes.data, ctypes.byref(rounds)) == 0
        g = good.value
        assert frame_of[:g].tolist() == want_f.tolist() and units[:g].tobytes() == want_u.tobytes() and (quant[:g] == want_q).all()


def test_threaded_job_executor_gives_the_same_tables():
    rng = np.random.default_rng(51)
    frames, _ = bw.gen_stream(rng, 90, version="1", mode=1, bitrate_idx=9)
    data, packets = _packets_of(frames[3:])  # joined late: silent jobs, scfsi re-reads across the thread cut points
    md, jobs, frame_of, _ = frontend.entropy_plan(data, packets)
    want = frontend.entropy_run_cpu(md, jobs)
    for threads in (2, 3, 8, 0, 500):
        got = frontend.entropy_run_cpu(md, jobs, threads=threads)
        assert got[0].tobytes() == want[0].tobytes() and (got[1] == want[1]).all() and (got[2] == want[2]).all()


def test_every_mpeg2_scalefac_compress_value():
    """All 512 values of the 9-bit field, on an ordinary channel and on the intensity channel (three partition tables each, with
    their switch points at 400 / 500 and 360 / 488), over random block types.  (A mutant moving one switch point survived the
    randomised streams.)"""
    rng = np.random.default_rng(81)
    for mode in (1, 0):  # joint stereo (channel 1 is the intensity channel when the mode extension says so), plain stereo
        frames, truth = bw.gen_stream(rng, 1024, version="2", mode=mode, bitrate_idx=12, rate_idx=0, fill=(0.2, 0.5), rich=False,
                                      force_sfc=lambda k, gr, ch: k % 512, force_mode_ext=lambda k: 1 + 2 * (k // 512))  # intensity on, mid/side off then on
        seen = {(t["granules"][0][1]["scalefac_compress"], bool(t["mode"] == 1 and t["mode_ext"] & 1)) for t in truth}
        assert seen == {(v, mode == 1) for v in range(512)}
        assert all(_run_both(frames, f"sfc sweep mode {mode}"))
        cfe = frontend.Mp3Frontend()
        for f, t in zip(frames, truth):
            units, quant, _ = cfe.decode(f)
            for ch in range(2):
                g = t["granules"][0][ch]
                assert [int(x) for x in units[0, ch]["scalefacs"]] == g["scalefacs"] and bool(int(units[0, ch]["flags"]) & nat.F_PREFLAG) == bool(g["preflag"])
