"""AAC-LC entropy front-end (`symgpu_aac_fe_*`, SURVEY §8f N1 for the AAC path) against oracle/aac_frontend_oracle.py (the reference's
sequence, numpy f32 arithmetic) and against an independent raw_data_block writer's ground truth.  Every value is a short chain of
single IEEE operations on table entries, so the bar is bit equality.  The oracle is pinned by the reference's own unit test for
this path (ics/mod.rs:612-635) and by the writer.  CPU only."""
import numpy as np
import pytest

import symphonia_b200 as sb
from oracle import aac_frontend_oracle as ao
from symphonia_b200 import _native as nat
from symphonia_b200 import frontend
from symphonia_b200.engine import SymgpuError
from tests import _aac_bitstream as ab
from tests import _oracle

RATES = [44100, 48000, 8000, 96000, 22050, 32000, 16000, 64000, 11025, 24000, 88200, 12000]


def u32(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- tables ---------------------------------------------------------------------------------------------------------------

def test_tables_equal_the_oracle_and_the_closed_forms():
    p43, normal, intensity = frontend.aac_tables()
    assert np.array_equal(u32(p43), u32(np.array(ao.POW43)))
    assert np.array_equal(u32(normal), u32(np.array(ao.NORMAL_SCF)))
    assert np.array_equal(u32(intensity), u32(np.array(ao.INTENSITY_SCF)))
    # the scale-factor tables are the correctly rounded powers of two -- whichever libm routine computes them
    assert np.array_equal(u32(normal), u32(np.array([ab.scale_normal(i) for i in range(256)])))
    assert np.array_equal(u32(intensity), u32(np.array([ab.scale_intensity(i) for i in range(256)])))
    assert float(normal[156]) == 1.0 and float(intensity[155]) == 1.0 and float(normal[160]) == 2.0 and float(intensity[159]) == 0.5
    # x^(4/3) is the C library's powf with the exponent 4/3 rounded to f32 (1.3333334), as in the reference: never more than one
    # unit in the last place from the correctly rounded value of THAT power -- so 8^(4/3) is one step above 16
    cr = np.array([ab.pow43_correctly_rounded(i) for i in range(8192)], dtype=np.float32)
    off = np.abs(p43.view(np.int32).astype(np.int64) - cr.view(np.int32))
    assert off.max() <= 1 and int((off != 0).sum()) <= 32
    assert float(p43[0]) == 0.0 and float(p43[1]) == 1.0 and float(p43[8]) == float(np.nextafter(np.float32(16), np.float32(17)))


# ---- the reference's unit test for this path ------------------------------------------------------------------------------

def test_section_data_rejects_excess_zero_length_sections():
    # ics/mod.rs decode_section_data_rejects_excess_zero_length_sections: long window, one group, max_sfb = 1, an all-zero stream
    # (codebook 0, length 0, for ever) must fail cleanly after MAX_SFBS sections
    ics = ao.Ics(ao.L48, ao.S48)
    ics.long_win, ics.window_groups, ics.max_sfb = True, 1, 1
    with pytest.raises(ao.AacError) as e:
        ics.decode_section_data(ao.BitsLtr(bytes(65 * 9 // 8 + 1)))
    assert e.value.kind == ao.DECODE
    # the same through the front-end: SCE, global gain, ics_info (long, max_sfb = 1), then zeros
    w = ab.BitWriterMsb()
    w.put(0, 3), w.put(0, 4), w.put(100, 8), w.put(0, 1), w.put(0, 2), w.put(0, 1), w.put(1, 6), w.put(0, 1)
    pkt = w.bytes() + bytes(80)
    with pytest.raises(ao.AacError):
        ao.AacFrontend(44100, 1).decode(pkt)
    fe = frontend.AacFrontend(44100, 1)
    with pytest.raises(SymgpuError) as e2:
        fe.decode(pkt)
    assert e2.value.status == 1
    fe.close()


# ---- oracle and front-end against the writer's ground truth ---------------------------------------------------------------

def _check_channel(truth, want, units, tns, coeffs, c, what):
    for key in ("window_sequence", "window_shape", "prev_window_shape"):
        assert truth is None or truth[key] == want[key], (what, key)
        assert want[key] == int(units[c][key]), (what, key)
    if truth is not None:
        assert np.array_equal(u32(truth["coeffs"]), u32(want["coeffs"])), what
        assert len(truth["tns"]) == len(want["tns"]), what
        for a, b in zip(truth["tns"], want["tns"]):
            assert tuple(a[:4]) == tuple(b[:4]) and np.array_equal(u32(np.array(a[4])), u32(np.array(b[4]))), what
    assert np.array_equal(u32(coeffs[c]), u32(want["coeffs"])), what
    assert len(want["tns"]) == int(units[c]["n_tns"]), what
    for j, b in enumerate(want["tns"]):
        r = tns[int(units[c]["tns_first"]) - what[-1] + j]
        assert (int(r["start"]), int(r["end"]), int(r["order"]), int(r["direction"])) == tuple(b[:4]), what
        assert np.array_equal(u32(r["lpc"]), u32(np.array(b[4]))), what


def _streams(base, n):
    for seed in range(n):
        rng = np.random.default_rng(base + seed)
        ch = 1 if seed % 4 == 3 else 2
        layout = ["sce", "sce"] if (seed % 7 == 5 and ch == 2) else None
        yield seed, ab.Stream(rng, rate=RATES[seed % len(RATES)], channels=ch, layout=layout), rng


def test_oracle_and_frontend_equal_the_writer_truth():
    kinds = set()
    for seed, s, _ in _streams(100, 48):
        o, fe = ao.AacFrontend(s.rate, s.channels), frontend.AacFrontend(s.rate, s.channels)
        for k in range(6):
            pkt, truth = s.packet()
            covered, want = o.decode(pkt)
            assert covered == s.channels
            base = 1000 * k
            units, tns, coeffs = fe.decode(pkt, tns_base=base)
            for c in range(s.channels):
                _check_channel(truth[c], want[c], units, tns, coeffs, c, (seed, k, c, base))
                kinds.add((truth[c]["window_sequence"], bool(truth[c]["tns"])))
            if s.channels == 1:
                assert not coeffs[1].any() and units[1].tobytes() == bytes(16)
            assert int(units["n_tns"].sum()) == len(tns)
        fe.close()
    assert len(kinds) == 8  # every window sequence with and without TNS


def test_truncated_and_damaged_packets():
    """The reference fails a packet at the first read past its end or the first rule it breaks; what it had changed by then (window
    history, noise generator) stays changed.  Both sides see the same packets in the same order, so the histories stay in step."""
    refused = accepted = 0
    for seed, s, rng in _streams(200, 24):
        o, fe = ao.AacFrontend(s.rate, s.channels), frontend.AacFrontend(s.rate, s.channels)
        for k in range(12):
            pkt, _ = s.packet()
            mode = k % 3
            if mode == 1:
                pkt = pkt[:int(rng.integers(0, len(pkt)))]
            elif mode == 2:
                b = bytearray(pkt)
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(len(b)))] ^= 1 << int(rng.integers(8))
                pkt = bytes(b)
            try:
                covered, want = o.decode(pkt)
                status = 0 if covered == s.channels else 2
            except ao.AacError as e:
                status = 1 if e.kind == ao.DECODE else 2
            if status:
                with pytest.raises(SymgpuError) as e2:
                    fe.decode(pkt)
                assert e2.value.status == status, (seed, k, mode)
                refused += 1
            else:
                units, tns, coeffs = fe.decode(pkt)
                for c in range(s.channels):
                    _check_channel(None, want[c], units, tns, coeffs, c, (seed, k, c, 0))
                accepted += 1
        fe.close()
    assert refused > 60 and accepted > 120


def test_reset_forgets_the_window_history():
    for seed, s, _ in _streams(300, 6):
        o, fe = ao.AacFrontend(s.rate, s.channels), frontend.AacFrontend(s.rate, s.channels)
        for k in range(8):
            if k == 4:
                o.reset(), fe.reset(), s.forget_windows()
            pkt, truth = s.packet()
            _, want = o.decode(pkt)
            units, tns, coeffs = fe.decode(pkt)
            for c in range(s.channels):
                _check_channel(truth[c], want[c], units, tns, coeffs, c, (seed, k, c, 0))
                if k == 4:
                    assert int(units[c]["prev_window_shape"]) == 0
        fe.close()


def test_element_layout_rules():
    rng = np.random.default_rng(5)
    stereo = ab.Stream(rng, 44100, 2)
    mono = ab.Stream(rng, 44100, 1)
    two = ab.Stream(rng, 44100, 2, layout=["sce", "sce"])
    # a channel pair in a mono stream: too many channels (mod.rs:123-124)
    for cls in (lambda: ao.AacFrontend(44100, 1), lambda: frontend.AacFrontend(44100, 1)):
        with pytest.raises((ao.AacError, SymgpuError)) as e:
            cls().decode(stereo.packet(extras=False)[0])
        assert getattr(e.value, "kind", None) == ao.DECODE or getattr(e.value, "status", None) == 1
    # one single-channel element in a stereo stream: decodes in the reference (one plane rendered); refused here as unsupported
    covered, _ = ao.AacFrontend(44100, 2).decode(mono.packet(extras=False)[0])
    assert covered == 1
    fe = frontend.AacFrontend(44100, 2)
    with pytest.raises(SymgpuError) as e:
        fe.decode(mono.packet(extras=False)[0])
    assert e.value.status == 2
    # ... after which the first element of the stream is pinned as a single channel: a pair in its place is a decode error (:117-121)
    o = ao.AacFrontend(44100, 2)
    o.decode(two.packet(extras=False)[0])
    with pytest.raises(ao.AacError):
        o.decode(stereo.packet(extras=False)[0])
    with pytest.raises(SymgpuError) as e:
        fe.decode(stereo.packet(extras=False)[0])
    assert e.value.status == 1
    fe.close()
    # coupling channel and program config elements are unsupported (:154-157, :177-180); more than two channels at create
    for eid in (2, 5):
        w = ab.BitWriterMsb()
        w.put(eid, 3), w.put(0, 13)
        fe = frontend.AacFrontend(48000, 2)
        with pytest.raises(SymgpuError) as e:
            fe.decode(w.bytes())
        assert e.value.status == 2
        with pytest.raises(ao.AacError) as e3:
            ao.AacFrontend(48000, 2).decode(w.bytes())
        assert e3.value.kind == ao.UNSUPPORTED
        fe.close()
    for ch in (0, 3, 6):
        with pytest.raises(SymgpuError) as e:
            frontend.AacFrontend(44100, ch)
        assert e.value.status == 2
    # an empty packet and a lone terminator carry no element: nothing covered
    fe = frontend.AacFrontend(44100, 2)
    for pkt in (b"", b"\xe0"):
        with pytest.raises(SymgpuError) as e:
            fe.decode(pkt)
        assert e.value.status == 2
    fe.close()


def test_predictor_and_gain_control_bits():
    # predictor_data_present in a long-window ics_info is unsupported (ltp.rs:20-54), gain control data a decode error (ics/mod.rs:438-441)
    def sce(predictor, gain):
        w = ab.BitWriterMsb()
        w.put(0, 3), w.put(0, 4), w.put(120, 8)
        w.put(0, 1), w.put(0, 2), w.put(1, 1), w.put(0, 6), w.put(predictor, 1)   # ics_info, max_sfb = 0
        w.put(0, 1), w.put(0, 1), w.put(gain, 1)                                  # pulse, tns, gain control
        w.put(7, 3)
        return w.bytes()
    fe = frontend.AacFrontend(44100, 1)
    units, tns, coeffs = fe.decode(sce(0, 0))
    assert int(units[0]["window_shape"]) == 1 and len(tns) == 0 and not coeffs.any()
    for pkt, status, kind in ((sce(1, 0), 2, ao.UNSUPPORTED), (sce(0, 1), 1, ao.DECODE)):
        with pytest.raises(SymgpuError) as e:
            fe.decode(pkt)
        assert e.value.status == status
        with pytest.raises(ao.AacError) as e2:
            ao.AacFrontend(44100, 1).decode(pkt)
        assert e2.value.kind == kind
    fe.close()


# ---- packets -> front-end -> synthesis oracle ------------------------------------------------------------------------------

def test_front_end_output_feeds_the_synthesis_stage():
    lib = _oracle.load()
    for seed, s, _ in _streams(400, 6):
        fe = frontend.AacFrontend(s.rate, s.channels)
        n = 10
        units = np.zeros((n, 2), dtype=nat.AAC_UNIT_DTYPE)
        coeffs = np.zeros((n, 2, 1024), dtype=np.float32)
        tns = []
        for k in range(n):
            u, t, c = fe.decode(s.packet()[0], tns_base=sum(len(x) for x in tns))
            units[k], coeffs[k] = u, c
            tns.append(t)
        tns = np.concatenate(tns) if tns else np.zeros(0, dtype=nat.AAC_TNS_DTYPE)
        fe.close()
        # the host entry point's descriptor check accepts what the front-end emits
        assert sb.lib().symgpu_aac_units_check(units.ctypes.data, tns.ctypes.data if len(tns) else None, len(tns), n) == 0
        runs = np.zeros(1, dtype=nat.AAC_RUN_DTYPE)
        runs[0]["n_frames"], runs[0]["channels"] = n, s.channels
        rc, pcm = _oracle.aac_batch(lib, units, tns, coeffs, runs, 1)
        assert rc == 0 and pcm.shape[0] == n
        assert np.abs(pcm[np.isfinite(pcm)]).max() > 0


def test_escape_prefix_lengths():
    """read_escape (ics/mod.rs:598-607): up to eight ones are a prefix, nine or more a decode error, ones to the end of the packet
    the end of the bitstream -- including runs longer than one machine word."""
    def packet(n_ones, tail_zero=True, word=0):
        w = ab.BitWriterMsb()
        w.put(0, 3), w.put(0, 4), w.put(140, 8)                                   # SCE, global gain
        w.put(0, 1), w.put(0, 2), w.put(0, 1), w.put(1, 6), w.put(0, 1)           # ics_info: long, sine, max_sfb 1
        w.put(11, 4), w.put(1, 5)                                                 # one section: book 11, one band
        w.huff("scf", 60)                                                         # scale factor = global gain
        w.put(0, 1), w.put(0, 1), w.put(0, 1)                                     # no pulse, no TNS, no gain control
        w.huff("11", 17 * 16 + 1), w.put(1, 1), w.put(0, 1)                       # (16, 1): signs -, +
        for _ in range(n_ones):
            w.put(1, 1)
        if tail_zero:
            w.put(0, 1)
            w.put(word, min(n_ones, 8) + 4)
            w.huff("11", 0)                                                       # the band's second pair: (0, 0)
            w.put(7, 3)
        else:
            while len(w.bits) % 8:                                                # ones all the way to the end of the packet
                w.put(1, 1)
        return w.bytes()
    for n_ones, tail, status in ((0, True, 0), (8, True, 0), (9, True, 1), (56, True, 1), (57, True, 1), (58, True, 1), (130, True, 1),
                                 (40, False, 1), (57, False, 1), (64, False, 1), (200, False, 1)):
        pkt = packet(n_ones, tail, word=5)
        fe, o = frontend.AacFrontend(44100, 1), ao.AacFrontend(44100, 1)
        try:
            _, want = o.decode(pkt)
            got_status = 0
        except ao.AacError as e:
            got_status = 1 if e.kind == ao.DECODE else 2
        assert got_status == status, (n_ones, tail)
        if status:
            with pytest.raises(SymgpuError) as e2:
                fe.decode(pkt)
            assert e2.value.status == status, (n_ones, tail)
        else:
            units, tns, coeffs = fe.decode(pkt)
            assert np.array_equal(u32(coeffs[0]), u32(want[0]["coeffs"]))
            v = (1 << (n_ones + 4)) + 5
            assert float(coeffs[0][0]) == -float(ab.pow43(v) * ab.scale_normal(140)) and float(coeffs[0][1]) == float(ab.scale_normal(140))
        fe.close()


# ---- boundary cases found by mutating the front-end (tools/mutate_frontend.py): each pins one rule at its edge ---------------

def _sce(gg, max_sfb, sections, scf=(), pulse=None, tns=None, spectral=(), short=False, trailing=((7, 3),), rate=44100, channels=1):
    """A hand-built single_channel_element: sections [(book, length)], scf [('pcm', v) | ('d', diff)], pulse (start, [(off, amp)]),
    tns [(value, width)] after the presence bit, spectral [(book, index)], trailing [(value, width)].  Returns the packet."""
    w = ab.BitWriterMsb()
    w.put(0, 3), w.put(0, 4), w.put(gg, 8)
    if short:
        w.put(0, 1), w.put(2, 2), w.put(0, 1), w.put(max_sfb, 4)
        for _ in range(7):
            w.put(1, 1)                                    # one group of eight windows
    else:
        w.put(0, 1), w.put(0, 2), w.put(0, 1), w.put(max_sfb, 6), w.put(0, 1)
    for cb, n in sections:
        w.put(cb, 4), w.put(n, 3 if short else 5)
    for kind, v in scf:
        w.put(v, 9) if kind == "pcm" else w.huff("scf", v + 60)
    if pulse is None:
        w.put(0, 1)
    else:
        w.put(1, 1), w.put(len(pulse[1]) - 1, 2), w.put(pulse[0], 6)
        for off, amp in pulse[1]:
            w.put(off, 5), w.put(amp, 4)
    if tns is None:
        w.put(0, 1)
    else:
        w.put(1, 1)
        for v, width in tns:
            w.put(v, width)
    w.put(0, 1)
    for book, idx in spectral:
        w.huff(book, idx)
    for v, width in trailing:
        w.put(v, width)
    return w.bytes()


def _both_status(pkt, rate=44100, channels=1):
    """(status, oracle result or None, front-end result or None); the two must agree."""
    fe, o = frontend.AacFrontend(rate, channels), ao.AacFrontend(rate, channels)
    try:
        covered, want = o.decode(pkt)
        status = 0 if covered == channels else 2
    except ao.AacError as e:
        status, want = (1 if e.kind == ao.DECODE else 2), None
    try:
        got = fe.decode(pkt)
        got_status = 0
    except SymgpuError as e:
        got, got_status = None, e.status
    fe.close()
    assert got_status == status
    if status == 0:
        for c in range(channels):
            assert np.array_equal(u32(got[2][c]), u32(want[c]["coeffs"]))
    return status, want, got


QUAD_ZERO = ("1", 40)   # book 1, the all-zero quad (digits 1 1 1 1)


def test_element_loop_stops_with_three_bits_left():
    # mod.rs:131 `while bs.bits_left() > 3`: 29 bits of element, 3 zero bits of padding -- which would read as another SCE id
    pkt = _sce(120, 0, [], trailing=())
    assert len(pkt) == 4
    assert _both_status(pkt)[0] == 0
    # four bits left: the loop goes round again, reads element id 0 and runs out of data
    assert _both_status(_sce(120, 0, [], trailing=((0, 1),)) + b"")[0] == 0          # 30 bits + 2 padding
    assert _both_status(_sce(120, 0, [], trailing=((0, 4),)) + b"\x00")[0] == 1       # a whole zero byte more: SCE, then nothing


def test_section_rules_at_their_edges():
    # 64 empty sections are the limit: the 65th is refused even though it would be valid (ics/mod.rs:243-246)
    assert _both_status(_sce(120, 1, [(0, 0)] * 63 + [(0, 1)]))[0] == 0
    assert _both_status(_sce(120, 1, [(0, 0)] * 64 + [(0, 1)]))[0] == 1
    # book 12 is reserved (:251-253)
    assert _both_status(_sce(120, 2, [(12, 2)], scf=[("d", 0), ("d", 0)]))[0] == 1    # complete but for the band type
    assert _both_status(_sce(120, 2, [(15, 2)], scf=[("d", 0), ("d", 0)]))[0] == 0
    assert _both_status(_sce(120, 2, [(0, 2)]))[0] == 0
    # a section may end at max_sfb, not beyond (:265)
    assert _both_status(_sce(120, 2, [(0, 3)]))[0] == 1
    # max_sfb may name every band, not one more (:292-300): 14 short bands at 48 kHz
    assert _both_status(_sce(120, 14, [(0, 6), (0, 6), (0, 2)], short=True), rate=48000)[0] == 0
    assert _both_status(_sce(120, 15, [(0, 6), (0, 6), (0, 3)], short=True), rate=48000)[0] == 1


def test_scale_factor_ranges_at_their_edges():
    # normal: index 255 is the last valid (ics/mod.rs:344-347)
    st, want, _ = _both_status(_sce(255, 1, [(1, 1)], scf=[("d", 0)], spectral=[QUAD_ZERO]))
    assert st == 0
    assert _both_status(_sce(255, 1, [(1, 1)], scf=[("d", 1)], spectral=[QUAD_ZERO]))[0] == 1
    assert _both_status(_sce(0, 1, [(1, 1)], scf=[("d", -1)], spectral=[QUAD_ZERO]))[0] == 1
    # noise: global gain 100 -> 110, first value 9 bits offset by 256 (:326-336); index 0 valid, -1 not
    assert _both_status(_sce(100, 1, [(13, 1)], scf=[("pcm", 146)]))[0] == 0
    assert _both_status(_sce(100, 1, [(13, 1)], scf=[("pcm", 145)]))[0] == 1
    # intensity: starts at 155; 255 valid, 256 not (:316-323)
    assert _both_status(_sce(100, 2, [(15, 2)], scf=[("d", 60), ("d", 40)]))[0] == 0
    assert _both_status(_sce(100, 2, [(15, 2)], scf=[("d", 60), ("d", 41)]))[0] == 1


def test_tns_order_and_pulse_rules_at_their_edges():
    # long window, AAC-LC: order 12 is the limit (tns.rs:118-128, :52): n_filt 1, coef_res 0, length 10, order, direction 0, compress 0, 3-bit values
    def tns(order):
        return [(1, 2), (0, 1), (10, 6), (order, 5), (0, 1), (0, 1)] + [(3, 3)] * order
    st, want, got = _both_status(_sce(120, 4, [(0, 4)], tns=tns(12)))
    assert st == 0 and len(got[1]) == 1 and int(got[1][0]["order"]) == 12
    assert _both_status(_sce(120, 4, [(0, 4)], tns=tns(13)))[0] == 1
    # pulse data in a short window is refused (ics/mod.rs:427)
    assert _both_status(_sce(120, 1, [(0, 1)], pulse=(0, [(1, 1)]), short=True))[0] == 1
    assert _both_status(_sce(120, 1, [(0, 1)], pulse=(0, [(1, 1)])))[0] == 0
    # pulses stop at line 1024 (pulse.rs:78-80): last band of the 48 kHz table starts at 928; 31 + 31 + 31 + 3 = 96
    n_bands = len(ao.L48) - 1
    st, want, got = _both_status(_sce(150, n_bands, [(0, 30), (0, n_bands - 30)], pulse=(n_bands - 1, [(31, 3), (31, 2), (31, 1), (3, 7)])), rate=48000)
    assert st == 0 and np.count_nonzero(got[2][0]) == 0      # zero scale: every restored line is a signed zero
    assert [bool(np.signbit(got[2][0][928 + k])) for k in (31, 62, 93)] == [True, True, True] and not np.signbit(got[2][0][1023])


def test_ms_mask_three_is_refused():
    def cpe(mask):
        w = ab.BitWriterMsb()
        w.put(1, 3), w.put(0, 4), w.put(1, 1)                                            # CPE, common window
        w.put(0, 1), w.put(0, 2), w.put(0, 1), w.put(0, 6), w.put(0, 1)                   # ics_info, max_sfb 0
        w.put(mask, 2)
        for _ in range(2):
            w.put(120, 8), w.put(0, 1), w.put(0, 1), w.put(0, 1)                          # gain, no pulse / TNS / gain control
        w.put(7, 3)
        return w.bytes()
    assert _both_status(cpe(2), channels=2)[0] == 0
    assert _both_status(cpe(3), channels=2)[0] == 1


# ---- AudioSpecificConfig ---------------------------------------------------------------------------------------------------------

def _asc_bits(aot, rate_idx, ch_cfg, rate=None, tail=()):
    w = ab.BitWriterMsb()
    if aot < 31:
        w.put(aot, 5)
    else:
        w.put(31, 5), w.put(aot - 32, 6)
    w.put(rate_idx, 4)
    if rate_idx == 15:
        w.put(rate, 24)
    w.put(ch_cfg, 4)
    for v, width in tail:
        w.put(v, width)
    return w


def _asc_both(blob):
    try:
        want = ao.read_asc(blob)
        status = 0
    except ao.AacError as e:
        want, status = None, (1 if e.kind == ao.DECODE else 2)
    try:
        got = frontend.aac_asc_parse(blob)
        got_status = 0
    except SymgpuError as e:
        got, got_status = None, e.status
    assert got_status == status, blob.hex()
    if status == 0:
        for k, v in want.items():
            assert int(got[k]) == v, (blob.hex(), k)
        verdict = ao.decoder_accepts(want)
        try:
            fe = frontend.AacFrontend(extra_data=blob)
            assert verdict is None and (fe.sample_rate, fe.channels) == (want["sample_rate"], want["channels"])
            fe.close()
        except SymgpuError as e:
            assert verdict == ao.UNSUPPORTED and e.status == 2, blob.hex()
    return status, want


def test_audio_specific_config_known_answers():
    # the two configurations every AAC-LC encoder writes: 0x1210 = LC, 44.1 kHz, stereo; 0x1190 = LC, 48 kHz, stereo
    st_, a = _asc_both(bytes.fromhex("1210"))
    assert st_ == 0 and (a["object_type"], a["sample_rate"], a["channels"], a["samples"], a["sbr_present"]) == (2, 44100, 2, 1024, 0)
    st_, a = _asc_both(bytes.fromhex("1190"))
    assert st_ == 0 and (a["object_type"], a["sample_rate"], a["channels"]) == (2, 48000, 2)
    st_, a = _asc_both(bytes.fromhex("1208"))   # LC, 44.1 kHz, mono
    assert st_ == 0 and a["channels"] == 1
    # HE-AAC, explicit hierarchical signalling: SBR (5), core 24 kHz, stereo, extension rate 48 kHz, then LC + GASpecificConfig
    w = _asc_bits(5, 6, 2, tail=((3, 4), (2, 5), (0, 1), (0, 1), (0, 1)))
    st_, a = _asc_both(w.bytes())
    assert st_ == 0 and (a["object_type"], a["sbr_present"], a["ext_sample_rate"], a["sample_rate"]) == (2, 1, 48000, 24000)
    with pytest.raises(SymgpuError) as e:       # ... which the LC decoder refuses as too complex
        frontend.AacFrontend(extra_data=w.bytes())
    assert e.value.status == 2
    # LC followed by the backward-compatible SBR signalling (sync 0x2b7) is only looked at after an explicit prefix: plain LC stays LC
    w = _asc_bits(2, 4, 2, tail=((0, 3), (0x2B7, 11), (5, 5), (1, 1), (3, 4)))
    st_, a = _asc_both(w.bytes())
    assert st_ == 0 and a["sbr_present"] == 0
    # escapes: explicit 24-bit rate; 960-sample frames refused by the decoder; channel configuration 0 needs a program config element
    st_, a = _asc_both(_asc_bits(2, 15, 2, rate=44100, tail=((0, 3),)).bytes())
    assert st_ == 0 and a["sample_rate"] == 44100
    assert _asc_both(_asc_bits(2, 15, 2, rate=0, tail=((0, 3),)).bytes())[0] == 1
    st_, a = _asc_both(_asc_bits(2, 4, 2, tail=((1, 1), (0, 2))).bytes())
    assert st_ == 0 and a["samples"] == 960
    assert _asc_both(_asc_bits(2, 4, 0, tail=((0, 3),)).bytes())[0] == 2
    assert _asc_both(_asc_bits(2, 13, 2, tail=((0, 3),)).bytes())[0] == 1          # reserved rate index
    assert _asc_both(_asc_bits(2, 4, 9, tail=((0, 3),)).bytes())[0] == 1           # reserved channel configuration
    assert _asc_both(b"\x12")[0] == 1                                               # ends inside the configuration
    with pytest.raises(SymgpuError) as e:                                           # aac/mod.rs:60: at least two bytes
        frontend.AacFrontend(extra_data=b"\x12")
    assert e.value.status == 1


def test_audio_specific_config_every_object_type_and_random_tails():
    rng = np.random.default_rng(77)
    seen = set()
    for aot in list(range(0, 31)) + list(range(32, 48)) + [95]:
        for trial in range(12):
            ch = int(rng.integers(0, 8)) if trial else 2
            tail = [(int(rng.integers(2)), 1) for _ in range(int(rng.integers(0, 40)))]
            blob = _asc_bits(aot, int(rng.choice([3, 4, 6, 11, 15])), ch, rate=int(rng.integers(1, 1 << 24)), tail=tail).bytes()
            st_, a = _asc_both(blob)
            seen.add((aot, st_))
    assert {s for _, s in seen} == {0, 1, 2}
    assert (2, 0) in seen and (8, 2) in seen and (39, 2) in seen and (4, 0) in seen
    # random bytes
    for _ in range(400):
        _asc_both(rng.integers(0, 256, int(rng.integers(0, 9)), dtype=np.uint8).tobytes())


def test_audio_specific_config_extension_edges():
    """Cases the randomised tails did not pin (found by mutating the parser, tools/mutate_frontend.py asc)."""
    # ER AAC LD (23) with the extension flag: three resilience bits are skipped before extensionFlag3 (mod.rs:307-314)
    st_, a = _asc_both(_asc_bits(23, 4, 2, tail=((0, 1), (0, 1), (1, 1), (7, 3), (0, 1), (0, 2))).bytes())
    assert st_ == 0 and a["object_type"] == 23
    # backward-compatible signalling behind an explicit SBR prefix: sync 0x2b7, SBR, sbr_present = 0 switches SBR OFF again (:390-404) ...
    core = ((3, 4), (2, 5), (0, 3))                                   # extension rate 48 kHz, object type LC, GASpecificConfig
    st_, a = _asc_both(_asc_bits(5, 6, 2, tail=core + ((0x2B7, 11), (5, 5), (0, 1))).bytes())
    assert st_ == 0 and a["sbr_present"] == 0 and a["has_ext"] == 1 and ao.decoder_accepts(a) is None
    st_, a = _asc_both(_asc_bits(5, 6, 2, tail=core + ((0x2B6, 11), (5, 5), (0, 1))).bytes())
    assert st_ == 0 and a["sbr_present"] == 1
    # ... and with sbr_present = 1 a second sync (0x548) carries the PS flag, if twelve bits are left for it (:396-402)
    st_, a = _asc_both(_asc_bits(5, 6, 2, tail=core + ((0x2B7, 11), (5, 5), (1, 1), (3, 4), (0x548, 11), (1, 1))).bytes())
    assert st_ == 0 and (a["sbr_present"], a["ps_present"]) == (1, 1)
    st_, a = _asc_both(_asc_bits(5, 6, 2, tail=core + ((0x2B7, 11), (5, 5), (1, 1), (3, 4), (0x549, 11), (1, 1))).bytes())
    assert st_ == 0 and a["ps_present"] == 0
    # (exactly twelve bits left are enough for it: a configuration with the 14-bit core-coder delay ends on a byte boundary here)
    w = _asc_bits(5, 6, 2, tail=((3, 4), (2, 5), (0, 1), (1, 1), (0, 14), (0, 1), (0x2B7, 11), (5, 5), (1, 1), (3, 4), (0x548, 11), (1, 1)))
    assert len(w.bits) == 72
    st_, a = _asc_both(w.bytes())
    assert st_ == 0 and a["ps_present"] == 1
    # exactly sixteen bits behind the configuration are enough to look for the sync -- and then the flag behind it is missing
    found = 0
    for inner in (2, 6, 17, 20):
        for depends in (0, 1):
            for ext in (0, 1):
                tail = ((3, 4), (inner, 5), (0, 1), (depends, 1)) + (((0, 14),) if depends else ()) + ((ext, 1),)
                if inner in (6, 20):
                    tail += ((0, 3),)                                   # layer number
                if ext:
                    tail += (((0, 3),) if inner in (17, 20) else ()) + ((0, 1),)   # resilience flags, extensionFlag3
                if inner in (17, 20):
                    tail += ((0, 2),)                                   # epConfig
                w = _asc_bits(5, 6, 2, tail=tail)
                if len(w.bits) % 8:
                    continue
                found += 1
                blob = _asc_bits(5, 6, 2, tail=tail + ((0x2B7, 11), (5, 5))).bytes()
                assert len(blob) * 8 == len(w.bits) + 16
                assert _asc_both(blob)[0] == 1
                assert _asc_both(blob + b"\x00")[0] == 0
    assert found


# ---- the blocks of a stream as independent jobs ---------------------------------------------------------------------------------

def test_blocks_as_independent_jobs_equal_the_serial_front_end():
    """symgpu_aac_fe_decode_packets_jobs: every block from a fresh state, noise generators jumped ahead over the prefix sums of the
    draws, window history chained afterwards -- the same bits as the serial front-end, for any thread count."""
    done = noisy = 0
    for seed, s, _ in _streams(600, 20):
        pk = [s.packet()[0] for _ in range(14)]
        blob = b"".join(pk)
        table = np.zeros(len(pk), dtype=nat.PIECE_DTYPE)
        table["len"] = [len(p) for p in pk]
        table["offset"] = np.concatenate([[0], np.cumsum(table["len"][:-1], dtype=np.uint64)])
        fe = frontend.AacFrontend(s.rate, s.channels)
        units, tns, coeffs, frame_of = fe.decode_packets(blob, table, tns_base=7)
        fe.close()
        assert len(units) == len(pk)
        for threads in (1, 3, 8):
            got = frontend.aac_decode_packets_jobs(s.rate, s.channels, blob, table, tns_base=7, threads=threads)
            assert got is not None, seed                     # (the writer keeps pulses inside the coded bands: no stale scale is read)
            assert got[0].tobytes() == units.tobytes() and got[1].tobytes() == tns.tobytes(), (seed, threads)
            assert np.array_equal(u32(got[2]), u32(coeffs)), (seed, threads)
        done += 1
        o = ao.AacFrontend(s.rate, s.channels)
        for p in pk[:-1]:
            o.decode(p)
        noisy += int(any(q.lcg.state != 0x1F2E3D4C for q in o.pairs))   # noise was drawn before the last block: its start state is a jump
    assert done == 20 and noisy >= 15
    # the generator's jump-ahead against stepping it
    lcg = ao.Lcg()
    states = [lcg.state]
    for _ in range(3000):
        lcg.next()
        states.append(lcg.state)
    a, c = 1664525, 1013904223

    def jump(s0, n):
        ra, rc, aa, cc = 1, 0, a, c
        while n:
            if n & 1:
                ra, rc = (ra * aa) % (1 << 32), (rc * aa + cc) % (1 << 32)
            cc, aa = (cc * aa + cc) % (1 << 32), (aa * aa) % (1 << 32)
            n >>= 1
        return (ra * s0 + rc) % (1 << 32)
    assert all(jump(states[0], n) == states[n] for n in (0, 1, 2, 3, 17, 1000, 2999, 3000))


def test_streams_that_need_the_serial_path_say_so():
    rng = np.random.default_rng(5)
    s = ab.Stream(rng, 44100, 2)
    pk = [s.packet()[0] for _ in range(8)]
    pk[3] = pk[3][:len(pk[3]) // 2]                       # a block that ends early: refused, and the reference's state carries on from its middle
    blob = b"".join(pk)
    table = np.zeros(len(pk), dtype=nat.PIECE_DTYPE)
    table["len"] = [len(p) for p in pk]
    table["offset"] = np.concatenate([[0], np.cumsum(table["len"][:-1], dtype=np.uint64)])
    assert frontend.aac_decode_packets_jobs(44100, 2, blob, table) is None
    # a layout that changes between blocks (two single-channel elements, then a pair)
    two, pair = ab.Stream(rng, 44100, 2, layout=["sce", "sce"]), ab.Stream(rng, 44100, 2)
    pk = [two.packet(extras=False)[0], pair.packet(extras=False)[0]]
    blob = b"".join(pk)
    table = np.zeros(2, dtype=nat.PIECE_DTYPE)
    table["len"] = [len(p) for p in pk]
    table["offset"] = [0, len(pk[0])]
    assert frontend.aac_decode_packets_jobs(44100, 2, blob, table) is None
    # a pulse that lands above the coded bands reads a scale an earlier block left behind
    w = _sce(150, 2, [(1, 2)], scf=[("d", 0), ("d", 0)], pulse=(10, [(0, 5)]), spectral=[QUAD_ZERO, QUAD_ZERO])
    table = np.zeros(1, dtype=nat.PIECE_DTYPE)
    table["len"] = len(w)
    assert frontend.aac_decode_packets_jobs(44100, 1, w, table) is None
    assert frontend.aac_decode_packets_jobs(44100, 1, _sce(150, 2, [(1, 2)], scf=[("d", 0), ("d", 0)], pulse=(1, [(0, 5)]), spectral=[QUAD_ZERO, QUAD_ZERO]), table) is not None
