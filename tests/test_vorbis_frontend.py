"""Vorbis entropy front-end (`symgpu_vorbis_fe_*`, SURVEY §8f N1 for the Vorbis path): codebooks, floor-1 packet decode, residue
types 0 / 1 / 2 and the packet-level steps up to inverse coupling, against oracle/vorbis_frontend_oracle.py (the reference's
sequence incl. its bit-cache behaviour on packets that end early) and against an independent stream writer's ground truth.
The oracle itself is pinned to the reference's codebook unit tests (codebook.rs:402-485).  VQ vectors are sums of single IEEE
operations on exact inputs, so the bar is bit equality.  CPU only."""
import numpy as np
import pytest

import symphonia_b200 as sb  # noqa: F401  (builds / loads the library)
from oracle import packetizer_oracle as po
from oracle import vorbis_frontend_oracle as vo
from symphonia_b200 import frontend, packetizer
from symphonia_b200.engine import SymgpuError
from tests import _oracle
from tests import _vorbis_bitstream as vb


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---- the oracle against the reference's own known answers -----------------------------------------------------------------

def test_oracle_ilog_known_answers():
    # codebook.rs verify_ilog
    assert [po.ilog(v) for v in (0, 1, 2, 3, 4, 7)] == [0, 1, 2, 2, 3, 3]


def test_oracle_lookup1_values_known_answers():
    # codebook.rs verify_lookup1_values: equal to the naive search (largest x with x^dims <= entries), overflow = "too large"
    def naive(entries, dims):
        x = 1
        if dims > 0:
            while x ** dims <= entries and x ** dims < (1 << 32):
                x += 1
        return x - 1

    for entries, dims in ((1, 0), (0, 1), (1, 1), (361, 2), (560, 3), (3, 950), (0xFFFF, 0xFF), (0, 65535), (1, 65535),
                          (0xFFFFFF, 65535)):
        assert vo.lookup1_values(entries, dims) == naive(entries, dims), (entries, dims)
    assert vo.lookup1_values(361, 2) == 19 and vo.lookup1_values(560, 3) == 8 and vo.lookup1_values(3, 950) == 1


def test_oracle_synthesize_codewords_known_answers():
    # codebook.rs verify_synthesize_codewords, verify_synthesize_codewords_overspecified
    assert vo.synthesize_codewords([2, 4, 4, 4, 4, 2, 3, 3]) == [0, 0x4, 0x5, 0x6, 0x7, 0x2, 0x6, 0x7]
    for lens in ([1, 1, 1], [1, 1, 32]):
        with pytest.raises(po.ReaderError):
            vo.synthesize_codewords(lens)
    # under-specified trees are refused as well (codebook.rs:199-207)
    with pytest.raises(po.ReaderError):
        vo.synthesize_codewords([1, 2])


def test_oracle_codewords_equal_the_specification_for_random_trees():
    # the writer assigns codewords by the specification's rule (lowest free leaf), the oracle by the reference's table walk
    rng = np.random.default_rng(7)
    for _ in range(200):
        lens = vb.random_lengths(rng, int(rng.integers(2, 60)))
        if rng.random() < 0.4:  # sparse: unused entries in between
            lens = [x for n in lens for x in ([n] + [0] * int(rng.integers(0, 2)))]
        assert vo.synthesize_codewords(lens) == [c for c in vb.canonical_codewords(lens) if c is not None]


def test_oracle_float32_unpack():
    # Vorbis I 9.2.2: mantissa * 2^(exponent - 788), sign in bit 31
    assert float(vo.float32_unpack(vb.pack_float32(1, 788))) == 1.0
    assert float(vo.float32_unpack(vb.pack_float32(3, 787, True))) == -1.5
    assert float(vo.float32_unpack(vb.pack_float32(0x1FFFFF, 788 - 21))) == 0x1FFFFF / 2.0 ** 21
    assert float(vo.float32_unpack(0)) == 0.0


# ---- oracle vs the writer's ground truth ---------------------------------------------------------------------------------

def _same(a, b, what):
    assert bool(a["block_flag"]) == bool(b["block_flag"]) and bool(a["prev_block_flag"]) == bool(b["prev_block_flag"]), what
    assert [bool(x) for x in a["do_not_decode"]] == [bool(x) for x in b["do_not_decode"]], what
    assert list(a["floor"]) == list(b["floor"]), what
    assert np.array_equal(a["floor_y"], b["floor_y"]), what
    assert np.array_equal(bits(a["residue"]), bits(b["residue"])), what


@pytest.mark.parametrize("residue_type", [0, 1, 2])
def test_oracle_equals_writer_truth(residue_type):
    for seed in range(14):
        rng = np.random.default_rng(1000 * residue_type + seed)
        s = vb.Stream(rng, channels=1 if seed % 5 == 4 else 2, residue_type=residue_type, per_word=1)
        o = vo.VorbisFrontend(s.ident, s.setup)
        slot = (1 << s.bs_exp[1]) >> 1
        for k in range(10):
            pkt, truth = s.packet()
            _same(o.decode(pkt, slot), truth, (residue_type, seed, k))


# ---- the C++ front-end vs the oracle ------------------------------------------------------------------------------------

def _as_dict(unit, floor_y, residue):
    fl = [None if int(v) == 0xFFFF else int(v) for v in unit["floor"]]
    return dict(block_flag=int(unit["block_flag"]), prev_block_flag=int(unit["prev_block_flag"]),
                do_not_decode=[int(v) for v in unit["do_not_decode"]], floor=fl, floor_y=floor_y, residue=residue)


def _both(fe, o, pkt, slot, what):
    """Decodes `pkt` with both; equal results or both refuse."""
    try:
        want = o.decode(pkt, slot)
    except po.ReaderError:
        with pytest.raises(SymgpuError) as e:
            fe.decode(pkt, slot)
        assert e.value.status == 1, what
        return None
    got = _as_dict(*fe.decode(pkt, slot))
    # a channel the oracle reports as not decoded carries no floor index on either side
    _same(got, want, what)
    return want


def test_frontend_equals_oracle_on_clean_streams():
    for seed in range(36):
        rng = np.random.default_rng(5000 + seed)
        s = vb.Stream(rng, channels=1 if seed % 6 == 5 else 2)
        fe, o = frontend.VorbisFrontend(s.ident, s.setup), vo.VorbisFrontend(s.ident, s.setup)
        slot = fe.slot
        assert slot == (1 << s.bs_exp[1]) >> 1
        for k in range(8):
            pkt, _ = s.packet()
            _both(fe, o, pkt, slot, (seed, k))
        fe.close()


def test_frontend_config_equals_setup_parse():
    for seed in range(12):
        s = vb.Stream(np.random.default_rng(6000 + seed))
        fe = frontend.VorbisFrontend(s.ident, s.setup)
        ident = packetizer.vorbis_ident(s.ident)
        info, floors = packetizer.vorbis_setup_parse(s.setup, ident)
        assert len(fe.floors) == len(floors) == len(s.floors) and fe.floors.tobytes() == floors.tobytes()
        assert int(fe.stream["bs0_exp"]) == s.bs_exp[0] and int(fe.stream["bs1_exp"]) == s.bs_exp[1]
        assert int(fe.stream["channels"]) == s.channels and bool(fe.stream["coupled"]) == s.coupled
        fe.close()


def test_packets_that_end_early_and_damaged_packets():
    """A Vorbis packet may end anywhere: floors read so far stand, the residue keeps what was decoded, a failed floor read marks
    the channel unused (floor.rs:661-667, residue.rs:391-399).  The reference's reader works from a 64-bit cache, so what a
    failed read leaves behind is part of the behaviour; the C++ front-end follows it."""
    n_cut = n_flip = 0
    for seed in range(16):
        rng = np.random.default_rng(7000 + seed)
        s = vb.Stream(rng)
        fe, o = frontend.VorbisFrontend(s.ident, s.setup), vo.VorbisFrontend(s.ident, s.setup)
        for k in range(10):
            pkt, _ = s.packet()
            mode = k % 3
            if mode == 1 and len(pkt) > 1:
                pkt = pkt[:int(rng.integers(0, len(pkt)))]
                n_cut += 1
            elif mode == 2:
                b = bytearray(pkt)
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(len(b)))] ^= 1 << int(rng.integers(8))
                pkt = bytes(b)
                n_flip += 1
            _both(fe, o, pkt, fe.slot, (seed, k, mode))
        fe.close()
    assert n_cut > 30 and n_flip > 30


def test_empty_and_non_audio_packets_are_refused():
    s = vb.Stream(np.random.default_rng(1))
    fe = frontend.VorbisFrontend(s.ident, s.setup)
    for pkt in (b"", b"\x01", s.ident, s.setup):
        with pytest.raises(SymgpuError) as e:
            fe.decode(pkt)
        assert e.value.status == 1
    fe.close()


def test_reset_forgets_the_previous_block():
    rng = np.random.default_rng(11)
    while True:
        s = vb.Stream(rng)
        flags = {f for f, _ in s.modes}
        if len(flags) == 2:
            break
    fe = frontend.VorbisFrontend(s.ident, s.setup)
    seen = []
    for k in range(30):
        pkt, _ = s.packet()
        if k % 7 == 3:
            fe.reset()
            unit, _, _ = fe.decode(pkt)
            assert int(unit["prev_block_flag"]) == int(unit["block_flag"])  # dsp.prev_block_flag.unwrap_or(block_flag), lib.rs:298
        else:
            unit, _, _ = fe.decode(pkt)
            if seen:
                assert int(unit["prev_block_flag"]) == seen[-1]
        seen.append(int(unit["block_flag"]))
    fe.close()


def test_bad_headers_are_refused_at_create():
    s = vb.Stream(np.random.default_rng(3))
    with pytest.raises(SymgpuError):
        frontend.VorbisFrontend(s.ident[:-1], s.setup)
    with pytest.raises(SymgpuError):
        frontend.VorbisFrontend(s.ident, s.setup[:len(s.setup) // 2])
    # a setup header whose first codebook has lost its sync pattern
    bad = bytearray(s.setup)
    bad[8] ^= 0xFF
    with pytest.raises(SymgpuError):
        frontend.VorbisFrontend(s.ident, bytes(bad))
    # damaged setup headers: the front-end and the oracle agree on accept / refuse
    rng = np.random.default_rng(4)
    agree = 0
    for _ in range(60):
        b = bytearray(s.setup)
        b[int(rng.integers(7, len(b)))] ^= 1 << int(rng.integers(8))
        try:
            vo.VorbisFrontend(s.ident, bytes(b))
            ok = True
        except (po.ReaderError, vo.End):
            ok = False
        try:
            frontend.VorbisFrontend(s.ident, bytes(b)).close()
            got = True
        except SymgpuError:
            got = False
        assert got == ok
        agree += 1
    assert agree == 60


# ---- packets -> front-end -> synthesis oracle: the descriptors are what the synthesis stage takes ---------------------------

def _chain_workload(seed, n_packets=24, floor_posts_min=2):
    rng = np.random.default_rng(seed)
    s = vb.Stream(rng, bs_exp=(8, 11))
    fe = frontend.VorbisFrontend(s.ident, s.setup)
    units = np.zeros(n_packets, dtype=sb._native.VORBIS_UNIT_DTYPE)
    floor_y = np.zeros((n_packets, 2, 65), dtype=np.uint16)
    residue = np.zeros((n_packets, 2, fe.slot), dtype=np.float32)
    for k in range(n_packets):
        pkt, _ = s.packet()
        units[k], floor_y[k], residue[k] = fe.decode(pkt)
    runs = np.zeros(1, dtype=sb._native.VORBIS_RUN_DTYPE)
    runs["n_packets"] = n_packets
    streams = np.array([fe.stream], dtype=sb._native.VORBIS_STREAM_DTYPE)
    wl = dict(streams=streams, floors=fe.floors.copy(), units=units, floor_y=floor_y, residue=residue, runs=runs, slot=fe.slot)
    fe.close()
    return wl


def test_front_end_output_feeds_the_synthesis_oracle():
    lib = _oracle.load()
    for seed in range(6):
        wl = _chain_workload(9000 + seed)
        rc, pcm = _oracle.vorbis_batch(lib, wl)
        assert rc == 0
        assert np.isfinite(pcm).all()
        assert np.abs(pcm).max() > 0.0


# ---- boundary cases found by mutating the front-end (tools/mutate_frontend.py) -----------------------------------------------

def test_mode_number_beyond_the_mode_list_is_refused():
    # lib.rs:158-163: with three modes the field is two bits wide and the value 3 names no mode
    seed = 0
    while True:
        s = vb.Stream(np.random.default_rng(8000 + seed))
        if len(s.modes) == 3:
            break
        seed += 1
    fe, o = frontend.VorbisFrontend(s.ident, s.setup), vo.VorbisFrontend(s.ident, s.setup)
    for mode, ok in ((0, True), (2, True), (3, False)):
        pkt = bytes([mode << 1]) + bytes(40)       # audio packet, mode number, everything after it zero
        if ok:
            _both(fe, o, pkt, fe.slot, ("mode", mode))
        else:
            with pytest.raises(po.ReaderError):
                o.decode(pkt, fe.slot)
            with pytest.raises(SymgpuError) as e:
                fe.decode(pkt)
            assert e.value.status == 1
    fe.close()


def test_residue_that_begins_beyond_the_short_block():
    # residue.rs:150-160: begin and end are both limited to the block's vector, a residue that starts beyond a short block codes
    # nothing there (and everything it has in the long block)
    for seed in range(6):
        rng = np.random.default_rng(8100 + seed)
        s = vb.Stream(rng, bs_exp=(7, 9), residue_begin=96, residue_type=seed % 3)   # short vectors are 64 long (128 interleaved)
        fe, o = frontend.VorbisFrontend(s.ident, s.setup), vo.VorbisFrontend(s.ident, s.setup)
        flags = set()
        for k in range(12):
            pkt, truth = s.packet()
            want = _both(fe, o, pkt, fe.slot, (seed, k))
            flags.add(bool(want["block_flag"]))
            if s.per_word == 1:
                _same(want, truth, (seed, k, "truth"))
        fe.close()


def _patched_book(monkeypatch, which, **change):
    """Streams whose `which`-th codebook is built with changed arguments."""
    real, count = vb.Book, [0]

    def make(rng, entries, dims, vq=None, style=None, **kw):
        count[0] += 1
        if count[0] == which:
            if "vq" in change and vq is not None:
                vq = (change["vq"], vq[1])
            style = change.get("style", style)
        return real(rng, entries, dims, vq=vq, style=style, **kw)
    monkeypatch.setattr(vb, "Book", make)


def test_unknown_lookup_type_is_refused(monkeypatch):
    # codebook.rs:290-352: lookup types 0, 1 and 2 exist
    s = vb.Stream(np.random.default_rng(8200))
    first_vq = next(k for k, b in enumerate(s.books) if b.vq is not None) + 1
    _patched_book(monkeypatch, first_vq, vq=3)
    bad = vb.Stream(np.random.default_rng(8200))
    with pytest.raises(po.ReaderError):
        vo.VorbisFrontend(bad.ident, bad.setup)
    with pytest.raises(SymgpuError) as e:
        frontend.VorbisFrontend(bad.ident, bad.setup)
    assert e.value.status == 1


def test_over_and_under_specified_codebooks_are_refused():
    # codebook.rs:112-210: a plain-coded book's lengths sit in 5-bit fields; all ones = length 2 for every entry
    for seed in range(200):
        s = vb.Stream(np.random.default_rng(8300 + seed))
        w = vb.BitWriterRtl()
        b0 = s.books[0]
        # is book 0 plain-coded (ordered flag 0, sparse flag 0)?
        head = int.from_bytes(s.setup[7:], "little") >> 8
        if (head >> 64) & 3 != 0 or b0.entries < 6:
            continue
        for new_len in (2, 5):                     # 6+ entries of length 2: over-specified; of length 5 (< 32 entries): under-specified
            if new_len == 5 and b0.entries >= 32:
                continue
            body = int.from_bytes(s.setup[7:], "little")
            at = 8 + 66
            for k in range(b0.entries):
                body &= ~(31 << (at + 5 * k))
                body |= (new_len - 1) << (at + 5 * k)
            setup = s.setup[:7] + body.to_bytes(len(s.setup) - 7, "little")
            with pytest.raises(po.ReaderError):
                vo.VorbisFrontend(s.ident, setup)
            with pytest.raises(SymgpuError) as e:
                frontend.VorbisFrontend(s.ident, setup)
            assert e.value.status == 1
        del w
        return
    raise AssertionError("no plain-coded first book among the seeds")


def test_packets_as_independent_jobs_equal_the_serial_front_end():
    """symgpu_vorbis_fe_decode_packets_jobs: every thread its own front-end, packets in any order, previous block flags chained over the
    accepted packets afterwards -- the serial front-end's bits on clean, truncated and damaged packets alike."""
    refused = 0
    for seed in range(18):
        rng = np.random.default_rng(9500 + seed)
        s = vb.Stream(rng, channels=1 if seed % 6 == 5 else 2, residue_type=seed % 3)
        pk = []
        for k in range(16):
            p, _ = s.packet()
            if k % 4 == 1 and len(p) > 2:
                p = p[:int(rng.integers(1, len(p)))]
            elif k % 4 == 2:
                b = bytearray(p)
                b[int(rng.integers(len(b)))] ^= 1 << int(rng.integers(8))
                p = bytes(b)
            elif k == 7:
                p = b"\x01" + p                               # not an audio packet: refused
            pk.append(p)
        blob = b"".join(pk)
        table = np.zeros(len(pk), dtype=sb._native.PIECE_DTYPE)
        table["len"] = [len(p) for p in pk]
        table["offset"] = np.concatenate([[0], np.cumsum(table["len"][:-1], dtype=np.uint64)])
        fe = frontend.VorbisFrontend(s.ident, s.setup)
        units, fy, res, keep = fe.decode_packets(blob, table)
        slot = fe.slot
        fe.close()
        refused += len(pk) - len(keep)
        for threads in (1, 3, 8):
            ju, jf, jr, acc = frontend.vorbis_decode_packets_jobs(s.ident, s.setup, blob, table, slot, threads=threads)
            assert acc.tolist() == keep.tolist(), (seed, threads)
            assert ju[acc].tobytes() == units.tobytes() and jf[acc].tobytes() == fy.tobytes(), (seed, threads)
            assert np.array_equal(bits(jr[acc]), bits(res)), (seed, threads)
            gone = np.setdiff1d(np.arange(len(pk)), acc)
            assert not ju[gone].tobytes().strip(b"\0")
    assert refused >= 18
