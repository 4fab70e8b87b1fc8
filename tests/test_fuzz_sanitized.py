"""Untrusted input: every host-side parser (packetisers, MP3 / Layer I-II / FLAC / Vorbis / AAC front-ends, plan + jobs, Vorbis mapping) built
with AddressSanitizer + UndefinedBehaviorSanitizer and driven with mutated streams (tests/cpp/fuzz_frontends.cpp).  Any
out-of-bounds access, signed overflow, misaligned access or leak aborts the driver.  A 60 000-input run was clean when this
was written; the suite runs a shorter one."""
import os
import subprocess

import numpy as np

from symphonia_b200 import workloads
from tests import _flac_bitstream as fw
from tests import _mp3_bitstream as bw
from tests import _mpa12_bitstream as b12
from tests import _aac_bitstream as ab
from tests import _streams as st
from tests import _vorbis_bitstream as vb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _seeds(rng):
    seeds = {}
    fr, _ = bw.gen_stream(rng, 12, version="1", mode=1, bitrate_idx=9)
    seeds["mp3_v1"] = st.mpa_tag_frame(rng, dict(version="1", layer=3, bitrate_idx=9, rate_idx=0, mode=1)) + b"".join(fr)
    seeds["mp3_v2"] = b"".join(bw.gen_stream(rng, 12, version="2", mode=3, bitrate_idx=6)[0])
    seeds["mp3_v25"] = b"".join(bw.gen_stream(rng, 12, version="2.5", mode=1, bitrate_idx=8)[0])
    seeds["mp2"] = b"".join(b12.gen_layer2_frame(rng, "1", 12, 0, 1, mode_ext=k % 4)[0] for k in range(8))
    seeds["mp1"] = b"".join(b12.gen_layer1_frame(rng, "1", 9, 1, 0)[0] for _ in range(8))
    seeds["adts"] = b"".join(st.adts_frame(rng, int(rng.integers(50, 300)), protected=bool(k % 2)) for k in range(20))
    ident = st.vorbis_ident(channels=2)
    setup, modes = st.vorbis_setup(rng, channels=2)
    pk = [ident, b"\x03vorbis" + bytes(30), setup] + [st.vorbis_audio_packet(rng, len(modes))[0] for _ in range(30)]
    seeds["ogg"] = b"".join(st.ogg_paginate(9, pk[:1], rng, eos=False) + st.ogg_paginate(9, pk[1:], rng, max_segments=20, first_sequence=1, bos=False))
    seeds["vsetup"] = setup
    seeds["vsetup_valid"] = st.vorbis_setup_valid(rng, channels=2)[0]
    frames, subs, samples = workloads.flac_batch(8, 576, seed=3, bps=16, channels=2)
    order = [f for f in range(8) if f % 7] + [0]
    fp = [fw.write_frame(rng, frames[f], subs[int(frames[f]["first_subframe"]):int(frames[f]["first_subframe"]) + 2], samples, k, stream_bps=16)
          for k, f in enumerate(order)]
    seeds["flac"] = fw.native_file(fp, fw.stream_info_block(576, 576, 44100, 2, 16, 0))
    seeds["flac_frame"] = fp[1]
    for k, rtype in enumerate((0, 1, 2)):
        vs = vb.Stream(np.random.default_rng(40 + k), residue_type=rtype)
        parts = [vs.ident, vs.setup] + [vs.packet()[0] for _ in range(10)]
        seeds[f"vorbis_fe{rtype}"] = b"VFE1" + b"".join(len(q).to_bytes(2, "little") + q for q in parts)
    for k in range(3):
        a = ab.Stream(np.random.default_rng(60 + k), rate=[44100, 48000, 8000][k], channels=2 - k % 2)
        parts = [a.packet()[0] for _ in range(8)]
        seeds[f"aac_fe{k}"] = b"AFE1" + bytes([k, 1 - k % 2]) + b"".join(len(q).to_bytes(2, "little") + q for q in parts)
    # hand-built edge packets (pulses that run to line 1024, the last valid scale-factor indices, 64 sections, TNS order 12)
    from tests.test_aac_frontend import _sce
    n_bands = 49
    edge = [_sce(150, n_bands, [(0, 30), (0, n_bands - 30)], pulse=(n_bands - 1, [(31, 3), (31, 2), (31, 1), (3, 7)])),
            _sce(255, 1, [(1, 1)], scf=[("d", 0)], spectral=[("1", 40)]), _sce(100, 2, [(15, 2)], scf=[("d", 60), ("d", 40)]),
            _sce(120, 1, [(0, 0)] * 63 + [(0, 1)]),
            _sce(120, 4, [(0, 4)], tns=[(1, 2), (0, 1), (10, 6), (12, 5), (0, 1), (0, 1)] + [(3, 3)] * 12)]
    seeds["aac_fe_edge"] = b"AFE1" + bytes([1, 0]) + b"".join(len(q).to_bytes(2, "little") + q for q in edge)
    return seeds


def test_parsers_under_address_and_ub_sanitizers(tmp_path):
    csrc = os.path.join(ROOT, "symphonia_b200", "csrc")
    exe = str(tmp_path / "fuzz_frontends")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-ffp-contract=off",
                           "-I/usr/local/cuda/include", "-o", exe, os.path.join(ROOT, "tests", "cpp", "fuzz_frontends.cpp")] +
                          [os.path.join(csrc, f) for f in ("mp3_frontend.cpp", "mpa12_frontend.cpp", "flac_frontend.cpp", "vorbis_frontend.cpp", "aac_frontend.cpp", "packetizer.cpp", "tables.cpp")])
    paths = []
    for name, blob in _seeds(np.random.default_rng(1)).items():
        path = str(tmp_path / (name + ".bin"))
        with open(path, "wb") as f:
            f.write(blob)
        paths.append(path)
    env = dict(os.environ, FUZZ_ITERS="250", ASAN_OPTIONS="detect_leaks=1:abort_on_error=1")
    res = subprocess.run([exe] + paths, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "no sanitizer report" in res.stdout and "4518 inputs" in res.stdout
