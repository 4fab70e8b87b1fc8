#!/usr/bin/env python3
"""End to end from FILE BYTES (SURVEY §8f N1 + N2 in front of the synthesis kernel), BASELINE shape: 64 streams x 128
frames of 128 kbit/s joint stereo.  Two ways to the same PCM, host wall clock around the synchronous host calls:

  device front-end   Engine.mp3_decode_files_host: side-information pass on the CPU, Huffman decode + synthesis on the GPU
  CPU front-end      Mp3Frontend.decode_packets per stream (one core here), then Engine.mp3_synth_host_quantized

NOT part of the driver contract (bench.py is); written in round 1 for the first GPU call of round 2.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import symphonia_b200 as sb  # noqa: E402
from symphonia_b200 import _native as nat  # noqa: E402
from symphonia_b200 import frontend, packetizer  # noqa: E402
from tests import _mp3_bitstream as bw  # noqa: E402


def main():
    S, F = 64, 128
    rng = np.random.default_rng(7)
    # a handful of distinct streams, reused round-robin (the Python writer is slow; every stream slot still has its own state)
    base = []
    for k in range(4):
        frames, _ = bw.gen_stream(rng, F, version="1", mode=1, bitrate_idx=9, fill=(0.85, 1.0), pair_blocks=True)
        data = b"".join(frames)
        base.append((data, packetizer.mpa_index(data)[1]))
    files = [(base[s % 4][0], base[s % 4][1], s) for s in range(S)]
    n_bytes = sum(len(d) for d, _, _ in files)
    out = {"workload": f"MP3 128 kbit/s joint stereo, {S} streams x {F} frames from file bytes", "file_bytes": n_bytes}
    with sb.Engine(0) as eng:
        eng.mp3_streams_alloc(S)
        # ---- CPU front-end + quantised entry
        t0 = time.perf_counter()
        units, quant, runs, at = [], [], np.zeros(S, dtype=nat.MP3_RUN_DTYPE), 0
        for s, (data, packets, slot) in enumerate(files):
            u, q, frame_of, info = frontend.Mp3Frontend().decode_packets(data, packets)
            units.append(u), quant.append(q)
            runs[s] = (slot, at, len(u), int(info["granules"]), int(info["channels"]), 0)
            at += len(u)
        units, quant = np.concatenate(units).reshape(-1), np.concatenate(quant)
        t_fe = time.perf_counter() - t0
        ref = eng.mp3_synth_host_quantized(units, quant, runs)
        best = 1e9
        for _ in range(10):
            eng.mp3_streams_alloc(S)
            t = time.perf_counter()
            ref = eng.mp3_synth_host_quantized(units, quant, runs)
            best = min(best, time.perf_counter() - t)
        out["cpu_front_end_s_one_core"] = t_fe
        out["synth_host_quantized_ms"] = best * 1e3
        # ---- device front-end
        try:
            eng.mp3_streams_alloc(S)
            got, good, frame_of, rounds = eng.mp3_decode_files_host(files)
            out["device_path_bit_identical"] = bool((got.view(np.uint32) == ref.view(np.uint32)).all())
            best = 1e9
            for _ in range(10):
                eng.mp3_streams_alloc(S)
                t = time.perf_counter()
                eng.mp3_decode_files_host(files)
                best = min(best, time.perf_counter() - t)
            out["decode_files_host_ms"] = best * 1e3
            out["rounds"] = rounds
            audio = at * 1152 / 44100.0
            out["e2e_audio_s_per_s_device_front_end"] = audio / best
            out["e2e_audio_s_per_s_cpu_front_end_one_core"] = audio / (t_fe + out["synth_host_quantized_ms"] * 1e-3)
        except Exception as e:  # the device path is experimental: report, do not hide the other numbers
            out["device_path_error"] = repr(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
