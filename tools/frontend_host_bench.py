#!/usr/bin/env python3
"""Host-side throughput of the packetisers and the MP3 entropy front-end: no GPU involved.
Synthetic streams from the test writers (dense frames: 85-100 % of every frame's bit budget used).  One JSON line per
case; `python tools/frontend_host_bench.py > profiles/<round>_frontend_host.jsonl`."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from symphonia_b200 import frontend, packetizer  # noqa: E402
from tests import _mp3_bitstream as bw  # noqa: E402
from tests import _streams as st  # noqa: E402

RATES = [44100, 48000, 32000, 22050, 24000, 16000, 11025, 12000, 8000]


def best(fn, reps=5):
    t = 1e9
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t = min(t, time.perf_counter() - t0)
    return t, out


def main():
    rng = np.random.default_rng(5)
    cores = os.cpu_count()
    for label, kw in (("mp3 128k joint stereo 44.1k", dict(version="1", mode=1, bitrate_idx=9)),
                      ("mp3 320k stereo 44.1k", dict(version="1", mode=0, bitrate_idx=14)),
                      ("mp3 64k mono 22.05k (MPEG-2)", dict(version="2", mode=3, bitrate_idx=8))):
        frames, _ = bw.gen_stream(rng, 200, fill=(0.85, 1.0), **kw)
        data = b"".join(frames) * 50
        t_index, (track, packets) = best(lambda: packetizer.mpa_index(data))
        fe = frontend.Mp3Frontend()

        def serial():
            fe.reset()
            return fe.decode_packets(data, packets)
        t_serial, (u, q, frame_of, info) = best(serial, 3)
        t_plan, (md, jobs, _, _) = best(lambda: frontend.entropy_plan(data, packets))
        t_jobs, _ = best(lambda: frontend.entropy_run_cpu(md, jobs), 3)
        n = len(frame_of)
        spf = 1152 if kw["version"] == "1" else 576
        sec = n * spf / RATES[int(info["sample_rate_idx"])]
        print(json.dumps({"case": label, "frames": n, "file_bytes": len(data), "host_cores_total": cores, "threads_used": 1,
                          "index_frames_per_s": n / t_index, "serial_frontend_frames_per_s": n / t_serial, "serial_frontend_audio_s_per_s": sec / t_serial,
                          "plan_frames_per_s": n / t_plan, "plan_audio_s_per_s": sec / t_plan, "jobs_cpu_frames_per_s": n / t_jobs,
                          "bytes_per_frame": {"file": len(data) / n, "main_data": md.size / n, "jobs": 256, "quant_i16": 4608, "spectra_f32": 9216},
                          "nonzero_lines_per_frame": float(np.count_nonzero(q)) / n}))
    exe = os.path.join(ROOT, "tests", "cpp", "packetizer_host")
    if os.path.exists(exe):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            ad = b"".join(st.adts_frame(rng, int(rng.integers(250, 450))) for _ in range(2000)) * 40
            pk = [rng.integers(0, 256, int(rng.integers(30, 600)), dtype=np.uint8).tobytes() for _ in range(60000)]
            og = b"".join(st.ogg_paginate(3, pk, np.random.default_rng(2), max_segments=40))
            for mode, blob in (("bench-adts", ad), ("bench-ogg", og)):
                path = os.path.join(tmp, "x.bin")
                with open(path, "wb") as f:
                    f.write(blob)
                line = subprocess.run([exe, mode, path, "10"], capture_output=True, text=True).stdout.strip()
                print(json.dumps({"case": mode, "result": line}))


if __name__ == "__main__":
    main()
