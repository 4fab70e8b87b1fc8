"""File bytes in, gapless interleaved samples out: the public call for MPEG audio files (what a user of the reference does with
`MpaReader` + `MpaDecoder` + `copy_to_slice_interleaved`).  Packetiser and entropy front-end on the CPU,
synthesis and the output stage (N3) on the GPU.  Every device entry point used here is part of the round-1 GPU parity suite."""
import numpy as np

from . import _native as nat
from . import frontend, packetizer


def mpeg_audio_plan(data):
    """CPU half: (kind, payload, runs, spans, sample_rate, channels, total_frames).  kind 3: payload = (units, quant); kind 1 / 2:
    payload = sub-band samples.  spans carry the packetiser's trims (encoder delay / padding from a LAME tag, or the end trim of
    an extrapolated length) and where every packet's surviving frames go in the output."""
    track, packets = packetizer.mpa_index(data)
    layer = int(track["layer"])
    if layer == 3:
        units, quant, frame_of, info = frontend.Mp3Frontend().decode_packets(data, packets)
        payload = (units.reshape(-1), quant)
        n, per = len(units), 1152 if int(info["granules"]) == 2 else 576
        runs = np.zeros(1, dtype=nat.MP3_RUN_DTYPE)
        runs[0] = (0, 0, n, int(info["granules"]), int(info["channels"]), 0)
    else:
        sub, frame_of, info = frontend.mpa12_decode_packets(data, packets, layer)
        payload = sub
        n, per = len(sub), 32 * sub.shape[-1]
        runs = np.zeros(1, dtype=nat.MPA12_RUN_DTYPE)
        runs[0] = (0, 0, n, int(info["channels"]), (0, 0, 0))
    kept = packets[frame_of]
    spans = np.zeros(n, dtype=nat.PCM_SPAN_DTYPE)
    spans["src"] = np.arange(n, dtype=np.uint64) * 2304          # every frame slot holds 2 planes of 1152 floats
    spans["plane_stride"], spans["frames"] = 1152, per
    spans["trim_start"] = np.minimum(kept["trim_start"], per)
    spans["trim_end"] = np.minimum(kept["trim_end"], per - spans["trim_start"])
    left = per - spans["trim_start"].astype(np.int64) - spans["trim_end"].astype(np.int64)
    spans["dst_frame"] = np.concatenate([[0], np.cumsum(left)[:-1]]).astype(np.uint64) if n else 0
    return layer, payload, runs, spans, int(info["sample_rate"]) if n else int(track["sample_rate"]), int(info["channels"]) if n else int(track["channels"]), int(left.sum())


def decode_mpeg_audio(engine, data, fmt=nat.FMT_S16, stream=0):
    """(samples [frames, channels] of `fmt`, sample_rate).  Layers I-III; one stream slot of `engine` is used and reset first."""
    layer, payload, runs, spans, rate, channels, total = mpeg_audio_plan(data)
    runs["stream"] = stream
    if len(spans) == 0:
        return np.zeros((0, channels), dtype=nat.FMT_NUMPY[fmt]), rate
    engine.mp3_stream_reset(stream)
    pcm = engine.mp3_synth_host_quantized(payload[0], payload[1], runs) if layer == 3 else engine.mpa12_synth_host(payload, runs)
    return engine.pcm_pack_host(pcm, spans, channels, fmt, total), rate
