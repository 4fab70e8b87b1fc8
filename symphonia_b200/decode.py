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


def ogg_vorbis_plan(data, serial=None):
    """CPU half for a Vorbis-in-Ogg file: pages -> packets (symgpu_ogg_index) -> identification / setup headers -> entropy front-end
    (symgpu_vorbis_fe_*) -> the synthesis stage's batch, plus the reader's time line: per-packet duration and leading discard
    (mappings/vorbis.rs:45-107) and the end trim against each page's granule position (symphonia-format-ogg/src/logical.rs:164-302;
    a page whose predecessor completed no packet starts at end - total duration, a stream whose audio sits on one page starts at
    -discard when that leaves padding).  Returns dict(stream, floors, units, floor_y, residue, runs, slot, spans, channels,
    sample_rate, total_frames).  Packets the front-end refuses are dropped, as a caller of the reference drops a DecodeError."""
    packets, pieces = packetizer.ogg_index(data)
    if len(packets) == 0:
        raise ValueError("no Ogg packets")
    serial = int(packets["serial"][0]) if serial is None else serial
    mine = packets[packets["serial"] == serial]
    blobs = [packetizer.gather(data, pk, pieces) for pk in mine]
    ident_b = blobs[0]
    ident = packetizer.vorbis_ident(ident_b)
    at = 1
    while at < len(blobs) and not (len(blobs[at]) >= 7 and blobs[at][0] == 5 and blobs[at][1:7] == b"vorbis"):
        at += 1
    if at == len(blobs):
        raise ValueError("no Vorbis setup header")
    setup_b = blobs[at]
    n_modes, mask = packetizer.vorbis_setup_modes(setup_b, ident)
    audio = [(pk, b) for pk, b in zip(mine[at + 1:], blobs[at + 1:]) if len(b) and (b[0] & 1) == 0]
    dur, discard, _ = packetizer.vorbis_packet_durations(ident, n_modes, mask, [b for _, b in audio])
    dur, discard = dur.astype(np.int64), discard.astype(np.int64)
    trim_end = packetizer.ogg_page_end_trims([int(pk["page_sequence"]) for pk, _ in audio], [int(pk["page_absgp"]) for pk, _ in audio],
                                             dur, discard).astype(np.int64)
    fe = frontend.VorbisFrontend(ident_b, setup_b)
    slot = fe.slot
    units, fy, res, keep = [], [], [], []
    for k, (_, b) in enumerate(audio):
        try:
            u, y, r = fe.decode(b)
        except frontend.SymgpuError:
            continue
        units.append(u), fy.append(y), res.append(r), keep.append(k)
    n = len(units)
    stream, floors = np.array([fe.stream], dtype=nat.VORBIS_STREAM_DTYPE), fe.floors.copy()
    fe.close()
    bs = {0: 1 << int(ident["bs0_exp"]), 1: 1 << int(ident["bs1_exp"])}
    spans = np.zeros(n, dtype=nat.PCM_SPAN_DTYPE)
    total = 0
    for o, k in enumerate(keep):
        frames = (bs[int(units[o]["prev_block_flag"])] + bs[int(units[o]["block_flag"])]) // 4
        if o == 0:
            ts, te = frames, 0  # the first packet after a reset is silenced in gapless mode (codec-vorbis lib.rs:318-322)
        else:
            ts = min(int(discard[k]), frames)
            te = min(int(trim_end[k]), frames - ts)
        spans[o] = (o * 2 * slot, slot, frames, ts, te, total)
        total += frames - ts - te
    runs = np.zeros(1, dtype=nat.VORBIS_RUN_DTYPE)
    runs["n_packets"] = n
    return dict(stream=stream, floors=floors, units=np.array(units, dtype=nat.VORBIS_UNIT_DTYPE).reshape(n),
                floor_y=np.array(fy, dtype=np.uint16).reshape(n, 2, 65), residue=np.array(res, dtype=np.float32).reshape(n, 2, slot),
                runs=runs, slot=slot, spans=spans, channels=int(ident["channels"]), sample_rate=int(ident["sample_rate"]), total_frames=total)


def decode_ogg_vorbis(engine, data, fmt=nat.FMT_S16, serial=None):
    """(samples [frames, channels] of `fmt`, sample_rate) of one Vorbis logical stream (mono / stereo, floor 1: what the synthesis
    kernel takes).  Registers the stream as slot 0 of `engine` with its floors from index 0."""
    plan = ogg_vorbis_plan(data, serial)
    if len(plan["units"]) == 0:
        return np.zeros((0, plan["channels"]), dtype=nat.FMT_NUMPY[fmt]), plan["sample_rate"]
    engine.vorbis_streams_set(plan["stream"])
    engine.vorbis_floors_set(plan["floors"])
    pcm = engine.vorbis_synth_host(plan["units"], plan["floor_y"], plan["residue"], plan["runs"], plan["slot"])
    return engine.pcm_pack_host(pcm, plan["spans"], plan["channels"], fmt, plan["total_frames"]), plan["sample_rate"]


def adts_aac_plan(data):
    """CPU half for an ADTS file: frames (symgpu_adts_index: header rules of adts.rs:130-309) -> raw_data_block payloads -> AAC-LC
    entropy front-end (symgpu_aac_fe_*) -> the synthesis stage's batch.  The reader gives every frame 1024 samples and trims
    nothing.  Returns dict(units [n,2], tns, coeffs [n,2,1024], runs, spans, channels, sample_rate, total_frames).  Packets the
    front-end refuses are dropped, as a caller of the reference drops a DecodeError; the stream's parameters are the first
    frame's (AdtsReader::try_new)."""
    packets, _ = packetizer.adts_index(data)
    if len(packets) == 0:
        raise ValueError("no ADTS frames")
    rate, channels = int(packets[0]["sample_rate"]), int(packets[0]["channels"])
    if channels not in (1, 2):
        raise ValueError("channel configuration outside AAC-LC mono / stereo")
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    fe = frontend.AacFrontend(rate, channels)
    units, tns, coeffs, n_tns = [], [], [], 0
    for pk in packets:
        try:
            u, t, c = fe.decode(buf[int(pk["offset"]):int(pk["offset"]) + int(pk["size"])].tobytes(), tns_base=n_tns)
        except frontend.SymgpuError:
            continue
        units.append(u), tns.append(t), coeffs.append(c)
        n_tns += len(t)
    fe.close()
    n = len(units)
    runs = np.zeros(1, dtype=nat.AAC_RUN_DTYPE)
    runs[0]["n_frames"], runs[0]["channels"] = n, channels
    spans = np.zeros(n, dtype=nat.PCM_SPAN_DTYPE)
    spans["src"] = np.arange(n, dtype=np.uint64) * 2048
    spans["plane_stride"], spans["frames"] = 1024, 1024
    spans["dst_frame"] = np.arange(n, dtype=np.uint64) * 1024
    return dict(units=np.array(units, dtype=nat.AAC_UNIT_DTYPE).reshape(n, 2), tns=np.concatenate(tns) if n else np.zeros(0, dtype=nat.AAC_TNS_DTYPE),
                coeffs=np.array(coeffs, dtype=np.float32).reshape(n, 2, 1024), runs=runs, spans=spans, channels=channels, sample_rate=rate,
                total_frames=1024 * n)


def decode_adts_aac(engine, data, fmt=nat.FMT_S16, stream=0):
    """(samples [frames, channels] of `fmt`, sample_rate) of an ADTS AAC-LC file; stream slot `stream` of `engine` is reset first."""
    plan = adts_aac_plan(data)
    if len(plan["units"]) == 0:
        return np.zeros((0, plan["channels"]), dtype=nat.FMT_NUMPY[fmt]), plan["sample_rate"]
    plan["runs"]["stream"] = stream
    engine.aac_stream_reset(stream)
    pcm = engine.aac_synth_host(plan["units"], plan["tns"], plan["coeffs"], plan["runs"])
    return engine.pcm_pack_host(pcm, plan["spans"], plan["channels"], fmt, plan["total_frames"]), plan["sample_rate"]
