"""Loader for liboracle.so (ORACLE = test infrastructure; never imported by the product).

Builds it with oracle/Makefile when the prebuilt file is missing (gcc is present both in the build
container and on the GPU box).
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

c_f32p = ctypes.POINTER(ctypes.c_float)


class Mp3State(ctypes.Structure):
    _fields_ = [("overlap", ctypes.c_float * (2 * 32 * 18)),
                ("v_vec", ctypes.c_float * (2 * 16 * 64)),
                ("v_front", ctypes.c_int32 * 2)]


class AacState(ctypes.Structure):
    _fields_ = [("delay", ctypes.c_float * 1024)]


class VorbisState(ctypes.Structure):
    _fields_ = [("overlap", ctypes.c_float * (2 * 4096))]


class VorbisMcState(ctypes.Structure):
    _fields_ = [("overlap", ctypes.c_float * (8 * 4096))]


def build(arch=None, out=None):
    cmd = ["make", "-C", ODIR, "-s"]
    if arch:
        cmd.append("ARCH=" + arch)
    if out:
        cmd.append("OUT=" + out)
    subprocess.check_call(cmd)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def load(path=None):
    path = path or os.path.join(ODIR, "_build", "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    lib.oracle_mp3_tables.restype = ctypes.c_size_t
    lib.oracle_mp3_tables.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.oracle_mp3_pow43.restype = ctypes.c_float
    lib.oracle_mp3_imdct_window.restype = c_f32p
    lib.oracle_mp3_batch.restype = ctypes.c_int
    lib.oracle_mp3_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_void_p]
    lib.oracle_mp3_batch_mt.restype = ctypes.c_int
    lib.oracle_mp3_batch_mt.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
    lib.oracle_mp3_frame.restype = ctypes.c_int
    lib.oracle_mp3_frame.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int, ctypes.c_int]
    lib.oracle_mp3_polyphase.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p]
    lib.oracle_imdct.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    lib.oracle_fft_inplace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.oracle_aac_synth.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p]
    lib.oracle_aac_tns.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    lib.oracle_aac_batch.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int]
    lib.oracle_aac_window.restype = c_f32p
    lib.oracle_vorbis_floor1.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    lib.oracle_vorbis_batch.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p,
                                                                ctypes.c_int]
    lib.oracle_vorbis_window.restype = c_f32p
    lib.oracle_vorbis_inverse_db.restype = ctypes.c_float
    lib.oracle_conv_s16.restype = ctypes.c_int16
    lib.oracle_conv_s24.restype = ctypes.c_int32
    lib.oracle_conv_s32.restype = ctypes.c_int32
    lib.oracle_conv_u8.restype = ctypes.c_uint8
    for fn in (lib.oracle_conv_s16, lib.oracle_conv_s24, lib.oracle_conv_s32, lib.oracle_conv_u8):
        fn.argtypes = [ctypes.c_float]
    lib.oracle_pcm_pack.restype = ctypes.c_int
    lib.oracle_pcm_pack.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_uint32] * 4 + [ctypes.c_int,
                                                                                                 ctypes.c_void_p]
    return lib


def pcm_pack(lib, pcm, spans, channels, fmt, out_frames, plane_stride=0, frames=0, n_spans=None):
    """oracle_pcm_pack with the same calling convention as Engine.pcm_pack_host."""
    from symphonia_b200._native import FMT_NUMPY, PCM_SPAN_DTYPE
    pcm = np.ascontiguousarray(pcm, dtype=np.float32)
    if spans is not None:
        spans = np.ascontiguousarray(spans, dtype=PCM_SPAN_DTYPE)
        n_spans = len(spans)
    out = np.zeros((out_frames, channels), dtype=FMT_NUMPY[fmt])
    rc = lib.oracle_pcm_pack(ptr(pcm), ptr(spans) if spans is not None else None, n_spans, channels, plane_stride,
                             frames, fmt, ptr(out))
    assert rc == 0
    return out


def mpa12_batch(lib, subbands, runs, n_streams, states=None):
    """oracle_mpa12_batch with fresh (or the given) per-stream synthesis state."""
    if states is None:
        states = (Mp3State * n_streams)()
    subbands = np.ascontiguousarray(subbands, dtype=np.float32)
    runs = np.ascontiguousarray(runs)
    pcm = np.zeros((subbands.shape[0], 2, 1152), dtype=np.float32)
    lib.oracle_mpa12_batch.restype = ctypes.c_int
    lib.oracle_mpa12_batch.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    rc = lib.oracle_mpa12_batch(ctypes.byref(states), ptr(subbands), ptr(runs), len(runs), subbands.shape[-1], ptr(pcm))
    return rc, pcm, states


def mp3_batch(lib, units, spectra, runs, n_streams):
    """Runs the oracle over a batch laid out as symgpu_mp3_synth_host expects; fresh stream state."""
    n_frames = spectra.shape[0]
    states = (Mp3State * n_streams)()
    pcm = np.zeros((n_frames, 2, 1152), dtype=np.float32)
    units = np.ascontiguousarray(units)
    spectra = np.ascontiguousarray(spectra, dtype=np.float32)
    runs = np.ascontiguousarray(runs)
    rc = lib.oracle_mp3_batch(ctypes.byref(states), ptr(units), ptr(spectra), ptr(runs),
                              ctypes.c_uint32(len(runs)), ptr(pcm))
    return rc, pcm, states


def aac_batch(lib, units, tns, coeffs, runs, n_streams, n_threads=1):
    n_frames = coeffs.shape[0]
    states = (AacState * (2 * n_streams))()
    pcm = np.zeros((n_frames, 2, 1024), dtype=np.float32)
    units, tns, runs = np.ascontiguousarray(units), np.ascontiguousarray(tns), np.ascontiguousarray(runs)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float32)
    rc = lib.oracle_aac_batch(ctypes.byref(states), ptr(units), ptr(tns) if len(tns) else None, ptr(coeffs), ptr(runs),
                              ctypes.c_uint32(len(runs)), ptr(pcm), n_threads)
    return rc, pcm


def vorbis_batch(lib, wl, n_threads=1):
    """wl: the dict returned by symphonia_b200.workloads.vorbis_batch."""
    n_streams = len(wl["streams"])
    states = (VorbisState * n_streams)()
    pcm = np.zeros((len(wl["units"]), 2, wl["slot"]), dtype=np.float32)
    arrs = {k: np.ascontiguousarray(wl[k]) for k in ("streams", "floors", "units", "floor_y", "residue", "runs")}
    rc = lib.oracle_vorbis_batch(ctypes.byref(states), ptr(arrs["streams"]), ptr(arrs["floors"]), ptr(arrs["units"]),
                                 ptr(arrs["floor_y"]), ptr(arrs["residue"]), ptr(arrs["runs"]),
                                 ctypes.c_uint32(len(arrs["runs"])), ctypes.c_uint32(wl["slot"]), ptr(pcm), n_threads)
    return rc, pcm


def vorbis_mc_batch(lib, wl):
    """wl: the dict returned by symphonia_b200.workloads.vorbis_mc_batch."""
    states = (VorbisMcState * len(wl["streams"]))()
    C = int(wl["channels"])
    pcm = np.zeros((len(wl["units"]), C, wl["slot"]), dtype=np.float32)
    arrs = {k: np.ascontiguousarray(wl[k]) for k in ("streams", "floors", "units", "floor_y", "residue", "runs")}
    lib.oracle_vorbis_mc_batch.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_uint32] * 3 + [ctypes.c_void_p]
    rc = lib.oracle_vorbis_mc_batch(ctypes.byref(states), ptr(arrs["streams"]), ptr(arrs["floors"]), ptr(arrs["units"]), ptr(arrs["floor_y"]),
                                    ptr(arrs["residue"]), ptr(arrs["runs"]), len(arrs["runs"]), C, int(wl["slot"]), ptr(pcm))
    return rc, pcm
