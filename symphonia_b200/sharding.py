"""Multi-GPU host logic: the path shards by STREAM, with no data-path collective.

  * shard_streams      stream i -> rank i mod world, keeps per-stream state GPU-local
  * shard_runs         the runs of a batch that belong to this rank, re-based onto a compact local batch
  * broadcast_tables   the one collective of the path: rank 0's host-built table blob to every rank
"""
import ctypes

import numpy as np

from . import _native


def shard_streams(n_streams, rank, world):
    """Indices of the streams rank `rank` owns."""
    return np.arange(rank, n_streams, world, dtype=np.int64)


def shard_runs(runs, rank, world, frames_key="n_frames", first_key="first_frame"):
    """Selects the runs whose stream belongs to `rank` and lays their frames out contiguously.

    Returns (local_runs, frame_index): frame_index[k] is the global frame/packet index of local frame k, so
    `units[frame_index]`, `spectra[frame_index]` are the rank's compact inputs."""
    runs = np.asarray(runs)
    mine = runs[(runs["stream"] % world) == rank].copy()
    idx = []
    pos = 0
    for r in mine:
        n = int(r[frames_key])
        idx.append(np.arange(int(r[first_key]), int(r[first_key]) + n, dtype=np.int64))
        r[first_key] = pos
        pos += n
    mine["stream"] = mine["stream"] // world      # local state slot
    return mine, (np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64))


def tables_blob():
    lib = _native.lib()
    n = lib.symgpu_tables_host_blob(None, 0)
    blob = np.zeros(n, dtype=np.uint8)
    lib.symgpu_tables_host_blob(blob.ctypes.data_as(ctypes.c_void_p), n)
    return blob


def broadcast_tables(dist, device=None, src=0):
    """Broadcasts rank `src`'s table blob (torch.distributed, NCCL on GPUs / gloo on CPU) and returns it."""
    import torch
    blob = tables_blob()
    t = torch.from_numpy(blob)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
