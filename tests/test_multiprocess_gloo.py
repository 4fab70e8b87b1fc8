"""CPU tier, world_size 2 over gloo: the host-side logic of the multi-GPU path (stream sharding and
the table-blob broadcast -- the only collective this path has)."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["SYMGPU_ROOT"])
    from symphonia_b200 import sharding, workloads
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    blob = sharding.tables_blob()
    if rank != 0:
        blob[:] = 0                      # a rank whose host libm disagreed would be overwritten
    import torch
    t = torch.from_numpy(blob)
    dist.broadcast(t, src=0)
    ref = sharding.tables_blob()
    assert (t.numpy() == ref).all(), "broadcast blob differs from rank 0's tables"
    units, spectra, runs = workloads.mp3_batch(7, 3, seed=77)
    local_runs, idx = sharding.shard_runs(runs, rank, world)
    owned = sharding.shard_streams(7, rank, world)
    assert sorted(runs["stream"][(runs["stream"] % world) == rank].tolist()) == owned.tolist()
    assert len(idx) == 3 * len(owned) and (np.diff(local_runs["first_frame"]) == 3).all()
    assert (local_runs["stream"] == np.arange(len(owned))).all()
    # every frame of the batch is owned by exactly one rank
    mask = torch.zeros(len(units), dtype=torch.int32)
    mask[torch.from_numpy(idx)] = 1
    dist.all_reduce(mask)
    assert (mask == 1).all()
    # whole-job throughput aggregation used by bench.py: max of times, sum of work
    tt = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    assert tt.item() == float(world)
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_sharding_and_table_broadcast(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, SYMGPU_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2


FILES_WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, os.environ["SYMGPU_ROOT"])
    from symphonia_b200 import decode, sharding
    from tests.test_zz_many_files import _files
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    files = _files()                                  # the same corpus on every rank (seeded)
    mine = sharding.shard_streams(len(files), rank, world)
    plans, batches = decode.plan_files([files[i] for i in mine], threads=2)
    # every file is planned by exactly one rank, and the ranks' audio adds up to the corpus planned in one piece on rank 0
    owner = torch.zeros(len(files), dtype=torch.int32)
    owner[torch.from_numpy(mine)] = 1
    dist.all_reduce(owner)
    assert (owner == 1).all()
    frames = torch.tensor([sum(p["total_frames"] for p in plans)], dtype=torch.int64)
    dist.all_reduce(frames)
    if rank == 0:
        whole, _ = decode.plan_files(files, threads=2)
        assert frames.item() == sum(p["total_frames"] for p in whole)
        # a file's plan does not depend on which batch it is planned in
        for k, i in enumerate(mine):
            a, b = plans[k], whole[i]
            assert a["kind"] == b["kind"] and a["total_frames"] == b["total_frames"] and a["spans"].tobytes() == b["spans"].tobytes()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_two_rank_file_corpus(tmp_path):
    """BASELINE config 4's shape on the host side: a mixed corpus of files sharded by stream over two ranks, no data-path collective."""
    script = tmp_path / "files_worker.py"
    script.write_text(FILES_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, SYMGPU_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert out.stdout.count("ok") == 2
