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
