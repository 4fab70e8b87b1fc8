"""The C++ host-side mirror of the reference's decoder plug-in interface (include/symgpu/decoder.hpp):
CPU tier checks registry tiers / error mapping / loud failure without a GPU; the GPU tier decodes a
stream packet by packet through AudioDecoder::decode (BASELINE config 0, "plumbing") and compares
with the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "decoder_host")


def _build():
    src = os.path.join(ROOT, "tests", "cpp", "decoder_host.cpp")
    lib = os.path.join(ROOT, "symphonia_b200", "libsymgpu.so")
    hdr = os.path.join(ROOT, "include", "symgpu", "decoder.hpp")
    hdr2 = os.path.join(ROOT, "include", "symgpu", "packetizer.hpp")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(lib), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-o", EXE, src, "-L" + os.path.dirname(lib), "-lsymgpu",
                               "-Wl,-rpath," + os.path.dirname(lib)])
    return EXE


def test_registry_and_errors_cpu():
    out = subprocess.run([_build(), "registry"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "registry: ok" in out.stdout


@pytest.mark.gpu
def test_packet_by_packet_decode_matches_oracle(tmp_path, oracle):
    from symphonia_b200 import workloads
    from tests import _oracle
    F = 24
    units, spectra, runs = workloads.mp3_batch(1, F, seed=321)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, 1)
    blob = b"".join(units[f].tobytes() + spectra[f].tobytes() for f in range(F))
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.write_bytes(blob)
    res = subprocess.run([_build(), "decode", str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    print(res.stdout.strip())
    got = np.frombuffer(outp.read_bytes(), dtype=np.float32).reshape(F, 2, 1152)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.gpu
def test_sixty_four_decoder_threads_share_one_context(tmp_path, oracle):
    """The server shape: 64 single-stream decoders on 64 threads, ONE context; every decode() is a submit + wait and
    the context batches whatever the threads have in flight.  Bit-exact per stream, and the launches really are shared."""
    import re
    from symphonia_b200 import workloads
    from tests import _oracle
    S, F = 64, 12
    units, spectra, runs = workloads.mp3_batch(S, F, seed=4242)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, S)
    assert rc == 0
    blob = b"".join(units[f].tobytes() + spectra[f].tobytes() for f in range(S * F))   # stream-major: frames of a run are consecutive
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.write_bytes(blob)
    res = subprocess.run([_build(), "threads", str(S), str(inp), str(outp)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    got = np.frombuffer(outp.read_bytes(), dtype=np.float32).reshape(S * F, 2, 1152)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    m = re.search(r"batches (\d+) frames (\d+)", res.stdout)
    assert m and int(m.group(2)) == S * F
    assert int(m.group(1)) < S * F, "some launches carried the packets of several decoders: " + res.stdout


def _build_nccl():
    src = os.path.join(ROOT, "tests", "cpp", "nccl_tables.cpp")
    exe = os.path.join(ROOT, "tests", "cpp", "nccl_tables")
    lib = os.path.join(ROOT, "symphonia_b200", "libsymgpu.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", "-I/usr/local/cuda/include", "-o", exe, src,
                               "-L" + os.path.dirname(lib), "-lsymgpu", "-L/usr/local/cuda/lib64", "-lcudart", "-ldl",
                               "-Wl,-rpath," + os.path.dirname(lib) + ":/usr/local/cuda/lib64"])
    return exe


@pytest.mark.gpu
def test_native_table_broadcast_over_nccl(tmp_path, oracle):
    """symgpu_tables_broadcast with a communicator made in C++ (ncclCommInitAll): two ranks when the box has two GPUs -- rank 1
    starts from zeroed tables and must decode bit-exactly afterwards -- else a 1-rank communicator over the same code path."""
    from symphonia_b200 import workloads
    from tests import _oracle
    F = 10
    units, spectra, runs = workloads.mp3_batch(1, F, seed=777)
    rc, want, _ = _oracle.mp3_batch(oracle, units, spectra, runs, 1)
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    inp.write_bytes(b"".join(units[f].tobytes() + spectra[f].tobytes() for f in range(F)))
    res = subprocess.run([_build_nccl(), str(inp), str(outp)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    got = np.frombuffer(outp.read_bytes(), dtype=np.float32).reshape(F, 2, 1152)
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), res.stdout
