#!/usr/bin/env python
"""Per-phase instruction / stall-sample breakdown of a kernel of mp3_kernel_v2.cu (or any .cu with `// PHASE:` markers)
from an ncu report.  usage: tools/phase_breakdown_v2.py <report.ncu-rep> [libsymgpu.so] [source.cu]
Joins the SASS listing of ncu's source page (per-instruction executed counts and stall samples) with nvdisasm -gi line info of
the same function in the .so; an instruction belongs to the phase whose marker precedes its OUTERMOST inlined-at line."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mangled_fragment(name):
    """'void symgpu::mp3v2_synth_kernel<(int)12, (int)9>(...)' -> 'mp3v2_synth_kernelILi12ELi9E'"""
    m = re.search(r"(\w+)<(.*?)>\(", name)
    if not m:
        return re.search(r"(\w+)\(", name).group(1)
    args = re.findall(r"\(int\)(\d+)|\(bool\)(\d)", m.group(2))
    frag = m.group(1) + "I" + "".join(f"Li{a}E" if a else f"Lb{b}E" for a, b in args)
    return frag


def lines_of(so, frag, base):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    out = []
    for cubin in sorted(os.listdir(tmp)):
        sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        cur, infn = 0, False
        for ln in sass.splitlines():
            if ".section" in ln and ".text." in ln:
                infn = frag in ln
            m = re.match(r'\s*//## File ".*%s", line (\d+)\s*$' % re.escape(base), ln)
            if m:
                cur = int(m.group(1))
            elif "//## File" in ln and "inlined at" not in ln:
                cur = -1
            if infn and re.match(r"^\s*/\*[0-9a-f]{4,6}\*/", ln):
                out.append((cur, ln.split("*/", 1)[1].strip()))
        if out:
            break
    return out


def main():
    rep = sys.argv[1]
    so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "symphonia_b200/libsymgpu.so")
    src = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "symphonia_b200/csrc/mp3_kernel_v2.cu")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "")
    lines = lines_of(so, mangled_fragment(name), os.path.basename(src))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, body = rows[hdr_i], rows[hdr_i + 1:]
    if len(body) != len(lines):
        print(f"instruction count mismatch: report {len(body)} vs cubin {len(lines)} (different build?)")
        return 1
    col = {n: hdr.index(n) for n in hdr}
    marks = [(i, m.group(1).strip()) for i, ln in enumerate(open(src).read().splitlines(), 1)
             for m in [re.search(r"// PHASE: (.+)$", ln)] if m]

    def phase(line):
        name = "?"
        for first, nm in marks:
            if line >= first:
                name = nm
        return name

    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for (line, _), r in zip(lines, body):
        a = agg.setdefault(phase(line), {"inst": 0, "samples": 0, **{s: 0 for s in stalls}})
        a["inst"] += int(r[col["Instructions Executed"]])
        a["samples"] += int(r[col["# Samples"]])
        for s in stalls:
            a[s] += int(r[col[s]])
    ti = sum(a["inst"] for a in agg.values())
    ts = sum(a["samples"] for a in agg.values())
    print(f"# {name}")
    print(f"{'phase':28s} {'inst':>10s} {'%':>5s} {'samples':>8s} {'%':>5s}  top stalls")
    for p, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
        top = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:5]
        tops = " ".join(f"{n}:{100 * v / max(a['samples'], 1):.0f}%" for v, n in top)
        print(f"{p:28s} {a['inst']:10d} {100 * a['inst'] / ti:5.1f} {a['samples']:8d} {100 * a['samples'] / ts:5.1f}  {tops}")
    print(f"{'total':28s} {ti:10d}       {ts:8d}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
