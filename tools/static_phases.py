#!/usr/bin/env python
"""Static per-phase SASS instruction counts of a .cu file compiled to a cubin (no GPU needed).
usage: tools/static_phases.py <file.cubin> <source.cu> [kernel-substring]
Buckets every SASS instruction by the `// PHASE: name` marker that precedes its source line (nvdisasm -gi, outermost frame)."""
import collections
import re
import subprocess
import sys


def main():
    cubin, src = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else None
    marks = []
    for i, ln in enumerate(open(src).read().splitlines(), 1):
        m = re.search(r"// PHASE: (.+)$", ln)
        if m:
            marks.append((i, m.group(1).strip()))
    sass = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    base = src.split("/")[-1]
    cur, infn = 0, want is None
    counts = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for ln in sass.splitlines():
        if ".section" in ln and ".text." in ln:
            infn = want is None or want in ln
        # with -gi an instruction is preceded by its inline chain, innermost first: the LAST line of the group is
        # the outermost frame (a line of the kernel body), which is what the phase markers bracket
        m = re.match(r'\s*//## File ".*%s", line (\d+)\s*$' % re.escape(base), ln)
        if m:
            cur = int(m.group(1))
        elif "//## File" in ln and "inlined at" not in ln:
            cur = -1
        if infn and re.match(r"^\s*/\*[0-9a-f]{4,6}\*/", ln):
            phase = "?"
            for first, name in marks:
                if first <= cur:
                    phase = name
            body = ln.split("*/", 1)[1].strip()
            op = body.split()[1] if body.startswith("@") else body.split()[0]
            op = op.split(".")[0].rstrip(";")
            counts[phase] += 1
            ops[phase][op] += 1
    total = sum(counts.values())
    for phase, n in counts.most_common():
        top = " ".join(f"{k}:{v}" for k, v in ops[phase].most_common(8))
        print(f"{phase:28s} {n:6d} {100.0 * n / total:5.1f}%  {top}")
    print(f"{'total':28s} {total:6d}  ({total * 16 / 1024:.1f} KB)")


if __name__ == "__main__":
    main()
