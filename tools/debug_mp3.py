import sys, numpy as np
sys.path.insert(0, '.')
import symphonia_b200 as sb
from symphonia_b200 import workloads
from tests import _oracle
orc = _oracle.load()
eng = sb.Engine(0)
def case(S, F, **kw):
    units, spectra, runs = workloads.mp3_batch(S, F, **kw)
    rc, want, _ = _oracle.mp3_batch(orc, units, spectra, runs, S)
    eng.mp3_streams_alloc(S)
    got = eng.mp3_synth_host(units, spectra, runs)
    g = got.view(np.uint32); w = want.view(np.uint32)
    bad = (g != w)
    print(f"case S={S} F={F} {kw}: bad {bad.sum()} / {bad.size}; maxabs diff {np.nanmax(np.abs(got-want)):.3e}  max |want| {np.abs(want).max():.3e}")
    if bad.sum():
        b = bad.reshape(S*F, 2, 2, 18, 32)   # frame, ch, gr, slot, i
        per = b.sum(axis=(3,4))
        for f in range(min(S*F, 12)):
            print("  frame", f, "bad per (ch,gr):", per[f].tolist(), "bt", units['block_type'][f].tolist(), "flags", units['flags'][f].tolist())
        f, c, gq, sl, i = [x[0] for x in np.nonzero(b)]
        print("  first bad: frame", f, "ch", c, "gr", gq, "slot", sl, "i", i, got.reshape(S*F,2,2,18,32)[f,c,gq,sl,i], want.reshape(S*F,2,2,18,32)[f,c,gq,sl,i])
        print("  bad per slot (frame %d ch %d gr %d):" % (f,c,gq), b[f,c,gq].sum(axis=1).tolist())
        print("  bad per i    :", b[f,c,gq].sum(axis=0).tolist())
case(1, 1, seed=1, joint=False, block_switching=False)
case(1, 3, seed=1, joint=False, block_switching=False)
case(1, 12, seed=1, joint=False, block_switching=False)
case(2, 6, seed=2, joint=True, block_switching=False)
case(2, 6, seed=3, joint=False, block_switching=True)
case(2, 6, seed=4)
