"""numpy dtypes mirroring the AAC / Vorbis structs of include/symgpu.h + f64 reference transforms
used by the oracle known-answer tests."""
import numpy as np

AAC_UNIT = np.dtype([("window_sequence", "u1"), ("window_shape", "u1"), ("prev_window_shape", "u1"), ("n_tns", "u1"),
                     ("tns_first", "<u4"), ("reserved", "<u4", (2,))])
AAC_TNS = np.dtype([("start", "<u2"), ("end", "<u2"), ("order", "u1"), ("direction", "u1"), ("reserved", "<u2"),
                    ("lpc", "<f4", (20,))])
AAC_RUN = np.dtype([("stream", "<u4"), ("first_frame", "<u4"), ("n_frames", "<u4"), ("channels", "u1"),
                    ("reserved", "u1", (3,))])
VORBIS_FLOOR1 = np.dtype([("multiplier", "u1"), ("n_posts", "u1"), ("x_list", "<u2", (65,)), ("low", "u1", (65,)),
                          ("high", "u1", (65,)), ("sort_order", "u1", (65,)), ("reserved", "u1", (5,))])
VORBIS_STREAM = np.dtype([("bs0_exp", "u1"), ("bs1_exp", "u1"), ("channels", "u1"), ("coupled", "u1")])
VORBIS_UNIT = np.dtype([("block_flag", "u1"), ("prev_block_flag", "u1"), ("do_not_decode", "u1", (2,)),
                        ("floor", "<u2", (2,)), ("reserved", "u1", (8,))])
VORBIS_RUN = np.dtype([("stream", "<u4"), ("first_packet", "<u4"), ("n_packets", "<u4"), ("reserved", "<u4")])
assert AAC_UNIT.itemsize == 16 and AAC_TNS.itemsize == 88 and AAC_RUN.itemsize == 16
assert VORBIS_FLOOR1.itemsize == 332 and VORBIS_STREAM.itemsize == 4 and VORBIS_UNIT.itemsize == 16


def mdct_forward(block, n):
    """X[j] = sum_i block[i] cos(pi/(4n) (2i+1+n)(2j+1)), block of 2n samples -> n coefficients (f64)."""
    i = np.arange(2 * n)[:, None]
    j = np.arange(n)[None, :]
    return (block[:, None] * np.cos(np.pi / (4 * n) * ((2 * i + 1 + n) * (2 * j + 1)))).sum(axis=0)


def find_neighbors(x_list, i):
    """floor.rs:748-773 (low_neighbor / high_neighbor of the Vorbis I spec, 9.2.4-9.2.5)."""
    bound = x_list[i]
    low, high = 0, 0xFFFFFFFF
    res = [0, 0]
    for k in range(i):
        xv = x_list[k]
        if low < xv < bound:
            low, res[0] = xv, k
        if bound < xv < high:
            high, res[1] = xv, k
    return res


def make_floor1_setup(x_list, multiplier):
    s = np.zeros((), dtype=VORBIS_FLOOR1)
    n = len(x_list)
    s["multiplier"] = multiplier
    s["n_posts"] = n
    s["x_list"][:n] = x_list
    for i in range(n):
        lo, hi = find_neighbors(list(x_list), i)
        s["low"][i], s["high"][i] = lo, hi
    s["sort_order"][:n] = sorted(range(n), key=lambda k: x_list[k])
    return s
