"""numpy dtypes mirroring the AAC / Vorbis structs of include/symgpu.h + f64 reference transforms
used by the oracle known-answer tests."""
import numpy as np

from symphonia_b200._native import (AAC_RUN_DTYPE as AAC_RUN, AAC_TNS_DTYPE as AAC_TNS, AAC_UNIT_DTYPE as AAC_UNIT,
                                     VORBIS_FLOOR1_DTYPE as VORBIS_FLOOR1, VORBIS_RUN_DTYPE as VORBIS_RUN,
                                     VORBIS_STREAM_DTYPE as VORBIS_STREAM, VORBIS_UNIT_DTYPE as VORBIS_UNIT)
from symphonia_b200.workloads import find_neighbors, make_floor1_setup  # noqa: F401


def mdct_forward(block, n):
    """X[j] = sum_i block[i] cos(pi/(4n) (2i+1+n)(2j+1)), block of 2n samples -> n coefficients (f64)."""
    i = np.arange(2 * n)[:, None]
    j = np.arange(n)[None, :]
    return (block[:, None] * np.cos(np.pi / (4 * n) * ((2 * i + 1 + n) * (2 * j + 1)))).sum(axis=0)


