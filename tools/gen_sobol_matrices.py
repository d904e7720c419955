"""Generate tungsten_b200/data/sobol_1024x32.u32: Sobol' direction matrices for 1024 dimensions.

The reference samples its Sobol' points with Joe & Kuo's D(6) direction numbers
(new-joe-kuo-6.21201; reference: src/thirdparty/sobol/sobol.cpp:36, 1024 dims x 52 columns of
uint32, of which only columns 0..31 can be reached because SobolPathSampler's index is a uint32,
src/core/sampling/SobolPathSampler.hpp:20-23,67-72).  scipy ships the same Joe-Kuo table, so the
matrices are regenerated here from scipy instead of being copied out of the reference;
tests/test_sobol_table.py checks the blob's checksum and (when /root/reference is mounted) that it
equals the reference's table column for column.

Layout: little-endian uint32 [1024][32]; entry [d][i] is XORed into the result when bit i of the
sample index is set.
"""
import hashlib
import os
import sys

import numpy as np
from scipy.stats import qmc


def generate():
    s = qmc.Sobol(d=1024, scramble=False, bits=32)
    v = np.ascontiguousarray(s._sv, dtype="<u4")
    assert v.shape == (1024, 32)
    return v


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "tungsten_b200", "data", "sobol_1024x32.u32")
    v = generate()
    v.tofile(out)
    print(out, v.nbytes, "bytes sha256", hashlib.sha256(v.tobytes()).hexdigest())
