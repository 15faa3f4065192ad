#!/usr/bin/env python3
"""Check the restatement of glibc's expf used by the sort-based X-ray (csrc/kernels.cu: expf_glibc) against this
machine's libm: the same table / coefficients / operation order in numpy float64.  Prints the mismatch counts of the
restatement and of the correctly rounded value (float(exp(double)))."""
import ctypes
import math
import struct
from decimal import Decimal, getcontext

import numpy as np

getcontext().prec = 60
N = 32
T = []
for i in range(N):
    bits = struct.unpack("<Q", struct.pack("<d", float(Decimal(2) ** (Decimal(i) / Decimal(N)))))[0]
    T.append((bits - (i << 47)) & 0xffffffffffffffff)
C0 = float.fromhex("0x1.c6af84b912394p-5") / N / N / N
C1 = float.fromhex("0x1.ebfce50fac4f3p-3") / N / N
C2 = float.fromhex("0x1.62e42ff0c52d6p-1") / N
INVLN2N = float.fromhex("0x1.71547652b82fep+0") * N
SHIFT = float.fromhex("0x1.8p+52")


def expf_restated(x32):
    z = np.float64(INVLN2N) * np.float64(x32)
    kd = z + SHIFT
    ki = struct.unpack("<Q", struct.pack("<d", float(kd)))[0]
    r = z - (kd - SHIFT)
    s = struct.unpack("<d", struct.pack("<Q", (T[ki % N] + (ki << 47)) & 0xffffffffffffffff))[0]
    y = (np.float64(C0) * r + np.float64(C1)) * (r * r) + (np.float64(C2) * r + 1.0)
    return np.float32(y * s)


if __name__ == "__main__":
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float; libm.expf.argtypes = [ctypes.c_float]
    xs = (-np.random.default_rng(0).uniform(0, 25, 200000)).astype(np.float32)
    bad = cr = 0
    for x in xs:
        g = np.float32(libm.expf(float(x))).view(np.uint32)
        bad += g != expf_restated(x).view(np.uint32)
        cr += g != np.float32(math.exp(float(x))).view(np.uint32)
    print(f"restated algorithm: {bad} mismatches of {len(xs)}; correctly rounded value: {cr} mismatches")
