"""Host (pure Python / numpy) restatement of csrc/philox.cuh for the tests."""
import math

import numpy as np

M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c = [int(x) & M32 for x in ctr]
    k0, k1 = int(key[0]) & M32, int(key[1]) & M32
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c[3] ^ k1) & M32, p0 & M32]
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c


def philox_key(seed, chain, stream):
    return (seed + chain * 0x9E3779B97F4A7C15 + stream * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF


def normal4(seed, chain, idx4, step, stream):
    """The four N(0,1) draws of lane idx4 (float64 Box-Muller of the same uniforms)."""
    key = philox_key(seed, chain, stream)
    c = philox4x32_10([idx4 & M32, idx4 >> 32, step & M32, step >> 32], [key & M32, key >> 32])
    out = []
    for a, b in ((c[0], c[1]), (c[2], c[3])):
        u1 = float(np.float32(a)) * 2.3283064365386963e-10 + 2.3283064365386963e-10
        u2 = float(np.float32(b)) * 2.3283064365386963e-10
        r = math.sqrt(-2.0 * math.log(u1))
        out += [r * math.cos(2 * math.pi * u2), r * math.sin(2 * math.pi * u2)]
    return out
