"""numpy Philox4x32-10 — independent restatement used to check odeintr_fill_state_uniform()."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        with np.errstate(over="ignore"):
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) * (
        1.0 / 9007199254740992.0)


def uniform_state(instances, n, seed, lo, hi):
    """x[v, j] for global instance indices `instances` (any order) — same numbers as the device kernel."""
    g = np.asarray(instances, dtype=np.uint64)
    x = np.empty((n, g.size), dtype=np.float64)
    c0 = (g & MASK).astype(np.uint32)
    c1 = (g >> np.uint64(32)).astype(np.uint32)
    for v0 in range(0, n, 2):
        r = philox4x32_10(c0, c1, np.full(g.size, v0 // 2, np.uint32), np.zeros(g.size, np.uint32),
                          seed & 0xFFFFFFFF, seed >> 32)
        x[v0] = lo[v0] + (hi[v0] - lo[v0]) * u53(r[0], r[1])
        if v0 + 1 < n:
            x[v0 + 1] = lo[v0 + 1] + (hi[v0 + 1] - lo[v0 + 1]) * u53(r[2], r[3])
    return x
