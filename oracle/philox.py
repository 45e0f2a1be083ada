"""TEST INFRASTRUCTURE — not part of the product path.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123),
restated with numpy. The reference itself draws from numpy PCG64 streams (ma_frozen_lake.py:260-261,
qlearning.py:120-142) which a counter-based GPU generator cannot reproduce; parity is therefore by INJECTION:
the same Philox words drive the CUDA kernels, the C oracle and (through oracle/ref_harness.py's stand-in rng
objects) the live reference code.

Draw layout for lockstep iteration t, global instance i, agent a:
    (w0, w1, w2, w3) = philox4x32_10(counter=(t & 0xffffffff, t >> 32, i, a), key=(seed_lo, seed_hi))
    w0 explore test  u = w0 / 2^32 < epsilon          (qlearning.py:120)
    w1 random action (w1 * 4) >> 32                    (qlearning.py:122)
    w2 tie-break     maxs[(w2 * len(maxs)) >> 32]      (qlearning.py:141-142)
    w3 slip outcome  cdf.searchsorted(w3 / 2^32, 'right')  (ma_frozen_lake.py:261 ; ma_office.py:378)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays of uint32 (broadcastable). Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))


def draws(seed, t, n_instances, n_agents, instance_offset=0):
    """uint32 [N, A, 4] draw words of lockstep iteration t."""
    i = (np.arange(n_instances, dtype=np.uint64) + np.uint64(instance_offset))[:, None]
    a = np.arange(n_agents, dtype=np.uint64)[None, :]
    i, a = np.broadcast_arrays(i, a)
    w = philox4x32_10(np.full(i.shape, t & 0xFFFFFFFF, dtype=np.uint64), np.full(i.shape, (t >> 32) & 0xFFFFFFFF, dtype=np.uint64),
                      i, a, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=-1)


# Random123 known-answer vectors (kat_vectors, philox4x32 10 rounds)
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]

if __name__ == "__main__":
    for ctr, key, want in KAT:
        got = tuple(int(x) for x in philox4x32_10(*ctr, *key))
        print([hex(g) for g in got], got == want)
