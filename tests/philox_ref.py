"""Test-only numpy restatement of Philox4x32-10 (Salmon et al., SC'11) and of the sampler's uniform stream
(art_planner_b200/csrc/artp_sampler.cuh: counter = (idx_lo, idx_hi, block, "ARTP"), key = seed)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
TAG = 0x41525450


def philox4x32_10(ctr, key):
    """ctr: [n, 4] uint32, key: (k0, k1) ints -> [n, 4] uint32."""
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = key
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0)) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1)) & mask
        n3 = p0 & mask
        c = [n0, n1, n2, n3]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=1).astype(np.uint32)


def sampler_uniforms(seed: int, first: int, n: int) -> np.ndarray:
    idx = np.arange(first, first + n, dtype=np.uint64)
    u = np.empty((n, 6), np.float64)
    for b in range(3):
        ctr = np.stack([(idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32),
                        np.full(n, b, np.uint32), np.full(n, TAG, np.uint32)], axis=1)
        w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.uint64)
        u[:, 2 * b] = ((w[:, 1] << np.uint64(32) | w[:, 0]) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
        u[:, 2 * b + 1] = ((w[:, 3] << np.uint64(32) | w[:, 2]) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0
    return u
