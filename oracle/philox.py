"""Host restatement of the in-kernel noise stream (TEST INFRASTRUCTURE ONLY).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11) keyed by the rollout seed with counter (env lo, env hi, step, 0), then
Box-Muller on pairs -- the arithmetic of philox_normal4 in
trajopt_grpo_b200/csrc/tg_mlp.cuh.  The integer part is bit-exact; the normals
agree to float32 rounding of logf/sincospif.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0 & MASK, p1 & MASK, n2 & MASK, p0 & MASK
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_known_answer():
    """Random123 KAT: counter 0, key 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    z = np.zeros(1, np.uint64)
    return [int(v[0]) for v in philox4x32_10(z, z, z, z, 0, 0)]


def normals(seed, N, T, A):
    """noise[T, A, N] float32, the stream tg_rollout(noise=NULL, seed) consumes."""
    n = np.arange(N, dtype=np.uint64)
    out = np.zeros((T, A, N), np.float32)
    k24 = np.float32(2.0 ** -24)
    for t in range(T):
        c = philox4x32_10(n & MASK, n >> np.uint64(32), np.full(N, t, np.uint64), np.zeros(N, np.uint64),
                          seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        z = []
        for i in range(2):
            u1 = ((c[2 * i] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * k24
            u2 = ((c[2 * i + 1] >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * k24
            r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
            ang = (np.float64(2.0) * u2.astype(np.float64)) * np.pi
            z += [(r * np.cos(ang).astype(np.float32)).astype(np.float32),
                  (r * np.sin(ang).astype(np.float32)).astype(np.float32)]
        for j in range(A):
            out[t, j] = z[j]
    return out
