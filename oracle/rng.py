"""Random streams for the oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).

Two providers with one interface:

``LegacyRNG``  consumes NumPy's *global* legacy ``RandomState`` in exactly the order
the reference does: per transition ``D`` normals
(``quadpotential.py:200-203`` via ``base_hmc.py:135``), then per doubling one uniform
for the direction (``nuts.py:177``), one uniform per completed inner merge in
post-order (``nuts.py:375``), one uniform for the top-level pick (``nuts.py:290``).
It exists solely to reproduce ``tests/test_step.py``'s golden traces.

``PhiloxRNG``  is the stream shared with the CUDA engine: Philox4x32-10 keyed by the
chain's 64-bit seed and *addressed by counter* (transition, purpose, a, b), so the
recursive oracle and the iterative device kernel draw identical numbers no matter in
which order they ask.  The device twin is ``pymc3_b200/csrc/philox.cuh``.
"""
import numpy as np

PURPOSE_MOMENTUM = 0
PURPOSE_DIRECTION = 1
PURPOSE_MERGE = 2
PURPOSE_TOP = 3
PURPOSE_HMC_JITTER = 4
PURPOSE_HMC_ACCEPT = 5

_M0 = 0xD2511F53
_M1 = 0xCD9E8D57
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """Philox4x32 with 10 rounds (Salmon et al. 2011).  Pure-Python ints."""
    c0, c1, c2, c3 = (int(c) & _MASK for c in counter)
    k0, k1 = (int(k) & _MASK for k in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def philox4x32_10_vec(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over uint64-held 32-bit lanes (NumPy)."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(int(k0) & _MASK)
    k1 = np.uint64(int(k1) & _MASK)
    m0, m1 = np.uint64(_M0), np.uint64(_M1)
    mask, sh = np.uint64(_MASK), np.uint64(32)
    for _ in range(10):
        p0 = m0 * c0
        p1 = m1 * c2
        hi0, lo0 = p0 >> sh, p0 & mask
        hi1, lo1 = p1 >> sh, p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + np.uint64(_W0)) & mask
        k1 = (k1 + np.uint64(_W1)) & mask
    return c0, c1, c2, c3


def u53(hi, lo):
    """Two 32-bit words -> double in [0, 1) with 53 random bits."""
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    return ((hi >> np.uint64(5)) * 67108864.0 + (lo >> np.uint64(6))) / 9007199254740992.0


class PhiloxRNG:
    """Counter-addressed stream; key = the chain's 64-bit seed."""

    kind = "philox"

    def __init__(self, seed):
        seed = int(seed)
        self.key = (seed & _MASK, (seed >> 32) & _MASK)

    def seed(self, seed):
        self.__init__(seed)

    def _uniform(self, t, purpose, a=0, b=0):
        r = philox4x32_10((t, purpose, a, b), self.key)
        return float(u53(r[0], r[1]))

    def momentum(self, t, n):
        """n standard normals: Box-Muller on pairs, element i uses counter a=i>>1."""
        npair = (n + 1) // 2
        a = np.arange(npair, dtype=np.uint64)
        r0, r1, r2, r3 = philox4x32_10_vec(t, PURPOSE_MOMENTUM, a, 0, *self.key)
        u1 = 1.0 - u53(r0, r1)          # (0, 1]
        u2 = u53(r2, r3)
        rad = np.sqrt(-2.0 * np.log(u1))
        ang = 2.0 * np.pi * u2
        z = np.empty(2 * npair)
        z[0::2] = rad * np.cos(ang)
        z[1::2] = rad * np.sin(ang)
        return z[:n]

    def direction_u(self, t, depth):
        return self._uniform(t, PURPOSE_DIRECTION, depth)

    def merge_u(self, t, depth, level, leaf):
        return self._uniform(t, PURPOSE_MERGE, depth, (level << 16) | leaf)

    def top_u(self, t, depth):
        return self._uniform(t, PURPOSE_TOP, depth)

    def hmc_jitter(self, t, elow=0.85, ehigh=1.15):
        return elow + (ehigh - elow) * self._uniform(t, PURPOSE_HMC_JITTER)

    def hmc_accept_u(self, t):
        return self._uniform(t, PURPOSE_HMC_ACCEPT)


class LegacyRNG:
    """NumPy global RandomState, consumed in the reference's order."""

    kind = "legacy"

    def __init__(self, seed=None):
        if seed is not None:
            self.seed(seed)

    def seed(self, seed):
        np.random.seed(seed)        # sampling.py:883-884

    def momentum(self, t, n):
        return np.random.normal(size=n)      # quadpotential.py:200-203, 378-380

    def direction_u(self, t, depth):
        return np.random.uniform()           # nuts.py:30-33, 177

    def merge_u(self, t, depth, level, leaf):
        return np.random.uniform()           # nuts.py:375

    def top_u(self, t, depth):
        return np.random.uniform()           # nuts.py:290

    def hmc_jitter(self, t, elow=0.85, ehigh=1.15):
        return np.random.uniform(elow, ehigh)   # hmc.py:26-27

    def hmc_accept_u(self, t):
        return np.random.rand()              # hmc.py:136
