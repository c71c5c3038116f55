"""Host-side Xoshiro256++ (Julia's `Random.Xoshiro`) and the seeding policy for walker streams.

The device owns the per-walker streams once `kdsl_set_rng` has copied them in; this module only
creates states and lets a host-side `MCContext.rng` draw numbers exactly like the reference's
`rand(ctx.rng)` / `sample(ctx.rng, .)`.
"""
from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


def _rotl(x: int, k: int) -> int:
    return ((x << k) | (x >> (64 - k))) & _M


def splitmix64(x: int):
    """one SplitMix64 step: returns (new_state, output)"""
    x = (x + 0x9E3779B97F4A7C15) & _M
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M
    return x, z ^ (z >> 31)


class Xoshiro:
    """xoshiro256++ with Julia's output conventions."""

    def __init__(self, seed=None, state=None):
        if state is not None:
            self.s = [int(v) & _M for v in state]
            assert len(self.s) == 4 and any(self.s), "need 4 words, not all zero"
        else:
            x = (0 if seed is None else int(seed)) & _M
            self.s = []
            for _ in range(4):
                x, out = splitmix64(x)
                self.s.append(out)

    def next_u64(self) -> int:
        s0, s1, s2, s3 = self.s
        res = (_rotl((s0 + s3) & _M, 23) + s0) & _M
        t = (s1 << 17) & _M
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= t
        s3 = _rotl(s3, 45)
        self.s = [s0, s1, s2, s3]
        return res

    def rand(self) -> float:
        """rand(rng)::Float64"""
        return (self.next_u64() >> 11) * 2.0 ** -53

    def rand_index(self, n: int) -> int:
        """rand(rng, 1:n) -- SamplerRangeNDL; one draw is consumed even for n == 1"""
        x = self.next_u64()
        m = x * n
        lo = m & _M
        if lo < n:
            t = ((1 << 64) - n) % n
            while lo < t:
                x = self.next_u64()
                m = x * n
                lo = m & _M
        return (m >> 64) + 1

    def sample(self, a):
        """StatsBase.sample(rng, a)"""
        return a[self.rand_index(len(a)) - 1]


def walker_states(seed: int, n_walkers: int, first_walker: int = 0) -> np.ndarray:
    """uint64 [n_walkers][4]: walker w's state is four outputs of the SplitMix64 stream started at
    mix(seed) XOR mix(~w_global), mix = one SplitMix64 output.  Seed and global walker index are hashed
    JOINTLY: seed 0 is as good as any other, and distinct (seed, walker) pairs do not collide the way the
    multiplicative rule seed * (1 + w) did ((2, 1) == (4, 0); seed 0 gave every walker the same stream).
    `first_walker` offsets the global walker index so that ranks of a multi-GPU job get disjoint streams."""
    out = np.zeros((n_walkers, 4), dtype=np.uint64)
    _, key = splitmix64(int(seed) & _M)
    for w in range(n_walkers):
        _, wk = splitmix64(~(first_walker + w) & _M)
        x = key ^ wk
        for q in range(4):
            x, v = splitmix64(x)
            out[w, q] = v
    return out
