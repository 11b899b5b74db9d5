"""Seed-compatible latent noise sources: the three generators the reference's Swift pipeline offers
(``StableDiffusionRNG``: numpyRNG / torchRNG / nvidiaRNG, ``StableDiffusionPipeline.swift:438-447``) so that a given
``--seed`` reproduces the same initial latents as the reference CLIs.  Host-side, vectorised numpy; results are the
float64 streams the Swift sources produce (the pipeline casts them to float32 / float16 like the reference).

* ``NumPyRandomSource``  -- ``numpy.random.seed(seed); numpy.random.randn`` (MT19937 + polar method), the stream the
  Python pipeline itself uses (``pipeline.py:331,725-726``; ``NumPyRandomSource.swift:13-119``).
* ``TorchRandomSource``  -- ``torch.manual_seed(seed); torch.randn(n)`` on the CPU: MT19937 32-bit outputs, 24-bit
  uniforms, Box-Muller over blocks of 16 (``TorchRandomSource.swift:116-152``).
* ``NvRandomSource``     -- Philox-4x32-10 keyed by the seed, counter (offset, 0, index, 0), Box-Muller on the first two
  words (``NvRandomSource.swift:10-91``: "consistent with NVIDIA curandom").
"""
from __future__ import annotations

import numpy as np

_U32 = np.uint32
_U64 = np.uint64


class NumPyRandomSource:
    def __init__(self, seed: int):
        self._rs = np.random.RandomState(int(seed) & 0xFFFFFFFF)

    def normal_array(self, count: int, mean: float = 0.0, stdev: float = 1.0) -> np.ndarray:
        return self._rs.standard_normal(int(count)) * stdev + mean


class _MT19937:
    """32-bit Mersenne Twister with the reference's ``init_genrand`` seeding (TorchRandomSource.swift:31-41)."""
    N, M = 624, 397

    def __init__(self, seed: int):
        key = np.empty(self.N, dtype=np.uint64)
        s = int(seed) & 0xFFFFFFFF
        for i in range(self.N):
            key[i] = s
            s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
        self.key = key.astype(_U32)
        self.pos = self.N

    def _twist(self):
        k, n, m = self.key, self.N, self.M
        upper, lower, a = _U32(0x80000000), _U32(0x7FFFFFFF), _U32(0x9908B0DF)

        def mix(cur, nxt, far):
            y = (cur & upper) | (nxt & lower)
            return far ^ (y >> _U32(1)) ^ np.where((y & _U32(1)) != 0, a, _U32(0))

        # entries [0, n-m) read old values at i + m; the rest read values already updated in this pass
        k[: n - m] = mix(k[: n - m], k[1: n - m + 1], k[m:])
        lo = n - m
        while lo < n - 1:
            hi = min(lo + (n - m), n - 1)
            k[lo:hi] = mix(k[lo:hi], k[lo + 1: hi + 1], k[lo - (n - m): hi - (n - m)])
            lo = hi
        k[n - 1] = mix(k[n - 1: n], k[0:1], k[m - 1: m])[0]
        self.pos = 0

    def uint32(self, count: int) -> np.ndarray:
        out = np.empty(count, dtype=_U32)
        done = 0
        while done < count:
            if self.pos == self.N:
                self._twist()
            take = min(count - done, self.N - self.pos)
            y = self.key[self.pos: self.pos + take].copy()
            self.pos += take
            y ^= y >> _U32(11)
            y ^= (y << _U32(7)) & _U32(0x9D2C5680)
            y ^= (y << _U32(15)) & _U32(0xEFC60000)
            y ^= y >> _U32(18)
            out[done: done + take] = y
            done += take
        return out


class TorchRandomSource:
    def __init__(self, seed: int):
        self._mt = _MT19937(seed)
        self._next_gauss = None

    def _next_double(self, count: int) -> np.ndarray:
        w = self._mt.uint32(2 * count).astype(_U64)
        a = (w[0::2] << _U64(32)) | w[1::2]
        return (a & _U64(9007199254740991)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def _next_float(self, count: int) -> np.ndarray:
        return (self._mt.uint32(count) & _U32(16777215)).astype(np.float64) * (1.0 / 16777216.0)

    def _next_gauss_scalar(self) -> float:
        if self._next_gauss is not None:
            g, self._next_gauss = self._next_gauss, None
            return g
        u1 = float(self._next_double(1)[0])
        u2 = 1.0 - float(self._next_double(1)[0])
        radius = np.sqrt(-2.0 * np.log(u2))
        theta = 2.0 * np.pi * u1
        self._next_gauss = float(radius * np.sin(theta))
        return float(radius * np.cos(theta))

    @staticmethod
    def _box_muller16(block: np.ndarray) -> np.ndarray:
        """block [..., 16] of uniforms -> normals: pairs (j, j + 8), cos into the first half, sin into the second."""
        u1 = 1.0 - block[..., :8]
        u2 = block[..., 8:]
        radius = np.sqrt(-2.0 * np.log(u1))
        theta = 2.0 * np.pi * u2
        return np.concatenate([radius * np.cos(theta), radius * np.sin(theta)], axis=-1)

    def normal_array(self, count: int, mean: float = 0.0, stdev: float = 1.0) -> np.ndarray:
        count = int(count)
        if count < 16:  # torch draws these one by one from the scalar Box-Muller (with its cached second value)
            return np.array([self._next_gauss_scalar() * stdev + mean for _ in range(count)], dtype=np.float64)
        data = self._next_float(count)
        full = count - count % 16
        data[:full] = self._box_muller16(data[:full].reshape(-1, 16)).reshape(-1) * stdev + mean
        if count % 16:
            # the last 16 entries are redrawn from 53-bit uniforms and transformed again (TorchRandomSource.swift:136-149)
            tail = self._next_double(16)
            data[count - 16:] = self._box_muller16(tail) * stdev + mean
        return data


class NvRandomSource:
    _M0, _M1 = _U64(0xD2511F53), _U64(0xCD9E8D57)
    _W0, _W1 = 0x9E3779B9, 0xBB67AE85

    def __init__(self, seed: int):
        self.seed = int(seed) & 0xFFFFFFFF
        self.offset = 0

    def normal_array(self, count: int, mean: float = 0.0, stdev: float = 1.0) -> np.ndarray:
        count = int(count)
        c0 = np.full(count, self.offset & 0xFFFFFFFF, dtype=_U64)
        c1 = np.zeros(count, dtype=_U64)
        c2 = np.arange(count, dtype=_U64)
        c3 = np.zeros(count, dtype=_U64)
        self.offset += 1
        k0, k1 = self.seed & 0xFFFFFFFF, self.seed >> 32
        mask = _U64(0xFFFFFFFF)
        for r in range(10):
            v1, v2 = c0 * self._M0, c2 * self._M1
            c0, c1, c2, c3 = ((v2 >> _U64(32)) ^ c1 ^ _U64(k0)) & mask, v2 & mask, ((v1 >> _U64(32)) ^ c3 ^ _U64(k1)) & mask, v1 & mask
            if r < 9:
                k0, k1 = (k0 + self._W0) & 0xFFFFFFFF, (k1 + self._W1) & 0xFFFFFFFF
        u = c0.astype(np.float64) / 4294967296.0 + (1.0 / 8589934592.0)
        v = c1.astype(np.float64) * (np.pi / 2147483648.0) + (np.pi / 4294967296.0)
        return stdev * np.sqrt(-2.0 * np.log(u)) * np.sin(v) + mean


_SOURCES = {"numpy": NumPyRandomSource, "numpyRNG": NumPyRandomSource, "torch": TorchRandomSource,
            "torchRNG": TorchRandomSource, "nvidia": NvRandomSource, "nvidiaRNG": NvRandomSource}


def random_source(rng: str, seed: int):
    """``StableDiffusionPipeline.randomSource(from:seed:)`` (StableDiffusionPipeline.swift:438-447)."""
    try:
        return _SOURCES[rng](seed)
    except KeyError:
        raise ValueError(f"unknown rng {rng!r}; expected one of numpy, torch, nvidia") from None
