"""TEST INFRASTRUCTURE (oracle): restatement of numpy.random.randn of the legacy global generator, the source of
randomness of the reference's Gaussian-noise step (modules/preprocessor.py:115-119).

Follows numpy's published C sources: numpy/random/src/mt19937/mt19937.c (mt19937_gen: the recurrence and tempering),
mt19937.h (mt19937_next_double: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53) and
numpy/random/src/legacy/legacy-distributions.c (legacy_gauss: Marsaglia's polar method, f*x2 returned first and f*x1
kept for the next call).  Pinned by tests/test_oracle_spec.py against numpy.random itself (values and final state)."""
import math

import numpy as np

_N, _LAG = 624, 227


def _twist(u, v):
    y = (u & 0x80000000) | (v & 0x7FFFFFFF)
    return (y >> 1) ^ (0x9908B0DF if y & 1 else 0)


def _temper(y):
    y = y.astype(np.uint64)
    y ^= y >> 11
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= y >> 18
    return (y & 0xFFFFFFFF).astype(np.uint32)


def mt_extend(key, total):
    """Raw (untempered) MT19937 words: the 624 of `key` followed by the next total-624 of the recurrence."""
    raw = [int(k) for k in key]
    for k in range(_N, total):
        raw.append(raw[k - _LAG] ^ _twist(raw[k - _N], raw[k - _N + 1]))
    return np.array(raw, dtype=np.uint32)


def randn_replay(state, n):
    """(values, state after) of numpy.random.randn(n) started from `state` = numpy.random.get_state()."""
    name, key, pos, has_gauss, cached = state
    off = 1 if has_gauss else 0
    out = np.empty(n)
    if n == 0:
        return out, state
    if off:
        out[0] = cached
    fresh = n - off
    if fresh == 0:
        return out, (name, key, pos, 0, 0.0)
    pairs = (fresh + 1) // 2
    attempts = int(pairs / 0.7853 + 8 * math.sqrt(pairs) + 64)
    total = (pos + 4 * attempts + _N - 1) // _N * _N + _N
    raw = mt_extend(key, total)
    w = _temper(raw[pos:pos + 4 * attempts]).reshape(attempts, 4)

    def dbl(a, b):
        return ((a >> 5).astype(np.float64) * 67108864.0 + (b >> 6).astype(np.float64)) / 9007199254740992.0

    x1 = 2.0 * dbl(w[:, 0], w[:, 1]) - 1.0
    x2 = 2.0 * dbl(w[:, 2], w[:, 3]) - 1.0
    r2 = x1 * x1 + x2 * x2
    idx = np.nonzero((r2 < 1.0) & (r2 != 0.0))[0][:pairs]
    assert len(idx) == pairs
    # libm's scalar log, as legacy_gauss calls it (numpy's vectorised np.log differs in the last bit now and then)
    f = np.array([math.sqrt(-2.0 * math.log(v) / v) for v in r2[idx]])
    g = np.empty(2 * pairs)
    g[0::2] = f * x2[idx]
    g[1::2] = f * x1[idx]
    out[off:] = g[:fresh]
    q = pos + 4 * (int(idx[-1]) + 1)
    block = (q - 1) // _N
    left_over = fresh % 2
    return out, (name, raw[_N * block:_N * block + _N].copy(), q - _N * block, left_over, float(g[-1]) if left_over else 0.0)


def add_gaussian_noise(mat, sigma, state):
    """modules/preprocessor.py:115-119 with the generator state made explicit; returns (uint8 image, state after)."""
    noise, after = randn_replay(state, mat.size)
    out = mat + noise.reshape(mat.shape) * sigma
    return np.clip(out, 0., 255.).astype(np.uint8), after
