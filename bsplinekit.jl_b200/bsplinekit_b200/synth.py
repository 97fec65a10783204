"""Deterministic synthetic inputs.

u(s, i) = (splitmix64(s * 0x9E3779B97F4A7C15 + i) >> 11) * 2^-53 — an integer hash plus one
exact float operation, so numpy (host) and torch (device) produce identical bits and every GPU
can generate its own slice in place.
"""
import numpy as np

_G = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB
_M64 = (1 << 64) - 1


def u_numpy(seed, start, n):
    """u(seed, i) for i in [start, start + n) as float64."""
    with np.errstate(over="ignore"):
        i = np.arange(start, start + n, dtype=np.uint64)
        z = np.uint64((seed * _G) & _M64) + i
        z = z + np.uint64(_G)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def _s64(v):
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, r):
    import torch
    return torch.bitwise_and(torch.bitwise_right_shift(z, r), (1 << (64 - r)) - 1)


def u_torch(seed, start, n, device="cuda", dtype=None, chunk=1 << 26):
    """Same numbers as u_numpy, generated on `device` (int64 two's-complement arithmetic)."""
    import torch
    dtype = dtype or torch.float64
    out = torch.empty(n, dtype=dtype, device=device)
    base = _s64(seed * _G + _G)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        z = torch.arange(start + c0, start + c0 + m, dtype=torch.int64, device=device) + base
        z = (z ^ _lsr(z, 30)) * _s64(_C1)
        z = (z ^ _lsr(z, 27)) * _s64(_C2)
        z = z ^ _lsr(z, 31)
        out[c0:c0 + m] = (_lsr(z, 11).to(torch.float64) * (2.0 ** -53)).to(dtype)
    return out


def breakpoints(seed, n):
    """Non-uniform strictly increasing breakpoints on [0, 1]:
    xi_i = (i + 0.4 (2 u(seed, i) - 1)) / (n - 1), ends pinned to 0 and 1."""
    i = np.arange(n, dtype=np.float64)
    xi = (i + 0.4 * (2.0 * u_numpy(seed, 0, n) - 1.0)) / (n - 1)
    xi[0] = 0.0
    xi[-1] = 1.0
    return xi
