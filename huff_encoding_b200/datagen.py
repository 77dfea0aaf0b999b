"""Integer-only synthetic inputs for BASELINE.json's configs.

Every generator is pure integer arithmetic on h(i) = splitmix64(seed + i), so the numpy (host) and torch (device)
versions produce identical bytes.  Nothing here reads the reference or the oracle.

    english  : order-0 sampling of a fixed 32-symbol letter/space/punctuation table   (configs[0], configs[3])
    uniform  : byte = h(i) & 0xFF                                                      (configs[1])
    zipf     : rank k in 1..256 with P ~ k^-1.2, symbol = k-1                          (configs[2])
    fibonacci: 224 symbols x 1 + 32 symbols with Fibonacci weights (40-bit codes)      (configs[4])
"""
from __future__ import annotations

import numpy as np

SEED_BASE = 0x5EED0000
_M64 = (1 << 64) - 1


# ------------------------------------------------------------------ splitmix64
def splitmix64_np(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _to_i64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


def splitmix64_torch(x):
    """x: int64 tensor holding the u64 bit pattern.  Logical shifts are emulated with masks."""
    import torch  # noqa: F401

    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    x = x + _to_i64(0x9E3779B97F4A7C15)
    x = (x ^ lsr(x, 30)) * _to_i64(0xBF58476D1CE4E5B9)
    x = (x ^ lsr(x, 27)) * _to_i64(0x94D049BB133111EB)
    x = x ^ lsr(x, 31)
    return x


# ------------------------------------------------------------------ tables (u32 cumulative thresholds on h>>32)
# A fixed 32-symbol English-like table: (symbol, weight in 1/10000).  Frequencies are rounded textbook letter
# frequencies scaled to leave room for space and punctuation; what matters is that the table is fixed.
_ENGLISH = [
    (" ", 1700), ("e", 1000), ("t", 750), ("a", 650), ("o", 620), ("i", 580), ("n", 560), ("s", 520),
    ("h", 500), ("r", 480), ("d", 350), ("l", 330), ("u", 230), ("c", 220), ("m", 200), ("w", 190),
    ("f", 180), ("g", 160), ("y", 160), ("p", 150), ("b", 120), ("v", 80), ("k", 60), (",", 60),
    (".", 55), ("\n", 40), ("x", 15), ("j", 12), ("q", 10), ("z", 8), ("'", 6), ("-", 4),
]


def english_table() -> tuple[np.ndarray, np.ndarray]:
    sym = np.array([ord(c) for c, _ in _ENGLISH], dtype=np.uint8)
    w = np.array([w for _, w in _ENGLISH], dtype=np.uint64)
    cum = np.cumsum(w)
    thr = (cum * np.uint64(1 << 32)) // np.uint64(cum[-1])      # exclusive upper bounds on h>>32; last == 2^32
    return sym, thr.astype(np.uint64)


def zipf_table(s_num: int = 12, s_den: int = 10) -> tuple[np.ndarray, np.ndarray]:
    """P(k) ~ k^-1.2 for k = 1..256.  The float pow is evaluated once, rounded to integers and FROZEN as u32
    thresholds; data generation itself is integer-only."""
    k = np.arange(1, 257, dtype=np.float64)
    p = k ** (-(s_num / s_den))
    w = np.maximum(1, np.round(p / p.sum() * (1 << 32))).astype(np.uint64)
    cum = [int(c) for c in np.cumsum(w)]                      # exact integers: cum * 2^32 can exceed 64 bits
    thr = np.array([(c << 32) // cum[-1] for c in cum], dtype=np.uint64)
    return np.arange(256, dtype=np.uint8), thr


def _sample_np(n: int, seed: int, offset: int, sym: np.ndarray, thr: np.ndarray) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 22
    for s in range(0, n, step):
        m = min(step, n - s)
        idx = np.arange(offset + s, offset + s + m, dtype=np.uint64) + np.uint64(seed)
        u = splitmix64_np(idx) >> np.uint64(32)
        out[s:s + m] = sym[np.searchsorted(thr, u, side="right")]
    return out


def _sample_torch(n: int, seed: int, offset: int, sym: np.ndarray, thr: np.ndarray, device):
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    sym_t = torch.from_numpy(sym.astype(np.int64)).to(device)
    thr_t = torch.from_numpy(thr.astype(np.int64)).to(device)
    step = 1 << 26
    for s in range(0, n, step):
        m = min(step, n - s)
        idx = torch.arange(offset + s, offset + s + m, dtype=torch.int64, device=device) + seed
        h = splitmix64_torch(idx)
        u = (h >> 32) & 0xFFFFFFFF
        out[s:s + m] = sym_t[torch.searchsorted(thr_t, u, right=True)].to(torch.uint8)
    return out


# ------------------------------------------------------------------ public generators
def seed_for(config_index: int) -> int:
    return SEED_BASE + config_index


def english(n: int, seed: int = SEED_BASE + 0, offset: int = 0, device=None):
    sym, thr = english_table()
    return _sample_np(n, seed, offset, sym, thr) if device is None else _sample_torch(n, seed, offset, sym, thr, device)


def zipf(n: int, seed: int = SEED_BASE + 2, offset: int = 0, device=None, s: tuple[int, int] = (12, 10)):
    """Zipf letters with exponent s[0]/s[1] (default 1.2 = BASELINE.json configs[2]; 1.5 and 2.0 give long-tailed
    code sets whose rare letters get 13..16-bit codes)."""
    sym, thr = zipf_table(*s)
    return _sample_np(n, seed, offset, sym, thr) if device is None else _sample_torch(n, seed, offset, sym, thr, device)


def uniform(n: int, seed: int = SEED_BASE + 1, offset: int = 0, device=None):
    if device is None:
        out = np.empty(n, dtype=np.uint8)
        step = 1 << 22
        for s in range(0, n, step):
            m = min(step, n - s)
            idx = np.arange(offset + s, offset + s + m, dtype=np.uint64) + np.uint64(seed)
            out[s:s + m] = (splitmix64_np(idx) & np.uint64(0xFF)).astype(np.uint8)
        return out
    import torch
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 26
    for s in range(0, n, step):
        m = min(step, n - s)
        idx = torch.arange(offset + s, offset + s + m, dtype=torch.int64, device=device) + seed
        out[s:s + m] = (splitmix64_torch(idx) & 0xFF).to(torch.uint8)
    return out


def fibonacci_weights(n_fib: int = 32, n_ones: int = 224, first: int = 13) -> np.ndarray:
    """256-bin weights: `n_ones` symbols with weight 1 and `n_fib` symbols with weights F_first, F_first+1, ...
    (F_1 = F_2 = 1).  With the defaults the total is 1 836 311 750 and the longest code has 40 bits."""
    fib = [1, 1]
    while len(fib) < first + n_fib:
        fib.append(fib[-1] + fib[-2])
    w = np.zeros(256, dtype=np.uint64)
    w[:n_ones] = 1
    for j in range(n_fib):
        w[n_ones + j] = fib[first - 1 + j]
    return w


def from_weights_runs(w: np.ndarray, device=None):
    """Contiguous runs: byte b repeated w[b] times, ascending b."""
    w = np.asarray(w, dtype=np.int64)
    if device is None:
        return np.repeat(np.arange(256, dtype=np.uint8), w)
    import torch
    return torch.repeat_interleave(torch.arange(256, dtype=torch.uint8, device=device),
                                   torch.from_numpy(w).to(device))


def from_weights_permuted(w: np.ndarray, seed: int = SEED_BASE + 4):
    """Index-permuted variant (host only): a bijective integer mix of the run layout."""
    runs = from_weights_runs(w)
    n = runs.size
    key = splitmix64_np(np.arange(n, dtype=np.uint64) + np.uint64(seed))
    return runs[np.argsort(key, kind="stable")]
