#!/usr/bin/env python
"""CPU emulation of the index algebra of qpsk_b200/csrc/fft.cuh (no GPU needed).

Walks every stage of the Stockham schedule exactly as the kernel's threads do -- same radix sequence,
same skewed shared-memory addresses formed as `per-thread base + compile-time offset`, same twiddle
table layout, same factored remainder-stage twiddles (table[m][j] * W_P^(m t)) and the same signed-
permutation table for the powers of W32 -- with numpy doing the small DFTs, and compares the result
with numpy.fft.  It proves the structure; the arithmetic itself is checked on the GPU (tests/test_fft_gpu.py).
"""
import numpy as np


TW_ON_STORE = False       # QPSK_FFT_TW_ON_STORE
P_BIG = {2048: 64, 4096: 64, 8192: 64}      # QPSK_FFT_P2048 / QPSK_FFT_P4096 / QPSK_FFT_P8192


def ppt(n):
    if n in P_BIG:
        return P_BIG[n]
    return 32 if n >= 512 else (16 if n >= 256 else (8 if n >= 8 else n))


def radix(rem, p):
    return p // 4 if (p == 32 and rem == 2 * p) else (p if rem >= p else rem)


def tw_cols(ns, tpf, p):
    return ns if (ns < tpf or p > 32) else tpf


def cos32(k):
    return np.cos(2 * np.pi * k / 32)


def w32_table(K):
    """the (base, swap, sign) decomposition of cmul_w32<K>; returns the complex constant it multiplies by"""
    k, q, r = K & 31, (K & 31) // 8, (K & 31) % 8
    if r == 0:
        return [1, -1j, -1, 1j][q]
    if r <= 4:
        c, s = cos32(r), np.sin(2 * np.pi * r / 32)
        bp, bm = (c, s), (c, -s)
        base, sw, sg = [(bm, 0, 1), (bp, 1, -1), (bm, 0, -1), (bp, 1, 1)][q]
    else:
        c, s = cos32(8 - r), np.sin(2 * np.pi * (8 - r) / 32)
        bp, bm = (c, s), (c, -s)
        base, sw, sg = [(bm, 1, -1), (bp, 0, -1), (bm, 1, 1), (bp, 0, 1)][q]
    u, v = (base[1], base[0]) if sw else base
    return sg * (u + 1j * v)


def emulate(n, x, split_half=None):
    """split_half = None: an n-point transform.  0 / 1: the even / odd half of a 2n-point burst whose radix-2 step has
    already been taken (x = x0 + x1 resp. x0 - x1): the odd half multiplies input r of stage 0 by W128^r and uses the
    twiddle base w_2n^(2k + 1) in the second stage (qpsk_fft_split)."""
    if split_half is None and n == 8192 and ppt(n) == 64 and ppt(4096) == 64:
        out = np.zeros(n, complex)
        out[0::2] = emulate(4096, x[:4096] + x[4096:], 0)
        out[1::2] = emulate(4096, x[:4096] - x[4096:], 1)
        return out
    p = ppt(n)
    tpf = n // p
    S = 64 if p >= 64 else (32 if p >= 32 else 16)
    lin = n >= 256
    skew = lambda i: i + i // S
    off = lambda c: c + c // S
    # twiddle tables, host layout
    stages, ns = [], 1
    while ns < n:
        r = radix(n // ns, p)
        stages.append((ns, r))
        ns *= r
    tables = {}
    tws = lambda ns_: TW_ON_STORE and p >= 64 and ns_ == p
    for (ns, r) in stages:
        if tws(ns):
            tables[ns] = np.array([[np.exp(-2j * np.pi * q * k / (ns * r)) for k in range(r)] for q in range(1, ns)])
        elif p >= 64 and ns == 64 and r == 64:
            hb = 0 if split_half is None else split_half
            tables[ns] = np.array([[np.exp(-2j * np.pi * ((8 * (i + 1)) if i < 7 else (i - 6)) * (2 * k + hb) / (2 * ns * r)) for k in range(64)] for i in range(14)])
        elif p >= 64 and r == 2 and ns > tpf and (2 * ns) // tpf == 64:
            tables[ns] = np.array([[np.exp(-2j * np.pi * k / (ns * r)) for k in range(tpf)]])
        elif ns > 1:
            cols = tw_cols(ns, tpf, p)
            tables[ns] = np.array([[np.exp(-2j * np.pi * m * k / (ns * r)) for k in range(cols)] for m in range(1, r)])
    base = 0
    sdat = np.zeros(skew(n) + 8, complex)
    pts = np.zeros((tpf, p), complex)
    for si, (ns, r) in enumerate(stages):
        first, last = si == 0, ns * r == n
        nb, stride = p // r, n // r
        kt = tw_cols(ns, tpf, p)
        fact = last and ns > tpf and p <= 32
        perbf = ns > tpf and not fact
        for j in range(tpf):
            rd = skew(base + j)
            for t in range(nb):
                v = np.zeros(r, complex)
                for rr in range(r):
                    c = t * tpf + rr * stride
                    if first:
                        v[rr] = x[j + c]
                        if split_half == 1:
                            assert nb == 1 and stride == 64
                            v[rr] *= np.exp(-2j * np.pi * rr / 128)
                    elif lin:
                        assert rd + off(c) == skew(base + j + c), (n, ns, r, j, c)
                        v[rr] = sdat[rd + off(c)]
                    else:
                        v[rr] = sdat[skew(base + j + c)]
                if p >= 64 and ns == 64 and r == 64 and not tws(ns):
                    for m in range(1, r):
                        if m // 8:
                            v[m] *= tables[ns][m // 8 - 1][j % 64]
                        if m % 8:
                            v[m] *= tables[ns][7 + m % 8 - 1][j % 64]
                elif p >= 64 and r == 2 and ns > tpf and (2 * ns) // tpf == 64:
                    v[1] *= tables[ns][0][j] * np.exp(-2j * np.pi * t / 64)
                elif ns > 1 and not tws(ns):
                    k = (j + t * tpf) if perbf else j % kt
                    for m in range(1, r):
                        v[m] *= tables[ns][m - 1][k]
                        if fact and t > 0:
                            v[m] *= w32_table((m * t * (32 // p)) & 31)
                o = np.fft.fft(v)
                if first and not last and tws(r):
                    rn = stages[si + 1][1]
                    assert nb == 1
                    for q in range(1, r):
                        o[q] *= tables[r][q - 1][(j * r * rn) // n]
                pts[j, t * r:(t + 1) * r] = o
        if not last:
            new = np.zeros_like(sdat)
            for j in range(tpf):
                dyn = j * r if ns == 1 else (j // ns) * ns * r + (j % ns)
                wr = skew(base + dyn)
                for t in range(nb):
                    for q in range(r):
                        c = t * tpf * r + q * ns
                        jj = j + t * tpf
                        true = (jj // ns) * ns * r + jj % ns + q * ns
                        if lin:
                            assert wr + off(c) == skew(base + true), (n, ns, r, j, t, q)
                            new[wr + off(c)] = pts[j, t * r + q]
                        else:
                            new[skew(base + true)] = pts[j, t * r + q]
            sdat = new
    nsl, rlast = stages[-1]
    out = np.zeros(n, complex)
    for j in range(tpf):
        for i in range(p):
            t, q = i // rlast, i % rlast
            out[(j + t * tpf) + q * nsl] = pts[j, i]
    return out


def main():
    for K in range(32):
        assert abs(w32_table(K) - np.exp(-2j * np.pi * K / 32)) < 1e-12, K
    rng = np.random.default_rng(0)
    for pbig in (16, 32, 64):
        P_BIG[2048] = P_BIG[4096] = P_BIG[8192] = pbig
        for n in (2048, 4096, 8192):
            x = rng.normal(size=n) + 1j * rng.normal(size=n)
            err = np.max(np.abs(emulate(n, x) - np.fft.fft(x))) / np.max(np.abs(np.fft.fft(x)))
            assert err < 1e-10, (n, pbig, err)
            print("n = %5d  P = %d ok" % (n, pbig))
    for lg in range(1, 14):
        n = 1 << lg
        x = rng.normal(size=n) + 1j * rng.normal(size=n)
        err = np.max(np.abs(emulate(n, x) - np.fft.fft(x))) / np.max(np.abs(np.fft.fft(x)))
        assert err < 1e-10, (n, err)
        print("n = %5d  ok  (rel err %.1e)" % (n, err))


if __name__ == "__main__":
    main()
