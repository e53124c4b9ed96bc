"""Executable specification of the order-independent ("closed form") algorithms the CUDA engine
implements, in slow plain Python. TEST INFRASTRUCTURE: it exists so that the *algorithms* can be
checked against the oracle on CPU (tests/test_algo_model.py) independently of the kernels; the
kernels in nlzm_b200/csrc mirror these functions one to one. Never imported by the product.

  bt4_model   exhaustive BT4 == "for every length L, nearest earlier position sharing >= L bytes"
              computed by a bottom-up divide and conquer over position ranges with
              previous/next-greater-position chains in suffix-rank order
  ht_model    HT2 / HT3 cell contents resolved by "last writer" recursion instead of a table
  rk_model    RK256 hash, slot lists, raw hits, and the sparse carry state machine
"""
from __future__ import annotations

import bisect
import numpy as np

NONE = 0xFFFFFFFF
MATCH_MAX = 264
HASH_MUL = 987660757
ADDH = 0x2F0FD693
M32 = 0xFFFFFFFF


def clamp(v, lo, hi):
    return lo if v < lo else hi if v > hi else v


def match_min(d):
    return 2 + (d >= 256) + (d >= 4096) + (d >= (1 << 20))


class Geometry:
    """Encoder geometry that is part of the matcher semantics (NLZM.cpp:1716-1725,1782-1798)."""

    def __init__(self, flen, hist_bits):
        hb = hist_bits
        while hb > 10 and flen < (1 << (hb - 1)):
            hb -= 1
        self.flen, self.hb, self.W = flen, hb, 1 << hb
        self.frame_bits = clamp(hb - 2, 14, 17)
        self.cs = ((1 << self.frame_bits) * 15) // 16 - 0x200
        self.ht3_bits = 12 + clamp(hb, 15, 17) - 15
        self.bt_bits = 13 + clamp(hb, 16, 20) - 16
        self.rk_bits = 15 + clamp(hb, 16, 22) - 16
        # ring shifts: simulate encode_file's chunk loop once (O(chunks))
        self.shift_starts = []     # absolute chunk starts at which a shift happened
        hist_pos, a0 = 0, 0
        while a0 < flen:
            if hist_pos >= 2 * self.W:
                hist_pos -= self.W
                self.shift_starts.append(a0)
            step = min(self.cs, flen - a0)
            hist_pos += step
            a0 += step

    def epoch(self, a):
        """number of ring shifts that happened at or before position a's chunk"""
        return bisect.bisect_right(self.shift_starts, a)

    def P(self, a):
        return a - self.W * self.epoch(a)

    def rem(self, a):
        k = a // self.cs
        return min((k + 1) * self.cs + MATCH_MAX + 1, self.flen) - a


def lcp(x, p0, p1, cap):
    m = 0
    n = len(x)
    while m < cap and p1 + m < n and x[p0 + m] == x[p1 + m]:
        m += 1
    return m


# --------------------------------------------------------------------------------------------
# BT4, exhaustive: divide and conquer with greater-position chains
# --------------------------------------------------------------------------------------------

def suffix_ranks(x, depth=MATCH_MAX):
    n = len(x)
    b = bytes(x)
    order = sorted(range(n), key=lambda i: (b[i:i + depth], i))
    rank = [0] * n
    for r, i in enumerate(order):
        rank[i] = r
    return rank


def bt4_model(x, hist_bits, min_lcp=4):
    """Returns candidate tuples (a, dist, len) — a superset of the staircase that the final
    dominance filter (oracle.records_to_csr) reduces to exactly the BT4 staircase."""
    n = len(x)
    g = Geometry(n, hist_bits)
    W = g.W
    rank = suffix_ranks(x)
    cur = list(range(n))
    PG = [NONE] * n
    NG = [NONE] * n
    best = [min_lcp - 1] * n
    out = []
    h = 1
    while h < n:
        nxt = cur[:]
        for base in range(0, n, 2 * h):
            L = cur[base:base + h]
            R = cur[base + h:base + 2 * h]
            if not R:
                continue
            mid = base + h
            Lr = [rank[q] for q in L]
            Rr = [rank[q] for q in R]
            merged = sorted(L + R, key=lambda q: rank[q])
            nxt[base:base + len(merged)] = merged
            # queries: right elements against the left child
            for a in R:
                cap = min(MATCH_MAX, n - a)
                if n - a < 4 or best[a] >= cap or a - (mid - 1) > W - 1:
                    continue
                t = bisect.bisect_left(Lr, rank[a])
                nb = best[a]
                for start, ptr in ((L[t - 1] if t > 0 else NONE, PG), (L[t] if t < len(L) else NONE, NG)):
                    c = start
                    while c != NONE:
                        l = lcp(x, c, a, cap)
                        if l <= best[a]:
                            break
                        d = a - c
                        if d <= W - 1 and l >= match_min(d):
                            out.append((a, d, l))
                        nb = max(nb, l)
                        c = ptr[c]
                best[a] = nb
            # pointer maintenance: left elements see the right child
            for q in L:
                u = bisect.bisect_left(Rr, rank[q])
                if u > 0:
                    rl = R[u - 1]
                    if PG[q] == NONE or rank[rl] > rank[PG[q]]:
                        PG[q] = rl
                if u < len(R):
                    rr = R[u]
                    if NG[q] == NONE or rank[rr] < rank[NG[q]]:
                        NG[q] = rr
        cur = nxt
        h *= 2
    return out


def bt4_short_model(x, hist_bits):
    """lcp 2..3 candidates that share a BT4 bucket by hash collision (only possible when the
    bucket index has fewer than 16 bits, i.e. hist_bits < 19)."""
    n = len(x)
    g = Geometry(n, hist_bits)
    out = []
    if g.bt_bits >= 16:
        return out
    sh = 32 - g.bt_bits

    def bucket(i):
        v = x[i] | (x[i + 1] << 8) | (x[i + 2] << 16) | (x[i + 3] << 24)
        return ((v * HASH_MUL) & M32) >> sh
    bk = [bucket(i) for i in range(n - 3)]
    for a in range(n - 3):
        cap = min(MATCH_MAX, n - a)
        seen = 0
        for q in range(a - 1, max(-1, a - 4096, a - g.W), -1):
            if bk[q] != bk[a] or x[q] != x[a] or x[q + 1] != x[a + 1]:
                continue
            l = lcp(x, q, a, min(cap, 4))
            if l >= 4 or l <= seen:
                continue
            if l >= match_min(a - q):
                out.append((a, a - q, l))
                seen = l
            if seen == 3:
                break
    return out


# --------------------------------------------------------------------------------------------
# HT2 / HT3: last-writer resolution
# --------------------------------------------------------------------------------------------

def ht_model(x, hist_bits, which):
    n = len(x)
    g = Geometry(n, hist_bits)
    hb, W = g.hb, g.W
    rows, bits, nbytes = (1, 12, 2) if which == 2 else (2, g.ht3_bits, 3)
    cmask = (1 << (32 - hb)) - 1
    called = [a for a in range(n) if g.rem(a) >= 4]
    hashes = {}
    by_bucket = {}
    for a in called:
        v = 0
        for i in range(nbytes):
            v |= x[a + i] << (8 * i)
        h = (v * HASH_MUL) & M32
        hashes[a] = h
        by_bucket.setdefault(h >> (32 - bits), []).append(a)

    def entry(a):
        return (g.P(a) | ((hashes[a] & cmask) << hb)) & M32

    def last_before(bucket, t):
        lst = by_bucket.get(bucket)
        if not lst:
            return -1
        i = bisect.bisect_left(lst, t)
        return lst[i - 1] if i else -1

    def value(c, t):
        """raw content of cell c just before the access at position t"""
        while True:
            a_c = last_before(c, t)
            a_l = last_before(c - 1, t) if (rows == 2 and c >= 1) else -1
            w = max(a_c, a_l)
            if c == 0:
                # MatchFinderHT::Shift as written clears cell 0 at every ring shift
                i = bisect.bisect_right(g.shift_starts, t)
                clr = g.shift_starts[i - 1] if i else -1
                if clr > w:
                    return NONE
            if w < 0:
                return NONE
            if a_c == w:
                return entry(a_c)
            c, t = c - 1, a_l

    out = []
    for a in called:
        h = hashes[a]
        b = h >> (32 - bits)
        chk = h & cmask
        P = g.P(a)
        cap = min(g.rem(a), MATCH_MAX)
        base = a - P          # absolute offset of shifted coordinate 0 in a's epoch
        best = 1
        for i in range(rows):
            row = value(b + i, a)
            if best < cap and (row >> hb) == chk:
                sp = row & (W - 1)
                if sp < P and P - sp <= W - 1:
                    m = lcp(x, sp + base, a, cap)
                    if m > best and m >= match_min(P - sp):
                        out.append((a, P - sp, m))
                        best = m
    return out


# --------------------------------------------------------------------------------------------
# RK256
# --------------------------------------------------------------------------------------------

def rk_hashes(x):
    """H(s) = sum x[s+i] * ADDH^(256-i) mod 2^32 for every s with s+256 <= n (NLZM.cpp:798-799)."""
    n = len(x)
    if n < 256:
        return []
    remh = pow(ADDH, 256, 1 << 32)
    h = 0
    for i in range(256):
        h = ((x[i] + h) * ADDH) & M32
    out = [h]
    for s in range(1, n - 255):
        h = ((x[s + 255] + h - x[s - 1] * remh) * ADDH) & M32
        out.append(h)
    return out


def rk_model(x, hist_bits):
    n = len(x)
    g = Geometry(n, hist_bits)
    hb, W = g.hb, g.W
    cmask = (1 << (32 - hb)) - 1
    H = rk_hashes(x)
    called = [a for a in range(n) if g.rem(a) >= 256]
    sh = 32 - g.rk_bits
    # table build: aligned blocks grouped by slot, in position order
    by_slot = {}
    for a in called:
        if a % 256 == 0:
            by_slot.setdefault(H[a] >> sh, []).append(a)
    # raw hits: every position looks up the last aligned block inserted before it
    hits = []
    for a in called:
        lst = by_slot.get(H[a] >> sh, [])
        i = bisect.bisect_left(lst, a)
        if i:
            b = lst[i - 1]
            e = (g.P(b) | ((H[b] << hb) & M32)) & M32
        else:
            e = NONE      # an empty slot is the raw word 0xFFFFFFFF and is matched like any entry
        sp = e & (W - 1)
        P = g.P(a)
        if (e >> hb) == (H[a] & cmask) and sp < P and P - sp <= W - 1:
            m = lcp(x, sp + (a - P), a, g.rem(a) & 0xFFFF)
            if m >= match_min(P - sp):
                hits.append((a, P - sp, m))
    # sparse carry state machine over the hit list (NLZM.cpp:1056-1069,1090-1107)
    out = []
    cl = 0
    ca = cd = 0     # carry start (absolute) and distance
    cep = -1        # epoch of the carry start
    for (a, d, m) in hits:
        alive = cl > 0 and g.epoch(a) == cep and a - ca < cl
        if alive and cl >= 256:
            continue                       # lookups are suppressed under a long carried match
        if alive and m < cl:
            continue                       # a new hit must be at least as long as the carry's original length
        # close the previous carry interval, open a new one
        if cl > 0:
            out.append((ca, cd, cl, min(a + 1, _carry_end(g, ca, cl, cep))))  # (A) still fires at a
        ca, cd, cl, cep = a, d, m, g.epoch(a)
    if cl > 0:
        out.append((ca, cd, cl, _carry_end(g, ca, cl, cep)))
    # expand carry intervals into per-position candidates
    cands = []
    for (a0, d, m, end) in out:
        for a in range(a0, end):
            if g.rem(a) < 256:
                break
            r = m - (a - a0)
            if r >= match_min(d):
                cands.append((a, d, min(r, MATCH_MAX)))
    return cands


def _carry_end(g, ca, cl, cep):
    """first position at which the carry started at ca (length cl) is no longer alive"""
    end = ca + cl
    i = bisect.bisect_right(g.shift_starts, ca)
    if i < len(g.shift_starts):
        end = min(end, g.shift_starts[i])
    return end
