"""Seeded synthetic corpora for the five BASELINE.json configs.

No real corpus is available offline, so every test / bench input is generated here.
Pure numpy, vectorised, deterministic for a given (kind, n, seed).

  text(n, seed)        C1/C2 "enwik8-shaped": Zipf vocabulary, word-bigram dependence, light markup
  text(n, seed, drift) C3    "enwik9-shaped": + topic drift and occasional article-level repeats
  longrange(n, seed)   C4    mutated duplicated blocks at large distances (RK256 stress)
  mixed(n, seed)       C5    uniform random runs interleaved with structured records
"""
from __future__ import annotations

import numpy as np

_LETTERS = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
_LETTER_P = np.array([12.7, 9.1, 8.2, 7.5, 7.0, 6.7, 6.3, 6.1, 6.0, 4.3, 4.0, 2.8, 2.8, 2.4, 2.4, 2.2,
                      2.0, 2.0, 1.9, 1.5, 1.0, 0.8, 0.15, 0.15, 0.1, 0.07])
_LETTER_P = _LETTER_P / _LETTER_P.sum()

_MARKUP = [b"[[", b"]]", b"''", b"==", b"{{", b"}}", b"&amp;", b"&quot;", b"<ref>", b"</ref>", b"|",
           b"*", b"#", b"http://www.", b".com", b".org", b"<br>", b"&lt;", b"&gt;", b"\n\n", b"\n"]


def _vocabulary(rng: np.random.Generator, n_words: int):
    """Random 'words' (letters by English frequency), a few capitalised / numeric / markup tokens."""
    lens = np.clip(rng.poisson(4.2, n_words) + 1, 1, 14).astype(np.int64)
    # frequent words are short (Zipf's law of abbreviation)
    lens[:64] = np.clip(lens[:64] // 2, 1, 4)
    offs = np.zeros(n_words + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    blob = rng.choice(_LETTERS, size=int(offs[-1]), p=_LETTER_P)
    # capitalise ~8 % of the words
    cap = rng.random(n_words) < 0.08
    first = offs[:-1][cap]
    blob[first] = blob[first] - 32
    # numbers
    num = np.flatnonzero(rng.random(n_words) < 0.03)
    for w in num[:2000]:
        blob[offs[w]:offs[w + 1]] = rng.integers(48, 58, offs[w + 1] - offs[w])
    # markup tokens replace a handful of mid-frequency words
    words = [bytes(blob[offs[i]:offs[i + 1]]) for i in range(min(n_words, 400))]
    for j, m in enumerate(_MARKUP):
        words[20 + 7 * j] = m
    head = b"".join(words)
    head_lens = np.array([len(w) for w in words], dtype=np.int64)
    blob = np.concatenate([np.frombuffer(head, dtype=np.uint8), blob[offs[len(words)]:]])
    lens = np.concatenate([head_lens, lens[len(words):]])
    offs = np.zeros(n_words + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    return blob, offs, lens


def _zipf_cdf(n_words: int, s: float = 1.05):
    w = 1.0 / np.power(np.arange(1, n_words + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def _word_stream(rng, n_tokens, cdf, succ, follow_p, perm=None):
    """Order-2-ish Markov stream: with probability follow_p the next word is one of the 8 fixed
    successors of the previous word, otherwise an independent Zipf draw."""
    z = np.searchsorted(cdf, rng.random(n_tokens)).astype(np.int64)
    np.minimum(z, len(cdf) - 1, out=z)
    if perm is not None:
        z = perm[z]
    follow = rng.random(n_tokens) < follow_p
    follow[0] = False
    pick = rng.integers(0, succ.shape[1], n_tokens)
    w = z.copy()
    # resolve runs of `follow` flags left to right, one run position per pass (runs are short)
    pending = np.flatnonzero(follow)
    prev_resolved = ~follow
    for _ in range(64):
        if pending.size == 0:
            break
        ready = prev_resolved[pending - 1]
        idx = pending[ready]
        w[idx] = succ[w[idx - 1], pick[idx]]
        prev_resolved[idx] = True
        pending = pending[~ready]
    if pending.size:  # absurdly long run: cut it
        w[pending] = z[pending]
    return w


def _render(words, blob, offs, lens, rng):
    """Concatenate words with separators into a byte array."""
    seps = np.full(words.size, 32, dtype=np.uint8)
    r = rng.random(words.size)
    seps[r < 0.06] = ord(",")
    seps[r < 0.035] = ord(".")
    seps[r < 0.006] = ord("\n")
    wl = lens[words] + 1
    dst = np.zeros(words.size + 1, dtype=np.int64)
    np.cumsum(wl, out=dst[1:])
    total = int(dst[-1])
    src_start = offs[words]
    idx = np.repeat(src_start - dst[:-1], wl) + np.arange(total, dtype=np.int64)
    out = blob[np.minimum(idx, blob.size - 1)]
    out[dst[1:] - 1] = seps
    # ", " and ". " look more like prose: the separator is followed by the next word directly,
    # which is fine for match statistics.
    return out


def text(n: int, seed: int = 1, drift: bool = False, n_words: int = 50000) -> np.ndarray:
    """n bytes of English-like text (uint8 array)."""
    rng = np.random.default_rng([seed, 0x7E87])
    blob, offs, lens = _vocabulary(rng, n_words)
    cdf = _zipf_cdf(n_words)
    succ = np.searchsorted(cdf, rng.random((n_words, 8)) ** 1.3).astype(np.int64)
    np.minimum(succ, n_words - 1, out=succ)
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    piece = 8 << 20
    topic = None
    articles = []  # (start, length) of earlier spans, for repeats
    while pos < n:
        want = min(piece, n - pos)
        if drift:
            # topic drift: permute the mid/low-frequency part of the vocabulary per piece
            topic = np.arange(n_words, dtype=np.int64)
            lo = 200
            width = 6000
            start = int(rng.integers(lo, n_words - width))
            seg = topic[start:start + width].copy()
            rng.shuffle(seg)
            topic[lo:lo + width], topic[start:start + width] = seg, topic[lo:lo + width].copy()
        n_tok = int(want / 5.2) + 64
        w = _word_stream(rng, n_tok, cdf, succ, 0.55, topic)
        chunk = _render(w, blob, offs, lens, rng)
        while chunk.size < want:
            w = _word_stream(rng, n_tok // 4 + 64, cdf, succ, 0.55, topic)
            chunk = np.concatenate([chunk, _render(w, blob, offs, lens, rng)])
        out[pos:pos + want] = chunk[:want]
        if drift and pos > 0:
            # article-level repeats: copy a few earlier spans (2-64 KB) into this piece
            for _ in range(int(rng.integers(0, 3))):
                ln = int(rng.integers(2 << 10, 64 << 10))
                if ln >= want or pos < ln:
                    continue
                src = int(rng.integers(0, pos - ln + 1))
                dst = pos + int(rng.integers(0, want - ln))
                out[dst:dst + ln] = out[src:src + ln]
        pos += want
    return out


def longrange(n: int, seed: int = 3, max_dist: int | None = None) -> np.ndarray:
    """Base text blocks re-inserted at large distances with byte mutations and a few indels.

    Block sizes and distances scale with n so that small test inputs keep the same structure:
    blocks of n/512 .. n/8, distances up to max_dist (default ~n/2)."""
    rng = np.random.default_rng([seed, 0x10C6])
    if max_dist is None:
        max_dist = n // 2
    out = np.empty(n, dtype=np.uint8)
    base_len = max(n // 4, 4096)
    out[:base_len] = text(base_len, seed=seed + 100)
    pos = base_len
    rates = [5e-5, 5e-4, 2e-3]
    while pos < n:
        blk = int(rng.integers(max(n // 512, 512), max(n // 8, 1024)))
        blk = min(blk, n - pos)
        if rng.random() < 0.2:
            # fresh text
            out[pos:pos + blk] = text(blk, seed=int(rng.integers(1 << 30)))
        else:
            dist = int(rng.integers(blk, max(min(max_dist, pos), blk + 1) + 1))
            dist = min(dist, pos)
            src = pos - dist
            seg = out[src:src + blk].copy() if src + blk <= pos else np.resize(out[src:pos], blk)
            rate = rates[int(rng.integers(0, 3))]
            nm = rng.binomial(blk, rate)
            if nm:
                at = rng.integers(0, blk, nm)
                seg[at] = rng.integers(0, 256, nm, dtype=np.uint8)
            if rng.random() < 0.3 and blk > 64:
                # an insert/delete shift in the middle
                cut = int(rng.integers(16, blk - 16))
                k = int(rng.integers(1, 9))
                seg = np.concatenate([seg[:cut], seg[cut + k:], seg[:k]])
            out[pos:pos + blk] = seg[:blk]
        pos += blk
    return out


def _records(rng, n: int) -> np.ndarray:
    """Fixed-layout records: counters, timestamps, low-entropy fields, zero padding."""
    rec_len = int(rng.choice([64, 96, 128, 192, 256]))
    n_rec = n // rec_len + 1
    rec = np.zeros((n_rec, rec_len), dtype=np.uint8)
    ids = np.arange(n_rec, dtype=np.uint64) + np.uint64(rng.integers(1 << 20))
    ts = np.uint64(1_600_000_000) + np.cumsum(rng.integers(0, 5, n_rec)).astype(np.uint64)
    rec[:, 0:4] = np.frombuffer(b"REC\x01", dtype=np.uint8)
    for b in range(4):
        rec[:, 4 + b] = ((ids >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)        # LE counter
        rec[:, 8 + b] = ((ts >> np.uint64(8 * (3 - b))) & np.uint64(0xFF)).astype(np.uint8)   # BE timestamp
    rec[:, 12] = rng.integers(0, 4, n_rec)                  # low-entropy enum
    rec[:, 13] = rng.integers(0, 2, n_rec) * 255
    field = rng.integers(0, 16, (n_rec, 8)).astype(np.uint8) + 65
    rec[:, 16:24] = field
    if rec_len >= 96:
        rec[:, 32:48] = rng.integers(0, 256, (n_rec, 16), dtype=np.uint8) * (rng.random((n_rec, 1)) < 0.25)
    # rest stays zero (padding)
    return rec.reshape(-1)[:n]


def mixed(n: int, seed: int = 4) -> np.ndarray:
    """~50 % uniform random bytes, ~50 % structured records, interleaved in runs (1-8 MB at full
    size; n/256 .. n/32 for small n)."""
    rng = np.random.default_rng([seed, 0x5EED])
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    lo, hi = max(min(1 << 20, n // 256), 256), max(min(8 << 20, n // 32), 512)
    flip = bool(rng.integers(0, 2))
    while pos < n:
        run = min(int(rng.integers(lo, hi + 1)), n - pos)
        if flip:
            out[pos:pos + run] = rng.integers(0, 256, run, dtype=np.uint8)
        else:
            out[pos:pos + run] = _records(rng, run)
        flip = not flip
        pos += run
    return out


def make(kind: str, n: int, seed: int | None = None) -> np.ndarray:
    if kind == "text":
        return text(n, 1 if seed is None else seed)
    if kind == "text_drift":
        return text(n, 2 if seed is None else seed, drift=True)
    if kind == "longrange":
        return longrange(n, 3 if seed is None else seed)
    if kind == "mixed":
        return mixed(n, 4 if seed is None else seed)
    if kind == "random":
        return np.random.default_rng(5 if seed is None else seed).integers(0, 256, n, dtype=np.uint8)
    if kind == "zeros":
        return np.zeros(n, dtype=np.uint8)
    raise ValueError(kind)
