"""Synthetic inputs of the BASELINE.md configurations: a counter-based generator (SplitMix64 of
seed and position), i.i.d. uniform ACGT, upper case; foreground sequences get one of 8 fixed 8-mers
planted at a uniform position with probability 1/2."""
import numpy as np

_MOTIFS = [b"ACGTGCAT", b"TTGACGCA", b"GGATCCAA", b"CATTAGCG", b"AGCTTCGA", b"TGCAGGTC", b"CCGATAAG", b"GTACCTGA"]
_LETTERS = np.frombuffer(b"ACGT", dtype=np.uint8)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def random_bases(n_bases, seed, offset=0):
    """uint8 ASCII array of n_bases letters; base i depends only on (seed, offset + i)"""
    out = np.empty(n_bases, dtype=np.uint8)
    chunk = 1 << 24
    with np.errstate(over="ignore"):
        salt = np.uint64(seed) * np.uint64(0x632BE59BD9B4E019)
        for a in range(0, n_bases, chunk):
            b = min(n_bases, a + chunk)
            idx = np.arange(offset + a, offset + b, dtype=np.uint64)
            key = _splitmix64(idx * np.uint64(2) + salt)
            out[a:b] = _LETTERS[(key >> np.uint64(62)).astype(np.int64)]
    return out


def sequences(n, length, seed, planted=False, first=0):
    """(buffer, offsets) of sequences first .. first+n-1 of the stream with this seed"""
    buf = random_bases(n * length, seed, offset=first * length)
    if planted and length >= 8 and n > 0:
        with np.errstate(over="ignore"):
            ids = np.arange(first, first + n, dtype=np.uint64)
            r = _splitmix64(ids * np.uint64(7) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03) + np.uint64(1))
        plant = (r & np.uint64(1)) == np.uint64(1)
        motif = ((r >> np.uint64(1)) & np.uint64(7)).astype(np.int64)
        pos = ((r >> np.uint64(8)) % np.uint64(length - 8 + 1)).astype(np.int64)
        m = np.stack([np.frombuffer(x, dtype=np.uint8) for x in _MOTIFS])
        rows = np.nonzero(plant)[0]
        if len(rows):
            base = rows * length + pos[rows]
            buf[(base[:, None] + np.arange(8)[None, :]).ravel()] = m[motif[rows]].ravel()
    off = np.arange(n + 1, dtype=np.int64) * length
    return buf, off


def training_set(n_fg, n_bg, length, first_fg=0, first_bg=0):
    """fg (seed 1, planted) then bg (seed 2): (buffer, offsets, labels)"""
    fb, fo = sequences(n_fg, length, 1, planted=True, first=first_fg)
    bb, bo = sequences(n_bg, length, 2, planted=False, first=first_bg)
    buf = np.concatenate([fb, bb])
    off = np.concatenate([fo, bo[1:] + fo[-1]])
    labels = np.concatenate([np.ones(n_fg, dtype=np.uint8), np.zeros(n_bg, dtype=np.uint8)])
    return buf, off, labels
