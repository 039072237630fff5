"""Shared parity inputs, all small enough for seconds on CPU."""
import numpy as np

from tilespmv_b200 import generators as g


def _unsorted(case):
    """Same matrix with the columns of every row reversed: exercises 'in-tile order is input order'."""
    m, n, rp, ci, v = case
    ci2, v2 = ci.copy(), v.copy()
    for i in range(m):
        lo, hi = rp[i], rp[i + 1]
        ci2[lo:hi] = ci[lo:hi][::-1]
        v2[lo:hi] = v[lo:hi][::-1]
    return m, n, rp, ci2, v2


def _ragged(case, rowA, colA):
    """Crop to rowA x colA (not multiples of 16): partial last block row / tile column."""
    m, n, rp, ci, v = case
    keep = np.zeros(len(ci), bool)
    new_rp = np.zeros(rowA + 1, np.int32)
    for i in range(rowA):
        sel = ci[rp[i]:rp[i + 1]] < colA
        keep[rp[i]:rp[i + 1]] = sel
        new_rp[i + 1] = new_rp[i] + sel.sum()
    return rowA, colA, new_rp, ci[keep], v[keep]


def _empty_rows(m=64, n=64):
    """Only rows 17 and 40 populated; whole block rows and the matrix tail are empty."""
    rp = np.zeros(m + 1, np.int32)
    cols, vals = [], []
    for i in range(m):
        if i == 17:
            cols += [0, 5, 33, 63]
        if i == 40:
            cols += list(range(16, 48))
        rp[i + 1] = len(cols)
    return m, n, rp, np.array(cols, np.int32), (np.arange(len(cols)) % 10).astype(np.float64)


def hyb_rich(nbr=48, ntc=40, rowA=None, colA=None, seed=11, val_mode=1):
    """Tiles shaped to hit the reference's dormant HYB rule (csr2tile.h:279-316: row-length variation >= 1.0
    and <= 4 entries past the I/O-optimal ELL width) next to near misses, ELL, CSR and COO tiles."""
    rng = np.random.default_rng(seed)
    rowA = nbr * 16 if rowA is None else rowA
    colA = ntc * 16 if colA is None else colA
    rows = [dict() for _ in range(rowA)]  # global row -> {col: 1}
    for b in range(nbr):
        rowlen = min(16, rowA - b * 16)
        if rowlen <= 0:
            break
        for tc in rng.choice(ntc, size=min(ntc, 6), replace=False):
            collen = min(16, colA - tc * 16)
            if collen <= 0:
                continue
            kind = rng.integers(0, 5)
            cells = set()
            if kind <= 1:  # k rows with one entry + up to 5 extras in one or two rows
                k = int(rng.integers(min(8, rowlen), rowlen + 1))
                for r in rng.choice(rowlen, size=k, replace=False):
                    cells.add((int(r), int(rng.integers(collen))))
                hot = rng.choice(rowlen, size=int(rng.integers(1, 3)), replace=False)
                for _ in range(int(rng.integers(0, 6))):
                    cells.add((int(rng.choice(hot)), int(rng.integers(collen))))
            elif kind == 2:  # scattered
                for _ in range(int(rng.integers(13, 48))):
                    cells.add((int(rng.integers(rowlen)), int(rng.integers(collen))))
            elif kind == 3:  # even rows: ELL
                w = int(rng.integers(1, 4))
                for r in range(rowlen):
                    for c in rng.choice(collen, size=min(w, collen), replace=False):
                        cells.add((r, int(c)))
            else:  # very sparse: COO
                for _ in range(int(rng.integers(1, 12))):
                    cells.add((int(rng.integers(rowlen)), int(rng.integers(collen))))
            for r, c in cells:
                rows[b * 16 + r][tc * 16 + c] = 1
    rp = np.zeros(rowA + 1, np.int32)
    cols = []
    for i, d in enumerate(rows):
        cols.extend(sorted(d))
        rp[i + 1] = len(cols)
    ci = np.array(cols, np.int32)
    if val_mode == 1:
        v = (np.arange(len(ci)) % 10).astype(np.float64)
    else:
        v = rng.uniform(-1, 1, len(ci))
    return rowA, colA, rp, ci, v


# inputs for the non-default HYB rule (TILESPMV_ENABLE_HYB), checked against the reference built with that rule
HYB_CASES = {
    "hyb_rich": lambda: hyb_rich(),
    "hyb_rich_real": lambda: hyb_rich(seed=12, val_mode=0),
    "hyb_ragged": lambda: hyb_rich(nbr=12, ntc=12, rowA=12 * 16 - 3, colA=12 * 16 - 5, seed=19),  # a HYB tile in the 13-row last block row
    "rmat_12": lambda: g.rmat(12, val_mode=1),
    "rmat_10_real": lambda: g.rmat(10, val_mode=0),
    "seven_formats": lambda: g.seven_formats(),
    "banded_2k_real": lambda: g.banded(2048, val_mode=0),
}

def _hub_rows(m=256, n=8192, seed=21):
    """A few very long rows among very sparse ones: row 3 has 5000 entries (cut into pieces, every piece one
    long local row), row 100 has 40 (long, inside a whole block row), row 101 has 31 (just short), the rest 1-2."""
    rng = np.random.default_rng(seed)
    rp = np.zeros(m + 1, np.int32)
    cols = []
    for i in range(m):
        k = {3: 5000, 100: 40, 101: 31, 200: 700}.get(i, int(rng.integers(1, 3)))
        cols.extend(sorted(rng.choice(n, size=k, replace=False).tolist()))
        rp[i + 1] = len(cols)
    ci = np.array(cols, np.int32)
    return m, n, rp, ci, rng.uniform(-1, 1, len(ci))


CASES = {
    "hub_rows": _hub_rows,
    "seven_formats": lambda: g.seven_formats(),
    "lap2d_64": lambda: g.lap2d(64, val_mode=1),
    "lap3d27_24": lambda: g.lap3d27(24, val_mode=1),
    "banded_8k": lambda: g.banded(8192, val_mode=1),
    "banded_8k_real": lambda: g.banded(8192, val_mode=0),
    "band_contig_8k": lambda: g.band_contig(8192, val_mode=1),
    "rmat_12": lambda: g.rmat(12, val_mode=1),
    "rmat_12_real": lambda: g.rmat(12, val_mode=0),
    "uniform_8k": lambda: g.uniform(8192, val_mode=1),
    "banded_2k_real": lambda: g.banded(2048, val_mode=0),
    "rmat_10_real": lambda: g.rmat(10, val_mode=0),
    "uniform_2k": lambda: g.uniform(2048, val_mode=1),
    "band_unsorted": lambda: _unsorted(g.band_contig(2048, hb=20, val_mode=0)),
    "rmat_unsorted": lambda: _unsorted(g.rmat(10, val_mode=0)),
    "ragged_band": lambda: _ragged(g.band_contig(1024, hb=18, val_mode=0), 1003, 1001),
    "ragged_seven": lambda: _ragged(g.seven_formats(), 29, 37),
    "ragged_rmat": lambda: _ragged(g.rmat(10, val_mode=1), 1000, 999),
    "empty_rows": _empty_rows,
    "empty_matrix": lambda: (32, 32, np.zeros(33, np.int32), np.zeros(0, np.int32), np.zeros(0)),
    "dense_48": lambda: g.band_contig(48, hb=48, val_mode=0),
}

# larger inputs: used by the oracle-vs-reference test (CPU) and the full-size GPU tests
BIG_CASES = {
    "lap2d_256": lambda: g.lap2d(256, val_mode=1),
    "lap3d27_48": lambda: g.lap3d27(48, val_mode=1),
    "banded_64k": lambda: g.banded(65536, val_mode=0),
    "band_contig_64k": lambda: g.band_contig(65536, val_mode=1),
    "rmat_15": lambda: g.rmat(15, val_mode=0),
    "uniform_64k": lambda: g.uniform(65536, val_mode=1),
}


def random_case(rng):
    """A random CSR matrix mixing the structures that select different tile formats: dense blocks, full rows / columns
    inside a tile, bands, scattered entries, hub rows, empty rows; dimensions are not multiples of 16."""
    m, n = int(rng.integers(1, 400)), int(rng.integers(1, 400))
    A = np.zeros((m, n), bool)
    for _ in range(int(rng.integers(0, 6))):  # dense-ish blocks
        r0, c0 = int(rng.integers(0, m)), int(rng.integers(0, n))
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        A[r0:r0 + h, c0:c0 + w] |= rng.random((min(h, m - r0), min(w, n - c0))) < rng.choice([0.3, 0.8, 1.0])
    for _ in range(int(rng.integers(0, 4))):  # full rows / columns (DenseRow / DenseCol tiles)
        if rng.random() < 0.5:
            A[int(rng.integers(0, m)), :] = True
        else:
            A[:, int(rng.integers(0, n))] = True
    if rng.random() < 0.5:  # band
        hb = int(rng.integers(1, 24))
        i, j = np.indices((m, n))
        A |= (np.abs(i - j) <= hb) & (rng.random((m, n)) < rng.choice([0.4, 1.0]))
    A |= rng.random((m, n)) < rng.choice([0.0, 0.002, 0.02, 0.1])  # scattered
    if rng.random() < 0.3:
        A[int(rng.integers(0, m)):, :] = False  # empty tail
    rows, cols = np.nonzero(A)
    rp = np.zeros(m + 1, np.int32)
    np.add.at(rp, rows + 1, 1)
    rp = np.cumsum(rp).astype(np.int32)
    v = rng.uniform(-1, 1, len(cols))
    return m, n, rp, cols.astype(np.int32), v


def x_for(n, mode, dtype=np.float64, seed=7):
    """mode 1: x[i] = i % 10 like main.cu:93-97; mode 0: seeded uniform(-1,1)."""
    if mode == 1:
        return (np.arange(n) % 10).astype(dtype)
    return np.random.default_rng(seed).uniform(-1, 1, n).astype(dtype)
