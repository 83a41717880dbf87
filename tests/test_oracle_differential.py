"""Differential fuzz of the C++ oracle's encoding/csv restatement (oracle/oracle.cpp) against the independent
pure-Python restatement in tests/pyref_csv.py: random byte soups over the characters that drive the reader's state
machine, under every combination of LazyQuotes / TrimLeadingSpace / Comment / FieldsPerRecord.  Two restatements
written differently agreeing on 3.6 x 10^4 adversarial inputs is the strongest pin of the oracle available without a Go
toolchain."""
import itertools
import random

import pytest

from oracle import oracle as orc
from tests.pyref_csv import Reader

ALPHABET = [b"a", b"b", b",", b",", b'"', b'"', b"\n", b"\n", b"\r", b"\r\n", b" ", b"#", b"\t", b'""', b'",', b'"\n',
            b"\xc3\xa9", b"\xe2\x82\xac", b"\xff", b"\x00"]  # multi-byte UTF-8, an invalid byte and NUL are ordinary field bytes
COMBOS = list(itertools.product([False, True], [False, True], ["", "#"], [0, -1, 2]))


@pytest.mark.parametrize("ci", range(len(COMBOS)))
def test_oracle_reader_matches_python_restatement(ci):
    lazy, trim, comment, fpr = COMBOS[ci]
    rng = random.Random(1000 + ci)
    o = orc.Opts(comment=comment, fields_per_record=fpr, lazy_quotes=lazy, trim_leading_space=trim)
    for it in range(1500):
        n = rng.randrange(0, 40)
        data = b"".join(rng.choice(ALPHABET) for _ in range(n))
        want = Reader(data, comment=comment.encode(), fields_per_record=fpr, lazy_quotes=lazy, trim_leading_space=trim).read_all()
        got = orc.csv_records(data, o)
        assert got == want, (data, o, got, want)


# ------------------------------------------------------------------ index / find / join / except (csvplus.go:529-608, :707-807, :870-920)
def _py_sorted(rows, cols):
    return sorted(rows, key=lambda r: tuple(r[c].encode() for c in cols))  # strings.Compare column by column, stable


def _py_join(probe, idx_sorted, idx_cols, on, anti=False):
    out = []
    for p in probe:
        key = [p[c] for c in on]
        hits = [r for r in idx_sorted if [r[c] for c in idx_cols[:len(key)]] == key]  # run of equal prefixes, sorted order
        if anti:
            if not hits:
                out.append(dict(p))
        else:
            out += [{**r, **p} for r in hits]  # mergeRows: the probe row's values win name collisions
    return out


def _dump(rows, cols):
    # same quoting rule as csv.Writer for the characters the generator uses (none need quotes)
    return ("\n".join([",".join(cols)] + [",".join(r[c] for c in cols) for r in rows]) + "\n").encode()


@pytest.mark.parametrize("seed", range(12))
def test_oracle_index_join_match_python_restatement(seed):
    rng = random.Random(seed)
    word = lambda: "".join(rng.choice("ab0") for _ in range(rng.randrange(0, 4)))
    nkeys = rng.choice([1, 2, 3])
    kcols = ["k%d" % i for i in range(nkeys)]
    irows = [{**{k: word() for k in kcols}, "v": str(i), "shared": "i%d" % i} for i in range(rng.randrange(1, 60))]
    non = rng.randrange(1, nkeys + 1)  # join on a prefix of the key columns
    pcols = ["p%d" % i for i in range(non)]
    prows = [{**{c: word() for c in pcols}, "pid": str(j), "shared": "p%d" % j} for j in range(rng.randrange(1, 80))]
    oi = orc.take_rows(irows).index_on(*kcols)
    srt = _py_sorted(irows, kcols)
    assert oi.rows().to_csv(*kcols, "v")[0] == _dump(srt, kcols + ["v"])
    op = orc.take_rows(prows)
    allcols = kcols + ["v"] + pcols + ["pid", "shared"]
    assert op.join(oi, *pcols).to_csv(*allcols)[0] == _dump(_py_join(prows, srt, kcols, pcols), allcols)
    assert op.except_(oi, *pcols).to_csv(*pcols, "pid")[0] == _dump(_py_join(prows, srt, kcols, pcols, anti=True), pcols + ["pid"])
    # Find: rows whose key prefix equals the values (csvplus.go:625-627, :870-891)
    vals = [rng.choice(irows)[k] for k in kcols[:non]]
    want = [r for r in srt if [r[k] for k in kcols[:non]] == vals]
    assert oi.find(*vals).to_csv(*kcols, "v")[0] == _dump(want, kcols + ["v"])


# ------------------------------------------------------------------ ResolveDuplicates (csvplus.go:810-867), including §Q1
def _py_dedup(rows, cols, resolve):
    """index-based restatement of indexImpl.dedup; resolve(group) -> row or None ("empty row": drop the group).
    Keeps the reference's quirk: the scan for the next duplicate copies rows[lower-1] only while lower < len(rows),
    so a trailing run of singletons loses its last row whenever at least one duplicate group was resolved."""
    key = lambda r: [r[c] for c in cols]
    rows = list(rows)
    n = len(rows)
    lower = 1
    while lower < n and key(rows[lower - 1]) != key(rows[lower]):
        lower += 1
    if lower >= n:
        return rows
    dest = lower - 1
    while lower < n:
        upper = lower
        while upper < n and key(rows[upper]) == key(rows[lower]):
            upper += 1
        row = resolve(rows[lower - 1: upper])
        lower = upper + 1
        if row is not None:
            rows[dest] = row; dest += 1
        while lower < n:
            if key(rows[lower - 1]) == key(rows[lower]):
                break
            rows[dest] = rows[lower - 1]
            lower += 1; dest += 1
    return rows[:dest]


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("mode", ["min", "drop", "first"])
def test_oracle_dedup_matches_python_restatement(seed, mode):
    rng = random.Random(100 + seed)
    rows = [{"k": rng.choice("abcdefgh"[: rng.randrange(1, 9)]), "w": "%02d" % rng.randrange(50), "n": str(i)} for i in range(rng.randrange(1, 40))]
    oi = orc.take_rows(rows).index_on("k")
    oi.dedup(mode, "w")
    srt = _py_sorted(rows, ["k"])
    resolve = {"drop": lambda g: None, "first": lambda g: g[0],
               "min": lambda g: min(g, key=lambda r: r["w"].encode())}[mode]  # min() keeps the first of equal minima, like the oracle's strict <
    want = _py_dedup(srt, ["k"], resolve)
    assert oi.rows().to_csv("k", "w", "n")[0] == _dump(want, ["k", "w", "n"])
