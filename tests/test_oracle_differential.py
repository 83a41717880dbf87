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

ALPHABET = [b"a", b"b", b",", b",", b'"', b'"', b"\n", b"\n", b"\r", b"\r\n", b" ", b"#", b"\t", b'""', b'",', b'"\n']
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
