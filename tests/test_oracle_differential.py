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
