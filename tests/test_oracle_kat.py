"""Pins the CPU oracle against every literal known-answer vector available without a Go toolchain:
  * SURVEY.md App. A.3 — restatement of Go's encoding/csv reader_test table;
  * the reference's own literal KATs: TestRow (csvplus_test.go:49-116), TestIndexImpl (:198-246),
    TestErrors message strings (:808-909).
"""
import pytest

from oracle import oracle as orc

O = orc.Opts

# (input, opts, expected records or None, expected error)
KATS = [
    (b"a,b,c\n", O(), [[b"a", b"b", b"c"]], None),
    (b"a,b\r\nc,d\r\n", O(), [[b"a", b"b"], [b"c", b"d"]], None),
    (b"a,b\rc,d\r\n", O(), [[b"a", b"b\rc", b"d"]], None),
    (b"a,b,c", O(), [[b"a", b"b", b"c"]], None),
    (b"a;b;c\n", O(comma=";"), [[b"a", b"b", b"c"]], None),
    (b'"two\nline","one line","three\nline\nfield"', O(), [[b"two\nline", b"one line", b"three\nline\nfield"]], None),
    (b"a,b,c\n\nd,e,f\n\n", O(), [[b"a", b"b", b"c"], [b"d", b"e", b"f"]], None),
    (b" a,  b,   c\n", O(trim_leading_space=True), [[b"a", b"b", b"c"]], None),
    (b" a,  b,   c\n", O(), [[b" a", b"  b", b"   c"]], None),
    (b"#1,2,3\na,b,c\n#comment", O(comment="#"), [[b"a", b"b", b"c"]], None),
    (b"#1,2,3\na,b,c", O(), [[b"#1", b"2", b"3"], [b"a", b"b", b"c"]], None),
    (b'a "word","1"2",a","b', O(lazy_quotes=True), [[b'a "word"', b'1"2', b'a"', b"b"]], None),
    (b'a "word","1"2",a"', O(lazy_quotes=True), [[b'a "word"', b'1"2', b'a"']], None),
    (b'a""b,c', O(lazy_quotes=True), [[b'a""b', b"c"]], None),
    (b'a""b,c', O(), [], "bare_quote"),
    (b' "a"," b",c', O(trim_leading_space=True), [[b"a", b" b", b"c"]], None),
    (b'a "word","b"', O(), [], "bare_quote"),
    (b'"a word",b"', O(), [], "bare_quote"),
    (b'"a "word","b"', O(), [], "quote"),
    (b"a,b,c\nd,e", O(), [[b"a", b"b", b"c"]], "field_count"),
    (b"a,b,c\n\nd,e", O(), [[b"a", b"b", b"c"]], "field_count"),
    (b"a,b,c\nd,e", O(fields_per_record=-1), [[b"a", b"b", b"c"], [b"d", b"e"]], None),
    (b"a,b,c,", O(), [[b"a", b"b", b"c", b""]], None),
    (b"a,b,c,\n", O(), [[b"a", b"b", b"c", b""]], None),
    (b"a,b,c, ", O(trim_leading_space=True), [[b"a", b"b", b"c", b""]], None),
    (b"a,b,c, \n", O(trim_leading_space=True), [[b"a", b"b", b"c", b""]], None),
    (b"a,b,\nc,d,e", O(trim_leading_space=True), [[b"a", b"b", b""], [b"c", b"d", b"e"]], None),
    (b'x,,\nx,y,\nx,,z\n,,\n"x","",""\n"","",""', O(),
     [[b"x", b"", b""], [b"x", b"y", b""], [b"x", b"", b"z"], [b"", b"", b""], [b"x", b"", b""], [b"", b"", b""]], None),
    (b'A,"Hello\r\nHi",B\r\n', O(), [[b"A", b"Hello\nHi", b"B"]], None),
    (b"field1,field2\r", O(), [[b"field1", b"field2"]], None),
    (b'"field"\r', O(), [[b"field"]], None),
    (b'"field"\r\r', O(), [], "quote"),
    (b"field\rfield\r", O(), [[b"field\rfield"]], None),
    (b"field\r\rfield\r\r", O(), [[b"field\r\rfield\r"]], None),
    (b"field\r\r\nfield\r\r\n", O(), [[b"field\r"], [b"field\r"]], None),
    (b"field\r\r\n\rfield\r\r\n\r", O(), [[b"field\r"], [b"\rfield\r"]], None),
    (b"field\r\r\n\r\rfield\r\r\n\r\r", O(), [[b"field\r"], [b"\r\rfield\r"], [b"\r"]], None),
    (b"field1,field2\r\r\n\r\rfield1,field2\r\r\n\r\r,", O(),
     [[b"field1", b"field2\r"], [b"\r\rfield1", b"field2\r"], [b"\r\r", b""]], None),
    ("a£b,c£ \td,e\n€ comment\n".encode(), O(comma="£", comment="€", trim_leading_space=True),
     [[b"a", b"b,c", b"d,e"]], None),
    (b'"foo"bar"\r\n', O(), [], "quote"),
    (b'"foo"bar"\r\n', O(lazy_quotes=True), [[b'foo"bar']], None),
    (b'""""""""', O(), [[b'"""']], None),
    (b'"""""""', O(), [], "quote"),
    (b'"""""""', O(lazy_quotes=True), [[b'"""']], None),
    (b'"abc', O(), [], "quote"),
    (b'"abc', O(lazy_quotes=True), [[b"abc"]], None),
    (b'a,"b\nc"d,e', O(), [], "quote"),
    (b"", O(), [], None),
    (b"\n\n\r\n", O(), [], None),
    (b"\r", O(), [], None),
    (b"a,b\n", O(comma='"'), [], "invalid_delim"),
    (b"a,b\n", O(comma=",", comment=","), [], "invalid_delim"),
]


@pytest.mark.parametrize("i", range(len(KATS)))
def test_encoding_csv_kat(i):
    data, opts, want, err = KATS[i]
    recs, e = orc.csv_records(data, opts)
    assert e == err
    assert recs == want


def test_row_string_kat():
    # csvplus_test.go:50-54, :80, :112
    r = orc.take_rows([{"id": "12345", "Name": "John", "Surname": "Doe"}])
    assert r.row_string(0) == '{ "Name" : "John", "Surname" : "Doe", "id" : "12345" }'
    assert r.select("Name").row_string(0) == '{ "Name" : "John" }'
    # csvplus_test.go:85-103: Select / SelectValues missing column errors
    s = r.select("xxx", "zzz")
    assert s.error.endswith('missing column "xxx"')
    s = r.select("id", "zzz")
    assert s.error.endswith('missing column "zzz"')
    assert r.select("id").row_string(0) == '{ "id" : "12345" }'


def test_index_impl_kat():
    # csvplus_test.go:198-246
    rows = [
        {"x": "1", "y": "2", "z": "3", "junk": "zzz"}, {"x": "5", "y": "6", "z": "8", "junk": "nnn"},
        {"x": "0", "y": "5", "z": "3", "junk": "xxx"}, {"x": "8", "y": "9", "z": "1", "junk": "aaa"},
        {"x": "7", "y": "4", "z": "0", "junk": "bbb"}, {"x": "5", "y": "6", "z": "9", "junk": "iii"},
        {"x": "2", "y": "6", "z": "7", "junk": "mmm"},
    ]
    ix = orc.take_rows(rows).index_on("x", "y", "z")
    f = ix.find("1", "2", "3").to_dicts()
    assert f == [{b"x": b"1", b"y": b"2", b"z": b"3", b"junk": b"zzz"}]
    f = ix.find("5", "6", "8").to_dicts()
    assert len(f) == 1 and f[0][b"junk"] == b"nnn"
    f = ix.find("5", "6").to_dicts()
    assert len(f) == 2 and all(r[b"x"] == b"5" and r[b"y"] == b"6" for r in f)
    assert [r[b"x"] for r in ix.rows().to_dicts()] == [b"0", b"1", b"2", b"5", b"5", b"7", b"8"]


PEOPLE = b"id,name,surname,born\n" + b"".join(
    b"%d,%s,%s,%d\n" % (i * 3 + j, n, s, 1950 + i + j)
    for i, n in enumerate([b"Amelia", b"Olivia", b"Jack"]) for j, s in enumerate([b"Smith", b"Jones", b"Taylor"]))


def test_errors_kat():
    # csvplus_test.go:810-815
    r = orc.reader_rows(PEOPLE, select=["id", "name", "xxx"])
    assert r.error == "row 1: column not found: xxx"
    # :826-833
    src = orc.reader_rows(PEOPLE, select=["id", "name", "surname"])
    with pytest.raises(orc.OracleError, match='missing column "xxx" while creating an index'):
        src.index_on("name", "xxx")
    # :836-841
    with pytest.raises(orc.OracleError, match="duplicate value while creating unique index: { \"name\" : \"Amelia\" }"):
        src.unique_index_on("name")
    # :845-863 dedup to one row per name
    ix = src.index_on("name")
    ix.dedup("first")
    assert len(ix) == 3
    # :886-908
    r = orc.reader_rows(PEOPLE, expect={"name": 1, "surname": 3})
    assert r.error == 'row 1: misplaced column "surname": expected at pos. 3, but found at pos. 2'
    r = orc.reader_rows(PEOPLE, expect={"name": 1, "surname": 25})
    assert r.error == 'row 1: misplaced column "surname": expected at pos. 25, but found at pos. 2'


def test_dedup_tail_quirk():
    # trailing singleton after a duplicate group is lost (csvplus.go:851-864)
    def keys_after(keys):
        ix = orc.take_rows([{"k": k, "v": str(i)} for i, k in enumerate(keys)]).index_on("k")
        ix.dedup("first")
        return [r[b"k"].decode() for r in ix.rows().to_dicts()]
    assert keys_after("aab") == ["a"]
    assert keys_after("aabc") == ["a", "b"]
    assert keys_after("abbcdde") == ["a", "b", "c", "d"]
    assert keys_after("abb") == ["a", "b"]
    assert keys_after("abbcdd") == ["a", "b", "c", "d"]
    assert keys_after("abc") == ["a", "b", "c"]


def test_reader_semantics():
    # header = row 1; field count locked by header
    r = orc.reader_rows(b"a,b\n1,2\n3\n4,5\n")
    assert len(r) == 1 and r.error == "row 3: wrong number of fields"
    r = orc.reader_rows(b"a,b\n1,2\n\n3,\"x\"y\n")
    assert len(r) == 1 and r.error == 'row 3: extraneous or missing " in quoted-field'
    r = orc.reader_rows(b"")
    assert r.error == "row 1: EOF"
    r = orc.reader_rows(b"a,b\n1,2\n3\n", orc.Opts(fields_per_record=-1), select=["b"])
    assert r.values("b") == [b"2", b""] and r.error is None
    r = orc.reader_rows(b"1,2\n3,4\n", assume={"x": 0, "y": 1})
    assert r.values("y") == [b"2", b"4"]
    r = orc.reader_rows(b"1,2\n3,4\n", assume={"x": 0, "y": 5})
    assert r.error == 'row 1: column not found: "y" (5)'


def test_join_and_csv():
    cust = orc.reader_rows(b"id,name\n1,Ann\n2,Bob\n3,Cy\n").unique_index_on("id")
    orders = orc.reader_rows(b"oid,cust_id,qty\n10,2,5\n11,9,1\n12,1,7\n13,2,2\n")
    j = orders.join(cust, "cust_id")
    assert j.values("name") == [b"Bob", b"Ann", b"Bob"]
    assert j.values("oid") == [b"10", b"12", b"13"]
    out, err = j.to_csv("oid", "name", "qty")
    assert err is None and out == b"oid,name,qty\n10,Bob,5\n12,Ann,7\n13,Bob,2\n"
    out, err = orc.take_rows([{"a": 'x"y', "b": " lead"}, {"a": "", "b": "p,q"}]).to_csv("a", "b")
    assert out == b'a,b\n"x""y"," lead"\n,"p,q"\n'
    ex = orders.except_(cust, "cust_id")
    assert ex.values("oid") == [b"11"]


# Go encoding/csv Writer golden vectors (writer_test.go `writeTests`, the default-settings subset ToCsv uses: Comma ',',
# UseCRLF=false; go 1.23 per the reference's go.mod).  Each input record set is written through the oracle's ToCsv
# (csvplus.go:379-406) under column names c0..cN; the expected text is the header line + Go's expected output.
WRITER_KATS = [
    ([["abc"]], "abc\n"),
    ([['"abc"']], '"""abc"""\n'),
    ([['a"b']], '"a""b"\n'),
    ([['"a"b"']], '"""a""b"""\n'),
    ([[" abc"]], '" abc"\n'),
    ([["abc,def"]], '"abc,def"\n'),
    ([["abc", "def"]], "abc,def\n"),
    ([["abc"], ["def"]], "abc\ndef\n"),
    ([["abc\ndef"]], '"abc\ndef"\n'),
    ([["abc\rdef"]], '"abc\rdef"\n'),
    ([[""]], "\n"),
    ([["", ""]], ",\n"),
    ([["", "", ""]], ",,\n"),
    ([["", "", "a"]], ",,a\n"),
    ([["", "a", ""]], ",a,\n"),
    ([["", "a", "a"]], ",a,a\n"),
    ([["a", "", ""]], "a,,\n"),
    ([["a", "", "a"]], "a,,a\n"),
    ([["a", "a", ""]], "a,a,\n"),
    ([["a", "a", "a"]], "a,a,a\n"),
    ([["\\."]], '"\\."\n'),
]


@pytest.mark.parametrize("i", range(len(WRITER_KATS)))
def test_encoding_csv_writer_kat(i):
    recs, want = WRITER_KATS[i]
    names = ["c%d" % k for k in range(len(recs[0]))]
    out, err = orc.take_rows([dict(zip(names, r)) for r in recs]).to_csv(*names)
    assert err is None
    assert out == (",".join(names) + "\n" + want).encode()
