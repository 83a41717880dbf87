"""GPU parity of the fused parse (+select +filter) path against the CPU oracle, through the C ABI.

Mirrors the reference's TestSimpleDataSource / TestErrors (csvplus_test.go:118-151, :808-909) and adds
the encoding/csv edge cases the reference itself never exercises: quotes, "" escapes, CRLF,
blank lines, ragged rows, records crossing tile and window boundaries, error ordinals."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_table_equals_oracle, check_parity, gpu_ctx, people_csv, random_csv, run_both
from tests.test_oracle_kat import KATS

pytestmark = pytest.mark.gpu


def _default_opts(o):
    return not (o.comment or o.lazy_quotes or o.trim_leading_space)


@pytest.mark.parametrize("i", [i for i, k in enumerate(KATS) if _default_opts(k[1])])
def test_kat_vectors_through_reader(i):
    """every default-option vector of SURVEY App. A.3, first as Reader(header from row 1), then headerless"""
    data, opts, _, _ = KATS[i]
    check_parity(data, opts=opts)
    o2 = orc.Opts(comma=opts.comma, fields_per_record=-1)
    check_parity(data, opts=o2, assume={"a": 0, "b": 1, "c": 2, "d": 3})
    o3 = orc.Opts(comma=opts.comma, fields_per_record=opts.fields_per_record)
    check_parity(data, opts=o3, assume={"a": 0})


def test_simple_data_source():
    # csvplus_test.go:118-151: SelectColumns + Filter(Any(Like, Like))
    import csvplus_b200 as cp
    data = people_csv(120)
    hdr = sorted(["id", "name", "surname", "born"])
    src = cp.Take(cp.FromBytes(data).SelectColumns(*hdr)).Filter(cp.Any(cp.Like({"name": "Jack"}), cp.Like({"name": "Amelia"})))
    n = 0
    for row in src.ToRows():
        assert row["name"] in ("Jack", "Amelia")
        assert len(row) == 4 and sorted(row) == hdr
        n += 1
    assert n == 12 * 2


def test_reader_errors():
    # csvplus_test.go:810-815, :886-908 + App. A ordinals
    import csvplus_b200 as cp
    data = people_csv(120)
    with pytest.raises(cp.DataSourceError) as e:
        cp.Take(cp.FromBytes(data).SelectColumns("id", "name", "xxx")).ToRows()
    assert str(e.value).endswith("row 1: column not found: xxx")
    with pytest.raises(ValueError):  # duplicate column name panics, :818-823
        cp.FromBytes(data).SelectColumns("id", "name", "id")
    with pytest.raises(cp.DataSourceError) as e:
        cp.Take(cp.FromBytes(data).ExpectHeader({"name": 1, "surname": 3})).ToRows()
    assert str(e.value).endswith('row 1: misplaced column "surname": expected at pos. 3, but found at pos. 2')
    with pytest.raises(cp.DataSourceError) as e:
        cp.Take(cp.FromBytes(data).ExpectHeader({"name": 1, "surname": 25})).ToRows()
    assert str(e.value).endswith('row 1: misplaced column "surname": expected at pos. 25, but found at pos. 2')
    for bad in [b"a,b\n1,2\n3\n4,5\n", b'a,b\n1,2\n\n3,"x"y\n', b"", b'a,b\n1,2\n3,4"\n5,6\n', b'a,b\n"1,2\n3,4\n',
                b"a,b\n1,2,3\n", b"a\n\n\n\r\n"]:
        check_parity(bad)
    check_parity(b"a,b\n1,2\n3\n", opts=orc.Opts(fields_per_record=-1), select=["b"])
    check_parity(b"1,2\n3,4\n", assume={"x": 0, "y": 5})
    check_parity(b"1,2\n3,4\n", assume={"x": 0, "y": 1})
    check_parity(b"a,b\n1,2\n", opts=orc.Opts(fields_per_record=3))
    check_parity(b"a,b,c\n1,2\n", opts=orc.Opts(fields_per_record=3))
    check_parity(b"a,b\n1,2\n", opts=orc.Opts(comma="\n"))  # invalid delimiter


def test_rows_before_error_are_delivered():
    data = b"k,v\n" + b"".join(b"%d,x%d\n" % (i, i) for i in range(5000)) + b'5000,"bad"x\n' + b"5001,y\n" * 2000
    t, gerr, orows = run_both(data)
    assert str(gerr) == orows.error == 'row 5002: extraneous or missing " in quoted-field'
    assert_table_equals_oracle(t, orows)
    assert len(t) == 5000


@pytest.mark.parametrize("seed", range(8))
def test_random_adversarial_small(seed):
    data = random_csv(seed, nrows=400, ncols=4 + seed % 3, quoted_p=0.4, crlf_p=0.4, blank_p=0.2,
                      trailing_newline=seed % 2 == 0)
    check_parity(data)
    check_parity(data, select=["c2", "c0"])
    check_parity(data, select=["c1"], like={"c1": ""})


@pytest.mark.parametrize("seed", range(4))
def test_random_adversarial_multi_tile(seed):
    """several 32 KiB tiles, records crossing tile/window edges, quoted newlines spanning tiles"""
    data = random_csv(100 + seed, nrows=6000, ncols=5, quoted_p=0.35, long_p=0.004, crlf_p=0.3, blank_p=0.05)
    assert len(data) > 200_000
    check_parity(data)
    check_parity(data, select=["c4", "c1"])


@pytest.mark.parametrize("seed", range(3))
def test_random_ragged(seed):
    data = random_csv(200 + seed, nrows=3000, ncols=5, quoted_p=0.2, ragged_p=0.3)
    check_parity(data, opts=orc.Opts(fields_per_record=-1))
    check_parity(data, opts=orc.Opts(fields_per_record=-1), select=["c3", "c0"])
    check_parity(data)  # auto field count: first ragged row is the error


def test_error_inside_big_input_picks_first():
    good = random_csv(7, nrows=20000, ncols=4, quoted_p=0.3)
    bad1 = good + b'x,y"z,1,2\n' + random_csv(8, nrows=20000, ncols=4, quoted_p=0.3, header=False) + b'"open,1,2,3\n'
    check_parity(bad1)


def test_quote_parity_across_many_tiles():
    """a quoted field longer than several tiles: chain-1 parity carry and the slow path past the window"""
    big = b"A" * 150_000 + b'""' + b"\n" * 10 + b"B" * 50_000
    data = b'k,v\n1,"' + big + b'"\n2,plain\n3,"x\r\ny"\n' + b"".join(b"%d,z\n" % i for i in range(4, 3000))
    t, _ = check_parity(data)
    assert len(t) == 2999


def test_delimiter_and_filters():
    data = b"a;b;c\n1;x;y\n2;x;z\n3;w;y\n"
    check_parity(data, opts=orc.Opts(comma=";"), like={"b": "x"})
    check_parity(data, opts=orc.Opts(comma=";"), select=["c", "a"], like={"c": "y", "a": "3"})
    check_parity(data, opts=orc.Opts(comma=";"), select=["c", "a"], like={"zzz": "y"})  # missing column => false
    check_parity(data.replace(b";", b"\t"), opts=orc.Opts(comma="\t"))


def test_predicate_combinators():
    import csvplus_b200 as cp
    data = people_csv(120)
    o = orc.reader_rows(data)
    cases = [
        (cp.Any(cp.Like({"name": "Jack"}), cp.Like({"surname": "Smith"})), orc.Any(orc.Like({"name": "Jack"}), orc.Like({"surname": "Smith"}))),
        (cp.All(cp.Like({"name": "Jack"}), cp.Not(cp.Like({"surname": "Smith"}))), orc.All(orc.Like({"name": "Jack"}), orc.Not(orc.Like({"surname": "Smith"})))),
        (cp.Not(cp.Any(cp.Like({"name": "Jack"}), cp.Like({"name": "Ava", "surname": "Jones"}))),
         orc.Not(orc.Any(orc.Like({"name": "Jack"}), orc.Like({"name": "Ava", "surname": "Jones"})))),
        (cp.All(), orc.All()), (cp.Any(), orc.Any()),
    ]
    for gp, op in cases:
        t, err = cp.parse_csv(gpu_ctx(), data, pred=gp)
        assert err is None
        assert_table_equals_oracle(t, o.filter(op))
        # the same predicate on a materialised table (cpb_table_filter)
        t0, _ = cp.parse_csv(gpu_ctx(), data)
        assert_table_equals_oracle(t0.filter(gp), o.filter(op))


def test_synthetic_people_config1():
    """BASELINE config 1 shape at 200 k rows: Take(FromFile).Filter(Like{name: Amelia}).ToRows, all 6 columns"""
    import csvplus_b200 as cp
    ctx = gpu_ctx()
    buf = ctx.gen_csv("people", (0, 200_000))
    host = buf.to_host()
    t, err = cp.parse_csv(ctx, buf, pred=cp.Like({"name": "Amelia"}))
    assert err is None
    o = orc.reader_rows(host, pred=orc.Like({"name": "Amelia"}))
    assert 15_000 < len(o) < 25_000
    assert_table_equals_oracle(t, o)
    # config 2 shape: SelectColumns(name, surname, id) + Filter
    t, err = cp.parse_csv(ctx, buf, spec=[("name", -1), ("surname", -1), ("id", -1)], pred=cp.Like({"name": "Amelia"}))
    o = orc.reader_rows(host, select=["name", "surname", "id"], pred=orc.Like({"name": "Amelia"}))
    assert_table_equals_oracle(t, o)
    # host-memory entry (H2D inside the call) gives the same table
    t2, _ = cp.parse_csv(ctx, host, spec=[("name", -1), ("surname", -1), ("id", -1)], pred=cp.Like({"name": "Amelia"}))
    assert_table_equals_oracle(t2, o)


def test_full_size_properties():
    """BASELINE config 2 at 20 M rows (≈0.9 GB): size-independent properties instead of the oracle:
    row count conservation across a partition of predicates, and per-column byte checksums."""
    import csvplus_b200 as cp
    ctx = gpu_ctx()
    n = 20_000_000
    buf = ctx.gen_csv("people", (0, n))
    names = ["Amelia", "Olivia", "Emily", "Ava", "Isla", "Oliver", "Jack", "Harry", "Jacob", "Charlie"]
    total = 0
    for nm in names:
        t, err = cp.parse_csv(ctx, buf, spec=[("name", -1), ("id", -1)], pred=cp.Like({"name": nm}))
        assert err is None
        total += len(t)
        off, data = t.column("name", 0, min(len(t), 1000))
        assert data.tobytes() == nm.encode() * min(len(t), 1000)
    assert total == n
    t, err = cp.parse_csv(ctx, buf, spec=[("id", -1)])
    assert err is None and len(t) == n
    # ids are the row ordinals: total digits = sum of decimal lengths of 0..n-1
    nb = sum((min(n, 10 ** (d + 1)) - 10 ** d) * (d + 1) for d in range(8) if 10 ** d < n) + 1
    off, data = t.column("id", n - 3, n)
    assert data.tobytes() == b"%d%d%d" % (n - 3, n - 2, n - 1)
    import ctypes as C
    nbytes = C.c_uint64()
    ctx.lib.cpb_table_col_bytes(ctx.h, t.h, 0, 0, n, C.byref(nbytes))
    assert nbytes.value == nb


# ------------------------------------------------------------------ general reader options (parse_general.cu)
def _general_opts(o):
    return bool(o.comment or o.lazy_quotes or o.trim_leading_space)


@pytest.mark.parametrize("i", [i for i, k in enumerate(KATS) if _general_opts(k[1])])
def test_kat_vectors_general_options(i):
    """SURVEY App. A.3 vectors that use CommentChar / LazyQuotes / TrimLeadingSpace"""
    data, opts, _, _ = KATS[i]
    check_parity(data, opts=opts)
    check_parity(data, opts=orc.Opts(**{**opts.__dict__, "fields_per_record": -1}), assume={"a": 0, "b": 1, "c": 2, "d": 3})


def _mutate(data: bytes, seed: int, comments=False, stray_quotes=False, spaces=False) -> bytes:
    import random
    rng = random.Random(seed)
    out = bytearray()
    for ln in data.split(b"\n"):
        if comments and rng.random() < 0.15:
            out += b"#" + rng.choice([b' a "comment" line', b'"unbalanced', b"", b"x,y,z"]) + rng.choice([b"\n", b"\r\n"])
        if spaces and ln:
            ln = b",".join(rng.choice([b"", b" ", b"\t ", b"  "]) + f for f in ln.split(b","))
        if stray_quotes and ln and rng.random() < 0.3:
            p = rng.randrange(len(ln) + 1)
            ln = ln[:p] + b'"' + ln[p:]
        out += ln + b"\n"
    return bytes(out[:-1])


@pytest.mark.parametrize("seed", range(6))
def test_random_general_options(seed):
    base = random_csv(300 + seed, nrows=1500, ncols=4, quoted_p=0.35, crlf_p=0.3, blank_p=0.1)
    # comment lines (may contain quotes), default quote rules
    d1 = _mutate(base, seed, comments=True)
    check_parity(d1, opts=orc.Opts(comment="#"))
    check_parity(d1, opts=orc.Opts(comment="#"), select=["c2", "c0"])
    # lazy quotes: stray quotes anywhere never fail; field counts still checked
    d2 = _mutate(base, seed, stray_quotes=True)
    check_parity(d2, opts=orc.Opts(lazy_quotes=True, fields_per_record=-1))
    check_parity(d2, opts=orc.Opts(lazy_quotes=True))
    check_parity(d2, opts=orc.Opts())  # the same bytes under strict rules: first error must agree
    # leading white space, with and without trimming; all three options together
    d3 = _mutate(base, seed, spaces=True)
    check_parity(d3, opts=orc.Opts(trim_leading_space=True, fields_per_record=-1))
    check_parity(d3, opts=orc.Opts(fields_per_record=-1))
    d4 = _mutate(base, seed, comments=True, stray_quotes=True, spaces=True)
    check_parity(d4, opts=orc.Opts(comment="#", lazy_quotes=True, trim_leading_space=True, fields_per_record=-1), select=["c1", "c3"],
                 like={"c1": ""})
    check_parity(d4, opts=orc.Opts(comment="#", lazy_quotes=True, trim_leading_space=True))


def test_general_options_big_and_edges():
    big = random_csv(77, nrows=40000, ncols=5, quoted_p=0.3, long_p=0.002, crlf_p=0.2)
    check_parity(_mutate(big, 1, comments=True), opts=orc.Opts(comment="#"))
    for data in [b"", b"#only comment", b"#c\n\n#d\r\n", b"a,b\n#x\n1,2\n", b" a , b \n 1, \"q\" \n", b'a\n"x\r', b"\r", b"a,b\r\n\r\n#z\r\n1,2"]:
        for o in (orc.Opts(comment="#"), orc.Opts(lazy_quotes=True), orc.Opts(trim_leading_space=True),
                  orc.Opts(comment="#", lazy_quotes=True, trim_leading_space=True)):
            check_parity(data, opts=o)
            check_parity(data, opts=orc.Opts(**{**o.__dict__, "fields_per_record": -1}), assume={"x": 0, "y": 1})


def _route(fn):
    """runs fn with kernel statistics on; returns (result, True if the general multi-pass path ran)"""
    ctx = gpu_ctx()
    ctx.stats(enable=True, reset=True)
    try:
        r = fn()
        st = ctx.stats()
    finally:
        ctx.stats(enable=False)
    assert any(k.startswith("csv_scan") or k.startswith("general") for k in st), st
    return r, any(k.startswith("general_rec") for k in st)


def _clean_with_options(nrows=30000, seed=5, comments=True, spaces=True, crlf=True):
    """a quote-free file in the people shape with comment lines, leading white space and CRLF sprinkled in"""
    import random
    rng = random.Random(seed)
    out = [b"id,name,surname,born"]
    for ln in people_csv(nrows, seed).split(b"\n")[1:-1]:
        if comments and rng.random() < 0.05:
            out.append(b"#" + rng.choice([b" note", b"", b"x,y,z,w,v", b" , , ,"]))
        if spaces and rng.random() < 0.5:
            ln = b",".join(rng.choice([b"", b" ", b"\t", b" \t ", b"\r"]) + f for f in ln.split(b","))
        out.append(ln + (b"\r" if crlf and rng.random() < 0.3 else b""))
    return b"\n".join(out) + b"\n"


@pytest.mark.parametrize("opts", [dict(comment="#"), dict(trim_leading_space=True), dict(lazy_quotes=True),
                                  dict(comment="#", trim_leading_space=True, lazy_quotes=True),
                                  dict(comment="#", trim_leading_space=True, fields_per_record=-1)])
def test_reader_options_take_the_single_pass_scan_when_the_input_allows(opts):
    """CommentChar / TrimLeadingSpace / LazyQuotes on input that does not need the general path (no quote inside a comment
    line, no lazy quote): the fused scan runs, results equal the oracle (csvplus.go:971-993)"""
    data = _clean_with_options(comments="comment" in opts, spaces=True)
    o = orc.Opts(**opts)
    (_, general) = _route(lambda: check_parity(data, opts=o))
    assert not general
    (_, general) = _route(lambda: check_parity(data, opts=o, select=["surname", "id"], like={"surname": "Evans"}))
    assert not general
    # a few quoted fields (valid under the strict rules, leading white space before the quote when trimming): still one pass
    q = data.replace(b",Evans,", b', "Ev""ans",' if o.trim_leading_space else b',"Ev""ans",', 40)
    (_, general) = _route(lambda: check_parity(q, opts=o))
    assert not general


def test_reader_options_fall_back_to_the_general_path_when_they_must():
    data = _clean_with_options(nrows=20000, comments=True, spaces=False)
    # a comment line that holds a quote: the parity shortcut does not hold
    lines = data.split(b"\n")
    lines.insert(len(lines) // 2, b'# he said "hi')
    d1 = b"\n".join(lines)
    (_, general) = _route(lambda: check_parity(d1, opts=orc.Opts(comment="#")))
    assert general
    # a lazy quote: strict rules fail on it, LazyQuotes keeps it as data
    d2 = data.replace(b",Smith,", b',Sm"ith,', 3)
    (_, general) = _route(lambda: check_parity(d2, opts=orc.Opts(comment="#", lazy_quotes=True)))
    assert general
    check_parity(d2, opts=orc.Opts(comment="#"))  # strict: the first bare quote is the error, rows before it delivered
    # TrimLeadingSpace with a delimiter that is itself white space eats empty fields: general path
    d3 = b"a\tb\tc\n1\t\t3\n \t4\t5\t6\n"
    (_, general) = _route(lambda: check_parity(d3, opts=orc.Opts(comma="\t", trim_leading_space=True, fields_per_record=-1)))
    assert general


# ------------------------------------------------------------------ multi-byte reader runes (csrc/subst.cu)
UNISPACES = ["\u0085", "\u00a0", "\u1680", "\u2000", "\u2005", "\u200a", "\u2028", "\u2029", "\u202f", "\u205f", "\u3000"]


def test_every_kat_runs_on_the_gpu():
    """all vectors of the restated encoding/csv table take part in the two parametrised tests above (round 1 skipped the
    multi-byte Delimiter / CommentChar one)"""
    assert all(_default_opts(k[1]) or _general_opts(k[1]) for k in KATS)
    assert any(ord(k[1].comma) >= 0x80 for k in KATS)


@pytest.mark.parametrize("comma", ["£", "€", "\u00a0", "𝄞"])
def test_multibyte_delimiter(comma):
    import random
    rng = random.Random(ord(comma[0]))
    cb = comma.encode()
    rows = [cb.join([b"a", b"b", b"c"])]
    for i in range(3000):
        f = []
        for _ in range(3):
            kind = rng.randrange(6)
            if kind == 0:
                f.append(b'"q' + cb + b'uoted ""x"" ' + cb + b'"')       # the delimiter inside quotes is data
            elif kind == 1:
                f.append(b'"line\nbreak"')
            elif kind == 2:
                f.append(cb[:1] + b"x" if len(cb) > 1 else b"y")          # a lone lead byte is data
            else:
                f.append(bytes(rng.choice(b"abc,;\xc3\xa9 01") for _ in range(rng.randrange(0, 9))))
        rows.append(cb.join(f))
    data = b"\n".join(rows) + b"\n"
    check_parity(data, opts=orc.Opts(comma=comma))
    check_parity(data, opts=orc.Opts(comma=comma), select=["c", "a"], like={"a": "abc"})
    check_parity(data, opts=orc.Opts(comma=comma, lazy_quotes=True, fields_per_record=-1))


def test_multibyte_comment_and_unicode_spaces():
    import random
    rng = random.Random(9)
    sp = [s.encode() for s in UNISPACES] + [b" ", b"\t"]
    lines = [b"a,b,c"]
    for i in range(4000):
        if rng.random() < 0.1:
            lines.append("€ comment, with \"quotes".encode())
            continue
        f = []
        for _ in range(3):
            lead = b"".join(rng.choice(sp) for _ in range(rng.randrange(0, 3)))
            body = rng.choice([b"v%d" % i, b'"q,%d"' % i, b"x" + rng.choice(sp) + b"y", b"", rng.choice(sp)[:1] + b"z"])
            f.append(lead + body)
        lines.append(b",".join(f))
    data = b"\n".join(lines) + b"\n"
    o = orc.Opts(comment="€", trim_leading_space=True, fields_per_record=-1)
    check_parity(data, opts=o)
    check_parity(data, opts=o, select=["b"], like={"b": "x\u2028y"})
    check_parity(data, opts=orc.Opts(comment="€", fields_per_record=-1))            # untrimmed: the spaces are data
    check_parity(data, opts=orc.Opts(trim_leading_space=True, fields_per_record=-1))  # '€' lines are data now
    # a delimiter that is itself multi-byte together with everything else
    d2 = data.replace(b",", "£".encode())
    check_parity(d2, opts=orc.Opts(comma="£", comment="€", trim_leading_space=True, lazy_quotes=True, fields_per_record=-1))
