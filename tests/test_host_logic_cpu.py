"""Host-side logic of the API mirror that needs no GPU: argument validation (the reference's panics,
csvplus.go:381/:495/:513/:549/:592/:710/:715/:1000/:1005/:1022/:1041/:1048/:1281), predicate semantics and
their lowering to the cpb_pred tree, plan construction."""
import ctypes as C

import pytest

import csvplus_b200 as cp
from csvplus_b200 import _abi


def test_reader_option_panics():
    r = cp.FromBytes(b"a,b\n1,2\n")
    with pytest.raises(ValueError, match="empty header spec"):
        r.SelectColumns()
    with pytest.raises(ValueError, match="header spec: duplicate column name: a"):
        r.SelectColumns("a", "b", "a")
    with pytest.raises(ValueError, match="Empty header spec"):
        r.AssumeHeader({})
    with pytest.raises(ValueError, match="header spec: negative index for column x"):
        r.AssumeHeader({"x": -1})
    with pytest.raises(ValueError, match="empty header spec"):
        r.ExpectHeader({})
    assert r.Delimiter(";").CommentChar("#").LazyQuotes().TrimLeadingSpace().NumFieldsAny() is r
    assert (r.delimiter, r.comment, r.lazyQuotes, r.trimLeadingSpace, r.numFields) == (";", "#", True, True, -1)
    assert r.NumFieldsAuto().numFields == 0 and r.NumFields(7).numFields == 7
    r.ExpectHeader({"a": 0, "b": -1})
    assert r.headerFromFirstRow and r.header == {"a": 0, "b": -1}
    r.AssumeHeader({"a": 0})
    assert not r.headerFromFirstRow


def test_datasource_argument_panics():
    src = cp.Take(cp.FromBytes(b"a,b\n1,2\n"))
    with pytest.raises(ValueError, match=r"no columns specified in SelectColumns\(\)"):
        src.SelectColumns()
    with pytest.raises(ValueError, match=r"no columns specified in DropColumns\(\)"):
        src.DropColumns()
    with pytest.raises(ValueError, match=r"empty column list in ToCsv\(\) function"):
        src.ToCsv(None)
    with pytest.raises(ValueError, match=r"empty column list in CreateIndex\(\)"):
        src.IndexOn()
    with pytest.raises(ValueError, match=r"duplicate column name\(s\) in CreateIndex\(\)"):
        src.UniqueIndexOn("a", "a")
    with pytest.raises(ValueError, match=r"empty match row in Like\(\) predicate"):
        cp.Like({})

    class FakeIndex:
        columns = ["k"]
    with pytest.raises(ValueError, match=r"too many source columns in Join\(\)"):
        src.Join(FakeIndex(), "a", "b")
    with pytest.raises(ValueError, match=r"too many source columns in Except\(\)"):
        src.Except(FakeIndex(), "a", "b")
    with pytest.raises(TypeError):
        cp.Take(42)


def test_plans_are_lazy_and_immutable():
    src = cp.Take(cp.FromBytes(b"a,b\n1,2\n"))
    a = src.Filter(cp.Like({"a": "1"}))
    b = a.SelectColumns("b").Top(3).Drop(1)
    assert len(src._ops) == 0 and len(a._ops) == 1 and [o[0] for o in b._ops] == ["filter", "select", "top", "drop"]


def test_predicate_semantics_match_reference():
    row = {"name": "Amelia", "surname": "Smith"}
    assert cp.Like({"name": "Amelia"})(row) and not cp.Like({"name": "Amelia", "xxx": "1"})(row)  # missing column => false
    assert cp.All(cp.Like({"name": "Amelia"}), cp.Like({"surname": "Smith"}))(row)
    assert not cp.All(cp.Like({"name": "Amelia"}), cp.Like({"surname": "Jones"}))(row)
    assert cp.Any(cp.Like({"name": "x"}), cp.Like({"surname": "Smith"}))(row) and not cp.Any()(row) and cp.All()(row)
    assert cp.Not(cp.Like({"name": "x"}))(row)
    opaque = cp.Any(cp.Like({"name": "x"}), lambda r: r["surname"] == "Smith")  # mixing in a Python closure => opaque
    assert not isinstance(opaque, cp.Predicate) and opaque(row)
    assert cp.Not(lambda r: False)(row)


def test_predicate_lowering_to_abi_tree():
    p = cp.Not(cp.Any(cp.Like({"name": "Jack"}), cp.All(cp.Like({"name": "Ava", "surname": "Jones"}))))
    assert p.lowerable()
    keep = []
    c = p._c(keep)
    assert c.op == 3 and c.n == 1
    any_ = c.children[0].contents
    assert any_.op == 2 and any_.n == 2
    like = any_.children[0].contents
    assert like.op == 0 and like.n == 1 and C.string_at(like.keys[0].ptr, like.keys[0].len) == b"name"
    assert C.string_at(like.values[0].ptr, like.values[0].len) == b"Jack"
    all_ = any_.children[1].contents
    assert all_.op == 1 and all_.children[0].contents.n == 2
    assert C.sizeof(_abi.Pred) == 32


def test_error_types_format_like_reference():
    e = cp.DataSourceError(7, "wrong number of fields")
    assert str(e) == "row 7: wrong number of fields" and e.Line == 7


def test_go_json_string_escaping():
    """ToJSON writes strings the way encoding/json does with SetEscapeHTML(false) (csvplus.go:446-474)"""
    from csvplus_b200.api import _go_json_string as g
    assert g('plain') == b'"plain"'
    assert g('q"b\\s') == b'"q\\"b\\\\s"'
    assert g("a\nb\rc\td\x00e\x1f") == b'"a\\nb\\rc\\td\\u0000e\\u001f"'
    assert g("<&>") == b'"<&>"'                               # no HTML escaping
    assert g(" x ") == b'"\\u2028x\\u2029"'          # always escaped
    assert g("é€𝄞") == "\"é€𝄞\"".encode()                      # valid UTF-8 verbatim
    assert g(b"\xff\xc3(".decode("utf-8", "surrogateescape")) == b'"\\ufffd\\ufffd("'  # every invalid byte -> U+FFFD
    assert g(b"\xe2\x82".decode("utf-8", "surrogateescape")) == b'"\\ufffd\\ufffd"'    # truncated sequence
