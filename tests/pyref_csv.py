"""A second, independent restatement of Go's encoding/csv Reader (go 1.23: readLine / readRecord / ReadAll, SURVEY
App. A) in plain Python, written from the published algorithm and structured differently from oracle/oracle.cpp
(slice-based, like the Go source).  TEST INFRASTRUCTURE ONLY: tests/test_oracle_differential.py fuzzes the C++ oracle
against it; nothing in the product imports it.  ASCII comma/comment only (the only case the CUDA path takes)."""

SPACES = b" \t\n\v\f\r"  # unicode.IsSpace restricted to ASCII (+ U+0085/U+00A0 are multi-byte in UTF-8: not produced by the fuzzer)


def _length_nl(b: bytes) -> int:
    return 1 if b.endswith(b"\n") else 0


class Reader:
    def __init__(self, data: bytes, comma=b",", comment=b"", fields_per_record=0, lazy_quotes=False, trim_leading_space=False):
        self.data, self.pos = data, 0
        self.comma, self.comment = comma, comment
        self.fpr, self.lazy, self.trim = fields_per_record, lazy_quotes, trim_leading_space

    def read_line(self):
        """-> (line, eof).  bufio.ReadSlice('\\n') + the trailing-\\r-before-EOF and \\r\\n -> \\n rules."""
        if self.pos >= len(self.data):
            return b"", True
        j = self.data.find(b"\n", self.pos)
        eof = j < 0
        line = self.data[self.pos:] if eof else self.data[self.pos: j + 1]
        self.pos += len(line)
        if eof and line.endswith(b"\r"):  # drop trailing \r before EOF (len(line) > 0 here)
            line = line[:-1]
        if len(line) >= 2 and line.endswith(b"\r\n"):
            line = line[:-2] + b"\n"
        # Go returns err == nil for a non-empty last line without newline; the NEXT call reports io.EOF
        return line, False

    def read_record(self):
        """-> (fields or None at EOF, error name or None)"""
        while True:
            line, eof = self.read_line()
            if eof:
                return None, None
            if self.comment and line[:1] == self.comment:
                continue
            if len(line) == _length_nl(line):
                continue
            break
        fields, err = [], None
        err_read_eof = False
        while True:  # parseField
            if self.trim:
                i = 0
                while i < len(line) and line[i] in SPACES:
                    i += 1
                line = line[i:]  # (i == len(line): the whole rest, newline included, was white space)
            if len(line) == 0 or line[:1] != b'"':
                i = line.find(self.comma)
                field = line[:i] if i >= 0 else line[: len(line) - _length_nl(line)]
                if not self.lazy and b'"' in field:
                    err = "bare_quote"
                    break
                fields.append(field)
                if i >= 0:
                    line = line[i + 1:]
                    continue
                break
            else:
                line = line[1:]
                buf = b""
                done_field = False
                while True:
                    i = line.find(b'"')
                    if i >= 0:
                        buf += line[:i]
                        line = line[i + 1:]
                        if line[:1] == b'"':
                            buf += b'"'
                            line = line[1:]
                        elif line[:1] == self.comma:
                            line = line[1:]
                            fields.append(buf)
                            done_field = True
                            break
                        elif _length_nl(line) == len(line):
                            fields.append(buf)
                            break
                        elif self.lazy:
                            buf += b'"'
                        else:
                            err = "quote"
                            break
                    elif len(line) > 0:
                        buf += line
                        if err_read_eof:
                            break
                        line, eof = self.read_line()
                        if eof:
                            line, err_read_eof = b"", False  # Go maps io.EOF to nil here; the next round sees len(line) == 0
                    else:
                        if not self.lazy:
                            err = "quote"
                            break
                        fields.append(buf)
                        break
                if done_field:
                    continue
                break
        if err is None:
            if self.fpr > 0:
                if len(fields) != self.fpr:
                    err = "field_count"
            elif self.fpr == 0:
                self.fpr = len(fields)
        return fields, err

    def read_all(self):
        recs = []
        while True:
            rec, err = self.read_record()
            if err is not None:
                return recs, err
            if rec is None:
                return recs, None
            recs.append(rec)
