// csvplus.hpp — header-only C++ mirror of the csvplus Go API (csvplus.go) over the C ABI of
// include/csvplus_b200.h.  The reference's language (Go) has no toolchain in this image; the reference
// is compiled code, so the host side above the C ABI is written in C++ with the reference's names,
// argument meaning and error text.  A DataSource is a plan; sinks lower it to C-ABI calls (CUDA kernels).
// Opaque std::function closures (Filter(func), Map) run on the host at a materialisation boundary
// and the rows are uploaded again with TakeRows semantics — exactly what the cgo wrapper does for Go
// closures.  There is no CPU fallback: without a GPU every sink throws.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../include/csvplus_b200.h"

namespace csvplus {

using Row = std::map<std::string, std::string>;  // map[string]string, csvplus.go:59
using RowFunc = std::function<void(const Row&)>;

struct DataSourceError : std::runtime_error {  // csvplus.go:1230-1238
    uint64_t Line; std::string Err;
    DataSourceError(uint64_t line, const std::string& err)
        : std::runtime_error("row " + std::to_string(line) + ": " + err), Line(line), Err(err) {}
};
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

namespace detail {
inline cpb_ctx* ctx() {
    static cpb_ctx* c = [] {
        cpb_ctx* h = nullptr;
        if (cpb_init(0, &h) != CPB_OK) throw Error("csvplus: no usable CUDA device (there is no CPU fallback)");
        return h;
    }();
    return c;
}
inline void check(int st, const cpb_error& e) {
    if (st == CPB_OK) return;
    if (st == CPB_ERR_DATA && e.has_line) throw DataSourceError(e.line, e.msg);
    throw Error(e.msg[0] ? e.msg : cpb_last_error(ctx()));
}
struct Strs {
    std::vector<std::string> keep; std::vector<cpb_str> v;
    explicit Strs(const std::vector<std::string>& s) : keep(s) { for (auto& x : keep) v.push_back(cpb_str{x.data(), x.size()}); if (v.empty()) v.push_back(cpb_str{nullptr, 0}); }
};
}  // namespace detail

// maps nbytes of device memory into the context's pool ahead of time (cpb_pool_reserve); optional
inline void Reserve(uint64_t nbytes) { if (cpb_pool_reserve(detail::ctx(), nbytes) != CPB_OK) throw Error(cpb_last_error(detail::ctx())); }

// ---------------------------------------------------------------- predicates (csvplus.go:1243-1293)
struct Predicate {
    int op = -1;  // CPB_PRED_*; -1 = opaque host function
    Row match; std::vector<Predicate> kids; std::function<bool(const Row&)> fn;
    bool operator()(const Row& r) const {
        switch (op) {
            case CPB_PRED_LIKE: for (auto& kv : match) { auto it = r.find(kv.first); if (it == r.end() || it->second != kv.second) return false; } return true;
            case CPB_PRED_ALL: for (auto& k : kids) if (!k(r)) return false; return true;
            case CPB_PRED_ANY: for (auto& k : kids) if (k(r)) return true; return false;
            case CPB_PRED_NOT: return !kids[0](r);
            default: return fn(r);
        }
    }
    bool lowerable() const { if (op < 0) return false; for (auto& k : kids) if (!k.lowerable()) return false; return true; }
};
inline Predicate Like(const Row& match) {
    if (match.empty()) throw std::logic_error("empty match row in Like() predicate");
    Predicate p; p.op = CPB_PRED_LIKE; p.match = match; return p;
}
inline Predicate All(std::vector<Predicate> k) { Predicate p; p.op = CPB_PRED_ALL; p.kids = std::move(k); return p; }
inline Predicate Any(std::vector<Predicate> k) { Predicate p; p.op = CPB_PRED_ANY; p.kids = std::move(k); return p; }
inline Predicate Not(Predicate k) { Predicate p; p.op = CPB_PRED_NOT; p.kids = {std::move(k)}; return p; }
inline Predicate Func(std::function<bool(const Row&)> f) { Predicate p; p.fn = std::move(f); return p; }

namespace detail {
struct CPred {  // owns the cpb_pred tree
    std::vector<std::unique_ptr<CPred>> kids; std::vector<const cpb_pred*> kid_ptrs;
    std::vector<std::string> ks, vs; std::vector<cpb_str> kstr, vstr; cpb_pred p{};
    explicit CPred(const Predicate& q) {
        p.op = q.op;
        if (q.op == CPB_PRED_LIKE) {
            for (auto& kv : q.match) { ks.push_back(kv.first); vs.push_back(kv.second); }
            for (size_t i = 0; i < ks.size(); i++) { kstr.push_back(cpb_str{ks[i].data(), ks[i].size()}); vstr.push_back(cpb_str{vs[i].data(), vs[i].size()}); }
            p.n = (int)ks.size(); p.keys = kstr.data(); p.values = vstr.data();
        } else {
            for (auto& k : q.kids) { kids.emplace_back(new CPred(k)); kid_ptrs.push_back(&kids.back()->p); }
            p.n = (int)kids.size(); p.children = kid_ptrs.data();
        }
    }
};
}  // namespace detail

// ---------------------------------------------------------------- Table / Index handles
class Index;
class Table {
public:
    explicit Table(cpb_table* h = nullptr) : h_(h, cpb_table_free) {}
    cpb_table* get() const { return h_.get(); }
    int64_t size() const { return h_ ? cpb_table_num_rows(h_.get()) : 0; }
    std::vector<std::string> columns() const {
        std::vector<std::string> out;
        for (int i = 0; h_ && i < cpb_table_num_cols(h_.get()); i++) { cpb_str s; cpb_table_col_name(h_.get(), i, &s); out.emplace_back(s.ptr, s.len); }
        return out;
    }
    std::vector<Row> rows() const {
        std::vector<Row> out((size_t)size());
        auto cols = columns();
        for (size_t c = 0; c < cols.size(); c++) {
            uint64_t nb = 0;
            cpb_table_col_bytes(detail::ctx(), h_.get(), (int)c, 0, size(), &nb);
            std::vector<int64_t> off((size_t)size() + 1); std::vector<uint8_t> data(nb + 1);
            if (cpb_table_fetch_column(detail::ctx(), h_.get(), (int)c, 0, size(), off.data(), data.data(), data.size()) != CPB_OK) throw Error("fetch failed");
            for (size_t i = 0; i < out.size(); i++) out[i][cols[c]] = std::string((const char*)data.data() + off[i], (size_t)(off[i + 1] - off[i]));
        }
        return out;
    }
    static Table from_rows(const std::vector<Row>& rows) {  // TakeRows, csvplus.go:218
        std::vector<std::string> cols;
        if (!rows.empty()) for (auto& kv : rows[0]) cols.push_back(kv.first);
        std::vector<std::vector<int64_t>> offs(cols.size()); std::vector<std::string> datas(cols.size());
        for (size_t c = 0; c < cols.size(); c++) {
            offs[c].push_back(0);
            for (auto& r : rows) { auto it = r.find(cols[c]); if (it == r.end()) throw Error("TakeRows on the device requires rows with identical column sets"); datas[c] += it->second; offs[c].push_back((int64_t)datas[c].size()); }
            datas[c] += '\0';
        }
        detail::Strs names(cols);
        std::vector<const int64_t*> op; std::vector<const uint8_t*> dp;
        for (size_t c = 0; c < cols.size(); c++) { op.push_back(offs[c].data()); dp.push_back((const uint8_t*)datas[c].data()); }
        cpb_table* h = nullptr;
        if (cpb_table_from_host(detail::ctx(), (int)cols.size(), names.v.data(), op.data(), dp.data(), (int64_t)rows.size(), &h) != CPB_OK) throw Error("from_host failed");
        return Table(h);
    }
private:
    std::shared_ptr<cpb_table> h_;
};

class DataSource;
class Index {  // csvplus.go:610-653
public:
    Index(cpb_index* h, std::vector<std::string> cols) : h_(h, cpb_index_free), columns_(std::move(cols)) {}
    cpb_index* get() const { return h_.get(); }
    const std::vector<std::string>& columns() const { return columns_; }
    int64_t size() const { return cpb_index_num_rows(h_.get()); }
    Table table() const { cpb_table* t = nullptr; cpb_index_table(detail::ctx(), h_.get(), &t); return Table(t); }
    inline DataSource Find(const std::vector<std::string>& values) const;
    Index SubIndex(const std::vector<std::string>& values) const {
        if (values.size() >= columns_.size()) throw std::logic_error("too many values in SubIndex()");
        detail::Strs v(values); cpb_index* s = nullptr;
        if (cpb_index_sub(detail::ctx(), h_.get(), v.v.data(), (int)values.size(), &s) != CPB_OK) throw Error(cpb_last_error(detail::ctx()));
        return Index(s, std::vector<std::string>(columns_.begin() + values.size(), columns_.end()));
    }
    // ResolveDuplicates, csvplus.go:651: resolve returns the position (within the group) of the row to keep, or -1
    void ResolveDuplicates(const std::function<long(const std::vector<Row>&)>& resolve, bool bug_compatible = true) {
        int64_t ng = 0, *lo = nullptr, *hi = nullptr;
        if (cpb_index_dup_groups(detail::ctx(), h_.get(), &ng, &lo, &hi) != CPB_OK) throw Error(cpb_last_error(detail::ctx()));
        std::vector<int64_t> keep((size_t)ng + 1);
        Table t = table();
        auto all = t.rows();
        for (int64_t g = 0; g < ng; g++) {
            std::vector<Row> grp(all.begin() + lo[g], all.begin() + hi[g]);
            long pick = resolve(grp);
            keep[(size_t)g] = pick < 0 ? -1 : lo[g] + pick;
        }
        cpb_free(lo); cpb_free(hi);
        if (cpb_index_dedup_apply(detail::ctx(), h_.get(), ng, keep.data(), bug_compatible) != CPB_OK) throw Error(cpb_last_error(detail::ctx()));
    }
private:
    std::shared_ptr<cpb_index> h_; std::vector<std::string> columns_;
};

// ---------------------------------------------------------------- Reader (csvplus.go:922-1076)
class Reader {
public:
    explicit Reader(std::function<std::string()> source) : source_(std::move(source)) {}
    Reader& Delimiter(char32_t c) { delimiter_ = c; return *this; }
    Reader& CommentChar(char32_t c) { comment_ = c; return *this; }
    Reader& LazyQuotes() { lazy_ = true; return *this; }
    Reader& TrimLeadingSpace() { trim_ = true; return *this; }
    Reader& NumFields(int n) { num_fields_ = n; return *this; }
    Reader& NumFieldsAuto() { return NumFields(0); }
    Reader& NumFieldsAny() { return NumFields(-1); }
    Reader& SelectColumns(const std::vector<std::string>& names) {
        if (names.empty()) throw std::logic_error("empty header spec");
        header_.clear();
        for (auto& n : names) { for (auto& h : header_) if (h.first == n) throw std::logic_error("header spec: duplicate column name: " + n); header_.emplace_back(n, -1); }
        from_first_row_ = true; return *this;
    }
    Reader& ExpectHeader(const std::map<std::string, int>& spec) {
        if (spec.empty()) throw std::logic_error("empty header spec");
        header_.assign(spec.begin(), spec.end()); from_first_row_ = true; return *this;
    }
    Reader& AssumeHeader(const std::map<std::string, int>& spec) {
        if (spec.empty()) throw std::logic_error("Empty header spec");
        for (auto& kv : spec) if (kv.second < 0) throw std::logic_error("header spec: negative index for column " + kv.first);
        header_.assign(spec.begin(), spec.end()); from_first_row_ = false; return *this;
    }
    // Reader.Iterate lowered to cpb_parse_csv; *err receives the DataSourceError raised after the delivered rows
    Table parse(const Predicate* pred, std::unique_ptr<DataSourceError>* err) const {
        std::string data = source_();
        cpb_reader_opts o{}; o.delimiter = delimiter_; o.comment = comment_; o.num_fields = num_fields_;
        o.lazy_quotes = lazy_; o.trim_leading_space = trim_; o.header_from_first_row = from_first_row_;
        std::vector<cpb_header_col> spec;
        for (auto& h : header_) spec.push_back(cpb_header_col{cpb_str{h.first.data(), h.first.size()}, h.second, 0});
        std::unique_ptr<detail::CPred> cp; if (pred) cp.reset(new detail::CPred(*pred));
        cpb_table* t = nullptr; cpb_error e{};
        int st = cpb_parse_csv(detail::ctx(), data.data(), data.size(), 0, &o, spec.data(), (int)spec.size(), cp ? &cp->p : nullptr, &t, &e);
        if (st == CPB_ERR_DATA && t && e.has_line) { err->reset(new DataSourceError(e.line, e.msg)); return Table(t); }
        detail::check(st, e);
        return Table(t);
    }
private:
    std::function<std::string()> source_;
    uint32_t delimiter_ = ',', comment_ = 0; int num_fields_ = 0; bool lazy_ = false, trim_ = false, from_first_row_ = true;
    std::vector<std::pair<std::string, int>> header_;
};
Reader FromFile(const std::string& name);  // csvplus.go:950 (defined below)
inline Reader FromString(std::string data) { return Reader([data] { return data; }); }

// ---------------------------------------------------------------- DataSource (csvplus.go:207-608)
class DataSource {
    struct Op { int kind; Predicate pred; std::vector<std::string> cols; std::shared_ptr<Index> index; uint64_t n = 0; std::function<Row(const Row&)> map; };
    enum { FILTER, SELECT, DROPCOLS, JOIN, EXCEPT, TOP, DROP, MAP };
public:
    explicit DataSource(Reader r) : reader_(new Reader(std::move(r))) {}
    explicit DataSource(Table t) : table_(new Table(std::move(t))) {}
    DataSource Filter(Predicate p) const { return with({FILTER, std::move(p)}); }
    DataSource Map(std::function<Row(const Row&)> f) const { Op o{MAP}; o.map = std::move(f); return with(std::move(o)); }
    DataSource Top(uint64_t n) const { Op o{TOP}; o.n = n; return with(std::move(o)); }
    DataSource Drop(uint64_t n) const { Op o{DROP}; o.n = n; return with(std::move(o)); }
    DataSource SelectColumns(const std::vector<std::string>& c) const { if (c.empty()) throw std::logic_error("no columns specified in SelectColumns()"); Op o{SELECT}; o.cols = c; return with(std::move(o)); }
    DataSource DropColumns(const std::vector<std::string>& c) const { if (c.empty()) throw std::logic_error("no columns specified in DropColumns()"); Op o{DROPCOLS}; o.cols = c; return with(std::move(o)); }
    DataSource Join(const Index& ix, const std::vector<std::string>& c = {}) const {
        if (c.size() > ix.columns().size()) throw std::logic_error("too many source columns in Join()");
        Op o{JOIN}; o.cols = c; o.index.reset(new Index(ix)); return with(std::move(o));
    }
    DataSource Except(const Index& ix, const std::vector<std::string>& c = {}) const {
        if (c.size() > ix.columns().size()) throw std::logic_error("too many source columns in Except()");
        Op o{EXCEPT}; o.cols = c; o.index.reset(new Index(ix)); return with(std::move(o));
    }
    // sinks
    void operator()(const RowFunc& fn) const { std::unique_ptr<DataSourceError> err; Table t = run(&err); for (auto& r : t.rows()) fn(r); if (err) throw *err; }
    std::vector<Row> ToRows() const { std::vector<Row> out; (*this)([&](const Row& r) { out.push_back(r); }); return out; }
    void ToCsv(std::ostream& out, const std::vector<std::string>& columns) const {
        if (columns.empty()) throw std::logic_error("empty column list in ToCsv() function");
        std::unique_ptr<DataSourceError> err; Table t = run(&err);
        detail::Strs c(columns); void* bytes = nullptr; uint64_t n = 0; cpb_error e{};
        detail::check(cpb_table_to_csv(detail::ctx(), t.get(), c.v.data(), (int)columns.size(), &bytes, &n, &e), e);
        out.write((const char*)bytes, (std::streamsize)n); cpb_host_free(detail::ctx(), bytes);
        if (err) throw *err;
    }
    Index IndexOn(const std::vector<std::string>& cols) const { return index(cols, false); }
    Index UniqueIndexOn(const std::vector<std::string>& cols) const { return index(cols, true); }
private:
    DataSource with(Op o) const { DataSource d = *this; d.ops_.push_back(std::move(o)); return d; }
    Index index(const std::vector<std::string>& cols, bool unique) const {
        if (cols.empty()) throw std::logic_error("empty column list in CreateIndex()");
        for (size_t i = 0; i < cols.size(); i++) for (size_t j = i + 1; j < cols.size(); j++) if (cols[i] == cols[j]) throw std::logic_error("duplicate column name(s) in CreateIndex()");
        std::unique_ptr<DataSourceError> err; Table t = run(&err); if (err) throw *err;
        detail::Strs c(cols); cpb_index* ix = nullptr; cpb_error e{};
        detail::check(cpb_index_build(detail::ctx(), t.get(), c.v.data(), (int)cols.size(), unique, &ix, &e), e);
        return Index(ix, cols);
    }
    Table run(std::unique_ptr<DataSourceError>* err) const {
        size_t first = 0; Table t;
        if (reader_) {
            const Predicate* fused = (!ops_.empty() && ops_[0].kind == FILTER && ops_[0].pred.lowerable()) ? &ops_[0].pred : nullptr;
            t = reader_->parse(fused, err); if (fused) first = 1;
        } else t = *table_;
        for (size_t i = first; i < ops_.size(); i++) {
            const Op& o = ops_[i]; cpb_table* out = nullptr; cpb_error e{}; detail::Strs c(o.cols);
            switch (o.kind) {
                case FILTER:
                    if (o.pred.lowerable()) { detail::CPred cp(o.pred); detail::check(cpb_table_filter(detail::ctx(), t.get(), &cp.p, &out), e); }
                    else { std::vector<Row> keep; for (auto& r : t.rows()) if (o.pred(r)) keep.push_back(r); t = Table::from_rows(keep); continue; }
                    break;
                case MAP: { std::vector<Row> m; for (auto& r : t.rows()) m.push_back(o.map(r)); t = Table::from_rows(m); continue; }
                case SELECT: detail::check(cpb_table_select(detail::ctx(), t.get(), c.v.data(), (int)o.cols.size(), &out, &e), e); break;
                case DROPCOLS: detail::check(cpb_table_drop(detail::ctx(), t.get(), c.v.data(), (int)o.cols.size(), &out), e); break;
                case JOIN: detail::check(cpb_join(detail::ctx(), t.get(), o.index->get(), c.v.data(), (int)o.cols.size(), &out, &e), e); break;
                case EXCEPT: detail::check(cpb_except(detail::ctx(), t.get(), o.index->get(), c.v.data(), (int)o.cols.size(), &out, &e), e); break;
                case TOP: if (*err && (uint64_t)t.size() > o.n) err->reset(); detail::check(cpb_table_slice(detail::ctx(), t.get(), 0, (int64_t)o.n, &out), e); break;
                case DROP: detail::check(cpb_table_slice(detail::ctx(), t.get(), (int64_t)o.n, t.size(), &out), e); break;
            }
            t = Table(out);
        }
        return t;
    }
    std::shared_ptr<Reader> reader_; std::shared_ptr<Table> table_; std::vector<Op> ops_;
};

inline DataSource Take(const Reader& r) { return DataSource(r); }                          // csvplus.go:252
inline DataSource Take(const Index& ix) { return DataSource(ix.table()); }
inline DataSource TakeRows(const std::vector<Row>& rows) { return DataSource(Table::from_rows(rows)); }  // :218
inline DataSource Index::Find(const std::vector<std::string>& values) const {
    if (values.size() > columns_.size()) throw std::logic_error("too many columns in indexImpl.find()");
    detail::Strs v(values); cpb_table* t = nullptr;
    if (cpb_index_find(detail::ctx(), h_.get(), v.v.data(), (int)values.size(), &t) != CPB_OK) throw Error(cpb_last_error(detail::ctx()));
    return DataSource(Table(t));
}
inline Reader FromFile(const std::string& name) {
    return Reader([name] {
        FILE* f = fopen(name.c_str(), "rb");
        if (!f) throw DataSourceError(1, std::string("open: ") + strerror(errno));  // mapError, csvplus.go:1216-1220
        std::string s; char buf[1 << 16]; size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
        fclose(f); return s;
    });
}

}  // namespace csvplus
