// rules.hpp -- host-side compiler for the bidding.cl predicate language.
//
// The reference evaluates a rule by `sublis`-ing 13 variables into a quoted form
// and `eval`-ing it (good-opening?, bidding.cl:60-74).  Here the same forms are
// read as S-expressions and compiled to a small warp-uniform bytecode that the
// kernels interpret over an 11-entry feature vector (bytecode.cuh), plus a
// conservative range table (ranges implied by the top-level conjuncts) that the
// sampling kernel uses as its cheap first-stage test.
#pragma once
#include <cstdint>
#include <cstdlib>
#include "bytecode.h"
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bmc {

struct RuleError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- S-expressions -----------------------------------------------------------
struct Sx {
    enum Kind { NIL, SYM, INT, STR, LIST } kind = NIL;
    std::string sym;          // upper-cased symbol / string body
    long num = 0;
    std::vector<Sx> items;    // LIST
    bool dotted = false;      // (a b . c): last item is the tail
    bool is_sym(const char *s) const { return kind == SYM && sym == s; }
    bool is_list() const { return kind == LIST; }
};

class SxReader {
public:
    explicit SxReader(const std::string &text) : s_(text) {}
    Sx read()
    {
        skip();
        if (i_ >= s_.size()) throw RuleError("unexpected end of form");
        char c = s_[i_];
        if (c == '(') {
            ++i_;
            Sx l;
            l.kind = Sx::LIST;
            for (;;) {
                skip();
                if (i_ >= s_.size()) throw RuleError("unbalanced '('");
                if (s_[i_] == ')') { ++i_; break; }
                if (s_[i_] == '.' && i_ + 1 < s_.size() && (isspace((unsigned char)s_[i_ + 1]) || s_[i_ + 1] == '(')) {
                    ++i_;
                    Sx tail = read();
                    skip();
                    if (i_ >= s_.size() || s_[i_] != ')') throw RuleError("bad dotted pair");
                    ++i_;
                    if (tail.kind == Sx::LIST && !tail.dotted) {       // (a . (b c)) == (a b c)
                        for (auto &t : tail.items) l.items.push_back(t);
                    } else if (tail.kind != Sx::NIL) {
                        l.items.push_back(tail);
                        l.dotted = true;
                    }
                    break;
                }
                l.items.push_back(read());
            }
            if (l.items.empty()) l.kind = Sx::NIL;
            return l;
        }
        if (c == ')') throw RuleError("unexpected ')'");
        if (c == '\'') { ++i_; return wrap("QUOTE"); }
        if (c == '`') { ++i_; return wrap("BACKQUOTE"); }
        if (c == '#' && i_ + 1 < s_.size() && s_[i_ + 1] == '\'') { i_ += 2; return wrap("FUNCTION"); }
        if (c == '"') {
            ++i_;
            Sx v;
            v.kind = Sx::STR;
            while (i_ < s_.size() && s_[i_] != '"') {
                if (s_[i_] == '\\' && i_ + 1 < s_.size()) ++i_;
                v.sym.push_back(s_[i_++]);
            }
            if (i_ >= s_.size()) throw RuleError("unterminated string");
            ++i_;
            return v;
        }
        size_t j = i_;
        while (j < s_.size() && !isspace((unsigned char)s_[j]) && s_[j] != '(' && s_[j] != ')' && s_[j] != '\'' && s_[j] != '"')
            ++j;
        std::string tok = s_.substr(i_, j - i_);
        i_ = j;
        if (tok.empty()) throw RuleError("cannot read token");
        Sx v;
        char *end = nullptr;
        long n = strtol(tok.c_str(), &end, 10);
        if (end && *end == 0) { v.kind = Sx::INT; v.num = n; return v; }
        for (auto &ch : tok) ch = (char)toupper((unsigned char)ch);
        if (tok == "NIL") return v;
        v.kind = Sx::SYM;
        v.sym = tok;
        return v;
    }
    bool eof() { skip(); return i_ >= s_.size(); }
private:
    Sx wrap(const char *head)
    {
        Sx l, h;
        l.kind = Sx::LIST;
        h.kind = Sx::SYM;
        h.sym = head;
        l.items.push_back(h);
        l.items.push_back(read());
        return l;
    }
    void skip()
    {
        while (i_ < s_.size()) {
            if (isspace((unsigned char)s_[i_])) ++i_;
            else if (s_[i_] == ';') { while (i_ < s_.size() && s_[i_] != '\n') ++i_; }
            else break;
        }
    }
    std::string s_;
    size_t i_ = 0;
};

// ---- range table -------------------------------------------------------------
struct SeatRanges {
    int lo[F_COUNT], hi[F_COUNT];     // inclusive, over features 0..9; F_BAL unused
    int bal_req = 0;                  // 0 any, 1 balanced, 2 not balanced
    SeatRanges()
    {
        for (int i = 0; i < F_COUNT; ++i) { lo[i] = 0; hi[i] = 63; }
    }
    bool empty() const
    {
        for (int i = 0; i < F_COUNT; ++i) if (lo[i] > hi[i]) return true;
        return false;
    }
};

// ---- compiler ------------------------------------------------------------------
class RuleCompiler {
public:
    // Compile one boolean FORM; appends to `code` (terminated by OP_END).
    // If `ranges` is given, top-level conjunct comparisons against constants are
    // folded into it; *exact is set when the ranges capture the whole form.
    void compile(const Sx &form, std::vector<uint32_t> &code, SeatRanges *ranges = nullptr, bool *exact = nullptr)
    {
        size_t start = code.size();
        expr(form, code);
        code.push_back(enc(OP_END));
        // A program that can raise the error evaluates with Lisp's short-circuit order made explicit: the
        // three-valued and / or / not (bytecode.h), so that an illegal form only counts where it is reached.
        bool has_err = false;
        for (size_t i = start; i < code.size(); ++i) has_err = has_err || (code[i] & 15u) == OP_ERR;
        // boolean stack lives in one 32-bit register, two bits per slot
        int depth = 0, maxd = 0;
        for (size_t i = start; i < code.size(); ++i) {
            uint32_t op = code[i] & 15u;
            if (has_err && (op == OP_AND || op == OP_OR || op == OP_NOT))
                code[i] = (code[i] & ~15u) | (op == OP_AND ? OP_AND3 : op == OP_OR ? OP_OR3 : OP_NOT3);
            if (op == OP_CMPC || op == OP_CMPF || op == OP_TRUE || op == OP_FALSE || op == OP_ERR) ++depth;
            else if (op == OP_AND || op == OP_OR) --depth;
            if (depth > maxd) maxd = depth;
        }
        if (maxd > 15) throw RuleError("rule nests too deeply (boolean stack > 15)");
        if (ranges) {
            bool ex = true;
            absorb(form, *ranges, ex);
            if (exact) *exact = ex;
        }
    }

private:
    static int feature_of(const std::string &s)
    {
        static const char *names[] = {"C", "D", "H", "S", "CPOWER", "DPOWER", "HPOWER", "SPOWER", "HCP", "LOOSERS", "BALANCED"};
        for (int i = 0; i < F_COUNT; ++i) if (s == names[i]) return i;
        if (s == "LOSERS") return F_LOS;
        return -1;
    }
    struct Operand { bool is_const; long k; int f; };

    static long const_value(const Sx &x, bool &ok)
    {
        if (x.kind == Sx::INT) { ok = true; return x.num; }
        if (x.is_list() && !x.items.empty() && x.items[0].kind == Sx::SYM &&
            (x.items[0].sym == "+" || x.items[0].sym == "-" || x.items[0].sym == "*")) {
            char op = x.items[0].sym[0];
            long acc = op == '*' ? 1 : 0;
            for (size_t i = 1; i < x.items.size(); ++i) {
                bool o;
                long v = const_value(x.items[i], o);
                if (!o) { ok = false; return 0; }
                if (op == '+') acc += v;
                else if (op == '*') acc *= v;
                else acc = (i == 1 && x.items.size() > 2) ? v : acc - v;
            }
            ok = true;
            return acc;
        }
        ok = false;
        return 0;
    }
    Operand operand(const Sx &x)
    {
        bool ok;
        long k = const_value(x, ok);
        if (ok) return {true, k, 0};
        if (x.kind == Sx::SYM) {
            int f = feature_of(x.sym);
            if (f >= 0 && f != F_BAL) return {false, 0, f};
        }
        throw RuleError("operand is neither an integer nor a hand variable: " + show(x));
    }
    static std::string show(const Sx &x)
    {
        switch (x.kind) {
        case Sx::NIL: return "NIL";
        case Sx::SYM: return x.sym;
        case Sx::INT: return std::to_string(x.num);
        case Sx::STR: return "\"" + x.sym + "\"";
        default: {
            std::string s = "(";
            for (size_t i = 0; i < x.items.size(); ++i) s += (i ? " " : "") + show(x.items[i]);
            return s + ")";
        }
        }
    }
    static Cmp flip(Cmp c)
    {
        switch (c) { case CMP_LT: return CMP_GT; case CMP_LE: return CMP_GE; case CMP_GT: return CMP_LT; case CMP_GE: return CMP_LE; default: return c; }
    }
    static bool cmp_const(Cmp c, long a, long b)
    {
        switch (c) { case CMP_LT: return a < b; case CMP_LE: return a <= b; case CMP_GT: return a > b; case CMP_GE: return a >= b; case CMP_EQ: return a == b; default: return a != b; }
    }
    static bool parse_cmp(const std::string &s, Cmp &c)
    {
        if (s == "<") c = CMP_LT; else if (s == "<=") c = CMP_LE; else if (s == ">") c = CMP_GT;
        else if (s == ">=") c = CMP_GE; else if (s == "=") c = CMP_EQ; else if (s == "/=") c = CMP_NE; else return false;
        return true;
    }
    // emit (a CMP b)
    void emit_cmp(Cmp c, const Operand &a, const Operand &b, std::vector<uint32_t> &code)
    {
        if (a.is_const && b.is_const) { code.push_back(enc(cmp_const(c, a.k, b.k) ? OP_TRUE : OP_FALSE)); return; }
        if (!a.is_const && !b.is_const) { code.push_back(enc(OP_CMPF, c, a.f, b.f)); return; }
        if (a.is_const) { emit_cmp(flip(c), b, a, code); return; }
        // feature vs const; features live in 0..63, fold out-of-range constants
        long k = b.k;
        if (k < 0 || k > 255) {
            bool v = cmp_const(c, k < 0 ? 0 : 63, k);     // any feature value gives the same answer
            code.push_back(enc(v ? OP_TRUE : OP_FALSE));
            return;
        }
        code.push_back(enc(OP_CMPC, c, a.f, 0, (int)k));
    }
    // (find-if (curry #'OP n) SEQ) -> OR over elements of (n OP elem)
    bool try_find_if(const Sx &x, std::vector<uint32_t> &code)
    {
        if (x.items.size() != 3 || !x.items[0].is_sym("FIND-IF")) return false;
        const Sx &fn = x.items[1];
        if (!fn.is_list() || fn.items.size() != 3 || !fn.items[0].is_sym("CURRY")) throw RuleError("unsupported find-if predicate: " + show(x));
        const Sx &f = fn.items[1];
        Cmp c;
        if (!f.is_list() || f.items.size() != 2 || !f.items[0].is_sym("FUNCTION") || f.items[1].kind != Sx::SYM || !parse_cmp(f.items[1].sym, c))
            throw RuleError("unsupported find-if predicate: " + show(x));
        Operand n = operand(fn.items[2]);
        // sequence
        const Sx &seq = x.items[2];
        int base, from = 0, to = 4;
        auto base_of = [&](const Sx &s) -> int {
            if (s.is_sym("LENS")) return F_C;
            if (s.is_sym("POWER")) return F_CP;
            return -1;
        };
        if ((base = base_of(seq)) >= 0) {
        } else if (seq.is_list() && (seq.items.size() == 3 || seq.items.size() == 4) && seq.items[0].is_sym("SUBSEQ") &&
                   (base = base_of(seq.items[1])) >= 0) {
            bool ok;
            from = (int)const_value(seq.items[2], ok);
            if (!ok) throw RuleError("subseq bounds must be integers: " + show(x));
            if (seq.items.size() == 4) { to = (int)const_value(seq.items[3], ok); if (!ok) throw RuleError("subseq bounds must be integers"); }
            if (from < 0 || to > 4 || from > to) throw RuleError("subseq bounds out of range: " + show(x));
        } else
            throw RuleError("find-if over an unsupported sequence: " + show(x));
        if (from == to) { code.push_back(enc(OP_FALSE)); return true; }
        for (int i = from; i < to; ++i) {
            emit_cmp(c, n, Operand{false, 0, base + i}, code);
            if (i > from) code.push_back(enc(OP_OR));
        }
        return true;
    }
    void expr(const Sx &x, std::vector<uint32_t> &code)
    {
        switch (x.kind) {
        case Sx::NIL: code.push_back(enc(OP_FALSE)); return;
        case Sx::INT: case Sx::STR: code.push_back(enc(OP_TRUE)); return;     // any non-nil object is true
        case Sx::SYM:
            if (x.sym == "T") { code.push_back(enc(OP_TRUE)); return; }
            if (x.sym == "BALANCED") { code.push_back(enc(OP_CMPC, CMP_NE, F_BAL, 0, 0)); return; }
            if (feature_of(x.sym) >= 0) { code.push_back(enc(OP_TRUE)); return; }   // a number: non-nil
            throw RuleError("unknown variable: " + x.sym);
        case Sx::LIST: break;
        }
        const Sx &h = x.items[0];
        // A list whose head is not a symbol (the un-spliced rule list of bidding.cl:249) is an "illegal
        // function call" -- but only if evaluation reaches it: it compiles to the ERROR value, which joins
        // the chain like any other operand; the three-valued and / or / not (bytecode.h) let it surface
        // exactly where Lisp's short-circuit order would evaluate it.
        if (h.kind != Sx::SYM) { code.push_back(enc(OP_ERR)); return; }
        const std::string &op = h.sym;
        size_t n = x.items.size() - 1;
        if (op == "AND" || op == "OR") {
            bool is_and = op == "AND";
            if (n == 0) { code.push_back(enc(is_and ? OP_TRUE : OP_FALSE)); return; }
            for (size_t i = 1; i <= n; ++i) {
                expr(x.items[i], code);
                if (i > 1) code.push_back(enc(is_and ? OP_AND : OP_OR));
            }
            return;
        }
        if (op == "NOT" || op == "NULL") {
            if (n != 1) throw RuleError("not takes one argument");
            expr(x.items[1], code);
            code.push_back(enc(OP_NOT));
            return;
        }
        Cmp c;
        if (parse_cmp(op, c)) {
            if (n < 1) throw RuleError("comparison needs arguments");
            if (n == 1) { operand(x.items[1]); code.push_back(enc(OP_TRUE)); return; }
            if (c == CMP_NE && n > 2) throw RuleError("/= with more than two arguments is not supported");
            for (size_t i = 1; i < n; ++i) {
                emit_cmp(c, operand(x.items[i]), operand(x.items[i + 1]), code);
                if (i > 1) code.push_back(enc(OP_AND));
            }
            return;
        }
        if (op == "COND") {
            // (cond (t1 e1..) (t2 e2..)) = (t1 & e1) | (!t1 & t2 & e2) | ...
            if (n == 0) { code.push_back(enc(OP_FALSE)); return; }
            // build right-to-left as nested if: if t1 then e1 else rest
            cond_from(x, 1, code);
            return;
        }
        if (op == "IF") {
            if (n < 2 || n > 3) throw RuleError("if takes 2 or 3 arguments");
            // (t & a) | (!t & b)
            expr(x.items[1], code); expr(x.items[2], code); code.push_back(enc(OP_AND));
            expr(x.items[1], code); code.push_back(enc(OP_NOT));
            if (n == 3) expr(x.items[3], code); else code.push_back(enc(OP_FALSE));
            code.push_back(enc(OP_AND));
            code.push_back(enc(OP_OR));
            return;
        }
        if (op == "FIND-IF") { try_find_if(x, code); return; }
        if (op == "QUOTE") { expr_quoted(x.items.size() > 1 ? x.items[1] : Sx{}, code); return; }
        throw RuleError("form outside the bidding.cl predicate language: " + show(x));
    }
    void expr_quoted(const Sx &x, std::vector<uint32_t> &code) { code.push_back(enc(x.kind == Sx::NIL ? OP_FALSE : OP_TRUE)); }
    void cond_from(const Sx &x, size_t i, std::vector<uint32_t> &code)
    {
        if (i >= x.items.size()) { code.push_back(enc(OP_FALSE)); return; }
        const Sx &cl = x.items[i];
        if (!cl.is_list() || cl.items.empty()) throw RuleError("bad cond clause");
        // value of clause = last body form, or the test itself
        expr(cl.items[0], code);
        if (cl.items.size() > 1) { expr(cl.items.back(), code); code.push_back(enc(OP_AND)); }
        if (i + 1 < x.items.size()) {
            expr(cl.items[0], code);
            code.push_back(enc(OP_NOT));
            cond_from(x, i + 1, code);
            code.push_back(enc(OP_AND));
            code.push_back(enc(OP_OR));
        }
    }

    // ---- range absorption over top-level conjuncts -----------------------------
    void tighten(SeatRanges &r, int f, Cmp c, long k)
    {
        switch (c) {
        case CMP_LT: if (k - 1 < r.hi[f]) r.hi[f] = (int)(k - 1); break;
        case CMP_LE: if (k < r.hi[f]) r.hi[f] = (int)k; break;
        case CMP_GT: if (k + 1 > r.lo[f]) r.lo[f] = (int)(k + 1); break;
        case CMP_GE: if (k > r.lo[f]) r.lo[f] = (int)k; break;
        case CMP_EQ: if (k > r.lo[f]) r.lo[f] = (int)k; if (k < r.hi[f]) r.hi[f] = (int)k; break;
        default: break;
        }
    }
    void absorb(const Sx &x, SeatRanges &r, bool &exact)
    {
        if (x.kind == Sx::SYM && x.sym == "T") return;
        if (x.kind == Sx::SYM && x.sym == "BALANCED") {
            if (r.bal_req == 2) { r.lo[F_HCP] = 1; r.hi[F_HCP] = 0; }   // contradiction
            r.bal_req = 1;
            // necessary shape conditions of balanced? (bidding.cl:42-44)
            for (int s = 0; s < 4; ++s) {
                if (r.lo[s] < 2) r.lo[s] = 2;
                int mx = s >= 2 ? 4 : 5;
                if (r.hi[s] > mx) r.hi[s] = mx;
            }
            return;
        }
        if (!x.is_list() || x.items.empty() || x.items[0].kind != Sx::SYM) { exact = false; return; }
        const std::string &op = x.items[0].sym;
        if (op == "AND") {
            for (size_t i = 1; i < x.items.size(); ++i) absorb(x.items[i], r, exact);
            return;
        }
        if (op == "NOT" && x.items.size() == 2 && x.items[1].is_sym("BALANCED")) {
            if (r.bal_req == 1) { r.lo[F_HCP] = 1; r.hi[F_HCP] = 0; }
            r.bal_req = 2;
            return;
        }
        Cmp c;
        if (parse_cmp(op, c) && x.items.size() == 3 && c != CMP_NE) {
            try {
                Operand a = operand(x.items[1]), b = operand(x.items[2]);
                if (!a.is_const && b.is_const) { tighten(r, a.f, c, clampk(b.k)); return; }
                if (a.is_const && !b.is_const) { tighten(r, b.f, flip(c), clampk(a.k)); return; }
            } catch (const RuleError &) {
            }
        }
        exact = false;
    }
    static long clampk(long k) { return k < -1 ? -1 : k > 64 ? 64 : k; }
};

// A rule table "(((2 C) . FORM) ...)": rows[i] = (bid text, program)
struct SchemeRows {
    std::vector<std::string> bids;
    std::vector<std::vector<uint32_t>> progs;
    std::vector<Sx> forms;             // the rows' rule forms (NIL for a NIL row): the NVRTC path re-reads them
};

inline std::string sx_to_string(const Sx &x)
{
    switch (x.kind) {
    case Sx::NIL: return "NIL";
    case Sx::SYM: return x.sym;
    case Sx::INT: return std::to_string(x.num);
    case Sx::STR: return x.sym;
    default: {
        std::string s = "(";
        for (size_t i = 0; i < x.items.size(); ++i) s += (i ? " " : "") + sx_to_string(x.items[i]);
        return s + ")";
    }
    }
}

inline SchemeRows compile_scheme(const std::string &text)
{
    SxReader rd(text);
    Sx table = rd.read();
    if (table.is_list() && table.items.size() == 2 && table.items[0].is_sym("QUOTE")) table = table.items[1];
    SchemeRows out;
    if (table.kind == Sx::NIL) return out;
    if (!table.is_list()) throw RuleError("rule table must be a list of (bid . form) rows");
    RuleCompiler rc;
    for (const Sx &row : table.items) {
        if (row.kind == Sx::NIL) {               // a NIL row (further-bid emits them): never matches
            out.bids.push_back("NIL");
            out.progs.push_back({enc(OP_FALSE), enc(OP_END)});
            out.forms.push_back(Sx{});
            continue;
        }
        if (!row.is_list() || row.items.empty()) throw RuleError("bad rule row: " + sx_to_string(row));
        // (bid . form): after reading, a proper-list form is spliced: (bid op args...)
        Sx form;
        if (row.dotted) {
            form = row.items.back();          // (bid . atom)
            if (row.items.size() != 2) throw RuleError("bad rule row: " + sx_to_string(row));
        } else if (row.items.size() > 1) {
            form.kind = Sx::LIST;
            form.items.assign(row.items.begin() + 1, row.items.end());
        }                                       // ((bid)) == (bid . nil): never matches
        std::vector<uint32_t> code;
        rc.compile(form, code);
        out.bids.push_back(sx_to_string(row.items[0]));
        out.progs.push_back(std::move(code));
        out.forms.push_back(form);
    }
    return out;
}

}  // namespace bmc
