#include "sqlexpr.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace tg {

namespace {

struct Tok {
    enum K { END, NUM_I, NUM_F, STR, IDENT, QIDENT, OP, LP, RP, COMMA } k;
    std::string s;
    int64_t i = 0;
    double f = 0;
};

static std::string upper(std::string s) {
    for (auto& c : s) c = (char)toupper((unsigned char)c);
    return s;
}

struct Lexer {
    const std::string& t;
    size_t p = 0;
    explicit Lexer(const std::string& text) : t(text) {}
    [[noreturn]] void err(const std::string& m) { throw Error(TG_ERR_UNSUPPORTED, "SQL parser error: " + m); }
    Tok next() {
        while (p < t.size() && isspace((unsigned char)t[p])) ++p;
        Tok k;
        if (p >= t.size()) {
            k.k = Tok::END;
            return k;
        }
        char c = t[p];
        if (isdigit((unsigned char)c) || (c == '.' && p + 1 < t.size() && isdigit((unsigned char)t[p + 1]))) {
            size_t s = p;
            bool is_f = false;
            while (p < t.size() && isdigit((unsigned char)t[p])) ++p;
            if (p < t.size() && t[p] == '.') {
                is_f = true;
                ++p;
                while (p < t.size() && isdigit((unsigned char)t[p])) ++p;
            }
            if (p < t.size() && (t[p] == 'e' || t[p] == 'E')) {
                size_t q = p + 1;
                if (q < t.size() && (t[q] == '+' || t[q] == '-')) ++q;
                if (q < t.size() && isdigit((unsigned char)t[q])) {
                    is_f = true;
                    p = q;
                    while (p < t.size() && isdigit((unsigned char)t[p])) ++p;
                }
            }
            std::string num = t.substr(s, p - s);
            if (is_f) {
                k.k = Tok::NUM_F;
                k.f = strtod(num.c_str(), nullptr);
            } else {
                errno = 0;
                long long v = strtoll(num.c_str(), nullptr, 10);
                if (errno == ERANGE) {
                    k.k = Tok::NUM_F;
                    k.f = strtod(num.c_str(), nullptr);
                } else {
                    k.k = Tok::NUM_I;
                    k.i = v;
                }
            }
            return k;
        }
        if (isalpha((unsigned char)c) || c == '_') {
            size_t s = p;
            while (p < t.size() && (isalnum((unsigned char)t[p]) || t[p] == '_' || t[p] == '.')) ++p;
            k.k = Tok::IDENT;
            k.s = t.substr(s, p - s);
            return k;
        }
        if (c == '"') {
            ++p;
            std::string s;
            while (p < t.size()) {
                if (t[p] == '"') {
                    if (p + 1 < t.size() && t[p + 1] == '"') {
                        s += '"';
                        p += 2;
                        continue;
                    }
                    break;
                }
                s += t[p++];
            }
            if (p >= t.size()) err("unterminated quoted identifier");
            ++p;
            k.k = Tok::QIDENT;
            k.s = s;
            return k;
        }
        if (c == '\'') {
            ++p;
            std::string s;
            while (p < t.size()) {
                if (t[p] == '\'') {
                    if (p + 1 < t.size() && t[p + 1] == '\'') {
                        s += '\'';
                        p += 2;
                        continue;
                    }
                    break;
                }
                s += t[p++];
            }
            if (p >= t.size()) err("unterminated string literal");
            ++p;
            k.k = Tok::STR;
            k.s = s;
            return k;
        }
        if (c == '(') { ++p; k.k = Tok::LP; return k; }
        if (c == ')') { ++p; k.k = Tok::RP; return k; }
        if (c == ',') { ++p; k.k = Tok::COMMA; return k; }
        static const char* ops2[] = {"<=", ">=", "<>", "!=", "=="};
        for (const char* o : ops2) {
            if (t.compare(p, 2, o) == 0) {
                p += 2;
                k.k = Tok::OP;
                k.s = o;
                return k;
            }
        }
        if (strchr("+-*/%<>=", c)) {
            ++p;
            k.k = Tok::OP;
            k.s = std::string(1, c);
            return k;
        }
        err(std::string("unexpected character '") + c + "'");
    }
};

struct Parser {
    Lexer lx;
    Tok cur;
    explicit Parser(const std::string& t) : lx(t) { cur = lx.next(); }
    void adv() { cur = lx.next(); }
    bool is_kw(const char* kw) const { return cur.k == Tok::IDENT && upper(cur.s) == kw; }
    bool is_op(const char* o) const { return cur.k == Tok::OP && cur.s == o; }
    [[noreturn]] void err(const std::string& m) { throw Error(TG_ERR_UNSUPPORTED, "SQL parser error: " + m); }

    static ExprP mk(Expr::Kind k) {
        auto e = std::make_shared<Expr>();
        e->kind = k;
        return e;
    }
    static ExprP bin(const std::string& op, ExprP a, ExprP b) {
        auto e = mk(Expr::BINARY);
        e->s = op;
        e->args = {a, b};
        return e;
    }
    static ExprP un(const std::string& op, ExprP a) {
        auto e = mk(Expr::UNARY);
        e->s = op;
        e->args = {a};
        return e;
    }

    ExprP parse_or() {
        ExprP l = parse_and();
        while (is_kw("OR")) {
            adv();
            l = bin("OR", l, parse_and());
        }
        return l;
    }
    ExprP parse_and() {
        ExprP l = parse_not();
        while (is_kw("AND")) {
            adv();
            l = bin("AND", l, parse_not());
        }
        return l;
    }
    ExprP parse_not() {
        if (is_kw("NOT")) {
            adv();
            return un("NOT", parse_not());
        }
        return parse_cmp();
    }
    ExprP parse_cmp() {
        ExprP l = parse_add();
        while (true) {
            if (cur.k == Tok::OP && (cur.s == "=" || cur.s == "==" || cur.s == "<>" || cur.s == "!=" ||
                                     cur.s == "<" || cur.s == "<=" || cur.s == ">" || cur.s == ">=")) {
                std::string op = cur.s;
                if (op == "==") op = "=";
                if (op == "!=") op = "<>";
                adv();
                l = bin(op, l, parse_add());
                continue;
            }
            if (is_kw("IS")) {
                adv();
                bool neg = false;
                if (is_kw("NOT")) {
                    neg = true;
                    adv();
                }
                auto e = mk(Expr::IS);
                if (is_kw("NULL")) e->s = neg ? "NOTNULL" : "NULL";
                else if (is_kw("TRUE")) e->s = neg ? "NOTTRUE" : "TRUE";
                else if (is_kw("FALSE")) e->s = neg ? "NOTFALSE" : "FALSE";
                else err("expected NULL, TRUE or FALSE after IS");
                adv();
                e->args = {l};
                l = e;
                continue;
            }
            bool neg = false;
            if (is_kw("NOT")) {
                // lookahead: NOT BETWEEN / NOT IN
                size_t save_p = lx.p;
                Tok save = cur;
                adv();
                if (is_kw("BETWEEN") || is_kw("IN") || is_kw("LIKE")) neg = true;
                else {
                    lx.p = save_p;
                    cur = save;
                    break;
                }
            }
            if (is_kw("LIKE")) {  // pattern: a string literal (`%`, `_`, backslash escape); evaluated as a virtual column
                adv();
                ExprP pat = parse_add();
                if (pat->kind != Expr::LIT_S) err("the LIKE pattern must be a string literal");
                auto e = mk(Expr::FUNC);
                e->s = "LIKE";
                e->args = {l, pat};
                l = neg ? un("NOT", e) : ExprP(e);
                continue;
            }
            if (is_kw("BETWEEN")) {
                adv();
                ExprP lo = parse_add();
                if (!is_kw("AND")) err("expected AND in BETWEEN");
                adv();
                ExprP hi = parse_add();
                ExprP e = bin("AND", bin(">=", l, lo), bin("<=", l, hi));
                l = neg ? un("NOT", e) : e;
                continue;
            }
            if (is_kw("IN")) {
                adv();
                if (cur.k != Tok::LP) err("expected ( after IN");
                adv();
                ExprP e;
                while (true) {
                    ExprP item = parse_add();
                    ExprP eq = bin("=", l, item);
                    e = e ? bin("OR", e, eq) : eq;
                    if (cur.k == Tok::COMMA) {
                        adv();
                        continue;
                    }
                    break;
                }
                if (cur.k != Tok::RP) err("expected ) after IN list");
                adv();
                l = neg ? un("NOT", e) : e;
                continue;
            }
            break;
        }
        return l;
    }
    ExprP parse_add() {
        ExprP l = parse_mul();
        while (is_op("+") || is_op("-")) {
            std::string op = cur.s;
            adv();
            l = bin(op, l, parse_mul());
        }
        return l;
    }
    ExprP parse_mul() {
        ExprP l = parse_unary();
        while (is_op("*") || is_op("/") || is_op("%")) {
            std::string op = cur.s;
            adv();
            l = bin(op, l, parse_unary());
        }
        return l;
    }
    ExprP parse_unary() {
        if (is_op("-")) {
            adv();
            ExprP a = parse_unary();
            if (a->kind == Expr::LIT_I) {
                a->i = (int64_t)(0 - (uint64_t)a->i);
                return a;
            }
            if (a->kind == Expr::LIT_F) {
                a->f = -a->f;
                return a;
            }
            return un("NEG", a);
        }
        if (is_op("+")) {
            adv();
            return parse_unary();
        }
        return parse_primary();
    }
    // CASE [operand] WHEN c THEN x [WHEN ..] [ELSE y] END -> FUNC "CASE", args = c1, x1, .., cn, xn [, y]; i = 1 when ELSE is there.
    // The simple form compares the operand with every WHEN value.
    ExprP parse_case() {
        ExprP operand;
        if (!is_kw("WHEN")) operand = parse_or();
        auto e = mk(Expr::FUNC);
        e->s = "CASE";
        if (!is_kw("WHEN")) err("expected WHEN after CASE");
        while (is_kw("WHEN")) {
            adv();
            ExprP c = parse_or();
            if (operand) c = bin("=", operand, c);
            if (!is_kw("THEN")) err("expected THEN");
            adv();
            e->args.push_back(c);
            e->args.push_back(parse_or());
        }
        if (is_kw("ELSE")) {
            adv();
            e->args.push_back(parse_or());
            e->i = 1;
        }
        if (!is_kw("END")) err("expected END of CASE");
        adv();
        return e;
    }
    // CAST(x AS type): the floating-point types and the integer types (of an integer operand)
    ExprP parse_cast() {
        adv();  // (
        auto e = mk(Expr::FUNC);
        e->args.push_back(parse_or());
        if (!is_kw("AS")) err("expected AS in CAST");
        adv();
        if (cur.k != Tok::IDENT) err("expected a type name in CAST");
        const std::string ty = upper(cur.s);
        adv();
        if (ty == "DOUBLE" && is_kw("PRECISION")) adv();
        if (cur.k == Tok::LP) {  // (precision [, scale])
            while (cur.k != Tok::RP && cur.k != Tok::END) adv();
            if (cur.k == Tok::RP) adv();
        }
        if (ty == "DOUBLE" || ty == "FLOAT" || ty == "REAL" || ty == "FLOAT8" || ty == "FLOAT4") e->s = "CAST_F64";
        else if (ty == "BIGINT" || ty == "INT" || ty == "INTEGER" || ty == "SMALLINT" || ty == "TINYINT" || ty == "INT8" || ty == "INT4") e->s = "CAST_I64";
        else err("CAST to " + ty + " is not supported");
        if (cur.k != Tok::RP) err("expected ) after CAST");
        adv();
        return e;
    }
    ExprP parse_primary() {
        if (cur.k == Tok::NUM_I) {
            auto e = mk(Expr::LIT_I);
            e->i = cur.i;
            adv();
            return e;
        }
        if (cur.k == Tok::NUM_F) {
            auto e = mk(Expr::LIT_F);
            e->f = cur.f;
            adv();
            return e;
        }
        if (cur.k == Tok::STR) {
            auto e = mk(Expr::LIT_S);
            e->s = cur.s;
            adv();
            return e;
        }
        if (cur.k == Tok::LP) {
            adv();
            ExprP e = parse_or();
            if (cur.k != Tok::RP) err("expected )");
            adv();
            return e;
        }
        if (cur.k == Tok::QIDENT) {
            auto e = mk(Expr::COL);
            e->s = cur.s;
            adv();
            return e;
        }
        if (cur.k == Tok::IDENT) {
            std::string name = cur.s, up = upper(cur.s);
            if (up == "TRUE" || up == "FALSE") {
                auto e = mk(Expr::LIT_B);
                e->b = up == "TRUE";
                adv();
                return e;
            }
            if (up == "NULL") {
                adv();
                return mk(Expr::LIT_NULL);
            }
            adv();
            if ((up == "DATE" || up == "TIMESTAMP") && cur.k == Tok::STR) {  // typed literal: resolved against the column it meets
                auto e = mk(Expr::FUNC);
                e->s = up + "_LITERAL";
                auto lit = mk(Expr::LIT_S);
                lit->s = cur.s;
                e->args.push_back(lit);
                adv();
                return e;
            }
            if (up == "INTERVAL" && cur.k == Tok::STR) {  // INTERVAL '1 day' / INTERVAL '1' DAY: folded with the constant it meets
                auto e = mk(Expr::FUNC);
                e->s = "INTERVAL_LITERAL";
                auto lit = mk(Expr::LIT_S);
                lit->s = cur.s;
                adv();
                if (cur.k == Tok::IDENT) {
                    std::string u = upper(cur.s);
                    if (!u.empty() && u.back() == 'S') u.pop_back();
                    if (u == "YEAR" || u == "MONTH" || u == "WEEK" || u == "DAY" || u == "HOUR" || u == "MINUTE" || u == "SECOND" ||
                        u == "MILLISECOND" || u == "MICROSECOND" || u == "NANOSECOND") {
                        lit->s += " " + cur.s;
                        adv();
                    }
                }
                e->args.push_back(lit);
                return e;
            }
            if ((up == "CURRENT_TIMESTAMP" || up == "CURRENT_DATE") && cur.k != Tok::LP) {  // (SQL spells these without parentheses)
                auto e = mk(Expr::FUNC);
                e->s = up;
                return e;
            }
            if (up == "CASE") return parse_case();
            if (up == "CAST" && cur.k == Tok::LP) return parse_cast();
            if (cur.k == Tok::LP) {
                adv();
                auto e = mk(Expr::FUNC);
                e->s = up;
                if (cur.k != Tok::RP) {
                    while (true) {
                        e->args.push_back(parse_or());
                        if (cur.k == Tok::COMMA) {
                            adv();
                            continue;
                        }
                        break;
                    }
                }
                if (cur.k != Tok::RP) err("expected ) after function arguments");
                adv();
                return e;
            }
            auto e = mk(Expr::COL);
            // unquoted identifiers are folded to lower case by the SQL parser (DataFusion normalises)
            std::string low = name;
            for (auto& ch : low) ch = (char)tolower((unsigned char)ch);
            // qualified name t.c -> keep the column part
            size_t dot = low.rfind('.');
            e->s = dot == std::string::npos ? low : low.substr(dot + 1);
            return e;
        }
        err("unexpected token");
    }
};

// ---------------- compiler ----------------
struct Operand {
    uint8_t kind;   // PK_*
    uint16_t idx;   // temp or tile column
    uint64_t imm;
    PredType type;
};

struct Compiler {
    const ColumnResolver& resolve;
    std::vector<PredInstr>& code;
    bool ntemp_used[4] = {false, false, false, false};
    bool btemp_used[4] = {false, false, false, false};

    int alloc(bool* used) {
        for (int k = 0; k < 4; ++k)
            if (!used[k]) {
                used[k] = true;
                return k;
            }
        throw Error(TG_ERR_UNSUPPORTED, "SQL expression too deeply nested for the predicate engine");
    }
    void release(const Operand& o) {
        if (o.kind == PK_NTEMP) ntemp_used[o.idx] = false;
        if (o.kind == PK_BTEMP) btemp_used[o.idx] = false;
    }
    static uint64_t d2u(double d) {
        uint64_t u;
        memcpy(&u, &d, 8);
        return u;
    }
    static bool is_imm(const Operand& o) { return o.kind == PK_IMM || o.kind == PK_BIMM; }
    Operand none() { return Operand{PK_NONE, 0, 0, PT_NULL}; }
    Operand null_num() { return Operand{PK_NNULL, 0, 0, PT_NULL}; }
    Operand null_bool() { return Operand{PK_BNULL, 0, 0, PT_BOOL}; }

    // result_bool: destination register file
    Operand emit(uint8_t op, Operand a, Operand b, PredType type) {
        const bool result_bool = type == PT_BOOL;
        if (is_imm(a) && is_imm(b)) {
            // at most one immediate per instruction: materialise `a`
            a = emit(a.kind == PK_BIMM ? PO_MOVB : PO_MOVN, a, none(), a.type);
        }
        release(a);
        release(b);
        int dst = alloc(result_bool ? btemp_used : ntemp_used);
        PredInstr ins{};
        ins.op = op;
        ins.dst = (uint8_t)dst;
        ins.a_kind = a.kind;
        ins.a_idx = a.idx;
        ins.b_kind = b.kind;
        ins.b_idx = b.idx;
        ins.imm = is_imm(a) ? a.imm : b.imm;
        if ((int)code.size() >= SCAN_MAX_CODE) throw Error(TG_ERR_UNSUPPORTED, "SQL expression too long");
        code.push_back(ins);
        return Operand{(uint8_t)(result_bool ? PK_BTEMP : PK_NTEMP), (uint16_t)dst, 0, type};
    }

    Operand to_f64(Operand o) {
        if (o.type == PT_F64 || o.type == PT_NULL) return o;
        if (o.type != PT_I64) throw Error(TG_ERR_TYPE_MISMATCH, "cannot use a boolean value in arithmetic");
        if (o.kind == PK_IMM) return Operand{PK_IMM, 0, d2u((double)(int64_t)o.imm), PT_F64};
        if (o.kind == PK_COL_I64) return Operand{PK_COL_I64_AS_F64, o.idx, 0, PT_F64};
        return emit(PO_I2F, o, none(), PT_F64);
    }
    // a NULL literal adopts the register file its context needs
    Operand as_bool(Operand o) {
        if (o.type == PT_NULL) return null_bool();
        if (o.type != PT_BOOL) throw Error(TG_ERR_TYPE_MISMATCH, "expected a boolean operand");
        return o;
    }

    // static type of an expression (what gen() will produce), needed before the arms of CASE / COALESCE are generated
    static PredType unify(PredType a, PredType b) {
        if (a == PT_NULL) return b;
        if (b == PT_NULL) return a;
        if (a == PT_BOOL || b == PT_BOOL) {
            if (a != b) throw Error(TG_ERR_TYPE_MISMATCH, "CASE / COALESCE arms mix boolean and numeric values");
            return PT_BOOL;
        }
        return a == PT_F64 || b == PT_F64 ? PT_F64 : PT_I64;
    }
    PredType type_of(const ExprP& e) {
        switch (e->kind) {
            case Expr::LIT_I: return PT_I64;
            case Expr::LIT_F: return PT_F64;
            case Expr::LIT_B: return PT_BOOL;
            case Expr::LIT_NULL: return PT_NULL;
            case Expr::LIT_S: throw Error(TG_ERR_UNSUPPORTED, "string literals are only supported in comparisons with a Utf8 column");
            case Expr::COL: {
                const ColumnBinding b = resolve(e->s);
                if (b.dtype == TG_INT64) return PT_I64;
                if (b.dtype == TG_FLOAT64) return PT_F64;
                if (b.dtype == TG_BOOL) return PT_BOOL;
                throw Error(TG_ERR_UNSUPPORTED, "column '" + e->s + "' has a type the predicate engine does not support");
            }
            case Expr::UNARY: return e->s == "NOT" ? PT_BOOL : type_of(e->args[0]);
            case Expr::IS: return PT_BOOL;
            case Expr::BINARY: {
                const std::string& op = e->s;
                if (!(op == "+" || op == "-" || op == "*" || op == "/" || op == "%")) return PT_BOOL;
                const PredType a = type_of(e->args[0]), b = type_of(e->args[1]);
                if (a == PT_NULL || b == PT_NULL) return PT_NULL;
                return a == PT_F64 || b == PT_F64 ? PT_F64 : PT_I64;
            }
            case Expr::FUNC: {
                if (e->s == "ABS" && e->args.size() == 1) return type_of(e->args[0]);
                if (e->s == "CAST_F64") return PT_F64;
                if (e->s == "CAST_I64") return PT_I64;
                if (e->s == "COALESCE" || e->s == "CASE") {
                    PredType t = PT_NULL;
                    const size_t n_arm = e->s == "CASE" ? (e->args.size() - (size_t)e->i) / 2 : 0;
                    for (size_t k = 0; k < e->args.size(); ++k) {
                        if (e->s == "CASE" && k < 2 * n_arm && k % 2 == 0) continue;  // a WHEN condition
                        t = unify(t, type_of(e->args[k]));
                    }
                    return t;
                }
                throw Error(TG_ERR_UNSUPPORTED, "function " + e->s + " is not supported by the predicate engine");
            }
        }
        throw Error(TG_ERR_INTERNAL, "bad expression node");
    }
    static int height(const ExprP& e) {
        int h = 0;
        for (auto& a : e->args) h = std::max(h, height(a));
        return h + 1;
    }
    Operand as_num(Operand o, PredType t) {
        if (o.type == PT_BOOL) throw Error(TG_ERR_TYPE_MISMATCH, "expected a numeric operand");
        return t == PT_F64 ? to_f64(o) : o;
    }
    static ExprP node(Expr::Kind k, const std::string& s, std::vector<ExprP> args) {
        auto e = std::make_shared<Expr>();
        e->kind = k;
        e->s = s;
        e->args = std::move(args);
        return e;
    }

    Operand gen(const ExprP& e) {
        switch (e->kind) {
            case Expr::LIT_I: return Operand{PK_IMM, 0, (uint64_t)e->i, PT_I64};
            case Expr::LIT_F: return Operand{PK_IMM, 0, d2u(e->f), PT_F64};
            case Expr::LIT_B: return Operand{PK_BIMM, 0, e->b ? 1ull : 0ull, PT_BOOL};
            case Expr::LIT_NULL: return null_num();
            case Expr::LIT_S:
                throw Error(TG_ERR_UNSUPPORTED, "string literals are only supported in comparisons with a Utf8 column");
            case Expr::COL: {
                ColumnBinding b = resolve(e->s);
                if (b.dtype == TG_INT64) return Operand{PK_COL_I64, (uint16_t)b.tile_col, 0, PT_I64};
                if (b.dtype == TG_FLOAT64) return Operand{PK_COL_F64, (uint16_t)b.tile_col, 0, PT_F64};
                if (b.dtype == TG_BOOL) return Operand{PK_COL_BOOL, (uint16_t)b.tile_col, 0, PT_BOOL};
                throw Error(TG_ERR_UNSUPPORTED, "column '" + e->s + "' has a type the predicate engine does not support");
            }
            case Expr::UNARY: {
                Operand a = gen(e->args[0]);
                if (e->s == "NOT") return emit(PO_NOT, as_bool(a), none(), PT_BOOL);
                if (a.type == PT_I64) return emit(PO_NEG_I, a, none(), PT_I64);
                if (a.type == PT_F64) return emit(PO_NEG_F, a, none(), PT_F64);
                if (a.type == PT_NULL) return a;
                throw Error(TG_ERR_TYPE_MISMATCH, "unary minus requires a numeric operand");
            }
            case Expr::IS: {
                Operand a = gen(e->args[0]);
                const bool is_b = a.type == PT_BOOL;
                if (e->s == "NULL") return emit(is_b ? PO_ISNULL_B : PO_ISNULL_N, a, none(), PT_BOOL);
                if (e->s == "NOTNULL") return emit(is_b ? PO_ISNOTNULL_B : PO_ISNOTNULL_N, a, none(), PT_BOOL);
                a = as_bool(a);
                if (e->s == "TRUE") return emit(PO_ISTRUE, a, none(), PT_BOOL);
                if (e->s == "FALSE") return emit(PO_ISFALSE, a, none(), PT_BOOL);
                Operand t = emit(e->s == "NOTTRUE" ? PO_ISTRUE : PO_ISFALSE, a, none(), PT_BOOL);
                return emit(PO_NOT, t, none(), PT_BOOL);
            }
            case Expr::FUNC: {
                if (e->s == "ABS" && e->args.size() == 1) {
                    Operand a = gen(e->args[0]);
                    if (a.type == PT_I64) return emit(PO_ABS_I, a, none(), PT_I64);
                    if (a.type == PT_F64) return emit(PO_ABS_F, a, none(), PT_F64);
                    if (a.type == PT_NULL) return a;
                    throw Error(TG_ERR_TYPE_MISMATCH, "ABS requires a numeric operand");
                }
                if (e->s == "CAST_F64" && e->args.size() == 1) return to_f64(gen(e->args[0]));
                if (e->s == "CAST_I64" && e->args.size() == 1) {
                    Operand a = gen(e->args[0]);
                    if (a.type == PT_I64 || a.type == PT_NULL) return a;
                    throw Error(TG_ERR_UNSUPPORTED, "CAST to an integer type is supported for integer operands only");
                }
                if (e->s == "COALESCE" && !e->args.empty()) {
                    const PredType t = type_of(e);
                    if (t == PT_BOOL) throw Error(TG_ERR_UNSUPPORTED, "COALESCE over boolean operands is not supported");
                    if (t == PT_NULL) return null_num();
                    Operand r = as_num(gen(e->args[0]), t);
                    for (size_t k = 1; k < e->args.size(); ++k) {
                        Operand b = as_num(gen(e->args[k]), t);
                        r = emit(PO_COALESCE_N, r, b, t);
                    }
                    return r;
                }
                if (e->s == "CASE" && e->args.size() >= 2) {
                    const size_t n_arm = (e->args.size() - (size_t)e->i) / 2;
                    const PredType t = type_of(e);
                    auto is_true = [&](const ExprP& c) { return node(Expr::IS, "TRUE", {c}); };
                    if (t == PT_BOOL || t == PT_NULL) {
                        // (c IS TRUE AND x) OR (NOT (c IS TRUE) AND rest): three-valued AND / OR give exactly the arm's value
                        ExprP rest = e->i ? e->args.back() : node(Expr::LIT_NULL, "", {});
                        for (size_t k = n_arm; k-- > 0;) {
                            const ExprP c = is_true(e->args[2 * k]);
                            rest = node(Expr::BINARY, "OR", {node(Expr::BINARY, "AND", {c, e->args[2 * k + 1]}),
                                                             node(Expr::BINARY, "AND", {node(Expr::UNARY, "NOT", {c}), rest})});
                        }
                        return as_bool(gen(rest));
                    }
                    // numeric arms: COALESCE(x where c IS TRUE, rest where NOT (c IS TRUE)); a NULL arm stays NULL because the
                    // other side is masked out on the same rows
                    Operand rest = e->i ? as_num(gen(e->args.back()), t) : null_num();
                    for (size_t k = n_arm; k-- > 0;) {
                        const ExprP c = is_true(e->args[2 * k]);
                        Operand cv = gen(c);
                        Operand x = as_num(gen(e->args[2 * k + 1]), t);
                        Operand keep_x = emit(PO_KEEPIF_N, x, cv, t);
                        Operand ncv = gen(node(Expr::UNARY, "NOT", {c}));
                        Operand keep_r = emit(PO_KEEPIF_N, rest, ncv, t);
                        rest = emit(PO_COALESCE_N, keep_x, keep_r, t);
                    }
                    return rest;
                }
                throw Error(TG_ERR_UNSUPPORTED, "function " + e->s + " is not supported by the predicate engine");
            }
            case Expr::BINARY: {
                const std::string& op = e->s;
                // the taller operand first (Sethi-Ullman order): its result then occupies ONE temporary while the other is
                // generated — right-leaning chains (desugared CASE / IN lists) stay within the four temporaries
                const bool right_first = height(e->args[1]) > height(e->args[0]);
                if (op == "AND" || op == "OR") {
                    Operand a, b;
                    if (right_first) {
                        b = as_bool(gen(e->args[1]));
                        a = as_bool(gen(e->args[0]));
                    } else {
                        a = as_bool(gen(e->args[0]));
                        b = as_bool(gen(e->args[1]));
                    }
                    return emit(op == "AND" ? PO_AND : PO_OR, a, b, PT_BOOL);
                }
                Operand a, b;
                if (right_first) {
                    b = gen(e->args[1]);
                    a = gen(e->args[0]);
                } else {
                    a = gen(e->args[0]);
                    b = gen(e->args[1]);
                }
                const bool arith = op == "+" || op == "-" || op == "*" || op == "/" || op == "%";
                if (a.type == PT_NULL || b.type == PT_NULL) {
                    release(a);
                    release(b);
                    if (arith) return null_num();
                    return null_bool();  // comparison with NULL is NULL
                }
                if (a.type == PT_BOOL || b.type == PT_BOOL) {
                    if (arith || a.type != b.type)
                        throw Error(TG_ERR_TYPE_MISMATCH, "cannot apply '" + op + "' to boolean and numeric operands");
                    if (op == "=") return emit(PO_EQ_B, a, b, PT_BOOL);
                    if (op == "<>") return emit(PO_NE_B, a, b, PT_BOOL);
                    throw Error(TG_ERR_UNSUPPORTED, "ordering comparison of booleans is not supported");
                }
                const bool use_f = a.type == PT_F64 || b.type == PT_F64;
                if (use_f) {
                    a = to_f64(a);
                    b = to_f64(b);
                }
                if (arith) {
                    uint8_t o;
                    if (use_f) {
                        if (op == "%") throw Error(TG_ERR_UNSUPPORTED, "% on floating point operands is not supported");
                        o = op == "+" ? PO_ADD_F : op == "-" ? PO_SUB_F : op == "*" ? PO_MUL_F : PO_DIV_F;
                    } else {
                        o = op == "+" ? PO_ADD_I : op == "-" ? PO_SUB_I : op == "*" ? PO_MUL_I
                          : op == "/" ? PO_DIV_I : PO_MOD_I;
                    }
                    return emit(o, a, b, use_f ? PT_F64 : PT_I64);
                }
                uint8_t o;
                if (use_f)
                    o = op == "=" ? PO_EQ_F : op == "<>" ? PO_NE_F : op == "<" ? PO_LT_F
                      : op == "<=" ? PO_LE_F : op == ">" ? PO_GT_F : PO_GE_F;
                else
                    o = op == "=" ? PO_EQ_I : op == "<>" ? PO_NE_I : op == "<" ? PO_LT_I
                      : op == "<=" ? PO_LE_I : op == ">" ? PO_GT_I : PO_GE_I;
                return emit(o, a, b, PT_BOOL);
            }
        }
        throw Error(TG_ERR_INTERNAL, "bad expression node");
    }
};

}  // namespace

ExprP parse_sql_expr(const std::string& text) {
    Parser p(text);
    ExprP e = p.parse_or();
    if (p.cur.k != Tok::END) p.err("unexpected trailing input");
    return e;
}

void collect_columns(const ExprP& e, std::vector<std::string>& out) {
    if (!e) return;
    if (e->kind == Expr::COL) {
        for (auto& c : out)
            if (c == e->s) return;
        out.push_back(e->s);
    }
    for (auto& a : e->args) collect_columns(a, out);
}

static void flatten(const ExprP& e, const std::string& op, std::vector<ExprP>& out) {
    if (e->kind == Expr::BINARY && e->s == op) {
        flatten(e->args[0], op, out);
        flatten(e->args[1], op, out);
    } else {
        out.push_back(e);
    }
}

bool try_compile_terms(const ExprP& e, const ColumnResolver& resolve, std::vector<ScanTerm>& terms, bool& is_or) {
    std::vector<ExprP> leaves;
    is_or = e->kind == Expr::BINARY && e->s == "OR";
    flatten(e, is_or ? "OR" : "AND", leaves);
    if (leaves.size() > (size_t)SCAN_UNIT_TERMS) return false;
    auto d2u = [](double d) {
        uint64_t u;
        memcpy(&u, &d, 8);
        return u;
    };
    std::vector<ScanTerm> out;
    for (auto& l : leaves) {
        ScanTerm t{};
        if (l->kind == Expr::IS && (l->s == "NULL" || l->s == "NOTNULL") && l->args[0]->kind == Expr::COL) {
            ColumnBinding b = resolve(l->args[0]->s);
            t.col = b.tile_col;
            t.kind = l->s == "NULL" ? TK_ISNULL : TK_NOTNULL;
            out.push_back(t);
            continue;
        }
        if (l->kind != Expr::BINARY) return false;
        const std::string& op = l->s;
        int mask = op == "<" ? 1 : op == "<=" ? 3 : op == "=" ? 2 : op == ">=" ? 6 : op == ">" ? 4 : op == "<>" ? 13 : -1;
        if (mask < 0) return false;
        ExprP c = l->args[0], v = l->args[1];
        if (c->kind != Expr::COL) {
            std::swap(c, v);
            // literal op col  ==  col (flipped op) literal
            mask = (mask & 2) | ((mask & 1) << 2) | ((mask & 4) >> 2) | (mask & 8);
        }
        if (c->kind != Expr::COL || (v->kind != Expr::LIT_I && v->kind != Expr::LIT_F)) return false;
        ColumnBinding b = resolve(c->s);
        t.col = b.tile_col;
        t.cmp_mask = mask;
        if (b.dtype == TG_FLOAT64) {
            t.kind = TK_F64;
            t.imm = d2u(v->kind == Expr::LIT_I ? (double)v->i : v->f);
        } else if (b.dtype == TG_INT64) {
            if (v->kind == Expr::LIT_I) {
                t.kind = TK_I64;
                t.imm = (uint64_t)v->i;
            } else {
                t.kind = TK_I64_AS_F64;
                t.imm = d2u(v->f);
            }
        } else {
            return false;
        }
        out.push_back(t);
    }
    terms = out;
    return true;
}

void compile_predicate(const ExprP& e, const ColumnResolver& resolve, std::vector<PredInstr>& code) {
    Compiler c{resolve, code};
    Operand r = c.gen(e);
    if (r.type == PT_NULL) r = c.null_bool();
    if (r.type != PT_BOOL) throw Error(TG_ERR_TYPE_MISMATCH, "predicate must be a boolean expression");
    // the predicate's value must end in B0
    if (!(r.kind == PK_BTEMP && r.idx == 0)) {
        c.release(r);
        PredInstr ins{};
        ins.op = PO_MOVB;
        ins.dst = 0;
        ins.a_kind = r.kind;
        ins.a_idx = r.idx;
        ins.b_kind = PK_NONE;
        ins.imm = r.imm;
        code.push_back(ins);
    }
}

}  // namespace tg
