// Predicate grammar for `satisfies` / ComplianceAnalyzer: parser + compiler to the scan kernel's
// 3-address code. Declared subset of SQL (SURVEY §7 hard part d): column refs, numeric / boolean /
// NULL literals, + - * / %, unary -, comparisons, AND / OR / NOT, IS [NOT] NULL, IS [NOT] TRUE|FALSE,
// [NOT] BETWEEN, [NOT] IN (...), ABS(); over Utf8 columns: the six comparisons with a string literal or another Utf8
// column, [NOT] LIKE 'pattern', LENGTH / CHAR_LENGTH / CHARACTER_LENGTH / OCTET_LENGTH (materialised as virtual columns
// by the engine before the scan, engine.cu rewrite_string_compares); CASE (searched and simple form; boolean arms desugar to
// three-valued AND / OR, numeric arms use PO_KEEPIF_N + PO_COALESCE_N), COALESCE over numeric operands, CAST to the
// floating-point types (and to integer types of integer operands); string / DATE '..' / TIMESTAMP '..' literals compared with a
// date or timestamp column are cast to the column's type by the engine (engine.cu). Anything else -> TG_ERR_UNSUPPORTED.
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "common.hpp"
#include "scan_defs.h"

namespace tg {

struct Expr;
using ExprP = std::shared_ptr<Expr>;

struct Expr {
    enum Kind { LIT_I, LIT_F, LIT_S, LIT_B, LIT_NULL, COL, UNARY, BINARY, IS, FUNC } kind;
    int64_t i = 0;
    double f = 0;
    bool b = false;
    std::string s;   // column name, string literal, operator or function name
    std::vector<ExprP> args;
};

ExprP parse_sql_expr(const std::string& text);                 // throws Error(TG_ERR_UNSUPPORTED,...)
void collect_columns(const ExprP& e, std::vector<std::string>& out);

enum PredType { PT_I64, PT_F64, PT_BOOL, PT_NULL };

struct ColumnBinding {
    int tile_col;  // index in the scan tile
    int dtype;     // tg_dtype
};
// resolve(name) returns the binding or throws Error(TG_ERR_COLUMN_NOT_FOUND,...)
using ColumnResolver = std::function<ColumnBinding(const std::string&)>;

// Fast path: if the predicate is an AND (or an OR) of at most SCAN_UNIT_TERMS terms of the form
// `col cmp literal`, `literal cmp col`, `col IS [NOT] NULL`, fills `terms` (col = resolver's tile_col) and
// returns true; otherwise returns false and the caller compiles the general code.
bool try_compile_terms(const ExprP& e, const ColumnResolver& resolve, std::vector<ScanTerm>& terms, bool& is_or);

// Appends instructions to `code`; result of the predicate ends in temp 0.
void compile_predicate(const ExprP& e, const ColumnResolver& resolve, std::vector<PredInstr>& code);

}  // namespace tg
