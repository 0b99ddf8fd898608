// Internal types shared by the host library. Not part of the C ABI.
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/termgpu.h"

namespace tg {

struct Error {
    tg_status code;
    std::string msg;
    Error(tg_status c, std::string m) : code(c), msg(std::move(m)) {}
};

void set_last_error(const std::string& msg);
tg_status fail(tg_status code, const std::string& msg);

// ---- Rust-compatible formatting (messages must match the reference byte for byte) ----
std::string fmt_f64(double v);                 // `{}` / `{v}` Display of f64
std::string fmt_f64_prec(double v, int prec);  // `{:.N}`
std::string fmt_i64(int64_t v);
std::string json_f64(double v);               // serde_json / ryu rendering of an f64

// ---- Assertion (constraints/assertion.rs) ----
bool assertion_evaluate(const tg_assertion& a, double value);
std::string assertion_description(const tg_assertion& a);

// ---- LogicalOperator (core/logical.rs) ----
bool logical_evaluate(int op, int n, const std::vector<bool>& results);
std::string logical_description(int op, int n);

// ---- SqlSecurity (security.rs) ----
void validate_identifier(const std::string& id);        // throws Error(TG_ERR_SECURITY)
void validate_regex_pattern_text(const std::string& p); // length / NUL / ReDoS substrings
void validate_sql_expression(const std::string& e);     // custom_sql.rs:100-190

std::string format_pattern(int format_kind, const char* arg, int flag);
std::string format_name(int format_kind);
std::string format_description(int format_kind, const std::string& pattern, const char* arg, int flag);

}  // namespace tg
