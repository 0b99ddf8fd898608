// O(1) host logic restated from the reference: Rust number formatting, Assertion, LogicalOperator,
// SqlSecurity, FormatType patterns. No CUDA here.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>

#include "common.hpp"

namespace tg {

static thread_local std::string g_last_error;

void set_last_error(const std::string& msg) { g_last_error = msg; }
tg_status fail(tg_status code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

// Rust's `impl Display for f64`: shortest digits that round-trip, never scientific notation,
// integral values print without a fractional part, NaN -> "NaN", infinities -> "inf"/"-inf".
std::string fmt_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    // shortest round-trip digits (scientific), then laid out positionally with zero padding — Rust pads the
    // shortest digits with zeros instead of printing the exact binary expansion of huge / tiny values
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string sci(buf, r.ptr);
    bool neg = false;
    size_t i = 0;
    if (sci[0] == '-') {
        neg = true;
        i = 1;
    }
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t k = i; k < epos; ++k)
        if (sci[k] != '.') digits += sci[k];
    const int exp10 = std::stoi(sci.substr(epos + 1));
    std::string out;
    const int point = exp10 + 1;  // digits before the decimal point
    if (point <= 0) {
        out = "0.";
        out.append((size_t)(-point), '0');
        out += digits;
    } else if ((size_t)point >= digits.size()) {
        out = digits;
        out.append((size_t)point - digits.size(), '0');
    } else {
        out = digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
    }
    if (out.find('.') != std::string::npos) {
        while (!out.empty() && out.back() == '0') out.pop_back();
        if (!out.empty() && out.back() == '.') out.pop_back();
    }
    if (out.empty()) out = "0";
    return neg ? "-" + out : out;
}

// serde_json's f64 (ryu::Buffer::format_finite): shortest round-trip digits; positional for decimal exponents
// -5 < kk <= 16 with at least ".0", otherwise d[.ddd]e[-]x; non-finite values serialise as null.
std::string json_f64(double v) {
    if (std::isnan(v) || std::isinf(v)) return "null";
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
    std::string sci(buf, r.ptr);
    bool neg = false;
    size_t i = 0;
    if (sci[0] == '-') {
        neg = true;
        i = 1;
    }
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t k = i; k < epos; ++k)
        if (sci[k] != '.') digits += sci[k];
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int exp10 = std::stoi(sci.substr(epos + 1));
    const int length = (int)digits.size();
    const int kk = exp10 + 1;          // position of the decimal point relative to the first digit
    const int k = kk - length;         // value = digits * 10^k
    std::string out;
    if (0 <= k && kk <= 16) {
        out = digits + std::string((size_t)k, '0') + ".0";
    } else if (0 < kk && kk <= 16) {
        out = digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
    } else if (-5 < kk && kk <= 0) {
        out = "0." + std::string((size_t)(-kk), '0') + digits;
    } else if (length == 1) {
        out = digits + "e" + std::to_string(kk - 1);
    } else {
        out = digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    }
    return neg ? "-" + out : out;
}

// `{:.N}`: exact decimal expansion rounded half-to-even at N places — what glibc printf does.
std::string fmt_f64_prec(double v, int prec) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[512];
    snprintf(buf, sizeof(buf), "%.*f", prec, v);
    return buf;
}

std::string fmt_i64(int64_t v) { return std::to_string(v); }

// constraints/assertion.rs:48-61
bool assertion_evaluate(const tg_assertion& a, double value) {
    const double EPSILON = 1e-10;
    switch (a.kind) {
        case TG_ASSERT_EQUALS: return std::fabs(value - a.a) < EPSILON;
        case TG_ASSERT_NOT_EQUALS: return std::fabs(value - a.a) >= EPSILON;
        case TG_ASSERT_GREATER_THAN: return value > a.a;
        case TG_ASSERT_GREATER_THAN_OR_EQUAL: return value >= a.a;
        case TG_ASSERT_LESS_THAN: return value < a.a;
        case TG_ASSERT_LESS_THAN_OR_EQUAL: return value <= a.a;
        case TG_ASSERT_BETWEEN: return value >= a.a && value <= a.b;
        case TG_ASSERT_NOT_BETWEEN: return value < a.a || value > a.b;
    }
    return false;
}

// constraints/assertion.rs:64-75
std::string assertion_description(const tg_assertion& a) {
    switch (a.kind) {
        case TG_ASSERT_EQUALS: return "equals " + fmt_f64(a.a);
        case TG_ASSERT_NOT_EQUALS: return "not equals " + fmt_f64(a.a);
        case TG_ASSERT_GREATER_THAN: return "greater than " + fmt_f64(a.a);
        case TG_ASSERT_GREATER_THAN_OR_EQUAL: return "greater than or equal to " + fmt_f64(a.a);
        case TG_ASSERT_LESS_THAN: return "less than " + fmt_f64(a.a);
        case TG_ASSERT_LESS_THAN_OR_EQUAL: return "less than or equal to " + fmt_f64(a.a);
        case TG_ASSERT_BETWEEN: return "between " + fmt_f64(a.a) + " and " + fmt_f64(a.b);
        case TG_ASSERT_NOT_BETWEEN: return "not between " + fmt_f64(a.a) + " and " + fmt_f64(a.b);
    }
    return "?";
}

// core/logical.rs:69-89
bool logical_evaluate(int op, int n, const std::vector<bool>& results) {
    if (results.empty()) {
        switch (op) {
            case TG_OP_ALL: return true;
            case TG_OP_ANY: return false;
            case TG_OP_EXACTLY: return n == 0;
            case TG_OP_AT_LEAST: return n == 0;
            case TG_OP_AT_MOST: return true;
        }
        return false;
    }
    size_t true_count = std::count(results.begin(), results.end(), true);
    switch (op) {
        case TG_OP_ALL: return true_count == results.size();
        case TG_OP_ANY: return true_count > 0;
        case TG_OP_EXACTLY: return true_count == (size_t)n;
        case TG_OP_AT_LEAST: return true_count >= (size_t)n;
        case TG_OP_AT_MOST: return true_count <= (size_t)n;
    }
    return false;
}

// core/logical.rs:92-100
std::string logical_description(int op, int n) {
    switch (op) {
        case TG_OP_ALL: return "all";
        case TG_OP_ANY: return "any";
        case TG_OP_EXACTLY: return "exactly " + std::to_string(n);
        case TG_OP_AT_LEAST: return "at least " + std::to_string(n);
        case TG_OP_AT_MOST: return "at most " + std::to_string(n);
    }
    return "?";
}

static std::string to_lower_ascii(std::string s) {
    for (auto& c : s) c = (char)tolower((unsigned char)c);
    return s;
}
static std::string to_upper_ascii(std::string s) {
    for (auto& c : s) c = (char)toupper((unsigned char)c);
    return s;
}

static bool ident_start(char c) { return isalpha((unsigned char)c) || c == '_' || c == '"'; }
static bool ident_cont(char c) { return isalnum((unsigned char)c) || c == '_' || c == '"'; }

// security.rs:89-137 validate_identifier + :190-255 check_dangerous_patterns
void validate_identifier(const std::string& id) {
    if (id.empty()) throw Error(TG_ERR_SECURITY, "SQL identifier cannot be empty");
    if (id.size() > 128) throw Error(TG_ERR_SECURITY, "SQL identifier too long (max 128 characters)");
    if (id.find('\0') != std::string::npos)
        throw Error(TG_ERR_SECURITY, "SQL identifier cannot contain null bytes");
    // ^[a-zA-Z_"][a-zA-Z0-9_"]*(\.[a-zA-Z_"][a-zA-Z0-9_"]*)*$
    bool ok = true;
    size_t i = 0;
    while (true) {
        if (i >= id.size() || !ident_start(id[i])) {
            ok = false;
            break;
        }
        ++i;
        while (i < id.size() && ident_cont(id[i])) ++i;
        if (i == id.size()) break;
        if (id[i] != '.') {
            ok = false;
            break;
        }
        ++i;
    }
    if (!ok)
        throw Error(TG_ERR_SECURITY,
                    "Invalid SQL identifier format: '" + id +
                        "'. Identifiers must start with a letter or underscore and contain only "
                        "letters, numbers, underscores, and dots");
    std::string lower = to_lower_ascii(id);
    for (const char* p : {";", "--", "/*", "*/"}) {
        if (lower.find(p) != std::string::npos)
            throw Error(TG_ERR_SECURITY,
                        std::string("SQL identifier contains dangerous character sequence: '") + p + "'");
    }
    if (lower.rfind("xp_", 0) == 0 || lower.rfind("sp_", 0) == 0)
        throw Error(TG_ERR_SECURITY, "SQL identifier looks like a system stored procedure");
    static const char* injection[] = {"union ",  "union_",  "select ", "select_",  "insert ",  "insert_",
                                      "update ", "update_", "delete ", "delete_",  "drop ",    "drop_",
                                      "create ", "alter ",  "exec ",   "execute ", "declare ", "cursor ",
                                      "fetch ",  "open ",   "close "};
    for (const char* p : injection) {
        if (lower.find(p) != std::string::npos) {
            std::string kw(p);
            while (!kw.empty() && kw.back() == '_') kw.pop_back();
            while (!kw.empty() && kw.back() == ' ') kw.pop_back();
            throw Error(TG_ERR_SECURITY,
                        "SQL identifier contains suspicious SQL keyword pattern: '" + kw + "'");
        }
    }
}

// security.rs:152-183 (text-level checks; syntax validity is checked by the DFA compiler) + :258-281
void validate_regex_pattern_text(const std::string& p) {
    if (p.size() > 1000) throw Error(TG_ERR_SECURITY, "Regex pattern too long (max 1000 characters)");
    if (p.find('\0') != std::string::npos)
        throw Error(TG_ERR_SECURITY, "Regex pattern cannot contain null bytes");
    for (const char* d : {"(.*)*", "(.*)+", "(a+)+", "(a*)*"}) {
        if (p.find(d) != std::string::npos)
            throw Error(TG_ERR_SECURITY, "Regex pattern might cause ReDoS attack");
    }
}

static bool is_word(char c) { return isalnum((unsigned char)c) || c == '_'; }

// constraints/custom_sql.rs:100-190
void validate_sql_expression(const std::string& sql) {
    std::string up = to_upper_ascii(sql);
    static const char* kws[] = {"DROP",     "DELETE",   "INSERT",    "UPDATE", "CREATE",      "ALTER", "TRUNCATE",
                                "GRANT",    "REVOKE",   "EXECUTE",   "EXEC",   "CALL",        "MERGE", "REPLACE",
                                "RENAME",   "MODIFY",   "SET",       "COMMIT", "ROLLBACK",    "SAVEPOINT",
                                "BEGIN",    "START",    "TRANSACTION", "LOCK", "UNLOCK"};
    // The reference iterates a HashSet (unspecified order); report the first keyword in source order.
    size_t best_pos = std::string::npos;
    const char* best_kw = nullptr;
    for (const char* kw : kws) {
        size_t L = strlen(kw), pos = 0;
        while ((pos = up.find(kw, pos)) != std::string::npos) {
            bool lb = pos == 0 || !is_word(up[pos - 1]);
            bool rb = pos + L >= up.size() || !is_word(up[pos + L]);
            if (lb && rb) {
                if (pos < best_pos) {
                    best_pos = pos;
                    best_kw = kw;
                }
                break;
            }
            pos += 1;
        }
    }
    if (best_kw)
        throw Error(TG_ERR_VALIDATION,
                    std::string("SQL expression contains forbidden operation: ") + best_kw);
    if (sql.find(';') != std::string::npos)
        throw Error(TG_ERR_VALIDATION, "SQL expression cannot contain semicolons");
    if (sql.find("--") != std::string::npos || sql.find("/*") != std::string::npos ||
        sql.find("*/") != std::string::npos)
        throw Error(TG_ERR_VALIDATION, "SQL expression cannot contain comments");
}

// constraints/format.rs:217-307 — the pattern strings are data the engine must agree on with the
// reference byte for byte (they are the regex the user's column is checked against).
std::string format_pattern(int kind, const char* arg, int flag) {
    std::string a = arg ? arg : "";
    switch (kind) {
        case TG_FMT_REGEX: return a;
        case TG_FMT_EMAIL:
            return R"(^[a-zA-Z0-9.!#$%&'*+/=?^_`{|}~-]+@[a-zA-Z0-9](?:[a-zA-Z0-9-]{0,61}[a-zA-Z0-9])?(?:\.[a-zA-Z0-9](?:[a-zA-Z0-9-]{0,61}[a-zA-Z0-9])?)*$)";
        case TG_FMT_URL:
            if (flag)
                return R"(^https?://(?:localhost|(?:[a-zA-Z0-9.-]+\.?[a-zA-Z]{2,}|(?:\d{1,3}\.){3}\d{1,3}))(?::\d+)?(?:/[^\s]*)?$)";
            return R"(^https?://[a-zA-Z0-9.-]+\.[a-zA-Z]{2,}(?::\d+)?(?:/[^\s]*)?$)";
        case TG_FMT_CREDIT_CARD:
            return R"(^(?:4[0-9]{12}(?:[0-9]{3})?|5[1-5][0-9]{14}|3[47][0-9]{13}|3[0-9]{13}|6(?:011|5[0-9]{2})[0-9]{12})$|^(?:\d{4}[-\s]?){3}\d{4}$)";
        case TG_FMT_PHONE:
            if (arg && (a == "US" || a == "CA"))
                return R"(^(\+?1[-.\s]?)?\(?([0-9]{3})\)?[-.\s]?([0-9]{3})[-.\s]?([0-9]{4})$)";
            if (arg && a == "UK")
                return R"(^(\+44\s?)?(?:\(?0\d{4}\)?\s?\d{6}|\(?0\d{3}\)?\s?\d{7}|\(?0\d{2}\)?\s?\d{8})$)";
            if (arg && a == "DE") return R"(^(\+49\s?)?(?:\(?0\d{2,5}\)?\s?\d{4,12})$)";
            if (arg && a == "FR") return R"(^(\+33\s?)?(?:\(?0\d{1}\)?\s?\d{8})$)";
            return R"(^[\+]?[1-9][\d]{0,15}$)";
        case TG_FMT_POSTAL_CODE:
            if (a == "US") return R"(^\d{5}(-\d{4})?$)";
            if (a == "CA") return R"(^[A-Za-z]\d[A-Za-z][ -]?\d[A-Za-z]\d$)";
            if (a == "UK") return R"(^[A-Z]{1,2}\d[A-Z\d]?\s?\d[A-Z]{2}$)";
            if (a == "DE") return R"(^\d{5}$)";
            if (a == "FR") return R"(^\d{5}$)";
            if (a == "JP") return R"(^\d{3}-\d{4}$)";
            if (a == "AU") return R"(^\d{4}$)";
            return R"(^[A-Za-z0-9\s-]{3,10}$)";
        case TG_FMT_UUID:
            return R"(^[0-9a-fA-F]{8}-[0-9a-fA-F]{4}-[1-5][0-9a-fA-F]{3}-[89abAB][0-9a-fA-F]{3}-[0-9a-fA-F]{12}$)";
        case TG_FMT_IPV4:
            return R"(^(?:(?:25[0-5]|2[0-4][0-9]|[01]?[0-9][0-9]?)\.){3}(?:25[0-5]|2[0-4][0-9]|[01]?[0-9][0-9]?)$)";
        case TG_FMT_IPV6:
            return R"(^([0-9a-fA-F]{0,4}:){1,7}([0-9a-fA-F]{0,4})?$|^::$|^::1$|^([0-9a-fA-F]{1,4}:)*::([0-9a-fA-F]{1,4}:)*[0-9a-fA-F]{1,4}$)";
        case TG_FMT_JSON: return R"(^\s*[\{\[].*[\}\]]\s*$)";
        case TG_FMT_ISO8601:
            return R"(^\d{4}-\d{2}-\d{2}T\d{2}:\d{2}:\d{2}(?:\.\d+)?(?:Z|[+-]\d{2}:\d{2})$)";
        case TG_FMT_SSN:
            return R"(^(00[1-9]|0[1-9][0-9]|[1-5][0-9]{2}|6[0-5][0-9]|66[0-5]|667|66[89]|6[7-9][0-9]|[7-8][0-9]{2})-?(0[1-9]|[1-9][0-9])-?(000[1-9]|00[1-9][0-9]|0[1-9][0-9]{2}|[1-9][0-9]{3})$)";
    }
    throw Error(TG_ERR_INVALID_ARG, "unknown format kind");
}

// constraints/format.rs:310-326
std::string format_name(int kind) {
    switch (kind) {
        case TG_FMT_REGEX: return "regex";
        case TG_FMT_EMAIL: return "email";
        case TG_FMT_URL: return "url";
        case TG_FMT_CREDIT_CARD: return "credit_card";
        case TG_FMT_PHONE: return "phone";
        case TG_FMT_POSTAL_CODE: return "postal_code";
        case TG_FMT_UUID: return "uuid";
        case TG_FMT_IPV4: return "ipv4";
        case TG_FMT_IPV6: return "ipv6";
        case TG_FMT_JSON: return "json";
        case TG_FMT_ISO8601: return "iso8601_datetime";
        case TG_FMT_SSN: return "social_security_number";
    }
    return "?";
}

// constraints/format.rs:329-361
std::string format_description(int kind, const std::string& pattern, const char* arg, int flag) {
    switch (kind) {
        case TG_FMT_REGEX: return "matches pattern '" + pattern + "'";
        case TG_FMT_EMAIL: return "are valid email addresses";
        case TG_FMT_URL: return flag ? "are valid URLs (including localhost)" : "are valid URLs";
        case TG_FMT_CREDIT_CARD:
            return flag ? "contain credit card number patterns" : "are valid credit card numbers";
        case TG_FMT_PHONE:
            return arg ? std::string("are valid ") + arg + " phone numbers" : "are valid phone numbers";
        case TG_FMT_POSTAL_CODE: return std::string("are valid ") + (arg ? arg : "") + " postal codes";
        case TG_FMT_UUID: return "are valid UUIDs";
        case TG_FMT_IPV4: return "are valid IPv4 addresses";
        case TG_FMT_IPV6: return "are valid IPv6 addresses";
        case TG_FMT_JSON: return "are valid JSON documents";
        case TG_FMT_ISO8601: return "are valid ISO 8601 date-time strings";
        case TG_FMT_SSN: return "contain Social Security Number patterns";
    }
    return "?";
}

}  // namespace tg
