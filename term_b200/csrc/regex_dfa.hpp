// Host regex -> byte DFA compiler for the string kernel (K2).
//
// Semantics follow the Rust `regex` crate as the reference uses it through DataFusion's `~` / `~*`
// (constraints/format.rs:756-776 -> arrow-string regexp_is_match -> Regex::is_match): unanchored SEARCH
// unless the pattern anchors itself, Unicode mode (classes and `.` range over scalar values encoded as
// UTF-8), `$`/`^` match only at the very end/start of the haystack (no multi-line), `.` excludes \n.
// Unsupported syntax (look-around and back-references do not exist in the crate either; here also \b,
// (?m), (?x), class set operations) raises TG_ERR_UNSUPPORTED so nothing silently diverges.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace tg {

struct Dfa {
    // state ids: 0 = DEAD (no match possible), 1 = MATCH (absorbing, match already found), 2.. normal
    uint32_t n_states = 0;
    uint32_t n_classes = 0;
    uint32_t start = 0;
    uint8_t class_of[256];
    std::vector<uint16_t> next;        // [n_states * n_classes]
    std::vector<uint8_t> accept_end;   // [n_states] 1 if the haystack ending here is a match
    bool match(const uint8_t* s, int64_t len) const;
};

constexpr uint32_t DFA_DEAD = 0, DFA_MATCH = 1;
constexpr uint32_t DFA_MAX_STATES = 8192;

// throws tg::Error (TG_ERR_SECURITY "Invalid regex pattern: ..." for syntax errors like the reference,
// TG_ERR_UNSUPPORTED for valid-but-unsupported constructs or state blow-up)
Dfa compile_regex(const std::string& pattern, bool case_insensitive);

}  // namespace tg
