#include "regex_dfa.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <unordered_map>

#include "common.hpp"

namespace tg {

namespace {

#include "unicode_tables.inc"

constexpr uint32_t MAX_CP = 0x10FFFF;

// ------------------------------------------------------------------ code point sets ----
struct CharSet {
    std::vector<std::pair<uint32_t, uint32_t>> r;  // sorted, disjoint, non-adjacent
    void add(uint32_t lo, uint32_t hi) { r.emplace_back(lo, hi); }
    void normalize() {
        std::sort(r.begin(), r.end());
        std::vector<std::pair<uint32_t, uint32_t>> o;
        for (auto& x : r) {
            if (!o.empty() && x.first <= o.back().second + 1) o.back().second = std::max(o.back().second, x.second);
            else o.push_back(x);
        }
        r.swap(o);
    }
    void add_table(const uint32_t (*t)[2], size_t n) {
        for (size_t i = 0; i < n; ++i) add(t[i][0], t[i][1]);
    }
    void add_set(const CharSet& o) {
        for (auto& x : o.r) r.push_back(x);
    }
    void negate() {
        normalize();
        std::vector<std::pair<uint32_t, uint32_t>> o;
        uint32_t next = 0;
        for (auto& x : r) {
            if (x.first > next) o.emplace_back(next, x.first - 1);
            next = x.second + 1;
        }
        if (next <= MAX_CP) o.emplace_back(next, MAX_CP);
        r.swap(o);
        // surrogates are not scalar values
        remove_range(0xD800, 0xDFFF);
    }
    void remove_range(uint32_t lo, uint32_t hi) {
        std::vector<std::pair<uint32_t, uint32_t>> o;
        for (auto& x : r) {
            if (x.second < lo || x.first > hi) {
                o.push_back(x);
                continue;
            }
            if (x.first < lo) o.emplace_back(x.first, lo - 1);
            if (x.second > hi) o.emplace_back(hi + 1, x.second);
        }
        r.swap(o);
    }
    void intersect(const CharSet& o) {  // A & B = ~(~A | ~B)
        CharSet a = *this, b = o;
        a.negate();
        b.negate();
        a.add_set(b);
        a.negate();
        r = a.r;
    }
    void subtract(const CharSet& o) {  // A - B = A & ~B
        CharSet b = o;
        b.negate();
        intersect(b);
    }
    bool contains(uint32_t c) const {
        for (auto& x : r)
            if (c >= x.first && c <= x.second) return true;
        return false;
    }
    // simple case folding (the crate's (?i) in Unicode mode): every member of the orbit of every code point of the set,
    // from the generated orbit table (K / k / KELVIN SIGN, s / S / LONG S, é / É, σ / ς / Σ, ..)
    void case_fold() {
        normalize();
        std::vector<std::pair<uint32_t, uint32_t>> extra;
        const size_t n_pairs = sizeof(UNI_CASE_PAIRS) / sizeof(UNI_CASE_PAIRS[0]);
        for (auto& x : r) {
            // first pair whose code point is >= x.first
            size_t lo = 0, hi = n_pairs;
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                if (UNI_CASE_PAIRS[mid][0] < x.first) lo = mid + 1;
                else hi = mid;
            }
            for (size_t i = lo; i < n_pairs && UNI_CASE_PAIRS[i][0] <= x.second; ++i) extra.emplace_back(UNI_CASE_PAIRS[i][1], UNI_CASE_PAIRS[i][1]);
        }
        for (auto& e : extra) r.push_back(e);
        normalize();
    }
};

// ------------------------------------------------------------------ AST ----
struct Node;
using NodeP = std::shared_ptr<Node>;
struct Node {
    enum K { EMPTY, SET, CAT, ALT, REP, BOL, EOL, MBOL, MEOL, WORDB, NWORDB } k = EMPTY;
    CharSet set;
    std::vector<NodeP> kids;
    int min = 0, max = -1;
};

[[noreturn]] void syntax_error(const std::string& m) {
    throw Error(TG_ERR_SECURITY, "Invalid regex pattern: regex parse error: " + m);
}
[[noreturn]] void unsupported(const std::string& m) {
    throw Error(TG_ERR_UNSUPPORTED, "regex construct not supported by the DFA engine: " + m);
}

struct Parser {
    std::vector<uint32_t> cp;  // pattern as code points
    size_t p = 0;
    bool icase, dotall = false, multiline = false, verbose = false;
    int depth = 0;

    Parser(const std::string& pat, bool ic) : icase(ic) {
        // decode UTF-8
        size_t i = 0;
        while (i < pat.size()) {
            unsigned char c = pat[i];
            uint32_t v;
            int n;
            if (c < 0x80) { v = c; n = 1; }
            else if ((c >> 5) == 6) { v = c & 0x1F; n = 2; }
            else if ((c >> 4) == 14) { v = c & 0x0F; n = 3; }
            else if ((c >> 3) == 30) { v = c & 0x07; n = 4; }
            else syntax_error("pattern is not valid UTF-8");
            if (i + n > pat.size()) syntax_error("pattern is not valid UTF-8");
            for (int k = 1; k < n; ++k) v = (v << 6) | (pat[i + k] & 0x3F);
            cp.push_back(v);
            i += n;
        }
    }
    bool eof() const { return p >= cp.size(); }
    uint32_t peek() const { return cp[p]; }
    bool looking_at(const char* s) const {
        size_t n = strlen(s);
        if (p + n > cp.size()) return false;
        for (size_t i = 0; i < n; ++i)
            if (cp[p + i] != (unsigned char)s[i]) return false;
        return true;
    }

    // (?x): whitespace and # comments between tokens are not part of the pattern
    void skip_verbose() {
        if (!verbose) return;
        while (!eof()) {
            const uint32_t c = peek();
            if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') ++p;
            else if (c == '#') {
                while (!eof() && peek() != '\n') ++p;
            } else break;
        }
    }

    static NodeP mk(Node::K k) {
        auto n = std::make_shared<Node>();
        n->k = k;
        return n;
    }
    NodeP set_node(CharSet s) {
        if (icase) s.case_fold();
        s.normalize();
        auto n = mk(Node::SET);
        n->set = std::move(s);
        return n;
    }
    NodeP lit(uint32_t c) {
        CharSet s;
        s.add(c, c);
        return set_node(std::move(s));
    }

    NodeP parse_alt() {
        std::vector<NodeP> alts;
        alts.push_back(parse_cat());
        while (!eof() && peek() == '|') {
            ++p;
            alts.push_back(parse_cat());
        }
        if (alts.size() == 1) return alts[0];
        auto n = mk(Node::ALT);
        n->kids = std::move(alts);
        return n;
    }
    NodeP parse_cat() {
        auto n = mk(Node::CAT);
        skip_verbose();
        while (!eof() && peek() != '|' && peek() != ')') {
            n->kids.push_back(parse_rep());
            skip_verbose();
        }
        if (n->kids.empty()) return mk(Node::EMPTY);
        if (n->kids.size() == 1) return n->kids[0];
        return n;
    }
    bool parse_int(int& v) {
        size_t s = p;
        long long x = 0;
        while (!eof() && peek() >= '0' && peek() <= '9') {
            x = x * 10 + (peek() - '0');
            if (x > 100000) syntax_error("repetition count too large");
            ++p;
        }
        if (p == s) return false;
        v = (int)x;
        return true;
    }
    NodeP parse_rep() {
        NodeP a = parse_atom();
        while (true) {
            skip_verbose();
            if (eof()) break;
            uint32_t c = peek();
            int mn, mx;
            if (c == '*') { mn = 0; mx = -1; ++p; }
            else if (c == '+') { mn = 1; mx = -1; ++p; }
            else if (c == '?') { mn = 0; mx = 1; ++p; }
            else if (c == '{') {
                size_t save = p;
                ++p;
                if (!parse_int(mn)) {
                    p = save;
                    syntax_error("repetition quantifier expects a valid decimal");
                }
                mx = mn;
                if (!eof() && peek() == ',') {
                    ++p;
                    if (!parse_int(mx)) mx = -1;
                }
                if (eof() || peek() != '}') syntax_error("unclosed counted repetition");
                ++p;
                if (mx != -1 && mx < mn) syntax_error("invalid repetition count range, the start must be <= the end");
            } else break;
            if (!eof() && peek() == '?') ++p;  // lazy: same language
            if (a->k == Node::BOL || a->k == Node::EOL || a->k == Node::MBOL || a->k == Node::MEOL || a->k == Node::WORDB || a->k == Node::NWORDB) {
                // repetition of an assertion: x{0,..} = optional (always satisfiable by empty); keep one copy if min>0
                if (mn == 0) a = mk(Node::EMPTY);
                continue;
            }
            if (a->k == Node::EMPTY && false) continue;
            auto r = mk(Node::REP);
            r->kids = {a};
            r->min = mn;
            r->max = mx;
            a = r;
        }
        return a;
    }

    void perl_class(uint32_t c, CharSet& s) {
        CharSet t;
        switch (c) {
            case 'd': case 'D': t.add_table(UNI_DIGIT, sizeof(UNI_DIGIT) / sizeof(UNI_DIGIT[0])); break;
            case 's': case 'S': t.add_table(UNI_SPACE, sizeof(UNI_SPACE) / sizeof(UNI_SPACE[0])); break;
            case 'w': case 'W': t.add_table(UNI_WORD, sizeof(UNI_WORD) / sizeof(UNI_WORD[0])); break;
        }
        if (c == 'D' || c == 'S' || c == 'W') t.negate();
        s.add_set(t);
    }
    // \p{Name}, \p{^Name}, \pL, \p{gc=Lu}, \p{Script=Greek}: General_Category values, binary properties and scripts of
    // the generated tables (loose matching: case, spaces, '_' and '-' are ignored, as in regex-syntax)
    void unicode_property(bool negated, CharSet& s) {
        std::string name;
        if (eof()) syntax_error("incomplete escape sequence, reached end of pattern prematurely");
        if (peek() == '{') {
            ++p;
            while (!eof() && peek() != '}') {
                if (peek() >= 0x80) syntax_error("Unicode property not found");
                name += (char)cp[p++];
            }
            if (eof()) syntax_error("unclosed Unicode class");
            ++p;
        } else {
            if (peek() >= 0x80) syntax_error("Unicode property not found");
            name = std::string(1, (char)cp[p++]);
        }
        if (!name.empty() && name[0] == '^') {
            negated = !negated;
            name.erase(0, 1);
        }
        std::string key, value = name;
        bool neq = false;
        size_t eq = name.find("!=");
        if (eq != std::string::npos) {
            key = name.substr(0, eq);
            value = name.substr(eq + 2);
            neq = true;
        } else if ((eq = name.find_first_of("=:")) != std::string::npos) {
            key = name.substr(0, eq);
            value = name.substr(eq + 1);
        }
        if (neq) negated = !negated;
        auto norm = [](const std::string& x) {
            std::string o;
            for (char ch : x)
                if (ch != ' ' && ch != '_' && ch != '-') o += (char)tolower((unsigned char)ch);
            return o;
        };
        const std::string k = norm(key), v = norm(value);
        char want = 0;  // 0 any kind; g general category, s script
        if (!k.empty()) {
            if (k == "gc" || k == "generalcategory") want = 'g';
            else if (k == "sc" || k == "script") want = 's';
            else if (k == "scx" || k == "scriptextensions") unsupported("Script_Extensions (\\p{scx=..})");
            else if (k == "age") unsupported("Unicode Age property");
            else syntax_error("Unicode property not found");
        }
        CharSet t;
        bool found = false;
        if (!want && v == "any") {
            t.add(0, MAX_CP);
            t.remove_range(0xD800, 0xDFFF);
            found = true;
        } else if (!want && v == "ascii") {
            t.add(0, 0x7F);
            found = true;
        } else if (!want && v == "assigned") {
            for (const UniProp& up : UNI_PROPS)
                if (std::string(up.name) == "cn") t.add_table(up.ranges, up.n);
            t.negate();
            found = true;
        } else {
            for (const UniProp& up : UNI_PROPS) {
                if (v != up.name) continue;
                if (want && up.kind != want) continue;
                t.add_table(up.ranges, up.n);
                found = true;
                break;
            }
        }
        if (!found) {
            // a syntactically fine name the tables do not carry (rarer scripts, Age, ..): say so instead of guessing
            unsupported("Unicode property '" + name + "' is not in the engine's tables");
        }
        if (negated) t.negate();
        s.add_set(t);
    }

    uint32_t parse_hex(int digits_fixed) {
        uint32_t v = 0;
        if (!eof() && peek() == '{') {
            ++p;
            int n = 0;
            while (!eof() && peek() != '}') {
                uint32_t c = peek();
                int d = c >= '0' && c <= '9' ? (int)(c - '0') : c >= 'a' && c <= 'f' ? (int)(c - 'a') + 10 : c >= 'A' && c <= 'F' ? (int)(c - 'A') + 10 : -1;
                if (d < 0) syntax_error("invalid hexadecimal digit");
                v = v * 16 + d;
                if (++n > 8) syntax_error("invalid hexadecimal literal");
                ++p;
            }
            if (eof() || n == 0) syntax_error("unclosed or empty hexadecimal literal");
            ++p;
        } else {
            for (int i = 0; i < digits_fixed; ++i) {
                if (eof()) syntax_error("incomplete hexadecimal escape");
                uint32_t c = peek();
                int d = c >= '0' && c <= '9' ? (int)(c - '0') : c >= 'a' && c <= 'f' ? (int)(c - 'a') + 10 : c >= 'A' && c <= 'F' ? (int)(c - 'A') + 10 : -1;
                if (d < 0) syntax_error("invalid hexadecimal digit");
                v = v * 16 + d;
                ++p;
            }
        }
        if (v > MAX_CP || (v >= 0xD800 && v <= 0xDFFF)) syntax_error("hexadecimal literal is not a Unicode scalar value");
        return v;
    }
    // after a backslash; returns true and sets `c` for a single literal, false if it added a class to `s`
    bool parse_escape(uint32_t& c, CharSet& s, bool in_class) {
        if (eof()) syntax_error("incomplete escape sequence, reached end of pattern prematurely");
        uint32_t e = cp[p++];
        switch (e) {
            case 'd': case 'D': case 's': case 'S': case 'w': case 'W': perl_class(e, s); return false;
            case 'n': c = '\n'; return true;
            case 't': c = '\t'; return true;
            case 'r': c = '\r'; return true;
            case 'f': c = '\f'; return true;
            case 'v': c = '\v'; return true;
            case 'a': c = 7; return true;
            case 'x': c = parse_hex(2); return true;
            case 'u': c = parse_hex(4); return true;
            case 'U': c = parse_hex(8); return true;
            case 'p': case 'P': unicode_property(e == 'P', s); return false;
            case 'b': case 'B':
                if (in_class) syntax_error("unrecognized escape sequence");
                if (e == 'b' && !eof() && peek() == '{') unsupported("special word boundary assertions (\\b{start}, ..)");
                c = e == 'b' ? 0xFFFFFFF2 : 0xFFFFFFF3;
                return true;
            case 'A': case 'z': case 'Z': case 'G': case 'K': case 'Q': case 'E': case 'C': case 'R': case 'X':
                if (!in_class && e == 'A') { c = 0xFFFFFFF0; return true; }
                if (!in_class && e == 'z') { c = 0xFFFFFFF1; return true; }
                syntax_error("unrecognized escape sequence");
            default: break;
        }
        if (e >= '0' && e <= '9') syntax_error("backreferences are not supported");
        if (e < 0x80 && (isalnum((int)e))) syntax_error("unrecognized escape sequence");
        c = e;  // escaped punctuation / non-ASCII
        return true;
    }

    // one item of a class union: a literal / range / escape class / POSIX class / nested class. Returns false at ']'
    // or at a set operator.
    bool parse_class_item(CharSet& s, bool first) {
        if (eof()) syntax_error("unclosed character class");
        uint32_t c = cp[p];
        if (c == ']' && !first) return false;
        if (!first && (looking_at("&&") || looking_at("--") || looking_at("~~"))) return false;
        if (c == '[') {
            if (looking_at("[:")) {
                // POSIX class
                size_t e = p + 2;
                bool pneg = false;
                if (e < cp.size() && cp[e] == '^') { pneg = true; ++e; }
                std::string name;
                while (e < cp.size() && cp[e] != ':') name += (char)cp[e++];
                if (e + 1 < cp.size() && cp[e] == ':' && cp[e + 1] == ']') {
                    CharSet t;
                    if (name == "alpha") { t.add('a', 'z'); t.add('A', 'Z'); }
                    else if (name == "digit") t.add('0', '9');
                    else if (name == "alnum") { t.add('a', 'z'); t.add('A', 'Z'); t.add('0', '9'); }
                    else if (name == "upper") t.add('A', 'Z');
                    else if (name == "lower") t.add('a', 'z');
                    else if (name == "space") { t.add('\t', '\r'); t.add(' ', ' '); }
                    else if (name == "blank") { t.add('\t', '\t'); t.add(' ', ' '); }
                    else if (name == "punct") { t.add('!', '/'); t.add(':', '@'); t.add('[', '`'); t.add('{', '~'); }
                    else if (name == "xdigit") { t.add('0', '9'); t.add('a', 'f'); t.add('A', 'F'); }
                    else if (name == "word") { t.add('a', 'z'); t.add('A', 'Z'); t.add('0', '9'); t.add('_', '_'); }
                    else if (name == "cntrl") { t.add(0, 0x1F); t.add(0x7F, 0x7F); }
                    else if (name == "print") t.add(' ', '~');
                    else if (name == "graph") t.add('!', '~');
                    else if (name == "ascii") t.add(0, 0x7F);
                    else syntax_error("unrecognized POSIX class");
                    if (pneg) t.negate();
                    s.add_set(t);
                    p = e + 2;
                    return true;
                }
            }
            // nested class: its set joins the union
            ++p;
            if (++depth > 200) syntax_error("exceeded the maximum number of nested parentheses/brackets");
            CharSet inner = parse_class_body();
            --depth;
            s.add_set(inner);
            return true;
        }
        uint32_t lo;
        ++p;
        if (c == '\\') {
            CharSet tmp;
            if (!parse_escape(lo, tmp, true)) {
                s.add_set(tmp);
                return true;
            }
        } else {
            lo = c;
        }
        // range?
        if (!eof() && peek() == '-' && p + 1 < cp.size() && cp[p + 1] != ']' && !looking_at("--")) {
            size_t save = p;
            ++p;
            uint32_t hi = cp[p++];
            if (hi == '\\') {
                CharSet tmp;
                if (!parse_escape(hi, tmp, true)) syntax_error("invalid character class range, the end must be a single character");
            } else if (hi == '[') {
                p = save;
                s.add(lo, lo);
                return true;
            }
            if (hi < lo) syntax_error("invalid character class range, the start must be <= the end");
            s.add(lo, hi);
        } else {
            s.add(lo, lo);
        }
        return true;
    }
    // after '[': '^'? union ( ('&&' | '--' | '~~') union )* ']' ; the operators associate to the left, juxtaposition
    // (union) binds tighter (regex-syntax's ClassSetBinaryOp)
    CharSet parse_class_body() {
        bool neg = false;
        if (!eof() && peek() == '^') {
            neg = true;
            ++p;
        }
        CharSet acc;
        bool have_acc = false, first = true;
        int pending = 0;  // 0 none, 1 &&, 2 --, 3 ~~
        while (true) {
            CharSet u;
            while (parse_class_item(u, first)) first = false;
            first = false;
            if (icase) u.case_fold();
            u.normalize();
            if (!have_acc) {
                acc = u;
                have_acc = true;
            } else if (pending == 1) {
                acc.intersect(u);
            } else if (pending == 2) {
                acc.subtract(u);
            } else {
                CharSet a = acc, b = u;
                a.subtract(u);
                b.subtract(acc);
                a.add_set(b);
                a.normalize();
                acc = a;
            }
            if (eof()) syntax_error("unclosed character class");
            if (cp[p] == ']') {
                ++p;
                break;
            }
            pending = looking_at("&&") ? 1 : looking_at("--") ? 2 : 3;
            p += 2;
        }
        if (neg) acc.negate();
        acc.normalize();
        return acc;
    }
    NodeP parse_class() {
        // '[' already consumed
        auto n = mk(Node::SET);
        n->set = parse_class_body();
        return n;
    }

    NodeP parse_group() {
        // '(' consumed
        bool save_icase = icase, save_dotall = dotall, save_multiline = multiline, save_verbose = verbose;
        if (!eof() && peek() == '?') {
            ++p;
            if (looking_at("P<") || (looking_at("<") && !looking_at("<=") && !looking_at("<!"))) {
                while (!eof() && peek() != '>') ++p;
                if (eof()) syntax_error("unclosed capture group name");
                ++p;
            } else if (looking_at("=") || looking_at("!") || looking_at("<=") || looking_at("<!")) {
                syntax_error("look-around, including look-ahead and look-behind, is not supported");
            } else {
                bool on = true;
                bool any = false;
                while (!eof() && peek() != ':' && peek() != ')') {
                    uint32_t f = cp[p++];
                    any = true;
                    if (f == '-') { on = false; continue; }
                    if (f == 'i') icase = on;
                    else if (f == 's') dotall = on;
                    else if (f == 'U') {}
                    else if (f == 'u') { if (!on) unsupported("(?-u) byte mode"); }
                    else if (f == 'm') multiline = on;
                    else if (f == 'x') verbose = on;
                    else if (f == 'R') { if (on) unsupported("(?R) CRLF mode"); }
                    else syntax_error("unrecognized flag");
                }
                if (eof()) syntax_error("unclosed group");
                if (peek() == ')') {
                    if (!any) syntax_error("missing flags");
                    ++p;
                    // flags apply to the rest of the enclosing group
                    return mk(Node::EMPTY);
                }
                ++p;  // ':'
            }
        }
        if (++depth > 200) syntax_error("exceeded the maximum number of nested parentheses/brackets");
        NodeP n = parse_alt();
        --depth;
        if (eof() || peek() != ')') syntax_error("unclosed group");
        ++p;
        icase = save_icase;
        dotall = save_dotall;
        multiline = save_multiline;
        verbose = save_verbose;
        return n;
    }

    NodeP parse_atom() {
        uint32_t c = cp[p++];
        switch (c) {
            case '(': return parse_group();
            case '[': return parse_class();
            case '.': {
                CharSet s;
                if (dotall) s.add(0, MAX_CP);
                else {
                    s.add(0, '\n' - 1);
                    s.add('\n' + 1, MAX_CP);
                }
                s.remove_range(0xD800, 0xDFFF);
                s.normalize();
                auto n = mk(Node::SET);
                n->set = std::move(s);
                return n;
            }
            case '^': return mk(multiline ? Node::MBOL : Node::BOL);
            case '$': return mk(multiline ? Node::MEOL : Node::EOL);
            case '*': case '+': case '?': syntax_error("repetition operator missing expression");
            case '{': syntax_error("repetition operator missing expression");
            case '\\': {
                uint32_t l;
                CharSet s;
                if (!parse_escape(l, s, false)) {
                    if (icase) s.case_fold();
                    s.normalize();
                    auto n = mk(Node::SET);
                    n->set = std::move(s);
                    return n;
                }
                if (l == 0xFFFFFFF0) return mk(Node::BOL);
                if (l == 0xFFFFFFF1) return mk(Node::EOL);
                if (l == 0xFFFFFFF2) return mk(Node::WORDB);
                if (l == 0xFFFFFFF3) return mk(Node::NWORDB);
                return lit(l);
            }
            default: return lit(c);
        }
    }

    NodeP parse() {
        // flags set by a bare (?i) at top level persist to the end of the pattern
        NodeP n = parse_alt();
        if (!eof()) {
            if (peek() == ')') syntax_error("unopened group");
            syntax_error("unexpected character");
        }
        return n;
    }
};

// ------------------------------------------------------------------ NFA ----
struct Trans {
    uint8_t lo, hi;
    int to;
};
// zero-width assertions that look at the neighbouring characters (everything but the plain ^ / $ of the haystack's ends)
enum AssertKind : uint8_t { AS_MBOL, AS_MEOL, AS_WORDB, AS_NWORDB };
struct NState {
    std::vector<Trans> t;
    std::vector<int> eps, eps_bol, eps_eol;
    std::vector<std::pair<uint8_t, int>> asserts;  // (AssertKind, target)
};
struct Nfa {
    std::vector<NState> st;
    bool uses_word_boundary = false, uses_multiline_eol = false;
    int add() {
        if (st.size() > 400000) throw Error(TG_ERR_UNSUPPORTED, "regex is too large for the DFA engine");
        st.emplace_back();
        return (int)st.size() - 1;
    }
};

struct Frag {
    int s, e;
};

void utf8_encode(uint32_t c, uint8_t* b, int& n) {
    if (c < 0x80) { b[0] = (uint8_t)c; n = 1; }
    else if (c < 0x800) { b[0] = 0xC0 | (c >> 6); b[1] = 0x80 | (c & 0x3F); n = 2; }
    else if (c < 0x10000) { b[0] = 0xE0 | (c >> 12); b[1] = 0x80 | ((c >> 6) & 0x3F); b[2] = 0x80 | (c & 0x3F); n = 3; }
    else { b[0] = 0xF0 | (c >> 18); b[1] = 0x80 | ((c >> 12) & 0x3F); b[2] = 0x80 | ((c >> 6) & 0x3F); b[3] = 0x80 | (c & 0x3F); n = 4; }
}

// split [lo,hi] (same encoded length, no surrogates) into sequences of byte ranges
void utf8_split(uint32_t lo, uint32_t hi, Nfa& nfa, int s, int e) {
    uint8_t a[4], b[4];
    int na, nb;
    utf8_encode(lo, a, na);
    utf8_encode(hi, b, nb);
    for (int i = 1; i < na; ++i) {
        uint32_t m = (1u << (6 * i)) - 1;
        if ((lo & ~m) != (hi & ~m)) {
            if ((lo & m) != 0) {
                utf8_split(lo, lo | m, nfa, s, e);
                utf8_split((lo | m) + 1, hi, nfa, s, e);
                return;
            }
            if ((hi & m) != m) {
                utf8_split(lo, (hi & ~m) - 1, nfa, s, e);
                utf8_split(hi & ~m, hi, nfa, s, e);
                return;
            }
        }
    }
    int cur = s;
    for (int i = 0; i < na; ++i) {
        int nxt = i == na - 1 ? e : nfa.add();
        nfa.st[cur].t.push_back(Trans{a[i], b[i], nxt});
        cur = nxt;
    }
}

void add_cp_range(uint32_t lo, uint32_t hi, Nfa& nfa, int s, int e) {
    static const uint32_t bounds[] = {0x7F, 0x7FF, 0xFFFF, MAX_CP};
    // drop surrogates
    if (lo <= 0xDFFF && hi >= 0xD800) {
        if (lo < 0xD800) add_cp_range(lo, 0xD7FF, nfa, s, e);
        if (hi > 0xDFFF) add_cp_range(0xE000, hi, nfa, s, e);
        return;
    }
    uint32_t start = lo;
    for (uint32_t bnd : bounds) {
        if (start > hi) break;
        if (start <= bnd) {
            uint32_t end = std::min(hi, bnd);
            utf8_split(start, end, nfa, s, e);
            start = end + 1;
        }
    }
}

Frag build(const NodeP& n, Nfa& nfa) {
    switch (n->k) {
        case Node::EMPTY: {
            int s = nfa.add(), e = nfa.add();
            nfa.st[s].eps.push_back(e);
            return {s, e};
        }
        case Node::SET: {
            int s = nfa.add(), e = nfa.add();
            for (auto& r : n->set.r) add_cp_range(r.first, r.second, nfa, s, e);
            return {s, e};
        }
        case Node::BOL: {
            int s = nfa.add(), e = nfa.add();
            nfa.st[s].eps_bol.push_back(e);
            return {s, e};
        }
        case Node::EOL: {
            int s = nfa.add(), e = nfa.add();
            nfa.st[s].eps_eol.push_back(e);
            return {s, e};
        }
        case Node::MBOL: case Node::MEOL: case Node::WORDB: case Node::NWORDB: {
            int s = nfa.add(), e = nfa.add();
            const uint8_t kind = n->k == Node::MBOL ? AS_MBOL : n->k == Node::MEOL ? AS_MEOL : n->k == Node::WORDB ? AS_WORDB : AS_NWORDB;
            nfa.st[s].asserts.emplace_back(kind, e);
            if (kind == AS_WORDB || kind == AS_NWORDB) nfa.uses_word_boundary = true;
            if (kind == AS_MEOL) nfa.uses_multiline_eol = true;
            return {s, e};
        }
        case Node::CAT: {
            Frag f = build(n->kids[0], nfa);
            for (size_t i = 1; i < n->kids.size(); ++i) {
                Frag g = build(n->kids[i], nfa);
                nfa.st[f.e].eps.push_back(g.s);
                f.e = g.e;
            }
            return f;
        }
        case Node::ALT: {
            int s = nfa.add(), e = nfa.add();
            for (auto& k : n->kids) {
                Frag g = build(k, nfa);
                nfa.st[s].eps.push_back(g.s);
                nfa.st[g.e].eps.push_back(e);
            }
            return {s, e};
        }
        case Node::REP: {
            int s = nfa.add();
            int cur = s;
            for (int i = 0; i < n->min; ++i) {
                Frag g = build(n->kids[0], nfa);
                nfa.st[cur].eps.push_back(g.s);
                cur = g.e;
            }
            if (n->max == -1) {
                // star
                int loop = nfa.add(), e = nfa.add();
                nfa.st[cur].eps.push_back(loop);
                Frag g = build(n->kids[0], nfa);
                nfa.st[loop].eps.push_back(g.s);
                nfa.st[loop].eps.push_back(e);
                nfa.st[g.e].eps.push_back(loop);
                return {s, e};
            }
            int e = nfa.add();
            for (int i = n->min; i < n->max; ++i) {
                Frag g = build(n->kids[0], nfa);
                nfa.st[cur].eps.push_back(g.s);
                nfa.st[cur].eps.push_back(e);
                cur = g.e;
            }
            nfa.st[cur].eps.push_back(e);
            return {s, e};
        }
    }
    throw Error(TG_ERR_INTERNAL, "bad regex node");
}

// ------------------------------------------------------------------ DFA ----
struct SetHash {
    size_t operator()(const std::vector<int>& v) const {
        size_t h = 1469598103934665603ull;
        for (int x : v) h = (h ^ (size_t)(x + 7)) * 1099511628211ull;
        return h;
    }
};

// ---- subset construction with one-character look-around ------------------------------------------------------------
// The items of a DFA state are NFA threads:
//   q >= 0            a thread at NFA state q
//   ITEM_START (-1)   marker "at the start of the haystack" (only in the initial state)
//   >= PROD_BASE      a thread at q that still owes a condition on the NEXT character (an "obligation"): the assertions
//                     that look ahead (\b, \B, (?m)$) cannot be decided when the thread passes them, so the thread goes
//                     on with the obligation attached and the character it consumes next is run, byte by byte, through
//                     a recogniser of the required class in lock step — product item (q, r) with r a state of that
//                     recogniser. When the recogniser accepts (at the character's last byte) the obligation is
//                     discharged and the item becomes the plain thread again; if it cannot move, the thread dies.
//                     A thread that reached the final state with an obligation waits there (the rest of the haystack
//                     does not matter for a search) until the next character settles it.
// What precedes the current position is known without extra state: the tracker (an NFA loop over \w | \W characters that
// is always part of the set) leaves its "a word character just ended" / "a non-word character just ended" state in the
// set exactly at character boundaries, and "the previous byte was \n" is known when the closure runs right after the step.
constexpr int ITEM_START = -1, PROD_BASE = 1 << 28;
enum Obl : uint8_t { OB_WORD = 0, OB_NONWORD = 1, OB_NEWLINE = 2 };  // NONWORD and NEWLINE are also satisfied by the end of the haystack

struct Look {
    // recognisers: start / accept state per obligation kind (built into the same Nfa, never entered by the pattern)
    int r_start[3] = {-1, -1, -1}, r_accept[3] = {-1, -1, -1};
    int t0 = -1, t_word = -1, t_nonword = -1;  // tracker
    int final_state = -1;
    struct Prod {
        int q, r;
        uint8_t kind;
    };
    std::vector<Prod> prods;
    std::map<std::tuple<int, int, int>, int> prod_ids;
    int prod(int q, int r, uint8_t kind) {
        auto key = std::make_tuple(q, r, (int)kind);
        auto it = prod_ids.find(key);
        if (it != prod_ids.end()) return it->second;
        if (prods.size() > 2000000) throw Error(TG_ERR_UNSUPPORTED, "regex is too large for the DFA engine");
        const int id = PROD_BASE + (int)prods.size();
        prods.push_back(Prod{q, r, kind});
        prod_ids.emplace(key, id);
        return id;
    }
    bool fresh(const Prod& pr) const { return pr.r == r_start[pr.kind]; }
};

struct Ctx {
    bool at_start = false, prev_nl = false;
    int prev_word = -1;  // 1 a word character just ended, 0 a non-word character (or the start), -1 inside a character
};

// both obligations must hold for the same next character: returns false when they cannot
bool combine_obl(uint8_t have, uint8_t need, uint8_t& out) {
    if (have == need) {
        out = have;
        return true;
    }
    if ((have == OB_NONWORD && need == OB_NEWLINE) || (have == OB_NEWLINE && need == OB_NONWORD)) {
        out = OB_NEWLINE;  // \n is a non-word character
        return true;
    }
    return false;
}

void closure(const Nfa& nfa, Look& L, std::vector<int>& set, const Ctx& ctx, bool eol) {
    std::vector<int> stack(set.begin(), set.end());
    std::unordered_map<int, char> seen;
    for (int x : set) seen[x] = 1;
    auto push = [&](int y) {
        if (seen.emplace(y, 1).second) {
            set.push_back(y);
            stack.push_back(y);
        }
    };
    while (!stack.empty()) {
        const int x = stack.back();
        stack.pop_back();
        if (x == ITEM_START) continue;
        int q;
        bool has_obl = false;
        uint8_t kind = 0;
        if (x >= PROD_BASE) {
            const Look::Prod pr = L.prods[x - PROD_BASE];
            if (!L.fresh(pr)) continue;  // in the middle of a character: nothing zero-width can happen
            q = pr.q;
            has_obl = true;
            kind = pr.kind;
        } else {
            q = x;
        }
        auto go = [&](int y) { push(has_obl ? L.prod(y, L.r_start[kind], kind) : y); };
        auto go_with = [&](int y, uint8_t need) {
            uint8_t k = need;
            if (has_obl && !combine_obl(kind, need, k)) return;
            push(L.prod(y, L.r_start[k], k));
        };
        const NState& st = nfa.st[q];
        for (int y : st.eps) go(y);
        if (ctx.at_start)
            for (int y : st.eps_bol) go(y);
        if (eol)
            for (int y : st.eps_eol) go(y);
        for (auto& as : st.asserts) {
            switch (as.first) {
                case AS_MBOL:
                    if (ctx.at_start || ctx.prev_nl) go(as.second);
                    break;
                case AS_MEOL: go_with(as.second, OB_NEWLINE); break;
                case AS_WORDB:
                    if (ctx.prev_word == 0) go_with(as.second, OB_WORD);
                    else if (ctx.prev_word == 1) go_with(as.second, OB_NONWORD);
                    break;
                case AS_NWORDB:
                    if (ctx.prev_word == 0) go_with(as.second, OB_NONWORD);
                    else if (ctx.prev_word == 1) go_with(as.second, OB_WORD);
                    break;
            }
        }
    }
    std::sort(set.begin(), set.end());
}

}  // namespace

bool Dfa::match(const uint8_t* s, int64_t len) const {
    uint32_t st = start;
    for (int64_t i = 0; i < len; ++i) {
        if (st == DFA_MATCH) return true;
        if (st == DFA_DEAD) return false;
        st = next[(size_t)st * n_classes + class_of[s[i]]];
    }
    return st == DFA_MATCH || accept_end[st];
}

static void add_char_class(const CharSet& cs, Nfa& nfa, int s, int e) {
    for (auto& r : cs.r) add_cp_range(r.first, r.second, nfa, s, e);
}

static Dfa compile_regex_uncached(const std::string& pattern, bool case_insensitive) {
    Parser ps(pattern, case_insensitive);
    NodeP ast = ps.parse();
    Nfa nfa;
    // unanchored search prefix: LOOP consumes any byte and re-enters the pattern (not at start any more)
    const int loop = nfa.add();
    nfa.st[loop].t.push_back(Trans{0, 255, loop});
    Frag f = build(ast, nfa);
    nfa.st[loop].eps.push_back(f.s);
    const int final_state = f.e;
    Look L;
    L.final_state = final_state;
    if (nfa.uses_word_boundary) {
        CharSet w, nw;
        w.add_table(UNI_WORD, sizeof(UNI_WORD) / sizeof(UNI_WORD[0]));
        w.normalize();
        nw = w;
        nw.negate();
        for (int k = 0; k < 2; ++k) {
            L.r_start[k] = nfa.add();
            L.r_accept[k] = nfa.add();
            add_char_class(k == OB_WORD ? w : nw, nfa, L.r_start[k], L.r_accept[k]);
        }
        L.t0 = nfa.add();
        L.t_word = nfa.add();
        L.t_nonword = nfa.add();
        add_char_class(w, nfa, L.t0, L.t_word);
        add_char_class(nw, nfa, L.t0, L.t_nonword);
        nfa.st[L.t_word].eps.push_back(L.t0);
        nfa.st[L.t_nonword].eps.push_back(L.t0);
    }
    if (nfa.uses_multiline_eol) {
        L.r_start[OB_NEWLINE] = nfa.add();
        L.r_accept[OB_NEWLINE] = nfa.add();
        nfa.st[L.r_start[OB_NEWLINE]].t.push_back(Trans{'\n', '\n', L.r_accept[OB_NEWLINE]});
        if (L.r_start[OB_NONWORD] < 0) {  // \B / \b absent: NONWORD can still arise? no - only from them; nothing to do
        }
    }

    // byte classes from all transition boundaries
    bool boundary[257];
    memset(boundary, 0, sizeof(boundary));
    boundary[0] = true;
    boundary['\n'] = boundary['\n' + 1] = true;  // the closure's "previous byte was a newline" context
    for (auto& s : nfa.st)
        for (auto& t : s.t) {
            boundary[t.lo] = true;
            boundary[(int)t.hi + 1] = true;
        }
    Dfa d;
    int ncls = -1;
    std::vector<uint8_t> rep;  // representative byte per class
    for (int b = 0; b < 256; ++b) {
        if (boundary[b]) {
            ++ncls;
            rep.push_back((uint8_t)b);
        }
        d.class_of[b] = (uint8_t)ncls;
    }
    ++ncls;

    std::unordered_map<std::vector<int>, uint32_t, SetHash> ids;
    std::vector<std::vector<int>> sets;
    std::vector<std::vector<uint32_t>> trans;  // raw ids (before dead/min), -1u for MATCH
    std::vector<uint8_t> acc_end;
    const uint32_t RAW_MATCH = 0xFFFFFFFFu;

    auto has_final = [&](const std::vector<int>& s) { return std::binary_search(s.begin(), s.end(), final_state); };
    // the haystack ends here: the final state, or the final state with an obligation the end satisfies
    auto accepts_at_end = [&](const std::vector<int>& s) {
        if (has_final(s)) return true;
        for (int x : s)
            if (x >= PROD_BASE) {
                const Look::Prod& pr = L.prods[x - PROD_BASE];
                if (pr.q == final_state && L.fresh(pr) && pr.kind != OB_WORD) return true;
            }
        return false;
    };
    auto intern = [&](std::vector<int>& s) -> uint32_t {
        if (has_final(s)) return RAW_MATCH;
        auto it = ids.find(s);
        if (it != ids.end()) return it->second;
        if (sets.size() >= DFA_MAX_STATES)
            throw Error(TG_ERR_UNSUPPORTED, "regex needs more than " + std::to_string(DFA_MAX_STATES) + " DFA states");
        uint32_t id = (uint32_t)sets.size();
        ids.emplace(s, id);
        sets.push_back(s);
        return id;
    };
    auto context_of = [&](const std::vector<int>& s, bool at_start, bool prev_nl) {
        Ctx c;
        c.at_start = at_start;
        c.prev_nl = prev_nl;
        if (at_start) c.prev_word = 0;
        else if (L.t0 < 0) c.prev_word = -1;
        else if (std::find(s.begin(), s.end(), L.t_word) != s.end()) c.prev_word = 1;
        else if (std::find(s.begin(), s.end(), L.t_nonword) != s.end()) c.prev_word = 0;
        return c;
    };

    std::vector<int> init = {ITEM_START, loop, f.s};  // the marker: "at the start of the haystack"
    if (L.t0 >= 0) init.push_back(L.t0);
    closure(nfa, L, init, context_of(init, true, false), false);
    uint32_t raw_start = intern(init);
    if (raw_start == RAW_MATCH) {
        // matches the empty prefix of every haystack
        d.n_states = 2;
        d.n_classes = 1;
        memset(d.class_of, 0, 256);
        d.next = {0, 1};
        d.accept_end = {0, 1};
        d.start = DFA_MATCH;
        return d;
    }
    for (size_t i = 0; i < sets.size(); ++i) {
        const std::vector<int> cur = sets[i];
        const bool at_start = !cur.empty() && cur[0] == ITEM_START;
        // accept if the haystack ends here. The set is already closed under everything but the end-of-haystack edges;
        // those do not depend on the previous character, except through assertions behind them, which see the same context
        // the state was closed with — recovered from the tracker items (the previous byte being \n only matters to (?m)^,
        // which is satisfied after the end of input exactly when it was at the closure: an un-passed (?m)^ stays un-passed)
        std::vector<int> e = cur;
        Ctx ectx = context_of(cur, at_start, false);
        closure(nfa, L, e, ectx, true);
        acc_end.push_back(accepts_at_end(e) ? 1 : 0);
        std::vector<uint32_t> row(ncls);
        for (int c = 0; c < ncls; ++c) {
            const uint8_t b = rep[c];
            std::vector<int> nxt;
            for (int x : cur) {
                if (x == ITEM_START) continue;
                if (x >= PROD_BASE) {
                    const Look::Prod pr = L.prods[x - PROD_BASE];
                    for (auto& rt : nfa.st[pr.r].t) {
                        if (b < rt.lo || b > rt.hi) continue;
                        const bool done = rt.to == L.r_accept[pr.kind];
                        if (pr.q == final_state) {  // the match is complete but for the obligation: any byte keeps it alive
                            nxt.push_back(done ? final_state : L.prod(final_state, rt.to, pr.kind));
                            continue;
                        }
                        for (auto& t : nfa.st[pr.q].t)
                            if (b >= t.lo && b <= t.hi) nxt.push_back(done ? t.to : L.prod(t.to, rt.to, pr.kind));
                    }
                    continue;
                }
                for (auto& t : nfa.st[x].t)
                    if (b >= t.lo && b <= t.hi) nxt.push_back(t.to);
            }
            std::sort(nxt.begin(), nxt.end());
            nxt.erase(std::unique(nxt.begin(), nxt.end()), nxt.end());
            closure(nfa, L, nxt, context_of(nxt, false, b == '\n'), false);
            row[c] = intern(nxt);
        }
        trans.push_back(std::move(row));
    }

    // ---- dead-state detection: states that can never reach MATCH or an accept-at-end state ----
    const size_t n = sets.size();
    std::vector<uint8_t> live(n, 0);
    bool changed = true;
    for (size_t i = 0; i < n; ++i)
        if (acc_end[i]) live[i] = 1;
    while (changed) {
        changed = false;
        for (size_t i = 0; i < n; ++i) {
            if (live[i]) continue;
            for (uint32_t t : trans[i])
                if (t == RAW_MATCH || live[t]) {
                    live[i] = 1;
                    changed = true;
                    break;
                }
        }
    }
    // ---- Moore minimisation over live states (+ DEAD, MATCH) ----
    // block ids: 0 DEAD, 1 MATCH, then by accept_end
    std::vector<uint32_t> block(n);
    for (size_t i = 0; i < n; ++i) block[i] = !live[i] ? 0 : (acc_end[i] ? 3 : 2);
    uint32_t nblocks = 0;  // forces a second pass: the seed partition may have fewer than 2 live blocks
    while (true) {
        std::map<std::vector<uint32_t>, uint32_t> sig_ids;
        std::vector<uint32_t> nb(n);
        uint32_t next_id = 2;
        for (size_t i = 0; i < n; ++i) {
            if (!live[i]) {
                nb[i] = 0;
                continue;
            }
            std::vector<uint32_t> sig;
            sig.reserve(ncls + 1);
            sig.push_back(block[i]);
            for (uint32_t t : trans[i]) sig.push_back(t == RAW_MATCH ? 1 : block[t]);
            auto it = sig_ids.find(sig);
            if (it == sig_ids.end()) it = sig_ids.emplace(std::move(sig), next_id++).first;
            nb[i] = it->second;
        }
        bool same = next_id == nblocks;
        block.swap(nb);
        nblocks = next_id;
        if (same) break;
    }
    d.n_states = nblocks;
    d.n_classes = (uint32_t)ncls;
    d.next.assign((size_t)nblocks * ncls, 0);
    d.accept_end.assign(nblocks, 0);
    for (int c = 0; c < ncls; ++c) {
        d.next[(size_t)DFA_DEAD * ncls + c] = DFA_DEAD;
        d.next[(size_t)DFA_MATCH * ncls + c] = DFA_MATCH;
    }
    d.accept_end[DFA_MATCH] = 1;
    for (size_t i = 0; i < n; ++i) {
        if (!live[i]) continue;
        uint32_t b = block[i];
        d.accept_end[b] = acc_end[i];
        for (int c = 0; c < ncls; ++c) {
            uint32_t t = trans[i][c];
            d.next[(size_t)b * ncls + c] = (uint16_t)(t == RAW_MATCH ? DFA_MATCH : block[t]);
        }
    }
    d.start = live[raw_start] ? block[raw_start] : DFA_DEAD;
    // merge byte classes with identical columns
    {
        std::map<std::vector<uint16_t>, uint8_t> col_ids;
        std::vector<uint8_t> remap(ncls);
        std::vector<std::vector<uint16_t>> cols;
        for (int c = 0; c < ncls; ++c) {
            std::vector<uint16_t> col(nblocks);
            for (uint32_t s = 0; s < nblocks; ++s) col[s] = d.next[(size_t)s * ncls + c];
            auto it = col_ids.find(col);
            if (it == col_ids.end()) {
                it = col_ids.emplace(col, (uint8_t)cols.size()).first;
                cols.push_back(col);
            }
            remap[c] = it->second;
        }
        const uint32_t nc2 = (uint32_t)cols.size();
        std::vector<uint16_t> nx((size_t)nblocks * nc2);
        for (uint32_t c = 0; c < nc2; ++c)
            for (uint32_t s = 0; s < nblocks; ++s) nx[(size_t)s * nc2 + c] = cols[c][s];
        d.next.swap(nx);
        for (int b = 0; b < 256; ++b) d.class_of[b] = remap[d.class_of[b]];
        d.n_classes = nc2;
    }
    return d;
}

// compiled DFAs are cached like the reference caches pattern strings (format.rs:183-184)
Dfa compile_regex(const std::string& pattern, bool case_insensitive) {
    static std::mutex mu;
    static std::map<std::pair<std::string, bool>, Dfa> cache;
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find({pattern, case_insensitive});
        if (it != cache.end()) return it->second;
    }
    Dfa d = compile_regex_uncached(pattern, case_insensitive);
    std::lock_guard<std::mutex> g(mu);
    cache[{pattern, case_insensitive}] = d;
    return d;
}

void regex_check_supported(const std::string& pattern, bool icase) { (void)compile_regex(pattern, icase); }

}  // namespace tg
