#include "regex_dfa.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
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
    bool contains(uint32_t c) const {
        for (auto& x : r)
            if (c >= x.first && c <= x.second) return true;
        return false;
    }
    // simple case folding: ASCII letters, plus the two non-ASCII code points that fold to ASCII letters
    void case_fold() {
        normalize();
        std::vector<std::pair<uint32_t, uint32_t>> extra;
        for (auto& x : r) {
            uint32_t lo = std::max<uint32_t>(x.first, 'a'), hi = std::min<uint32_t>(x.second, 'z');
            if (lo <= hi) extra.emplace_back(lo - 32, hi - 32);
            lo = std::max<uint32_t>(x.first, 'A');
            hi = std::min<uint32_t>(x.second, 'Z');
            if (lo <= hi) extra.emplace_back(lo + 32, hi + 32);
        }
        for (auto& e : extra) r.push_back(e);
        normalize();
        if (contains('k') || contains(0x212A)) {
            add('k', 'k');
            add('K', 'K');
            add(0x212A, 0x212A);
        }
        if (contains('s') || contains(0x17F)) {
            add('s', 's');
            add('S', 'S');
            add(0x17F, 0x17F);
        }
        normalize();
    }
};

// ------------------------------------------------------------------ AST ----
struct Node;
using NodeP = std::shared_ptr<Node>;
struct Node {
    enum K { EMPTY, SET, CAT, ALT, REP, BOL, EOL } k = EMPTY;
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
    bool icase, dotall = false;
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
        while (!eof() && peek() != '|' && peek() != ')') n->kids.push_back(parse_rep());
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
        while (!eof()) {
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
            if (a->k == Node::BOL || a->k == Node::EOL) {
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
            case 'p': case 'P': unsupported("Unicode property classes (\\p)");
            case 'b': case 'B':
                if (in_class) syntax_error("unrecognized escape sequence");
                unsupported("word boundary assertions (\\b, \\B)");
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

    NodeP parse_class() {
        // at '[' already consumed
        bool neg = false;
        if (!eof() && peek() == '^') {
            neg = true;
            ++p;
        }
        CharSet s;
        bool first = true;
        while (true) {
            if (eof()) syntax_error("unclosed character class");
            uint32_t c = cp[p];
            if (c == ']' && !first) {
                ++p;
                break;
            }
            first = false;
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
                        continue;
                    }
                }
                unsupported("nested character classes");
            }
            if (looking_at("&&") || looking_at("--") || looking_at("~~")) {
                // `--` could be a literal '-' followed by a range start only in odd patterns; the crate treats
                // these as set operators
                if (!(c == '-' && p + 2 < cp.size() && cp[p + 2] == ']' && false)) unsupported("character class set operations");
            }
            uint32_t lo;
            ++p;
            if (c == '\\') {
                CharSet tmp;
                if (!parse_escape(lo, tmp, true)) {
                    s.add_set(tmp);
                    continue;
                }
            } else {
                lo = c;
            }
            // range?
            if (!eof() && peek() == '-' && p + 1 < cp.size() && cp[p + 1] != ']') {
                size_t save = p;
                ++p;
                uint32_t hi = cp[p++];
                if (hi == '\\') {
                    CharSet tmp;
                    if (!parse_escape(hi, tmp, true)) syntax_error("invalid character class range, the end must be a single character");
                } else if (hi == '[') {
                    p = save;
                    s.add(lo, lo);
                    continue;
                }
                if (hi < lo) syntax_error("invalid character class range, the start must be <= the end");
                s.add(lo, hi);
            } else {
                s.add(lo, lo);
            }
        }
        if (icase) s.case_fold();
        if (neg) {
            s.negate();
        }
        s.normalize();
        auto n = mk(Node::SET);
        n->set = std::move(s);
        return n;
    }

    NodeP parse_group() {
        // '(' consumed
        bool save_icase = icase, save_dotall = dotall;
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
                    else if (f == 'm') { if (on) unsupported("(?m) multi-line mode"); }
                    else if (f == 'x') { if (on) unsupported("(?x) verbose mode"); }
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
            case '^': return mk(Node::BOL);
            case '$': return mk(Node::EOL);
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
struct NState {
    std::vector<Trans> t;
    std::vector<int> eps, eps_bol, eps_eol;
};
struct Nfa {
    std::vector<NState> st;
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

void closure(const Nfa& nfa, std::vector<int>& set, bool bol, bool eol, std::vector<uint8_t>& mark) {
    std::vector<int> stack(set.begin(), set.end());
    for (int x : set)
        if (x >= 0) mark[x] = 1;
    while (!stack.empty()) {
        int x = stack.back();
        stack.pop_back();
        if (x < 0) continue;
        auto push = [&](int y) {
            if (!mark[y]) {
                mark[y] = 1;
                set.push_back(y);
                stack.push_back(y);
            }
        };
        for (int y : nfa.st[x].eps) push(y);
        if (bol)
            for (int y : nfa.st[x].eps_bol) push(y);
        if (eol)
            for (int y : nfa.st[x].eps_eol) push(y);
    }
    for (int x : set)
        if (x >= 0) mark[x] = 0;
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

    // byte classes from all transition boundaries
    bool boundary[257];
    memset(boundary, 0, sizeof(boundary));
    boundary[0] = true;
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

    std::vector<uint8_t> mark(nfa.st.size(), 0);
    std::unordered_map<std::vector<int>, uint32_t, SetHash> ids;
    std::vector<std::vector<int>> sets;
    std::vector<std::vector<uint32_t>> trans;  // raw ids (before dead/min), -1u for MATCH
    std::vector<uint8_t> acc_end;
    const uint32_t RAW_MATCH = 0xFFFFFFFFu;

    auto has_final = [&](const std::vector<int>& s) { return std::binary_search(s.begin(), s.end(), final_state); };
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

    std::vector<int> init = {-1, loop, f.s};  // -1 marks "at start of haystack"
    closure(nfa, init, true, false, mark);
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
        const bool at_start = !cur.empty() && cur[0] == -1;
        // accept if the haystack ends here
        std::vector<int> e = cur;
        closure(nfa, e, at_start, true, mark);
        acc_end.push_back(has_final(e) ? 1 : 0);
        std::vector<uint32_t> row(ncls);
        for (int c = 0; c < ncls; ++c) {
            const uint8_t b = rep[c];
            std::vector<int> nxt;
            for (int x : cur) {
                if (x < 0) continue;
                for (auto& t : nfa.st[x].t)
                    if (b >= t.lo && b <= t.hi) nxt.push_back(t.to);
            }
            std::sort(nxt.begin(), nxt.end());
            nxt.erase(std::unique(nxt.begin(), nxt.end()), nxt.end());
            closure(nfa, nxt, false, false, mark);
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
