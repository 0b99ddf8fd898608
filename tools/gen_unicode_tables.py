#!/usr/bin/env python3
"""Generate term_b200/csrc/unicode_tables.inc: code point ranges of the Perl classes as the Rust
`regex` crate defines them in Unicode mode (regex-syntax: \\d = Nd, \\s = White_Space,
\\w = Alphabetic + M + Nd + Pc + Join_Control) and of the Unicode properties \\p{..} accepts there:
every General_Category value (short and long names, the one-letter groups and LC), the common binary
properties and the scripts listed below. Source of truth here is the `regex` PyPI module's property
tables (a newer Unicode than this Python's unicodedata). The Unicode version may differ from the
reference's regex-syntax 0.8.8 tables by a few newly assigned code points; parity corpora are ASCII
plus a pinned non-ASCII set.

    python tools/gen_unicode_tables.py term_b200/csrc/unicode_tables.inc
"""
import sys, unicodedata
import regex

ALL = "".join(chr(cp) for cp in range(0x110000) if not (0xD800 <= cp <= 0xDFFF))


def ranges_of(pattern):
    rx = regex.compile(pattern)
    out, start, prev = [], None, None
    for m in rx.finditer(ALL):
        cp = ord(m.group())
        if start is None:
            start = prev = cp
        elif cp == prev + 1:
            prev = cp
        else:
            out.append((start, prev))
            start = prev = cp
    if start is not None:
        out.append((start, prev))
    return out


GC = {  # short -> long (UAX #44)
    "Lu": "Uppercase_Letter", "Ll": "Lowercase_Letter", "Lt": "Titlecase_Letter", "Lm": "Modifier_Letter", "Lo": "Other_Letter",
    "Mn": "Nonspacing_Mark", "Mc": "Spacing_Mark", "Me": "Enclosing_Mark", "Nd": "Decimal_Number", "Nl": "Letter_Number",
    "No": "Other_Number", "Pc": "Connector_Punctuation", "Pd": "Dash_Punctuation", "Ps": "Open_Punctuation",
    "Pe": "Close_Punctuation", "Pi": "Initial_Punctuation", "Pf": "Final_Punctuation", "Po": "Other_Punctuation",
    "Sm": "Math_Symbol", "Sc": "Currency_Symbol", "Sk": "Modifier_Symbol", "So": "Other_Symbol", "Zs": "Space_Separator",
    "Zl": "Line_Separator", "Zp": "Paragraph_Separator", "Cc": "Control", "Cf": "Format", "Co": "Private_Use", "Cn": "Unassigned",
    "L": "Letter", "M": "Mark", "N": "Number", "P": "Punctuation", "S": "Symbol", "Z": "Separator", "C": "Other", "LC": "Cased_Letter",
}
GC_ALIASES = {"Nd": ["digit"], "P": ["punct"], "Cc": ["cntrl"], "M": ["Combining_Mark"]}
BINARY = ["Alphabetic", "Lowercase", "Uppercase", "White_Space", "Cased", "Case_Ignorable", "Hex_Digit", "ASCII_Hex_Digit", "Dash",
          "Ideographic", "Join_Control", "Math", "Noncharacter_Code_Point", "Quotation_Mark", "Emoji", "Extended_Pictographic",
          "ID_Start", "ID_Continue", "XID_Start", "XID_Continue", "Default_Ignorable_Code_Point", "Diacritic", "Extender"]
BINARY_ALIASES = {"White_Space": ["space", "WSpace"], "Alphabetic": ["Alpha"], "Lowercase": ["Lower"], "Uppercase": ["Upper"],
                  "Hex_Digit": ["Hex"], "ASCII_Hex_Digit": ["AHex"], "Ideographic": ["Ideo"], "Join_Control": ["Join_C"],
                  "Noncharacter_Code_Point": ["NChar"], "Quotation_Mark": ["QMark"], "ID_Start": ["IDS"], "ID_Continue": ["IDC"],
                  "XID_Start": ["XIDS"], "XID_Continue": ["XIDC"], "Default_Ignorable_Code_Point": ["DI"], "Diacritic": ["Dia"],
                  "Extender": ["Ext"], "Case_Ignorable": ["CI"], "Extended_Pictographic": ["ExtPict"]}
SCRIPTS = {"Latin": "Latn", "Greek": "Grek", "Cyrillic": "Cyrl", "Armenian": "Armn", "Hebrew": "Hebr", "Arabic": "Arab", "Syriac": "Syrc",
           "Thaana": "Thaa", "Devanagari": "Deva", "Bengali": "Beng", "Gurmukhi": "Guru", "Gujarati": "Gujr", "Oriya": "Orya",
           "Tamil": "Taml", "Telugu": "Telu", "Kannada": "Knda", "Malayalam": "Mlym", "Sinhala": "Sinh", "Thai": "Thai", "Lao": "Laoo",
           "Tibetan": "Tibt", "Myanmar": "Mymr", "Georgian": "Geor", "Hangul": "Hang", "Ethiopic": "Ethi", "Cherokee": "Cher",
           "Khmer": "Khmr", "Mongolian": "Mong", "Hiragana": "Hira", "Katakana": "Kana", "Bopomofo": "Bopo", "Han": "Hani",
           "Common": "Zyyy", "Inherited": "Zinh", "Braille": "Brai", "Coptic": "Copt", "Gothic": "Goth", "Runic": "Runr"}


def case_orbits():
    """simple case folding orbits (what regex-syntax's case folding table holds): code points connected by single-code-point
    lower / upper / title mappings; returns sorted (cp, other member) pairs for every cp whose orbit has more than one member"""
    parent = {}

    def find(x):
        while parent.get(x, x) != x:
            parent[x] = parent.get(parent[x], parent[x])
            x = parent[x]
        return x

    def union(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)

    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        for m in (ch.lower(), ch.upper(), ch.title()):
            if len(m) == 1 and m != ch:
                union(cp, ord(m))
    groups = {}
    for cp in list(parent.keys()):
        groups.setdefault(find(cp), set()).add(cp)
    for root in list(groups):
        groups[root].add(root)
    pairs = []
    for members in groups.values():
        if len(members) < 2:
            continue
        for a in members:
            for b in members:
                if a != b:
                    pairs.append((a, b))
    return sorted(pairs)


def norm(name):
    return "".join(ch for ch in name.lower() if ch not in " _-")


def main(path):
    tabs = {"UNI_DIGIT": ranges_of(r"\p{Nd}"), "UNI_SPACE": ranges_of(r"\p{White_Space}"),
            "UNI_WORD": ranges_of(r"[\p{Alphabetic}\p{M}\p{Nd}\p{Pc}\p{Join_Control}]")}
    props = []  # (table name, [normalised lookup names], kind)
    for short, long_ in GC.items():
        t = "UNI_GC_" + short.upper() + ("_GROUP" if len(short) == 1 or short == "LC" else "")
        tabs[t] = ranges_of(r"\p{gc=%s}" % short)
        props.append((t, [norm(short), norm(long_)] + [norm(a) for a in GC_ALIASES.get(short, [])], "gc"))
    for b in BINARY:
        t = "UNI_BP_" + b.upper()
        tabs[t] = ranges_of(r"\p{%s}" % b)
        props.append((t, [norm(b)] + [norm(a) for a in BINARY_ALIASES.get(b, [])], "bin"))
    for long_, short in SCRIPTS.items():
        t = "UNI_SC_" + long_.upper()
        tabs[t] = ranges_of(r"\p{Script=%s}" % long_)
        props.append((t, [norm(long_), norm(short)], "sc"))
    with open(path, "w") as f:
        f.write("// generated by tools/gen_unicode_tables.py (regex module %s; unicodedata %s) - do not edit\n" % (regex.__version__, unicodedata.unidata_version))
        for name, rs in tabs.items():
            f.write("static const uint32_t %s[][2] = {\n" % name)
            for i in range(0, len(rs), 6):
                f.write("    " + " ".join("{0x%X,0x%X}," % r for r in rs[i:i + 6]) + "\n")
            f.write("};\n")
        pairs = case_orbits()
        f.write("// simple case folding: (code point, another member of its orbit), sorted\n")
        f.write("static const uint32_t UNI_CASE_PAIRS[][2] = {\n")
        for i in range(0, len(pairs), 6):
            f.write("    " + " ".join("{0x%X,0x%X}," % r for r in pairs[i:i + 6]) + "\n")
        f.write("};\n")
        f.write("struct UniProp { const char* name; const uint32_t (*ranges)[2]; uint32_t n; char kind; };  // kind: g general category, b binary, s script\n")
        f.write("static const UniProp UNI_PROPS[] = {\n")
        for t, names, kind in props:
            for nm in dict.fromkeys(names):
                f.write('    {"%s", %s, sizeof(%s) / sizeof(%s[0]), \'%s\'},\n' % (nm, t, t, t, kind[0]))
        f.write("};\n")


if __name__ == "__main__":
    main(sys.argv[1])
