// Lexer + block/property parser with the reference's semantics
// (src/scene_file_parser/lex.rs:81-205, src/scene_file_parser/mod.rs:11-145, 219-296).
#include "parser.h"

#include <cerrno>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>

namespace cpbrt_host {

std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of('/');
    if (p == std::string::npos) return std::string();
    if (p == 0) return "/";
    return path.substr(0, p);
}
std::string join_path(const std::string& dir, const std::string& rel) {
    if (!rel.empty() && rel[0] == '/') return rel;   // Path::join with an absolute path replaces
    if (dir.empty()) return rel;
    if (dir.back() == '/') return dir + rel;
    return dir + "/" + rel;
}

namespace {
struct InternTable {
    std::deque<std::string> v;   // deque: references handed out stay valid as it grows
    std::map<std::string, uint32_t> idx;
    std::mutex mu;
};
InternTable& table() { static InternTable t; return t; }
}  // namespace
uint32_t intern(const std::string& s) {
    InternTable& t = table();
    std::lock_guard<std::mutex> g(t.mu);
    auto it = t.idx.find(s);
    if (it != t.idx.end()) return it->second;
    uint32_t id = (uint32_t)t.v.size();
    t.v.push_back(s);
    t.idx.emplace(s, id);
    return id;
}
const std::string& interned(uint32_t id) {
    InternTable& t = table();
    std::lock_guard<std::mutex> g(t.mu);
    return t.v[id];
}
const std::string& Basic::str() const { return interned(sid); }
const std::string& Basic::file() const { return interned(fid); }

static std::string read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw ParseError("cannot open scene file '" + path + "'");
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// Rust `str::parse::<i32>`: optional sign, at least one digit, nothing else, no overflow.
static bool parse_i32(const std::string& w, int32_t* out) {
    size_t i = 0;
    if (w.empty()) return false;
    if (w[0] == '+' || w[0] == '-') i = 1;
    if (i == w.size()) return false;
    long long v = 0;
    for (size_t k = i; k < w.size(); ++k) {
        if (w[k] < '0' || w[k] > '9') return false;
        v = v * 10 + (w[k] - '0');
        if (v > (long long)INT_MAX + 1) return false;
    }
    if (w[0] == '-') v = -v;
    if (v > INT_MAX || v < INT_MIN) return false;
    *out = (int32_t)v;
    return true;
}

// Rust `str::parse::<f32>` grammar (decimal literal with optional exponent, or inf/nan),
// value correctly rounded -- glibc strtof rounds correctly as well.
static bool parse_f32(const std::string& w, float* out) {
    size_t i = 0, n = w.size();
    if (n == 0) return false;
    if (w[i] == '+' || w[i] == '-') ++i;
    std::string rest = w.substr(i);
    for (auto& c : rest) c = (char)tolower((unsigned char)c);
    bool special = (rest == "inf" || rest == "infinity" || rest == "nan");
    if (!special) {
        size_t k = i, digits = 0;
        while (k < n && isdigit((unsigned char)w[k])) { ++k; ++digits; }
        if (k < n && w[k] == '.') {
            ++k;
            while (k < n && isdigit((unsigned char)w[k])) { ++k; ++digits; }
        }
        if (digits == 0) return false;
        if (k < n && (w[k] == 'e' || w[k] == 'E')) {
            ++k;
            if (k < n && (w[k] == '+' || w[k] == '-')) ++k;
            size_t ed = 0;
            while (k < n && isdigit((unsigned char)w[k])) { ++k; ++ed; }
            if (ed == 0) return false;
        }
        if (k != n) return false;
    }
    char* end = nullptr;
    *out = strtof(w.c_str(), &end);
    return end && *end == 0;
}

// Numbers are almost all of a big scene file (a million-triangle `trianglemesh` literal is 4.5 million of them), so
// the common shape -- [-]digits[.digits][e[+-]digits] -- is read straight from the text, without a std::string, a
// Token or strtof per number:
//   * all digits, fits i32                       -> Integer, as parse_i32
//   * mantissa <= 2^53 and |decimal exponent| <= 22 -> one IEEE double operation m * 10^e or m / 10^e (both operands
//     exact, so the double is the correctly rounded value), then double -> float.  The two roundings can only differ
//     from one correct rounding to f32 when the double lands EXACTLY on the midpoint of two neighbouring floats (its
//     low 29 mantissa bits are 0x10000000): those, subnormal / overflowing magnitudes and everything else that is not
//     the common shape take the strtof path (parse_f32) as before.
// Same values bit for bit (tests/lexer_number_check.cpp compares millions of literals with strtof); returns
// 0 = not a number, 1 = Integer, 2 = Float.
int parse_number(const char* b, size_t n, int32_t* iv, float* fv) {
    static const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    size_t k = 0;
    bool neg = false;
    if (k < n && (b[k] == '-' || b[k] == '+')) { neg = b[k] == '-'; ++k; }
    uint64_t m = 0; int sig = 0, int_digits = 0, frac_digits = 0; int dropped = 0; bool inexact = false;
    auto take = [&](char c) {
        if (sig < 19) { m = m * 10 + (uint64_t)(c - '0'); if (m) ++sig; }
        else { ++dropped; if (c != '0') inexact = true; }
    };
    while (k < n && b[k] >= '0' && b[k] <= '9') { take(b[k]); ++k; ++int_digits; }
    bool is_int = int_digits > 0 && k == n;
    if (is_int) {
        // parse_i32: no overflow
        if (sig <= 10 && dropped == 0 && m <= (neg ? 2147483648ull : 2147483647ull)) { *iv = neg ? (int32_t)(-(int64_t)m) : (int32_t)m; return 1; }
    }
    int dropped_int = dropped;
    if (k < n && b[k] == '.') { ++k; while (k < n && b[k] >= '0' && b[k] <= '9') { take(b[k]); ++k; ++frac_digits; } }
    if (int_digits + frac_digits == 0) goto slow;
    {
        int e10 = 0;
        if (k < n && (b[k] == 'e' || b[k] == 'E')) {
            ++k;
            bool eneg = false;
            if (k < n && (b[k] == '+' || b[k] == '-')) { eneg = b[k] == '-'; ++k; }
            int ed = 0, ev = 0;
            while (k < n && b[k] >= '0' && b[k] <= '9') { if (ev < 100000) ev = ev * 10 + (b[k] - '0'); ++k; ++ed; }
            if (ed == 0) return 0;
            e10 = eneg ? -ev : ev;
        }
        if (k != n) goto slow;
        // digits beyond the 19th were dropped: those before the point scale the value, any non-zero one makes it inexact
        const int frac_kept = frac_digits - (dropped - dropped_int);
        e10 += dropped_int - frac_kept;
        if (inexact) goto slow;
        if (m == 0) { *fv = neg ? -0.f : 0.f; return 2; }
        if (m > (1ull << 53) || e10 < -22 || e10 > 22) goto slow;
        const double d = e10 >= 0 ? (double)m * p10[e10] : (double)m / p10[-e10];
        if (!(d >= 1e-30 && d <= 1e30)) goto slow;
        uint64_t bits; memcpy(&bits, &d, 8);
        const uint32_t low = (uint32_t)(bits & 0x1FFFFFFFu);
        if (low >= 0x0FFFFFFFu && low <= 0x10000001u) goto slow;
        const float f = (float)d;
        *fv = neg ? -f : f;
        return 2;
    }
slow: {
        const std::string w(b, n);
        if (parse_i32(w, iv)) return 1;
        if (parse_f32(w, fv)) return 2;
        return 0;
    }
}

namespace {
struct Lexer {
    const std::string& s;
    std::string file;
    uint32_t fid;
    std::vector<Token> toks;
    size_t line = 1, column = 1, index = 0;

    Lexer(const std::string& text, const std::string& f) : s(text), file(f), fid(intern(f)) {}

    [[noreturn]] void fail(const std::string& msg) const {
        std::ostringstream o;
        o << "file [" << file << "] line " << line << " column " << column << ": " << msg;
        throw ParseError(o.str());
    }
    // next_equal (lex.rs:51-54): position of the next `c` strictly after index
    bool next_equal(char c, size_t* end) const {
        size_t p = s.find(c, index + 1);
        if (p == std::string::npos) return false;
        *end = p;
        return true;
    }
    // word (lex.rs:55-61)
    bool word(size_t* end) const {
        const char* d = s.data();
        for (size_t p = index + 1, n = s.size(); p < n; ++p) {
            const char ch = d[p];
            if (ch == ' ' || ch == '\n' || ch == '\t' || ch == ']' || ch == '\r') { *end = p; return true; }
        }
        return false;
    }
    // Between '[' and ']' the reference pushes ordinary tokens and drains them into
    // Token::Array at ']' (lex.rs:127-155); we append the reduced BasicType directly, which is
    // the same thing for every array BasicTypes::from_lex would accept.
    bool in_array = false;
    Basics cur_array;
    size_t arr_line = 0, arr_col = 0;
    void push(Token t) {
        if (in_array) {
            Basic b;
            if (t.kind == Token::String) { b.kind = Basic::Str; b.sid = intern(t.text); b.fid = fid; }
            else if (t.kind == Token::Integer) { b.kind = Basic::Int; b.i = t.i; }
            else if (t.kind == Token::Float) { b.kind = Basic::Flt; b.f = t.f; }
            else fail("Unexpected basic type");
            cur_array.push_back(b);
            return;
        }
        t.fid = fid; t.line = (uint32_t)line; t.column = (uint32_t)column;
        toks.push_back(std::move(t));
    }
    std::vector<Token> run() {
        while (index != s.size()) {
            char c = s[index];
            if (c == '\n') { ++index; ++line; column = 1; }
            else if (c == ' ' || c == '\t' || c == '\r') { ++index; ++column; }
            else if (c == '"') {
                size_t end;
                if (!next_equal('"', &end)) fail("Unpaired '\"'");
                std::string content = s.substr(index + 1, end - index - 1);
                bool is_include = false;
                if (!in_array && !toks.empty() && toks.back().kind == Token::Include) {
                    toks.pop_back();
                    std::vector<Token> inc = lex_file(join_path(dir_of(file), content));
                    for (auto& t : inc) toks.push_back(std::move(t));
                    is_include = true;
                }
                if (!is_include) { Token t; t.kind = Token::String; t.text = content; push(std::move(t)); }
                column += end - index + 1;
                index = end + 1;
            }
            else if (c == '#') {
                size_t end;
                if (!next_equal('\n', &end)) break;
                column += end - index;
                index = end;
            }
            else if (c == '[') {
                if (in_array) fail("nested '['");
                in_array = true; cur_array.clear(); arr_line = line; arr_col = column;
                // one allocation for the array instead of a dozen doublings (a mesh literal holds millions of
                // elements): elements are separated by white space, so the separators up to ']' bound their number
                if (const void* close = memchr(s.data() + index, ']', s.size() - index)) {
                    const char* b = s.data() + index + 1; const char* e = (const char*)close;
                    if (e - b > 4096) {
                        size_t seps = 0;
                        for (const char* q = b; q < e; ++q) seps += (*q == ' ' || *q == '\n' || *q == '\t' || *q == '\r') ? 1u : 0u;
                        cur_array.reserve(seps + 1);
                    }
                }
                ++column; ++index;
            }
            else if (c == ']') {
                if (!in_array) fail("Unpaired ']'");
                in_array = false;
                Token t; t.kind = Token::Array;
                t.array = std::move(cur_array);
                cur_array = Basics();
                t.fid = fid; t.line = (uint32_t)arr_line; t.column = (uint32_t)arr_col;
                toks.push_back(std::move(t));
                ++column; ++index;
            }
            else if (c >= 'A' && c <= 'Z') {
                size_t end;
                if (!word(&end)) break;
                std::string w = s.substr(index, end - index);
                Token t;
                auto ends_with = [&](const char* suf) {
                    size_t n = strlen(suf);
                    return w.size() >= n && w.compare(w.size() - n, n, suf) == 0;
                };
                if (w == "Include") t.kind = Token::Include;
                else if (ends_with("Begin")) { t.kind = Token::BlockBegin; t.text = w.substr(0, w.size() - 5); }
                else if (ends_with("End")) { t.kind = Token::BlockEnd; t.text = w.substr(0, w.size() - 3); }
                else { t.kind = Token::Type; t.text = w; }
                push(std::move(t));
                column += end - index;
                index = end;
            }
            else if (c == '-' || (c >= '0' && c <= '9')) {
                size_t end;
                if (!word(&end)) break;
                int32_t iv = 0; float fv = 0.f;
                const int kind = parse_number(s.data() + index, end - index, &iv, &fv);
                if (kind == 0) fail("Cant parse number \"" + s.substr(index, end - index) + "\"");
                if (in_array) {      // the bulk: no Token in between
                    Basic b; b.kind = kind == 1 ? Basic::Int : Basic::Flt; b.i = iv; b.f = fv;
                    cur_array.push_back(b);
                } else {
                    Token t; t.kind = kind == 1 ? Token::Integer : Token::Float; t.i = iv; t.f = fv;
                    push(std::move(t));
                }
                column += end - index;
                index = end;
            }
            else {
                fail(std::string("Cant parse char '") + c);
            }
        }
        return std::move(toks);
    }
};
}  // namespace

std::vector<Token> lex_text(const std::string& text, const std::string& file_label) {
    Lexer lx(text, file_label);
    return lx.run();
}
// Include nests through lex_file (lex.rs:98-105).  A file that includes itself, directly or through
// others, recurses until the stack is gone in the reference; here it is a parse error.
static thread_local int g_include_depth = 0;
std::vector<Token> lex_file(const std::string& path) {
    struct Depth { Depth() { ++g_include_depth; } ~Depth() { --g_include_depth; } } depth;
    if (g_include_depth > 32) throw ParseError("Include nested deeper than 32 files (a file including itself?): " + path);
    std::string text = read_file(path);
    return lex_text(text, path);
}

// ---- block / property layer -----------------------------------------------------------------

static Basic basic_from_token(const Token& t) {   // BasicType::from_lex (mod.rs:445-457)
    Basic b;
    switch (t.kind) {
        case Token::String: b.kind = Basic::Str; b.sid = intern(t.text); b.fid = t.fid; break;
        case Token::Integer: b.kind = Basic::Int; b.i = t.i; break;
        case Token::Float: b.kind = Basic::Flt; b.f = t.f; break;
        default: throw ParseError("Unexpected basic type in " + interned(t.fid) + " line " + std::to_string(t.line));
    }
    return b;
}

static Basics basics_from_tokens(std::deque<Token>& toks) {   // BasicTypes::from_lex (mod.rs:281-296)
    if (toks.empty()) throw ParseError("unexpected end of file while reading a value");
    Token t = std::move(toks.front());
    toks.pop_front();
    Basics r;
    if (t.kind == Token::Array) r = std::move(t.array);
    else r.push_back(basic_from_token(t));
    return r;
}

static Property property_from_tokens(std::deque<Token>& toks) {   // Property::from_lex (mod.rs:240-276)
    Property p;
    const Token& t = toks.front();
    if (t.kind == Token::String) {
        std::vector<std::string> words;
        std::istringstream ss(t.text);
        for (std::string w; ss >> w;) words.push_back(w);
        if (words.size() == 2) {
            static const char* typed[] = {"string", "float", "spectrum", "texture", "integer", "rgb", "point"};
            for (const char* ty : typed)
                if (words[0] == ty) {
                    p.typed = true; p.type_name = words[0]; p.name = words[1];
                    toks.pop_front();
                    p.values = basics_from_tokens(toks);
                    return p;
                }
        }
        p.values.push_back(basic_from_token(t));
        toks.pop_front();
        return p;
    }
    p.values = basics_from_tokens(toks);
    return p;
}

// Blocks nest by recursion here, in the scene builder and in the reference (BlockSegment::from_lex,
// SceneParseStack::parse); a file with hundreds of thousands of unclosed *Begin lines would take the
// stack with it, so the depth is bounded and exceeding it is a parse error.
static thread_local int g_block_depth = 0;
static const int kMaxBlockDepth = 512;
static Segment segment_from_tokens(std::deque<Token>& toks) {   // BlockSegment::from_lex (mod.rs:41-100)
    struct Depth { Depth() { ++g_block_depth; } ~Depth() { --g_block_depth; } } depth;
    if (g_block_depth > kMaxBlockDepth) throw ParseError("blocks nested deeper than " + std::to_string(kMaxBlockDepth));
    Token t = std::move(toks.front());
    toks.pop_front();
    Segment seg;
    if (t.kind == Token::Type) {
        seg.type = t.text;
        while (!toks.empty()) {
            Token::Kind k = toks.front().kind;
            if (k == Token::Type || k == Token::BlockBegin || k == Token::BlockEnd) break;
            seg.object.props.push_back(property_from_tokens(toks));
        }
        return seg;
    }
    if (t.kind == Token::BlockBegin) {
        seg.is_block = true;
        seg.type = t.text;
        if (!toks.empty() && toks.front().kind == Token::String) {
            seg.has_name = true; seg.name = toks.front().text; toks.pop_front();
        }
        while (!toks.empty()) {
            const Token& n = toks.front();
            if (n.kind == Token::BlockEnd) {
                if (n.text == seg.type) { toks.pop_front(); break; }
                throw ParseError("file [" + interned(n.fid) + "] line " + std::to_string(n.line) + ": Unpaired block end");
            }
            seg.children.push_back(segment_from_tokens(toks));
        }
        return seg;
    }
    throw ParseError("file [" + interned(t.fid) + "] line " + std::to_string(t.line) + ": Unexpected block segments");
}

std::vector<Segment> segments_from_tokens(std::vector<Token> tokens) {
    std::deque<Token> q(std::make_move_iterator(tokens.begin()), std::make_move_iterator(tokens.end()));
    std::vector<Segment> r;
    while (!q.empty()) r.push_back(segment_from_tokens(q));
    return r;
}

const char* PropertySet::get_name() const {
    if (props.empty() || props[0].values.empty()) return nullptr;
    const Basic& b = props[0].values.front();
    return b.kind == Basic::Str ? b.str().c_str() : nullptr;
}
const Property* PropertySet::find(const std::string& name) const {
    for (auto& p : props) if (p.typed && p.name == name) return &p;
    return nullptr;
}
bool PropertySet::take_front(size_t size, Basics* out) {
    out->clear();
    while (size > 0) {
        if (props.empty()) return false;
        Basics v = std::move(props.front().values);
        props.pop_front();
        size_t take = v.size() < size ? v.size() : size;
        for (size_t i = 0; i < take; ++i) out->push_back(v[i]);
        if (v.size() > size) throw ParseError("value list longer than the directive consumes");   // usize underflow in the reference
        size -= v.size();
    }
    return true;
}

bool basics_get_float(const Basics& b, float* out) {
    if (b.empty()) return false;
    if (b[0].kind == Basic::Flt) { *out = b[0].f; return true; }
    if (b[0].kind == Basic::Int) { *out = (float)b[0].i; return true; }
    return false;
}
bool basics_get_integer(const Basics& b, int32_t* out) {
    if (b.empty() || b[0].kind != Basic::Int) return false;
    *out = b[0].i;
    return true;
}
bool basics_get_floats(const Basics& b, std::vector<float>* out) {
    out->clear();
    for (auto& e : b) {
        if (e.kind == Basic::Flt) out->push_back(e.f);
        else if (e.kind == Basic::Int) out->push_back((float)e.i);
        else return false;
    }
    return true;
}
const Basic* basics_get_string(const Basics& b) {
    if (b.empty() || b[0].kind != Basic::Str) return nullptr;
    return &b[0];
}

}  // namespace cpbrt_host
