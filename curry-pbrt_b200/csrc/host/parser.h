// Scene-file front end: same token / block / property model as the reference
// (src/scene_file_parser/lex.rs, src/scene_file_parser/mod.rs) so that a .pbrt file the Rust
// host accepts is accepted here with the same meaning.  Errors are C++ exceptions inside the
// host library and are turned into CPBRT_ERR_PARSE at the C boundary (the reference panics).
#pragma once
#include <cstdint>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace cpbrt_host {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Token (lex.rs:22-31)
// BasicType (mod.rs:439-444), stored compactly: a 10M-triangle trianglemesh is ~10^8 of them.
// Strings (and the file a string token came from -- texture paths are resolved against it,
// mod.rs:300-306) live in a process-wide intern table.
struct Basic {
    enum Kind : uint8_t { Str, Flt, Int } kind = Int;
    float f = 0.f;
    int32_t i = 0;
    uint32_t sid = 0, fid = 0;       // Str: interned string / file ids
    const std::string& str() const;
    const std::string& file() const;
};
using Basics = std::vector<Basic>;   // BasicTypes
uint32_t intern(const std::string& s);
const std::string& interned(uint32_t id);

struct Token {
    enum Kind : uint8_t { Include, Type, BlockBegin, BlockEnd, String, Integer, Float, Array } kind;
    std::string text;          // Type / BlockBegin / BlockEnd / String payload
    int32_t i = 0;
    float f = 0.f;
    Basics array;              // Array payload (elements already reduced to BasicType)
    uint32_t fid = 0;          // interned name of the file the token came from
    uint32_t line = 0, column = 0;
};

// One numeric literal of the scene text (parser.cpp): 0 = not a number, 1 = Integer (*iv), 2 = Float (*fv).
int parse_number(const char* begin, size_t n, int32_t* iv, float* fv);

// parse_lex (lex.rs:202-205).  `text` may be given directly (file is then only a label).
std::vector<Token> lex_file(const std::string& path);
std::vector<Token> lex_text(const std::string& text, const std::string& file_label);

// Property (mod.rs:210-218)
struct Property {
    bool typed = false;
    std::string type_name, name;
    Basics values;
};

// PropertySet (mod.rs:151-209)
struct PropertySet {
    std::deque<Property> props;
    const char* get_name() const;   // first value of the first property, as string (or nullptr)
    const Property* find(const std::string& name) const;   // first TypedValue with that name
    // as_one_basic_types (mod.rs:191-201): pops whole properties from the front until `size`
    // values were gathered.  Returns false when the set runs dry.
    bool take_front(size_t size, Basics* out);
};

// BlockSegment (mod.rs:28-39)
struct Segment {
    bool is_block = false;
    std::string type;                 // block_type or object_type
    bool has_name = false;
    std::string name;                 // block name (ObjectBegin "name")
    std::vector<Segment> children;    // block_segments
    PropertySet object;               // object_value
};

// FileBlock::from_lex (mod.rs:11-19)
std::vector<Segment> segments_from_tokens(std::vector<Token> tokens);

// typed extraction helpers (mod.rs:300-339, 393-437)
bool basics_get_float(const Basics& b, float* out);      // get_float
bool basics_get_integer(const Basics& b, int32_t* out);  // get_integer
bool basics_get_floats(const Basics& b, std::vector<float>* out);   // get_floats
const Basic* basics_get_string(const Basics& b);         // get_string
std::string dir_of(const std::string& path);
std::string join_path(const std::string& dir, const std::string& rel);

}  // namespace cpbrt_host
