// toml_obj.cpp — parsers of the host front end: the TOML subset the scene files use and
// Wavefront OBJ/MTL.  The reference delegates both to crates (toml 0.4 + serde, tobj 0.1.6:
// description.rs:38,156); this is a from-scratch C++ reader of the same surface.
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "host_scene.h"

namespace lr {

// ============================================================================ TOML
const TomlValue* TomlValue::get(const std::string& key) const {
  for (const auto& kv : tab) if (kv.first == key) return &kv.second;
  return nullptr;
}
TomlValue* TomlValue::get(const std::string& key) {
  for (auto& kv : tab) if (kv.first == key) return &kv.second;
  return nullptr;
}
TomlValue& TomlValue::insert(const std::string& key) {
  tab.emplace_back(key, TomlValue{});
  return tab.back().second;
}

namespace {

struct TomlParser {
  const std::string& s;
  size_t pos = 0;
  int line = 1;
  std::string err;

  explicit TomlParser(const std::string& text) : s(text) {}

  bool fail(const std::string& m) { if (err.empty()) err = "TOML line " + std::to_string(line) + ": " + m; return false; }
  bool eof() const { return pos >= s.size(); }
  char peek() const { return eof() ? '\0' : s[pos]; }
  void skip_ws() { while (!eof() && (s[pos] == ' ' || s[pos] == '\t')) pos++; }
  void skip_comment() { if (peek() == '#') while (!eof() && s[pos] != '\n') pos++; }
  // whitespace, comments and newlines (inside arrays / between statements)
  void skip_blank() {
    while (!eof()) {
      const char c = s[pos];
      if (c == ' ' || c == '\t' || c == '\r') pos++;
      else if (c == '\n') { pos++; line++; }
      else if (c == '#') skip_comment();
      else break;
    }
  }
  bool end_of_line() {
    skip_ws(); skip_comment();
    if (eof()) return true;
    if (s[pos] == '\r') pos++;
    if (eof()) return true;
    if (s[pos] == '\n') { pos++; line++; return true; }
    return fail("unexpected trailing characters");
  }

  bool parse_key(std::string& key) {
    skip_ws();
    key.clear();
    if (peek() == '"' || peek() == '\'') return parse_string(key);
    while (!eof() && (std::isalnum((unsigned char)s[pos]) || s[pos] == '_' || s[pos] == '-')) key.push_back(s[pos++]);
    if (key.empty()) return fail("expected a key");
    return true;
  }
  bool parse_key_path(std::vector<std::string>& path) {
    path.clear();
    while (true) {
      std::string k;
      if (!parse_key(k)) return false;
      path.push_back(k);
      skip_ws();
      if (peek() == '.') { pos++; continue; }
      return true;
    }
  }
  bool parse_string(std::string& out) {
    const char q = s[pos++];
    out.clear();
    while (!eof() && s[pos] != q) {
      char c = s[pos++];
      if (c == '\n') return fail("newline in string");
      if (q == '"' && c == '\\') {
        if (eof()) return fail("bad escape");
        const char e = s[pos++];
        switch (e) {
          case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break;
          case '\\': c = '\\'; break; case '"': c = '"'; break; case 'b': c = '\b'; break; case 'f': c = '\f'; break;
          default: return fail("unsupported escape sequence");
        }
      }
      out.push_back(c);
    }
    if (eof()) return fail("unterminated string");
    pos++;
    return true;
  }
  bool parse_number_or_word(TomlValue& v) {
    const size_t start = pos;
    while (!eof() && (std::isalnum((unsigned char)s[pos]) || s[pos] == '+' || s[pos] == '-' || s[pos] == '.' || s[pos] == '_')) pos++;
    std::string tok = s.substr(start, pos - start);
    if (tok.empty()) return fail("expected a value");
    if (tok == "true" || tok == "false") { v.kind = TomlValue::BOOL; v.b = tok == "true"; return true; }
    std::string clean;
    for (char c : tok) if (c != '_') clean.push_back(c);
    const std::string body = (clean[0] == '+' || clean[0] == '-') ? clean.substr(1) : clean;
    if (body == "inf" || body == "nan") {
      v.kind = TomlValue::FLOAT;
      v.f = body == "inf" ? HUGE_VAL : NAN;
      if (clean[0] == '-') v.f = -v.f;
      return true;
    }
    const bool is_float = clean.find_first_of(".eE") != std::string::npos;
    char* end = nullptr;
    errno = 0;
    if (is_float) {
      v.kind = TomlValue::FLOAT;
      v.f = std::strtod(clean.c_str(), &end);
    } else {
      v.kind = TomlValue::INT;
      v.i = std::strtoll(clean.c_str(), &end, 10);
    }
    if (!end || *end != '\0') return fail("malformed number `" + tok + "`");
    return true;
  }
  bool parse_value(TomlValue& v) {
    skip_ws();
    const char c = peek();
    if (c == '"' || c == '\'') { v.kind = TomlValue::STRING; return parse_string(v.s); }
    if (c == '[') {
      pos++;
      v.kind = TomlValue::ARRAY;
      while (true) {
        skip_blank();
        if (peek() == ']') { pos++; return true; }
        TomlValue e;
        if (!parse_value(e)) return false;
        v.arr.push_back(std::move(e));
        skip_blank();
        if (peek() == ',') { pos++; continue; }
        if (peek() == ']') { pos++; return true; }
        return fail("expected `,` or `]` in array");
      }
    }
    if (c == '{') {
      pos++;
      v.kind = TomlValue::TABLE;
      skip_ws();
      if (peek() == '}') { pos++; return true; }
      while (true) {
        std::string k;
        if (!parse_key(k)) return false;
        skip_ws();
        if (peek() != '=') return fail("expected `=` in inline table");
        pos++;
        if (v.get(k)) return fail("duplicate key `" + k + "`");
        TomlValue e;
        if (!parse_value(e)) return false;
        v.insert(k) = std::move(e);
        skip_ws();
        if (peek() == ',') { pos++; continue; }
        if (peek() == '}') { pos++; return true; }
        return fail("expected `,` or `}` in inline table");
      }
    }
    return parse_number_or_word(v);
  }

  // walks `path` from the root; arrays of tables resolve to their last element
  TomlValue* descend(TomlValue& root, const std::vector<std::string>& path, size_t n) {
    TomlValue* cur = &root;
    for (size_t i = 0; i < n; i++) {
      TomlValue* next = cur->get(path[i]);
      if (!next) { next = &cur->insert(path[i]); next->kind = TomlValue::TABLE; }
      if (next->kind == TomlValue::ARRAY) {
        if (next->arr.empty() || next->arr.back().kind != TomlValue::TABLE) { fail("`" + path[i] + "` is not a table"); return nullptr; }
        next = &next->arr.back();
      } else if (next->kind != TomlValue::TABLE) { fail("`" + path[i] + "` is not a table"); return nullptr; }
      cur = next;
    }
    return cur;
  }

  bool parse(TomlValue& root) {
    root = TomlValue{};
    root.kind = TomlValue::TABLE;
    TomlValue* current = &root;
    while (true) {
      skip_blank();
      if (eof()) return true;
      if (peek() == '[') {
        pos++;
        const bool is_array = peek() == '[';
        if (is_array) pos++;
        std::vector<std::string> path;
        if (!parse_key_path(path)) return false;
        skip_ws();
        if (peek() != ']') return fail("expected `]`");
        pos++;
        if (is_array) { if (peek() != ']') return fail("expected `]]`"); pos++; }
        TomlValue* parent = descend(root, path, path.size() - 1);
        if (!parent) return false;
        const std::string& last = path.back();
        TomlValue* slot = parent->get(last);
        if (is_array) {
          if (!slot) { slot = &parent->insert(last); slot->kind = TomlValue::ARRAY; }
          if (slot->kind != TomlValue::ARRAY) return fail("`" + last + "` is not an array of tables");
          slot->arr.emplace_back();
          slot->arr.back().kind = TomlValue::TABLE;
          current = &slot->arr.back();
        } else {
          if (!slot) { slot = &parent->insert(last); slot->kind = TomlValue::TABLE; }
          if (slot->kind != TomlValue::TABLE) return fail("`" + last + "` is not a table");
          current = slot;
        }
        if (!end_of_line()) return false;
        continue;
      }
      std::vector<std::string> path;
      if (!parse_key_path(path)) return false;
      skip_ws();
      if (peek() != '=') return fail("expected `=` after key");
      pos++;
      TomlValue* target = current;
      if (path.size() > 1) {
        // dotted key: relative to the current table
        for (size_t i = 0; i + 1 < path.size(); i++) {
          TomlValue* next = target->get(path[i]);
          if (!next) { next = &target->insert(path[i]); next->kind = TomlValue::TABLE; }
          if (next->kind != TomlValue::TABLE) return fail("`" + path[i] + "` is not a table");
          target = next;
        }
      }
      if (target->get(path.back())) return fail("duplicate key `" + path.back() + "`");
      TomlValue v;
      if (!parse_value(v)) return false;
      target->insert(path.back()) = std::move(v);
      if (!end_of_line()) return false;
    }
  }
};

}  // namespace

int toml_parse(const std::string& text, TomlValue& root, std::string& err) {
  TomlParser p(text);
  if (!p.parse(root)) { err = p.err; return LR_ERR_PARSE; }
  return LR_OK;
}

// ============================================================================ OBJ / MTL
namespace {

std::string dir_of(const std::string& path) {
  const size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string() : path.substr(0, k + 1);
}

bool read_file(const std::string& path, std::string& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

inline const char* skip_sp(const char* p) { while (*p == ' ' || *p == '\t') p++; return p; }

void load_mtl(const std::string& path, std::vector<ObjMaterial>& mats) {
  std::string text;
  if (!read_file(path, text)) return;     // tobj reports an error; the scene then fails only if a material is needed
  std::istringstream in(text);
  std::string ln;
  while (std::getline(in, ln)) {
    if (!ln.empty() && ln.back() == '\r') ln.pop_back();
    const char* p = skip_sp(ln.c_str());
    if (std::strncmp(p, "newmtl", 6) == 0 && (p[6] == ' ' || p[6] == '\t')) {
      ObjMaterial m;
      m.name = skip_sp(p + 6);
      while (!m.name.empty() && (m.name.back() == ' ' || m.name.back() == '\t')) m.name.pop_back();
      mats.push_back(m);
    } else if (std::strncmp(p, "Kd", 2) == 0 && (p[2] == ' ' || p[2] == '\t') && !mats.empty()) {
      // a malformed Kd (fewer than 3 numbers) leaves the components it does not define at tobj's default of 0
      char* e = nullptr;
      const char* q = p + 2;
      float kd[3] = {0.0f, 0.0f, 0.0f};
      for (int i = 0; i < 3; i++) { kd[i] = std::strtof(q, &e); if (e == q) { kd[i] = 0.0f; break; } q = e; }
      for (int i = 0; i < 3; i++) mats.back().diffuse[i] = kd[i];
    }
  }
}

}  // namespace

// Semantics of tobj::load_obj as used by description.rs:164-197: a model per `o`/`g` group (and per
// `usemtl` change inside a group), polygons fan-triangulated, `Kd` -> diffuse.
int load_obj(const std::string& path, ObjFile& out) {
  std::string text;
  if (!read_file(path, text)) return fail(LR_ERR_IO, "File `" + path + "` is not found.");
  out = ObjFile{};
  std::vector<float> pos;
  ObjModel cur;
  cur.name = "unnamed_object";
  auto flush = [&]() {
    if (!cur.positions.empty()) out.models.push_back(cur);
    cur.positions.clear();
  };
  const char* p = text.c_str();
  const char* end = p + text.size();
  std::vector<long> corner;
  int lineno = 0;
  while (p < end) {
    lineno++;
    const char* eol = (const char*)std::memchr(p, '\n', end - p);
    if (!eol) eol = end;
    const char* q = skip_sp(p);
    if (q[0] == 'v' && (q[1] == ' ' || q[1] == '\t')) {
      // strtof skips '\n': a short line ("v 1 2") must not borrow the first number of the next line
      char* e = nullptr;
      q += 1;
      for (int i = 0; i < 3; i++) {
        q = skip_sp(q);
        const float x = q < eol ? std::strtof(q, &e) : 0.0f;
        if (q >= eol || e == q || e > eol) return fail(LR_ERR_PARSE, path + ":" + std::to_string(lineno) + ": malformed vertex");
        pos.push_back(x);
        q = e;
      }
    } else if (q[0] == 'f' && (q[1] == ' ' || q[1] == '\t')) {
      corner.clear();
      q += 1;
      while (true) {
        q = skip_sp(q);
        if (q >= eol || *q == '\r' || *q == '\n' || *q == '#') break;
        char* e = nullptr;
        long idx = std::strtol(q, &e, 10);
        if (e == q) return fail(LR_ERR_PARSE, path + ":" + std::to_string(lineno) + ": malformed face");
        const long nv = (long)(pos.size() / 3);
        if (idx < 0) idx = nv + idx; else idx -= 1;
        if (idx < 0 || idx >= nv) return fail(LR_ERR_PARSE, path + ":" + std::to_string(lineno) + ": face index out of range");
        corner.push_back(idx);
        q = e;
        while (q < eol && *q != ' ' && *q != '\t' && *q != '\r') q++;   // skip /vt/vn
      }
      if (corner.size() < 3) return fail(LR_ERR_PARSE, path + ":" + std::to_string(lineno) + ": face with fewer than 3 vertices");
      for (size_t k = 1; k + 1 < corner.size(); k++) {
        const long tri[3] = {corner[0], corner[k], corner[k + 1]};
        for (long c : tri) for (int a = 0; a < 3; a++) cur.positions.push_back(pos[3 * c + a]);
      }
    } else if ((q[0] == 'o' || q[0] == 'g') && (q[1] == ' ' || q[1] == '\t')) {
      flush();
      std::string name(skip_sp(q + 1), eol);
      while (!name.empty() && (name.back() == '\r' || name.back() == ' ')) name.pop_back();
      cur.name = name.empty() ? "unnamed_object" : name;
    } else if (std::strncmp(q, "usemtl", 6) == 0 && (q[6] == ' ' || q[6] == '\t')) {
      // tobj 0.1.6 looks up the first whitespace-delimited word after the keyword (`words.next()`)
      const char* w0 = skip_sp(q + 6);
      const char* w1 = w0;
      while (w1 < eol && *w1 != ' ' && *w1 != '\t' && *w1 != '\r') w1++;
      std::string name(w0, w1);
      int id = -1;
      for (size_t i = 0; i < out.materials.size(); i++) if (out.materials[i].name == name) id = (int)i;
      if (!cur.positions.empty() && id != cur.material_id) flush();
      cur.material_id = id;
    } else if (std::strncmp(q, "mtllib", 6) == 0 && (q[6] == ' ' || q[6] == '\t')) {
      std::string name(skip_sp(q + 6), eol);
      while (!name.empty() && (name.back() == '\r' || name.back() == ' ')) name.pop_back();
      load_mtl(dir_of(path) + name, out.materials);
    }
    p = eol + 1;
  }
  flush();
  return LR_OK;
}

}  // namespace lr
