// yaml_lite.hpp — the YAML subset the LidarOdometry pipeline files use (pipelines/lidar3d-default.yaml,
// pipelines/lidar3d-ndt.yaml): block mappings and sequences by indentation, single-line flow sequences/maps,
// quoted and plain scalars, '#' comments, '~' null; plus the mola_yaml pre-processor forms that appear in them:
//   ${VAR|default}   environment variable with default (docs/mola_lo_pipelines.rst:25-30)
//   $f{expr}         formula, kept as text and evaluated when the owning object is created
// Not a general YAML parser (no anchors, multi-line scalars, multi-document streams).
#pragma once
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mlo_host {

struct YamlNode {
  enum Type { Null, Scalar, Map, Seq } type = Null;
  std::string scalar;
  std::vector<std::pair<std::string, YamlNode>> map;
  std::vector<YamlNode> seq;

  bool isNull() const { return type == Null; }
  bool isMap() const { return type == Map; }
  bool isSeq() const { return type == Seq; }
  bool has(const std::string& k) const {
    for (auto& kv : map)
      if (kv.first == k) return true;
    return false;
  }
  const YamlNode& operator[](const std::string& k) const {
    for (auto& kv : map)
      if (kv.first == k) return kv.second;
    static const YamlNode null_node;
    return null_node;
  }
  const YamlNode& at(const std::string& k) const {
    if (!has(k)) throw std::runtime_error("yaml: missing required key '" + k + "'");
    return (*this)[k];
  }
  // scalar text with a formula wrapper $f{...} removed
  std::string str() const {
    std::string s = scalar;
    if (s.size() > 4 && s.compare(0, 3, "$f{") == 0 && s.back() == '}') s = s.substr(3, s.size() - 4);
    return s;
  }
  std::string str_or(const std::string& d) const { return type == Scalar ? str() : d; }
  double num(double d = 0.0) const {
    if (type != Scalar) return d;
    char* e = nullptr;
    const double v = std::strtod(scalar.c_str(), &e);
    while (e && (*e == ' ' || *e == '\t')) e++;
    // the whole scalar must be the number: '2000*K' or '30 deg' is a formula / a typo, not 2000 or 30
    if (e == scalar.c_str() || (e && *e != 0)) throw std::runtime_error("yaml: '" + scalar + "' is not a number");
    return v;
  }
  bool boolean(bool d = false) const {
    if (type != Scalar) return d;
    return scalar == "true" || scalar == "True" || scalar == "1" || scalar == "yes";
  }
};

namespace detail {

inline std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) a++;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) b--;
  return s.substr(a, b - a);
}

// ${VAR|default} -> getenv(VAR) or default, innermost first (defaults may nest further ${} / $f{} forms)
inline std::string expand_env(std::string s) {
  for (int guard = 0; guard < 1000; guard++) {
    const size_t a = s.rfind("${");
    if (a == std::string::npos) return s;
    const size_t b = s.find('}', a);
    if (b == std::string::npos) throw std::runtime_error("yaml: unterminated ${...}");
    // the matching brace: count nested braces opened after a (e.g. $f{...} inside the default)
    size_t depth = 0, end = a + 2;
    for (; end < s.size(); end++) {
      if (s[end] == '{') depth++;
      else if (s[end] == '}') {
        if (depth == 0) break;
        depth--;
      }
    }
    const std::string inner = s.substr(a + 2, end - a - 2);
    const size_t bar = inner.find('|');
    const std::string name = inner.substr(0, bar);
    std::string val;
    if (const char* e = std::getenv(name.c_str())) val = e;
    else if (bar != std::string::npos) val = inner.substr(bar + 1);
    else if (name == "CURRENT_YAML_FILE_PATH") val = ".";
    else throw std::runtime_error("yaml: environment variable '" + name + "' is not set and has no default");
    s = s.substr(0, a) + val + s.substr(end + 1);
  }
  throw std::runtime_error("yaml: ${} expansion does not terminate");
}

inline std::string strip_comment(const std::string& line) {
  char q = 0;
  for (size_t i = 0; i < line.size(); i++) {
    const char c = line[i];
    if (q) {
      if (c == q) q = 0;
    } else if (c == '\'' || c == '"') {
      q = c;
    } else if (c == '#' && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) {
      return line.substr(0, i);
    }
  }
  return line;
}

inline std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '\'' && s.back() == '\'') || (s.front() == '"' && s.back() == '"')))
    return s.substr(1, s.size() - 2);
  return s;
}

// split a flow collection body at top-level commas
inline std::vector<std::string> split_flow(const std::string& body) {
  std::vector<std::string> out;
  int depth = 0;
  char q = 0;
  std::string cur;
  for (char c : body) {
    if (q) {
      cur += c;
      if (c == q) q = 0;
      continue;
    }
    if (c == '\'' || c == '"') q = c;
    if (c == '[' || c == '{' || c == '(') depth++;
    if (c == ']' || c == '}' || c == ')') depth--;
    if (c == ',' && depth == 0) {
      out.push_back(trim(cur));
      cur.clear();
    } else {
      cur += c;
    }
  }
  if (!trim(cur).empty()) out.push_back(trim(cur));
  return out;
}

inline size_t find_key_colon(const std::string& s) {
  char q = 0;
  int depth = 0;
  for (size_t i = 0; i < s.size(); i++) {
    const char c = s[i];
    if (q) {
      if (c == q) q = 0;
      continue;
    }
    if (c == '\'' || c == '"') q = c;
    else if (c == '[' || c == '{' || c == '(') depth++;
    else if (c == ']' || c == '}' || c == ')') depth--;
    else if (c == ':' && depth == 0 && (i + 1 == s.size() || s[i + 1] == ' ')) return i;
  }
  return std::string::npos;
}

inline YamlNode parse_value(const std::string& raw);

inline YamlNode parse_flow(const std::string& v) {
  YamlNode n;
  if (v.front() == '[') {
    n.type = YamlNode::Seq;
    for (auto& it : split_flow(v.substr(1, v.size() - 2))) n.seq.push_back(parse_value(it));
  } else {
    n.type = YamlNode::Map;
    for (auto& it : split_flow(v.substr(1, v.size() - 2))) {
      const size_t c = find_key_colon(it);
      if (c == std::string::npos) throw std::runtime_error("yaml: bad flow map entry '" + it + "'");
      n.map.emplace_back(unquote(trim(it.substr(0, c))), parse_value(trim(it.substr(c + 1))));
    }
  }
  return n;
}

inline YamlNode parse_value(const std::string& raw) {
  const std::string v = trim(raw);
  YamlNode n;
  if (v.empty() || v == "~" || v == "null") return n;
  if ((v.front() == '[' && v.back() == ']') || (v.front() == '{' && v.back() == '}')) return parse_flow(v);
  n.type = YamlNode::Scalar;
  n.scalar = unquote(v);
  return n;
}

struct Line {
  int indent;
  std::string text;
};

inline YamlNode parse_block(const std::vector<Line>& L, size_t& i, int indent);

inline YamlNode parse_map_from(const std::vector<Line>& L, size_t& i, int indent, const std::string& first) {
  // `first` is the text of the first "key: value" of a map whose keys sit at `indent`
  YamlNode n;
  n.type = YamlNode::Map;
  std::string text = first;
  for (;;) {
    const size_t c = find_key_colon(text);
    if (c == std::string::npos) throw std::runtime_error("yaml: expected 'key: value' in '" + text + "'");
    const std::string key = unquote(trim(text.substr(0, c)));
    const std::string val = trim(text.substr(c + 1));
    if (!val.empty()) {
      n.map.emplace_back(key, parse_value(val));
    } else if (i < L.size() && L[i].indent > indent) {
      n.map.emplace_back(key, parse_block(L, i, L[i].indent));
    } else if (i < L.size() && L[i].indent == indent && L[i].text.compare(0, 2, "- ") == 0) {
      n.map.emplace_back(key, parse_block(L, i, indent));  // sequence at the same indent as its key
    } else {
      n.map.emplace_back(key, YamlNode{});
    }
    if (i >= L.size() || L[i].indent != indent || L[i].text.compare(0, 2, "- ") == 0 || L[i].text == "-") break;
    text = L[i].text;
    i++;
  }
  return n;
}

inline YamlNode parse_block(const std::vector<Line>& L, size_t& i, int indent) {
  if (i >= L.size()) return YamlNode{};
  if (L[i].text.compare(0, 2, "- ") == 0 || L[i].text == "-") {
    YamlNode n;
    n.type = YamlNode::Seq;
    while (i < L.size() && L[i].indent == indent && (L[i].text.compare(0, 2, "- ") == 0 || L[i].text == "-")) {
      const std::string rest = L[i].text.size() > 2 ? trim(L[i].text.substr(2)) : "";
      const int child_indent = indent + 2 + int(L[i].text.size() > 2 ? L[i].text.find_first_not_of(' ', 2) - 2 : 0);
      i++;
      if (rest.empty()) {
        n.seq.push_back(i < L.size() && L[i].indent > indent ? parse_block(L, i, L[i].indent) : YamlNode{});
      } else if (find_key_colon(rest) != std::string::npos && rest.front() != '{' && rest.front() != '[' &&
                 rest.front() != '\'' && rest.front() != '"') {
        n.seq.push_back(parse_map_from(L, i, child_indent, rest));
      } else {
        n.seq.push_back(parse_value(rest));
      }
    }
    return n;
  }
  const std::string first = L[i].text;
  i++;
  if (find_key_colon(first) == std::string::npos) return parse_value(first);  // a bare scalar block, e.g. "~"
  return parse_map_from(L, i, indent, first);
}

}  // namespace detail

inline YamlNode yaml_parse(const std::string& text_in) {
  std::vector<detail::Line> lines;
  std::istringstream is(text_in);
  std::string ln;
  while (std::getline(is, ln)) {
    // comments go first: a ${VAR} mentioned in a comment must not be expanded (nor fail for lack of a default)
    const std::string s = detail::expand_env(detail::strip_comment(ln));
    const std::string t = detail::trim(s);
    if (t.empty() || t == "---") continue;
    int ind = 0;
    while (ind < int(s.size()) && s[ind] == ' ') ind++;
    lines.push_back({ind, t});
  }
  size_t i = 0;
  if (lines.empty()) return YamlNode{};
  YamlNode root = detail::parse_block(lines, i, lines[0].indent);
  if (i != lines.size()) throw std::runtime_error("yaml: could not parse line '" + lines[i].text + "'");
  return root;
}

inline YamlNode yaml_load_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("yaml: cannot open '" + path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  return yaml_parse(ss.str());
}

}  // namespace mlo_host
