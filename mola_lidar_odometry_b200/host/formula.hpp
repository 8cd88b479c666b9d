// formula.hpp — runtime formulas of the pipeline YAML (mp2p_icp::Parameterizable / ParameterSource as used at
// module/src/LidarOdometry.cpp:284,356,1571-1635): parameters are expressions over variables that the caller
// pushes per scan / per ICP iteration (ADAPTIVE_THRESHOLD_SIGMA, ICP_ITERATION, ESTIMATED_SENSOR_MAX_RANGE, vx..wz,
// robot_x..robot_roll, ...).  Grammar seen in the reference pipelines (pipelines/lidar3d-default.yaml:44-48,190,
// 198,233,289,301-302,309-310,316): numbers, identifiers, + - * / ^, unary minus, parentheses, max(a,b), min(a,b),
// sqrt(x), abs(x).
#pragma once
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mlo_host {

class ParameterSource {
 public:
  ParameterSource() : id_(next_id()++) {}
  void updateVariable(const std::string& name, double v) {
    auto it = index_.find(name);
    if (it == index_.end()) {
      index_.emplace(name, int(vals_.size()));
      vals_.push_back(v);
    } else {
      vals_[size_t(it->second)] = v;
    }
  }
  bool has(const std::string& name) const { return index_.count(name) != 0; }
  double get(const std::string& name) const {
    auto it = index_.find(name);
    if (it == index_.end()) throw std::runtime_error("formula: undefined variable '" + name + "'");
    return vals_[size_t(it->second)];
  }
  std::map<std::string, double> getVariableValues() const {
    std::map<std::string, double> m;
    for (const auto& kv : index_) m[kv.first] = vals_[size_t(kv.second)];
    return m;
  }
  // slot access for compiled formulas: variables are never removed, so a slot stays valid for this source
  int slot(const std::string& name) const {
    auto it = index_.find(name);
    return it == index_.end() ? -1 : it->second;
  }
  double at(int slot) const { return vals_[size_t(slot)]; }
  void set(int slot, double v) { vals_[size_t(slot)] = v; }
  uint64_t id() const { return id_; }

 private:
  static std::atomic<uint64_t>& next_id() {  // process-wide: a cached slot is only trusted for the source it came from
    static std::atomic<uint64_t> n{1};
    return n;
  }
  uint64_t id_;
  std::map<std::string, int> index_;
  std::vector<double> vals_;
};

// An expression is compiled on first use into a postfix program (the pipelines re-evaluate the same few formulas for
// every ICP iteration of every scan); syntax errors and undefined variables surface at evaluation time.
class Formula {
 public:
  Formula() = default;
  explicit Formula(std::string expr) : expr_(std::move(expr)) {}
  explicit Formula(double constant) : expr_(std::to_string(constant)) {}
  const std::string& text() const { return expr_; }
  bool empty() const { return expr_.empty(); }

  double eval(const ParameterSource& ps) const {
    if (!compiled_) compile();
    double st[32];
    int sp = 0;
    for (Op& o : prog_) {
      switch (o.code) {
        case NUM: st[sp++] = o.num; break;
        case VAR: {
          if (o.src != ps.id() || o.slot < 0) {
            o.slot = ps.slot(o.name);
            o.src = ps.id();
            if (o.slot < 0) {
              o.src = 0;
              throw std::runtime_error("formula: undefined variable '" + o.name + "'");
            }
          }
          st[sp++] = ps.at(o.slot);
          break;
        }
        case ADD: sp--; st[sp - 1] += st[sp]; break;
        case SUB: sp--; st[sp - 1] -= st[sp]; break;
        case MUL: sp--; st[sp - 1] *= st[sp]; break;
        case DIV: sp--; st[sp - 1] /= st[sp]; break;
        case POW: sp--; st[sp - 1] = std::pow(st[sp - 1], st[sp]); break;
        case NEG: st[sp - 1] = -st[sp - 1]; break;
        case SQRT: st[sp - 1] = std::sqrt(st[sp - 1]); break;
        case ABS: st[sp - 1] = std::fabs(st[sp - 1]); break;
        case MAX: sp--; st[sp - 1] = std::fmax(st[sp - 1], st[sp]); break;
        case MIN: sp--; st[sp - 1] = std::fmin(st[sp - 1], st[sp]); break;
      }
    }
    return st[0];
  }

 private:
  enum Code { NUM, VAR, ADD, SUB, MUL, DIV, POW, NEG, SQRT, ABS, MAX, MIN };
  struct Op {
    Code code;
    double num = 0;
    std::string name;
    int slot = -1;
    uint64_t src = 0;
  };
  void compile() const {
    std::vector<Op> prog;
    Parser p{expr_, 0, &prog, 0, 0};
    p.expr();
    p.skip();
    if (p.pos != expr_.size()) throw std::runtime_error("formula: trailing characters in '" + expr_ + "'");
    if (p.max_depth > 32) throw std::runtime_error("formula: expression too deep: '" + expr_ + "'");
    prog_.swap(prog);
    compiled_ = true;
  }
  struct Parser {
    const std::string& s;
    size_t pos;
    std::vector<Op>* out;
    int depth, max_depth;
    void emit(Code c, int delta) {
      out->push_back(Op{c});
      depth += delta;
      if (depth > max_depth) max_depth = depth;
    }
    void skip() {
      while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos]))) pos++;
    }
    bool eat(char c) {
      skip();
      if (pos < s.size() && s[pos] == c) {
        pos++;
        return true;
      }
      return false;
    }
    void expr() {  // + -
      term();
      for (;;) {
        if (eat('+')) { term(); emit(ADD, -1); }
        else if (eat('-')) { term(); emit(SUB, -1); }
        else return;
      }
    }
    void term() {  // * /
      power();
      for (;;) {
        if (eat('*')) { power(); emit(MUL, -1); }
        else if (eat('/')) { power(); emit(DIV, -1); }
        else return;
      }
    }
    void power() {  // ^ (right associative)
      unary();
      if (eat('^')) { power(); emit(POW, -1); }
    }
    void unary() {
      if (eat('-')) { unary(); emit(NEG, 0); return; }
      if (eat('+')) { unary(); return; }
      atom();
    }
    void atom() {
      skip();
      if (pos >= s.size()) throw std::runtime_error("formula: unexpected end in '" + s + "'");
      if (eat('(')) {
        expr();
        if (!eat(')')) throw std::runtime_error("formula: missing ')' in '" + s + "'");
        return;
      }
      const unsigned char c = static_cast<unsigned char>(s[pos]);
      if (std::isdigit(c) || c == '.') {
        char* end = nullptr;
        const double v = std::strtod(s.c_str() + pos, &end);
        pos = size_t(end - s.c_str());
        emit(NUM, +1);
        out->back().num = v;
        return;
      }
      if (std::isalpha(c) || c == '_') {
        size_t b = pos;
        while (pos < s.size() && (std::isalnum(static_cast<unsigned char>(s[pos])) || s[pos] == '_')) pos++;
        const std::string id = s.substr(b, pos - b);
        if (eat('(')) {
          expr();
          if (id == "sqrt" || id == "abs") {
            if (!eat(')')) throw std::runtime_error("formula: missing ')' after " + id);
            emit(id == "sqrt" ? SQRT : ABS, 0);
            return;
          }
          if (id == "max" || id == "min") {
            if (!eat(',')) throw std::runtime_error("formula: " + id + " needs two arguments");
            expr();
            if (!eat(')')) throw std::runtime_error("formula: missing ')' after " + id);
            emit(id == "max" ? MAX : MIN, -1);
            return;
          }
          throw std::runtime_error("formula: unknown function '" + id + "'");
        }
        if (id == "true" || id == "false") {
          emit(NUM, +1);
          out->back().num = id == "true" ? 1.0 : 0.0;
          return;
        }
        emit(VAR, +1);
        out->back().name = id;
        return;
      }
      throw std::runtime_error(std::string("formula: unexpected character '") + s[pos] + "' in '" + s + "'");
    }
  };
  std::string expr_;
  mutable std::vector<Op> prog_;
  mutable bool compiled_ = false;
};

}  // namespace mlo_host
