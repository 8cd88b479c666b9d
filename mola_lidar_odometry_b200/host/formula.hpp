// formula.hpp — runtime formulas of the pipeline YAML (mp2p_icp::Parameterizable / ParameterSource as used at
// module/src/LidarOdometry.cpp:284,356,1571-1635): parameters are expressions over variables that the caller
// pushes per scan / per ICP iteration (ADAPTIVE_THRESHOLD_SIGMA, ICP_ITERATION, ESTIMATED_SENSOR_MAX_RANGE, vx..wz,
// robot_x..robot_roll, ...).  Grammar seen in the reference pipelines (pipelines/lidar3d-default.yaml:44-48,190,
// 198,233,289,301-302,309-310,316): numbers, identifiers, + - * / ^, unary minus, parentheses, max(a,b), min(a,b),
// sqrt(x), abs(x).
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>

namespace mlo_host {

class ParameterSource {
 public:
  void updateVariable(const std::string& name, double v) { vars_[name] = v; }
  bool has(const std::string& name) const { return vars_.count(name) != 0; }
  double get(const std::string& name) const {
    auto it = vars_.find(name);
    if (it == vars_.end()) throw std::runtime_error("formula: undefined variable '" + name + "'");
    return it->second;
  }
  const std::map<std::string, double>& getVariableValues() const { return vars_; }

 private:
  std::map<std::string, double> vars_;
};

class Formula {
 public:
  Formula() = default;
  explicit Formula(std::string expr) : expr_(std::move(expr)) {}
  explicit Formula(double constant) : expr_(std::to_string(constant)) {}
  const std::string& text() const { return expr_; }
  bool empty() const { return expr_.empty(); }

  double eval(const ParameterSource& ps) const {
    Parser p{expr_, 0, &ps};
    const double v = p.expr();
    p.skip();
    if (p.pos != expr_.size()) throw std::runtime_error("formula: trailing characters in '" + expr_ + "'");
    return v;
  }

 private:
  struct Parser {
    const std::string& s;
    size_t pos;
    const ParameterSource* ps;
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
    double expr() {  // + -
      double v = term();
      for (;;) {
        if (eat('+')) v += term();
        else if (eat('-')) v -= term();
        else return v;
      }
    }
    double term() {  // * /
      double v = power();
      for (;;) {
        if (eat('*')) v *= power();
        else if (eat('/')) v /= power();
        else return v;
      }
    }
    double power() {  // ^ (right associative)
      const double b = unary();
      if (eat('^')) return std::pow(b, power());
      return b;
    }
    double unary() {
      if (eat('-')) return -unary();
      if (eat('+')) return unary();
      return atom();
    }
    double atom() {
      skip();
      if (pos >= s.size()) throw std::runtime_error("formula: unexpected end in '" + s + "'");
      if (eat('(')) {
        const double v = expr();
        if (!eat(')')) throw std::runtime_error("formula: missing ')' in '" + s + "'");
        return v;
      }
      const unsigned char c = static_cast<unsigned char>(s[pos]);
      if (std::isdigit(c) || c == '.') {
        char* end = nullptr;
        const double v = std::strtod(s.c_str() + pos, &end);
        pos = size_t(end - s.c_str());
        return v;
      }
      if (std::isalpha(c) || c == '_') {
        size_t b = pos;
        while (pos < s.size() && (std::isalnum(static_cast<unsigned char>(s[pos])) || s[pos] == '_')) pos++;
        const std::string id = s.substr(b, pos - b);
        if (eat('(')) {
          const double a = expr();
          if (id == "sqrt" || id == "abs") {
            if (!eat(')')) throw std::runtime_error("formula: missing ')' after " + id);
            return id == "sqrt" ? std::sqrt(a) : std::fabs(a);
          }
          if (id == "max" || id == "min") {
            if (!eat(',')) throw std::runtime_error("formula: " + id + " needs two arguments");
            const double b2 = expr();
            if (!eat(')')) throw std::runtime_error("formula: missing ')' after " + id);
            return id == "max" ? std::fmax(a, b2) : std::fmin(a, b2);
          }
          throw std::runtime_error("formula: unknown function '" + id + "'");
        }
        if (id == "true") return 1.0;
        if (id == "false") return 0.0;
        return ps->get(id);
      }
      throw std::runtime_error(std::string("formula: unexpected character '") + s[pos] + "' in '" + s + "'");
    }
  };
  std::string expr_;
};

}  // namespace mlo_host
