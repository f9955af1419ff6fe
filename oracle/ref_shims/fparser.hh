// TEST INFRASTRUCTURE (oracle/): stand-in for the fparser library's <fparser.hh> (absent here), enough for the
// reference's ParsedFunction (src/02_calculus/function_parser/ParsedFunction.{hpp,cpp}): expressions over named
// variables with + - * / ^, parentheses, numbers, named constants and the usual one-argument functions.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

template <class Value_t>
class FunctionParserBase {
 public:
  FunctionParserBase() {}
  bool AddConstant(const std::string& name, Value_t v) { _consts[name] = v; return true; }
  // returns -1 on success, otherwise the position of the error (as fparser does)
  int Parse(const std::string& expr, const std::string& vars) {
    _vars.clear();
    std::string cur;
    for (char c : vars + ",") {
      if (c == ',') { if (!cur.empty()) _vars.push_back(cur); cur.clear(); }
      else if (!std::isspace((unsigned char)c)) cur += c;
    }
    _src = expr;
    _pos = 0;
    _err.clear();
    _root = parse_sum();
    skip();
    if (_err.empty() && _pos != _src.size()) _err = "unexpected character";
    if (!_err.empty()) { _root.reset(); return (int)_pos; }
    return -1;
  }
  const char* ErrorMsg() const { return _err.c_str(); }
  void Optimize() {}
  Value_t Eval(const Value_t* x) const { return _root ? _root->eval(x) : Value_t(0); }

 private:
  struct Node {
    char op = 0;                // 'n' number, 'v' variable, 'f' function, '+', '-', '*', '/', '^', 'u' unary minus
    Value_t num = 0;
    int var = 0;
    Value_t (*fn)(Value_t) = nullptr;
    std::shared_ptr<Node> a, b;
    Value_t eval(const Value_t* x) const {
      switch (op) {
        case 'n': return num;
        case 'v': return x[var];
        case 'f': return fn(a->eval(x));
        case 'u': return -a->eval(x);
        case '+': return a->eval(x) + b->eval(x);
        case '-': return a->eval(x) - b->eval(x);
        case '*': return a->eval(x) * b->eval(x);
        case '/': return a->eval(x) / b->eval(x);
        default: return std::pow(a->eval(x), b->eval(x));
      }
    }
  };
  typedef std::shared_ptr<Node> P;
  void skip() { while (_pos < _src.size() && std::isspace((unsigned char)_src[_pos])) _pos++; }
  static P bin(char op, P a, P b) { P n(new Node); n->op = op; n->a = a; n->b = b; return n; }
  P parse_sum() {
    P l = parse_prod();
    for (;;) {
      skip();
      if (_pos < _src.size() && (_src[_pos] == '+' || _src[_pos] == '-')) { const char op = _src[_pos++]; l = bin(op, l, parse_prod()); }
      else return l;
    }
  }
  P parse_prod() {
    P l = parse_unary();
    for (;;) {
      skip();
      if (_pos < _src.size() && (_src[_pos] == '*' || _src[_pos] == '/')) { const char op = _src[_pos++]; l = bin(op, l, parse_unary()); }
      else return l;
    }
  }
  P parse_unary() {
    skip();
    if (_pos < _src.size() && _src[_pos] == '-') { _pos++; P n(new Node); n->op = 'u'; n->a = parse_unary(); return n; }
    if (_pos < _src.size() && _src[_pos] == '+') { _pos++; return parse_unary(); }
    return parse_pow();
  }
  P parse_pow() {
    P base = parse_atom();
    skip();
    if (_pos < _src.size() && _src[_pos] == '^') { _pos++; return bin('^', base, parse_unary()); }
    return base;
  }
  P parse_atom() {
    skip();
    P n(new Node);
    n->op = 'n';
    if (_pos >= _src.size()) { _err = "unexpected end of expression"; return n; }
    const char c = _src[_pos];
    if (c == '(') {
      _pos++;
      P e = parse_sum();
      skip();
      if (_pos < _src.size() && _src[_pos] == ')') _pos++; else if (_err.empty()) _err = "missing ')'";
      return e;
    }
    if (std::isdigit((unsigned char)c) || c == '.') {
      char* end = nullptr;
      n->num = (Value_t)std::strtod(_src.c_str() + _pos, &end);
      _pos = (size_t)(end - _src.c_str());
      return n;
    }
    if (std::isalpha((unsigned char)c) || c == '_') {
      std::string id;
      while (_pos < _src.size() && (std::isalnum((unsigned char)_src[_pos]) || _src[_pos] == '_')) id += _src[_pos++];
      for (size_t k = 0; k < _vars.size(); k++)
        if (_vars[k] == id) { n->op = 'v'; n->var = (int)k; return n; }
      auto it = _consts.find(id);
      if (it != _consts.end()) { n->num = it->second; return n; }
      static const std::map<std::string, Value_t (*)(Value_t)> fns = {
          {"sin", [](Value_t v) { return std::sin(v); }},   {"cos", [](Value_t v) { return std::cos(v); }},   {"tan", [](Value_t v) { return std::tan(v); }},
          {"exp", [](Value_t v) { return std::exp(v); }},   {"log", [](Value_t v) { return std::log(v); }},   {"sqrt", [](Value_t v) { return std::sqrt(v); }},
          {"abs", [](Value_t v) { return std::fabs(v); }},  {"sinh", [](Value_t v) { return std::sinh(v); }}, {"cosh", [](Value_t v) { return std::cosh(v); }},
          {"tanh", [](Value_t v) { return std::tanh(v); }}, {"asin", [](Value_t v) { return std::asin(v); }}, {"acos", [](Value_t v) { return std::acos(v); }},
          {"atan", [](Value_t v) { return std::atan(v); }}};
      auto f = fns.find(id);
      skip();
      if (f != fns.end() && _pos < _src.size() && _src[_pos] == '(') {
        n->op = 'f';
        n->fn = f->second;
        n->a = parse_atom();
        return n;
      }
      if (_err.empty()) _err = "unknown identifier '" + id + "'";
      return n;
    }
    if (_err.empty()) _err = "unexpected character";
    return n;
  }
  std::vector<std::string> _vars;
  std::map<std::string, Value_t> _consts;
  std::string _src, _err;
  size_t _pos = 0;
  P _root;
};
typedef FunctionParserBase<double> FunctionParser;
