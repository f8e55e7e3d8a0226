// expr.cpp — host-side compiler for source-term strings (gsFunctionExpr replacement).
//
// The reference evaluates the source term f and boundary data through exprtk
// (gsFunctionExpr.hpp:513-533, one point at a time under an omp critical).  exprtk
// cannot run on the device, so the strings the BASELINE configs use are compiled into a
// reverse-polish program that K0 interprets per quadrature point (option b).
// Grammar (precedence as exprtk):  expr := term (('+'|'-') term)*
//   term := unary (('*'|'/') unary)*     unary := ('-'|'+') unary | power
//   power := atom ('^' unary)?           atom := number | x | y | z | pi | f '(' expr ')' | '(' expr ')'
#include "../../include/gsb200.h"
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace gsb {
void set_error(const char *fmt, ...);
}

namespace {

struct Parser {
    const char *s;
    std::vector<int32_t> ops;
    std::vector<double> consts;
    std::string err;
    int depth, maxdepth;
    // one entry per value currently on the evaluation stack: where its code starts and, if it
    // is a compile-time constant, its value (constant sub-expressions such as 3*pi^2 are folded
    // so the device interpreter does not evaluate pow() at every quadrature point)
    struct Item { size_t start; bool is_const; double value; bool pi_scaled; size_t pi_mul_at; std::vector<int32_t> rest; };
    std::vector<Item> items;

    explicit Parser(const char *str) : s(str), depth(0), maxdepth(0) {}
    void skip() { while (*s && std::isspace((unsigned char)*s)) ++s; }
    void push(int d) { depth += d; if (depth > maxdepth) maxdepth = depth; }
    static double apply1(int op, double a) {
        switch (op) {
        case GSB200_OP_NEG: return -a; case GSB200_OP_SIN: return std::sin(a); case GSB200_OP_COS: return std::cos(a);
        case GSB200_OP_TAN: return std::tan(a); case GSB200_OP_EXP: return std::exp(a); case GSB200_OP_LOG: return std::log(a);
        case GSB200_OP_SQRT: return std::sqrt(a); case GSB200_OP_ABS: return std::fabs(a); case GSB200_OP_TANH: return std::tanh(a);
        case GSB200_OP_SINH: return std::sinh(a); case GSB200_OP_COSH: return std::cosh(a); case GSB200_OP_SQR: return a * a;
        case GSB200_OP_SINPI: return std::sin(3.14159265358979323846 * a); case GSB200_OP_COSPI: return std::cos(3.14159265358979323846 * a);
        }
        return NAN;
    }
    static double apply2(int op, double a, double b) {
        switch (op) {
        case GSB200_OP_ADD: return a + b; case GSB200_OP_SUB: return a - b; case GSB200_OP_MUL: return a * b;
        case GSB200_OP_DIV: return a / b; case GSB200_OP_POW: return std::pow(a, b);
        }
        return NAN;
    }
    void raw_const(double v) {
        size_t k = 0;
        for (; k < consts.size(); ++k) if (std::memcmp(&consts[k], &v, sizeof v) == 0) break;
        if (k == consts.size()) consts.push_back(v);
        ops.push_back(GSB200_OP_CONST); ops.push_back((int32_t)k);
    }
    void emit(int op) {
        const bool leaf = op == GSB200_OP_X || op == GSB200_OP_Y || op == GSB200_OP_Z;
        const bool binary = op == GSB200_OP_ADD || op == GSB200_OP_SUB || op == GSB200_OP_MUL || op == GSB200_OP_DIV || op == GSB200_OP_POW;
        if (leaf) { Item it = {ops.size(), false, 0.0, false, 0, std::vector<int32_t>()}; items.push_back(it); ops.push_back(op); return; }
        if (binary) {
            Item b = items.back(); items.pop_back();
            Item &a = items.back();
            if (a.is_const && b.is_const) { a.value = apply2(op, a.value, b.value); ops.resize(a.start); raw_const(a.value); return; }
            if (op == GSB200_OP_POW && b.is_const && b.value == 2.0) { ops.resize(b.start); ops.push_back(GSB200_OP_SQR); a.is_const = false; a.pi_scaled = false; return; }
            const double PI = 3.14159265358979323846264338328;
            if (op == GSB200_OP_MUL && (a.is_const != b.is_const) && (a.is_const ? a.value : b.value) == PI) {
                // remember how to strip the factor pi again should sin()/cos() be applied to this product
                std::vector<int32_t> rest;
                if (a.is_const) rest.assign(ops.begin() + b.start, ops.end());       // pi * E : E follows the constant
                else rest.assign(ops.begin() + a.start, ops.begin() + b.start);       // E * pi
                a.pi_scaled = true; a.pi_mul_at = a.start; a.rest = rest;
                a.is_const = false; ops.push_back(op); return;
            }
            a.is_const = false; a.pi_scaled = false; ops.push_back(op); return;
        }
        Item &a = items.back();   // unary
        if (a.is_const) { a.value = apply1(op, a.value); ops.resize(a.start); raw_const(a.value); return; }
        if ((op == GSB200_OP_SIN || op == GSB200_OP_COS) && a.pi_scaled) {
            // sin(pi*E): drop the multiplication by pi and use the range-reduction-free sinpi/cospi
            ops.resize(a.pi_mul_at);
            ops.insert(ops.end(), a.rest.begin(), a.rest.end());
            ops.push_back(op == GSB200_OP_SIN ? GSB200_OP_SINPI : GSB200_OP_COSPI);
            a.pi_scaled = false;
            return;
        }
        a.pi_scaled = false;
        ops.push_back(op);
    }
    void emit_const(double v) {
        Item it = {ops.size(), true, v, false, 0, std::vector<int32_t>()}; items.push_back(it);
        raw_const(v); push(1);
    }
    bool expr() {
        if (!term()) return false;
        for (;;) {
            skip();
            if (*s == '+') { ++s; if (!term()) return false; emit(GSB200_OP_ADD); push(-1); }
            else if (*s == '-') { ++s; if (!term()) return false; emit(GSB200_OP_SUB); push(-1); }
            else return true;
        }
    }
    bool term() {
        if (!unary()) return false;
        for (;;) {
            skip();
            if (*s == '*') { ++s; if (!unary()) return false; emit(GSB200_OP_MUL); push(-1); }
            else if (*s == '/') { ++s; if (!unary()) return false; emit(GSB200_OP_DIV); push(-1); }
            else return true;
        }
    }
    bool unary() {
        skip();
        if (*s == '-') { ++s; if (!unary()) return false; emit(GSB200_OP_NEG); return true; }
        if (*s == '+') { ++s; return unary(); }
        return power();
    }
    bool power() {
        if (!atom()) return false;
        skip();
        if (*s == '^') { ++s; if (!unary()) return false; emit(GSB200_OP_POW); push(-1); }
        return true;
    }
    bool atom() {
        skip();
        if (*s == '(') {
            ++s; if (!expr()) return false; skip();
            if (*s != ')') { err = "expected ')'"; return false; }
            ++s; return true;
        }
        if (std::isdigit((unsigned char)*s) || *s == '.') {
            char *end = 0; const double v = std::strtod(s, &end);
            if (end == s) { err = "bad number"; return false; }
            s = end; emit_const(v); return true;
        }
        if (std::isalpha((unsigned char)*s) || *s == '_') {
            std::string id;
            while (std::isalnum((unsigned char)*s) || *s == '_') id += (char)std::tolower((unsigned char)*s++);
            skip();
            if (*s == '(') {
                static const struct { const char *n; int op; } fn[] = {
                    {"sin", GSB200_OP_SIN}, {"cos", GSB200_OP_COS}, {"tan", GSB200_OP_TAN}, {"exp", GSB200_OP_EXP},
                    {"log", GSB200_OP_LOG}, {"sqrt", GSB200_OP_SQRT}, {"abs", GSB200_OP_ABS}, {"tanh", GSB200_OP_TANH},
                    {"sinh", GSB200_OP_SINH}, {"cosh", GSB200_OP_COSH}};
                int op = -1;
                for (size_t k = 0; k < sizeof fn / sizeof fn[0]; ++k) if (id == fn[k].n) op = fn[k].op;
                if (op < 0) { err = "unknown function '" + id + "'"; return false; }
                ++s; if (!expr()) return false; skip();
                if (*s != ')') { err = "expected ')'"; return false; }
                ++s; emit(op); return true;
            }
            if (id == "x") { emit(GSB200_OP_X); push(1); return true; }
            if (id == "y") { emit(GSB200_OP_Y); push(1); return true; }
            if (id == "z") { emit(GSB200_OP_Z); push(1); return true; }
            if (id == "pi") { emit_const(3.14159265358979323846264338328); return true; }
            err = "unknown symbol '" + id + "'"; return false;
        }
        err = std::string("unexpected character '") + *s + "'"; return false;
    }
};

} // namespace

extern "C" int gsb200_expr_compile(const char *expr, int32_t *ops, int32_t ops_cap, int32_t *nops,
                                   double *consts, int32_t consts_cap, int32_t *nconsts)
{
    if (!expr || !ops || !consts || !nops || !nconsts) { gsb::set_error("expr_compile: null argument"); return GSB200_EINVAL; }
    Parser P(expr);
    bool ok = P.expr();
    P.skip();
    if (ok && *P.s) { ok = false; P.err = std::string("trailing input at '") + P.s + "'"; }
    if (!ok) { gsb::set_error("cannot parse '%s': %s", expr, P.err.c_str()); return GSB200_EUNSUPPORTED; }
    if ((int)P.ops.size() > ops_cap || (int)P.ops.size() > GSB200_PROGRAM_MAX_OPS || (int)P.consts.size() > consts_cap ||
        P.maxdepth > GSB200_PROGRAM_MAX_STACK) {
        gsb::set_error("expression '%s' too long for the device program limits", expr); return GSB200_EUNSUPPORTED; }
    std::memcpy(ops, P.ops.data(), P.ops.size() * sizeof(int32_t));
    std::memcpy(consts, P.consts.data(), P.consts.size() * sizeof(double));
    *nops = (int32_t)P.ops.size(); *nconsts = (int32_t)P.consts.size();
    return GSB200_OK;
}
