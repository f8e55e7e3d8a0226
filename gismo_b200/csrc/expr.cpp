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

    explicit Parser(const char *str) : s(str), depth(0), maxdepth(0) {}
    void skip() { while (*s && std::isspace((unsigned char)*s)) ++s; }
    void push(int d) { depth += d; if (depth > maxdepth) maxdepth = depth; }
    void emit(int op) { ops.push_back(op); }
    void emit_const(double v) {
        size_t k = 0;
        for (; k < consts.size(); ++k) if (std::memcmp(&consts[k], &v, sizeof v) == 0) break;
        if (k == consts.size()) consts.push_back(v);
        ops.push_back(GSB200_OP_CONST); ops.push_back((int32_t)k); push(1);
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
