// terms.cuh — compile-time description of the sum-factorisation sweeps: static_for, the term tables (which
// derivative flags of the swept direction go with which input/output component) and the output-group helpers.
// Included by kernels.cuh (inside namespace gsb) AND embedded as text for NVRTC (jit.cuh): self-contained, no
// standard headers; needs only the GSB_DEVICE / GSB_CX macros.
template <int I> struct IC { static constexpr int value = I; };
template <int I, int N, class F>
GSB_DEVICE void static_for(F &&f)
{
    if constexpr (I < N) {
        f(IC<I>{});
        static_for<I + 1, N>(f);
    }
}

// ------------------------------------------------------------------------------------
// Term tables.  A term (o, c, a, b) adds  B^(a)_owner(q) * B^(b)_partner(q) * in_c(q)  to out_o.
// a/b = 0: value, 1: first derivative in the swept direction.
#define GSB_PK(o, c, a, b) (((o) << 8) | ((c) << 2) | ((a) << 1) | (b))
template <class D>
struct TermOps {
    static GSB_CX int o(int k) { return D::pk(k) >> 8; }
    static GSB_CX int c(int k) { return (D::pk(k) >> 2) & 63; }
    static GSB_CX int a(int k) { return (D::pk(k) >> 1) & 1; }
    static GSB_CX int b(int k) { return D::pk(k) & 1; }
    // first term contributing to z[o][b] ?
    static GSB_CX bool first(int k) { for (int j = 0; j < k; ++j) if (o(j) == o(k) && b(j) == b(k)) return false; return true; }
    static GSB_CX bool has(int oo, int bb) { for (int j = 0; j < D::NT; ++j) if (o(j) == oo && b(j) == bb) return true; return false; }
    // output that holds the same number when owner and partner swap roles (symmetric coefficient tensor):
    // identity unless the table overrides it
    static GSB_CX int omirror(int oo) { return oo; }
    // output order in which the window kernel forms its groups: consecutive outputs of this order share inputs
    static GSB_CX int order(int i) { return i; }
    // symmetry of the form (window kernels): an output that is symmetric under the exchange of owner and partner is stored for
    // delta >= 0 only; a (virtual) input c is backed by the stored component in_src(c) and read directly (mode 1), at the mirrored
    // pair (i + delta, -delta) (mode 2), or mirrored only where delta < 0 (mode 0, symmetric component)
    // first sweeps store whole rows of deltas at the owner's exit for every degree (measured: S1 4.0 -> 3.55 ms at p=3; the second
    // sweep loses as much through register pressure, so it holds pairs back only where the quadrature size demands it)
    static GSB_CX bool whole_rows() { return false; }
    static GSB_CX bool out_sym(int) { return false; }
    static GSB_CX int in_src(int cc) { return cc; }
    static GSB_CX int in_srcm(int cc) { return D::in_src(cc); }     // the stored component behind input cc where it is read at the mirrored pair
    static GSB_CX int in_mode(int) { return 1; }
    static GSB_CX bool uses_c(unsigned omask, int cc) { for (int j = 0; j < D::NT; ++j) if (((omask >> o(j)) & 1u) && c(j) == cc) return true; return false; }
};
// symmetric coefficient tensor, 3-D (Poisson): D = {00,01,02,11,12,22}
struct T3SymS1 : TermOps<T3SymS1> { enum { NIN = 6, NOUT = 8, NT = 8 };
    static GSB_CX bool whole_rows() { return true; }
    static GSB_CX int pk(int k) { const int v[8] = {GSB_PK(0,0,1,1), GSB_PK(1,1,1,0), GSB_PK(2,1,0,1), GSB_PK(3,3,0,0),
                                                    GSB_PK(4,2,1,0), GSB_PK(5,2,0,1), GSB_PK(6,4,0,0), GSB_PK(7,5,0,0)}; return v[k]; }
    static GSB_CX int order(int i) { const int v[8] = {1, 2, 0, 3, 4, 5, 6, 7}; return v[i]; } };
// outputs g = 2*(a==2)+(b==2): flags still needed in direction 2
struct T3SymS2 : TermOps<T3SymS2> { enum { NIN = 8, NOUT = 4, NT = 9 };
    static GSB_CX int pk(int k) { const int v[9] = {GSB_PK(0,0,0,0), GSB_PK(0,1,0,1), GSB_PK(0,2,1,0), GSB_PK(0,3,1,1),
                                                    GSB_PK(1,4,0,0), GSB_PK(2,5,0,0), GSB_PK(1,6,1,0), GSB_PK(2,6,0,1),
                                                    GSB_PK(3,7,0,0)}; return v[k]; }
    static GSB_CX int omirror(int oo) { return oo == 1 ? 2 : (oo == 2 ? 1 : oo); }      // g = 2*alpha2 + beta2: swap the flags
    static GSB_CX int order(int i) { const int v[4] = {0, 3, 1, 2}; return v[i]; } };
// The same sweep on a first-sweep output stored for delta >= 0 only (every component; 8 x (p+1) instead of 8 x (2p+1) rows per
// function: 43 % less to write and to read back).  A pair with delta < 0 is read at its mirror (i + delta, -delta), where owner and
// partner have swapped roles: the components with one derivative flag trade places (T3SymS1's outputs 1 <-> 2, 4 <-> 5).
struct T3SymS2U : TermOps<T3SymS2U> { enum { NIN = 8, NOUT = 4, NT = 9 };
    static GSB_CX int pk(int k) { return T3SymS2::pk(k); }
    static GSB_CX int omirror(int oo) { return T3SymS2::omirror(oo); }
    static GSB_CX int order(int i) { return T3SymS2::order(i); }
    static GSB_CX int in_mode(int) { return 0; }
    static GSB_CX int in_srcm(int cc) { return cc == 1 ? 2 : (cc == 2 ? 1 : (cc == 4 ? 5 : (cc == 5 ? 4 : cc))); } };
// last direction of any gradient-gradient form: in_g, g = 2*a+b
struct TLast : TermOps<TLast> { enum { NIN = 4, NOUT = 1, NT = 4 };
    static GSB_CX int pk(int k) { const int v[4] = {GSB_PK(0,0,0,0), GSB_PK(0,1,0,1), GSB_PK(0,2,1,0), GSB_PK(0,3,1,1)}; return v[k]; } };
// general (non-symmetric) tensor, 3-D: c = 3a+b
struct T3GenS1 : TermOps<T3GenS1> { enum { NIN = 9, NOUT = 9, NT = 9 };
    static GSB_CX bool whole_rows() { return true; }
    static GSB_CX int pk(int k) { const int v[9] = {GSB_PK(0,0,1,1), GSB_PK(1,1,1,0), GSB_PK(2,2,1,0), GSB_PK(3,3,0,1), GSB_PK(4,4,0,0),
                                                    GSB_PK(5,5,0,0), GSB_PK(6,6,0,1), GSB_PK(7,7,0,0), GSB_PK(8,8,0,0)}; return v[k]; } };
struct T3GenS2 : TermOps<T3GenS2> { enum { NIN = 9, NOUT = 4, NT = 9 };
    static GSB_CX int pk(int k) { const int v[9] = {GSB_PK(0,0,0,0), GSB_PK(0,1,0,1), GSB_PK(1,2,0,0), GSB_PK(0,3,1,0), GSB_PK(0,4,1,1),
                                                    GSB_PK(1,5,1,0), GSB_PK(2,6,0,0), GSB_PK(2,7,0,1), GSB_PK(3,8,0,0)}; return v[k]; } };
// 2-D: symmetric D = {00,01,11}; general c = 2a+b.  Outputs g = 2*(a==1)+(b==1).
struct T2SymS1 : TermOps<T2SymS1> { enum { NIN = 3, NOUT = 4, NT = 4 };
    static GSB_CX bool whole_rows() { return true; }
    static GSB_CX int pk(int k) { const int v[4] = {GSB_PK(0,0,1,1), GSB_PK(1,1,1,0), GSB_PK(2,1,0,1), GSB_PK(3,2,0,0)}; return v[k]; }
    static GSB_CX int order(int i) { const int v[4] = {1, 2, 0, 3}; return v[i]; } };
struct T2GenS1 : TermOps<T2GenS1> { enum { NIN = 4, NOUT = 4, NT = 4 };
    static GSB_CX bool whole_rows() { return true; }
    static GSB_CX int pk(int k) { const int v[4] = {GSB_PK(0,0,1,1), GSB_PK(1,1,1,0), GSB_PK(2,2,0,1), GSB_PK(3,3,0,0)}; return v[k]; } };
// mass-type form: one scalar density, no derivatives, every direction
struct TMass : TermOps<TMass> { enum { NIN = 1, NOUT = 1, NT = 1 };
    static GSB_CX int pk(int) { return GSB_PK(0,0,0,0); } };


template <class T, int NG> GSB_CX unsigned group_mask(int gi)
{
    unsigned m = 0;
    for (int i = gi * NG; i < (gi + 1) * NG && i < T::NOUT; ++i) m |= 1u << T::order(i);
    return m;
}
GSB_CX int mask_rank(unsigned m, int o) { int r = 0; for (int i = 0; i < o; ++i) if ((m >> i) & 1u) ++r; return r; }
GSB_CX int mask_count(unsigned m) { int r = 0; for (int i = 0; i < 32; ++i) if ((m >> i) & 1u) ++r; return r; }

template <class T> GSB_CX int used_count(unsigned m) { int n = 0; for (int c = 0; c < T::NIN; ++c) if (T::uses_c(m, c)) ++n; return n; }
template <class T> GSB_CX int used_rank(unsigned m, int cc) { int n = 0; for (int c = 0; c < cc; ++c) if (T::uses_c(m, c)) ++n; return n; }
