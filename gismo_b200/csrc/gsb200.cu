// gsb200.cu — host orchestration + C ABI (include/gsb200.h) of the B200 assembly path.
// Compiled by nvcc for sm_100a into gismo_b200/csrc/libgsb200.so.  No torch, no Eigen.
#include "launch.h"
#include <algorithm>
#include <cstdarg>
#include <string>
#include <vector>
#ifndef GSB200_EMULATE
#include <cub/device/device_scan.cuh>
#include <cuda.h>   // CUtensorMap + enums only; the encoder is fetched with cudaGetDriverEntryPoint
#endif

#ifndef GSB200_EMULATE
#include "jit.cuh"
#include <dlfcn.h>
#include <sys/mman.h>
#include <atomic>
#include <memory>
#include <mutex>
#include <set>
#include <thread>
#endif

namespace gsb {

static thread_local char g_err[1024] = "";
static thread_local int g_launches = 0;      // per host thread: two threads driving two assemblers do not see each other's planning pass
void set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
void note_launch() { ++g_launches; }
static thread_local bool g_dry = false;
bool dry_run() { return g_dry; }
#ifndef GSB200_EMULATE
int grant_dynamic_smem(const void *kfn, size_t smem)
{
    if (smem <= 48 * 1024) return 0;
    static std::mutex mu; static std::set<std::pair<int, const void *>> granted;       // per (device, kernel)
    int dev = 0; cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (granted.count({dev, kfn})) return 0;
    GSB_TRY(dev_check(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024), "cudaFuncSetAttribute"));
    granted.insert({dev, kfn});
    return 0;
}
#endif


// Gauss-Legendre rule on [-1,1] (the reference tabulates the same numbers to 30 digits,
// gsGaussRule.hpp:218-547): Newton on P_n in long double, rounded to double.
static void gauss_rule(int n, std::vector<double> &x, std::vector<double> &w)
{
    x.assign(n, 0.0); w.assign(n, 0.0);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(pi * (i + 0.75L) / (n + 0.5L)), dp = 1;
        for (int it = 0; it < 100; ++it) {
            long double p0 = 1, p1 = z;
            for (int k = 2; k <= n; ++k) { const long double p2 = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
            dp = n * (z * p1 - p0) / (z * z - 1);
            const long double dz = p1 / dp;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        long double p0 = 1, p1 = z;
        for (int k = 2; k <= n; ++k) { const long double p2 = ((2 * k - 1) * z * p1 - (k - 1) * p0) / k; p0 = p1; p1 = p2; }
        dp = n * (z * p1 - p0) / (z * z - 1);
        const long double ww = 2 / ((1 - z * z) * dp * dp);
        x[i] = (double)(-z); x[n - 1 - i] = (double)z;
        w[i] = w[n - 1 - i] = (double)ww;
    }
    if (n % 2) x[n / 2] = 0.0;
}

template <class T>
static int upload(T **dptr, const std::vector<T> &h, stream_t s)
{
    GSB_TRY(dev_malloc((void **)dptr, h.size() * sizeof(T)));
    if (!h.empty()) GSB_TRY(dev_h2d(*dptr, h.data(), h.size() * sizeof(T), s));
    return 0;
}

// ------------------------------------------------------------------ per-direction tables
struct Dir1D {
    int p = 0, q = 0, nfun = 0, nel = 0, Q = 0;
    std::vector<int> span, first, nexit, plo, phi, ffirst, flast;
    int *d_first = 0, *d_nexit = 0, *d_plo = 0, *d_phi = 0, *d_ffirst = 0, *d_flast = 0;
    double2 *d_tab = 0, *d_tabl = 0; double *d_upt = 0, *d_hpt = 0, *d_gw = 0, *d_gwp = 0;
    double2 *d_gtab = 0; int *d_gfirst = 0; int pg1 = 0, ngeo = 0;
    // basis values at the two ends of the parameter interval (Neumann sides): [0] lower, [1] upper
    double bval[2][GSB_MAXP + 1]; int bfirst[2]; double2 bgeo[2][GSB_MAXP + 1]; int bgfirst[2];
    void release() {
        dev_free(d_first); dev_free(d_nexit); dev_free(d_plo); dev_free(d_phi); dev_free(d_ffirst); dev_free(d_flast);
        dev_free(d_tab); dev_free(d_tabl); dev_free(d_upt); dev_free(d_hpt); dev_free(d_gw); dev_free(d_gwp); dev_free(d_gtab); dev_free(d_gfirst);
    }
};

static int build_dir(Dir1D &d, const double *kn, int nk, int p, int q, const double *gkn, int gnk, int gp, stream_t s)
{
    d.p = p; d.q = q; d.nfun = nk - p - 1;
    if (d.nfun < 1) { set_error("knot vector too short"); return GSB200_EINVAL; }
    for (int i = 1; i < nk; ++i) if (kn[i] < kn[i - 1]) { set_error("knot vector not sorted"); return GSB200_EINVAL; }
    for (int e = p; e < nk - p - 1; ++e) if (kn[e] < kn[e + 1]) d.span.push_back(e);
    d.nel = (int)d.span.size();
    if (!d.nel) { set_error("knot vector has no element"); return GSB200_EINVAL; }
    d.Q = d.nel * q;
    const int p1 = p + 1;
    d.first.resize(d.nel); d.nexit.resize(d.nel);
    d.ffirst.assign(d.nfun, d.nel); d.flast.assign(d.nfun, -1);
    for (int e = 0; e < d.nel; ++e) {
        d.first[e] = d.span[e] - p;
        for (int a = 0; a < p1; ++a) { const int f = d.first[e] + a; d.ffirst[f] = std::min(d.ffirst[f], e); d.flast[f] = std::max(d.flast[f], e); }
    }
    for (int e = 0; e < d.nel; ++e) d.nexit[e] = (e + 1 < d.nel) ? d.first[e + 1] - d.first[e] : p1;
    d.plo.resize(d.nfun); d.phi.resize(d.nfun);
    for (int f = 0; f < d.nfun; ++f) {
        if (d.flast[f] < 0) { set_error("basis function %d has empty support (knot multiplicity > degree+1?)", f); return GSB200_EINVAL; }
        d.plo[f] = d.first[d.ffirst[f]]; d.phi[f] = d.first[d.flast[f]] + p;
    }
    for (int side = 0; side < 2; ++side) {      // boundary evaluations for the Neumann load
        const double ub = side ? kn[nk - p - 1] : kn[p];
        const int s = side ? d.span.back() : d.span.front();
        double val[GSB_MAXP + 1], der[GSB_MAXP + 1];
        bspline_ders(kn, p, s, ub, val, der);
        for (int a = 0; a <= p; ++a) d.bval[side][a] = val[a];
        d.bfirst[side] = s - p;
        int gs = gp;                              // geometry span containing ub
        if (side) { gs = gnk - gp - 2; while (gkn[gs] == gkn[gs + 1]) --gs; } else { while (gkn[gs] == gkn[gs + 1]) ++gs; }
        bspline_ders(gkn, gp, gs, ub, val, der);
        for (int a = 0; a <= gp; ++a) d.bgeo[side][a] = make_double2(val[a], der[a]);
        d.bgfirst[side] = gs - gp;
    }
    std::vector<double> gx, gw;
    gauss_rule(q, gx, gw);
    std::vector<double> knv(kn, kn + nk), gknv(gkn, gkn + gnk);
    double *d_kn = 0, *d_gkn = 0, *d_gx = 0; int *d_span = 0;
    GSB_TRY(upload(&d_kn, knv, s)); GSB_TRY(upload(&d_gkn, gknv, s)); GSB_TRY(upload(&d_gx, gx, s));
    GSB_TRY(upload(&d_span, d.span, s)); GSB_TRY(upload(&d.d_gw, gw, s));
    GSB_TRY(upload(&d.d_first, d.first, s)); GSB_TRY(upload(&d.d_nexit, d.nexit, s));
    GSB_TRY(upload(&d.d_plo, d.plo, s)); GSB_TRY(upload(&d.d_phi, d.phi, s));
    GSB_TRY(upload(&d.d_ffirst, d.ffirst, s)); GSB_TRY(upload(&d.d_flast, d.flast, s));
    GSB_TRY(dev_malloc((void **)&d.d_tab, sizeof(double2) * (size_t)d.Q * p1));
    GSB_TRY(dev_malloc((void **)&d.d_tabl, sizeof(double2) * (size_t)d.Q * p1));
    GSB_TRY(dev_malloc((void **)&d.d_upt, sizeof(double) * (size_t)d.Q));
    GSB_TRY(dev_malloc((void **)&d.d_hpt, sizeof(double) * (size_t)d.Q));
    GSB_TRY(dev_malloc((void **)&d.d_gwp, sizeof(double) * (size_t)d.Q));
    BasisTableArgs B; B.knots = d_kn; B.span = d_span; B.gnodes = d_gx; B.gweights = d.d_gw; B.p = p; B.nel = d.nel; B.q = q;
    B.tab = d.d_tab; B.tabl = d.d_tabl; B.upt = d.d_upt; B.hpt = d.d_hpt; B.gwp = d.d_gwp;
    GSB_LAUNCH(k_basis_table, dim3((d.Q + 127) / 128), dim3(128), s, B);
    d.pg1 = gp + 1; d.ngeo = gnk - gp - 1;
    GSB_TRY(dev_malloc((void **)&d.d_gtab, sizeof(double2) * (size_t)d.Q * d.pg1));
    GSB_TRY(dev_malloc((void **)&d.d_gfirst, sizeof(int) * (size_t)d.Q));
    GeoTableArgs G; G.knots = d_gkn; G.nknots = gnk; G.p = gp; G.npts = d.Q; G.upt = d.d_upt; G.gtab = d.d_gtab; G.gfirst = d.d_gfirst;
    GSB_LAUNCH(k_geo_table, dim3((d.Q + 127) / 128), dim3(128), s, G);
    GSB_TRY(dev_last_error("table kernels"));
    GSB_TRY(dev_sync(s));
    dev_free(d_kn); dev_free(d_gkn); dev_free(d_gx); dev_free(d_span);
    return 0;
}

// segments of a sweep: functions [xa,xb) split into nseg contiguous exit ranges
static std::vector<int> make_segments(const Dir1D &d, int xa, int xb, int nseg)
{
    std::vector<int> s;
    nseg = std::max(1, std::min(nseg, xb - xa));
    for (int k = 0; k < nseg; ++k) {
        const int xs = xa + (int)((i64)(xb - xa) * k / nseg), xe = xa + (int)((i64)(xb - xa) * (k + 1) / nseg);
        if (xe <= xs) continue;
        s.push_back(d.ffirst[xs]); s.push_back(d.flast[xe - 1] + 1); s.push_back(xs); s.push_back(xe);
    }
    return s;
}

struct PatchDev {
    int dim = 0;
    Dir1D dir[3];
    i64 nb = 0, ngeo_total = 0;
    int *d_dofmap = 0; double *d_coefs = 0, *d_weights = 0;
    unsigned char *d_colflag = 0; unsigned *d_st = 0, *d_st2 = 0; int nrun = 1; i64 *d_ownrec = 0;
    int own_lo = 0, own_hi = 0;     // owner range along the last direction on this rank
    double *d_lc = 0;               // line coefficients of the geometry (fused.cuh, k_line_coefs), 0: fused first sweep not available
    std::vector<int> sufmin;        // sufmin[x] = smallest free column with a preimage in last-direction layers >= x (scalar spaces): columns below it are final once layers < x are done
    void release() { for (int k = 0; k < 3; ++k) dir[k].release(); dev_free(d_dofmap); dev_free(d_coefs); dev_free(d_weights); dev_free(d_colflag); dev_free(d_st); dev_free(d_st2); dev_free(d_ownrec); dev_free(d_lc); }
};

} // namespace gsb

using namespace gsb;
namespace gsb { void destroy_comm(void *comm); struct ncclUniqueIdPod { char internal[128]; }; }

struct gsb200_assembler {
    int device = 0, dim = 0, form = 0, ncomp = 1, nfree = 0, nfixed = 0, nrhs = 1, rhs_kind = 0, rank = 0, nranks = 1;
    double coef[4] = {0, 0, 0, 0};
    std::vector<PatchDev> patches;
    double *d_fixed = 0, *d_rhs = 0, *d_values = 0;
    i64 *d_colptr = 0; int *d_inner = 0; int *d_npre = 0;
    i64 nnz = 0;
    std::vector<DevProgram> progs; std::vector<void *> prog_bufs;
    std::vector<HostProgram> progs_host;   // the same source-term programs, for the NVRTC path (jit.cuh)
    int jit_launches = 0;                  // geometry launches of the last assemble() that ran the compiled source term
    struct NeumannSide { int patch, side, ndata; DevProgram prog[3]; };
    std::vector<NeumannSide> neumann; double *d_face = 0; size_t face_cap = 0;
    stream_t stream = 0;
    bool pattern_built = false, assembled = false, any_generic = false;
    i64 ws_limit = 0; void *ws[4] = {0, 0, 0, 0}; size_t ws_size[4] = {0, 0, 0, 0};      // one workspace per lane (patches of a multi-patch problem run on up to 4 streams)
#ifndef GSB200_EMULATE
    cudaStream_t lanes[4] = {0, 0, 0, 0}; cudaEvent_t lane_ev[4] = {0, 0, 0, 0};
#endif
    gsb200_timings tm;
    int *d_seg = 0; size_t seg_cap = 0; std::vector<int> seg_host; size_t seg_used = 0;
    bool plan_valid = false; i64 plan_limit = 0; size_t ev_used = 0;
#ifndef GSB200_EMULATE
    std::vector<cudaEvent_t> ev; std::vector<int> ev_tag;   // tag: 0 geometry, 1..3 sweeps, 4 rhs, 5 total-begin, 6 total-end
#endif
    int *d_outer32 = 0;                 // narrowed column pointers for the host-side gsSparseMatrix
#ifndef GSB200_EMULATE
    cudaStream_t copy_stream = 0; cudaEvent_t ev_done = 0;
    void *stage[4] = {0, 0, 0, 0}; cudaEvent_t stage_ev[4] = {0, 0, 0, 0};     // pinned staging ring for pageable destinations
    std::vector<cudaEvent_t> chunk_ev;  // recorded behind the last kernel of every chunk: finished column ranges travel while later chunks integrate
#endif
    int deliver_chunks = 1, plan_chunks = 1;      // chunks the last direction is cut into for streamed delivery (requested / planned)
    std::vector<int> chunk_cols; std::vector<i64> chunk_off;   // per chunk: first column / value offset that is NOT yet final behind it
    // ---- consumer side (consumer.cuh)
    std::vector<int> patch_owner;               // rank that integrates each patch (several patches: greedy balance by element count)
    std::vector<int> mid_dof;                   // per (patch, component): global DOF of a function in the middle of the owned part, -1 if none
    std::vector<int> coupled_runs;              // [a, b) pairs: runs of columns with pre-images in more than one patch
    std::vector<i64> coupled_off;               // their value offsets (after the pattern is built)
    void *comm = 0; bool comm_owned = false;    // ncclComm_t
    gsb200_allreduce_fn ar_fn = 0; void *ar_ctx = 0;
    i64 xchg_bytes = 0; int xchg_calls = 0;
    unsigned char *d_reg = 0; int *d_regoff = 0, *d_reglen = 0; int reg_ntab = 0, reg_stride = 1; bool spmv_ready = false;
    int own_c0 = 0, own_c1 = 0, need_lo = 0, need_hi = 0; bool own_contig = true, cg_halo_mode = false; double cg_loop_ms = 0;
    double *cgv[8] = {0, 0, 0, 0, 0, 0, 0, 0}, *cg_b = 0;      // CG work vectors x r z p q d h + scalars; user rhs
    ~gsb200_assembler() {
        dev_sync(stream);                  // frees below are not ordered behind this stream's kernels
        for (auto &p : patches) p.release();
        dev_free(d_fixed); dev_free(d_rhs); dev_free(d_values); dev_free(d_colptr); dev_free(d_inner); dev_free(d_npre);
        for (void *b : prog_bufs) dev_free(b);
        for (int k = 0; k < 4; ++k) dev_free(ws[k]);
        dev_free(d_seg); dev_free(d_face); dev_free(d_outer32);
#ifndef GSB200_EMULATE
        for (int k = 0; k < 4; ++k) { if (lanes[k]) { cudaStreamSynchronize(lanes[k]); cudaStreamDestroy(lanes[k]); } if (lane_ev[k]) cudaEventDestroy(lane_ev[k]); }
#endif
#ifndef GSB200_EMULATE
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (ev_done) cudaEventDestroy(ev_done);
        for (int k = 0; k < 4; ++k) { if (stage[k]) cudaFreeHost(stage[k]); if (stage_ev[k]) cudaEventDestroy(stage_ev[k]); }
        for (auto e : chunk_ev) cudaEventDestroy(e);
#endif
        for (int k = 0; k < 8; ++k) dev_free(cgv[k]);
        dev_free(cg_b); dev_free(d_reg); dev_free(d_regoff); dev_free(d_reglen);
#ifndef GSB200_EMULATE
        for (auto e : ev) cudaEventDestroy(e);
        if (comm && comm_owned) gsb::destroy_comm(comm);
#endif
    }
};

namespace gsb {

static int num_nodes(double quA, int quB, int p) { return (int)(quA * p + quB + 0.5); }

static int select_device(int device)
{
#ifndef GSB200_EMULATE
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { set_error("no CUDA device available (%s): the B200 path has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e)); return GSB200_ENODEVICE; }
    if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return GSB200_EINVAL; }
    return dev_check(cudaSetDevice(device), "cudaSetDevice");
#else
    (void)device; return 0;
#endif
}

static int scan_lengths(gsb200_assembler *a, unsigned long long *d_len, i64 *d_ptr, int n)
{
#ifndef GSB200_EMULATE
    void *tmp = 0; size_t bytes = 0;
    GSB_TRY(dev_check(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_len, d_ptr, n + 1, a->stream), "cub scan size"));
    GSB_TRY(dev_malloc(&tmp, bytes));
    int rc = dev_check(cub::DeviceScan::ExclusiveSum(tmp, bytes, d_len, d_ptr, n + 1, a->stream), "cub scan");
    if (!rc) rc = dev_sync(a->stream);
    dev_free(tmp);
    ++g_launches;
    return rc;
#else
    GSB_LAUNCH(k_scan_serial, dim3(1), dim3(1), a->stream, d_len, d_ptr, n);
    return 0;
#endif
}

static void fill_pat_args(const gsb200_assembler *a, const PatchDev &P, PatArgs &A)
{
    A.dim = P.dim; A.ncomp = a->ncomp;
    for (int k = 0; k < 3; ++k) {
        A.n[k] = k < P.dim ? P.dir[k].nfun : 1; A.p[k] = k < P.dim ? P.dir[k].p : 0;
        A.plo[k] = k < P.dim ? P.dir[k].d_plo : 0; A.phi[k] = k < P.dim ? P.dir[k].d_phi : 0;
    }
    A.dofmap = P.d_dofmap; A.nb = P.nb; A.nfree = a->nfree; A.own_lo = P.own_lo; A.own_hi = P.own_hi;
    A.npre = a->d_npre; A.colflag = P.d_colflag; A.st = P.d_st; A.st2 = P.d_st2; A.nrun = P.nrun;
}

static int build_pattern(gsb200_assembler *a)
{
    const int N = a->nfree;
    stream_t s = a->stream;
#ifndef GSB200_EMULATE
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s);
#endif
    unsigned long long *d_len = 0; int *d_cursor = 0; unsigned char *d_gneed = 0;
    struct Guard {          // the temporaries and the timing events go on every return path
        unsigned long long *&len; int *&cursor; unsigned char *&gneed; stream_t s;
#ifndef GSB200_EMULATE
        cudaEvent_t e0, e1;
#endif
        ~Guard() { dev_sync(s); dev_free(len); dev_free(cursor); dev_free(gneed);
#ifndef GSB200_EMULATE
                   cudaEventDestroy(e0); cudaEventDestroy(e1);
#endif
        }
    } guard{d_len, d_cursor, d_gneed, s
#ifndef GSB200_EMULATE
            , e0, e1
#endif
    };
    GSB_TRY(dev_malloc((void **)&d_len, sizeof(unsigned long long) * (size_t)(N + 1)));
    GSB_TRY(dev_memset(d_len, 0, sizeof(unsigned long long) * (size_t)(N + 1), s));
    GSB_TRY(dev_malloc((void **)&d_cursor, sizeof(int) * (size_t)(N + 1)));
    GSB_TRY(dev_memset(d_cursor, 0, sizeof(int) * (size_t)(N + 1), s));
    GSB_TRY(dev_malloc((void **)&d_gneed, (size_t)N + 1));
    GSB_TRY(dev_memset(d_gneed, 0, (size_t)N + 1, s));
    GSB_TRY(dev_sync(s));
    dev_free(a->d_colptr); dev_free(a->d_inner); dev_free(a->d_values); a->d_colptr = 0; a->d_inner = 0; a->d_values = 0;
    GSB_TRY(dev_malloc((void **)&a->d_colptr, sizeof(i64) * (size_t)(N + 1)));
    for (auto &P : a->patches) {
        PatArgs A; fill_pat_args(a, P, A); A.len = d_len; A.colptr = 0; A.inner = 0; A.cursor = d_cursor; A.gneed = d_gneed;
        const i64 nt = P.nb * a->ncomp;
        GSB_LAUNCH(k_pattern<0>, dim3((unsigned)((nt + 127) / 128)), dim3(128), s, A);
    }
    GSB_TRY(scan_lengths(a, d_len, a->d_colptr, N));
    i64 nnz_ub = 0;
    GSB_TRY(dev_d2h(&nnz_ub, a->d_colptr + N, sizeof(i64), s));
    GSB_TRY(dev_malloc((void **)&a->d_inner, sizeof(int) * (size_t)std::max<i64>(nnz_ub, 1)));
    for (auto &P : a->patches) {
        PatArgs A; fill_pat_args(a, P, A); A.len = d_len; A.colptr = a->d_colptr; A.inner = a->d_inner; A.cursor = d_cursor; A.gneed = d_gneed;
        const i64 nt = P.nb * a->ncomp;
#ifndef GSB200_EMULATE
        int maxlen = a->ncomp;
        for (int k = 0; k < P.dim; ++k) maxlen *= 2 * P.dir[k].p + 1;
        const int stride = maxlen | 1;
        const size_t smem = (size_t)32 * stride * sizeof(int);
        if (smem <= 200 * 1024) {      // coalesced fill through per-lane shared-memory rows
            GSB_TRY(grant_dynamic_smem((const void *)k_pattern_staged, 200 * 1024));
            k_pattern_staged<<<(unsigned)((nt + 31) / 32), 32, smem, s>>>(A, stride); note_launch();
        } else
#endif
        GSB_LAUNCH(k_pattern<1>, dim3((unsigned)((nt + 127) / 128)), dim3(128), s, A);
    }
    GSB_LAUNCH(k_pat_sort, dim3((N + 127) / 128), dim3(128), s, N, d_gneed, a->d_colptr, a->d_inner, d_len);
    // coupled columns may have shrunk: rescan and compact
    std::vector<unsigned char> gneed(N + 1);
    GSB_TRY(dev_d2h(gneed.data(), d_gneed, (size_t)N + 1, s));
    bool coupled = false; a->any_generic = false;
    for (int g = 0; g < N; ++g) { if (gneed[g] == 2) coupled = true; }
    if (coupled) {
        i64 *d_newptr = 0; int *d_newinner = 0;
        GSB_TRY(dev_malloc((void **)&d_newptr, sizeof(i64) * (size_t)(N + 1)));
        GSB_TRY(scan_lengths(a, d_len, d_newptr, N));
        i64 nnz = 0;
        GSB_TRY(dev_d2h(&nnz, d_newptr + N, sizeof(i64), s));
        GSB_TRY(dev_malloc((void **)&d_newinner, sizeof(int) * (size_t)std::max<i64>(nnz, 1)));
        GSB_LAUNCH(k_pat_compact, dim3((N + 127) / 128), dim3(128), s, N, a->d_colptr, d_newptr, a->d_inner, d_newinner);
        GSB_TRY(dev_sync(s));
        dev_free(a->d_colptr); dev_free(a->d_inner);
        a->d_colptr = d_newptr; a->d_inner = d_newinner; a->nnz = nnz;
    } else a->nnz = nnz_ub;
    // any column on the generic (search + atomic) path?  then values must start from zero
    for (auto &P : a->patches) {
        std::vector<unsigned char> fl((size_t)P.nb * a->ncomp);
        GSB_TRY(dev_d2h(fl.data(), P.d_colflag, fl.size(), s));
        for (unsigned char f : fl) if (f == 2) { a->any_generic = true; break; }
    }
    for (auto &P : a->patches) {       // packed owner records for the final sweep
        const i64 nt = P.nb * a->ncomp;
        dev_free(P.d_ownrec); P.d_ownrec = 0;
        GSB_TRY(dev_malloc((void **)&P.d_ownrec, sizeof(i64) * (size_t)nt));
        GSB_LAUNCH(k_owner_records, dim3((unsigned)((nt + 127) / 128)), dim3(128), s, nt, N, P.d_dofmap, P.d_colflag, a->d_colptr, P.d_ownrec);
    }
    GSB_TRY(dev_malloc((void **)&a->d_values, sizeof(double) * (size_t)std::max<i64>(a->nnz, 1)));
    GSB_TRY(dev_memset(a->d_values, 0, sizeof(double) * (size_t)std::max<i64>(a->nnz, 1), s));
    GSB_TRY(dev_last_error("pattern kernels"));
    GSB_TRY(dev_sync(s));
#ifndef GSB200_EMULATE
    cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(&a->tm.pattern_ms, e0, e1);
#endif
    a->coupled_off.clear();
    for (int c : a->coupled_runs) { i64 off = 0; GSB_TRY(dev_d2h(&off, a->d_colptr + c, sizeof(i64), s)); a->coupled_off.push_back(off); }
    a->pattern_built = true; a->plan_valid = false; a->spmv_ready = false;
    return 0;
}

static void stage_io(int kind, int stage, int *nin, int *nout)
{
    if (kind == KIND_MASS) { *nin = 1; *nout = 1; return; }
    if (stage == 2) { *nin = 4; *nout = 1; return; }
    if (stage == 0) { *nin = kind == KIND_SYM ? 6 : 9; *nout = kind == KIND_SYM ? 8 : 9; return; }
    if (stage == 1) { *nin = kind == KIND_SYM ? 8 : 9; *nout = 4; return; }
    *nin = kind == KIND_SYM ? 3 : 4; *nout = 4;
}

static int upload_segments(gsb200_assembler *a, const std::vector<int> &seg, size_t *offset_ints)
{
    // segments of all sweeps of one chunk live in one small device buffer, appended
    if ((*offset_ints + seg.size()) > a->seg_cap) { set_error("segment buffer overflow"); return GSB200_EINVAL; }
    if (dry_run()) {
        if (a->seg_host.size() < *offset_ints + seg.size()) a->seg_host.resize(*offset_ints + seg.size());
        std::copy(seg.begin(), seg.end(), a->seg_host.begin() + *offset_ints);
    }
    *offset_ints += seg.size();
    return 0;
}

#ifndef GSB200_EMULATE
static void mark(gsb200_assembler *a, int tag)
{
    if (dry_run()) return;
    if (a->ev_used == a->ev.size()) { cudaEvent_t e; cudaEventCreate(&e); a->ev.push_back(e); a->ev_tag.push_back(tag); }
    a->ev_tag[a->ev_used] = tag;
    cudaEventRecord(a->ev[a->ev_used++], a->stream);
}
#else
static void mark(gsb200_assembler *, int) {}
#endif

static int nseg_for(i64 threads_per_seg, int nfun, int p1)
{
    // enough segments to put ~300k threads in flight, but keep each segment >= 4 spans long
    const i64 target = 300000;
    i64 n = (target + threads_per_seg - 1) / std::max<i64>(threads_per_seg, 1);
    n = std::min<i64>(n, std::max(1, nfun / (4 * p1)));
    return (int)std::max<i64>(1, n);
}

static int assemble_pass(gsb200_assembler *a)
{
    stream_t s = a->stream;
    const int N = a->nfree;
    g_launches = 0;
    a->jit_launches = 0;
    a->ev_used = 0;
    memset(a->tm.sweep_bytes, 0, sizeof a->tm.sweep_bytes); memset(a->tm.sweep_flops, 0, sizeof a->tm.sweep_flops);
    a->tm.nchunks = 0;
    mark(a, 5);
    if (!dry_run()) {
        GSB_TRY(dev_memset(a->d_rhs, 0, sizeof(double) * (size_t)N * a->nrhs, s));
        if (a->any_generic) GSB_TRY(dev_memset(a->d_values, 0, sizeof(double) * (size_t)std::max<i64>(a->nnz, 1), s));
    }

    const int dim = a->dim, L = dim - 1;
    const int kind = a->form == GSB200_FORM_MASS ? KIND_MASS : (a->form == GSB200_FORM_POISSON ? KIND_SYM : KIND_GEN);
    const int nblocks = a->form == GSB200_FORM_ELASTICITY ? a->ncomp * a->ncomp : 1;
    const int nf = a->rhs_kind == GSB200_RHS_PROGRAM ? (int)a->progs.size() : 0;
    int ncD, no1, no2 = 0, tmp;
    if (dim == 3) { stage_io(kind, 0, &ncD, &no1); stage_io(kind, 1, &tmp, &no2); }
    else stage_io(kind, 3, &ncD, &no1);

    // workspace budget
    i64 limit = a->plan_valid ? a->plan_limit : a->ws_limit;
#ifndef GSB200_EMULATE
    if (limit <= 0) { size_t fr = 0, tot = 0; cudaMemGetInfo(&fr, &tot); limit = (i64)((fr + dev_pool_idle() + a->ws_size[0] + a->ws_size[1] + a->ws_size[2] + a->ws_size[3]) * 0.85); }      // (what the pool holds idle is ours to reuse)
#else
    if (limit <= 0) limit = (i64)1 << 30;
#endif
    a->plan_limit = limit;
    // Lanes: the patches of a multi-patch problem are independent up to the atomically summed interface columns, and their kernels
    // are short (config 4: 216 launches of ~0.6 ms per step, 17 % of the warp slots busy): they go round-robin onto up to four
    // streams, each with its own workspace, so that one patch's ramp-up and tail overlap the others' (GSB200_LANES=1: one stream).
    int NL = 1;
#ifndef GSB200_EMULATE
    {
        static const int lanes_env = [] { const char *e = getenv("GSB200_LANES"); return e ? atoi(e) : 4; }();
        int owned = 0; for (auto &P : a->patches) if (P.own_hi > P.own_lo) ++owned;
        NL = std::max(1, std::min(std::min(lanes_env, 4), owned));
        if (NL > 1 && !dry_run()) {
            for (int l = 1; l < NL; ++l) {
                if (!a->lanes[l]) GSB_TRY(dev_check(cudaStreamCreateWithFlags(&a->lanes[l], cudaStreamNonBlocking), "lane stream"));
                if (!a->lane_ev[l]) GSB_TRY(dev_check(cudaEventCreateWithFlags(&a->lane_ev[l], cudaEventDisableTiming), "lane event"));
            }
            if (!a->lane_ev[0]) GSB_TRY(dev_check(cudaEventCreateWithFlags(&a->lane_ev[0], cudaEventDisableTiming), "lane event"));
            GSB_TRY(dev_check(cudaEventRecord(a->lane_ev[0], a->stream), "event record"));      // the zeroed rhs / values are ready
            for (int l = 1; l < NL; ++l) GSB_TRY(dev_check(cudaStreamWaitEvent(a->lanes[l], a->lane_ev[0], 0), "stream wait"));
        }
    }
#endif
    limit /= NL;
    const bool timed = NL == 1;         // per-stage events only make sense on one stream
    int owned_seen = 0;

    size_t segoff = 0;   // every sweep of every chunk gets its own slice of the segment buffer
    for (size_t ip = 0; ip < a->patches.size(); ++ip) {
        PatchDev &P = a->patches[ip];
        if (P.own_hi <= P.own_lo) continue;
        const int lane = owned_seen++ % NL;
#ifndef GSB200_EMULATE
        stream_t s = lane ? a->lanes[lane] : a->stream;
#else
        (void)lane;
#endif
        const Dir1D &d0 = P.dir[0], &d1 = P.dir[1], &dL = P.dir[L];
        const i64 Q0 = d0.Q, Q1 = dim == 3 ? d1.Q : 1;
        const i64 NI0 = (i64)d0.nfun * (2 * d0.p + 1), NI1 = dim == 3 ? (i64)d1.nfun * (2 * d1.p + 1) : 1;
        const i64 n0 = d0.nfun, n1 = dim == 3 ? d1.nfun : 1;
        const i64 W0 = 2 * d0.p + 1, W1 = dim == 3 ? 2 * d1.p + 1 : 1;
        // layout of A1 (3-D): 1 (default) = blocked by last-direction element (A1[o][i0][q1][e2][d0][t]: the second sweep reads whole runs; the
        // first one stores whole rows of deltas at the owner's exit), 2 = A1[o][i0][d0][q1][q2] (coalesced first-sweep stores, the second
        // sweep gathers q-point pieces; measured slower, profiles/r01b_layout_experiments.txt).  GSB200_A1BLK overrides.
        const int a1_env = [] { const char *e = getenv("GSB200_A1BLK"); return e ? atoi(e) : -1; }();      // read at every assembly: the tests switch it
        const int a1_mode = dim != 3 ? 0 : ((a1_env >= 0 && a1_env != 3) ? a1_env : 1);      // (3 = 1 + rows stored for delta >= 0 only, see a1_half)
        // the gathered reads exist in the window kernel only (q = p+1 points); any other rule keeps the blocked layout, which the
        // generic kernels address through the same strides
        const bool a1_gather = a1_mode == 2 && dim == 3 && d1.q == d1.p + 1, a1_blk = a1_mode == 1 || (a1_mode == 2 && !a1_gather);
        // K0 fused into the first sweep (fused.cuh): D and F stay in shared memory.  Needs the p+1-point rule in direction 0 (window
        // accumulators) and a geometry whose last-direction functions per column tile fit the line-coefficient buffer.
        static const bool fuse_env = [] { const char *e = getenv("GSB200_FUSE"); return !e || atoi(e) > 0; }();
        const bool fused = fuse_env && P.d_lc && d0.q == d0.p + 1 && dL.q == d0.q && d0.p >= 1 && d0.p <= 4 && nf <= GSB_FUSE_MAXF;     // (the rows layout is chosen below when the next two sweeps are fused as well)
        // second + last sweep fused (fused23.cuh): A1 in rows layout, A2 never exists.  3-D gradient forms, degrees 1..3, p+1-point rules.
        const bool s23 = fused && dim == 3 && a1_env < 0 && d1.q == d1.p + 1 && dL.q == dL.p + 1 && d1.p == dL.p && s23_available(kind, d1.p + 1);
        // A1 stored for delta >= 0 only, A1[o][q1][e2][i0][d0 = 0..p][t] (terms.cuh T3SymS2U): symmetric 3-D forms at degree 3 with the
        // 4-point rule (a span's points fill a 32-byte sector); the second sweep reads delta < 0 at the mirrored pair, a few rows away
        // in the same (q1, e2) block.
        // Opt-in (GSB200_A1BLK=3): measured on B200 at config 2 the first sweep does not get faster (5.4 ms either way: it is bound by
        // instruction issue and shared-memory traffic, not by its stores) and the second sweep loses 1.1 ms (6.0 vs 4.9 ms: the mirrored
        // 32-byte pieces double the L2 requests per warp): profiles/r02_a1_half_experiment.txt
        const bool a1_half = fused && !s23 && dim == 3 && kind == KIND_SYM && a1_env == 3 && d0.p == 3 && d1.p == 3 && d1.q == 4 && dL.q == 4;
        const i64 A1I0 = a1_half ? (i64)d0.nfun * (d0.p + 1) : NI0;        // (function, stored delta) pairs of direction 0
        const i64 a1_pad = a1_half ? 64 : 0;
        const int nfv = std::max(nf, 1);
        // doubles of workspace per last-direction quadrature point
        i64 perq = (fused ? 0 : ncD * Q0 * Q1 + nf * Q0 * Q1) + no1 * A1I0 * Q1 + (dim == 3 && !s23 ? no2 * NI1 * NI0 : 0) + nfv * n0 * Q1 + (dim == 3 ? n1 * n0 : 0);
        i64 maxpts = limit / (perq * 8);
        const i64 minpts = (i64)(dL.p + 1) * dL.q;
        if (maxpts < minpts) { set_error("workspace limit %lld B too small: one slab of patch %zu needs %lld B", (long long)limit, ip, (long long)(perq * 8 * minpts)); return GSB200_ENOMEM; }
        int x_lo = P.own_lo;
        // streamed delivery (single scalar patch): at most ceil(owned layers / chunks) layers per chunk
        const bool stream_cols = a->plan_chunks > 1 && a->patches.size() == 1 && !P.sufmin.empty();
        const int x_cap = stream_cols ? std::max(dL.p + 1, (P.own_hi - P.own_lo + a->plan_chunks - 1) / a->plan_chunks) : P.own_hi - P.own_lo;
        while (x_lo < P.own_hi) {
            // largest chunk [x_lo,x_hi) whose element footprint fits
            int x_hi = x_lo + 1;
            while (x_hi < P.own_hi && x_hi - x_lo < x_cap && (i64)(dL.flast[x_hi] - dL.ffirst[x_lo] + 1) * dL.q <= maxpts) ++x_hi;
            const int eL0 = dL.ffirst[x_lo], eL1 = dL.flast[x_hi - 1] + 1, ELc = eL1 - eL0;
            const i64 QLc = (i64)ELc * dL.q;
            const size_t need = (size_t)(perq * QLc) * 8 + 6 * 256 + (size_t)a1_pad * 8;
            if (need > a->ws_size[lane]) {
                GSB_TRY(dev_sync(s));
                dev_free(a->ws[lane]); a->ws[lane] = 0; a->ws_size[lane] = 0;
                GSB_TRY(dev_malloc(&a->ws[lane], need)); a->ws_size[lane] = need;
            }
            double *w = (double *)a->ws[lane];
            auto carve = [&](i64 count) { double *p = w; w += (count + 31) / 32 * 32; return p; };   // 256-byte aligned pieces
            double *D = carve(fused ? 0 : ncD * Q0 * Q1 * QLc);
            double *A1 = carve(no1 * A1I0 * Q1 * QLc + a1_pad) + a1_pad;      // (the mirrored reads of the first functions reach a few doubles below)
            double *A2 = carve(dim == 3 && !s23 ? no2 * NI1 * NI0 * QLc : 0);
            double *F = carve(fused ? 0 : nf * Q0 * Q1 * QLc);
            double *V1 = carve(nfv * n0 * Q1 * QLc);
            double *V2 = carve(dim == 3 ? n1 * n0 * QLc : 0);
            const i64 npts = Q0 * Q1 * QLc;
            ++a->tm.nchunks;

            for (int blk = 0; blk < nblocks; ++blk) {
                const int brow = nblocks == 1 ? 0 : blk / a->ncomp, bcol = nblocks == 1 ? 0 : blk % a->ncomp;
                // ---------------- K0 arguments
                GeoArgs G; memset(&G, 0, sizeof G);
                G.dim = dim;
                for (int k = 0; k < dim; ++k) {
                    const Dir1D &d = P.dir[k];
                    G.qn[k] = (k == L) ? (int)QLc : d.Q; G.qoff[k] = (k == L) ? eL0 * d.q : 0;
                    G.gtab[k] = d.d_gtab; G.gfirst[k] = d.d_gfirst; G.pg1[k] = d.pg1; G.ngeo[k] = d.ngeo;
                    G.hpt[k] = d.d_hpt; G.gwp[k] = d.d_gwp;
                }
                G.coefs = P.d_coefs; G.weights = P.d_weights; G.ngeo_total = P.ngeo_total;
                G.form = a->form; G.brow = brow; G.bcol = bcol; G.lambda = a->coef[0]; G.mu = a->coef[1];
                G.symD = kind == KIND_SYM;
                G.D = D; G.dstride = npts;
                const bool with_load = blk == 0 && nf > 0;
                if (with_load) { G.F = F; G.fstride = npts; G.nf = nf; for (int c = 0; c < nf; ++c) G.prog[c] = a->progs[c]; }
                const int pgl = P.dir[L].pg1;
                const bool rat = P.d_weights != 0, hot = a->form == GSB200_FORM_POISSON && G.symD;
                if (!fused) {
                    if (timed) mark(a, 0);
                    static const int gblk = [] { const char *e = getenv("GSB200_GEO_BLOCK"); const int v = e ? atoi(e) : 128; return (v == 64 || v == 256) ? v : 128; }();
                    const dim3 gg((unsigned)((QLc + gblk - 1) / gblk), dim == 3 ? (unsigned)Q1 : (unsigned)Q0, dim == 3 ? (unsigned)Q0 : 1u);
                    if (gg.y > 65535u || gg.z > 65535u) { set_error("more than 65535 quadrature points per direction"); return GSB200_EUNSUPPORTED; }
#ifndef GSB200_EMULATE
#define GSB_GEOL(D_, PG_, R_, F_) { cudaKernel_t jk = (G.F && !dry_run()) ? jit_geometry_kernel(a->progs_host, a->device, D_, PG_, R_, F_) : 0; \
                                    if (jk) { void *kargs[] = {(void *)&G}; GSB_TRY(dev_check(cudaLaunchKernel((const void *)jk, gg, dim3(gblk), kargs, 0, s), "launch of the compiled geometry kernel")); note_launch(); ++a->jit_launches; } \
                                    else { auto kfn = k_geometry_line<D_, PG_, R_, F_>; GSB_LAUNCH(kfn, gg, dim3(gblk), s, G); } }
#else
#define GSB_GEOL(D_, PG_, R_, F_) { auto kfn = k_geometry_line<D_, PG_, R_, F_>; GSB_LAUNCH(kfn, gg, dim3(gblk), s, G); }
#endif
#define GSB_GEOL_D(D_) { if (hot && !rat && pgl == 2) GSB_GEOL(D_, 2, false, 1) else if (hot && !rat && pgl == 3) GSB_GEOL(D_, 3, false, 1) \
                         else if (rat) GSB_GEOL(D_, 0, true, 0) else GSB_GEOL(D_, 0, false, 0) }
                    if (dim == 2) GSB_GEOL_D(2) else GSB_GEOL_D(3)
#undef GSB_GEOL_D
#undef GSB_GEOL
                }

                // ---------------- sweeps
                FinalArgs Fa; memset(&Fa, 0, sizeof Fa);
                Fa.dim = dim; Fa.L = L;
                for (int k = 0; k < 3; ++k) {
                    Fa.n[k] = k < dim ? P.dir[k].nfun : 1; Fa.p[k] = k < dim ? P.dir[k].p : 0;
                    Fa.plo[k] = k < dim ? P.dir[k].d_plo : 0; Fa.phi[k] = k < dim ? P.dir[k].d_phi : 0;
                }
                Fa.dofmap = P.d_dofmap; Fa.nb = P.nb; Fa.brow = brow; Fa.bcol = bcol;
                Fa.colflag = P.d_colflag; Fa.ownrec = P.d_ownrec; Fa.st = P.d_st; Fa.st2 = P.d_st2; Fa.nrun = P.nrun;
                Fa.colptr = a->d_colptr; Fa.inner = a->d_inner; Fa.values = a->d_values;
                Fa.rhs = a->d_rhs; Fa.fixed = a->d_fixed; Fa.nfree = N; Fa.nfixed = a->nfixed; Fa.nrhs = a->nrhs;

                auto base_args = [&](const Dir1D &d) {
                    SweepArgs A; memset(&A, 0, sizeof A);
                    A.first = d.d_first; A.nexit = d.d_nexit; A.tab = d.d_tab; A.tabl = d.d_tabl; A.q = d.q; A.p = d.p; A.fin = Fa;
                    A.out_bq = 1; A.out_od = 1; A.d_off = d.p;
                    A.pf_dist = 0;                                   // (L2 prefetch hints: no gain once cp.async is in, profiles/r01b_layout_experiments.txt)
                    A.wb_stores = (dL.q * 8) % 32 != 0;              // pieces that do not fill 32-byte sectors wait in L2 for their neighbours
                    return A;
                };
                auto account = [&](int slot, const SweepArgs &A, const std::vector<int> &seg, i64 fpp, double nin, double nout, i64 npairs_out) {
                    i64 pts = 0;
                    for (size_t k = 0; k < seg.size(); k += 4) pts += (i64)(seg[k + 1] - seg[k]) * A.q;
                    a->tm.sweep_flops[slot] += fpp * A.ncol * pts;
                    a->tm.sweep_bytes[slot] += (i64)(8.0 * (nin * (double)A.ncol * (double)pts + nout * (double)npairs_out * (double)A.ncol));
                };
                // first sweep (direction 0): fused with K0, or the window / generic kernel reading D
                auto first_sweep = [&](SweepArgs &A, int stage, i64 ncolL, i64 nrows) -> int {
                    const int nseg = nseg_for(A.ncol * (d0.p + 1), d0.nfun, d0.p + 1);
                    std::vector<int> seg = make_segments(d0, 0, d0.nfun, nseg);
                    A.seg = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg, &segoff));
                    if (timed) mark(a, 1);
                    i64 fpp = 0; int nin, nout;
                    stage_io(kind, stage, &nin, &nout);
                    if (fused) {
                        FusedArgs FA; memset(&FA, 0, sizeof FA);
                        FA.G = G; FA.G.D = 0; FA.G.F = 0;
                        FA.first = d0.d_first; FA.nexit = d0.d_nexit; FA.tab = d0.d_tab; FA.seg = A.seg; FA.lc = P.d_lc; FA.lc_nL = dL.ngeo;
                        FA.ncolL = (int)ncolL; FA.nrows = (int)nrows;
                        FA.out = A.out; FA.out_cs = A.out_cs; FA.out_fs = A.out_fs; FA.out_bq = A.out_bq; FA.out_bs = A.out_bs; FA.out_is = A.out_is;
                        FA.d_off = A.d_off; FA.out_ds = A.out_ds; FA.half_rows = A.half_out;
                        FA.nf = with_load ? nf : 0; FA.v1 = V1; FA.v1_fs = A.ncol; FA.v1_cs = n0 * A.ncol;
                        { FusedCtx fc; fc.progs = &a->progs_host; fc.device = a->device; fc.jit_launches = &a->jit_launches;
                          GSB_TRY(launch_fused(fc, kind, dim, d0.p + 1, FA, (int)seg.size() / 4, s, hot, rat, pgl, s23 || a1_gather, &fpp)); }
                        account(0, A, seg, fpp, 0, nout, A.half_out ? A1I0 : NI0);
                    } else {
                        GSB_TRY(dispatch_sweep(kind, stage, d0.p + 1, A, (int)seg.size() / 4, s, &fpp));
                        account(0, A, seg, fpp, nin, nout, NI0);
                    }
                    return 0;
                };
                i64 fpp = 0; int nin, nout;
                if (dim == 3) {
                    {   // S1: direction 0
                        SweepArgs A = base_args(d0);
                        A.in = D; A.in_cs = npts; A.in_es = (i64)d0.q * Q1 * QLc; A.in_ts = Q1 * QLc; A.in_os = 0; A.in_is = 1; A.e_in0 = 0;
                        A.ncol = Q1 * QLc; A.ninner = A.ncol;
                        A.out = A1; A.out_cs = NI0 * Q1 * QLc; A.out_fs = (2 * d0.p + 1) * Q1 * QLc; A.out_ds = Q1 * QLc; A.out_os = 0; A.out_bq = A.ncol + 1; A.out_bs = 0; A.out_is = 1;
                        if (a1_blk && !s23) {    // A1[o][i0][q1][e2][d0][t]: the second sweep then reads AND writes whole (d0, t) runs
                            A.out_fs = Q1 * ELc * W0 * dL.q; A.out_ds = dL.q; A.out_bq = dL.q; A.out_bs = W0 * dL.q; A.out_is = 1;
                        }
                        if (a1_half) {           // A1[o][q1][e2][i0][d0 >= 0][t]
                            const i64 P1q = (i64)(d0.p + 1) * dL.q;
                            A.out_cs = A1I0 * Q1 * QLc; A.out_fs = P1q; A.out_ds = dL.q; A.out_bq = dL.q; A.out_bs = n0 * P1q; A.out_is = 1;
                            A.half_out = 1; A.d_off = 0;
                        }
                        GSB_TRY(first_sweep(A, 0, QLc, Q1));
                    }
                    if (s23) {   // S2 + S3 in one kernel: rows of A2 go from the direction-1 warps to the direction-2 warps through shared memory
                        S23Args SA; memset(&SA, 0, sizeof SA);
                        SA.first1 = d1.d_first; SA.nexit1 = d1.d_nexit; SA.tab1 = d1.d_tab;
                        SA.first2 = dL.d_first; SA.tabl2 = dL.d_tabl; SA.ffirst2 = dL.d_ffirst; SA.flast2 = dL.d_flast;
                        SA.e_in0 = eL0; SA.a1 = A1; SA.a1_cs = NI0 * Q1 * QLc; SA.a1_row = Q1 * QLc; SA.a1_q1 = QLc;
                        SA.n0 = (int)n0; SA.W0 = (int)W0; SA.fin = Fa;
                        // tiles of the last direction: as few as keep every tile within S23_NEMAX spans
                        std::vector<int> tiles; int ne_max = 0;
                        for (int nt = 1;; ++nt) {
                            tiles = make_segments(dL, x_lo, x_hi, nt);
                            ne_max = 0;
                            for (size_t k = 0; k < tiles.size(); k += 4) ne_max = std::max(ne_max, tiles[k + 1] - tiles[k]);
                            if (ne_max <= S23_NEMAX || nt >= x_hi - x_lo) break;
                        }
                        if (ne_max > S23_NEMAX) { set_error("a function of the last direction spans more than %d elements", S23_NEMAX); return GSB200_EUNSUPPORTED; }
                        const i64 ncta = NI0 * (i64)(tiles.size() / 4);
                        int nseg1 = (int)std::max<i64>(1, std::min<i64>((2 * 148 + ncta - 1) / ncta, std::max(1, d1.nfun / (4 * (d1.p + 1)))));
                        std::vector<int> seg1 = make_segments(d1, 0, d1.nfun, nseg1);
                        SA.tiles = a->d_seg + segoff; GSB_TRY(upload_segments(a, tiles, &segoff));
                        SA.seg1 = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg1, &segoff));
                        if (timed) mark(a, 2);
                        i64 fpp2 = 0, fpp3 = 0;
                        const dim3 grid((unsigned)NI0, (unsigned)(tiles.size() / 4), (unsigned)(seg1.size() / 4));
                        GSB_TRY(launch_s23(kind, d1.p + 1, SA, grid, ne_max, s, &fpp2, &fpp3));
                        stage_io(kind, 1, &nin, &nout);
                        {   // accounting: direction-1 work per tile point, direction-2 work per pair and span; bytes: A1 read once per tile span, K written once
                            i64 pts2 = 0; for (size_t k = 0; k < tiles.size(); k += 4) pts2 += (i64)(tiles[k + 1] - tiles[k]) * dL.q;
                            i64 pts1 = 0; for (size_t k = 0; k < seg1.size(); k += 4) pts1 += (i64)(seg1[k + 1] - seg1[k]) * d1.q;
                            a->tm.sweep_flops[1] += fpp2 * NI0 * pts2 * pts1 + fpp3 * NI0 * NI1 * pts2;
                            a->tm.sweep_bytes[1] += (i64)(8.0 * (nin * (double)NI0 * (double)pts2 * (double)pts1 + (double)NI0 * (double)NI1 * (double)(x_hi - x_lo) * (2 * dL.p + 1)));
                        }
                    } else {
                    {   // S2: direction 1
                        SweepArgs A = base_args(d1);
                        A.in = A1; A.in_cs = NI0 * Q1 * QLc; A.in_es = (i64)d1.q * QLc; A.in_ts = QLc; A.in_os = Q1 * QLc; A.in_is = 1; A.e_in0 = 0;
                        A.ncol = NI0 * QLc; A.ninner = QLc;
                        // A2[g][i1][e2][i0][d1][d0][t]: the last sweep then writes (d1,d0)-contiguous runs of each CSC column
                        A.out = A2; A.out_cs = NI1 * ELc * NI0 * dL.q; A.out_fs = (i64)ELc * NI0 * W1 * dL.q; A.out_ds = (i64)W0 * dL.q;
                        A.out_od = W0; A.out_dshift = 0; A.mirror = 0; A.out_nprev = d0.nfun;
                        A.out_os = (i64)W1 * W0 * dL.q; A.out_os2 = dL.q; A.out_bq = dL.q; A.out_bs = NI0 * W1 * dL.q; A.out_is = 1;
                        if (a1_blk) {    // thread = (i0; e2, d0, t): contiguous in A1 and, per element, in A2
                            A.in_os = Q1 * ELc * W0 * dL.q; A.in_is = 1; A.in_ts = ELc * W0 * dL.q; A.in_es = (i64)d1.q * A.in_ts;
                            A.ncol = n0 * ELc * W0 * dL.q; A.ninner = ELc * W0 * dL.q;
                            A.out_od = 1; A.out_dshift = 0; A.out_os = W1 * W0 * dL.q; A.out_os2 = 0; A.out_bq = W0 * dL.q; A.out_bs = NI0 * W1 * dL.q; A.out_is = 1;
                        }
                        if (a1_half) {   // thread = (i0; e2, d0, t); stored rows [q1][e2][i0][d0 >= 0][t]; delta < 0 at the mirrored pair (i0 + d0, -d0)
                            const i64 P1q = (i64)(d0.p + 1) * dL.q;
                            A.in = A1 - (i64)d0.p * dL.q;       // slot index d0 + p in the thread mapping, slot d0 in memory
                            A.in_cs = A1I0 * Q1 * QLc; A.in_ts = ELc * n0 * P1q; A.in_es = (i64)d1.q * A.in_ts; A.in_os = P1q; A.in_is = 1;
                            A.in_bq2 = W0 * dL.q; A.in_bs2 = n0 * P1q; A.in_bq = dL.q; A.in_bs = dL.q;
                            A.mir_q = dL.q; A.mir_w = (int)W0; A.mir_p = d0.p; A.mir_n = (int)n0;
                        }
                        if (a1_gather) { // thread = (i0; e2, d0, t) as above, but A1 keeps whole q2 rows per (i0, d0)
                            A.in_os = W0 * Q1 * QLc; A.in_is = 1; A.in_ts = QLc; A.in_es = (i64)d1.q * QLc;
                            A.in_bq2 = W0 * dL.q; A.in_bs2 = dL.q; A.in_bq = dL.q; A.in_bs = Q1 * QLc;
                            A.ncol = n0 * ELc * W0 * dL.q; A.ninner = ELc * W0 * dL.q;
                            A.out_od = 1; A.out_dshift = 0; A.out_os = W1 * W0 * dL.q; A.out_os2 = 0; A.out_bq = W0 * dL.q; A.out_bs = NI0 * W1 * dL.q; A.out_is = 1;
                        }
                        const int nseg = nseg_for(A.ncol * (d1.p + 1), d1.nfun, d1.p + 1);
                        std::vector<int> seg = make_segments(d1, 0, d1.nfun, nseg);
                        A.seg = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg, &segoff));
                        if (timed) mark(a, 2);
                        GSB_TRY(dispatch_sweep(kind, a1_half ? 4 : 1, d1.p + 1, A, (int)seg.size() / 4, s, &fpp));
                        stage_io(kind, 1, &nin, &nout);
                        account(1, A, seg, fpp, a1_half ? nin * (double)(d0.p + 1) / (double)W0 : nin, nout, NI1);      // (every stored value is read once from HBM)
                    }
                    {   // S3: direction 2, scatter into the CSC arrays
                        SweepArgs A = base_args(dL);
                        A.in = A2; A.in_cs = NI1 * ELc * NI0 * dL.q; A.in_es = NI0 * W1 * dL.q; A.in_ts = 1; A.in_os = (i64)ELc * NI0 * W1 * dL.q; A.in_is = dL.q; A.e_in0 = eL0;
                        A.ncol = NI1 * NI0; A.ninner = NI0 * W1;     // outer = i1, inner = (i0, d1, d0)
                        const int nseg = nseg_for(A.ncol, x_hi - x_lo, dL.p + 1);
                        std::vector<int> seg = make_segments(dL, x_lo, x_hi, nseg);
                        A.seg = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg, &segoff));
                        if (timed) mark(a, 3);
                        GSB_TRY(dispatch_sweep(kind, 2, dL.p + 1, A, (int)seg.size() / 4, s, &fpp));
                        stage_io(kind, 2, &nin, &nout); account(2, A, seg, fpp, nin, nout, (i64)(x_hi - x_lo) * (2 * dL.p + 1));
                    }
                    }
                } else {
                    {   // S1: direction 0
                        SweepArgs A = base_args(d0);
                        A.in = D; A.in_cs = npts; A.in_es = (i64)d0.q * QLc; A.in_ts = QLc; A.in_os = 0; A.in_is = 1; A.e_in0 = 0;
                        A.ncol = QLc; A.ninner = QLc;
                        A.out = A1; A.out_cs = (i64)ELc * NI0 * dL.q; A.out_fs = (i64)(2 * d0.p + 1) * dL.q; A.out_ds = dL.q; A.out_os = 0; A.out_bq = dL.q; A.out_bs = NI0 * dL.q; A.out_is = 1;
                        GSB_TRY(first_sweep(A, 3, QLc, 1));
                    }
                    {   // S2: direction 1, scatter
                        SweepArgs A = base_args(dL);
                        A.in = A1; A.in_cs = (i64)ELc * NI0 * dL.q; A.in_es = NI0 * dL.q; A.in_ts = 1; A.in_os = 0; A.in_is = dL.q; A.e_in0 = eL0;
                        A.ncol = NI0; A.ninner = NI0;
                        const int nseg = nseg_for(A.ncol, x_hi - x_lo, dL.p + 1);
                        std::vector<int> seg = make_segments(dL, x_lo, x_hi, nseg);
                        A.seg = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg, &segoff));
                        if (timed) mark(a, 2);
                        GSB_TRY(dispatch_sweep(kind, 2, dL.p + 1, A, (int)seg.size() / 4, s, &fpp));
                        stage_io(kind, 2, &nin, &nout); account(1, A, seg, fpp, nin, nout, (i64)(x_hi - x_lo) * (2 * dL.p + 1));
                    }
                }
                // ---------------- K3: load vector (once per chunk)
                if (with_load) {
                    if (timed) mark(a, 4);
                    for (int c = 0; c < nf; ++c) {
                        double *V1c = V1 + (fused ? (i64)c * n0 * Q1 * QLc : 0);      // the fused first sweep produced all components at once
                        const int rcol = a->form == GSB200_FORM_ELASTICITY ? 0 : c;   // rhs column
                        const int comp = a->form == GSB200_FORM_ELASTICITY ? c : 0;   // dof component
                        VSweepArgs V; memset(&V, 0, sizeof V);
                        auto vbase = [&](const Dir1D &d) { V.ffirst = d.d_ffirst; V.flast = d.d_flast; V.first = d.d_first; V.tab = d.d_tab; V.q = d.q; V.p1 = d.p + 1; };
                        // window kernel when the rule has p+1 points (reads its input once), else one thread per (column, function)
                        auto vlaunch = [&](const Dir1D &d) -> int {
                            const int P1 = d.p + 1, nx = V.x_hi - V.x_lo;
                            if (d.q != P1 || P1 < 2 || P1 > 5) {
                                GSB_LAUNCH(k_vsweep, dim3((unsigned)((V.ncol + 127) / 128), nx), dim3(128), s, V);
                                return 0;
                            }
                            const int nseg = std::min(nseg_for(V.ncol, nx, P1), 65535);
                            std::vector<int> seg = make_segments(d, V.x_lo, V.x_hi, nseg);
                            const int *dseg = a->d_seg + segoff; GSB_TRY(upload_segments(a, seg, &segoff));
                            const dim3 grid((unsigned)((V.ncol + 127) / 128), (unsigned)(seg.size() / 4));
                            switch (P1) {
                            case 2: GSB_LAUNCH(k_vsweepw<2>, grid, dim3(128), s, V, d.d_tabl, d.d_nexit, dseg); break;
                            case 3: GSB_LAUNCH(k_vsweepw<3>, grid, dim3(128), s, V, d.d_tabl, d.d_nexit, dseg); break;
                            case 4: GSB_LAUNCH(k_vsweepw<4>, grid, dim3(128), s, V, d.d_tabl, d.d_nexit, dseg); break;
                            default: GSB_LAUNCH(k_vsweepw<5>, grid, dim3(128), s, V, d.d_tabl, d.d_nexit, dseg); break;
                            }
                            return 0;
                        };
                        if (dim == 3) {
                            if (!fused) {
                                vbase(d0); V.final_ = 0; V.x_lo = 0; V.x_hi = d0.nfun; V.e_in0 = 0;
                                V.in = F + c * npts; V.in_qs = Q1 * QLc; V.in_os = 0; V.in_is = 1; V.ncol = Q1 * QLc; V.ninner = V.ncol;
                                V.out = V1c; V.out_fs = Q1 * QLc; V.out_os = 0; V.out_is = 1;
                                GSB_TRY(vlaunch(d0));
                            }
                            vbase(d1); V.final_ = 0; V.e_in0 = 0; V.x_lo = 0; V.x_hi = d1.nfun;
                            V.in = V1c; V.in_qs = QLc; V.in_os = Q1 * QLc; V.in_is = 1; V.ncol = n0 * QLc; V.ninner = QLc;
                            V.out = V2; V.out_fs = n0 * QLc; V.out_os = QLc; V.out_is = 1;
                            GSB_TRY(vlaunch(d1));
                            vbase(dL); V.x_lo = x_lo; V.x_hi = x_hi; V.e_in0 = eL0; V.final_ = 1;
                            V.in = V2; V.in_qs = 1; V.in_os = n0 * QLc; V.in_is = QLc; V.ncol = n1 * n0; V.ninner = n0;
                            V.n0 = (int)n0; V.n1 = (int)n1; V.dimlow = 2; V.dofmap = P.d_dofmap + (i64)comp * P.nb; V.rhs = a->d_rhs + (i64)rcol * N; V.nfree = N;
                            GSB_TRY(vlaunch(dL));
                        } else {
                            if (!fused) {
                                vbase(d0); V.final_ = 0; V.x_lo = 0; V.x_hi = d0.nfun; V.e_in0 = 0;
                                V.in = F + c * npts; V.in_qs = QLc; V.in_os = 0; V.in_is = 1; V.ncol = QLc; V.ninner = QLc;
                                V.out = V1c; V.out_fs = QLc; V.out_os = 0; V.out_is = 1;
                                GSB_TRY(vlaunch(d0));
                            }
                            vbase(dL); V.x_lo = x_lo; V.x_hi = x_hi; V.e_in0 = eL0; V.final_ = 1;
                            V.in = V1c; V.in_qs = 1; V.in_os = 0; V.in_is = QLc; V.ncol = n0; V.ninner = n0;
                            V.n0 = (int)n0; V.n1 = 1; V.dimlow = 1; V.dofmap = P.d_dofmap + (i64)comp * P.nb; V.rhs = a->d_rhs + (i64)rcol * N; V.nfree = N;
                            GSB_TRY(vlaunch(dL));
                        }
                    }
                }
            }
            if (stream_cols) {
                if (dry_run()) a->chunk_cols.push_back(x_hi < P.own_hi ? P.sufmin[x_hi] : N);
#ifndef GSB200_EMULATE
                else {
                    const size_t k = (size_t)a->tm.nchunks - 1;
                    while (a->chunk_ev.size() <= k) { cudaEvent_t e; GSB_TRY(dev_check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event")); a->chunk_ev.push_back(e); }
                    GSB_TRY(dev_check(cudaEventRecord(a->chunk_ev[k], s), "event record"));
                }
#endif
            }
            x_lo = x_hi;
        }
    }
#ifndef GSB200_EMULATE
    if (NL > 1 && !dry_run())
        for (int l = 1; l < NL; ++l) {
            GSB_TRY(dev_check(cudaEventRecord(a->lane_ev[l], a->lanes[l]), "event record"));
            GSB_TRY(dev_check(cudaStreamWaitEvent(a->stream, a->lane_ev[l], 0), "stream wait"));
        }
#endif
    // ---------------- Neumann boundary load (rank 0 of a multi-rank run adds it once per side it owns)
    for (const auto &ns : a->neumann) {
        PatchDev &P = a->patches[ns.patch];
        const int dir = (ns.side - 1) / 2, upper = (ns.side - 1) % 2;
        // the side belongs to the rank owning the boundary function layer (last direction) or, for other
        // directions, it is split by the same slab ownership through the dof rows: keep it simple and exact:
        // only the rank that owns the patch's first slab assembles side loads
        if (!(P.own_lo == 0 && P.own_hi > 0)) continue;
        FaceArgs F; memset(&F, 0, sizeof F);
        FaceLoadArgs Lg; memset(&Lg, 0, sizeof Lg);
        F.dim = dim; F.dir = dir; F.upper = upper; Lg.dim = dim; Lg.dir = dir;
        i64 npt = 1, nfn = 1;
        for (int k = 0; k < dim; ++k) {
            const Dir1D &d = P.dir[k];
            F.qn[k] = d.Q; F.gtab[k] = d.d_gtab; F.gfirst[k] = d.d_gfirst; F.pg1[k] = d.pg1; F.ngeo[k] = d.ngeo; F.hpt[k] = d.d_hpt; F.gwp[k] = d.d_gwp;
            Lg.nfun[k] = d.nfun; Lg.p1[k] = d.p + 1; Lg.q[k] = d.q; Lg.Q[k] = d.Q; Lg.ffirst[k] = d.d_ffirst; Lg.flast[k] = d.d_flast; Lg.tab[k] = d.d_tab;
            if (k != dir) { npt *= d.Q; nfn *= d.nfun; }
        }
        const Dir1D &dd = P.dir[dir];
        for (int k2 = 0; k2 < dd.pg1; ++k2) F.bgeo[k2] = dd.bgeo[upper][k2];
        F.bgfirst = dd.bgfirst[upper];
        for (int k2 = 0; k2 <= dd.p; ++k2) Lg.bval[k2] = dd.bval[upper][k2];
        Lg.bfirst = dd.bfirst[upper]; Lg.nb1 = dd.p + 1;
        F.coefs = P.d_coefs; F.weights = P.d_weights; F.ngeo_total = P.ngeo_total;
        F.ndata = ns.ndata; for (int c = 0; c < ns.ndata; ++c) F.prog[c] = ns.prog[c];
        if ((size_t)npt > a->face_cap) { GSB_TRY(dev_sync(s)); dev_free(a->d_face); a->d_face = 0; GSB_TRY(dev_malloc((void **)&a->d_face, sizeof(double) * (size_t)npt)); a->face_cap = (size_t)npt; }
        F.Fb = a->d_face; Lg.Fb = a->d_face; Lg.dofmap = P.d_dofmap; Lg.rhs = a->d_rhs; Lg.nfree = N;
        if (dim == 2) { GSB_LAUNCH(k_face_geometry<2>, dim3((unsigned)((npt + 127) / 128)), dim3(128), s, F); GSB_LAUNCH(k_face_load<2>, dim3((unsigned)((nfn + 127) / 128)), dim3(128), s, Lg); }
        else { GSB_LAUNCH(k_face_geometry<3>, dim3((unsigned)((npt + 127) / 128)), dim3(128), s, F); GSB_LAUNCH(k_face_load<3>, dim3((unsigned)((nfn + 127) / 128)), dim3(128), s, Lg); }
    }
    mark(a, 6);
    if (dry_run()) return 0;
    GSB_TRY(dev_last_error("assembly kernels"));
    a->tm.launches = g_launches;
    a->assembled = true;
    return 0;
}

// The launch sequence (chunks, segments, workspace) is static for a built pattern: the first
// call walks it once without launching to size the workspace and upload every segment table in
// one copy; afterwards assemble() only enqueues kernels, so the stream never waits on the host.
static int assemble(gsb200_assembler *a)
{
    if (a->plan_valid && a->plan_chunks != a->deliver_chunks) a->plan_valid = false;
    if (!a->plan_valid) {
        a->plan_chunks = a->deliver_chunks;
        g_dry = true;
        a->seg_host.clear(); a->chunk_cols.clear(); a->chunk_off.clear();
        const int rc = assemble_pass(a);
        g_dry = false;
        if (rc) return rc;
        if (!a->seg_host.empty()) GSB_TRY(dev_h2d(a->d_seg, a->seg_host.data(), a->seg_host.size() * sizeof(int), a->stream));
        GSB_TRY(dev_sync(a->stream));
        for (int c : a->chunk_cols) {       // value offsets behind which everything is final once the chunk is done
            i64 off = 0; GSB_TRY(dev_d2h(&off, a->d_colptr + c, sizeof(i64), a->stream));
            a->chunk_off.push_back(off);
        }
        a->plan_valid = true;
    }
    return assemble_pass(a);
}

static void finish_timings(gsb200_assembler *a)
{
#ifndef GSB200_EMULATE
    gsb200_timings &t = a->tm;
    t.geometry_ms = t.rhs_ms = t.total_ms = 0; for (int k = 0; k < 3; ++k) t.sweep_ms[k] = 0;
    for (size_t i = 0; i + 1 < a->ev_used; ++i) {
        float ms = 0; cudaEventElapsedTime(&ms, a->ev[i], a->ev[i + 1]);
        const int tag = a->ev_tag[i];
        if (tag == 0) t.geometry_ms += ms; else if (tag >= 1 && tag <= 3) t.sweep_ms[tag - 1] += ms; else if (tag == 4) t.rhs_ms += ms;
    }
    if (a->ev_used >= 2) cudaEventElapsedTime(&t.total_ms, a->ev.front(), a->ev[a->ev_used - 1]);
#else
    (void)a;
#endif
}


#ifndef GSB200_EMULATE
// ------------------------------------------------------------------ device -> host delivery
// Page-locked destinations (cudaHostAlloc / cudaHostRegister) are written by the copy engine directly.  Pageable ones (an Eigen
// matrix: gsSparseMatrix::valuePtr()) go through a ring of pinned staging buffers that host threads drain while the next chunk is in
// flight: PCIe rate instead of the driver's single-threaded bounce copy.
static const size_t STAGE_CHUNK = (size_t)32 << 20;
static bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
static int staged_d2h(gsb200_assembler *a, void *dst_, const void *src_, size_t bytes, cudaStream_t cs)
{
    const int NB = 4;
    for (int k = 0; k < NB; ++k) {
        if (!a->stage[k]) GSB_TRY(dev_check(cudaHostAlloc(&a->stage[k], STAGE_CHUNK, cudaHostAllocDefault), "cudaHostAlloc (staging)"));
        if (!a->stage_ev[k]) GSB_TRY(dev_check(cudaEventCreateWithFlags(&a->stage_ev[k], cudaEventDisableTiming), "event"));
    }
    char *dst = (char *)dst_; const char *src = (const char *)src_;
    const size_t nchunk = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const int T = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
    std::unique_ptr<std::atomic<int>[]> ready(new std::atomic<int>[nchunk]), done(new std::atomic<int>[nchunk]);
    for (size_t i = 0; i < nchunk; ++i) { ready[i].store(0); done[i].store(0); }
    std::atomic<int> failed(0);
    std::vector<std::thread> workers;
    for (int t = 0; t < T; ++t)
        workers.emplace_back([&, t] {
            for (size_t i = 0; i < nchunk; ++i) {
                while (!ready[i].load(std::memory_order_acquire)) { if (failed.load()) return; std::this_thread::yield(); }
                const size_t off = i * STAGE_CHUNK, n = std::min(STAGE_CHUNK, bytes - off);
                const size_t lo = n * t / T, hi = n * (t + 1) / T;
                memcpy(dst + off + lo, (const char *)a->stage[i % NB] + lo, hi - lo);
                done[i].fetch_add(1, std::memory_order_release);
            }
        });
    size_t issued = 0, completed = 0; int rc = 0;
    while (completed < nchunk && !rc) {
        while (issued < nchunk && issued < completed + NB && (issued < (size_t)NB || done[issued - NB].load(std::memory_order_acquire) == T)) {
            const size_t off = issued * STAGE_CHUNK, n = std::min(STAGE_CHUNK, bytes - off);
            rc = dev_check(cudaMemcpyAsync(a->stage[issued % NB], src + off, n, cudaMemcpyDeviceToHost, cs), "D2H (staged)");
            if (!rc) rc = dev_check(cudaEventRecord(a->stage_ev[issued % NB], cs), "event record");
            if (rc) break;
            ++issued;
        }
        if (rc) break;
        if (issued == completed) { std::this_thread::yield(); continue; }       // the ring is full of chunks the workers still drain
        rc = dev_check(cudaEventSynchronize(a->stage_ev[completed % NB]), "event sync");
        if (!rc) { ready[completed].store(1, std::memory_order_release); ++completed; }
    }
    if (rc) failed.store(1);
    for (auto &w : workers) w.join();
    return rc;
}
// asynchronous for pinned destinations (caller synchronises `cs`), complete on return for pageable ones
static int d2h_any(gsb200_assembler *a, void *dst, const void *src, size_t bytes, cudaStream_t cs)
{
    if (!bytes) return 0;
    if (host_is_pinned(dst) || bytes < ((size_t)1 << 20)) return dev_check(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cs), "D2H");
    if (bytes >= ((size_t)64 << 20)) {
        // a freshly allocated multi-GB destination (Eigen's value / index arrays) is first touched by the copy: ask for transparent huge
        // pages on its 2 MB-aligned interior, 512 times fewer page faults where the system allows it (a hint: failure is ignored)
        const uintptr_t lo = ((uintptr_t)dst + ((uintptr_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1), hi = ((uintptr_t)dst + bytes) & ~(((uintptr_t)2 << 20) - 1);
        if (hi > lo) madvise((void *)lo, (size_t)(hi - lo), MADV_HUGEPAGE);
    }
    return staged_d2h(a, dst, src, bytes, cs);
}
#endif

} // namespace gsb

namespace gsb {
// a reverse-polish program (source term, boundary data, exact solution) on the device; short ones ride inside the kernel arguments
static int upload_device_program(gsb200_assembler *a, const gsb200_program &pr, DevProgram *out)
{
    if (pr.nops < 1 || pr.nops > GSB200_PROGRAM_MAX_OPS || !pr.ops) { set_error("source/boundary program has bad length %d", pr.nops); return GSB200_EINVAL; }
    std::vector<int> ops(pr.ops, pr.ops + pr.nops); std::vector<double> cs(pr.consts, pr.consts + pr.nconsts);
    int *d_ops = 0; double *d_cs = 0;
    GSB_TRY(upload(&d_ops, ops, a->stream)); a->prog_bufs.push_back(d_ops);
    GSB_TRY(upload(&d_cs, cs, a->stream)); a->prog_bufs.push_back(d_cs);
    DevProgram dp; memset(&dp, 0, sizeof dp); dp.ops = d_ops; dp.consts = d_cs; dp.nops = pr.nops;
    if (pr.nops <= GSB_INLINE_OPS && pr.nconsts <= GSB_INLINE_CONSTS) {
        dp.inl = 1;
        for (int k = 0; k < pr.nops; ++k) dp.iops[k] = (signed char)pr.ops[k];
        for (int k = 0; k < pr.nconsts; ++k) dp.iconsts[k] = pr.consts[k];
    }
    *out = dp;
    return 0;
}
} // namespace gsb

#include "consumer.cuh"

// ====================================================================== C ABI
extern "C" {

const char *gsb200_last_error(void) { return g_err; }
int gsb200_abi_version(void) { return GSB200_ABI_VERSION; }

int gsb200_device_count(int *count)
{
    if (!count) return GSB200_EINVAL;
#ifndef GSB200_EMULATE
    int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0; *count = n;
#else
    *count = 1;
#endif
    return GSB200_OK;
}

int gsb200_create(const gsb200_problem *pb, int device, gsb200_assembler **out)
{
    if (!pb || !out) { set_error("create: null argument"); return GSB200_EINVAL; }
    *out = 0;
    if (pb->abi_version != GSB200_ABI_VERSION) { set_error("ABI version mismatch: caller %d, library %d", pb->abi_version, GSB200_ABI_VERSION); return GSB200_EINVAL; }
    if (pb->npatches < 1 || !pb->patches || pb->nfree < 0 || pb->nfixed < 0 || pb->nrhs < 1) { set_error("create: malformed problem sizes"); return GSB200_EINVAL; }
    const int dim = pb->patches[0].space.dim;
    if (dim != 2 && dim != 3) { set_error("parametric dimension %d unsupported (2 or 3)", dim); return GSB200_EUNSUPPORTED; }
    if (pb->form != GSB200_FORM_POISSON && pb->form != GSB200_FORM_ELASTICITY && pb->form != GSB200_FORM_MASS) { set_error("unknown form %d", pb->form); return GSB200_EUNSUPPORTED; }
    const int ncomp_expected = pb->form == GSB200_FORM_ELASTICITY ? dim : 1;
    if (pb->ncomp != ncomp_expected) { set_error("form %d needs ncomp=%d, got %d", pb->form, ncomp_expected, pb->ncomp); return GSB200_EINVAL; }
    if (pb->form == GSB200_FORM_ELASTICITY && pb->nrhs != 1) { set_error("elasticity supports one right-hand side"); return GSB200_EUNSUPPORTED; }
    if (pb->rhs_kind == GSB200_RHS_SAMPLES) { set_error("sampled source terms are not implemented yet"); return GSB200_EUNSUPPORTED; }
    if (pb->nranks < 1 || pb->rank < 0 || pb->rank >= pb->nranks) { set_error("bad rank/nranks"); return GSB200_EINVAL; }
    GSB_TRY(select_device(device));

    gsb200_assembler *a = new gsb200_assembler();
    a->device = device; a->dim = dim; a->form = pb->form; a->ncomp = pb->ncomp; a->nfree = pb->nfree; a->nfixed = pb->nfixed;
    a->nrhs = pb->nrhs; a->rhs_kind = pb->rhs_kind; a->rank = pb->rank; a->nranks = pb->nranks;
    memcpy(a->coef, pb->coef, sizeof a->coef);
    memset(&a->tm, 0, sizeof a->tm);
    int rc = 0;
    a->patches.resize(pb->npatches);
    for (int ip = 0; ip < pb->npatches && !rc; ++ip) {
        const gsb200_patch &S = pb->patches[ip];
        PatchDev &P = a->patches[ip];
        if (S.space.dim != dim || S.geo.dim != dim) { set_error("patch %d: mixed dimensions", ip); rc = GSB200_EINVAL; break; }
        P.dim = dim; P.nb = 1; P.ngeo_total = 1;
        for (int k = 0; k < dim && !rc; ++k) {
            const int p = S.space.degree[k], gp = S.geo.degree[k];
            if (p < 1 || p > 4) { set_error("patch %d: degree %d unsupported (1..4)", ip, p); rc = GSB200_EUNSUPPORTED; break; }
            if (gp < 1 || gp > GSB_MAXP) { set_error("patch %d: geometry degree %d unsupported", ip, gp); rc = GSB200_EUNSUPPORTED; break; }
            if (!S.space.knots[k] || !S.geo.knots[k]) { set_error("patch %d: null knots", ip); rc = GSB200_EINVAL; break; }
            const int q = num_nodes(pb->quA, pb->quB, p);
            if (q < 1 || q > 16) { set_error("quadrature size %d unsupported", q); rc = GSB200_EUNSUPPORTED; break; }
            rc = build_dir(P.dir[k], S.space.knots[k], S.space.nknots[k], p, q, S.geo.knots[k], S.geo.nknots[k], gp, a->stream);
            if (rc) break;
            P.nb *= P.dir[k].nfun; P.ngeo_total *= P.dir[k].ngeo;
        }
        if (rc) break;
        if (!S.dofmap || !S.geo_coefs) { set_error("patch %d: null dofmap/coefs", ip); rc = GSB200_EINVAL; break; }
        std::vector<int> dm(S.dofmap, S.dofmap + P.nb * pb->ncomp);
        for (int g : dm) if (g < 0 || g >= pb->nfree + pb->nfixed) { set_error("patch %d: dof index %d out of range", ip, g); rc = GSB200_EINVAL; break; }
        if (rc) break;
        std::vector<double> cf(S.geo_coefs, S.geo_coefs + P.ngeo_total * dim);
        if ((rc = upload(&P.d_dofmap, dm, a->stream))) break;
        if ((rc = upload(&P.d_coefs, cf, a->stream))) break;
        if (S.geo_weights) { std::vector<double> wv(S.geo_weights, S.geo_weights + P.ngeo_total); if ((rc = upload(&P.d_weights, wv, a->stream))) break; }
        const int L = dim - 1;
        P.nrun = 1; for (int k = 1; k < dim; ++k) P.nrun *= 2 * P.dir[k].p + 1;
        {   // fused first sweep (fused.cuh)
            const Dir1D &dl = P.dir[L];
            // leading directions of the geometry contracted once per patch (they do not depend on the last direction)
            const int nfg = S.geo_weights ? dim + 1 : dim;
            const i64 lc_count = (i64)(dim == 3 ? P.dir[1].Q : 1) * P.dir[0].Q * dl.ngeo * nfg * dim;
            if (lc_count * 8 <= ((i64)1 << 30)) {      // a geometry as fine as the solution basis in the last direction would not pay: unfused path
                if ((rc = dev_malloc((void **)&P.d_lc, sizeof(double) * (size_t)lc_count))) break;
                LineCoefArgs LC; memset(&LC, 0, sizeof LC);
                LC.dim = dim; LC.rational = S.geo_weights ? 1 : 0; LC.Q0 = P.dir[0].Q; LC.Q1 = dim == 3 ? P.dir[1].Q : 1; LC.nL = dl.ngeo;
                for (int k = 0; k < dim - 1; ++k) { LC.gtab[k] = P.dir[k].d_gtab; LC.gfirst[k] = P.dir[k].d_gfirst; LC.pg1[k] = P.dir[k].pg1; LC.ngeo[k] = P.dir[k].ngeo; }
                LC.coefs = P.d_coefs; LC.weights = P.d_weights; LC.ngeo_total = P.ngeo_total; LC.lc = P.d_lc;
                const i64 nthreads = lc_count / dim;
                GSB_LAUNCH(k_line_coefs, dim3((unsigned)((nthreads + 127) / 128)), dim3(128), a->stream, LC);
            }
        }
        if ((rc = dev_malloc((void **)&P.d_colflag, (size_t)P.nb * pb->ncomp))) break;
        if ((rc = dev_memset(P.d_colflag, 0, (size_t)P.nb * pb->ncomp, a->stream))) break;
        if ((rc = dev_malloc((void **)&P.d_st, sizeof(unsigned) * (size_t)P.nb * pb->ncomp * P.nrun))) break;
        if (pb->npatches > 1 && (rc = dev_malloc((void **)&P.d_st2, sizeof(unsigned) * (size_t)P.nb * pb->ncomp * P.nrun))) break;
        // ownership: one patch -> slabs along the last direction; several -> whole patches, balanced by element count below
        const int nL = P.dir[L].nfun;
        if (pb->npatches == 1) { P.own_lo = (int)((i64)nL * pb->rank / pb->nranks); P.own_hi = (int)((i64)nL * (pb->rank + 1) / pb->nranks); }
        else { P.own_lo = 0; P.own_hi = nL; }
        if (pb->npatches == 1 && pb->ncomp == 1) {     // streamed delivery: which columns are final behind a layer of the last direction
            const i64 per = P.nb / nL;
            P.sufmin.assign((size_t)nL + 1, pb->nfree);
            for (int x = nL - 1; x >= 0; --x) {
                int m = P.sufmin[x + 1];
                for (i64 i = (i64)x * per; i < (i64)(x + 1) * per; ++i) if (dm[i] < m) m = dm[i];     // eliminated DOFs are >= nfree
                P.sufmin[x] = m;
            }
        }
    }
    if (!rc) {
        // several patches: longest-processing-time-first assignment (21 patches of yeti_mp2 over 8 GPUs), identical on every rank
        const int np = pb->npatches;
        a->patch_owner.assign(np, 0);
        if (np > 1) {
            std::vector<i64> cost(np); std::vector<int> order(np);
            for (int ip = 0; ip < np; ++ip) { i64 c = 1; for (int k = 0; k < dim; ++k) c *= a->patches[ip].dir[k].nel * (a->patches[ip].dir[k].p + 1); cost[ip] = c; order[ip] = ip; }
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
            std::vector<i64> load(pb->nranks, 0);
            for (int ip : order) { int best = 0; for (int r = 1; r < pb->nranks; ++r) if (load[r] < load[best]) best = r; a->patch_owner[ip] = best; load[best] += cost[ip]; }
            for (int ip = 0; ip < np; ++ip) if (a->patch_owner[ip] != pb->rank) a->patches[ip].own_hi = 0;
        }
        // reference columns of the regular-stencil SpMV; runs of coupled columns (more than one pre-image over all patches)
        std::vector<unsigned char> cnt((size_t)pb->nfree + 1, 0);
        a->mid_dof.assign((size_t)np * pb->ncomp, -1);
        for (int ip = 0; ip < np; ++ip) {
            const PatchDev &P = a->patches[ip]; const int *dmp = pb->patches[ip].dofmap;
            for (i64 i = 0; i < P.nb * pb->ncomp; ++i) { const int g = dmp[i]; if (g < pb->nfree && cnt[g] < 2) ++cnt[g]; }
            if (P.own_hi > P.own_lo) {
                i64 mid = 0, stride = 1;
                for (int k = 0; k < dim; ++k) { const int n = P.dir[k].nfun; const int m = k == dim - 1 ? (P.own_lo + P.own_hi) / 2 : n / 2; mid += stride * m; stride *= n; }
                for (int c = 0; c < pb->ncomp; ++c) { const int g = dmp[(i64)c * P.nb + mid]; a->mid_dof[(size_t)ip * pb->ncomp + c] = g < pb->nfree ? g : -1; }
            }
        }
        if (np > 1) for (int g = 0; g < pb->nfree; ) {
            if (cnt[g] < 2) { ++g; continue; }
            int e = g; while (e < pb->nfree && cnt[e] >= 2) ++e;
            a->coupled_runs.push_back(g); a->coupled_runs.push_back(e); g = e;
        }
    }
    if (!rc && pb->fixed) { std::vector<double> fx(pb->fixed, pb->fixed + (size_t)pb->nfixed * pb->nrhs); rc = upload(&a->d_fixed, fx, a->stream); }
    auto upload_program = [&](const gsb200_program &pr, DevProgram *out) -> int { return upload_device_program(a, pr, out); };
    if (!rc && pb->rhs_kind == GSB200_RHS_PROGRAM) {
        const int np = pb->form == GSB200_FORM_ELASTICITY ? pb->ncomp : pb->nrhs;
        if (!pb->rhs_programs) { set_error("rhs_kind=PROGRAM but no programs"); rc = GSB200_EINVAL; }
        for (int c = 0; c < np && !rc; ++c) {
            DevProgram dp; rc = upload_program(pb->rhs_programs[c], &dp);
            if (!rc) {
                a->progs.push_back(dp);
                HostProgram hp; hp.ops.assign(pb->rhs_programs[c].ops, pb->rhs_programs[c].ops + pb->rhs_programs[c].nops);
                hp.consts.assign(pb->rhs_programs[c].consts, pb->rhs_programs[c].consts + pb->rhs_programs[c].nconsts);
                a->progs_host.push_back(hp);
            }
        }
    }
    if (!rc && pb->nneumann > 0) {
        if (pb->ncomp != 1 || pb->nrhs != 1 || !pb->neumann) { set_error("Neumann sides need a scalar problem with one right-hand side"); rc = GSB200_EUNSUPPORTED; }
        for (int i = 0; i < pb->nneumann && !rc; ++i) {
            const gsb200_neumann &nm = pb->neumann[i];
            if (nm.patch < 0 || nm.patch >= pb->npatches || nm.side < 1 || nm.side > 2 * dim || (nm.ndata != 1 && nm.ndata != dim)) {
                set_error("Neumann side %d malformed (patch %d, side %d, ndata %d)", i, nm.patch, nm.side, nm.ndata); rc = GSB200_EINVAL; break; }
            gsb200_assembler::NeumannSide ns; ns.patch = nm.patch; ns.side = nm.side; ns.ndata = nm.ndata;
            for (int c = 0; c < nm.ndata && !rc; ++c) rc = upload_program(nm.data[c], &ns.prog[c]);
            if (!rc) a->neumann.push_back(ns);
        }
    }
    if (!rc) rc = dev_malloc((void **)&a->d_rhs, sizeof(double) * (size_t)std::max(1, pb->nfree) * pb->nrhs);
    if (!rc) rc = dev_malloc((void **)&a->d_npre, sizeof(int) * (size_t)(pb->nfree + 1));
    if (!rc) rc = dev_memset(a->d_npre, 0, sizeof(int) * (size_t)(pb->nfree + 1), a->stream);
    if (!rc) for (auto &P : a->patches) {
        const i64 nt = P.nb * pb->ncomp;
        GSB_LAUNCH(k_pat_preimages, dim3((unsigned)((nt + 127) / 128)), dim3(128), a->stream, P.d_dofmap, nt, pb->nfree, a->d_npre);
    }
    a->seg_cap = 1 << 20;
    if (!rc) rc = dev_malloc((void **)&a->d_seg, a->seg_cap * sizeof(int));
    if (!rc) rc = dev_sync(a->stream);
    if (rc) { delete a; return rc; }
    *out = a;
    return GSB200_OK;
}

void gsb200_destroy(gsb200_assembler *a)
{
    if (!a) return;
    select_device(a->device);
    delete a;               // buffers go back to the library's pool (recycled by the next assembler); gsb200_trim returns them to the driver
}

int gsb200_trim(int device)
{
    GSB_TRY(select_device(device));
    dev_trim();
    return GSB200_OK;
}

int gsb200_set_stream(gsb200_assembler *a, void *cuda_stream)
{
    if (!a) return GSB200_EINVAL;
#ifndef GSB200_EMULATE
    a->stream = (cudaStream_t)cuda_stream;
#else
    (void)cuda_stream;
#endif
    return GSB200_OK;
}

int gsb200_set_workspace_limit(gsb200_assembler *a, int64_t bytes) { if (!a) return GSB200_EINVAL; a->ws_limit = bytes; a->plan_valid = false; return GSB200_OK; }

int gsb200_build_pattern(gsb200_assembler *a)
{
    if (!a) { set_error("null assembler"); return GSB200_EINVAL; }
    GSB_TRY(select_device(a->device));
    return build_pattern(a);
}

int gsb200_assemble(gsb200_assembler *a)
{
    if (!a) { set_error("null assembler"); return GSB200_EINVAL; }
    if (!a->pattern_built) { set_error("gsb200_assemble called before gsb200_build_pattern"); return GSB200_ESTATE; }
    GSB_TRY(select_device(a->device));
    a->deliver_chunks = 1;      // device-resident result: no column ranges to send ahead
    return assemble(a);
}

int gsb200_synchronize(gsb200_assembler *a)
{
    if (!a) return GSB200_EINVAL;
    GSB_TRY(dev_sync(a->stream));
    finish_timings(a);
    return GSB200_OK;
}

int gsb200_nnz(const gsb200_assembler *a, int64_t *nnz)
{
    if (!a || !nnz) return GSB200_EINVAL;
    if (!a->pattern_built) { set_error("pattern not built"); return GSB200_ESTATE; }
    *nnz = a->nnz; return GSB200_OK;
}

int gsb200_device_view_get(const gsb200_assembler *a, gsb200_device_view *v)
{
    if (!a || !v) return GSB200_EINVAL;
    if (!a->pattern_built) { set_error("pattern not built"); return GSB200_ESTATE; }
    gsb200_assembler *am = const_cast<gsb200_assembler *>(a);
    GSB_TRY(select_device(a->device));
    GSB_TRY(spmv_prepare(am));      // extent of the stored columns
    v->nnz = a->nnz; v->ncols = a->nfree; v->col_begin = a->own_c0; v->col_end = a->own_c1;
    v->outer = (const int64_t *)a->d_colptr; v->inner = a->d_inner; v->values = a->d_values; v->rhs = a->d_rhs;
    return GSB200_OK;
}

int gsb200_timings_get(const gsb200_assembler *a, gsb200_timings *t)
{
    if (!a || !t) return GSB200_EINVAL;
    *t = a->tm; return GSB200_OK;
}

int gsb200_download_csc(gsb200_assembler *a, int32_t *outer, int32_t *inner, double *values)
{
    if (!a || !outer || !inner || !values) { set_error("download: null argument"); return GSB200_EINVAL; }
    if (!a->assembled) { set_error("download before assemble"); return GSB200_ESTATE; }
    if (a->nnz > 2147483647LL) { set_error("nnz = %lld exceeds the 32-bit index_t of gsSparseMatrix; use the device view", (long long)a->nnz); return GSB200_ERANGE; }
    std::vector<i64> ptr((size_t)a->nfree + 1);
    GSB_TRY(dev_d2h(ptr.data(), a->d_colptr, ptr.size() * sizeof(i64), a->stream));
    for (size_t i = 0; i < ptr.size(); ++i) outer[i] = (int32_t)ptr[i];
    GSB_TRY(dev_d2h(inner, a->d_inner, sizeof(int) * (size_t)a->nnz, a->stream));
    GSB_TRY(dev_d2h(values, a->d_values, sizeof(double) * (size_t)a->nnz, a->stream));
    finish_timings(a);
    return GSB200_OK;
}

int gsb200_download_rhs(gsb200_assembler *a, double *rhs)
{
    if (!a || !rhs) { set_error("download: null argument"); return GSB200_EINVAL; }
    if (!a->assembled) { set_error("download before assemble"); return GSB200_ESTATE; }
    return dev_d2h(rhs, a->d_rhs, sizeof(double) * (size_t)a->nfree * a->nrhs, a->stream);
}

int gsb200_download_pattern(gsb200_assembler *a, int32_t *outer, int32_t *inner)
{
    if (!a || !outer || !inner) { set_error("download_pattern: null argument"); return GSB200_EINVAL; }
    if (!a->pattern_built) { set_error("pattern not built"); return GSB200_ESTATE; }
    if (a->nnz > 2147483647LL) { set_error("nnz = %lld exceeds the 32-bit index_t of gsSparseMatrix; use the device view", (long long)a->nnz); return GSB200_ERANGE; }
    GSB_TRY(select_device(a->device));
#ifndef GSB200_EMULATE
    const int N = a->nfree;
    if (!a->copy_stream) GSB_TRY(dev_check(cudaStreamCreateWithFlags(&a->copy_stream, cudaStreamNonBlocking), "copy stream"));
    if (!a->d_outer32) GSB_TRY(dev_malloc((void **)&a->d_outer32, sizeof(int) * (size_t)(N + 1)));
    cudaStream_t cs = a->copy_stream;
    k_narrow_outer<<<(N + 1 + 255) / 256, 256, 0, cs>>>(N + 1, a->d_colptr, a->d_outer32); note_launch();
    GSB_TRY(d2h_any(a, outer, a->d_outer32, sizeof(int) * (size_t)(N + 1), cs));
    GSB_TRY(d2h_any(a, inner, a->d_inner, sizeof(int) * (size_t)a->nnz, cs));
    return dev_check(cudaStreamSynchronize(cs), "copy stream sync");
#else
    std::vector<i64> ptr((size_t)a->nfree + 1);
    GSB_TRY(dev_d2h(ptr.data(), a->d_colptr, ptr.size() * sizeof(i64), a->stream));
    for (size_t i = 0; i < ptr.size(); ++i) outer[i] = (int32_t)ptr[i];
    return dev_d2h(inner, a->d_inner, sizeof(int) * (size_t)a->nnz, a->stream);
#endif
}

int gsb200_set_fixed(gsb200_assembler *a, const double *fixed)
{
    if (!a) { set_error("null assembler"); return GSB200_EINVAL; }
    GSB_TRY(select_device(a->device));
    const size_t n = (size_t)a->nfixed * a->nrhs;
    if (!fixed) { GSB_TRY(dev_sync(a->stream)); dev_free(a->d_fixed); a->d_fixed = 0; return GSB200_OK; }
    if (!a->d_fixed) GSB_TRY(dev_malloc((void **)&a->d_fixed, sizeof(double) * std::max<size_t>(n, 1)));
    if (n) GSB_TRY(dev_h2d(a->d_fixed, fixed, sizeof(double) * n, a->stream));
    return dev_sync(a->stream);        // the caller's buffer may go away
}

// values (+ rhs) of a fresh assembly into host memory; with_pattern: the index arrays travel on the copy stream meanwhile
static int assemble_deliver(gsb200_assembler *a, int32_t *outer, int32_t *inner, double *values, double *rhs)
{
    if (a->nnz > 2147483647LL) { set_error("nnz = %lld exceeds the 32-bit index_t of gsSparseMatrix; use the device view", (long long)a->nnz); return GSB200_ERANGE; }
    GSB_TRY(select_device(a->device));
    {   // read at every call: the tests switch them
        const char *e = getenv("GSB200_DELIVER_CHUNKS"), *m = getenv("GSB200_DELIVER_MIN_NNZ");
        const int want = e ? atoi(e) : 8; const i64 min_nnz = m ? atoll(m) : ((i64)1 << 25);
        a->deliver_chunks = (a->patches.size() == 1 && !a->patches[0].sufmin.empty() && a->nnz >= min_nnz) ? std::max(1, want) : 1;
    }
#ifndef GSB200_EMULATE
    const int N = a->nfree;
    if (!a->copy_stream) GSB_TRY(dev_check(cudaStreamCreateWithFlags(&a->copy_stream, cudaStreamNonBlocking), "copy stream"));
    if (!a->ev_done) GSB_TRY(dev_check(cudaEventCreateWithFlags(&a->ev_done, cudaEventDisableTiming), "event"));
    cudaStream_t cs = a->copy_stream;
    // the kernels are enqueued first (asynchronous), so the index arrays travel WHILE the values are being integrated; for a single
    // scalar patch the last direction is cut into chunks and the columns a chunk completes travel while the next chunks integrate
    GSB_TRY(assemble(a));
    GSB_TRY(dev_check(cudaEventRecord(a->ev_done, a->stream), "event record"));
    if (outer && inner) {
        if (!a->d_outer32) GSB_TRY(dev_malloc((void **)&a->d_outer32, sizeof(int) * (size_t)(N + 1)));
        k_narrow_outer<<<(N + 1 + 255) / 256, 256, 0, cs>>>(N + 1, a->d_colptr, a->d_outer32); note_launch();
        GSB_TRY(d2h_any(a, outer, a->d_outer32, sizeof(int) * (size_t)(N + 1), cs));
        GSB_TRY(d2h_any(a, inner, a->d_inner, sizeof(int) * (size_t)a->nnz, cs));
    }
    i64 sent = 0;
    if (a->plan_chunks > 1 && a->chunk_off.size() == (size_t)a->tm.nchunks)
        for (size_t k = 0; k + 1 < a->chunk_off.size(); ++k) {
            const i64 upto = a->chunk_off[k];
            if (upto <= sent) continue;
            GSB_TRY(dev_check(cudaStreamWaitEvent(cs, a->chunk_ev[k], 0), "stream wait"));
            GSB_TRY(d2h_any(a, values + sent, a->d_values + sent, sizeof(double) * (size_t)(upto - sent), cs));
            sent = upto;
        }
    GSB_TRY(dev_check(cudaStreamWaitEvent(cs, a->ev_done, 0), "stream wait"));
    GSB_TRY(d2h_any(a, values + sent, a->d_values + sent, sizeof(double) * (size_t)(a->nnz - sent), cs));
    if (rhs) GSB_TRY(d2h_any(a, rhs, a->d_rhs, sizeof(double) * (size_t)N * a->nrhs, cs));
    GSB_TRY(dev_check(cudaStreamSynchronize(cs), "copy stream sync"));
    GSB_TRY(dev_sync(a->stream));
    finish_timings(a);
    return GSB200_OK;
#else
    GSB_TRY(assemble(a));
    if (outer && inner) GSB_TRY(gsb200_download_pattern(a, outer, inner));
    GSB_TRY(dev_d2h(values, a->d_values, sizeof(double) * (size_t)a->nnz, a->stream));
    return rhs ? gsb200_download_rhs(a, rhs) : GSB200_OK;
#endif
}

int gsb200_assemble_to_host(gsb200_assembler *a, int32_t *outer, int32_t *inner, double *values, double *rhs)
{
    if (!a || !outer || !inner || !values) { set_error("assemble_to_host: null argument"); return GSB200_EINVAL; }
    if (!a->pattern_built) { set_error("gsb200_assemble_to_host called before gsb200_build_pattern"); return GSB200_ESTATE; }
    return assemble_deliver(a, outer, inner, values, rhs);
}

int gsb200_assemble_values_to_host(gsb200_assembler *a, double *values, double *rhs)
{
    if (!a || !values) { set_error("assemble_values_to_host: null argument"); return GSB200_EINVAL; }
    if (!a->pattern_built) { set_error("gsb200_assemble_values_to_host called before gsb200_build_pattern"); return GSB200_ESTATE; }
    return assemble_deliver(a, 0, 0, values, rhs);
}

int gsb200_assemble_host(const gsb200_problem *pb, int device, int64_t *nnz, int32_t *outer, int32_t *inner, double *values, double *rhs)
{
    // stateless: the size query builds (and drops) the pattern; callers that want to keep it use an explicit handle
    // (gsb200_create / gsb200_build_pattern / gsb200_nnz / gsb200_assemble_to_host / gsb200_destroy), as the gismo shims do
    if (!pb || !nnz) { set_error("assemble_host: null argument"); return GSB200_EINVAL; }
    gsb200_assembler *a = 0;
    GSB_TRY(gsb200_create(pb, device, &a));
    int rc = gsb200_build_pattern(a);
    if (!rc) {
        const bool query = !outer || !inner || !values;
        if (!query && *nnz != 0 && *nnz != a->nnz) { set_error("assemble_host: buffers sized for nnz = %lld, the problem has %lld", (long long)*nnz, (long long)a->nnz); rc = GSB200_EINVAL; }
        *nnz = a->nnz;
        if (!rc && !query) rc = gsb200_assemble_to_host(a, outer, inner, values, rhs);
    }
    gsb200_destroy(a);
    return rc;
}

int gsb200_host_pin(void *p, int64_t bytes)
{
#ifndef GSB200_EMULATE
    if (!p || bytes <= 0) return GSB200_EINVAL;
    return dev_check(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault), "cudaHostRegister");
#else
    (void)p; (void)bytes; return GSB200_OK;
#endif
}
int gsb200_host_unpin(void *p)
{
#ifndef GSB200_EMULATE
    if (!p) return GSB200_EINVAL;
    return dev_check(cudaHostUnregister(p), "cudaHostUnregister");
#else
    (void)p; return GSB200_OK;
#endif
}

int gsb200_jit_launches(const gsb200_assembler *a, int *count)
{
    if (!a || !count) return GSB200_EINVAL;
    *count = a->jit_launches; return GSB200_OK;
}

int gsb200_jit_compile_check(const gsb200_program *progs, int nprogs, int dim, int pgl, int rational, int fspec, char *log, int log_cap)
{
    if (!progs || nprogs < 1 || nprogs > 3) { set_error("jit check: bad arguments"); return GSB200_EINVAL; }
#ifndef GSB200_EMULATE
    std::vector<HostProgram> hp(nprogs);
    for (int c = 0; c < nprogs; ++c) { hp[c].ops.assign(progs[c].ops, progs[c].ops + progs[c].nops); hp[c].consts.assign(progs[c].consts, progs[c].consts + progs[c].nconsts); }
    std::string src, clog; std::vector<char> cubin;
    if (!jit_build_source(hp, dim, pgl, rational != 0, fspec, src)) { set_error("jit check: program not translatable"); return GSB200_EUNSUPPORTED; }
    const bool ok = jit_compile(src, cubin, clog);
    if (log && log_cap > 0) { snprintf(log, (size_t)log_cap, "%s", clog.c_str()); }
    if (!ok) { set_error("jit check: NVRTC compile failed: %.800s", clog.c_str()); return GSB200_EUNSUPPORTED; }
    return GSB200_OK;
#else
    (void)dim; (void)pgl; (void)rational; (void)fspec; (void)log; (void)log_cap;
    set_error("no NVRTC in the interpreter build"); return GSB200_EUNSUPPORTED;
#endif
}

int gsb200_expr_eval_host(const gsb200_program *prog, double x, double y, double z, double *out)
{
    if (!prog || !out) return GSB200_EINVAL;
    // same interpreter source as the device (program_eval is host+device)
    DevProgram dp; memset(&dp, 0, sizeof dp); dp.ops = prog->ops; dp.consts = prog->consts; dp.nops = prog->nops;
#ifndef GSB200_EMULATE
    *out = program_eval(dp, x, y, z);
#else
    *out = program_eval(dp, x, y, z);
#endif
    return GSB200_OK;
}

} // extern "C"

#ifndef GSB200_EMULATE
#include "peaks.cuh"
#else
extern "C" int gsb200_measure_peaks(int, double *, double *, double *) { gsb::set_error("no device in the interpreter build"); return GSB200_ENODEVICE; }
#endif
