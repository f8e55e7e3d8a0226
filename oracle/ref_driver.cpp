/* ref_driver.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin C-callable driver around the UNMODIFIED reference (gismo/gismo v24.08.0), compiled
 * header-only from the sources where they lie under /root/reference by oracle/Makefile
 * into oracle/_ref/libgsref.so.  It builds one of the synthetic configurations of
 * SURVEY.md 8(d), runs the reference's own CPU assemblers
 *   path 0: gsPoissonAssembler<real_t>::assemble()           (gsPoissonAssembler.hpp:41-78)
 *   path 1: gsExprAssembler<real_t>::assemble(expr...)       (gsExprAssembler.h:754-833)
 * and hands back (a) the assembled CSC matrix + rhs and (b) the flattened POD problem
 * (gismo_b200/host/gsB200Flatten.h) so that tests can feed the very same inputs to the
 * C restatement (oracle/gsb_oracle.c) and to the CUDA path.  It is also the CPU baseline
 * ("kind": "reference") timed by bench.py.
 */
#include <gismo.h>
#include <gsAssembler/gsVisitorPoisson.h>
#include "../gismo_b200/host/gsB200Flatten.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace gismo;

extern "C" {

typedef struct gsref_config {
    int32_t dim;          /* 2 or 3                                                         */
    int32_t degree;       /* setDegree(degree) on the solution basis                        */
    int32_t nelem;        /* uniformRefine(nelem-1): elements per original span/direction   */
    int32_t geometry;     /* 0 unit box, 1 curved box, 2 NURBS quarter annulus (2D),
                             3 grid of boxes (multi-patch), 4 XML multipatch file,
                             5 complete BVP file as read by examples/poisson2_example.cpp:43-64
                               (geometry id 0, source id 1, boundary conditions id 2), nelem = #refinements */
    int32_t grid[3];      /* patches per direction for geometry 3                           */
    int32_t path;         /* 0 visitor (gsPoissonAssembler), 1 expression (gsExprAssembler)  */
    int32_t form;         /* 0 Poisson, 1 linear elasticity (path 1 only), 2 mass: path 0 = gsGenericAssembler::assembleMass
                             (gsVisitorMass), path 1 = u*u.tr()*meas(G)                                            */
    int32_t dir_values;   /* 100 homogeneous, 101 interpolation, 102 L2 projection          */
    int32_t threads;      /* OpenMP threads for the assembly (0 = leave default)            */
    int32_t degree_elevate; /* geometry 4: degreeElevate instead of setDegree when >0       */
    const char *rhs[3];   /* source term component expressions                              */
    const char *dir[3];   /* Dirichlet data component expressions                           */
    const char *xml;      /* geometry 4: file with a gsMultiPatch                           */
    double lambda, mu;
    int32_t neumann_mask; /* bit s set: boundary sides with index s (1..6) carry Neumann data `neu`  */
    int32_t neu_n;        /* 1: scalar flux (visitor path), dim: vector data dotted with the outer normal */
    const char *neu[3];
    int32_t degree_dir[3]; /* >0: degree of the solution basis in that direction (mixed degrees), applied after setDegree */
    int32_t nrhs;          /* path 0, form 0: right-hand-side columns (components of the source function), 0/1 = one */
    const char *exact;     /* path 1, form 0: exact solution; the system is solved (SimplicialLDLT) and the error norms are integrated
                              as in examples/poisson2_example.cpp:174-177 (gsExprEvaluator::integral)                  */
} gsref_config;

struct gsref_result {
    b200::gsB200Problem flat;
    gsSparseMatrix<real_t> K;
    gsMatrix<real_t> rhs;
    double assemble_seconds;
    int64_t elements, qpoints;
    std::vector<int32_t> neumann;   // (patch, side) pairs
    std::vector<std::string> rhs_text, neu_text;
    std::string err;
    gsMatrix<real_t> sol;           // solution coefficients (free DOFs) when `exact` was given
    double norms[4];                // int (u_ex-u_h)^2, int |grad(u_ex-u_h)|^2, int u_h^2, int |grad u_h|^2
};

static std::string g_err;
const char *gsref_last_error(void) { return g_err.c_str(); }

static gsMultiPatch<real_t> make_geometry(const gsref_config &c)
{
    gsMultiPatch<real_t> mp;
    switch (c.geometry) {
    case 0:
        if (c.dim == 2) mp.addPatch(gsNurbsCreator<>::BSplineSquare(1, 0, 0));
        else            mp.addPatch(gsNurbsCreator<>::BSplineCube(1, 0, 0, 0));
        break;
    case 1: {
        gsGeometry<real_t>::uPtr g;
        if (c.dim == 2) g = gsNurbsCreator<>::BSplineSquare(1, 0, 0);
        else            g = gsNurbsCreator<>::BSplineCube(1, 0, 0, 0);
        g->degreeElevate(1);
        g->uniformRefine(1);
        /* deterministic perturbation of the interior control points */
        gsMatrix<real_t> &C = g->coefs();
        const int n1 = 4; /* p=2, 2 elements -> 4 functions per direction */
        for (index_t i = 0; i < C.rows(); ++i) {
            int id = i; bool interior = true;
            for (int k = 0; k < c.dim; ++k) { const int ik = id % n1; id /= n1; if (ik == 0 || ik == n1 - 1) interior = false; }
            if (interior)
                for (int k = 0; k < c.dim; ++k) C(i, k) += 0.06 * std::sin(1.3 * i + 0.7 * k + 0.3);
        }
        mp.addPatch(give(g));
        break; }
    case 2:
        mp.addPatch(gsNurbsCreator<>::NurbsQuarterAnnulus(1, 2));
        break;
    case 3:
        if (c.dim == 2) mp = gsNurbsCreator<>::BSplineSquareGrid(c.grid[0], c.grid[1], 1.0);
        else            mp = gsNurbsCreator<>::BSplineCubeGrid(c.grid[0], c.grid[1], c.grid[2], 1.0);
        break;
    case 4: {
        gsFileData<real_t> fd(c.xml);
        fd.getFirst(mp);
        break; }
    case 5: {
        gsFileData<real_t> fd(c.xml);
        fd.getId(0, mp);
        return mp;   // topology comes from the file
    }
    default: GISMO_ERROR("unknown geometry kind");
    }
    mp.computeTopology();
    return mp;
}

void *gsref_run(const gsref_config *cfg)
{
    gsref_result *R = new gsref_result();
    try {
        const gsref_config &c = *cfg;
#ifdef _OPENMP
        if (c.threads > 0) omp_set_num_threads(c.threads);
#endif
        gsMultiPatch<real_t> mp = make_geometry(c);
        const int d = mp.parDim();
        gsMultiBasis<real_t> mb(mp, true);
        if (c.geometry == 5) {           // poisson2_example.cpp:69-81: setDegree(max + elevate), r x uniformRefine()
            mb.setDegree(mb.maxCwiseDegree() + c.degree_elevate);
            for (int r = 0; r < c.nelem; ++r) mb.uniformRefine();
        }
        else if (c.geometry == 4 && c.degree_elevate > 0) mb.degreeElevate(c.degree_elevate);
        else mb.setDegree(c.degree);
        for (int k = 0; k < d; ++k)      // mixed degrees: raise single directions
            if (c.degree_dir[k] > c.degree) mb.degreeElevate(c.degree_dir[k] - c.degree, k);
        if (c.geometry != 5 && c.nelem > 1) mb.uniformRefine(c.nelem - 1);

        const int ncomp = c.form == 1 ? d : 1;
        const int nfun = (c.form == 0 && c.path == 0 && c.nrhs > 1) ? c.nrhs : ncomp;     // components of the source / Dirichlet functions
        std::vector<std::string> fs, gs;
        for (int k = 0; k < nfun; ++k) { fs.push_back(c.rhs[k] ? c.rhs[k] : "0"); gs.push_back(c.dir[k] ? c.dir[k] : "0"); }
        gsFunctionExpr<real_t> f(fs, d), g(gs, d);

        std::vector<std::string> ns;
        for (int k = 0; k < std::max(1, c.neu_n); ++k) ns.push_back(c.neu[k] ? c.neu[k] : "0");
        gsFunctionExpr<real_t> gN(ns, d);
        gsBoundaryConditions<real_t> bc;
        for (gsMultiPatch<real_t>::const_biterator it = mp.bBegin(); it != mp.bEnd(); ++it) {
            if ((c.neumann_mask >> it->side().index()) & 1) {
                bc.addCondition(*it, condition_type::neumann, &gN, 0, false, -1);
                R->neumann.push_back(it->patch); R->neumann.push_back(it->side().index());
            } else bc.addCondition(*it, condition_type::dirichlet, &g, 0, false, -1);
        }
        if (c.geometry == 5) {           // source term and boundary conditions from the file
            gsFileData<real_t> fd(c.xml);
            fd.getId(1, f);
            bc = gsBoundaryConditions<real_t>();
            fd.getId(2, bc);
            R->neumann.clear();
            for (gsBoundaryConditions<real_t>::const_iterator it = bc.neumannSides().begin(); it != bc.neumannSides().end(); ++it) {
                R->neumann.push_back(it->patch()); R->neumann.push_back(it->side().index());
                const gsFunctionExpr<real_t> *fe = dynamic_cast<const gsFunctionExpr<real_t> *>(it->function().get());
                if (R->neu_text.empty() && fe) for (short_t k = 0; k < fe->targetDim(); ++k) R->neu_text.push_back(fe->expression(k));
            }
        } else if (c.neumann_mask) for (size_t k = 0; k < ns.size(); ++k) R->neu_text.push_back(ns[k]);
        for (short_t k = 0; k < f.targetDim(); ++k) R->rhs_text.push_back(f.expression(k));
        bc.setGeoMap(mp);

        gsStopwatch timer;
        gsOptionList opt;
        if (c.path == 0 && c.form == 2) {
            // gsGenericAssembler::assembleMass -> gsVisitorMass (gsGenericAssembler.hpp:37-60, gsVisitorMass.h:30-157); the load
            // vector of the same space comes from assembleMoments (gsVisitorMoments.h)
            gsGenericAssembler<real_t> A(mp, mb, gsGenericAssembler<real_t>::defaultOptions(), &bc);
            A.options().setInt("DirichletValues", c.dir_values);
            timer.restart();
            A.assembleMass();
            R->assemble_seconds = timer.stop();
            R->K = A.matrix();          // gsVisitorMass stores both triangles (gsVisitorMass.h:137-139)
            A.assembleMoments(f);
            R->rhs = A.rhs();
            opt = A.options();
            gsMatrix<real_t> nofixed;
            b200::flatten(mp, mb, A.system().colMapper(0), 1, nofixed, opt, GSB200_FORM_MASS, R->flat);
        } else if (c.path == 0) {
            GISMO_ENSURE(c.form == 0, "visitor path: Poisson or mass");
            gsPoissonAssembler<real_t> A(mp, mb, bc, f, dirichlet::elimination, iFace::glue);
            A.options().setInt("DirichletValues", c.dir_values);
            timer.restart();
            A.assemble();
            R->assemble_seconds = timer.stop();
            R->K = A.matrix();
            R->rhs = A.rhs();
            opt = A.options();
            b200::flatten(mp, mb, A.system().colMapper(0), 1, A.fixedDofs(0), opt, GSB200_FORM_POISSON, R->flat);
        } else {
            gsExprAssembler<real_t> A(1, 1);
            typedef gsExprAssembler<real_t>::geometryMap geometryMap;
            typedef gsExprAssembler<real_t>::space space;
            A.setIntegrationElements(mb);
            geometryMap G = A.getMap(mp);
            space u = A.getSpace(mb, ncomp);
            auto ff = A.getCoeff(f, G);
            u.setup(bc, (dirichlet::values)c.dir_values, 0);
            A.initSystem();
            timer.restart();
            if (c.form == 2)
                A.assemble(u * u.tr() * meas(G), u * ff * meas(G));
            else if (c.form == 0)
            {
                A.assemble(igrad(u, G) * igrad(u, G).tr() * meas(G), u * ff * meas(G));
                if (c.neumann_mask || c.geometry == 5) { auto g_N = A.getBdrFunction(G); A.assembleBdr(bc.get("Neumann"), u * g_N.tr() * nv(G)); }
            } else {
                auto pj = ijac(u, G);
                auto bl = c.lambda * idiv(u, G) * idiv(u, G).tr() * meas(G);
                auto bm = c.mu * ((pj.cwisetr() + pj) % pj.tr()) * meas(G);
                A.assemble(bl + bm, u * ff * meas(G));
            }
            R->assemble_seconds = timer.stop();
            R->K = A.matrix();
            R->rhs = A.rhs();
            opt = A.options();
            if (c.exact && c.form == 0) {
                gsSparseSolver<real_t>::SimplicialLDLT solver;
                solver.compute(A.matrix());
                R->sol = solver.solve(A.rhs());
                gsFunctionExpr<real_t> ex(c.exact, d);
                gsExprEvaluator<real_t> ev(A);
                auto u_sol = A.getSolution(u, R->sol);
                auto u_ex = ev.getVariable(ex, G);
                R->norms[0] = ev.integral((u_ex - u_sol).sqNorm() * meas(G));
                R->norms[1] = ev.integral((igrad(u_ex) - igrad(u_sol, G)).sqNorm() * meas(G));
                R->norms[2] = ev.integral(u_sol.sqNorm() * meas(G));
                R->norms[3] = ev.integral(igrad(u_sol, G).sqNorm() * meas(G));
            }
            b200::flatten(mp, mb, u.mapper(), ncomp, u.fixedPart(), opt,
                          c.form == 0 ? GSB200_FORM_POISSON : (c.form == 2 ? GSB200_FORM_MASS : GSB200_FORM_ELASTICITY), R->flat);
            R->flat.pb.coef[0] = c.lambda; R->flat.pb.coef[1] = c.mu;
        }
        R->flat.pb.nrhs = (c.path == 0 && c.form == 0) ? std::max(1, c.nrhs) : 1;
        R->K.makeCompressed();
        R->elements = 0; R->qpoints = 0;
        for (size_t k = 0; k != mb.nBases(); ++k) {
            const int64_t ne = mb.basis(k).numElements();
            int64_t q = 1;
            for (int i = 0; i < d; ++i) q *= (int64_t)(opt.askReal("quA", 1.0) * mb.basis(k).degree(i) + opt.askInt("quB", 1) + 0.5);
            R->elements += ne; R->qpoints += ne * q;
        }
    } catch (std::exception &e) {
        g_err = e.what();
        delete R;
        return NULL;
    }
    return R;
}

void gsref_free(void *h) { delete static_cast<gsref_result *>(h); }

/* sizes[0..7] = nfree, nfixed, nnz, npatches, ncomp, dim, elements, qpoints */
int gsref_sizes(void *h, int64_t *sizes, double *seconds, double *quA, int32_t *quB)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const gsb200_problem &pb = R->flat.pb;
    sizes[0] = pb.nfree; sizes[1] = pb.nfixed; sizes[2] = R->K.nonZeros(); sizes[3] = pb.npatches;
    sizes[4] = pb.ncomp; sizes[5] = pb.patches[0].space.dim; sizes[6] = R->elements; sizes[7] = R->qpoints;
    *seconds = R->assemble_seconds; *quA = pb.quA; *quB = pb.quB;
    return 0;
}

int gsref_csc(void *h, int32_t *outer, int32_t *inner, double *values, double *rhs, double *fixed)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const index_t n = R->K.cols(), nnz = R->K.nonZeros();
    std::copy(R->K.outerIndexPtr(), R->K.outerIndexPtr() + n + 1, outer);
    std::copy(R->K.innerIndexPtr(), R->K.innerIndexPtr() + nnz, inner);
    std::copy(R->K.valuePtr(), R->K.valuePtr() + nnz, values);
    std::copy(R->rhs.data(), R->rhs.data() + R->rhs.size(), rhs);
    if (fixed && R->flat.pb.fixed) std::copy(R->flat.fixed.begin(), R->flat.fixed.end(), fixed);
    return 0;
}

/* info[0..2] space degree, [3..5] space nknots, [6..8] geo degree, [9..11] geo nknots,
   [12] n_basis, [13] n_geo, [14] has_weights */
int gsref_patch_info(void *h, int patch, int32_t *info)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const gsb200_patch &P = R->flat.pb.patches[patch];
    int nb = 1, ng = 1;
    for (int k = 0; k < P.space.dim; ++k) {
        info[k] = P.space.degree[k]; info[3 + k] = P.space.nknots[k];
        info[6 + k] = P.geo.degree[k]; info[9 + k] = P.geo.nknots[k];
        nb *= P.space.nknots[k] - P.space.degree[k] - 1;
        ng *= P.geo.nknots[k] - P.geo.degree[k] - 1;
    }
    info[12] = nb; info[13] = ng; info[14] = P.geo_weights ? 1 : 0;
    return 0;
}

int gsref_patch_data(void *h, int patch, double *sk0, double *sk1, double *sk2, double *gk0, double *gk1,
                     double *gk2, double *coefs, double *weights, int32_t *dofmap)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const gsb200_problem &pb = R->flat.pb;
    const gsb200_patch &P = pb.patches[patch];
    double *sk[3] = {sk0, sk1, sk2}, *gk[3] = {gk0, gk1, gk2};
    int nb = 1, ng = 1;
    for (int k = 0; k < P.space.dim; ++k) {
        std::copy(P.space.knots[k], P.space.knots[k] + P.space.nknots[k], sk[k]);
        std::copy(P.geo.knots[k], P.geo.knots[k] + P.geo.nknots[k], gk[k]);
        nb *= P.space.nknots[k] - P.space.degree[k] - 1;
        ng *= P.geo.nknots[k] - P.geo.degree[k] - 1;
    }
    std::copy(P.geo_coefs, P.geo_coefs + (size_t)ng * P.space.dim, coefs);
    if (P.geo_weights) std::copy(P.geo_weights, P.geo_weights + ng, weights);
    std::copy(P.dofmap, P.dofmap + (size_t)nb * pb.ncomp, dofmap);
    return 0;
}

/* the reference's Gauss table (gsGaussRule.hpp:218-547), to pin the oracle's Newton nodes */
int gsref_gauss(int n, double *nodes, double *weights)
{
    gsVector<index_t> nn(1); nn[0] = n;
    gsGaussRule<real_t> rule(nn);
    for (int i = 0; i < n; ++i) { nodes[i] = rule.referenceNodes()(0, i); weights[i] = rule.referenceWeights()[i]; }
    return 0;
}

/* gsKnotVector::uniformRefine (gsKnotVector.hpp:1048-1063), to pin the host builders */
int gsref_uniform_refine(const double *knots, int nknots, int degree, int numKnots, double *out, int *nout)
{
    gsKnotVector<real_t> kv(degree, knots, knots + nknots);
    kv.uniformRefine(numKnots);
    std::copy(kv.data(), kv.data() + kv.size(), out);
    *nout = (int)kv.size();
    return 0;
}

/* solution coefficients (nfree) and the four norm integrals of a run with `exact`; returns 0 if there are none */
int gsref_solution(void *h, double *sol, double *norms4)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    if (R->sol.size() == 0) return 0;
    std::copy(R->sol.data(), R->sol.data() + R->sol.size(), sol);
    for (int k = 0; k < 4; ++k) norms4[k] = R->norms[k];
    return (int)R->sol.size();
}

/* which = 0: source-term component idx, 1: Neumann data component idx; returns the length (0 = none) */
int gsref_text(void *h, int which, int idx, char *buf, int cap)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const std::vector<std::string> &v = which ? R->neu_text : R->rhs_text;
    if (idx < 0 || idx >= (int)v.size()) return 0;
    snprintf(buf, cap, "%s", v[idx].c_str());
    return (int)v[idx].size();
}

/* (patch, side) pairs of the Neumann sides, in the order the reference visits them */
int gsref_neumann(void *h, int32_t *pairs, int32_t cap)
{
    gsref_result *R = static_cast<gsref_result *>(h);
    const int n = (int)R->neumann.size() / 2;
    for (int i = 0; i < 2 * n && i < cap; ++i) pairs[i] = R->neumann[i];
    return n;
}

int gsref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

} /* extern "C" */
