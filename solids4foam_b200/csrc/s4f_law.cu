// s4f_law.cu -- per-cell (and per-boundary-face) mechanicalLaw stress update, one thread per
// field entry, all whole-field temporaries of the reference fused away:
//   linearElastic                    ML/linearGeometryLaws/linearElastic/linearElastic.C:318-339
//   neoHookeanElastic                ML/nonLinearGeometryLaws/neoHookeanElastic/neoHookeanElastic.C:275-303
//   neoHookeanElasticMisesPlastic    ML/nonLinearGeometryLaws/neoHookeanElasticMisesPlastic/...C:991-1223
//   linearElasticMisesPlastic        ML/linearGeometryLaws/linearElasticMisesPlastic/...C:953-1078
// plus the total-Lagrangian flux tensor J Finv & sigma (nonLinGeomTotalLagTotalDispSolid.C:206, :225-232).
// The index set is cells [0,N) and boundary-value slots [bOff, bOff+B): OpenFOAM's field algebra
// evaluates the law on the boundary patches too (explicit loops :1120-1193 for the plastic law).
#include <cmath>

#include "s4f_ctx.h"
#include "s4f_dev.cuh"

namespace {

struct HardeningTable { int n; double eps[64]; double sig[64]; };

// s4f interpolationTable<scalar>::operator() with outOfBounds clamp, NUM/interpolationTable/interpolationTable.C:493-632
__device__ __forceinline__ double table_lookup(const HardeningTable& T, double x) {
    const int n = T.n;
    if (n <= 1) return T.sig[0];
    if (x < T.eps[0]) return T.sig[0];
    if (x >= T.eps[n - 1]) return T.sig[n - 1];
    int lo = 0, hi = 0;
    for (int i = 0; i < n; i++) {
        if (x >= T.eps[i]) { lo = hi = i; } else { hi = i; break; }
    }
    if (lo == hi) return T.sig[hi];
    return T.sig[lo] + (T.sig[hi] - T.sig[lo]) * (x - T.eps[lo]) / (T.eps[hi] - T.eps[lo]);
}

#define SQRT23 0.81649658092772603   // sqrt(2/3)

__device__ __forceinline__ double cur_yield(const HardeningTable& T, double epsPEq, double J) {
    return J * table_lookup(T, fmax(epsPEq, S4F_SMALL));                       // curYieldStress :139-150
}
__device__ __forceinline__ double yield_fn(const HardeningTable& T, double epsOld, double magS, double DLambda, double muBar, double J) {
    return magS - 2 * muBar * DLambda - SQRT23 * cur_yield(T, epsOld + SQRT23 * DLambda, J);   // yieldFunction :153-183
}
// newtonLoop :186-247 (LoopTol 1e-8, MaxNewtonIter 200, finiteDiff 0.25e-6)
__device__ void newton_loop(const HardeningTable& T, double& DLambda, double& curSigmaY, double epsOld, double magS, double muBar,
                            double J, double maxMagDEps) {
    int i = 0;
    double fTrial = yield_fn(T, epsOld, magS, DLambda, muBar, J);
    double residual = 1.0;
    do {
        const double fStep = yield_fn(T, epsOld, magS, DLambda + 0.25e-6, muBar, J);
        const double deriv = (fStep - fTrial) / 0.25e-6;
        residual = fTrial / deriv;
        DLambda -= residual;
        residual /= maxMagDEps;
        fTrial = yield_fn(T, epsOld, magS, DLambda, muBar, J);
    } while ((fabs(residual) > 1e-8) && ++i < 200);
    curSigmaY = cur_yield(T, epsOld + SQRT23 * DLambda, J) / J;
}
// Ibar: Rubin-Attia cubic enforcing det(bEbar) = 1, :250-395
__device__ __forceinline__ double ibar_of(const double* devB) {
    const double detd = s_det(devB), fac1 = 2.0 * s_magSqr(devB) / 3.0;
    double alpha1;
    if (fabs(fac1) < S4F_SMALL) alpha1 = 3.0;
    else {
        const double fac2 = (4.0 * (1.0 - detd)) / pow(fac1, 1.5);
        if (fac2 >= 1.0) alpha1 = 3.0 * sqrt(fac1) * cosh(acosh(fac2) / 3.0);
        else alpha1 = 3.0 * sqrt(fac1) * cos(acos(fac2) / 3.0);
    }
    return alpha1 / 3.0;
}

__device__ __forceinline__ int field_index(int t, int N, int bOff) { return (t < N) ? t : bOff + (t - N); }

template <int NC>
__device__ __forceinline__ void ld_soa(const double* __restrict__ f, int ld, int i, double* v) {
#pragma unroll
    for (int q = 0; q < NC; q++) v[q] = f[(size_t)q * ld + i];
}
template <int NC>
__device__ __forceinline__ void st_soa(double* __restrict__ f, int ld, int i, const double* v) {
#pragma unroll
    for (int q = 0; q < NC; q++) f[(size_t)q * ld + i] = v[q];
}

// ---- linearElastic: epsilon = symm(gradD); sigma = 2 mu dev(eps) + K tr(eps) I + sigma0 ---------
struct S6 { double v[6]; };
// M (nullable): the combined tensor sigma - gamma grad(D) of the right-hand side on orthogonal meshes (k_source_m), sigma_b on
// the boundary slots
__global__ void __launch_bounds__(S4F_BLOCK) k_law_linear_elastic(const double* __restrict__ gradD, double* __restrict__ sigma, int N,
                                                                  int bOff, int B, int ld, double mu, double K, S6 sigma0,
                                                                  double* __restrict__ M, double gamma, double* __restrict__ pExp) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = field_index(t, N, bOff);
        double g[9], e[6], de[6], s[6];
        ld_soa<9>(gradD, ld, i, g);
        t_symm(g, e);
        const double sh = K * s_tr(e);
        if (pExp) pExp[i] = sh;
        s_dev(e, de);
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] = 2.0 * mu * de[q] + sigma0.v[q];
        s[0] += sh; s[3] += sh; s[5] += sh;
        st_soa<6>(sigma, ld, i, s);
        if (M) {
            double s9[9]; s_to_t(s, s9);
            const double gm = (t < N) ? gamma : 0.0;
#pragma unroll
            for (int q = 0; q < 9; q++) s9[q] -= gm * g[q];
            st_soa<9>(M, ld, i, s9);
        }
    }
}

// ---- neoHookeanElastic ---------------------------------------------------------------------------
__global__ void __launch_bounds__(S4F_BLOCK) k_law_neo_hookean(const double* __restrict__ gradD, double* __restrict__ sigma,
                                                               double* __restrict__ Jout, int N, int bOff, int B, int ld, double mu, double K,
                                                               const double* __restrict__ Fold /* updated Lagrangian only */,
                                                               double* __restrict__ Fout, double* __restrict__ pExp) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = field_index(t, N, bOff);
        double g[9], Fm[9], FT[9], FFT[9], b[6], s[6];
        ld_soa<9>(gradD, ld, i, g);
        t_transpose(g, Fm); Fm[0] += 1; Fm[4] += 1; Fm[8] += 1;          // F = I + gradD.T()  mechanicalLaw.C:1130-1135
        if (Fold) {                                                       // UL: relF = I + gradDD.T(); F = relF & F.oldTime()  :1055-1072
            double Fo[9], rF[9];
            ld_soa<9>(Fold, ld, i, Fo);
#pragma unroll
            for (int q = 0; q < 9; q++) rF[q] = Fm[q];
            t_mul(rF, Fo, Fm);
            st_soa<9>(Fout, ld, i, Fm);
        }
        const double J = t_det(Fm);
        t_transpose(Fm, FT); t_mul(Fm, FT, FFT); t_symm(FFT, b);
        const double sc = pow(J, -2.0 / 3.0);
#pragma unroll
        for (int q = 0; q < 6; q++) b[q] *= sc;
        s_dev(b, s);
        const double sh = 0.5 * K * (pow(J, 2.0) - 1.0);
        if (pExp) pExp[i] = sh;
        const double rJ = 1.0 / J;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] *= mu;
        s[0] += sh; s[3] += sh; s[5] += sh;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] *= rJ;
        st_soa<6>(sigma, ld, i, s);
        Jout[i] = J;
    }
}

// ---- neoHookeanElasticMisesPlastic ---------------------------------------------------------------
// trial state shared by the two passes: F, J, relF = F & inv(F.old), relFbar, bEbarTrial = transform(relFbar, bEbar.old)
__device__ __forceinline__ void mises_trial(const double* __restrict__ gradD, const double* __restrict__ Fold, const double* __restrict__ Jold,
                                            const double* __restrict__ bEbarOld, int ld, int i, double* Fm, double& J, double* bt, int UL) {
    double g[9], Fo[9], Fi[9], relF[9], bo6[6], bo[9], t1[9], rT[9], t2[9];
    ld_soa<9>(gradD, ld, i, g);
    t_transpose(g, Fm); Fm[0] += 1; Fm[4] += 1; Fm[8] += 1;
    ld_soa<9>(Fold, ld, i, Fo);
    if (UL) {            // gradD is grad(DD) on the updated configuration: relF = I + gradDD.T(); F = relF & F.oldTime()
#pragma unroll
        for (int q = 0; q < 9; q++) relF[q] = Fm[q];
        t_mul(relF, Fo, Fm);
    } else { t_inv(Fo, Fi); t_mul(Fm, Fi, relF); }
    J = t_det(Fm);
    const double relJ = J / Jold[i];
    const double sc = pow(relJ, -1.0 / 3.0);
#pragma unroll
    for (int q = 0; q < 9; q++) relF[q] *= sc;
    ld_soa<6>(bEbarOld, ld, i, bo6);
    s_to_t(bo6, bo); t_mul(relF, bo, t1); t_transpose(relF, rT); t_mul(t1, rT, t2);
    t_symm(t2, bt);
}

struct FinMaxBE { OuterScalars* S; __device__ void operator()(const double* tot) const { S->maxMagBE = tot[0]; } };
__global__ void __launch_bounds__(S4F_BLOCK) k_mises_max_be(const double* __restrict__ gradD, const double* __restrict__ Fold,
                                                            const double* __restrict__ Jold, const double* __restrict__ bEbarOld, int N, int ld,
                                                            OuterScalars* S, RedCtx red, int UL) {
    double v[1] = {0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {   // gMax over the internal field :1030
        double Fm[9], J, bt[6];
        mises_trial(gradD, Fold, Jold, bEbarOld, ld, i, Fm, J, bt, UL);
        v[0] = fmax(v[0], sqrt(s_magSqr(bt)));
    }
    grid_reduce<1, OpMax>(v, red, FinMaxBE{S});
}

struct FinMat { OuterScalars* S; __device__ void operator()(const double* tot) const { S->matNum = tot[0]; S->matDen = tot[1]; } };

struct MisesPtrs {
    const double* gradD; const double* Fold; const double* Jold; const double* bEbarOld;
    const double* sigmaY; const double* epsPEq;
    double* F; double* J; double* bEbar; double* sigma; double* DSigmaY; double* DEpsPEq; double* DEpsP; double* DEpsPprev;
    double* DLambda; double* plasticN; double* pExp;
};
__global__ void __launch_bounds__(S4F_BLOCK) k_law_mises(MisesPtrs p, int N, int bOff, int B, int ld, double mu, double K, double Hp,
                                                         int consistent, double relax, HardeningTable T, OuterScalars* S,
                                                         RedCtx red, int UL) {
    const bool nonLinearPlasticity = T.n > 2;
    const double magHp = fabs(Hp);
    const double maxMagBE = fmax(S->maxMagBE, S4F_SMALL);
    double v[2] = {0, 0};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = field_index(t, N, bOff);
        double Fm[9], J, bt[6];
        mises_trial(p.gradD, p.Fold, p.Jold, p.bEbarOld, ld, i, Fm, J, bt, UL);
        double dv[6], sT[6];
        s_dev(bt, dv);
#pragma unroll
        for (int q = 0; q < 6; q++) sT[q] = mu * dv[q];
        const double Ibar = s_tr(bt) / 3.0, muBar = Ibar * mu;
        const double sigY = p.sigmaY[i];
        const double magS = sqrt(s_magSqr(sT));
        const double fTrial = magS - SQRT23 * J * sigY;
        double pn[6];
        ld_soa<6>(p.plasticN, ld, i, pn);
        if (magS > S4F_SMALL) {
#pragma unroll
            for (int q = 0; q < 6; q++) pn[q] = sT[q] / magS;
        }
        double DLambda = p.DLambda[i], DSigY = p.DSigmaY[i];
        if (fTrial < S4F_SMALL) { DSigY = 0; DLambda = 0; }
        else if (nonLinearPlasticity) {
            double curSigmaY = 0;
            newton_loop(T, DLambda, curSigmaY, p.epsPEq[i], magS, muBar, J, maxMagBE);
            DSigY = curSigmaY - sigY;
        } else {
            DLambda = fTrial / (2 * muBar);
            if (magHp > S4F_SMALL) { DLambda /= 1.0 + Hp / (3 * muBar); DSigY = SQRT23 * DLambda * Hp; }
        }
        double prev[6], dep[6], s[6], devB[6];
        ld_soa<6>(p.DEpsP, ld, i, prev);                 // DEpsilonP_.storePrevIter()
#pragma unroll
        for (int q = 0; q < 6; q++) {
            double x = Ibar * DLambda * pn[q];
            if (relax != 1.0) x = prev[q] + relax * (x - prev[q]);   // DEpsilonP_.relax()
            dep[q] = x;
            s[q] = sT[q] - 2 * mu * x;
            devB[q] = s[q] / mu;
        }
        const double Ib = consistent ? ibar_of(devB) : Ibar;
        double be[6];
#pragma unroll
        for (int q = 0; q < 6; q++) be[q] = devB[q];
        be[0] += Ib; be[3] += Ib; be[5] += Ib;
        const double sh = 0.5 * K * (pow(J, 2.0) - 1.0);
        if (p.pExp) p.pExp[i] = sh;
        s[0] += sh; s[3] += sh; s[5] += sh;
        const double rJ = 1.0 / J;
#pragma unroll
        for (int q = 0; q < 6; q++) s[q] *= rJ;
        st_soa<9>(p.F, ld, i, Fm); p.J[i] = J;
        st_soa<6>(p.bEbar, ld, i, be); st_soa<6>(p.sigma, ld, i, s);
        st_soa<6>(p.DEpsPprev, ld, i, prev); st_soa<6>(p.DEpsP, ld, i, dep); st_soa<6>(p.plasticN, ld, i, pn);
        p.DLambda[i] = DLambda; p.DSigmaY[i] = DSigY; p.DEpsPEq[i] = SQRT23 * DLambda;
        if (t < N) {   // residual(): :1502-1521, internal field
            double d[6];
#pragma unroll
            for (int q = 0; q < 6; q++) d[q] = dep[q] - prev[q];
            v[0] = fmax(v[0], sqrt(s_magSqr(d)));
            v[1] = fmax(v[1], S4F_SMALL + sqrt(s_magSqr(prev)));
        }
    }
    grid_reduce<2, OpMax>(v, red, FinMat{S});
}

// ---- linearElasticMisesPlastic ----------------------------------------------------------------------
__global__ void __launch_bounds__(S4F_BLOCK) k_lin_mises_max_eps(const double* __restrict__ gradD, int N, int ld, OuterScalars* S,
                                                                 RedCtx red) {
    double v[1] = {0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        double g[9], e[6];
        ld_soa<9>(gradD, ld, i, g); t_symm(g, e);
        v[0] = fmax(v[0], sqrt(s_magSqr(e)));
    }
    grid_reduce<1, OpMax>(v, red, FinMaxBE{S});
}
struct LinMisesPtrs {
    const double* gradD; const double* epsPOld; const double* sigmaYOld; const double* epsPEqOld;
    double* epsilon; double* sigma; double* sigmaY; double* DSigmaY; double* epsPEq; double* DEpsPEq; double* epsP; double* DEpsP;
    double* DEpsPprev; double* DLambda; double* plasticN; double* pExp;
};
__global__ void __launch_bounds__(S4F_BLOCK) k_law_lin_mises(LinMisesPtrs p, int N, int bOff, int B, int ld, double mu, double K, double Hp,
                                                             HardeningTable T, OuterScalars* S, RedCtx red) {
    const bool nonLinearPlasticity = T.n > 2;
    const double maxMagBE = fmax(S->maxMagBE, S4F_SMALL);
    double v[2] = {0, 0};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = field_index(t, N, bOff);
        double g[9], eps[6], e[6], epo[6], dpo[6], sT[6];
        ld_soa<9>(p.gradD, ld, i, g); t_symm(g, eps); s_dev(eps, e);
        ld_soa<6>(p.epsPOld, ld, i, epo); s_dev(epo, dpo);
#pragma unroll
        for (int q = 0; q < 6; q++) sT[q] = 2.0 * mu * (e[q] - dpo[q]);
        const double sYold = p.sigmaYOld[i];
        const double magS = sqrt(s_magSqr(sT));
        const double fT = magS - SQRT23 * sYold;
        double pn[6] = {1, 0, 0, 1, 0, 1};
        double DLambda = p.DLambda[i], DSigY = p.DSigmaY[i], sY = p.sigmaY[i];
        if (fT < S4F_SMALL) { DLambda = 0; DSigY = 0; sY = sYold; }
        else {
            if (magS > S4F_SMALL) {
#pragma unroll
                for (int q = 0; q < 6; q++) pn[q] = sT[q] / magS;
            }
            if (nonLinearPlasticity) {
                newton_loop(T, DLambda, sY, p.epsPEqOld[i], magS, mu, 1.0, maxMagBE);
                DSigY = sY - sYold;
            } else {
                DLambda = fT / (2 * mu);
                if (fabs(Hp) > S4F_SMALL) { DLambda /= 1.0 + Hp / (3 * mu); DSigY = SQRT23 * DLambda * Hp; sY = sYold + DSigY; }
            }
        }
        double prev[6], dep[6], ep[6], s[6];
        ld_soa<6>(p.DEpsP, ld, i, prev);
#pragma unroll
        for (int q = 0; q < 6; q++) { dep[q] = DLambda * pn[q]; ep[q] = epo[q] + dep[q]; s[q] = sT[q] - 2 * mu * dep[q]; }
        const double sh = K * s_tr(eps);
        if (p.pExp) p.pExp[i] = sh;
        s[0] += sh; s[3] += sh; s[5] += sh;
        st_soa<6>(p.epsilon, ld, i, eps); st_soa<6>(p.sigma, ld, i, s); st_soa<6>(p.epsP, ld, i, ep);
        st_soa<6>(p.DEpsPprev, ld, i, prev); st_soa<6>(p.DEpsP, ld, i, dep); st_soa<6>(p.plasticN, ld, i, pn);
        p.DLambda[i] = DLambda; p.DSigmaY[i] = DSigY; p.sigmaY[i] = sY;
        p.DEpsPEq[i] = SQRT23 * DLambda; p.epsPEq[i] = p.epsPEqOld[i] + SQRT23 * DLambda;
        if (t < N) {
            double d[6];
#pragma unroll
            for (int q = 0; q < 6; q++) d[q] = dep[q] - prev[q];
            v[0] = fmax(v[0], sqrt(s_magSqr(d)));
            v[1] = fmax(v[1], S4F_SMALL + sqrt(s_magSqr(prev)));
        }
    }
    grid_reduce<2, OpMax>(v, red, FinMat{S});
}

// ---- total-Lagrangian flux tensor: F = I + gradD.T(); Finv; J; T = J Finv & sigma ------------------
// gradSol (nullable): gradient of the SOLUTION field (grad(D) or grad(DD)); when given, cells get T - gamma gradSol, the combined
// tensor of the right-hand side on orthogonal meshes (k_source_m)
__global__ void __launch_bounds__(S4F_BLOCK) k_tl_flux_tensor(const double* __restrict__ gradD, const double* __restrict__ sigma,
                                                              double* __restrict__ T9, double* __restrict__ Finv, double* __restrict__ Jt,
                                                              int N, int bOff, int B, int ld, const double* __restrict__ gradSol, double gamma) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N + B; t += gridDim.x * blockDim.x) {
        const int i = field_index(t, N, bOff);
        double g[9], Fm[9], Fi[9], s6[6], s9[9], T[9];
        ld_soa<9>(gradD, ld, i, g);
        t_transpose(g, Fm); Fm[0] += 1; Fm[4] += 1; Fm[8] += 1;
        t_inv(Fm, Fi);
        const double J = t_det(Fm);
        ld_soa<6>(sigma, ld, i, s6); s_to_t(s6, s9);
        t_mul(Fi, s9, T);
#pragma unroll
        for (int q = 0; q < 9; q++) T[q] *= J;
        if (gradSol && t < N) {
            double gs[9]; ld_soa<9>(gradSol, ld, i, gs);
#pragma unroll
            for (int q = 0; q < 9; q++) T[q] -= gamma * gs[q];
        }
        st_soa<9>(T9, ld, i, T);
        Jt[i] = J;
        if (t >= N) st_soa<9>(Finv, ld, i, Fi);      // only the traction BC reads Finv, on boundary faces
    }
}

__global__ void k_axpy(double* __restrict__ y, const double* __restrict__ x, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) y[i] += x[i];
}

HardeningTable make_table(const s4fgpu_law& L) {
    HardeningTable T; T.n = L.nTable;
    for (int i = 0; i < 64; i++) { T.eps[i] = L.tableEps[i]; T.sig[i] = L.tableSigY[i]; }
    return T;
}

}  // namespace

// mechanicalModel::correct(sigma) (mechanicalModel.C:476-483), followed -- for the finite-strain
// solid models -- by the flux tensor the momentum equation needs, and the halo exchange of the result.
int s4f_law_correct(s4fgpu_ctx* c) {
    const int N = c->N, B = c->B, bOff = c->bOff(), ld = c->ld;
    const int grid = s4f_grid(c->numSMs, N + B), gridN = s4f_grid(c->numSMs, N);
    const s4fgpu_law& L = c->law;
    const int UL = c->UL() ? 1 : 0;
    const double* gD = UL ? c->gradD.p : c->gradForLaw();        // UL laws read grad(DD) (mechanicalLaw.C:1064-1072)
    if (L.solvePressureEqn && c->pExp.n != (size_t)ld) {
        S4F_CHECK_CUDA(c, c->pExp.alloc(ld)); S4F_CHECK_CUDA(c, c->sigmaHyd.alloc(ld)); S4F_CHECK_CUDA(c, c->pRatio.alloc(ld));
        S4F_CHECK_CUDA(c, c->gradP.alloc(3 * (size_t)ld)); S4F_CHECK_CUDA(c, c->eP.alloc((size_t)c->nEntries));
        for (DevBuf<double>* b : {&c->pDiag, &c->pRDiag, &c->pX, &c->pB}) S4F_CHECK_CUDA(c, b->alloc(3 * (size_t)ld));
    }
    double* pE = L.solvePressureEqn ? c->pExp.p : nullptr;
    if (L.kind == S4F_LAW_LINEAR_ELASTIC) {
        S6 s0; for (int q = 0; q < 6; q++) s0.v[q] = L.sigma0[q];
        const bool emitM = c->ctl.solidModel == S4F_MODEL_LIN_GEOM_TOTAL_DISP && !L.solvePressureEqn;
        k_law_linear_elastic<<<grid, S4F_BLOCK, 0, c->stream>>>(gD, c->sigma.p, N, bOff, B, ld, L.mu, L.K, s0, emitM ? c->T9.p : nullptr, c->gamma0(), pE);
        c->launches++;
        if (emitM) { c->mValid = true; S4F_CHECK_CUDA(c, cudaGetLastError()); return s4f_halo_exchange(c, c->T9.p, 9); }
    } else if (L.kind == S4F_LAW_NEO_HOOKEAN_ELASTIC) {
        k_law_neo_hookean<<<grid, S4F_BLOCK, 0, c->stream>>>(gD, c->sigma.p, c->lawJ.p, N, bOff, B, ld, L.mu, L.K, UL ? c->lawFold.p : nullptr,
                                                            c->lawF.p, pE);
        c->launches++;
    } else if (L.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC) {
        HardeningTable T = make_table(L);
        if (T.n > 2) {
            k_mises_max_be<<<gridN, S4F_BLOCK, 0, c->stream>>>(gD, c->lawFold.p, c->lawJold.p, c->bEbarOld.p, N, ld, c->outS.p, c->red(), UL);
            c->launches++;
        }
        MisesPtrs p{gD, c->lawFold.p, c->lawJold.p, c->bEbarOld.p, c->sigmaY.p, c->epsPEq.p, c->lawF.p, c->lawJ.p, c->bEbar.p, c->sigma.p,
                    c->DSigmaY.p, c->DEpsPEq.p, c->DEpsP.p, c->DEpsPprev.p, c->DLambda.p, c->plasticN.p, pE};
        k_law_mises<<<grid, S4F_BLOCK, 0, c->stream>>>(p, N, bOff, B, ld, L.mu, L.K, c->Hp, L.updateBEbarConsistent, L.DEpsilonPRelax, T, c->outS.p,
                                                      c->red(), UL);
        c->launches++;
    } else if (L.kind == S4F_LAW_LINEAR_ELASTIC_MISES_PLASTIC) {
        HardeningTable T = make_table(L);
        if (T.n > 2) {
            k_lin_mises_max_eps<<<gridN, S4F_BLOCK, 0, c->stream>>>(gD, N, ld, c->outS.p, c->red());
            c->launches++;
        }
        LinMisesPtrs p{gD, c->epsPOld.p, c->sigmaYOld.p, c->epsPEqOld.p, c->epsilon.p, c->sigma.p, c->sigmaY.p, c->DSigmaY.p, c->epsPEq.p,
                       c->DEpsPEq.p, c->epsP.p, c->DEpsP.p, c->DEpsPprev.p, c->DLambda.p, c->plasticN.p, pE};
        k_law_lin_mises<<<grid, S4F_BLOCK, 0, c->stream>>>(p, N, bOff, B, ld, L.mu, L.K, c->Hp, T, c->outS.p, c->red());
        c->launches++;
    } else {
        c->err = "unknown mechanical law"; return 1;
    }
    S4F_CHECK_CUDA(c, cudaGetLastError());
    if (L.solvePressureEqn) { int rp = s4f_pressure_smooth(c); if (rp) return rp; }
    if (c->finiteStrain()) {
        // UL: relF = I + gradDD.T() takes the place of F: fvc::div(relJ*relFinv & sigma), nonLinGeomUpdatedLagSolid.C:188
        k_tl_flux_tensor<<<grid, S4F_BLOCK, 0, c->stream>>>(gD, c->sigma.p, c->T9.p, c->Finv.p, c->Jt.p, N, bOff, B, ld,
                                                            c->gradD.p, c->gamma0());
        c->launches++;
        c->mValid = true;
        return s4f_halo_exchange(c, c->T9.p, 9);
    }
    c->mValid = false;          // other laws under the linear-geometry model: M is formed by k_make_m before the next right-hand side
    return s4f_halo_exchange(c, c->sigma.p, 6);
}

// Algorithmic bytes of one s4f_law_correct (every field read / written once per cell, fp64), for the roofline report:
//   linearElastic              grad(D) in | sigma, M out                                             72 + 48 + 72
//   neoHookeanElastic          grad(D) [, F.old] in | [F,] J, sigma out  (F only for the updated-Lagrangian model)  72 [+72 +72] + 8 + 48
//   neoHookeanElasticMisesPlastic  k_mises_max_be (table > 2 points): grad(D), F.old, J.old, bEbar.old  200
//                              k_law_mises: those + sigmaY, plasticN, DLambda, DSigmaY, epsPEq, DEpsP in  328
//                                           F, J, bEbar, sigma, DEpsP.prevIter, DEpsP, plasticN, DLambda, DSigmaY, DEpsPEq out  344
//   linearElasticMisesPlastic  grad(D), epsP.old, sigmaY.old, epsPEq.old, DEpsP, DLambda, plasticN in | epsilon, sigma, sigmaY,
//                              DSigmaY, epsPEq, DEpsPEq, epsP, DEpsP, DEpsP.prevIter, DLambda, plasticN out  (+ 72 for the max-epsilon pass)
//   finite-strain models       + k_tl_flux_tensor: grad(D), sigma [, grad(DD)] in | M, J out           72 + 48 [+72] + 72 + 8
double s4f_law_bytes(const s4fgpu_ctx* c) {
    const double n = (double)c->N + c->B;
    const s4fgpu_law& L = c->law;
    double b = 0;
    if (L.kind == S4F_LAW_LINEAR_ELASTIC) b = 72 + 48 + ((c->ctl.solidModel == S4F_MODEL_LIN_GEOM_TOTAL_DISP && !L.solvePressureEqn) ? 72 : 0);
    else if (L.kind == S4F_LAW_NEO_HOOKEAN_ELASTIC) b = 72 + (c->UL() ? 72 + 72 : 0) + 8 + 48;
    else if (L.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC) b = (L.nTable > 2 ? 200.0 * c->N / n : 0) + 328 + 344;
    else b = (L.nTable > 2 ? 72.0 * c->N / n : 0) + (72 + 48 + 8 + 8 + 48 + 8 + 48) + (48 + 48 + 8 + 8 + 8 + 8 + 48 + 48 + 48 + 8 + 48);
    if (c->finiteStrain()) b += 72 + 48 + (c->incremental() ? 72 : 0) + 72 + 8;      // Finv is written on the boundary slots only
    return b * n;
}

// Solver-level kinematics of the total-Lagrangian models: F = I + gradD.T(), Finv, J and the flux tensor
// J Finv & sigma (nonLinGeomTotalLagTotalDispSolid.C:225-232) from the current gradD and sigma.
int s4f_kinematics(s4fgpu_ctx* c) {
    if (!c->finiteStrain()) return 0;
    const int grid = s4f_grid(c->numSMs, c->N + c->B);
    k_tl_flux_tensor<<<grid, S4F_BLOCK, 0, c->stream>>>(c->UL() ? c->gradD.p : c->gradForLaw(), c->sigma.p, c->T9.p, c->Finv.p, c->Jt.p, c->N, c->bOff(), c->B, c->ld,
                                                        c->gradD.p, c->gamma0());
    c->launches++;
    c->mValid = true;
    S4F_CHECK_CUDA(c, cudaGetLastError());
    return s4f_halo_exchange(c, c->T9.p, 9);
}

// solidModel::updateTotalFields -> neoHookeanElasticMisesPlastic::updateTotalFields :1526-1536
namespace {
__global__ void k_rho_update(double* __restrict__ rho, const double* __restrict__ rhoO, const double* __restrict__ relJ, int N, int bOff, int B) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N + B) return;
    const int i = field_index(t, N, bOff);
    rho[i] = rhoO[i] / relJ[i];
}
}  // namespace

int s4f_update_total_fields_impl(s4fgpu_ctx* c) {
    if (c->UL()) {
        // after the loop (nonLinGeomUpdatedLagSolid.C:243): gradD() = fvc::grad(D().oldTime() + DD()); then
        // updateTotalFields :360-374: rho_ = rho_.oldTime()/relJ_.  The mesh motion is done by the host side of the
        // boundary (interpolate_to_points -> movePoints -> set_geometry / set_points).
        int rc = s4f_grad_calculated(c, c->Dtot.p, c->gradDtot.p); if (rc) return rc;
        k_rho_update<<<(c->N + c->B + 255) / 256, 256, 0, c->stream>>>(c->rhoF.p, c->rhoO.p, c->Jt.p, c->N, c->bOff(), c->B);
        c->launches++;
        c->matrixValid = false; c->histValid = false;
    }
    if (c->law.kind == S4F_LAW_NEO_HOOKEAN_MISES_PLASTIC) {
        const long long ld = c->ld;
        k_axpy<<<(unsigned)((ld + 255) / 256), 256, 0, c->stream>>>(c->sigmaY.p, c->DSigmaY.p, ld);
        k_axpy<<<(unsigned)((ld + 255) / 256), 256, 0, c->stream>>>(c->epsPEq.p, c->DEpsPEq.p, ld);
        k_axpy<<<(unsigned)((6 * ld + 255) / 256), 256, 0, c->stream>>>(c->epsP.p, c->DEpsP.p, 6 * ld);
        c->launches += 3;
        S4F_CHECK_CUDA(c, cudaGetLastError());
    }
    return 0;
}
