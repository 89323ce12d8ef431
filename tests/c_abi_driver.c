/* c_abi_driver.c -- a plain-C caller of libs4fgpu.so, the only non-Python user of the ABI before a real OpenFOAM build.
 *
 * It does what the OpenFOAM plugin's constructor + evolve() do (foam_plugin/gpuSolidBridge/gpuSolidBridge.C): builds the
 * lduAddressing, patches and geometry of a uniform hex box (cells i + nx (j + ny k), internal faces in upper-triangular
 * order, six patches), mirrors them through s4fgpu_set_mesh / s4fgpu_set_geometry, sets a linearElastic law, the
 * controls of linearGeometryTotalDisplacement and the cantilever boundary conditions, runs s4fgpu_evolve and writes D
 * (AoS, nCells x 3 doubles) to a binary file that tests/test_c_caller.py compares with the CPU oracle.
 *
 *   gcc -std=c99 -O2 -Iinclude tests/c_abi_driver.c -Lsolids4foam_b200 -ls4fgpu -lm -o c_abi_driver
 *   ./c_abi_driver nx ny nz out.bin
 * exit codes: 0 ok, 2 no CUDA device (the library has no CPU fallback), 1 any other error.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "s4fgpu.h"

#define CHECK(call)                                                                          \
    do {                                                                                     \
        int rc_ = (call);                                                                    \
        if (rc_ != 0) {                                                                      \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, s4fgpu_last_error(h));       \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 4, ny = argc > 2 ? atoi(argv[2]) : 2, nz = argc > 3 ? atoi(argv[3]) : 2;
    const char* out = argc > 4 ? argv[4] : "c_abi_driver_D.bin";
    const double L = 2.0, H = 1.0, W = 1.0, dx = L / nx, dy = H / ny, dz = W / nz;
    const int N = nx * ny * nz;
    const int Fx = (nx - 1) * ny * nz, Fy = nx * (ny - 1) * nz, Fz = nx * ny * (nz - 1), F = Fx + Fy + Fz;
    const int pSizeArr[6] = {ny * nz, ny * nz, nx * nz, nx * nz, nx * ny, nx * ny};   /* xMin xMax yMin yMax zMin zMax */
    int pStart[6], pKind[6], pNbr[6], B = 0, p, i, j, k, f, q;
    for (p = 0; p < 6; p++) { pStart[p] = B; B += pSizeArr[p]; pKind[p] = S4F_PATCH_GENERIC; pNbr[p] = -1; }

    int* owner = malloc(sizeof(int) * F); int* neighbour = malloc(sizeof(int) * F); int* faceCells = malloc(sizeof(int) * B);
    double* C = calloc(3 * (size_t)N, sizeof(double)); double* V = malloc(sizeof(double) * N);
    double* Sf = calloc(3 * (size_t)(F + B), sizeof(double)); double* magSf = calloc(F + B, sizeof(double));
    double* Cf = calloc(3 * (size_t)(F + B), sizeof(double)); double* w = malloc(sizeof(double) * (F + B));
    double* nod = malloc(sizeof(double) * (F + B)); double* corr = calloc(3 * (size_t)(F + B), sizeof(double));
    double* CnbrB = calloc(3 * (size_t)B, sizeof(double));

    for (k = 0; k < nz; k++) for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
        const int c = i + nx * (j + ny * k);
        C[3 * c] = (i + 0.5) * dx; C[3 * c + 1] = (j + 0.5) * dy; C[3 * c + 2] = (k + 0.5) * dz;
        V[c] = dx * dy * dz;
    }
    /* internal faces: owner ascending, per owner the neighbours ascending (c+1 < c+nx < c+nx*ny): upper-triangular order */
    f = 0;
    for (k = 0; k < nz; k++) for (j = 0; j < ny; j++) for (i = 0; i < nx; i++) {
        const int c = i + nx * (j + ny * k);
        const int has[3] = {i < nx - 1, j < ny - 1, k < nz - 1};
        const int nb[3] = {c + 1, c + nx, c + nx * ny};
        const double area[3] = {dy * dz, dx * dz, dx * dy}, dist[3] = {dx, dy, dz};
        for (q = 0; q < 3; q++) {
            if (!has[q]) continue;
            owner[f] = c; neighbour[f] = nb[q];
            Sf[3 * f + q] = area[q]; magSf[f] = area[q];
            Cf[3 * f] = C[3 * c]; Cf[3 * f + 1] = C[3 * c + 1]; Cf[3 * f + 2] = C[3 * c + 2];
            Cf[3 * f + q] += 0.5 * dist[q];
            w[f] = 0.5; nod[f] = 1.0 / dist[q];
            f++;
        }
    }
    if (f != F) { fprintf(stderr, "face count\n"); return 1; }
    /* boundary faces, patch by patch */
    {
        int b = 0;
        for (p = 0; p < 6; p++) {
            const int dir = p / 2, hi = p % 2;
            const double area = dir == 0 ? dy * dz : (dir == 1 ? dx * dz : dx * dy), half = 0.5 * (dir == 0 ? dx : (dir == 1 ? dy : dz));
            int a, bb;
            const int na = dir == 0 ? ny : nx, nb2 = dir == 2 ? ny : nz;
            for (bb = 0; bb < nb2; bb++) for (a = 0; a < na; a++) {
                int ci, cj, ck;
                if (dir == 0) { ci = hi ? nx - 1 : 0; cj = a; ck = bb; }
                else if (dir == 1) { ci = a; cj = hi ? ny - 1 : 0; ck = bb; }
                else { ci = a; cj = bb; ck = hi ? nz - 1 : 0; }
                const int c = ci + nx * (cj + ny * ck), g = F + b;
                faceCells[b] = c;
                Sf[3 * g + dir] = hi ? area : -area; magSf[g] = area;
                Cf[3 * g] = C[3 * c]; Cf[3 * g + 1] = C[3 * c + 1]; Cf[3 * g + 2] = C[3 * c + 2];
                Cf[3 * g + dir] += hi ? half : -half;
                w[g] = 1.0; nod[g] = 1.0 / half;
                CnbrB[3 * b] = Cf[3 * g]; CnbrB[3 * b + 1] = Cf[3 * g + 1]; CnbrB[3 * b + 2] = Cf[3 * g + 2];
                b++;
            }
        }
    }

    s4fgpu_handle h = NULL;
    if (s4fgpu_create(&h, 0) != 0) { fprintf(stderr, "%s\n", s4fgpu_last_error(NULL)); return 2; }
    const int solD[3] = {1, 1, 1};
    CHECK(s4fgpu_set_mesh(h, N, F, owner, neighbour, 6, pStart, pSizeArr, pKind, pNbr, faceCells, solD));
    CHECK(s4fgpu_set_geometry(h, C, V, Sf, magSf, Cf, w, nod, corr, CnbrB));

    s4fgpu_law law; memset(&law, 0, sizeof(law));
    const double E = 200e9, nu = 0.3;
    law.kind = S4F_LAW_LINEAR_ELASTIC; law.rho = 7800.0;
    law.mu = E / (2.0 * (1.0 + nu)); law.lambda = nu * E / ((1.0 + nu) * (1.0 - 2.0 * nu)); law.K = E / (3.0 * (1.0 - 2.0 * nu));
    law.updateBEbarConsistent = 1; law.DEpsilonPRelax = 1.0; law.pressureSmoothingScaleFactor = 100.0;
    CHECK(s4fgpu_set_law(h, &law));

    s4fgpu_controls ctl; memset(&ctl, 0, sizeof(ctl));
    ctl.solidModel = S4F_MODEL_LIN_GEOM_TOTAL_DISP; ctl.gradScheme = S4F_GRAD_LEAST_SQUARES; ctl.d2dt2Scheme = S4F_D2DT2_STEADY_STATE;
    ctl.stabilisation = S4F_STAB_RHIE_CHOW; ctl.stabScaleFactor = 0.1; ctl.relaxationMethod = S4F_RELAX_FIXED; ctl.fieldRelaxD = 0.9;
    ctl.solver = S4F_SOLVER_PCG; ctl.preconditioner = S4F_PRECOND_DIAGONAL; ctl.tolerance = 1e-13; ctl.relTol = 0.1; ctl.maxIter = 1000;
    ctl.nCorrectors = 20000; ctl.solutionTolerance = 1e-11; ctl.alternativeTolerance = 1e-11; ctl.materialTolerance = 1e-5;
    ctl.deltaT = 1.0; ctl.deltaT0 = 1.0; ctl.chebyshevDegree = 4; ctl.checkEvery = 4; ctl.gamgOverCorrection = 2.2;
    ctl.gamgSmootherDegree = 3; ctl.gamgCycle = 2; ctl.gamgSmootherRatio = 0.3;
    CHECK(s4fgpu_set_controls(h, &ctl));

    /* 0/D: xMin fixedDisplacement (0 0 0); xMax solidTraction (0 -1e6 0); the rest traction free */
    for (p = 0; p < 6; p++) {
        double* val = calloc(3 * (size_t)pSizeArr[p], sizeof(double));
        if (p == 1) for (i = 0; i < pSizeArr[p]; i++) val[3 * i + 1] = -1e6;
        CHECK(s4fgpu_set_bc(h, p, p == 0 ? S4F_BC_FIXED_DISPLACEMENT : S4F_BC_SOLID_TRACTION, val, NULL));
        free(val);
    }
    CHECK(s4fgpu_initialise(h));
    s4fgpu_stats st;
    CHECK(s4fgpu_evolve(h, &st));
    double* D = malloc(sizeof(double) * 3 * (size_t)N);
    CHECK(s4fgpu_download(h, S4F_FIELD_D, D));
    printf("converged %d after %d outer iterations, %lld PCG iterations, relative residual %.3e, launches %lld\n", st.converged, st.nCorr,
           st.totalInnerIterations, st.relResidual, s4fgpu_launch_count(h));
    FILE* fo = fopen(out, "wb");
    if (!fo || fwrite(D, sizeof(double), 3 * (size_t)N, fo) != 3 * (size_t)N) { fprintf(stderr, "cannot write %s\n", out); return 1; }
    fclose(fo);
    CHECK(s4fgpu_destroy(h));
    return st.converged ? 0 : 1;
}
