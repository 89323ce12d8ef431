// Micro-benchmark of mapping variants for the 3-component SELL-32 SpMV (design exploration; the winner is
// what s4f_pcg.cu / s4f_amg.cu ship).  Synthetic structured hex matrix in the library's row order.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o spmv_variants spmv_variants.cu
//   ./spmv_variants 800 100 100
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

__device__ __forceinline__ void warp_sum3(double v0, double v1, double v2, double* out) {
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(~0u, v0, o); v1 += __shfl_xor_sync(~0u, v1, o); v2 += __shfl_xor_sync(~0u, v2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, v0); atomicAdd(out + 1, v1); atomicAdd(out + 2, v2); }
}

// A: row per thread, unroll U, three vectors
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_rows(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                                     const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w,
                                                     int N, int ld, int nS, double* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    double v0 = 0, v1 = 0, v2 = 0;
    for (int s = warp; s < nS; s += nW) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0, a1 = 0, a2 = 0;
        const int* cp = col + base + lane; const double* ap = a + base + lane;
#pragma unroll U
        for (int k = 0; k < width; k++) { const int cc = cp[32 * k]; const double e = ap[32 * k]; a0 += e * p[cc]; a1 += e * p[cc + ld]; a2 += e * p[cc + 2 * ld]; }
        if (row < N) {
            double pp = p[row], ww = dg[row] * pp - a0; w[row] = ww; v0 += ww * pp;
            pp = p[row + ld]; ww = dg[row + ld] * pp - a1; w[row + ld] = ww; v1 += ww * pp;
            pp = p[row + 2 * ld]; ww = dg[row + 2 * ld] * pp - a2; w[row + 2 * ld] = ww; v2 += ww * pp;
        }
    }
    warp_sum3(v0, v1, v2, out);
}

// B: row per thread, entries fully unrolled in groups of 8 with predication (all loads of a group issued first)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_rows_full(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                                          const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w,
                                                          int N, int ld, int nS, double* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    double v0 = 0, v1 = 0, v2 = 0;
    for (int s = warp; s < nS; s += nW) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0, a1 = 0, a2 = 0;
        for (int k0 = 0; k0 < width; k0 += 8) {
            int cc[8]; double e[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { const bool ok = k0 + k < width; const int idx = base + 32 * (ok ? k0 + k : k0) + lane; cc[k] = col[idx]; e[k] = ok ? a[idx] : 0.0; }
#pragma unroll
            for (int k = 0; k < 8; k++) { a0 += e[k] * p[cc[k]]; a1 += e[k] * p[cc[k] + ld]; a2 += e[k] * p[cc[k] + 2 * ld]; }
        }
        if (row < N) {
            double pp = p[row], ww = dg[row] * pp - a0; w[row] = ww; v0 += ww * pp;
            pp = p[row + ld]; ww = dg[row + ld] * pp - a1; w[row + ld] = ww; v1 += ww * pp;
            pp = p[row + 2 * ld]; ww = dg[row + 2 * ld] * pp - a2; w[row + 2 * ld] = ww; v2 += ww * pp;
        }
    }
    warp_sum3(v0, v1, v2, out);
}

// C: entry per thread: block of 256 = 8 warps works on one slice, warp k takes entry k of the 32 rows,
//    partial sums meet in shared memory (double buffered), warps 0..2 finish one component each
__global__ void __launch_bounds__(256) k_entry(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                               const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w,
                                               int N, int ld, int nS, double* out) {
    __shared__ double sh[2][3][8][32];
    const int lane = threadIdx.x & 31, k = threadIdx.x >> 5;
    double v = 0;
    int buf = 0;
    for (int s = blockIdx.x; s < nS; s += gridDim.x, buf ^= 1) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0, a1 = 0, a2 = 0;
        for (int kk = k; kk < width; kk += 8) {
            const int idx = base + 32 * kk + lane; const int cc = col[idx]; const double e = a[idx];
            a0 += e * p[cc]; a1 += e * p[cc + ld]; a2 += e * p[cc + 2 * ld];
        }
        sh[buf][0][k][lane] = a0; sh[buf][1][k][lane] = a1; sh[buf][2][k][lane] = a2;
        __syncthreads();
        if (k < 3 && row < N) {
            double acc = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) acc += sh[buf][k][j][lane];
            const double pp = p[row + k * ld], ww = dg[row + k * ld] * pp - acc;
            w[row + k * ld] = ww; v += ww * pp;
        }
    }
    if (k < 3) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(~0u, v, o); if (lane == 0) atomicAdd(out + k, v); }
}

// D: component per warp (three consecutive warps share a slice)
__global__ void __launch_bounds__(192) k_cpw(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                             const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w,
                                             int N, int ld, int nS, double* out) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, q = wib % 3, sub = wib / 3;
    const double* pq = p + (size_t)q * ld; const double* dq = dg + (size_t)q * ld; double* wq = w + (size_t)q * ld;
    double v = 0;
    for (int s = blockIdx.x * 2 + sub; s < nS; s += gridDim.x * 2) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0;
        const int* cp = col + base + lane; const double* ap = a + base + lane;
#pragma unroll 4
        for (int k = 0; k < width; k++) a0 += ap[32 * k] * pq[cp[32 * k]];
        if (row < N) { const double pp = pq[row], ww = dq[row] * pp - a0; wq[row] = ww; v += ww * pp; }
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(~0u, v, o);
    if (lane == 0) atomicAdd(out + q, v);
}

// E: scalar SpMV (one vector) for reference
__global__ void __launch_bounds__(256) k_one(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                             const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w, int N, int nS) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nS; s += nW) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0;
        const int* cp = col + base + lane; const double* ap = a + base + lane;
#pragma unroll 2
        for (int k = 0; k < width; k++) a0 += ap[32 * k] * p[cp[32 * k]];
        if (row < N) w[row] = dg[row] * p[row] - a0;
    }
}

// F: row per thread, interleaved (AoS-3) vectors: one 24-byte gather per neighbour
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_rows_aos(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                                         const double* __restrict__ dg3, const double* __restrict__ p3, double* __restrict__ w3,
                                                         int N, int nS, double* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    double v0 = 0, v1 = 0, v2 = 0;
    for (int s = warp; s < nS; s += nW) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        double a0 = 0, a1 = 0, a2 = 0;
        const int* cp = col + base + lane; const double* ap = a + base + lane;
#pragma unroll U
        for (int k = 0; k < width; k++) { const int cc = cp[32 * k]; const double e = ap[32 * k]; const double* pp = p3 + 3 * (size_t)cc; a0 += e * pp[0]; a1 += e * pp[1]; a2 += e * pp[2]; }
        if (row < N) {
            const double* pp = p3 + 3 * (size_t)row; const double* dd = dg3 + 3 * (size_t)row; double* ww = w3 + 3 * (size_t)row;
            const double w0 = dd[0] * pp[0] - a0, w1 = dd[1] * pp[1] - a1, w2 = dd[2] * pp[2] - a2;
            ww[0] = w0; ww[1] = w1; ww[2] = w2; v0 += w0 * pp[0]; v1 += w1 * pp[1]; v2 += w2 * pp[2];
        }
    }
    warp_sum3(v0, v1, v2, out);
}

int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 800, ny = argc > 2 ? atoi(argv[2]) : 100, nz = argc > 3 ? atoi(argv[3]) : 100;
    const int N = nx * ny * nz, nS = (N + 31) / 32;
    const int nB = 2 * (ny * nz + nx * nz + nx * ny);
    const int ld = ((N + nB + 31) / 32) * 32;
    std::vector<int> cnt(N), sp(nS + 1, 0);
    auto nbrs = [&](int c, int* out) {
        const int i = c % nx, j = (c / nx) % ny, k = c / (nx * ny); int n = 0, b = 0;
        if (k > 0) out[n++] = c - nx * ny; if (j > 0) out[n++] = c - nx; if (i > 0) out[n++] = c - 1;
        if (i < nx - 1) out[n++] = c + 1; if (j < ny - 1) out[n++] = c + nx; if (k < nz - 1) out[n++] = c + nx * ny;
        b = (k == 0) + (j == 0) + (i == 0) + (i == nx - 1) + (j == ny - 1) + (k == nz - 1);
        return n + 100 * b;
    };
    int tmp[8];
    for (int c = 0; c < N; c++) { int r = nbrs(c, tmp); cnt[c] = r % 100 + r / 100; }
    for (int s = 0; s < nS; s++) { int w = 0; for (int r = s * 32; r < std::min(N, s * 32 + 32); r++) w = std::max(w, cnt[r]); sp[s + 1] = sp[s] + 32 * w; }
    const size_t nE = sp[nS];
    std::vector<int> col(nE); std::vector<double> a(nE, 0.0);
    double nnz = 0; int bslot = N;
    for (int s = 0; s < nS; s++) {
        const int w = (sp[s + 1] - sp[s]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            const int P = s * 32 + lane; int n = 0, nb = 0;
            if (P < N) { int r = nbrs(P, tmp); n = r % 100; nb = r / 100; }
            for (int k = 0; k < w; k++) {
                const size_t E = (size_t)sp[s] + 32 * k + lane;
                if (P >= N) { col[E] = 0; continue; }
                if (k < n) { col[E] = tmp[k]; a[E] = 1.0 + 1e-3 * ((P + tmp[k]) % 13); nnz += 1; }
                else if (k < n + nb) { col[E] = bslot++; a[E] = 0.0; }
                else col[E] = P;
            }
        }
    }
    std::vector<double> dg(3 * (size_t)ld, 7.0), p(3 * (size_t)ld, 0.0);
    for (int q = 0; q < 3; q++) for (int i = 0; i < N; i++) p[(size_t)q * ld + i] = 1.0 + 1e-3 * ((i * 7 + q) % 11);
    int *dsp, *dcol; double *da, *ddg, *dp, *dw, *dout;
    CK(cudaMalloc(&dsp, sp.size() * 4)); CK(cudaMalloc(&dcol, nE * 4)); CK(cudaMalloc(&da, nE * 8));
    CK(cudaMalloc(&ddg, dg.size() * 8)); CK(cudaMalloc(&dp, p.size() * 8)); CK(cudaMalloc(&dw, p.size() * 8)); CK(cudaMalloc(&dout, 64));
    CK(cudaMemcpy(dsp, sp.data(), sp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcol, col.data(), nE * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, a.data(), nE * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ddg, dg.data(), dg.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, p.data(), p.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemset(dw, 0, p.size() * 8)); CK(cudaMemset(dout, 0, 64));
    int dev = 0, nSM = 148; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev)); nSM = prop.multiProcessorCount;
    const double bytes3 = 12.0 * nnz + 72.125 * N, bytes1 = 12.0 * nnz + 24.125 * N;
    printf("N=%d nE=%zu nnz=%.0f algorithmic bytes: 3-vector %.1f MB, 1-vector %.1f MB, SMs %d\n", N, nE, nnz, bytes3 / 1e6, bytes1 / 1e6, nSM);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<double> ref(3 * (size_t)ld), got(3 * (size_t)ld);
    bool haveRef = false;
    auto run = [&](const char* name, double bytes, auto launch, bool check) {
        CK(cudaMemset(dw, 0, p.size() * 8));
        for (int i = 0; i < 3; i++) launch();
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int i = 0; i < reps; i++) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        CK(cudaGetLastError());
        double err = -1;
        if (check) {
            CK(cudaMemcpy(got.data(), dw, got.size() * 8, cudaMemcpyDeviceToHost));
            if (!haveRef) { ref = got; haveRef = true; err = 0; }
            else { err = 0; for (size_t i = 0; i < got.size(); i++) err = std::max(err, std::abs(got[i] - ref[i])); }
        }
        printf("%-28s %8.4f ms  %8.1f GB/s  (%.3f of 6557.8)  maxdiff %.2e\n", name, ms, bytes / ms / 1e6, bytes / ms / 1e6 / 6557.8, err);
    };
    const int gM = nSM * 8;
    run("rows u2 (56r)", bytes3, [&] { k_rows<2, 1><<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows u1", bytes3, [&] { k_rows<1, 1><<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows u4 minb6", bytes3, [&] { k_rows<4, 6><<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows u2 minb5", bytes3, [&] { k_rows<2, 5><<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows u7", bytes3, [&] { k_rows<7, 1><<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows full8 minb2", bytes3, [&] { k_rows_full<2><<<nSM * 2, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows full8 minb3", bytes3, [&] { k_rows_full<3><<<nSM * 3, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("rows full8 minb4", bytes3, [&] { k_rows_full<4><<<nSM * 4, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("entry/smem", bytes3, [&] { k_entry<<<nSM * 8, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("entry/smem x16", bytes3, [&] { k_entry<<<nSM * 16, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("component-per-warp", bytes3, [&] { k_cpw<<<nSM * 10, 192>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); }, true);
    run("3 x one-vector", bytes3, [&] { for (int q = 0; q < 3; q++) k_one<<<gM, 256>>>(dsp, dcol, da, ddg + (size_t)q * ld, dp + (size_t)q * ld, dw + (size_t)q * ld, N, nS); }, true);
    run("one-vector", bytes1, [&] { k_one<<<gM, 256>>>(dsp, dcol, da, ddg, dp, dw, N, nS); }, false);
    // interleaved layout
    {
        std::vector<double> p3(3 * (size_t)ld), d3(3 * (size_t)ld, 7.0);
        for (int q = 0; q < 3; q++) for (int i = 0; i < ld; i++) p3[3 * (size_t)i + q] = p[(size_t)q * ld + i];
        double *dp3, *dd3, *dw3; CK(cudaMalloc(&dp3, p3.size() * 8)); CK(cudaMalloc(&dd3, p3.size() * 8)); CK(cudaMalloc(&dw3, p3.size() * 8));
        CK(cudaMemcpy(dp3, p3.data(), p3.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dd3, d3.data(), p3.size() * 8, cudaMemcpyHostToDevice));
        run("rows AoS-3 u2", bytes3, [&] { k_rows_aos<2, 1><<<gM, 256>>>(dsp, dcol, da, dd3, dp3, dw3, N, nS, dout); }, false);
        run("rows AoS-3 u4", bytes3, [&] { k_rows_aos<4, 1><<<gM, 256>>>(dsp, dcol, da, dd3, dp3, dw3, N, nS, dout); }, false);
        run("rows AoS-3 u7", bytes3, [&] { k_rows_aos<7, 1><<<gM, 256>>>(dsp, dcol, da, dd3, dp3, dw3, N, nS, dout); }, false);
    }
    return 0;
}
