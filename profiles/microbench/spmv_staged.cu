// spmv_staged.cu -- round 2: the staging question (north_star: "SpMV and stencil loops use shared-memory or TMA staging of
// the cell values each tile reuses").  Three-component SELL-32 SpMV y = A x + sums (the shape of k_amul3 / k_amg_step):
//   gather8        the shipped mapping of round 1: row per thread, groups of eight entries, cached global gathers
//   gather exact   the same without the padded gathers of a partly filled group, row-local loads issued first
//   staged u32/u16 a block owns a TILE of R consecutive rows; the distinct columns its rows reference are a few contiguous
//                  segments of x (set-up pass per mesh, generic: no structured-grid knowledge), which one thread brings into
//                  shared memory with cp.async.bulk (TMA 1-D bulk copies completing on an mbarrier); the row loop then streams
//                  (local column, coefficient) and gathers from shared memory.  u16: the local column fits 16 bits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o spmv_staged spmv_staged.cu && ./spmv_staged 800 100 100
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

__device__ __forceinline__ void warp_sum3(double v0, double v1, double v2, double* out) {
    for (int o = 16; o > 0; o >>= 1) { v0 += __shfl_xor_sync(~0u, v0, o); v1 += __shfl_xor_sync(~0u, v1, o); v2 += __shfl_xor_sync(~0u, v2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, v0); atomicAdd(out + 1, v1); atomicAdd(out + 2, v2); }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather8(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
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

// no gathers for the padded tail of a group (the predicate is warp-uniform), row-local loads first
template <int MINB>
__global__ void __launch_bounds__(256, MINB) k_gather_exact(const int* __restrict__ sp, const int* __restrict__ col, const double* __restrict__ a,
                                                             const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w,
                                                             int N, int ld, int nS, double* out) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
    double v0 = 0, v1 = 0, v2 = 0;
    for (int s = warp; s < nS; s += nW) {
        const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
        const int r = row < N ? row : 0;
        const double p0 = p[r], p1 = p[r + ld], p2 = p[r + 2 * ld], d0 = dg[r], d1 = dg[r + ld], d2 = dg[r + 2 * ld];
        double a0 = 0, a1 = 0, a2 = 0;
        for (int k0 = 0; k0 < width; k0 += 8) {
            int cc[8]; double e[8];
            const int m = width - k0;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < m) { const int idx = base + 32 * (k0 + k) + lane; cc[k] = col[idx]; e[k] = a[idx]; }
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < m) { a0 += e[k] * p[cc[k]]; a1 += e[k] * p[cc[k] + ld]; a2 += e[k] * p[cc[k] + 2 * ld]; }
        }
        if (row < N) {
            double ww = d0 * p0 - a0; w[row] = ww; v0 += ww * p0;
            ww = d1 * p1 - a1; w[row + ld] = ww; v1 += ww * p1;
            ww = d2 * p2 - a2; w[row + 2 * ld] = ww; v2 += ww * p2;
        }
    }
    warp_sum3(v0, v1, v2, out);
}

// ---- staged ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t phase) {
    asm volatile(
        "{\n.reg .pred P1;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(b)),
        "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(b)) : "memory");
}

struct Seg { int gStart, len, sOff, pad; };

template <class LC, int R>
__global__ void __launch_bounds__(256, 2) k_staged(const int* __restrict__ sp, const LC* __restrict__ lcol, const double* __restrict__ a,
                                                   const double* __restrict__ dg, const double* __restrict__ p, double* __restrict__ w, int N, int ld,
                                                   int nS, const int* __restrict__ tileSegPtr, const Seg* __restrict__ segs, const int* __restrict__ ownOff,
                                                   int nTiles, int Smax, double* out) {
    extern __shared__ __align__(128) double xs[];       // [3][Smax]
    __shared__ __align__(8) uint64_t mbar;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    uint32_t phase = 0;
    double v0 = 0, v1 = 0, v2 = 0;
    const double* xs0 = xs; const double* xs1 = xs + Smax; const double* xs2 = xs + 2 * Smax;
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
        if (threadIdx.x == 0) {
            const int s0 = tileSegPtr[t], s1 = tileSegPtr[t + 1];
            uint32_t bytes = 0;
            for (int g = s0; g < s1; g++) bytes += 3u * 8u * (uint32_t)segs[g].len;
            mbar_expect_tx(&mbar, bytes);
            for (int g = s0; g < s1; g++) {
                const Seg sg = segs[g];
#pragma unroll
                for (int q = 0; q < 3; q++) bulk_g2s(xs + (size_t)q * Smax + sg.sOff, p + (size_t)q * ld + sg.gStart, 8u * (uint32_t)sg.len, &mbar);
            }
        }
        mbar_wait(&mbar, phase); phase ^= 1;
        const int own = ownOff[t];
        for (int sl = wib; sl < R / 32; sl += 8) {
            const int s = t * (R / 32) + sl;
            if (s >= nS) break;
            const int base = sp[s], width = (sp[s + 1] - base) >> 5, row = s * 32 + lane;
            const int r = row < N ? row : 0;
            const double d0 = dg[r], d1 = dg[r + ld], d2 = dg[r + 2 * ld];
            double a0 = 0, a1 = 0, a2 = 0;
            for (int k0 = 0; k0 < width; k0 += 8) {
                int cc[8]; double e[8];
                const int m = width - k0;
#pragma unroll
                for (int k = 0; k < 8; k++) if (k < m) { const int idx = base + 32 * (k0 + k) + lane; cc[k] = (int)lcol[idx]; e[k] = a[idx]; }
#pragma unroll
                for (int k = 0; k < 8; k++) if (k < m) { a0 += e[k] * xs0[cc[k]]; a1 += e[k] * xs1[cc[k]]; a2 += e[k] * xs2[cc[k]]; }
            }
            if (row < N) {
                const int lo = own + sl * 32 + lane;
                const double p0 = xs0[lo], p1 = xs1[lo], p2 = xs2[lo];
                double ww = d0 * p0 - a0; w[row] = ww; v0 += ww * p0;
                ww = d1 * p1 - a1; w[row + ld] = ww; v1 += ww * p1;
                ww = d2 * p2 - a2; w[row + 2 * ld] = ww; v2 += ww * p2;
            }
        }
        __syncthreads();      // the tile's values are dead: the next bulk copies may overwrite them
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

    // ---- set-up pass of the staged variant (generic: works from the rows alone) -------------------------------------
    auto build_tiles = [&](int R, std::vector<int>& tileSegPtr, std::vector<Seg>& segs, std::vector<int>& ownOff, std::vector<int>& lcol, int& Smax) {
        const int nTiles = (N + R - 1) / R;
        tileSegPtr.assign(nTiles + 1, 0); ownOff.assign(nTiles, 0); lcol.assign(nE, 0); segs.clear(); Smax = 0;
        std::vector<int> cols;
        for (int t = 0; t < nTiles; t++) {
            const int r0 = t * R, r1 = std::min(N, r0 + R);
            cols.clear();
            for (int r = r0; r < r1; r++) cols.push_back(r);
            for (int s = r0 / 32; s < (r1 + 31) / 32; s++) {
                const int w = (sp[s + 1] - sp[s]) / 32;
                for (int k = 0; k < w; k++) for (int lane = 0; lane < 32; lane++) {
                    const size_t E = (size_t)sp[s] + 32 * k + lane;
                    if (s * 32 + lane < N && a[E] != 0.0) cols.push_back(col[E]);
                }
            }
            std::sort(cols.begin(), cols.end()); cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
            // contiguous segments (gaps of up to 16 values are fetched as well), 16-byte aligned start and length
            int sOff = 0;
            const int segBegin = (int)segs.size();
            for (size_t i = 0; i < cols.size();) {
                size_t j = i;
                while (j + 1 < cols.size() && cols[j + 1] - cols[j] <= 16) j++;
                int g0 = cols[i] & ~1, g1 = (cols[j] + 2) & ~1;
                if (g1 > ld) g1 = ld;
                segs.push_back(Seg{g0, g1 - g0, sOff, 0});
                sOff += g1 - g0;
                i = j + 1;
            }
            tileSegPtr[t + 1] = (int)segs.size();
            Smax = std::max(Smax, sOff);
            auto local = [&](int c) {
                for (int g = segBegin; g < (int)segs.size(); g++) if (c >= segs[g].gStart && c < segs[g].gStart + segs[g].len) return segs[g].sOff + c - segs[g].gStart;
                return 0;
            };
            ownOff[t] = local(r0);
            for (int s = r0 / 32; s < (r1 + 31) / 32; s++) {
                const int w = (sp[s + 1] - sp[s]) / 32;
                for (int k = 0; k < w; k++) for (int lane = 0; lane < 32; lane++) {
                    const size_t E = (size_t)sp[s] + 32 * k + lane;
                    lcol[E] = (s * 32 + lane < N && a[E] != 0.0) ? local(col[E]) : ownOff[t];
                }
            }
        }
        return nTiles;
    };

    int *dsp, *dcol; double *da, *ddg, *dp, *dw, *dout;
    CK(cudaMalloc(&dsp, sp.size() * 4)); CK(cudaMalloc(&dcol, nE * 4)); CK(cudaMalloc(&da, nE * 8));
    CK(cudaMalloc(&ddg, dg.size() * 8)); CK(cudaMalloc(&dp, p.size() * 8)); CK(cudaMalloc(&dw, p.size() * 8)); CK(cudaMalloc(&dout, 64));
    CK(cudaMemcpy(dsp, sp.data(), sp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dcol, col.data(), nE * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(da, a.data(), nE * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(ddg, dg.data(), dg.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, p.data(), p.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemset(dw, 0, p.size() * 8)); CK(cudaMemset(dout, 0, 64));
    int nSM = 148; cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); nSM = prop.multiProcessorCount;
    const double bytes3 = 12.0 * nnz + 72.125 * N;
    printf("N=%d nE=%zu nnz=%.0f algorithmic bytes (int32 columns): %.1f MB, SMs %d\n", N, nE, nnz, bytes3 / 1e6, nSM);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<double> ref(3 * (size_t)ld), got(3 * (size_t)ld);
    bool haveRef = false;
    auto run = [&](const char* name, double bytes, auto launch) {
        CK(cudaMemset(dw, 0, p.size() * 8));
        for (int i = 0; i < 3; i++) launch();
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int i = 0; i < reps; i++) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        CK(cudaGetLastError());
        CK(cudaMemcpy(got.data(), dw, got.size() * 8, cudaMemcpyDeviceToHost));
        double err = 0;
        if (!haveRef) { ref = got; haveRef = true; }
        else for (size_t i = 0; i < got.size(); i++) err = std::max(err, std::abs(got[i] - ref[i]));
        printf("%-34s %8.4f ms  %8.1f GB/s (of its own bytes %.1f MB)  maxdiff %.2e\n", name, ms, bytes / ms / 1e6, bytes / 1e6, err);
    };
    run("gather8 minb4 (shipped r1)", bytes3, [&] { k_gather8<4><<<nSM * 4, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); });
    run("gather exact minb4", bytes3, [&] { k_gather_exact<4><<<nSM * 4, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); });
    run("gather exact minb3", bytes3, [&] { k_gather_exact<3><<<nSM * 3, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); });
    run("gather exact minb4 x2 grid", bytes3, [&] { k_gather_exact<4><<<nSM * 8, 256>>>(dsp, dcol, da, ddg, dp, dw, N, ld, nS, dout); });

    auto staged = [&](auto tag, int R, const char* label) {
        using LC = decltype(tag);
        std::vector<int> tileSegPtr, ownOff, lcol; std::vector<Seg> segs; int Smax = 0;
        const int nTiles = build_tiles(R, tileSegPtr, segs, ownOff, lcol, Smax);
        Smax = (Smax + 15) & ~15;
        if (sizeof(LC) == 2 && Smax > 65535) { printf("%s: tile too wide for 16-bit columns\n", label); return; }
        const size_t smem = 3 * (size_t)Smax * 8;
        std::vector<LC> lc(nE);
        for (size_t i = 0; i < nE; i++) lc[i] = (LC)lcol[i];
        int *dtp, *down; Seg* dsg; LC* dlc;
        CK(cudaMalloc(&dtp, tileSegPtr.size() * 4)); CK(cudaMalloc(&down, ownOff.size() * 4)); CK(cudaMalloc(&dsg, segs.size() * sizeof(Seg))); CK(cudaMalloc(&dlc, nE * sizeof(LC)));
        CK(cudaMemcpy(dtp, tileSegPtr.data(), tileSegPtr.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(down, ownOff.data(), ownOff.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dsg, segs.data(), segs.size() * sizeof(Seg), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dlc, lc.data(), nE * sizeof(LC), cudaMemcpyHostToDevice));
        const double by = (8.0 + sizeof(LC)) * nnz + 72.125 * N;
        char name[128];
        int perSM = (int)std::min<size_t>(8, (220 * 1024) / (smem + 1024));
        snprintf(name, sizeof(name), "%s R=%d smem %zu KB x%d segs/tile %.1f", label, R, smem / 1024, perSM, segs.size() / (double)nTiles);
        if (perSM < 1) { printf("%s: does not fit\n", name); return; }
        if (R == 1024) {
            CK(cudaFuncSetAttribute(k_staged<LC, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            run(name, by, [&] { k_staged<LC, 1024><<<nSM * perSM, 256, smem>>>(dsp, dlc, da, ddg, dp, dw, N, ld, nS, dtp, dsg, down, nTiles, Smax, dout); });
        } else if (R == 512) {
            CK(cudaFuncSetAttribute(k_staged<LC, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            run(name, by, [&] { k_staged<LC, 512><<<nSM * perSM, 256, smem>>>(dsp, dlc, da, ddg, dp, dw, N, ld, nS, dtp, dsg, down, nTiles, Smax, dout); });
        } else {
            CK(cudaFuncSetAttribute(k_staged<LC, 2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            run(name, by, [&] { k_staged<LC, 2048><<<nSM * perSM, 256, smem>>>(dsp, dlc, da, ddg, dp, dw, N, ld, nS, dtp, dsg, down, nTiles, Smax, dout); });
        }
        cudaFree(dtp); cudaFree(down); cudaFree(dsg); cudaFree(dlc);
    };
    staged((int)0, 1024, "staged u32");
    staged((unsigned short)0, 1024, "staged u16");
    staged((unsigned short)0, 512, "staged u16");
    staged((unsigned short)0, 2048, "staged u16");
    return 0;
}
