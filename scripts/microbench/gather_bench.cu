// gather_bench.cu -- which asynchronous mechanism feeds a random 80-byte row gather fastest on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_bench gather_bench.cu
//   ./gather_bench [n_rows=10000000] [L=17] [a=3.94] [reps=5]
// Workload: Y_i = sum_k val[k] * D[idx[k], :] over rows of L nonzeros (the gather pass CD = C*D of the inner iteration with the
// row-length structure taken out), D = n x 10 doubles.  Column distribution j = n*u^a: a = 1 is uniform, a = 3.94 reproduces
// the hub statistics of the C5 graph (52.7 % of the gathers in the top 8 % of the rows, 27.6 % in the top 0.8 %).
// Variants:
//   reg   register gathers (LDG.128), one lane group of 5 lanes per row, 8 nonzeros in flight per lane (round-1 structure)
//   bulk  per-warp tile pipeline: ptr/idx/val tile by cp.async.bulk, ROW GATHERS by cp.async.bulk (80 B each) -> mbarrier
//   ldgsts per-warp tile pipeline: ptr/idx/val tile by cp.async.bulk, row gathers by 16-byte cp.async + cp.async.mbarrier.arrive
//   stage per-warp tile pipeline: ptr/idx/val tile by cp.async.bulk, register gathers out of the staged indices
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int R = 10;
typedef long long i64;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void cpasync16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Tile { int row0, nrows, nz0, nnz; };

// ---------------------------------------------------------------- register gathers
template <int NB, int MINB>
__global__ void __launch_bounds__(256, MINB) k_reg(i64 nrows, const int *__restrict__ ptr, const int *__restrict__ idx, const double *__restrict__ val,
                                             const double *__restrict__ D, double *__restrict__ Y) {
    const int G = 5, gpb = 256 / G;
    const int gib = threadIdx.x / G, lg = threadIdx.x - gib * G;
    if (gib >= gpb) return;
    for (i64 i = (i64)blockIdx.x * gpb + gib; i < nrows; i += (i64)gridDim.x * gpb) {
        const int beg = ptr[i], end = ptr[i + 1];
        double2 acc = make_double2(0, 0);
        for (int k0 = beg; k0 < end; k0 += NB) {
            int c[NB]; double v[NB]; double2 g[NB];
#pragma unroll
            for (int j = 0; j < NB; j++) { const int k = k0 + j < end ? k0 + j : end - 1; c[j] = __ldg(idx + k); v[j] = __ldg(val + k); }
#pragma unroll
            for (int j = 0; j < NB; j++) g[j] = __ldg(reinterpret_cast<const double2 *>(D + (size_t)c[j] * R) + lg);
#pragma unroll
            for (int j = 0; j < NB; j++) if (k0 + j < end) { acc.x += v[j] * g[j].x; acc.y += v[j] * g[j].y; }
        }
        reinterpret_cast<double2 *>(Y + (size_t)i * R)[lg] = acc;
    }
}

// ---------------------------------------------------------------- per-warp tile pipeline
// MODE 1 bulk row gathers, 2 LDGSTS row gathers, 3 register gathers from staged indices
// per warp: rows[NRS][T][R] doubles | val[NIV][TC] doubles | idx[NIV][TC] ints | ptr[NIV][TC] ints | barriers
template <int T, int NIV, int NRS>
__host__ __device__ constexpr size_t warp_smem_bytes() { return (size_t)NRS * T * R * 8 + (size_t)NIV * (T + 8) * 16 + 128; }

template <int MODE, int T, int NIV, int NRS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tile(i64 ntiles, const Tile *__restrict__ tiles, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                      const double *__restrict__ val, const double *__restrict__ D, double *__restrict__ Y) {
    constexpr int TC = T + 8;   // staged copies start at an index rounded down to a multiple of 4 and are rounded up to 16 bytes
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr size_t rows_bytes = (size_t)NRS * T * R * 8, val_bytes = (size_t)NIV * TC * 8, i_bytes = (size_t)NIV * TC * 4;
    unsigned char *base = smem_raw + (size_t)warp * warp_smem_bytes<T, NIV, NRS>();
    double *rows_s = reinterpret_cast<double *>(base);
    double *val_s = reinterpret_cast<double *>(base + rows_bytes);
    int *idx_s = reinterpret_cast<int *>(base + rows_bytes + val_bytes);
    int *ptr_s = reinterpret_cast<int *>(base + rows_bytes + val_bytes + i_bytes);
    uint64_t *bar_iv = reinterpret_cast<uint64_t *>(base + rows_bytes + val_bytes + 2 * i_bytes);
    uint64_t *bar_g = bar_iv + NIV;
    if (lane == 0) {
        for (int s = 0; s < NIV; s++) mbar_init(bar_iv + s, 1);
        for (int s = 0; s < NRS; s++) mbar_init(bar_g + s, MODE == 2 ? 32 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const i64 gw = (i64)blockIdx.x * WARPS + warp, nw = (i64)gridDim.x * WARPS;
    const i64 nk = gw < ntiles ? (ntiles - gw + nw - 1) / nw : 0;   // tiles of this warp: gw, gw+nw, ...

    auto issue_iv = [&](i64 k) {   // lane 0 only
        const Tile t = tiles[gw + k * nw];
        const int s = (int)(k % NIV);
        const int a0 = t.nz0 & ~3, cnt = ((t.nz0 + t.nnz - a0) + 3) & ~3;
        const int p0 = t.row0 & ~3, pcnt = ((t.row0 + t.nrows + 1 - p0) + 3) & ~3;
        fence_proxy_async();
        mbar_expect_tx(bar_iv + s, (uint32_t)cnt * 12u + (uint32_t)pcnt * 4u);
        bulk_g2s(idx_s + s * TC, idx + a0, (uint32_t)cnt * 4u, bar_iv + s);
        bulk_g2s(val_s + s * TC, val + a0, (uint32_t)cnt * 8u, bar_iv + s);
        bulk_g2s(ptr_s + s * TC, ptr + p0, (uint32_t)pcnt * 4u, bar_iv + s);
    };
    auto issue_gathers = [&](i64 k) {   // whole warp
        const Tile t = tiles[gw + k * nw];
        const int s = (int)(k % NIV), sr = (int)(k % NRS);
        mbar_wait(bar_iv + s, (uint32_t)((k / NIV) & 1));
        const int *is = idx_s + s * TC + (t.nz0 & 3);
        double *rs = rows_s + (size_t)sr * T * R;
        if (MODE == 1) {
            if (lane == 0) { fence_proxy_async(); mbar_expect_tx(bar_g + sr, (uint32_t)t.nnz * (R * 8u)); }
            __syncwarp();
            for (int j = lane; j < t.nnz; j += 32) bulk_g2s(rs + (size_t)j * R, D + (size_t)is[j] * R, R * 8u, bar_g + sr);
        } else if (MODE == 2) {
            const int pieces = t.nnz * (R / 2);
            for (int e = lane; e < pieces; e += 32) {
                const int j = e / (R / 2), p = e - j * (R / 2);
                cpasync16(reinterpret_cast<double2 *>(rs) + e, reinterpret_cast<const double2 *>(D + (size_t)is[j] * R) + p);
            }
            cpasync_arrive_noinc(bar_g + sr);
        }
    };
    auto consume = [&](i64 k) {
        const Tile t = tiles[gw + k * nw];
        const int s = (int)(k % NIV), sr = (int)(k % NRS);
        const int *is = idx_s + s * TC + (t.nz0 & 3);
        const double *vs = val_s + s * TC + (t.nz0 & 3);
        const int *ps = ptr_s + s * TC + (t.row0 & 3);
        const double *rs = rows_s + (size_t)sr * T * R;
        const int g = lane / 5, lg = lane - g * 5;
        if (MODE == 3) mbar_wait(bar_iv + s, (uint32_t)((k / NIV) & 1));
        else mbar_wait(bar_g + sr, (uint32_t)((k / NRS) & 1));
        if (g < 6)
        for (int q = g; q < t.nrows; q += 6) {
            const int b = ps[q] - t.nz0, e = ps[q + 1] - t.nz0;
            double2 acc = make_double2(0, 0);
            if (MODE == 3) {
                for (int j0 = b; j0 < e; j0 += 8) {
                    double2 gq[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) { const int j = j0 + u < e ? j0 + u : e - 1; gq[u] = __ldg(reinterpret_cast<const double2 *>(D + (size_t)is[j] * R) + lg); }
#pragma unroll
                    for (int u = 0; u < 8; u++) if (j0 + u < e) { const double v = vs[j0 + u]; acc.x += v * gq[u].x; acc.y += v * gq[u].y; }
                }
            } else {
#pragma unroll 4
                for (int j = b; j < e; j++) {
                    const double v = vs[j];
                    const double2 x = reinterpret_cast<const double2 *>(rs + (size_t)j * R)[lg];
                    acc.x += v * x.x; acc.y += v * x.y;
                }
            }
            reinterpret_cast<double2 *>(Y + (size_t)(t.row0 + q) * R)[lg] = acc;
        }
        __syncwarp();
    };

    if (lane == 0) for (i64 k = 0; k < NIV && k < nk; k++) issue_iv(k);
    if (MODE != 3) for (i64 k = 0; k < NRS - 1 && k < nk; k++) issue_gathers(k);
    for (i64 k = 0; k < nk; k++) {
        if (MODE != 3 && k + NRS - 1 < nk) issue_gathers(k + NRS - 1);   // rows stage (k-1) % NRS was released by consume(k-1)
        consume(k);
        if (lane == 0 && k + NIV < nk) issue_iv(k + NIV);                // ptr/idx/val stage k % NIV is free now
    }
}

// ---------------------------------------------------------------- host
static double urand(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

template <int MODE, int T, int NIV, int NRS, int WARPS>
float run_tile(const char *name, int reps, i64 ntiles, const Tile *tiles, const int *ptr, const int *idx, const double *val, const double *D, double *Y,
               i64 nnz, const double *Yref, i64 ylen) {
    const size_t smem = warp_smem_bytes<T, NIV, NRS>() * WARPS;
    auto kern = k_tile<MODE, T, NIV, NRS, WARPS>;
    if (smem > 227 * 1024) { printf("%-28s T=%3d NRS=%d warps=%2d does not fit (smem %zu)\n", name, T, NRS, WARPS, smem); return 0; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, smem));
    if (occ < 1) { printf("%-28s does not fit (smem %zu)\n", name, smem); return 0; }
    const int grid = 148 * occ;
    CK(cudaMemset(Y, 0xff, ylen * sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < reps + 1; it++) {
        CK(cudaEventRecord(e0));
        kern<<<grid, WARPS * 32, smem>>>(ntiles, tiles, ptr, idx, val, D, Y);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    // compare with the reference result (same accumulation order per row -> bitwise)
    std::vector<double> h(1 << 20);
    CK(cudaMemcpy(h.data(), Y + (ylen - (i64)h.size()), h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<double> hr(1 << 20);
    CK(cudaMemcpy(hr.data(), Yref + (ylen - (i64)hr.size()), hr.size() * sizeof(double), cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < h.size(); i++) if (!(h[i] == hr[i])) bad++;
    CK(cudaMemcpy(h.data(), Y, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), Yref, hr.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); i++) if (!(h[i] == hr[i])) bad++;
    printf("%-28s T=%3d NIV=%d NRS=%d warps=%2d occ=%d smem/CTA=%6zu  %8.3f ms  %7.2f rows/ns  %7.1f GB/s gathered  mismatches=%zu\n", name, T, NIV, NRS, WARPS, occ,
           smem, best, nnz / (best * 1e6), nnz * 80.0 / (best * 1e6), bad);
    fflush(stdout);
    return best;
}

int main(int argc, char **argv) {
    const i64 n = argc > 1 ? atoll(argv[1]) : 10000000;
    const int L = argc > 2 ? atoi(argv[2]) : 17;
    const double a = argc > 3 ? atof(argv[3]) : 3.94;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    const i64 nnz = n * L;
    printf("n=%lld L=%d nnz=%lld a=%.2f\n", n, L, nnz, a);
    std::vector<int> hptr(n + 1 + 8), hidx(nnz + 16);
    std::vector<double> hval(nnz + 16, 0.0);
    uint64_t seed = 0x9E3779B97F4A7C15ull;
    for (i64 i = 0; i <= n; i++) hptr[i] = (int)(i * L);
    for (i64 i = n + 1; i < n + 9; i++) hptr[i] = (int)nnz;
    for (i64 k = 0; k < nnz; k++) {
        i64 j = (i64)(n * pow(urand(seed), a));
        hidx[k] = (int)(j < n ? j : n - 1);
        hval[k] = 0.5 + urand(seed);
    }
    for (i64 k = nnz; k < nnz + 16; k++) hidx[k] = 0;
    int *ptr, *idx; double *val, *D, *Y, *Yref;
    CK(cudaMalloc(&ptr, hptr.size() * 4)); CK(cudaMalloc(&idx, hidx.size() * 4)); CK(cudaMalloc(&val, hval.size() * 8));
    CK(cudaMalloc(&D, n * R * 8)); CK(cudaMalloc(&Y, n * R * 8)); CK(cudaMalloc(&Yref, n * R * 8));
    CK(cudaMemcpy(ptr, hptr.data(), hptr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(val, hval.data(), hval.size() * 8, cudaMemcpyHostToDevice));
    {
        std::vector<double> hD((size_t)n * R);
        for (size_t i = 0; i < hD.size(); i++) hD[i] = urand(seed) - 0.5;
        CK(cudaMemcpy(D, hD.data(), hD.size() * 8, cudaMemcpyHostToDevice));
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reference / variant "reg" with different register budgets
    auto run_reg = [&](const char *name, auto kern, int blocks_per_sm) {
        float best = 1e30f;
        for (int it = 0; it < reps + 1; it++) {
            CK(cudaEventRecord(e0));
            kern<<<148 * blocks_per_sm, 256>>>(n, ptr, idx, val, D, Yref);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it > 0 && ms < best) best = ms;
        }
        printf("%-28s %8.3f ms  %7.2f rows/ns  %7.1f GB/s gathered\n", name, best, nnz / (best * 1e6), nnz * 80.0 / (best * 1e6));
        fflush(stdout);
    };
    run_reg("reg NB=8 minb=8 (32 regs)", k_reg<8, 8>, 16);
    run_reg("reg NB=8 minb=1", k_reg<8, 1>, 8);
    run_reg("reg NB=8 minb=2", k_reg<8, 2>, 8);
    run_reg("reg NB=8 minb=3", k_reg<8, 3>, 6);
    run_reg("reg NB=8 minb=4", k_reg<8, 4>, 8);
    run_reg("reg NB=4 minb=4", k_reg<4, 4>, 8);
    run_reg("reg NB=4 minb=6", k_reg<4, 6>, 6);
    run_reg("reg NB=16 minb=2", k_reg<16, 2>, 4);
    run_reg("reg NB=8 minb=8 (32 regs)", k_reg<8, 8>, 16);
    auto make_tiles = [&](int T, Tile **dt) -> i64 {
        const int rpt = T / L > 0 ? T / L : 1;
        std::vector<Tile> ht;
        for (i64 r0 = 0; r0 < n; r0 += rpt) {
            const int nr = (int)((n - r0) < rpt ? (n - r0) : rpt);
            ht.push_back(Tile{(int)r0, nr, (int)(r0 * L), nr * L});
        }
        CK(cudaMalloc(dt, ht.size() * sizeof(Tile)));
        CK(cudaMemcpy(*dt, ht.data(), ht.size() * sizeof(Tile), cudaMemcpyHostToDevice));
        return (i64)ht.size();
    };
    const i64 ylen = n * R;
    Tile *t64, *t128, *t256;
    const i64 n64 = make_tiles(64, &t64), n128 = make_tiles(128, &t128), n256 = make_tiles(256, &t256);
#define RUN(MODE, T, NIV, NRS, W, tiles, nt) run_tile<MODE, T, NIV, NRS, W>(#MODE " " #T, reps, nt, tiles, ptr, idx, val, D, Y, nnz, Yref, ylen)
    Tile *t32; const i64 n32 = make_tiles(32, &t32);
    printf("-- bulk row gathers (cp.async.bulk 80 B per row)\n");
    RUN(1, 64, 4, 2, 8, t64, n64);
    RUN(1, 64, 4, 2, 14, t64, n64);
    RUN(1, 32, 4, 2, 16, t32, n32);
    RUN(1, 32, 4, 3, 16, t32, n32);
    RUN(1, 128, 4, 2, 7, t128, n128);
    printf("-- LDGSTS row gathers (cp.async 16 B pieces)\n");
    RUN(2, 64, 4, 2, 8, t64, n64);
    RUN(2, 64, 4, 2, 14, t64, n64);
    RUN(2, 64, 4, 3, 10, t64, n64);
    RUN(2, 32, 4, 2, 16, t32, n32);
    RUN(2, 32, 4, 3, 16, t32, n32);
    RUN(2, 32, 4, 4, 16, t32, n32);
    RUN(2, 128, 4, 2, 7, t128, n128);
    RUN(2, 128, 4, 3, 5, t128, n128);
    RUN(2, 256, 4, 2, 4, t256, n256);
    printf("-- register gathers out of staged indices\n");
    RUN(3, 64, 3, 1, 8, t64, n64);
    RUN(3, 64, 3, 1, 16, t64, n64);
    RUN(3, 128, 3, 1, 16, t128, n128);
    RUN(3, 128, 3, 1, 32, t128, n128);
    RUN(3, 32, 3, 1, 32, t32, n32);
    return 0;
}
