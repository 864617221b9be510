// l2_probe.cu -- ceilings and cache-policy experiments for the random row gather of the gather pass (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o l2_probe l2_probe.cu
//   ./l2_probe [fetch_granularity_bytes]
// Q1  DRAM random-read ceiling with NO index dependency (row ids from a hash): 64 / 80 / 128-byte rows, uniform over 0.8-1.3 GB
// Q2  hub-distributed gather (j = n*u^3.94) with L2 policies: none | evict_last hubs + evict_first tail | the same with a
//     persisting-L2 carve-out | stream access-policy window over the hub prefix
// Q3  red.global.add.f64 throughput into an L2-resident 64 MB region (symmetric push of the (hub,tail) block)
// Q4  effect of cudaLimitMaxL2FetchGranularity (argv[1]) on all of the above
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
typedef long long i64;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// Q1: every lane group of G lanes reads one row of G*16 bytes per step, NB independent rows in flight
template <int G, int NB>
__global__ void __launch_bounds__(256) k_rand_read(i64 nrows_table, int stride_dbl, i64 reads_per_group, const double *__restrict__ D, double *__restrict__ out) {
    const int gpb = 256 / G;
    const int gib = threadIdx.x / G, lg = threadIdx.x - gib * G;
    if (gib >= gpb) return;
    const uint32_t gid = blockIdx.x * gpb + gib;
    double2 acc = make_double2(0, 0);
    for (i64 s = 0; s < reads_per_group; s += NB) {
        double2 g[NB];
#pragma unroll
        for (int j = 0; j < NB; j++) {
            const uint32_t h = hash32(gid * 0x9E3779B9u + (uint32_t)(s + j) * 0x85EBCA6Bu);
            const i64 row = (i64)(((unsigned long long)h * (unsigned long long)nrows_table) >> 32);
            g[j] = __ldg(reinterpret_cast<const double2 *>(D + row * stride_dbl) + lg);
        }
#pragma unroll
        for (int j = 0; j < NB; j++) { acc.x += g[j].x; acc.y += g[j].y; }
    }
    if (acc.x == 1.2345e300) out[gid] = acc.x + acc.y;
}

__device__ __forceinline__ unsigned long long pol_last() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ unsigned long long pol_first() { unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ double2 ld_hint(const double *p, unsigned long long pol) {
    double2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ int ldi_hint(const int *p, unsigned long long pol) { int v; asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v; }

// Q2: CSR-less gather of L rows per output row from an index stream; HINT 0 none, 1 evict_last for col < hot / evict_first otherwise, 2 no_allocate tail
template <int HINT>
__global__ void __launch_bounds__(256, 4) k_gather(i64 nrows, int L, int hot, const int *__restrict__ idx, const double *__restrict__ D, double *__restrict__ Y) {
    const int G = 5, gpb = 256 / G, R = 10;
    const int gib = threadIdx.x / G, lg = threadIdx.x - gib * G;
    if (gib >= gpb) return;
    const unsigned long long pl = pol_last(), pf = pol_first();
    for (i64 i = (i64)blockIdx.x * gpb + gib; i < nrows; i += (i64)gridDim.x * gpb) {
        const int *ix = idx + i * L;
        double2 acc = make_double2(0, 0);
        for (int k0 = 0; k0 < L; k0 += 8) {
            int c[8]; double2 g[8];
#pragma unroll
            for (int j = 0; j < 8; j++) c[j] = HINT ? ldi_hint(ix + min(k0 + j, L - 1), pf) : __ldg(ix + min(k0 + j, L - 1));
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double *p = D + (size_t)c[j] * R + lg * 2;
                if (HINT == 0) g[j] = __ldg(reinterpret_cast<const double2 *>(p));
                else g[j] = ld_hint(p, c[j] < hot ? pl : pf);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) if (k0 + j < L) { acc.x += g[j].x; acc.y += g[j].y; }
        }
        reinterpret_cast<double2 *>(Y + (size_t)i * R)[lg] = acc;
    }
}

// Q3: every lane group adds one 80-byte row into a random row of an L2-resident accumulator
__global__ void __launch_bounds__(256) k_red(i64 nrows_acc, i64 pushes_per_group, double *__restrict__ acc) {
    const int G = 5, gpb = 256 / G;
    const int gib = threadIdx.x / G, lg = threadIdx.x - gib * G;
    if (gib >= gpb) return;
    const uint32_t gid = blockIdx.x * gpb + gib;
    for (i64 s = 0; s < pushes_per_group; s++) {
        const uint32_t h = hash32(gid * 0x9E3779B9u + (uint32_t)s * 0x85EBCA6Bu);
        const i64 row = (i64)(((unsigned long long)h * (unsigned long long)nrows_acc) >> 32);
        double *p = acc + row * 10 + lg * 2;
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(1.0) : "memory");
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + 1), "d"(2.0) : "memory");
    }
}

static double urand(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) * (1.0 / 9007199254740992.0); }

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < reps + 1; it++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    if (argc > 1 && atoi(argv[1]) > 0) {
        CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1])));
    }
    size_t gran = 0; CK(cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity));
    int maxp = 0, l2 = 0; CK(cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, 0)); CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0));
    printf("L2 fetch granularity limit = %zu B, L2 = %.1f MB, max persisting = %.1f MB\n", gran, l2 / 1048576.0, maxp / 1048576.0);
    const i64 n = 10000000;
    double *D, *Y, *out;
    CK(cudaMalloc(&D, (size_t)n * 16 * 8)); CK(cudaMalloc(&Y, (size_t)n * 10 * 8)); CK(cudaMalloc(&out, 1 << 24));
    CK(cudaMemset(D, 0, (size_t)n * 16 * 8));
    // ---- Q1
    {
        const i64 total_rows = 1 << 27;   // 134 M row reads
        auto run = [&](const char *name, auto kern, int G, int stride, i64 table_rows) {
            const int gpb = 256 / G, grid = 148 * 8;
            const i64 per = total_rows / ((i64)grid * gpb);
            float ms = time_ms([&] { kern<<<grid, 256>>>(table_rows, stride, per, D, out); }, 3);
            const double rows = (double)per * grid * gpb;
            printf("Q1 %-44s %7.3f ms  %6.2f rows/ns  %7.1f GB/s useful\n", name, ms, rows / (ms * 1e6), rows * G * 16 / (ms * 1e6));
        };
        run("64 B rows, 64 B stride, NB=8", k_rand_read<4, 8>, 4, 8, n);
        run("64 B rows, 64 B stride, NB=16", k_rand_read<4, 16>, 4, 8, n);
        run("80 B rows, 80 B stride, NB=8", k_rand_read<5, 8>, 5, 10, n);
        run("80 B rows, 80 B stride, NB=16", k_rand_read<5, 16>, 5, 10, n);
        run("80 B rows, 128 B stride, NB=8", k_rand_read<5, 8>, 5, 16, n);
        run("128 B rows, 128 B stride, NB=8", k_rand_read<8, 8>, 8, 16, n);
        run("32 B rows, 32 B stride, NB=16", k_rand_read<2, 16>, 2, 4, n);
        run("80 B rows, 80 B stride, NB=8, 1M-row table (L2)", k_rand_read<5, 8>, 5, 10, 1000000);
        run("80 B rows, 80 B stride, NB=16, 1M-row table (L2)", k_rand_read<5, 16>, 5, 10, 1000000);
    }
    // ---- Q2
    {
        const int L = 17; const i64 nnz = n * L;
        std::vector<int> hidx(nnz);
        uint64_t seed = 0x9E3779B97F4A7C15ull;
        for (i64 k = 0; k < nnz; k++) { i64 j = (i64)(n * pow(urand(seed), 3.94)); hidx[k] = (int)(j < n ? j : n - 1); }
        int *idx; CK(cudaMalloc(&idx, nnz * 4)); CK(cudaMemcpy(idx, hidx.data(), nnz * 4, cudaMemcpyHostToDevice));
        const int hot = 800000;
        auto rep = [&](const char *name, float ms) { printf("Q2 %-60s %7.3f ms  %6.2f rows/ns\n", name, ms, nnz / (ms * 1e6)); fflush(stdout); };
        rep("no hints", time_ms([&] { k_gather<0><<<148 * 8, 256>>>(n, L, hot, idx, D, Y); }, 3));
        rep("evict_last hubs (800k rows) / evict_first tail, no carve-out", time_ms([&] { k_gather<1><<<148 * 8, 256>>>(n, L, hot, idx, D, Y); }, 3));
        rep("evict_last hubs (400k rows) / evict_first tail, no carve-out", time_ms([&] { k_gather<1><<<148 * 8, 256>>>(n, L, hot / 2, idx, D, Y); }, 3));
        for (int mb : {32, 64, 79}) {
            const size_t want = std::min<size_t>((size_t)maxp, (size_t)mb << 20);
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
            char nm[128];
            snprintf(nm, sizeof nm, "carve-out %d MB + evict_last hubs (%dk rows) / evict_first tail", mb, (int)(want / 80 / 1000));
            rep(nm, time_ms([&] { k_gather<1><<<148 * 8, 256>>>(n, L, (int)(want / 80), idx, D, Y); }, 3));
            // access policy window on the stream: hub prefix persisting, everything else streaming
            cudaStream_t st; CK(cudaStreamCreate(&st));
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = D; av.accessPolicyWindow.num_bytes = want; av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            snprintf(nm, sizeof nm, "carve-out %d MB + stream access-policy window (persisting | streaming)", mb);
            rep(nm, time_ms([&] { k_gather<0><<<148 * 8, 256, 0, st>>>(n, L, hot, idx, D, Y); }, 3));
            av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            snprintf(nm, sizeof nm, "carve-out %d MB + stream access-policy window (persisting | normal)", mb);
            rep(nm, time_ms([&] { k_gather<0><<<148 * 8, 256, 0, st>>>(n, L, hot, idx, D, Y); }, 3));
            av.accessPolicyWindow.num_bytes = 0;
            CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
            CK(cudaStreamDestroy(st));
            CK(cudaCtxResetPersistingL2Cache());
        }
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        rep("no hints (carve-out back to 0)", time_ms([&] { k_gather<0><<<148 * 8, 256>>>(n, L, hot, idx, D, Y); }, 3));
        // hub columns only / tail columns only (what a two-phase pass would see): rewrite idx
        for (i64 k = 0; k < nnz; k++) { i64 j = (i64)(hot * urand(seed)); hidx[k] = (int)j; }
        CK(cudaMemcpy(idx, hidx.data(), nnz * 4, cudaMemcpyHostToDevice));
        rep("all columns uniform in the 800k hub prefix (64 MB, L2-resident)", time_ms([&] { k_gather<0><<<148 * 8, 256>>>(n, L, hot, idx, D, Y); }, 3));
        CK(cudaFree(idx));
    }
    // ---- Q3
    {
        const i64 acc_rows = 800000;
        const int grid = 148 * 8, gpb = 256 / 5;
        const i64 per = (i64)43600000 / ((i64)grid * gpb);
        float ms = time_ms([&] { k_red<<<grid, 256>>>(acc_rows, per, Y); }, 3);
        const double pushes = (double)per * grid * gpb;
        printf("Q3 red.global.add.f64: %.1f M row pushes (10 reds each) into 64 MB: %7.3f ms  %6.2f rows/ns  %6.1f Gred/s\n", pushes / 1e6, ms, pushes / (ms * 1e6), pushes * 10 / (ms * 1e6));
    }
    return 0;
}
