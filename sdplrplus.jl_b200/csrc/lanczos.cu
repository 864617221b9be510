// lanczos.cu -- q-step Lanczos on S for the dual (suboptimality) bound, run as
// repeated device SpMV with device-side alpha/beta (no host synchronisation
// inside the recurrence), optional full re-orthogonalisation, and the host-side
// smallest eigenvalue of the tridiagonal.
//
// Reference: src/coreop.jl:461-514 (approx_mineigval_lanczos): random unit
// start, three-term recurrence, NO re-orthogonalisation, break when
// beta_i < sqrt(n)*eps, shift by +1, smallest eigenvalue of the tridiagonal.
#include <math.h>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int LZ_L = 8;  // lanes per row in the SpMV

// w = S*v over one row class with LANES lanes per row (rows are binned by length, RowClasses of the full pattern:
// <= 32 nonzeros -> 4 lanes, <= 2048 -> one warp, longer -> 8 warps); optionally the class's share of alpha_i = <v, w>.
// v (n doubles) is L2-resident; the pattern (12 B per nonzero) streams once per step.
template <int LANES, bool WITH_ALPHA>
__global__ void __launch_bounds__(TPB) k_lz_spmv(const int *__restrict__ rows, i64 n_rows, const int *__restrict__ ptr,
                                                 const int *__restrict__ idx, const double *__restrict__ S,
                                                 const double *__restrict__ v, double *__restrict__ w,
                                                 const double *__restrict__ stop, double *partials, unsigned *ticket,
                                                 double *alpha_part) {
    if (stop[0] != 0.0) return;
    constexpr int GPB = TPB / LANES;   // row groups per CTA
    const int lg = threadIdx.x % LANES, gib = threadIdx.x / LANES;
    __shared__ double red[TPB / 32];
    double acc[1] = {0.0};
    for (i64 base = (i64)blockIdx.x * GPB; base < n_rows; base += (i64)gridDim.x * GPB) {  // CTA-uniform trip count
        const i64 q = base + gib;
        const bool live = q < n_rows;
        const i64 i = live ? (rows ? rows[q] : q) : 0;
        double t = 0.0;
        if (live) {
            const int beg = ptr[i], end = ptr[i + 1];
            int k = beg + lg;
            for (; k + 3 * LANES < end; k += 4 * LANES) {  // four independent index -> value chains per lane
                const int c0 = __ldg(idx + k), c1 = __ldg(idx + k + LANES), c2 = __ldg(idx + k + 2 * LANES), c3 = __ldg(idx + k + 3 * LANES);
                const double s0 = __ldg(S + k), s1 = __ldg(S + k + LANES), s2 = __ldg(S + k + 2 * LANES), s3 = __ldg(S + k + 3 * LANES);
                t += s0 * __ldg(v + c0) + s1 * __ldg(v + c1) + s2 * __ldg(v + c2) + s3 * __ldg(v + c3);
            }
            for (; k < end; k += LANES) t += __ldg(S + k) * __ldg(v + __ldg(idx + k));
        }
        if (LANES <= 32) {
#pragma unroll
            for (int o = LANES >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        } else {  // one CTA-wide group per row: combine the warps in a fixed order
            t = warp_sum(t);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
            __syncthreads();
            t = 0.0;
            if (threadIdx.x == 0)
                for (int q2 = 0; q2 < TPB / 32; q2++) t += red[q2];
        }
        if (live && lg == 0) {
            w[i] = t;
            if (WITH_ALPHA) acc[0] += t * v[i];
        }
    }
    if (WITH_ALPHA) grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { alpha_part[0] = s[0]; });
}

__global__ void k_lz_alpha_sum(const double *__restrict__ parts, int nparts, const double *__restrict__ stop, double *alpha_out) {
    if (stop[0] != 0.0) return;
    double a = 0.0;
    for (int c = 0; c < nparts; c++) a += parts[c];
    alpha_out[0] = a;
}

__global__ void k_lz_zero(i64 n, double *__restrict__ w, const double *__restrict__ stop) {
    if (stop[0] != 0.0) return;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) w[i] = 0.0;
}

// out = <x, y> (guarded by the stop flag)
__global__ void __launch_bounds__(TPB) k_lz_dot(i64 n, const double *__restrict__ x, const double *__restrict__ y,
                                                const double *__restrict__ stop, double *partials, unsigned *ticket, double *out) {
    if (stop[0] != 0.0) return;
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc[0] += x[i] * y[i];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

// w += B_k * (dot * D_k * y_gid)
__global__ void k_lz_lr_axpy(i64 n, const double *__restrict__ B, const double *__restrict__ dot, const double *__restrict__ Dg,
                             int k, const double *__restrict__ yv, int gid, const double *__restrict__ stop,
                             double *__restrict__ w) {
    if (stop[0] != 0.0) return;
    const double t = dot[0] * Dg[k] * yv[gid];
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) w[j] += B[j] * t;
}

// w -= alpha_i v + beta_{i-1} vp ; beta_i = ||w|| ; stop when beta_i < sqrt(n) eps
__global__ void __launch_bounds__(TPB) k_lz_update(i64 n, int step, const double *__restrict__ v, const double *__restrict__ vp,
                                                   double *__restrict__ w, double *ab /* alpha[q], beta[q] */, i64 q,
                                                   double *stop, double *partials, unsigned *ticket) {
    if (stop[0] != 0.0) return;
    const double a = ab[step];
    const double bprev = step > 0 ? ab[q + step - 1] : 0.0;
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double t = w[i];
        if (step == 0) t -= a * v[i];
        else t -= a * v[i] + bprev * vp[i];
        w[i] = t;
        acc[0] += t * t;
    }
    const double thresh = sqrt((double)n) * 2.220446049250313e-16;
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) {
        const double b = sqrt(s[0]);
        ab[q + step] = b;
        if (fabs(b) < thresh) stop[0] = (double)(step + 1);
    });
}

// w /= beta_i (then the host rotates the three buffers)
__global__ void k_lz_normalise(i64 n, int step, const double *__restrict__ ab, i64 q, const double *__restrict__ stop,
                               double *__restrict__ w) {
    if (stop[0] != 0.0) return;
    const double inv = ab[q + step];
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) w[i] /= inv;
}

// start vector: v = v0 / ||v0||
__global__ void __launch_bounds__(TPB) k_lz_norm2(i64 n, const double *__restrict__ x, double *partials, unsigned *ticket, double *out) {
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc[0] += x[i] * x[i];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = sqrt(s[0]); });
}
__global__ void k_lz_div(i64 n, const double *__restrict__ nrm, double *__restrict__ x) {
    const double d = nrm[0];
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) x[i] /= d;
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// seeded standard normals (counter-based hash + Box-Muller)
__global__ void k_lz_randn(i64 n, unsigned long long seed, double *__restrict__ x) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const unsigned long long a = splitmix64(seed ^ (2ull * (unsigned long long)i));
        const unsigned long long b = splitmix64(seed ^ (2ull * (unsigned long long)i + 1ull));
        const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740992.0);
        const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
        x[i] = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
}

// full re-orthogonalisation (option; the reference has none):
// coef[j] = <V_j, w>, j < k  (one CTA per basis vector)
__global__ void __launch_bounds__(TPB) k_lz_reorth_dots(i64 n, const double *__restrict__ basis, const double *__restrict__ w,
                                                        const double *__restrict__ stop, double *__restrict__ coef) {
    if (stop[0] != 0.0) return;
    const double *vj = basis + (size_t)blockIdx.x * n;
    double acc[1] = {0.0};
    for (i64 i = threadIdx.x; i < n; i += blockDim.x) acc[0] += vj[i] * w[i];
    block_sum<1>(acc);
    if (threadIdx.x == 0) coef[blockIdx.x] = acc[0];
}
__global__ void k_lz_reorth_apply(i64 n, int k, const double *__restrict__ basis, const double *__restrict__ coef,
                                  const double *__restrict__ stop, double *__restrict__ w) {
    if (stop[0] != 0.0) return;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double t = w[i];
        for (int j = 0; j < k; j++) t -= coef[j] * basis[(size_t)j * n + i];
        w[i] = t;
    }
}
// recompute beta after re-orthogonalisation
__global__ void __launch_bounds__(TPB) k_lz_beta(i64 n, int step, const double *__restrict__ w, double *ab, i64 q, double *stop,
                                                 double *partials, unsigned *ticket) {
    if (stop[0] != 0.0) return;
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) acc[0] += w[i] * w[i];
    const double thresh = sqrt((double)n) * 2.220446049250313e-16;
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) {
        const double b = sqrt(s[0]);
        ab[q + step] = b;
        if (fabs(b) < thresh) stop[0] = (double)(step + 1);
    });
}

}  // namespace

// smallest eigenvalue of SymTridiagonal(d, e) by Sturm-sequence bisection
double tridiag_mineig_host(const double *d, const double *e, i64 k) {
    if (k == 1) return d[0];
    double lo = INFINITY, hi = -INFINITY;
    for (i64 i = 0; i < k; i++) {
        const double rad = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < k - 1 ? fabs(e[i]) : 0.0);
        lo = std::min(lo, d[i] - rad);
        hi = std::max(hi, d[i] + rad);
    }
    for (int it = 0; it < 200; it++) {
        const double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        int cnt = 0;
        double qv = d[0] - mid;
        if (qv < 0) cnt++;
        for (i64 i = 1; i < k && cnt == 0; i++) {
            if (qv == 0.0) qv = 1e-300;
            qv = d[i] - mid - e[i - 1] * e[i - 1] / qv;
            if (qv < 0) cnt++;
        }
        if (cnt >= 1) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

int32_t lz_run(sdplrp_handle *h, i64 q, const double *v0_host, uint64_t seed, int reorth, double *alpha, double *beta, i64 *iters) {
    const i64 n = h->n;
    cudaStream_t st = h->stream;
    if (q > n - 1) q = n - 1;
    if (q < 1) q = 1;
    if (!h->lz_v) {
        SDP_CHECK(dev_alloc(h, &h->lz_v, n)); SDP_CHECK(dev_alloc(h, &h->lz_w, n)); SDP_CHECK(dev_alloc(h, &h->lz_vp, n));
    }
    if (h->lz_ab_len < 2 * q) { SDP_CHECK(dev_alloc(h, &h->lz_ab, 2 * q)); h->lz_ab_len = 2 * q; }
    if (reorth && h->lz_basis_len < q * n) { SDP_CHECK(dev_alloc(h, &h->lz_basis, q * n)); h->lz_basis_len = q * n; }
    double *v = h->lz_v, *w = h->lz_w, *vp = h->lz_vp, *ab = h->lz_ab;
    double *stop = h->dscal + SC_LANCZOS, *tmp = h->dscal + SC_LANCZOS + 1;
    double *coef = nullptr;
    if (reorth) { SDP_CHECK(dev_alloc(h, &coef, q)); }
    CUDA_TRY(h, cudaMemsetAsync(ab, 0, (size_t)(2 * q) * sizeof(double), st));
    CUDA_TRY(h, cudaMemsetAsync(stop, 0, sizeof(double), st));
    CUDA_TRY(h, cudaMemsetAsync(vp, 0, (size_t)n * sizeof(double), st));
    const int gs = grid_for(n, TPB, kRedBlocks);
    if (v0_host) {
        SDP_CHECK(perm_upload(h, v, v0_host, 1, false));  // the start vector follows the internal vertex order
    } else {
        k_lz_randn<<<gs, TPB, 0, st>>>(n, seed, v); KLAUNCH(h);
    }
    k_lz_norm2<<<gs, TPB, 0, st>>>(n, v, h->partials, h->ticket, tmp); KLAUNCH(h);
    k_lz_div<<<gs, TPB, 0, st>>>(n, tmp, v); KLAUNCH(h);
    const bool has_lr = !h->lr.empty();
    for (i64 i = 0; i < q; i++) {
        if (reorth) CUDA_TRY(h, cudaMemcpyAsync(h->lz_basis + (size_t)i * n, v, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (h->nA > 0) {
            // three row classes of the full pattern (class lists are null when every row is short: identity)
            const RowClasses &cls = h->full_cls;
            double *parts = h->dscal + SC_LANCZOS + 2;
            CUDA_TRY(h, cudaMemsetAsync(parts, 0, 3 * sizeof(double), st));
            for (int c = 0; c < 3; c++) {
                const i64 nr = cls.cnt[c];
                if (nr <= 0) continue;
                const int *rows = cls.list[c];
#define LZ_SPMV(L, A) k_lz_spmv<L, A><<<grid_for(nr, TPB / L, 16 * kNumSM), TPB, 0, st>>>(rows, nr, h->full_ptr, h->full_idx, h->S, v, w, stop, h->partials, h->ticket, parts + c)
                if (c == 0) { if (has_lr) LZ_SPMV(4, false); else LZ_SPMV(4, true); }
                else if (c == 1) { if (has_lr) LZ_SPMV(32, false); else LZ_SPMV(32, true); }
                else { if (has_lr) LZ_SPMV(256, false); else LZ_SPMV(256, true); }
#undef LZ_SPMV
                KLAUNCH(h);
            }
            if (!has_lr) k_lz_alpha_sum<<<1, 1, 0, st>>>(parts, 3, stop, ab + i);
        } else {
            k_lz_zero<<<gs, TPB, 0, st>>>(n, w, stop);
        }
        KLAUNCH(h);
        if (has_lr || h->nA <= 0) {
            for (const LowRank &L : h->lr)
                for (i64 k = 0; k < L.s; k++) {
                    k_lz_dot<<<gs, TPB, 0, st>>>(n, L.dB + k * n, v, stop, h->partials, h->ticket, tmp); KLAUNCH(h);
                    k_lz_lr_axpy<<<gs, TPB, 0, st>>>(n, L.dB + k * n, tmp, L.dD, (int)k, h->y, (int)L.gid, stop, w); KLAUNCH(h);
                }
            k_lz_dot<<<gs, TPB, 0, st>>>(n, v, w, stop, h->partials, h->ticket, ab + i); KLAUNCH(h);
        }
        k_lz_update<<<gs, TPB, 0, st>>>(n, (int)i, v, vp, w, ab, q, stop, h->partials, h->ticket); KLAUNCH(h);
        if (reorth) {
            k_lz_reorth_dots<<<(int)(i + 1), TPB, 0, st>>>(n, h->lz_basis, w, stop, coef); KLAUNCH(h);
            k_lz_reorth_apply<<<gs, TPB, 0, st>>>(n, (int)(i + 1), h->lz_basis, coef, stop, w); KLAUNCH(h);
            k_lz_beta<<<gs, TPB, 0, st>>>(n, (int)i, w, ab, q, stop, h->partials, h->ticket); KLAUNCH(h);
        }
        k_lz_normalise<<<gs, TPB, 0, st>>>(n, (int)i, ab, q, stop, w); KLAUNCH(h);
        // rotate: vp <- v, v <- w/beta, w <- old vp (scratch)
        double *t = vp; vp = v; v = w; w = t;
    }
    CUDA_TRY(h, cudaGetLastError());
    std::vector<double> hab((size_t)(2 * q));
    double hstop = 0.0;
    CUDA_TRY(h, cudaMemcpyAsync(hab.data(), ab, (size_t)(2 * q) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(&hstop, stop, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    if (coef) cudaFree(coef);
    for (i64 i = 0; i < q; i++) { alpha[i] = hab[(size_t)i]; beta[i] = hab[(size_t)(q + i)]; }
    *iters = hstop != 0.0 ? (i64)hstop : q;
    return SDPLRP_OK;
}
