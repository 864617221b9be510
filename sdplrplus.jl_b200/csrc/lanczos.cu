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
                                                 double *alpha_part, i64 row_off = 0) {
    if (stop[0] != 0.0) return;
    constexpr int GPB = TPB / LANES;   // row groups per CTA
    const int lg = threadIdx.x % LANES, gib = threadIdx.x / LANES;
    __shared__ double red[TPB / 32];
    double acc[1] = {0.0};
    for (i64 base = (i64)blockIdx.x * GPB; base < n_rows; base += (i64)gridDim.x * GPB) {  // CTA-uniform trip count
        const i64 q = base + gib;
        const bool live = q < n_rows;
        const i64 i = live ? (rows ? rows[q] : q + row_off) : 0;   // row_off: identity list restricted to a rank's rows
        double t = 0.0;
        if (live) {
            const int beg = ptr[i], end = ptr[i + 1];
            if (LANES <= 8) {
                // short rows: blocks of 2*LANES nonzeros, fully predicated, so that a row of <= 2*LANES nonzeros costs ONE
                // round trip for its indices / values and ONE for its gathers (the scalar tail loop below paid a dependent
                // idx -> v pair per LANES nonzeros: ncu long_scoreboard, profiles/r2_lanczos.md)
                for (int k0 = beg; k0 < end; k0 += 2 * LANES) {
                    const int ka = k0 + lg, kb = ka + LANES;
                    const bool oa = ka < end, ob = kb < end;
                    const int ca = oa ? __ldg(idx + ka) : 0, cb = ob ? __ldg(idx + kb) : 0;
                    const double sa = oa ? __ldg(S + ka) : 0.0, sb = ob ? __ldg(S + kb) : 0.0;
                    const double va = __ldg(v + ca), vb = __ldg(v + cb);
                    t += sa * va;
                    t += sb * vb;
                }
            } else {
            int k = beg + lg;
            for (; k + 3 * LANES < end; k += 4 * LANES) {  // four independent index -> value chains per lane
                const int c0 = __ldg(idx + k), c1 = __ldg(idx + k + LANES), c2 = __ldg(idx + k + 2 * LANES), c3 = __ldg(idx + k + 3 * LANES);
                const double s0 = __ldg(S + k), s1 = __ldg(S + k + LANES), s2 = __ldg(S + k + 2 * LANES), s3 = __ldg(S + k + 3 * LANES);
                t += s0 * __ldg(v + c0) + s1 * __ldg(v + c1) + s2 * __ldg(v + c2) + s3 * __ldg(v + c3);
            }
            for (; k < end; k += LANES) t += __ldg(S + k) * __ldg(v + __ldg(idx + k));
            }
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

// ---- row-partitioned q-step Lanczos (option "lanczos_dist", world > 1) ------------------------------------------------
// positions [q_lo, q_hi) of the three ascending class lists whose rows lie in [own_lo, own_hi); out = {lo0, hi0, lo1, hi1, lo2, hi2}
__global__ void k_lz_class_ranges(const int *l0, i64 n0, const int *l1, i64 n1, const int *l2, i64 n2, i64 own_lo, i64 own_hi,
                                  long long *out) {
    const int c = threadIdx.x;
    if (c >= 3) return;
    const int *list = c == 0 ? l0 : (c == 1 ? l1 : l2);
    const i64 nl = c == 0 ? n0 : (c == 1 ? n1 : n2);
    i64 lo = 0, hi = nl;
    if (!list) {  // identity list (every row in this class)
        lo = own_lo < nl ? own_lo : nl;
        hi = own_hi < nl ? own_hi : nl;
        if (hi < lo) hi = lo;
    } else {
        i64 a = 0, b = nl;
        while (a < b) { const i64 mid = a + ((b - a) >> 1); if ((i64)list[mid] < own_lo) a = mid + 1; else b = mid; }
        lo = a;
        b = nl;
        while (a < b) { const i64 mid = a + ((b - a) >> 1); if ((i64)list[mid] < own_hi) a = mid + 1; else b = mid; }
        hi = a;
    }
    out[2 * c] = lo;
    out[2 * c + 1] = hi;
}

// owned rows: w -= alpha_i v + beta_{i-1} vp, this rank's share of ||w||^2 -> out[0]
__global__ void __launch_bounds__(TPB) k_lz_update_part(i64 n_own, int step, const double *__restrict__ v, const double *__restrict__ vp,
                                                        double *__restrict__ w, const double *__restrict__ ab, i64 q,
                                                        const double *__restrict__ stop, double *partials, unsigned *ticket, double *out) {
    if (stop[0] != 0.0) return;
    const double a = ab[step];
    const double bprev = step > 0 ? ab[q + step - 1] : 0.0;
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n_own; i += (i64)gridDim.x * blockDim.x) {
        double t = w[i];
        if (step == 0) t -= a * v[i];
        else t -= a * v[i] + bprev * vp[i];
        w[i] = t;
        acc[0] += t * t;
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

// beta_i = sqrt(sum over the ranks), stop when beta_i < sqrt(n) eps (every rank takes the same decision from the same sum)
__global__ void k_lz_beta_finish(i64 n, int step, const double *__restrict__ sum, double *ab, i64 q, double *stop) {
    if (stop[0] != 0.0) return;
    const double b = sqrt(sum[0]);
    ab[q + step] = b;
    if (fabs(b) < sqrt((double)n) * 2.220446049250313e-16) stop[0] = (double)(step + 1);
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


// w = S*v (+ the low-rank terms y_g B D B' v) on the internal vertex order and alpha_out[0] = <v, w>; every kernel is
// guarded by the device stop flag.  S is the full pattern as last assembled; the three row classes of the pattern
// (class lists are null when every row is short: identity) each get the lane-group width that fits their rows.
static int32_t lz_apply(sdplrp_handle *h, const double *v, double *w, const double *stop, double *alpha_out) {
    const i64 n = h->n;
    cudaStream_t st = h->stream;
    const int gs = grid_for(n, TPB, kRedBlocks);
    const bool has_lr = !h->lr.empty();
    double *tmp = h->dscal + SC_LANCZOS + 1;
    if (h->nA > 0) {
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
        if (!has_lr) k_lz_alpha_sum<<<1, 1, 0, st>>>(parts, 3, stop, alpha_out);
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
        k_lz_dot<<<gs, TPB, 0, st>>>(n, v, w, stop, h->partials, h->ticket, alpha_out); KLAUNCH(h);
    }
    return SDPLRP_OK;
}

// The same recurrence with the rows of S, w and the vector updates divided among the ranks (option "lanczos_dist", the
// default for world > 1; the replicated operator does not scale: 8 s of a 55 s two-GPU solve of C5).  Per step: the SpMV over
// the owned rows of each class, alpha and ||w||^2 as partial sums + one scalar all-reduce each, and one all-gather of the
// new Lanczos vector (n doubles), which the next SpMV gathers from.  Every rank ends with the same alpha / beta (the
// all-reduced sums are identical on all ranks), so the decisions taken from the dual bound stay SPMD-consistent.
// Measured on 2 GPUs (profiles/r2_multigpu.md): the same alpha / beta as the replicated recurrence bit for bit, 0.59 -> 0.24 s per dual check.
static int32_t lz_run_dist(sdplrp_handle *h, i64 q, const double *v0_host, uint64_t seed, double *alpha, double *beta, i64 *iters) {
    const i64 n = h->n;
    cudaStream_t st = h->stream;
    if (q > n - 1) q = n - 1;
    if (q < 1) q = 1;
    if (!h->lz_v) {
        SDP_CHECK(dev_alloc(h, &h->lz_v, n)); SDP_CHECK(dev_alloc(h, &h->lz_w, n)); SDP_CHECK(dev_alloc(h, &h->lz_vp, n));
    }
    if (h->lz_ab_len < 2 * q) { SDP_CHECK(dev_alloc(h, &h->lz_ab, 2 * q)); h->lz_ab_len = 2 * q; }
    double *v = h->lz_v, *w = h->lz_w, *vp = h->lz_vp, *ab = h->lz_ab;
    double *stop = h->dscal + SC_LANCZOS, *tmp = h->dscal + SC_LANCZOS + 1, *parts = h->dscal + SC_LANCZOS + 2;
    double *bsum = h->dscal + SC_LANCZOS + 5;
    const i64 lo = h->row_lo, hi = h->row_hi, n_own = hi - lo;
    const bool has_lr = !h->lr.empty();
    // list positions of the owned rows in every class
    long long *d_rng = nullptr;
    SDP_CHECK(dev_alloc(h, &d_rng, 6));
    long long rng[6] = {0, 0, 0, 0, 0, 0};
    if (h->nA > 0) {
        const RowClasses &cls = h->full_cls;
        k_lz_class_ranges<<<1, 32, 0, st>>>(cls.list[0], cls.cnt[0], cls.list[1], cls.cnt[1], cls.list[2], cls.cnt[2], lo, hi, d_rng);
        KLAUNCH(h);
        cudaError_t e = cudaMemcpyAsync(rng, d_rng, sizeof(rng), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFree(d_rng); h->err = std::string("lanczos: ") + cudaGetErrorString(e); return SDPLRP_ERR_CUDA; }
    }
    cudaFree(d_rng);
    CUDA_TRY(h, cudaMemsetAsync(ab, 0, (size_t)(2 * q) * sizeof(double), st));
    CUDA_TRY(h, cudaMemsetAsync(stop, 0, sizeof(double), st));
    CUDA_TRY(h, cudaMemsetAsync(vp, 0, (size_t)n * sizeof(double), st));
    const int gs = grid_for(n, TPB, kRedBlocks);
    const int gso = grid_for(std::max<i64>(n_own, 1), TPB, kRedBlocks);
    if (v0_host) {
        SDP_CHECK(perm_upload(h, v, v0_host, 1, false));
    } else {
        k_lz_randn<<<gs, TPB, 0, st>>>(n, seed, v); KLAUNCH(h);
    }
    k_lz_norm2<<<gs, TPB, 0, st>>>(n, v, h->partials, h->ticket, tmp); KLAUNCH(h);   // replicated: same vector on every rank
    k_lz_div<<<gs, TPB, 0, st>>>(n, tmp, v); KLAUNCH(h);
    for (i64 i = 0; i < q; i++) {
        // w[own] = (S v)[own], this rank's share of alpha_i
        CUDA_TRY(h, cudaMemsetAsync(parts, 0, 3 * sizeof(double), st));
        if (h->nA > 0) {
            const RowClasses &cls = h->full_cls;
            for (int c = 0; c < 3; c++) {
                const i64 nr = rng[2 * c + 1] - rng[2 * c];
                if (nr <= 0) continue;
                const int *rows = cls.list[c] ? cls.list[c] + rng[2 * c] : nullptr;
                const i64 off = cls.list[c] ? 0 : rng[2 * c];
#define LZ_SPMV(L, A) k_lz_spmv<L, A><<<grid_for(nr, TPB / L, 16 * kNumSM), TPB, 0, st>>>(rows, nr, h->full_ptr, h->full_idx, h->S, v, w, stop, h->partials, h->ticket, parts + c, off)
                if (c == 0) { if (has_lr) LZ_SPMV(4, false); else LZ_SPMV(4, true); }
                else if (c == 1) { if (has_lr) LZ_SPMV(32, false); else LZ_SPMV(32, true); }
                else { if (has_lr) LZ_SPMV(256, false); else LZ_SPMV(256, true); }
#undef LZ_SPMV
                KLAUNCH(h);
            }
        } else if (n_own > 0) {
            k_lz_zero<<<gso, TPB, 0, st>>>(n_own, w + lo, stop); KLAUNCH(h);
        }
        if (has_lr || h->nA <= 0) {
            for (const LowRank &L : h->lr)
                for (i64 k = 0; k < L.s; k++) {
                    // <B_k, v> over all rows (v is replicated, so no exchange), the update on the owned rows
                    k_lz_dot<<<gs, TPB, 0, st>>>(n, L.dB + k * n, v, stop, h->partials, h->ticket, tmp); KLAUNCH(h);
                    if (n_own > 0) { k_lz_lr_axpy<<<gso, TPB, 0, st>>>(n_own, L.dB + k * n + lo, tmp, L.dD, (int)k, h->y, (int)L.gid, stop, w + lo); KLAUNCH(h); }
                }
            k_lz_dot<<<gso, TPB, 0, st>>>(n_own, v + lo, w + lo, stop, h->partials, h->ticket, ab + i); KLAUNCH(h);
        } else {
            k_lz_alpha_sum<<<1, 1, 0, st>>>(parts, 3, stop, ab + i); KLAUNCH(h);
        }
        SDP_CHECK(comm_reduce_ptr(h, ab + i, 1));
        CUDA_TRY(h, cudaMemsetAsync(bsum, 0, sizeof(double), st));
        k_lz_update_part<<<gso, TPB, 0, st>>>(n_own, (int)i, v + lo, vp + lo, w + lo, ab, q, stop, h->partials, h->ticket, bsum); KLAUNCH(h);
        SDP_CHECK(comm_reduce_ptr(h, bsum, 1));
        k_lz_beta_finish<<<1, 1, 0, st>>>(n, (int)i, bsum, ab, q, stop); KLAUNCH(h);
        if (n_own > 0) { k_lz_normalise<<<gso, TPB, 0, st>>>(n_own, (int)i, ab, q, stop, w + lo); KLAUNCH(h); }
        SDP_CHECK(comm_gather_rowvec(h, w));
        double *t = vp; vp = v; v = w; w = t;
    }
    CUDA_TRY(h, cudaGetLastError());
    std::vector<double> hab((size_t)(2 * q));
    double hstop = 0.0;
    CUDA_TRY(h, cudaMemcpyAsync(hab.data(), ab, (size_t)(2 * q) * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(&hstop, stop, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    for (i64 i = 0; i < q; i++) { alpha[i] = hab[(size_t)i]; beta[i] = hab[(size_t)(q + i)]; }
    *iters = hstop != 0.0 ? (i64)hstop : q;
    return SDPLRP_OK;
}

// Measured and NOT built in (profiles/r2_call9_lanczos_l2_window.md): a persisting L2 set-aside (64 MB) with a stream access-policy
// window over the gathered vector cuts the DRAM reads of a step from 4.74 to 2.88 GB and the step gets SLOWER (1.17 -> 1.24 ms;
// 79 MB: 1.42 ms): the SpMV is bound by the L1 gather-wavefront rate (one wavefront per gathered 8-byte entry), not by DRAM.
int32_t lz_run(sdplrp_handle *h, i64 q, const double *v0_host, uint64_t seed, int reorth, double *alpha, double *beta, i64 *iters) {
    if (h->world > 1 && h->lanczos_dist && !reorth) return lz_run_dist(h, q, v0_host, seed, alpha, beta, iters);
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
    for (i64 i = 0; i < q; i++) {
        if (reorth) CUDA_TRY(h, cudaMemcpyAsync(h->lz_basis + (size_t)i * n, v, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        SDP_CHECK(lz_apply(h, v, w, stop, ab + i));
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

// =====================================================================================================================
// High-precision smallest eigenvalues of S: SDP_S_eigval (src/coreop.jl:351-374) calls GenericArpack's
// `symeigs(op, nevs; which=:SA, ncv, tol, maxiter)` on x -> S*x + x, i.e. the implicitly restarted Lanczos method.
// Here: THICK-RESTART Lanczos (Wu & Simon 2000), which spans the same Krylov subspaces as implicit restarting with
// exact shifts, run entirely on the device:
//   * the basis V (ncv+1 vectors of n doubles) stays in HBM; a Lanczos step is one SpMV (the row-class kernels of the
//     q-step path above) plus two classical Gram-Schmidt passes against the whole basis (blocks of 8 basis vectors per
//     streaming pass, deterministic grid reductions), so the recurrence never loses orthogonality;
//   * alpha_j / beta_j and the breakdown flag stay in device memory: the host synchronises ONCE per restart cycle,
//     solves the (ncv x ncv) arrowhead + tridiagonal projected problem (cyclic Jacobi) and uploads the Ritz
//     coefficients; the basis is rotated in place by a shared-memory tiled kernel (rows are CTA-private).
// The shift by the identity of the reference cancels exactly inside the orthogonalisation; it is applied to the
// projected matrix on the host so that the ARPACK stopping rule  |beta_m * y_m,i| <= tol * max(eps^(2/3), |theta_i|)
// sees the same shifted Ritz values.
// =====================================================================================================================
namespace {

constexpr int TRL_NB = 8;     // basis vectors per Gram-Schmidt dot pass
constexpr int TRL_ROWS = 32;  // rows per tile of the basis rotation

// coef[b] = <V_b, w>, b < cnt <= TRL_NB  (V_b = basis + b*n)
__global__ void __launch_bounds__(TPB) k_trl_dots(i64 n, const double *__restrict__ basis, int cnt, const double *__restrict__ w,
                                                  const double *__restrict__ stop, double *partials, unsigned *ticket,
                                                  double *__restrict__ coef) {
    if (stop[0] != 0.0) return;
    double acc[TRL_NB];
#pragma unroll
    for (int b = 0; b < TRL_NB; b++) acc[b] = 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double wi = w[i];
#pragma unroll
        for (int b = 0; b < TRL_NB; b++)
            if (b < cnt) acc[b] += basis[(size_t)b * n + i] * wi;
    }
    grid_sum_finalize<TRL_NB>(acc, partials, ticket, [&](double (&s)[TRL_NB]) {
        for (int b = 0; b < cnt; b++) coef[b] = s[b];
    });
}

// alpha_j = c1[j] + c2[j] (the two Gram-Schmidt passes)
__global__ void k_trl_alpha(int j, const double *__restrict__ c1, const double *__restrict__ c2, const double *__restrict__ stop,
                            double *__restrict__ ab) {
    if (stop[0] != 0.0) return;
    ab[j] = c1[j] + c2[j];
}

// V[:, 0:k] = V[:, 0:m] * Y (Y: m x k row-major), in place.  One CTA stages TRL_ROWS rows of all m basis vectors in
// shared memory, then writes the k combinations of those rows back; no other CTA touches them.
__global__ void __launch_bounds__(256) k_trl_rotate(i64 n, int m, int k, double *__restrict__ V, const double *__restrict__ Y) {
    extern __shared__ double tile[];  // m x TRL_ROWS
    const int row = threadIdx.x & (TRL_ROWS - 1), cg = threadIdx.x / TRL_ROWS, ncg = 256 / TRL_ROWS;
    for (i64 base = (i64)blockIdx.x * TRL_ROWS; base < n; base += (i64)gridDim.x * TRL_ROWS) {
        const i64 i = base + row;
        __syncthreads();
        for (int j = cg; j < m; j += ncg) tile[j * TRL_ROWS + row] = i < n ? V[(size_t)j * n + i] : 0.0;
        __syncthreads();
        for (int c = cg; c < k; c += ncg) {
            double t = 0.0;
            for (int j = 0; j < m; j++) t += tile[j * TRL_ROWS + row] * __ldg(Y + (size_t)j * k + c);
            if (i < n) V[(size_t)c * n + i] = t;
        }
    }
}

// eigen-decomposition of a small dense symmetric matrix (cyclic Jacobi): A (m x m, row-major, destroyed) ->
// ascending eigenvalues `ev`, eigenvectors as the COLUMNS of Q (row-major m x m)
void jacobi_eigh(std::vector<double> &A, int m, std::vector<double> &ev, std::vector<double> &Q) {
    Q.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++) Q[(size_t)i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < m; p++) {
            diag += A[(size_t)p * m + p] * A[(size_t)p * m + p];
            for (int q = p + 1; q < m; q++) off += A[(size_t)p * m + q] * A[(size_t)p * m + q];
        }
        if (off <= 1e-32 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < m - 1; p++)
            for (int q = p + 1; q < m; q++) {
                const double apq = A[(size_t)p * m + q];
                if (apq == 0.0) continue;
                const double app = A[(size_t)p * m + p], aqq = A[(size_t)q * m + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < m; k++) {  // columns p, q of A
                    const double akp = A[(size_t)k * m + p], akq = A[(size_t)k * m + q];
                    A[(size_t)k * m + p] = c * akp - s * akq;
                    A[(size_t)k * m + q] = s * akp + c * akq;
                }
                for (int k = 0; k < m; k++) {  // rows p, q of A
                    const double apk = A[(size_t)p * m + k], aqk = A[(size_t)q * m + k];
                    A[(size_t)p * m + k] = c * apk - s * aqk;
                    A[(size_t)q * m + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < m; k++) {
                    const double qkp = Q[(size_t)k * m + p], qkq = Q[(size_t)k * m + q];
                    Q[(size_t)k * m + p] = c * qkp - s * qkq;
                    Q[(size_t)k * m + q] = s * qkp + c * qkq;
                }
            }
    }
    std::vector<int> order((size_t)m);
    for (int i = 0; i < m; i++) order[(size_t)i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * m + a] < A[(size_t)b * m + b]; });
    ev.resize((size_t)m);
    std::vector<double> Qs((size_t)m * m);
    for (int c = 0; c < m; c++) {
        const int src = order[(size_t)c];
        ev[(size_t)c] = A[(size_t)src * m + src];
        for (int k = 0; k < m; k++) Qs[(size_t)k * m + c] = Q[(size_t)k * m + src];
    }
    Q.swap(Qs);
}

}  // namespace

void dense_symeig_host(const double *A, i64 m, double *ev, double *Q) {
    std::vector<double> a(A, A + (size_t)(m * m)), e, q;
    jacobi_eigh(a, (int)m, e, q);
    for (i64 i = 0; i < m; i++) ev[i] = e[(size_t)i];
    if (Q) for (i64 i = 0; i < m * m; i++) Q[i] = q[(size_t)i];
}

int32_t lz_eigs(sdplrp_handle *h, i64 nev, i64 ncv, double tol, i64 maxiter, const double *v0_host, uint64_t seed, double *eigs,
                double *bounds, i64 *matvecs, i64 *restarts) {
    const i64 n = h->n;
    cudaStream_t st = h->stream;
    if (nev < 1 || nev > n) return fail(h, SDPLRP_ERR_ARG, "S_eigval: nev must be in [1, n]");
    i64 m = std::min<i64>(std::max<i64>(ncv, nev + 1), n);  // ARPACK: nev < ncv <= n
    if (m > 512) return fail(h, SDPLRP_ERR_ARG, "S_eigval: ncv > 512 is not supported");
    if (tol <= 0.0) tol = 2.220446049250313e-16;             // ARPACK: tol = 0 means machine precision
    const double eps23 = pow(2.220446049250313e-16, 2.0 / 3.0);
    if (maxiter < 1) maxiter = 1;
    const int gs = grid_for(n, TPB, kRedBlocks);

    const i64 need = (m + 1) * n;
    if (h->lz_basis_len < need) { SDP_CHECK(dev_alloc(h, &h->lz_basis, need)); h->lz_basis_len = need; }
    if (h->lz_ab_len < 2 * m) { SDP_CHECK(dev_alloc(h, &h->lz_ab, 2 * m)); h->lz_ab_len = 2 * m; }
    double *V = h->lz_basis, *ab = h->lz_ab;
    double *small = nullptr;   // c1 (m+1) | c2 (m+1) | Y (m*m)
    SDP_CHECK(dev_alloc(h, &small, 2 * (m + 1) + m * m));
    struct Guard { double *p; ~Guard() { if (p) cudaFree(p); } } guard{small};
    double *c1 = small, *c2 = small + (m + 1), *Yd = small + 2 * (m + 1);
    double *stop = h->dscal + SC_LANCZOS, *tmp = h->dscal + SC_LANCZOS + 1;
    const size_t rot_smem = (size_t)m * TRL_ROWS * sizeof(double);
    if (rot_smem > 48 * 1024) CUDA_TRY(h, cudaFuncSetAttribute(k_trl_rotate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rot_smem));

    CUDA_TRY(h, cudaMemsetAsync(stop, 0, sizeof(double), st));
    if (v0_host) {
        SDP_CHECK(perm_upload(h, V, v0_host, 1, false));
    } else {
        k_lz_randn<<<gs, TPB, 0, st>>>(n, seed, V); KLAUNCH(h);
    }
    k_lz_norm2<<<gs, TPB, 0, st>>>(n, V, h->partials, h->ticket, tmp); KLAUNCH(h);
    k_lz_div<<<gs, TPB, 0, st>>>(n, tmp, V); KLAUNCH(h);

    std::vector<double> theta, svec;          // kept Ritz values / coupling entries of the arrowhead (size k)
    std::vector<double> hab((size_t)(2 * m)), T, ev, Q, Yk;
    i64 k = 0, nmv = 0, nrestart = 0;
    double best_bound = INFINITY, last_theta = NAN;
    int stalled = 0;
    int32_t rc = SDPLRP_OK;
    for (;;) {
        CUDA_TRY(h, cudaMemsetAsync(ab, 0, (size_t)(2 * m) * sizeof(double), st));
        for (i64 j = k; j < m; j++) {
            const double *vj = V + (size_t)j * n;
            double *w = V + (size_t)(j + 1) * n;
            rc = lz_apply(h, vj, w, stop, tmp);  // w = S*v_j   (tmp: <v_j, w>, superseded by the Gram-Schmidt coefficients)
            if (rc != SDPLRP_OK) break;
            nmv++;
            for (int pass = 0; pass < 2; pass++) {  // classical Gram-Schmidt, twice
                double *c = pass == 0 ? c1 : c2;
                for (i64 b0 = 0; b0 <= j; b0 += TRL_NB) {
                    const int cnt = (int)std::min<i64>(TRL_NB, j + 1 - b0);
                    k_trl_dots<<<gs, TPB, 0, st>>>(n, V + (size_t)b0 * n, cnt, w, stop, h->partials, h->ticket, c + b0); KLAUNCH(h);
                }
                k_lz_reorth_apply<<<gs, TPB, 0, st>>>(n, (int)(j + 1), V, c, stop, w); KLAUNCH(h);
            }
            k_trl_alpha<<<1, 1, 0, st>>>((int)j, c1, c2, stop, ab); KLAUNCH(h);
            k_lz_beta<<<gs, TPB, 0, st>>>(n, (int)j, w, ab, m, stop, h->partials, h->ticket); KLAUNCH(h);
            k_lz_normalise<<<gs, TPB, 0, st>>>(n, (int)j, ab, m, stop, w); KLAUNCH(h);
        }
        if (rc != SDPLRP_OK) break;
        double hstop = 0.0;
        if (cudaMemcpyAsync(hab.data(), ab, (size_t)(2 * m) * sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(&hstop, stop, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {
            h->err = std::string("S_eigval: ") + cudaGetErrorString(cudaGetLastError());
            rc = SDPLRP_ERR_CUDA;
            break;
        }
        // an invariant subspace was found at step `hstop` (beta < sqrt(n)*eps): the projected problem is exact
        const i64 me = hstop != 0.0 ? (i64)hstop : m;
        if (hstop != 0.0) nmv -= (m - me);  // the guarded kernels of the later steps did nothing
        T.assign((size_t)(me * me), 0.0);
        for (i64 i = 0; i < k; i++) {
            T[(size_t)(i * me + i)] = theta[(size_t)i];
            T[(size_t)(i * me + k)] = T[(size_t)(k * me + i)] = svec[(size_t)i];
        }
        for (i64 j = k; j < me; j++) {
            T[(size_t)(j * me + j)] = hab[(size_t)j];
            if (j + 1 < me) T[(size_t)(j * me + j + 1)] = T[(size_t)((j + 1) * me + j)] = hab[(size_t)(m + j)];
        }
        const double beta_last = hab[(size_t)(m + me - 1)];
        jacobi_eigh(T, (int)me, ev, Q);
        const i64 nwant = std::min<i64>(nev, me);
        // ARPACK's rule on the shifted Ritz value (the reference's operator is S + I), with the attainable floor of the
        // projected eigenproblem (Jacobi: ~ m eps ||T||) so that tol = 0 ("machine precision", the DIMACS call) terminates
        const double tnorm = std::max(fabs(ev[0] + 1.0), fabs(ev[(size_t)(me - 1)] + 1.0));
        const double floor_ = 4.0 * (double)me * 2.220446049250313e-16 * tnorm;
        bool conv = true;
        double worst = 0.0;
        for (i64 i = 0; i < nwant; i++) {
            const double bound = fabs(beta_last * Q[(size_t)((me - 1) * me + i)]);
            if (bounds) bounds[i] = bound;
            eigs[i] = ev[(size_t)i];
            const double want = tol * std::max(eps23, fabs(ev[(size_t)i] + 1.0));
            if (!(bound <= std::max(want, floor_))) conv = false;
            worst = std::max(worst, bound);
        }
        // stagnation guard: the bounds have stopped improving and the wanted Ritz values have stopped moving
        const bool moved = fabs(ev[(size_t)(nwant - 1)] - last_theta) > 1e-13 * std::max(1.0, tnorm);
        if (worst < 0.5 * best_bound) { best_bound = worst; stalled = 0; }
        else if (!moved) stalled++;
        last_theta = ev[(size_t)(nwant - 1)];
        nrestart++;
        if (conv || hstop != 0.0 || nrestart >= maxiter || me < 2 || stalled >= 8) {
            for (i64 i = nwant; i < nev; i++) { eigs[i] = NAN; if (bounds) bounds[i] = NAN; }
            break;
        }
        // thick restart: keep the lowest kk Ritz vectors and the residual vector
        i64 kk = nev + ((me - nev) * 2) / 5;
        kk = std::max<i64>(nev, std::min<i64>(kk, me - 2));
        kk = std::max<i64>(1, std::min<i64>(kk, me - 1));
        Yk.resize((size_t)(me * kk));
        for (i64 j = 0; j < me; j++)
            for (i64 c = 0; c < kk; c++) Yk[(size_t)(j * kk + c)] = Q[(size_t)(j * me + c)];
        if (cudaMemcpyAsync(Yd, Yk.data(), Yk.size() * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) {
            h->err = "S_eigval: upload of the Ritz coefficients failed";
            rc = SDPLRP_ERR_CUDA;
            break;
        }
        k_trl_rotate<<<grid_for(n, TRL_ROWS, 8 * kNumSM), 256, rot_smem, st>>>(n, (int)me, (int)kk, V, Yd); KLAUNCH(h);
        if (cudaMemcpyAsync(V + (size_t)kk * n, V + (size_t)me * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) {   // Yk is host memory that the next cycle rewrites
            h->err = "S_eigval: basis restart failed";
            rc = SDPLRP_ERR_CUDA;
            break;
        }
        theta.assign(ev.begin(), ev.begin() + kk);
        svec.resize((size_t)kk);
        for (i64 i = 0; i < kk; i++) svec[(size_t)i] = beta_last * Q[(size_t)((me - 1) * me + i)];
        k = kk;
    }
    if (rc != SDPLRP_OK) return rc;
    CUDA_TRY(h, cudaGetLastError());
    if (matvecs) *matvecs = nmv;
    if (restarts) *restarts = nrestart;
    return SDPLRP_OK;
}
