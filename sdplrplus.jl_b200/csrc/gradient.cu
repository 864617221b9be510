// gradient.cu -- y formation, assembly of S = C - sum (lambda_i - sigma v_i) A_i
// on the aggregated pattern, and the SpMM  G = 2 * S * R  (+ low-rank terms);
// plus the S*x product used by Lanczos.
//
// Reference: src/coreop.jl:205-317 (At_preprocess_sparse!, copy2y_lambda_sub_pvio!,
// At_preprocess!, At! left/right, g!) and src/structs.jl:90-145 (BDB' mul!).
//
// Design (not a translation): the reference rebuilds every value of S on each
// call (a scatter over all E_c entries followed by an nnzF-long gather).  Here
// the contribution of the objective matrix is folded once into
// triuS_static, S keeps y_{m+1}*static resident in HBM, and each iteration only
// rewrites the "dynamic" slots that some constraint touches (n diagonal slots
// for MaxCut) through a deterministic slot->contributors gather (no atomics).
// The SpMM walks the symmetric pattern as CSR: one sub-warp lane group per row
// keeps the r-vector accumulator in registers and gathers rows of R with
// 128-bit loads; rows longer than kLongRowThreshold get a whole CTA.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = 256;

// y_i = -min(ub_i, lambda_i - sigma*raw_i), y_{m+1} = 1   (src/coreop.jl:229-236)
__global__ void k_form_y(i64 m, double sigma, const double *__restrict__ lambda, const double *__restrict__ ub,
                         const double *__restrict__ raw, double *__restrict__ y) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= m; i += (i64)gridDim.x * blockDim.x) {
        if (i == m) { y[i] = 1.0; continue; }
        y[i] = -fmin(ub[i], lambda[i] - sigma * raw[i]);
    }
}

// static part: S[k] = y_obj * triuS_static[mapped[k]]
__global__ void k_S_static(i64 nnzF, double yobj, const int *__restrict__ mapped, const double *__restrict__ st,
                           double *__restrict__ S) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < nnzF; k += (i64)gridDim.x * blockDim.x) {
        int t = mapped[k];
        S[k] = t >= 0 ? yobj * st[t] : 0.0;
    }
}

// dynamic slots: constraints first (in matrix order), objective last -- the
// accumulation order of the reference's CSC mat-vec (src/coreop.jl:221)
__global__ void k_S_dynamic(i64 nd, double yobj, const int *__restrict__ dyn_slot, const int *__restrict__ dyn_ptr,
                            const int *__restrict__ dyn_gid, const double *__restrict__ dyn_val,
                            const int *__restrict__ pos_a, const int *__restrict__ pos_b, const double *__restrict__ st,
                            const double *__restrict__ y, double *__restrict__ S, double *__restrict__ triu_out) {
    for (i64 d = blockIdx.x * (i64)blockDim.x + threadIdx.x; d < nd; d += (i64)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int j = dyn_ptr[d]; j < dyn_ptr[d + 1]; j++) s += dyn_val[j] * y[dyn_gid[j]];
        const int t = dyn_slot[d];
        s += yobj * st[t];
        if (triu_out) {
            triu_out[t] = s;
        } else {
            S[pos_a[d]] = s;
            const int pb = pos_b[d];
            if (pb >= 0) S[pb] = s;
        }
    }
}

__global__ void k_scale_copy(i64 len, double a, const double *__restrict__ x, double *__restrict__ y) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < len; k += (i64)gridDim.x * blockDim.x) y[k] = a * x[k];
}

// ---- SpMM ------------------------------------------------------------------
template <int VEC>
struct Acc;
template <>
struct Acc<1> {
    double v;
    __device__ __forceinline__ void zero() { v = 0.0; }
    __device__ __forceinline__ void fma(double s, const double *p) { v += s * __ldg(p); }
    __device__ __forceinline__ void add(const Acc &o) { v += o.v; }
    __device__ __forceinline__ void store(double *p, double sc) const { *p = sc * v; }
    __device__ __forceinline__ void shfl_add(int o) { v += __shfl_xor_sync(0xffffffffu, v, o); }
};
template <>
struct Acc<2> {
    double2 v;
    __device__ __forceinline__ void zero() { v.x = v.y = 0.0; }
    __device__ __forceinline__ void fma(double s, const double *p) {
        double2 x = ldg2(p);
        v.x += s * x.x; v.y += s * x.y;
    }
    __device__ __forceinline__ void add(const Acc &o) { v.x += o.v.x; v.y += o.v.y; }
    __device__ __forceinline__ void store(double *p, double sc) const {
        *reinterpret_cast<double2 *>(p) = make_double2(sc * v.x, sc * v.y);
    }
    __device__ __forceinline__ void shfl_add(int o) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
};

// Y[i,:] = scale * sum_k S[k] * X[idx[k],:]  for rows i in [lo,hi) with at most
// kLongRowThreshold nonzeros. One group of G lanes per row; lane lg owns the
// vector units lg, lg+G, ... (MAXU of them).
template <int VEC, int MAXU>
__global__ void __launch_bounds__(TPB) k_spmm_rows(i64 lo, i64 hi, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                   const double *__restrict__ S, const double *__restrict__ X,
                                                   double *__restrict__ Y, int r, int G, double scale) {
    const int nv = r / VEC;
    const int lg = threadIdx.x & (G - 1);
    const i64 group = ((i64)blockIdx.x * TPB + threadIdx.x) / G;
    const i64 n_groups = (i64)gridDim.x * TPB / G;
    for (i64 i = lo + group; i < hi; i += n_groups) {
        const int beg = ptr[i], end = ptr[i + 1];
        if (end - beg > kLongRowThreshold) continue;
        Acc<VEC> acc[MAXU];
#pragma unroll
        for (int u = 0; u < MAXU; u++) acc[u].zero();
        int k = beg;
        for (; k + 4 <= end; k += 4) {  // 4 independent gathers in flight per lane
            const int c0 = __ldg(idx + k), c1 = __ldg(idx + k + 1), c2 = __ldg(idx + k + 2), c3 = __ldg(idx + k + 3);
            const double s0 = __ldg(S + k), s1 = __ldg(S + k + 1), s2 = __ldg(S + k + 2), s3 = __ldg(S + k + 3);
#pragma unroll
            for (int u = 0; u < MAXU; u++) {
                const int c = lg + u * G;
                if (c < nv) {
                    acc[u].fma(s0, X + (size_t)c0 * r + c * VEC);
                    acc[u].fma(s1, X + (size_t)c1 * r + c * VEC);
                    acc[u].fma(s2, X + (size_t)c2 * r + c * VEC);
                    acc[u].fma(s3, X + (size_t)c3 * r + c * VEC);
                }
            }
        }
        for (; k < end; k++) {
            const int c0 = __ldg(idx + k);
            const double s0 = __ldg(S + k);
#pragma unroll
            for (int u = 0; u < MAXU; u++) {
                const int c = lg + u * G;
                if (c < nv) acc[u].fma(s0, X + (size_t)c0 * r + c * VEC);
            }
        }
#pragma unroll
        for (int u = 0; u < MAXU; u++) {
            const int c = lg + u * G;
            if (c < nv) acc[u].store(Y + (size_t)i * r + c * VEC, scale);
        }
    }
}

// one CTA per long row: groups stride over the nonzeros, then a fixed-order
// cross-group sum in shared memory
template <int VEC, int MAXU>
__global__ void __launch_bounds__(TPB) k_spmm_long(const int *__restrict__ rows, i64 lo, i64 hi, const int *__restrict__ ptr,
                                                   const int *__restrict__ idx, const double *__restrict__ S,
                                                   const double *__restrict__ X, double *__restrict__ Y, int r, int G,
                                                   double scale) {
    extern __shared__ double sm[];  // (TPB/G) * r
    const i64 i = rows[blockIdx.x];
    if (i < lo || i >= hi) return;
    const int nv = r / VEC;
    const int lg = threadIdx.x & (G - 1), grp = threadIdx.x / G, ng = TPB / G;
    Acc<VEC> acc[MAXU];
#pragma unroll
    for (int u = 0; u < MAXU; u++) acc[u].zero();
    for (int k = ptr[i] + grp; k < ptr[i + 1]; k += ng) {
        const int c0 = __ldg(idx + k);
        const double s0 = __ldg(S + k);
#pragma unroll
        for (int u = 0; u < MAXU; u++) {
            const int c = lg + u * G;
            if (c < nv) acc[u].fma(s0, X + (size_t)c0 * r + c * VEC);
        }
    }
#pragma unroll
    for (int u = 0; u < MAXU; u++) {
        const int c = lg + u * G;
        if (c < nv) acc[u].store(sm + (size_t)grp * r + c * VEC, 1.0);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < r; e += TPB) {
        double t = 0.0;
        for (int g = 0; g < ng; g++) t += sm[(size_t)g * r + e];
        Y[(size_t)i * r + e] = scale * t;
    }
}

// Y[j,:] += scale*coeff * sum_k XB[:,k] D[k] B[j,k]     (src/structs.jl:135-145)
__global__ void k_lr_apply(i64 lo, i64 hi, int r, int s, i64 n, const double *__restrict__ XB, const double *__restrict__ Dg,
                           const double *__restrict__ B, const double *__restrict__ y, int gid, double scale,
                           double *__restrict__ Y) {
    const double coeff = scale * y[gid];
    for (i64 e = lo * r + blockIdx.x * (i64)blockDim.x + threadIdx.x; e < hi * r; e += (i64)gridDim.x * blockDim.x) {
        const i64 j = e / r;
        const int i = (int)(e - j * r);
        double t = 0.0;
        for (int k = 0; k < s; k++) t += XB[k * r + i] * Dg[k] * __ldg(&B[j + k * n]);
        Y[e] += coeff * t;
    }
}

// ---- SpMV (Lanczos): y = S*x, 8 lanes per row ---------------------------------
constexpr int SPMV_L = 8;
__global__ void __launch_bounds__(TPB) k_spmv(i64 n, const int *__restrict__ ptr, const int *__restrict__ idx,
                                              const double *__restrict__ S, const double *__restrict__ x,
                                              double *__restrict__ y) {
    const int lg = threadIdx.x & (SPMV_L - 1);
    const i64 warp_global = (i64)blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
    const i64 n_warps = (i64)gridDim.x * (TPB / 32);
    const int g_in_warp = (threadIdx.x & 31) / SPMV_L;
    constexpr int gpw = 32 / SPMV_L;
    for (i64 base = warp_global * gpw; base < n; base += n_warps * gpw) {
        const i64 i = base + g_in_warp;
        double t = 0.0;
        if (i < n)
            for (int k = ptr[i] + lg; k < ptr[i + 1]; k += SPMV_L) t += __ldg(S + k) * __ldg(x + __ldg(idx + k));
#pragma unroll
        for (int o = SPMV_L >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (i < n && lg == 0) y[i] = t;
    }
}

// low-rank part of S*x for one column: t_k = D_k * coeff * <B[:,k], x>, y += B t
__global__ void __launch_bounds__(TPB) k_lr_dot(i64 n, int s, const double *__restrict__ B, const double *__restrict__ x,
                                                double *__restrict__ partials, unsigned *__restrict__ ticket,
                                                double *__restrict__ out) {
    // one launch per k (s is tiny); out[0] = <B[:,k], x>
    double acc[1] = {0.0};
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) acc[0] += B[j] * x[j];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&sv)[1]) { out[0] = sv[0]; });
    (void)s;
}
__global__ void k_lr_axpy_vec(i64 n, const double *__restrict__ B, const double *__restrict__ dot, const double *__restrict__ Dg,
                              int k, const double *__restrict__ yv, int gid, double *__restrict__ out) {
    const double t = dot[0] * Dg[k] * yv[gid];
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) out[j] += B[j] * t;
}

int pick_group(int nv) {
    int G = 1;
    while (G < nv && G < 32) G <<= 1;
    return G;
}

}  // namespace

int32_t grad_form_y(sdplrp_handle *h) {
    k_form_y<<<grid_for(h->m + 1, TPB, kRedBlocks), TPB, 0, h->stream>>>(h->m, h->sigma, h->lambda, h->lambda_ub, h->pvio_raw, h->y);
    KLAUNCH(h);
    h->y_obj = 1.0;
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t grad_assemble_S(sdplrp_handle *h) {
    if (h->nA <= 0) return SDPLRP_OK;
    cudaStream_t st = h->stream;
    const double yobj = (h->obj_mat >= 0) ? h->y_obj : 0.0;
    if (!h->S_static_valid || h->S_static_scale != yobj) {
        if (h->nnzF > 0) {
            k_S_static<<<grid_for(h->nnzF, TPB, 8 * kNumSM), TPB, 0, st>>>(h->nnzF, yobj, h->mapped, h->triuS_static, h->S);
            KLAUNCH(h);
        }
        h->S_static_valid = true;
        h->S_static_scale = yobj;
    }
    if (h->n_dyn > 0) {
        k_S_dynamic<<<grid_for(h->n_dyn, TPB, 8 * kNumSM), TPB, 0, st>>>(h->n_dyn, yobj, h->dyn_slot, h->dyn_ptr, h->dyn_gid, h->dyn_val,
                                                                        h->dyn_pos_a, h->dyn_pos_b, h->triuS_static, h->y, h->S, nullptr);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t grad_triuS(sdplrp_handle *h, double *out) {
    if (h->nnzT <= 0) return SDPLRP_OK;
    cudaStream_t st = h->stream;
    const double yobj = (h->obj_mat >= 0) ? h->y_obj : 0.0;
    k_scale_copy<<<grid_for(h->nnzT, TPB, 8 * kNumSM), TPB, 0, st>>>(h->nnzT, yobj, h->triuS_static, out);
    KLAUNCH(h);
    if (h->n_dyn > 0) {
        k_S_dynamic<<<grid_for(h->n_dyn, TPB, 8 * kNumSM), TPB, 0, st>>>(h->n_dyn, yobj, h->dyn_slot, h->dyn_ptr, h->dyn_gid, h->dyn_val,
                                                                        h->dyn_pos_a, h->dyn_pos_b, h->triuS_static, h->y, h->S, out);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

template <int VEC, int MAXU>
static int32_t spmm_launch(sdplrp_handle *h, const double *X, double *Y, double scale, int G) {
    cudaStream_t st = h->stream;
    const i64 rows = h->row_hi - h->row_lo;
    const int gpb = TPB / G;
    k_spmm_rows<VEC, MAXU><<<grid_for(rows, gpb, 32 * kNumSM), TPB, 0, st>>>(h->row_lo, h->row_hi, h->full_ptr, h->full_idx, h->S, X, Y, h->r, G, scale);
    KLAUNCH(h);
    if (h->n_long_rows > 0) {
        size_t smem = (size_t)gpb * h->r * sizeof(double);
        k_spmm_long<VEC, MAXU><<<(int)h->n_long_rows, TPB, smem, st>>>(h->long_rows, h->row_lo, h->row_hi, h->full_ptr, h->full_idx, h->S, X, Y, h->r, G, scale);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

// Y = scale * (X*S + sum_g y_g X B D B')  over the owned rows
int32_t grad_spmm(sdplrp_handle *h, const double *X, double *Y, double scale, bool /*want_norm*/) {
    const int r = h->r;
    const bool vec2 = (r % 2 == 0);
    const int nv = vec2 ? r / 2 : r;
    const int G = pick_group(nv);
    const int units = (nv + G - 1) / G;
    if (units > 4) return fail(h, SDPLRP_ERR_ARG, "rank too large for the SpMM kernel (r <= 256 even / 128 odd)");
    if (h->nA > 0) {
        if (vec2) {
            if (units == 1) SDP_CHECK((spmm_launch<2, 1>(h, X, Y, scale, G)));
            else SDP_CHECK((spmm_launch<2, 4>(h, X, Y, scale, G)));
        } else {
            if (units == 1) SDP_CHECK((spmm_launch<1, 1>(h, X, Y, scale, G)));
            else SDP_CHECK((spmm_launch<1, 4>(h, X, Y, scale, G)));
        }
    } else {
        CUDA_TRY(h, cudaMemsetAsync(Y + h->row_lo * r, 0, (size_t)(h->row_hi - h->row_lo) * r * sizeof(double), h->stream));
    }
    if (!h->lr.empty()) {
        SDP_CHECK(lr_scratch(h));
        for (const LowRank &L : h->lr) {
            SDP_CHECK(lr_project(h, L, X, h->lr_tmp));
            k_lr_apply<<<grid_for((h->row_hi - h->row_lo) * r, TPB, kRedBlocks), TPB, 0, h->stream>>>(
                h->row_lo, h->row_hi, r, (int)L.s, h->n, h->lr_tmp, L.dD, L.dB, h->y, (int)L.gid, scale, Y);
            KLAUNCH(h);
        }
        CUDA_TRY(h, cudaGetLastError());
    }
    return SDPLRP_OK;
}

// y = S*x (+ low rank), x and y are n x ncols column-major device arrays
int32_t grad_spmv(sdplrp_handle *h, const double *x, double *y, i64 ncols) {
    cudaStream_t st = h->stream;
    const i64 n = h->n;
    for (i64 q = 0; q < ncols; q++) {
        const double *xq = x + q * n;
        double *yq = y + q * n;
        if (h->nA > 0) {
            k_spmv<<<grid_for(n, TPB / SPMV_L, 16 * kNumSM), TPB, 0, st>>>(n, h->full_ptr, h->full_idx, h->S, xq, yq);
            KLAUNCH(h);
        } else {
            CUDA_TRY(h, cudaMemsetAsync(yq, 0, (size_t)n * sizeof(double), st));
        }
        for (const LowRank &L : h->lr) {
            for (i64 k = 0; k < L.s; k++) {
                k_lr_dot<<<kRedBlocks, TPB, 0, st>>>(n, (int)L.s, L.dB + k * n, xq, h->partials, h->ticket, h->dscal + SC_LANCZOS + 8);
                KLAUNCH(h);
                k_lr_axpy_vec<<<grid_for(n, TPB, kRedBlocks), TPB, 0, st>>>(n, L.dB + k * n, h->dscal + SC_LANCZOS + 8, L.dD, (int)k, h->y, (int)L.gid, yq);
                KLAUNCH(h);
            }
        }
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
