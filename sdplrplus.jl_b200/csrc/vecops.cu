// vecops.cu -- streaming kernels over the m-vectors (lambda, y, residuals,
// A_RD, A_DD): augmented-Lagrangian value, quartic coefficients, residual
// recurrence, feasibility norm, dual update, Armijo evaluation.
//
// Reference: src/coreop.jl:11-31 (f!), src/linesearch.jl:36-56, 118-124,
// 158-172, src/sdplr.jl:224-234, 358-362, src/coreop.jl:412 (dual value).
// All reductions are deterministic two-stage sums finalised on the device.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = kRedThreads;

// f! tail: raw[0:m] -= b; obj = raw[m]; L = obj + sum (yt^2 - lambda^2)/(2 sigma)
__global__ void __launch_bounds__(TPB) k_f_finish(i64 m, double sigma, const double *__restrict__ b,
                                                  const double *__restrict__ lambda, const double *__restrict__ ub,
                                                  double *__restrict__ raw, double *partials, unsigned *ticket,
                                                  double *__restrict__ dscal) {
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double v = raw[i] - b[i];
        raw[i] = v;
        const double l = lambda[i];
        const double yt = fmin(ub[i], l - sigma * v);
        acc[0] += (yt * yt - l * l) / (2.0 * sigma);
    }
    const double obj = raw[m];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) {
        dscal[SC_OBJ] = obj;
        dscal[SC_LVAL] = obj + s[0];
    });
}

// the eight dot products behind the five quartic coefficients (src/linesearch.jl:36-56)
__global__ void __launch_bounds__(TPB) k_biquadratic(i64 m, double sigma, const double *__restrict__ lambda,
                                                     const double *__restrict__ raw, const double *__restrict__ q1v,
                                                     const double *__restrict__ q2v, double *partials, unsigned *ticket,
                                                     double *__restrict__ dscal) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double l = lambda[i], q0 = raw[i], q1 = q1v[i], q2 = q2v[i];
        acc[0] += l * q0;
        acc[1] += q0 * q0;
        acc[2] += l * q1;
        acc[3] += q0 * q1;
        acc[4] += (l - sigma * q0) * q2;
        acc[5] += q1 * q1;
        acc[6] += q1 * q2;
        acc[7] += q2 * q2;
    }
    const double p0 = raw[m], p1 = q1v[m], p2 = q2v[m];
    grid_sum_finalize<8>(acc, partials, ticket, [&](double (&s)[8]) {
        dscal[SC_BQ + 0] = p0 - s[0] + sigma * s[1] / 2.0;
        dscal[SC_BQ + 1] = p1 - s[2] + sigma * s[3];
        dscal[SC_BQ + 2] = p2 - s[4] + sigma * s[5] / 2.0;
        dscal[SC_BQ + 3] = sigma * s[6];
        dscal[SC_BQ + 4] = sigma * s[7] / 2.0;
    });
}

// raw += a*(a*A_DD + A_RD) over all m+1 slots; obj = raw[m]  (src/linesearch.jl:118-119)
__global__ void k_commit(i64 m, double a, const double *__restrict__ q1v, const double *__restrict__ q2v,
                         double *__restrict__ raw, double *__restrict__ dscal) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= m; i += (i64)gridDim.x * blockDim.x) {
        const double v = raw[i] + a * (a * q2v[i] + q1v[i]);
        raw[i] = v;
        if (i == m) dscal[SC_OBJ] = v;
    }
}

// the m-vector part of the fused step/gradient pass (gradient.cu, k_step_grad) for the constraints that are NOT in the
// per-row lists, slots [c0, m), plus the objective slot m:  residual recurrence into the alternate buffer, y, and this
// range's share of ||max(raw, lb)||^2 -> *pn2_out (the row pass adds its own share)
__global__ void __launch_bounds__(TPB) k_tail_rest(i64 c0, i64 m, double a, double sigma, const double *__restrict__ q1v,
                                                   const double *__restrict__ q2v, const double *__restrict__ raw_in,
                                                   double *__restrict__ raw_out, const double *__restrict__ lambda,
                                                   const double *__restrict__ ub, const double *__restrict__ lb,
                                                   double *__restrict__ y, double *__restrict__ pn2_out,
                                                   double *partials, unsigned *ticket, double *__restrict__ dscal) {
    double acc[1] = {0.0};
    for (i64 i = c0 + blockIdx.x * (i64)blockDim.x + threadIdx.x; i <= m; i += (i64)gridDim.x * blockDim.x) {
        const double v = raw_in[i] + a * (a * q2v[i] + q1v[i]);
        raw_out[i] = v;
        if (i == m) { y[i] = 1.0; dscal[SC_OBJ] = v; continue; }
        y[i] = -fmin(ub[i], lambda[i] - sigma * v);
        const double w = fmax(v, lb[i]);
        acc[0] += w * w;
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { pn2_out[0] = s[0]; });
}

// ||max(raw, lb)||_2^2 (src/coreop.jl:340-347, src/sdplr.jl:230-234)
__global__ void __launch_bounds__(TPB) k_pnorm2(i64 m, const double *__restrict__ raw, const double *__restrict__ lb,
                                                double *partials, unsigned *ticket, double *__restrict__ dscal) {
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double v = fmax(raw[i], lb[i]);
        acc[0] += v * v;
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { dscal[SC_PNORM2] = s[0]; });
}

// lambda_i <- min(ub_i, lambda_i - sigma*raw_i)  (src/sdplr.jl:358-362)
__global__ void k_dual_update(i64 m, double sigma, const double *__restrict__ ub, const double *__restrict__ raw,
                              double *__restrict__ lambda) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x)
        lambda[i] = fmin(ub[i], lambda[i] - sigma * raw[i]);
}

// sharp AL at K step sizes plus the slope at 0 (src/linesearch.jl:158-172)
constexpr int ARM_K = 15;
struct ArmijoArgs { double a[ARM_K]; int k; };
__global__ void __launch_bounds__(TPB) k_armijo(i64 m, double sigma, ArmijoArgs args, const double *__restrict__ lambda,
                                                const double *__restrict__ ub, const double *__restrict__ raw,
                                                const double *__restrict__ q1v, const double *__restrict__ q2v,
                                                const double *__restrict__ y, double *partials, unsigned *ticket,
                                                double *__restrict__ out) {
    double acc[ARM_K + 1];
#pragma unroll
    for (int k = 0; k <= ARM_K; k++) acc[k] = 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) {
        const double l = lambda[i], u = ub[i], q0 = raw[i], q1 = q1v[i], q2 = q2v[i];
#pragma unroll
        for (int k = 0; k < ARM_K; k++) {
            if (k < args.k) {
                const double a = args.a[k];
                const double g = q0 + a * q1 + a * a * q2;
                const double t = fmin(u, l - sigma * g);
                acc[k] += (t * t - l * l) / (2.0 * sigma);
            }
        }
        acc[ARM_K] += y[i] * q1;
    }
    const double p0 = raw[m], p1 = q1v[m], p2 = q2v[m];
    grid_sum_finalize<ARM_K + 1>(acc, partials, ticket, [&](double (&s)[ARM_K + 1]) {
        for (int k = 0; k < args.k; k++) {
            const double a = args.a[k];
            out[k] = p0 + a * p1 + a * a * p2 + s[k];
        }
        out[ARM_K] = p1 + s[ARM_K];
    });
}

// sum y_i b_i over i < m
__global__ void __launch_bounds__(TPB) k_yb(i64 m, const double *__restrict__ y, const double *__restrict__ b,
                                            double *partials, unsigned *ticket, double *__restrict__ out) {
    double acc[1] = {0.0};
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < m; i += (i64)gridDim.x * blockDim.x) acc[0] += y[i] * b[i];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

inline int red_grid(i64 m) { return grid_for(m, TPB, kRedBlocks); }

}  // namespace

int32_t vec_f_finish(sdplrp_handle *h) {
    k_f_finish<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->sigma, h->b, h->lambda, h->lambda_ub, h->pvio_raw, h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_biquadratic(sdplrp_handle *h) {
    k_biquadratic<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->sigma, h->lambda, h->pvio_raw, h->A_RD, h->A_DD, h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_commit(sdplrp_handle *h, double alpha) {
    k_commit<<<red_grid(h->m + 1), TPB, 0, h->stream>>>(h->m, alpha, h->A_RD, h->A_DD, h->pvio_raw, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_pnorm2(sdplrp_handle *h) {
    k_pnorm2<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->pvio_raw, h->pvio_lb, h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_dual_update(sdplrp_handle *h) {
    k_dual_update<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->sigma, h->lambda_ub, h->pvio_raw, h->lambda);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_armijo(sdplrp_handle *h, const double *alphas, int k, double *L, double *slope) {
    int done = 0;
    double sl = 0.0;
    while (done < k || k == 0) {
        ArmijoArgs args;
        args.k = std::min(ARM_K, k - done);
        for (int q = 0; q < ARM_K; q++) args.a[q] = q < args.k ? alphas[done + q] : 0.0;
        k_armijo<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->sigma, args, h->lambda, h->lambda_ub, h->pvio_raw, h->A_RD, h->A_DD, h->y,
                                                      h->partials, h->ticket, h->dscal + SC_LANCZOS);
        KLAUNCH(h);
        CUDA_TRY(h, cudaGetLastError());
        SDP_CHECK(fetch_scalars(h, SC_LANCZOS, ARM_K + 1));
        for (int q = 0; q < args.k; q++) L[done + q] = h->hscal[SC_LANCZOS + q];
        sl = h->hscal[SC_LANCZOS + ARM_K];
        done += args.k;
        if (k == 0) break;
    }
    if (slope) *slope = sl;
    return SDPLRP_OK;
}

int32_t vec_dual_dot(sdplrp_handle *h, double *out) {
    k_yb<<<red_grid(h->m), TPB, 0, h->stream>>>(h->m, h->y, h->b, h->partials, h->ticket, h->dscal + SC_LANCZOS + 9);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    SDP_CHECK(fetch_scalars(h, SC_LANCZOS + 9, 1));
    *out = -h->hscal[SC_LANCZOS + 9];
    return SDPLRP_OK;
}

// see k_tail_rest; launched BEFORE the fused row pass (S_dyn needs the y of these constraints)
int32_t vec_tail_rest(sdplrp_handle *h, double alpha, const double *raw_in, double *raw_out, double *pn2_out) {
    const i64 c0 = h->n_sd;
    k_tail_rest<<<red_grid(h->m + 1 - c0), TPB, 0, h->stream>>>(c0, h->m, alpha, h->sigma, h->A_RD, h->A_DD, raw_in, raw_out, h->lambda,
                                                              h->lambda_ub, h->pvio_lb, h->y, pn2_out, h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
