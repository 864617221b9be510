// vecops.cu -- streaming kernels over the m-vectors (lambda, y, residuals,
// A_RD, A_DD): augmented-Lagrangian value, quartic coefficients, residual
// recurrence, feasibility norm, dual update, Armijo evaluation.
//
// Reference: src/coreop.jl:11-31 (f!), src/linesearch.jl:36-56, 118-124,
// 158-172, src/sdplr.jl:224-234, 358-362, src/coreop.jl:412 (dual value).
// All reductions are deterministic two-stage sums finalised on the device.
//
// Constraint slots and ranks.  The m-vectors are kept in the internal constraint order
// (preprocess.cu): the single-diagonal-entry constraints first, in internal row order, then every
// other constraint, then the objective slot m.  With several GPUs a rank OWNS the slots of the
// per-row constraints of its rows, [c_lo, c_hi), and every rank keeps an identical copy of the
// shared slots [n_sd, m] (their values are all-reduced where they are produced).  Slots of other
// ranks' rows are never read.  Element-wise kernels run over owned + shared slots; sums run over
// owned slots plus -- on rank 0 only -- the shared slots, are all-reduced as scalars and finished by
// a one-thread kernel, so no m-vector crosses NVLink inside the iteration.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = kRedThreads;

struct CRange {
    i64 a0, a1;  // owned per-row-constraint slots
    i64 b0, b1;  // shared slots this launch covers
    __host__ __device__ i64 len() const { return (a1 - a0) + (b1 - b0); }
    __device__ __forceinline__ i64 at(i64 k) const { return k < a1 - a0 ? a0 + k : b0 + (k - (a1 - a0)); }
};
// element-wise work: owned + every shared slot (each rank keeps the shared slots current)
CRange range_all(const sdplrp_handle *h) { return CRange{h->c_lo, h->c_hi, h->n_sd, h->m}; }
// sums: the shared slots are counted once, by rank 0
CRange range_sum(const sdplrp_handle *h) { return CRange{h->c_lo, h->c_hi, h->n_sd, h->rank == 0 ? h->m : h->n_sd}; }

#define FOR_SLOTS(i, R)                                                                                  \
    for (i64 k__ = blockIdx.x * (i64)blockDim.x + threadIdx.x, i = 0;                                     \
         k__ < (R).len() && ((i = (R).at(k__)), true); k__ += (i64)gridDim.x * blockDim.x)

// f! tail, part 1: raw[i] -= b[i] over owned + shared slots (element-wise)
__global__ void k_sub_b(CRange R, const double *__restrict__ b, double *__restrict__ raw) {
    FOR_SLOTS(i, R) raw[i] -= b[i];
}
// f! tail, part 2: this rank's share of sum (yt^2 - lambda^2)/(2 sigma) -> *out
__global__ void __launch_bounds__(TPB) k_f_sum(CRange R, double sigma, const double *__restrict__ lambda,
                                               const double *__restrict__ ub, const double *__restrict__ raw, double *partials,
                                               unsigned *ticket, double *__restrict__ out) {
    double acc[1] = {0.0};
    FOR_SLOTS(i, R) {
        const double l = lambda[i];
        const double yt = fmin(ub[i], l - sigma * raw[i]);
        acc[0] += (yt * yt - l * l) / (2.0 * sigma);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}
__global__ void k_f_final(i64 m, const double *__restrict__ raw, double *__restrict__ dscal) {
    const double obj = raw[m];
    dscal[SC_OBJ] = obj;
    dscal[SC_LVAL] = obj + dscal[SC_LVAL];
}

// the eight dot products behind the five quartic coefficients (src/linesearch.jl:36-56) -> dscal[SC_BQ .. SC_BQ+8)
__global__ void __launch_bounds__(TPB) k_biquadratic(CRange R, double sigma, const double *__restrict__ lambda,
                                                     const double *__restrict__ raw, const double *__restrict__ q1v,
                                                     const double *__restrict__ q2v, double *partials, unsigned *ticket,
                                                     double *__restrict__ dscal) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.0;
    FOR_SLOTS(i, R) {
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
    grid_sum_finalize<8>(acc, partials, ticket, [&](double (&s)[8]) {
#pragma unroll
        for (int k = 0; k < 8; k++) dscal[SC_BQ + k] = s[k];
    });
}
__global__ void k_biquadratic_final(i64 m, double sigma, const double *__restrict__ raw, const double *__restrict__ q1v,
                                    const double *__restrict__ q2v, double *__restrict__ dscal) {
    const double p0 = raw[m], p1 = q1v[m], p2 = q2v[m];
    double s[8];
    for (int k = 0; k < 8; k++) s[k] = dscal[SC_BQ + k];
    dscal[SC_BQ + 0] = p0 - s[0] + sigma * s[1] / 2.0;
    dscal[SC_BQ + 1] = p1 - s[2] + sigma * s[3];
    dscal[SC_BQ + 2] = p2 - s[4] + sigma * s[5] / 2.0;
    dscal[SC_BQ + 3] = sigma * s[6];
    dscal[SC_BQ + 4] = sigma * s[7] / 2.0;
}

// raw += a*(a*A_DD + A_RD) over owned + shared slots and the objective slot; obj = raw[m]  (src/linesearch.jl:118-119)
__global__ void k_commit(CRange R, i64 m, double a, const double *__restrict__ q1v, const double *__restrict__ q2v,
                         double *__restrict__ raw, double *__restrict__ dscal) {
    FOR_SLOTS(i, R) raw[i] += a * (a * q2v[i] + q1v[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double v = raw[m] + a * (a * q2v[m] + q1v[m]);
        raw[m] = v;
        dscal[SC_OBJ] = v;
    }
}

// the m-vector part of the fused step/gradient pass (gradient.cu, k_step_grad) for the constraints that are NOT in the
// per-row lists, shared slots [c0, m), plus the objective slot m:  residual recurrence into the alternate buffer, y, and
// this range's share of ||max(raw, lb)||^2 -> *pn2_out (the row pass adds its own share)
__global__ void __launch_bounds__(TPB) k_tail_rest(i64 c0, i64 m, double a, double sigma, const double *__restrict__ q1v,
                                                   const double *__restrict__ q2v, const double *__restrict__ raw_in,
                                                   double *__restrict__ raw_out, const double *__restrict__ lambda,
                                                   const double *__restrict__ ub, const double *__restrict__ lb,
                                                   double *__restrict__ y, double *__restrict__ pn2_out, double pn2_weight,
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
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { pn2_out[0] = pn2_weight * s[0]; });
}

// ||max(raw, lb)||_2^2 (src/coreop.jl:340-347, src/sdplr.jl:230-234), this rank's share
__global__ void __launch_bounds__(TPB) k_pnorm2(CRange R, const double *__restrict__ raw, const double *__restrict__ lb,
                                                double *partials, unsigned *ticket, double *__restrict__ dscal) {
    double acc[1] = {0.0};
    FOR_SLOTS(i, R) {
        const double v = fmax(raw[i], lb[i]);
        acc[0] += v * v;
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { dscal[SC_PNORM2] = s[0]; });
}

// lambda_i <- min(ub_i, lambda_i - sigma*raw_i)  (src/sdplr.jl:358-362)
__global__ void k_dual_update(CRange R, double sigma, const double *__restrict__ ub, const double *__restrict__ raw,
                              double *__restrict__ lambda) {
    FOR_SLOTS(i, R) lambda[i] = fmin(ub[i], lambda[i] - sigma * raw[i]);
}

// sharp AL at K step sizes plus the slope at 0 (src/linesearch.jl:158-172): this rank's partial sums -> out[0..ARM_K]
constexpr int ARM_K = 15;
struct ArmijoArgs { double a[ARM_K]; int k; };
__global__ void __launch_bounds__(TPB) k_armijo(CRange R, double sigma, ArmijoArgs args, const double *__restrict__ lambda,
                                                const double *__restrict__ ub, const double *__restrict__ raw,
                                                const double *__restrict__ q1v, const double *__restrict__ q2v,
                                                const double *__restrict__ y, double *partials, unsigned *ticket,
                                                double *__restrict__ out) {
    double acc[ARM_K + 1];
#pragma unroll
    for (int k = 0; k <= ARM_K; k++) acc[k] = 0.0;
    FOR_SLOTS(i, R) {
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
    grid_sum_finalize<ARM_K + 1>(acc, partials, ticket, [&](double (&s)[ARM_K + 1]) {
#pragma unroll
        for (int k = 0; k <= ARM_K; k++) out[k] = s[k];
    });
}
__global__ void k_armijo_final(i64 m, ArmijoArgs args, const double *__restrict__ raw, const double *__restrict__ q1v,
                               const double *__restrict__ q2v, double *__restrict__ out) {
    const double p0 = raw[m], p1 = q1v[m], p2 = q2v[m];
    for (int k = 0; k < args.k; k++) {
        const double a = args.a[k];
        out[k] = p0 + a * p1 + a * a * p2 + out[k];
    }
    out[ARM_K] = p1 + out[ARM_K];
}

// sum y_i b_i over i < m (this rank's share)
__global__ void __launch_bounds__(TPB) k_yb(CRange R, const double *__restrict__ y, const double *__restrict__ b,
                                            double *partials, unsigned *ticket, double *__restrict__ out) {
    double acc[1] = {0.0};
    FOR_SLOTS(i, R) acc[0] += y[i] * b[i];
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

// DIMACS sums (src/coreop.jl:431, 441-447): this rank's share of ||raw[1:m]||^2 and lambda'b -> out[0], out[1]
__global__ void __launch_bounds__(TPB) k_dimacs_sums(CRange R, const double *__restrict__ raw, const double *__restrict__ lambda,
                                                     const double *__restrict__ b, double *partials, unsigned *ticket,
                                                     double *__restrict__ out) {
    double acc[2] = {0.0, 0.0};
    FOR_SLOTS(i, R) {
        acc[0] += raw[i] * raw[i];
        acc[1] += lambda[i] * b[i];
    }
    grid_sum_finalize<2>(acc, partials, ticket, [&](double (&s)[2]) { out[0] = s[0]; out[1] = s[1]; });
}

// copy2y_lambda! (src/coreop.jl:238-246): y_i = -lambda_i over owned + shared slots, y_{m+1} = 1
__global__ void k_copy2y_lambda(CRange R, i64 m, const double *__restrict__ lambda, double *__restrict__ y) {
    FOR_SLOTS(i, R) y[i] = -lambda[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) y[m] = 1.0;
}

inline int red_grid(i64 m) { return grid_for(m, TPB, kRedBlocks); }

}  // namespace

int32_t vec_f_finish(sdplrp_handle *h) {
    const CRange all = range_all(h), sum = range_sum(h);
    k_sub_b<<<red_grid(all.len()), TPB, 0, h->stream>>>(all, h->b, h->pvio_raw);
    KLAUNCH(h);
    k_f_sum<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->sigma, h->lambda, h->lambda_ub, h->pvio_raw, h->partials, h->ticket,
                                                       h->dscal + SC_LVAL);
    KLAUNCH(h);
    SDP_CHECK(comm_reduce_scalars(h, SC_LVAL, 1));
    k_f_final<<<1, 1, 0, h->stream>>>(h->m, h->pvio_raw, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_biquadratic(sdplrp_handle *h) {
    const CRange sum = range_sum(h);
    k_biquadratic<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->sigma, h->lambda, h->pvio_raw, h->A_RD, h->A_DD, h->partials,
                                                             h->ticket, h->dscal);
    KLAUNCH(h);
    SDP_CHECK(comm_reduce_scalars(h, SC_BQ, 8));
    k_biquadratic_final<<<1, 1, 0, h->stream>>>(h->m, h->sigma, h->pvio_raw, h->A_RD, h->A_DD, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_commit(sdplrp_handle *h, double alpha) {
    const CRange all = range_all(h);
    k_commit<<<red_grid(all.len()), TPB, 0, h->stream>>>(all, h->m, alpha, h->A_RD, h->A_DD, h->pvio_raw, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_pnorm2(sdplrp_handle *h) {
    const CRange sum = range_sum(h);
    k_pnorm2<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->pvio_raw, h->pvio_lb, h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return comm_reduce_scalars(h, SC_PNORM2, 1);
}

int32_t vec_dual_update(sdplrp_handle *h) {
    const CRange all = range_all(h);
    k_dual_update<<<red_grid(all.len()), TPB, 0, h->stream>>>(all, h->sigma, h->lambda_ub, h->pvio_raw, h->lambda);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t vec_armijo(sdplrp_handle *h, const double *alphas, int k, double *L, double *slope) {
    int done = 0;
    double sl = 0.0;
    const CRange sum = range_sum(h);
    while (done < k || k == 0) {
        ArmijoArgs args;
        args.k = std::min(ARM_K, k - done);
        for (int q = 0; q < ARM_K; q++) args.a[q] = q < args.k ? alphas[done + q] : 0.0;
        k_armijo<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->sigma, args, h->lambda, h->lambda_ub, h->pvio_raw, h->A_RD, h->A_DD,
                                                           h->y, h->partials, h->ticket, h->dscal + SC_LANCZOS);
        KLAUNCH(h);
        SDP_CHECK(comm_reduce_scalars(h, SC_LANCZOS, ARM_K + 1));
        k_armijo_final<<<1, 1, 0, h->stream>>>(h->m, args, h->pvio_raw, h->A_RD, h->A_DD, h->dscal + SC_LANCZOS);
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
    const CRange sum = range_sum(h);
    k_yb<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->y, h->b, h->partials, h->ticket, h->dscal + SC_LANCZOS + 9);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    SDP_CHECK(comm_reduce_scalars(h, SC_LANCZOS + 9, 1));
    SDP_CHECK(fetch_scalars(h, SC_LANCZOS + 9, 1));
    *out = -h->hscal[SC_LANCZOS + 9];
    return SDPLRP_OK;
}

int32_t vec_copy2y_lambda(sdplrp_handle *h) {
    const CRange all = range_all(h);
    k_copy2y_lambda<<<red_grid(all.len()), TPB, 0, h->stream>>>(all, h->m, h->lambda, h->y);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    h->y_obj = 1.0;
    h->S_current = false;
    return SDPLRP_OK;
}

int32_t vec_dimacs_sums(sdplrp_handle *h, double *raw_norm2, double *lambda_b) {
    const CRange sum = range_sum(h);
    k_dimacs_sums<<<red_grid(sum.len()), TPB, 0, h->stream>>>(sum, h->pvio_raw, h->lambda, h->b, h->partials, h->ticket,
                                                            h->dscal + SC_LANCZOS + 9);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    SDP_CHECK(comm_reduce_scalars(h, SC_LANCZOS + 9, 2));
    SDP_CHECK(fetch_scalars(h, SC_LANCZOS + 9, 2));
    *raw_norm2 = h->hscal[SC_LANCZOS + 9];
    *lambda_b = h->hscal[SC_LANCZOS + 10];
    return SDPLRP_OK;
}

// see k_tail_rest; launched BEFORE the fused row pass (S_dyn needs the y of these constraints).  Every rank runs it
// (the shared slots stay current everywhere); its share of the norm is counted once, on rank 0.
int32_t vec_tail_rest(sdplrp_handle *h, double alpha, const double *raw_in, double *raw_out, double *pn2_out) {
    const i64 c0 = h->n_sd;
    k_tail_rest<<<red_grid(h->m + 1 - c0), TPB, 0, h->stream>>>(c0, h->m, alpha, h->sigma, h->A_RD, h->A_DD, raw_in, raw_out, h->lambda,
                                                              h->lambda_ub, h->pvio_lb, h->y, pn2_out, h->rank == 0 ? 1.0 : 0.0,
                                                              h->partials, h->ticket, h->dscal);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
