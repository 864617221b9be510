// lbfgs.cu -- the L-BFGS two-loop recursion and the dense BLAS-1 glue of the
// inner iteration as fused dot/axpy kernels with device-side scalars.
//
// Reference: src/lbfgs.jl:52-149 (lbfgs_clear!, lbfgs_dir!, lbfgs_update!) and
// src/sdplr.jl:201-205, 219, 224-228 (descent dot, fallback, Rt += alpha*dirt,
// gradient norm).
//
// Design (not a translation): the reference runs 2h dependent (dot -> host
// scalar -> axpy) pairs, 48*N bytes of traffic and 2h host round trips.  Here
// every axpy is fused with the dot the *next* step needs, rho/alpha/beta stay
// in device memory (the last CTA of each kernel finalises the sum), the
// negation, the y_next = -grad pre-store and the descent dot ride on the last
// axpy: 2h+1 launches, 35*N bytes, no host synchronisation until `descent`.
// The arithmetic is the literal two-loop (no H0 scaling, zeroed slots act as
// identity), only the summation order inside a dot differs.
//
// Default path ("lbfgs_kernel" = 1, numlbfgsvecs <= kGramMaxHist): the same recursion
// run on COEFFICIENTS.  Every vector the two-loop touches is a combination of the
// 2h+1 stored vectors [s_1..s_h, y_1..y_h, g], so the recursion only needs their
// pairwise dot products.  Those are all computed DIRECTLY from the stored vectors
// (never by recurrence): lbfgs_update! is one pass that writes the new pair
// (s = alpha*dir, y += g) and forms the dots of {s_new, y_new, g} with every stored
// vector (11N bytes), lbfgs_dir! is a 1-thread coefficient two-loop followed by one
// pass dir = -sum c_k b_k with the y_next = -g pre-store and the descent dot fused
// (11N bytes): 22N per iteration instead of 40N, 2 reductions instead of 2h+2.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = kRedThreads;

template <int VEC>
struct V;
template <>
struct V<1> {
    typedef double T;
    static __device__ __forceinline__ T ld(const double *p, i64 i) { return p[i]; }
    static __device__ __forceinline__ void st(double *p, i64 i, T v) { p[i] = v; }
    static __device__ __forceinline__ T axpy(double a, T x, T y) { return y + a * x; }
    static __device__ __forceinline__ T neg(T x) { return -x; }
    static __device__ __forceinline__ T scale(double a, T x) { return a * x; }
    static __device__ __forceinline__ double dot(T a, T b) { return a * b; }
};
template <>
struct V<2> {
    typedef double2 T;
    static __device__ __forceinline__ T ld(const double *p, i64 i) { return reinterpret_cast<const double2 *>(p)[i]; }
    static __device__ __forceinline__ void st(double *p, i64 i, T v) { reinterpret_cast<double2 *>(p)[i] = v; }
    static __device__ __forceinline__ T axpy(double a, T x, T y) { return make_double2(y.x + a * x.x, y.y + a * x.y); }
    static __device__ __forceinline__ T neg(T x) { return make_double2(-x.x, -x.y); }
    static __device__ __forceinline__ T scale(double a, T x) { return make_double2(a * x.x, a * x.y); }
    static __device__ __forceinline__ double dot(T a, T b) { return a.x * b.x + a.y * b.y; }
};

#define GRID_STRIDE(i, nu) for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < (nu); i += (i64)gridDim.x * blockDim.x)

// out_scalar = <x, y>
template <int VEC>
__global__ void __launch_bounds__(TPB) k_dot(i64 nu, const double *__restrict__ x, const double *__restrict__ y, double *partials,
                                             unsigned *ticket, double *out) {
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) acc[0] += V<VEC>::dot(V<VEC>::ld(x, i), V<VEC>::ld(y, i));
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

// one step of the two-loop recursion fused with the next step's dot.
// PHASE 1 (first loop, src/lbfgs.jl:93-101):  a = rho*dot_in; out = in - a*w;      a_j = a
// PHASE 2 (second loop, :104-112):            g = a_j - rho*dot_in; out = in + g*w
// PHASE 3 (last step): as PHASE 2, then out = -out (:115-117), ypre = -grad
//          (:121-123) and the fused dot is <out, grad> = descent (src/sdplr.jl:201)
template <int VEC, int PHASE>
__global__ void __launch_bounds__(TPB) k_two_loop(i64 nu, const double *in, const double *__restrict__ w,
                                                  const double *z, double *out, const double *__restrict__ rho_j,
                                                  double *a_j, const double *__restrict__ dot_in, double *dot_out,
                                                  const double *grad, double *ypre, double *partials, unsigned *ticket) {
    double coef;
    if (PHASE == 1) {
        const double a = rho_j[0] * dot_in[0];
        coef = -a;
        if (blockIdx.x == 0 && threadIdx.x == 0) a_j[0] = a;
    } else {
        coef = a_j[0] - rho_j[0] * dot_in[0];
    }
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T d = V<VEC>::axpy(coef, V<VEC>::ld(w, i), V<VEC>::ld(in, i));
        if (PHASE == 3) {
            d = V<VEC>::neg(d);
            typename V<VEC>::T g = V<VEC>::ld(grad, i);
            V<VEC>::st(ypre, i, V<VEC>::neg(g));
            acc[0] += V<VEC>::dot(d, g);
        } else {
            acc[0] += V<VEC>::dot(V<VEC>::ld(z, i), d);
        }
        V<VEC>::st(out, i, d);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { dot_out[0] = s[0]; });
}

// numlbfgsvecs == 0: the reference returns right after copyto!(dir, grad) (src/lbfgs.jl:84-90),
// i.e. dir = +grad and descent = ||grad||^2 >= 0, which then takes the fallback of src/sdplr.jl:202-205
template <int VEC>
__global__ void __launch_bounds__(TPB) k_copy_dir(i64 nu, const double *__restrict__ grad, double *__restrict__ dir,
                                                 double *partials, unsigned *ticket, double *dot_out) {
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T g = V<VEC>::ld(grad, i);
        V<VEC>::st(dir, i, g);
        acc[0] += V<VEC>::dot(g, g);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { dot_out[0] = s[0]; });
}

// lbfgs_update! (src/lbfgs.jl:129-149): s_j = alpha*dir, y_j += grad, rho_j = 1/<y_j,s_j>
template <int VEC>
__global__ void __launch_bounds__(TPB) k_update(i64 nu, double alpha, const double *__restrict__ dir,
                                                const double *__restrict__ grad, double *__restrict__ sj,
                                                double *__restrict__ yj, double *partials, unsigned *ticket, double *rho_j) {
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T s = V<VEC>::scale(alpha, V<VEC>::ld(dir, i));
        typename V<VEC>::T y = V<VEC>::axpy(1.0, V<VEC>::ld(grad, i), V<VEC>::ld(yj, i));
        V<VEC>::st(sj, i, s);
        V<VEC>::st(yj, i, y);
        acc[0] += V<VEC>::dot(y, s);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { rho_j[0] = s[0]; });
}
__global__ void k_invert(double *x) { x[0] = 1.0 / x[0]; }

template <int VEC>
__global__ void __launch_bounds__(TPB) k_axpy(i64 nu, double a, const double *__restrict__ x, double *__restrict__ y) {
    GRID_STRIDE(i, nu) V<VEC>::st(y, i, V<VEC>::axpy(a, V<VEC>::ld(x, i), V<VEC>::ld(y, i)));
}

// y1 += a*x1 ; y2 += a*x2  (R += alpha*D and CR += alpha*CD in one launch)
template <int VEC>
__global__ void __launch_bounds__(TPB) k_axpy2(i64 nu, double a, const double *__restrict__ x1, double *__restrict__ y1,
                                               const double *__restrict__ x2, double *__restrict__ y2) {
    GRID_STRIDE(i, nu) {
        V<VEC>::st(y1, i, V<VEC>::axpy(a, V<VEC>::ld(x1, i), V<VEC>::ld(y1, i)));
        V<VEC>::st(y2, i, V<VEC>::axpy(a, V<VEC>::ld(x2, i), V<VEC>::ld(y2, i)));
    }
}

// G = -G ; D = G   (src/sdplr.jl:203-204)
template <int VEC>
__global__ void __launch_bounds__(TPB) k_neg_copy(i64 nu, double *__restrict__ g, double *__restrict__ d) {
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T v = V<VEC>::neg(V<VEC>::ld(g, i));
        V<VEC>::st(g, i, v);
        V<VEC>::st(d, i, v);
    }
}

template <int VEC>
__global__ void __launch_bounds__(TPB) k_norm2(i64 nu, const double *__restrict__ x, double *partials, unsigned *ticket, double *out) {
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T v = V<VEC>::ld(x, i);
        acc[0] += V<VEC>::dot(v, v);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

template <int VEC>
__global__ void __launch_bounds__(TPB) k_dot2(i64 nu, const double *__restrict__ x, const double *__restrict__ y, double *partials,
                                              unsigned *ticket, double *out) {
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) acc[0] += V<VEC>::dot(V<VEC>::ld(x, i), V<VEC>::ld(y, i));
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&s)[1]) { out[0] = s[0]; });
}

// ---- coefficient-space ("Gram") two-loop ------------------------------------------------
constexpr int kGramMaxHist = 8;
constexpr int kGramNB = 2 * kGramMaxHist + 1;   // basis size bound
struct LbPtrs {
    const double *S[kGramMaxHist];
    const double *Y[kGramMaxHist];
};

// One pass over the basis [S_0..S_{M-1}, Y_0..Y_{M-1}, G] (slots in the ROTATED order the host passes, so that the
// slot being replaced is always position M-1):  dots of three probe vectors with every basis vector
// -> out[p*(2M+1) + k].
//   UPDATE:  position M-1 is replaced by the new pair first (S = alpha*D, Y += G; src/lbfgs.jl:142-148) and the
//            probes are {S_new, Y_new, G} -- static positions, no selects; every load precedes the two stores so
//            that the in-order issue never waits on a store operand with loads still unissued.
//   !UPDATE: probes p0..p2 (positions, -1 = unused) -- refresh path after uploads / out-of-order calls.
template <int VEC, int M, bool UPDATE>
__global__ void __launch_bounds__(TPB) k_gram_pass(i64 nu, double alpha, LbPtrs P, const double *__restrict__ G,
                                                   const double *__restrict__ D, double *Sj_out, double *Yj_out,
                                                   int p0, int p1, int p2, double *partials, unsigned *ticket, double *out) {
    constexpr int NB = 2 * M + 1;
    double acc[3 * NB];
#pragma unroll
    for (int k = 0; k < 3 * NB; k++) acc[k] = 0.0;
    GRID_STRIDE(i, nu) {
        typename V<VEC>::T b[NB];
#pragma unroll
        for (int k = 0; k < M; k++) {
            b[k] = (UPDATE && k == M - 1) ? V<VEC>::ld(D, i) : V<VEC>::ld(P.S[k], i);
            b[M + k] = V<VEC>::ld(P.Y[k], i);
        }
        b[2 * M] = V<VEC>::ld(G, i);
        typename V<VEC>::T pv[3];
        if (UPDATE) {
            b[M - 1] = V<VEC>::scale(alpha, b[M - 1]);
            b[2 * M - 1] = V<VEC>::axpy(1.0, b[2 * M], b[2 * M - 1]);
            V<VEC>::st(Sj_out, i, b[M - 1]);
            V<VEC>::st(Yj_out, i, b[2 * M - 1]);
            pv[0] = b[M - 1]; pv[1] = b[2 * M - 1]; pv[2] = b[2 * M];
        } else {
            const int probe[3] = {p0, p1, p2};
#pragma unroll
            for (int p = 0; p < 3; p++) {
                pv[p] = b[0];
#pragma unroll
                for (int k = 1; k < NB; k++)
                    if (k == probe[p]) pv[p] = b[k];
            }
        }
#pragma unroll
        for (int p = 0; p < 3; p++) {
#pragma unroll
            for (int k = 0; k < NB; k++) acc[p * NB + k] += V<VEC>::dot(pv[p], b[k]);
        }
    }
    grid_sum_finalize<3 * NB>(acc, partials, ticket, [&](double (&sv)[3 * NB]) {
#pragma unroll
        for (int k = 0; k < 3 * NB; k++) out[k] = sv[k];
    });
}

// scatter the probe rows into the symmetric Gram matrix (positions -> slots through `rot`: position q holds slot
// (rot + q) % M); rho_j = 1/<y_j, s_j> after an update (src/lbfgs.jl:146)
__global__ void k_gram_commit(int M, int rot, int p0, int p1, int p2, const double *__restrict__ tmp, double *__restrict__ gram,
                              int upd_j, double *__restrict__ rho) {
    const int NB = 2 * M + 1;
    const int probe[3] = {p0, p1, p2};
    auto true_index = [&](int b) { return b < M ? (rot + b) % M : (b < 2 * M ? M + (rot + b - M) % M : 2 * M); };
    for (int p = 0; p < 3; p++) {
        if (probe[p] < 0) continue;
        const int tp = true_index(probe[p]);
        for (int k = threadIdx.x; k < NB; k += blockDim.x) {
            const double v = tmp[p * NB + k];
            const int tk = true_index(k);
            gram[tp * NB + tk] = v;
            gram[tk * NB + tp] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && upd_j >= 0) rho[upd_j] = 1.0 / gram[(M + upd_j) * NB + upd_j];
}

// the two-loop recursion on coefficients (src/lbfgs.jl:93-117): dir = -sum_k c[k] b_k
__global__ void k_gram_coef(int M, int latest, const double *__restrict__ gram, const double *__restrict__ rho,
                            double *__restrict__ coef) {
    if (threadIdx.x != 0) return;
    const int NB = 2 * M + 1;
    double c[kGramNB], a[kGramMaxHist];
    for (int k = 0; k < NB; k++) c[k] = 0.0;
    c[2 * M] = 1.0;
    int order[kGramMaxHist];
    {
        int j = latest;  // 1-based, newest first
        for (int q = 0; q < M; q++) { order[q] = j - 1; j -= 1; if (j == 0) j = M; }
    }
    for (int q = 0; q < M; q++) {
        const int j = order[q];
        double sd = 0.0;
        for (int k = 0; k < NB; k++) sd += c[k] * gram[j * NB + k];
        a[j] = rho[j] * sd;
        c[M + j] -= a[j];
    }
    for (int q = M - 1; q >= 0; q--) {
        const int j = order[q];
        double yd = 0.0;
        for (int k = 0; k < NB; k++) yd += c[k] * gram[(M + j) * NB + k];
        c[j] += a[j] - rho[j] * yd;
    }
    for (int k = 0; k < NB; k++) coef[k] = c[k];
}

// dir = -sum_k c_k b_k ; y_pre = -g ; descent = <dir, g>
template <int VEC, int M>
__global__ void __launch_bounds__(TPB) k_gram_form(i64 nu, LbPtrs P, const double *__restrict__ G, const double *__restrict__ coef,
                                                   double *dir, double *ypre, int jpre,
                                                   double *partials, unsigned *ticket, double *dot_out) {
    constexpr int NB = 2 * M + 1;
    double c[NB];
#pragma unroll
    for (int k = 0; k < NB; k++) c[k] = coef[k];
    double acc[1] = {0.0};
    GRID_STRIDE(i, nu) {
        const typename V<VEC>::T g = V<VEC>::ld(G, i);
        typename V<VEC>::T d = V<VEC>::scale(c[2 * M], g);
#pragma unroll
        for (int k = 0; k < M; k++) {
            d = V<VEC>::axpy(c[k], V<VEC>::ld(P.S[k], i), d);
            d = V<VEC>::axpy(c[M + k], V<VEC>::ld(P.Y[k], i), d);
        }
        d = V<VEC>::neg(d);
        V<VEC>::st(dir, i, d);
        V<VEC>::st(ypre, i, V<VEC>::neg(g));  // slot jpre is read above before it is overwritten (same element, same thread)
        acc[0] += V<VEC>::dot(d, g);
    }
    grid_sum_finalize<1>(acc, partials, ticket, [&](double (&sv)[1]) { dot_out[0] = sv[0]; });
}

struct Slice {
    i64 off, len, nu;
    int vec;
    int grid;
};
Slice owned(const sdplrp_handle *h) {
    Slice s;
    s.off = h->row_lo * h->r;
    s.len = (h->row_hi - h->row_lo) * h->r;
    s.vec = (s.off % 2 == 0 && s.len % 2 == 0) ? 2 : 1;
    s.nu = s.len / s.vec;
    s.grid = grid_for(s.nu, TPB * 4, kRedBlocks);   // 4 CTAs per SM; measured on C5: 3 / 4 / 5 / 6 / 7 / 8 -> direction 1.40 / 1.35 / 1.34 / 1.74 / 1.44 / 1.36 ms, update 1.40 / 1.31 / 1.36 / 1.31 / 1.35 / 1.31 ms
    return s;
}

#define DISPATCH_VEC(sl, KERNEL, ...)                                              \
    do {                                                                           \
        if ((sl).vec == 2) KERNEL<2><<<(sl).grid, TPB, 0, h->stream>>>(__VA_ARGS__); \
        else KERNEL<1><<<(sl).grid, TPB, 0, h->stream>>>(__VA_ARGS__);             \
        KLAUNCH(h);                                                                \
    } while (0)

template <int PHASE>
void launch_two_loop(sdplrp_handle *h, const Slice &sl, const double *in, const double *w, const double *z, double *out,
                     const double *rho_j, double *a_j, const double *dot_in, double *dot_out, const double *grad, double *ypre) {
    if (sl.vec == 2)
        k_two_loop<2, PHASE><<<sl.grid, TPB, 0, h->stream>>>(sl.nu, in, w, z, out, rho_j, a_j, dot_in, dot_out, grad, ypre, h->partials, h->ticket);
    else
        k_two_loop<1, PHASE><<<sl.grid, TPB, 0, h->stream>>>(sl.nu, in, w, z, out, rho_j, a_j, dot_in, dot_out, grad, ypre, h->partials, h->ticket);
    KLAUNCH(h);
}


bool gram_enabled(const sdplrp_handle *h) { return h->lbfgs_kernel == 1 && h->hist >= 1 && h->hist <= kGramMaxHist; }

// position q of the kernel's basis holds slot (rot + q) % hist
LbPtrs gram_ptrs(const sdplrp_handle *h, i64 off, int rot = 0) {
    LbPtrs P = {};
    for (int q = 0; q < h->hist; q++) {
        const int k = (rot + q) % h->hist;
        P.S[q] = h->Sh[k] + off; P.Y[q] = h->Yh[k] + off;
    }
    return P;
}

// dots of up to three probe vectors with the whole basis (optionally writing the new pair of slot j first)
template <int M>
int32_t gram_pass_m(sdplrp_handle *h, const Slice &sl, bool update, double alpha, int j, int p0, int p1, int p2) {
    // update: rotate the slots so that slot j sits at position M-1 (static probe positions in the kernel)
    const int rot = update ? (j + 1) % M : 0;
    if (update) { p0 = M - 1; p1 = 2 * M - 1; p2 = 2 * M; }
    const LbPtrs P = gram_ptrs(h, sl.off, rot);
    const double *G = h->G + sl.off, *D = h->D + sl.off;
    double *Sj = update ? h->Sh[j] + sl.off : nullptr, *Yj = update ? h->Yh[j] + sl.off : nullptr;
    double *tmp = h->lb_small + kGramNB * kGramNB;
#define GP_LAUNCH(VECN, UPD) k_gram_pass<VECN, M, UPD><<<sl.grid, TPB, 0, h->stream>>>(sl.nu, alpha, P, G, D, Sj, Yj, p0, p1, p2, h->partials, h->ticket, tmp)
    if (sl.vec == 2) { if (update) GP_LAUNCH(2, true); else GP_LAUNCH(2, false); }
    else { if (update) GP_LAUNCH(1, true); else GP_LAUNCH(1, false); }
#undef GP_LAUNCH
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    SDP_CHECK(comm_reduce_ptr(h, tmp, 3 * (2 * M + 1)));
    k_gram_commit<<<1, 32, 0, h->stream>>>(M, rot, p0, p1, p2, tmp, h->lb_small, update ? j : -1, h->dscal + SC_RHO);
    KLAUNCH(h);
    return SDPLRP_OK;
}
int32_t gram_pass(sdplrp_handle *h, const Slice &sl, bool update, double alpha, int j, int p0, int p1, int p2) {
    switch (h->hist) {
    case 1: return gram_pass_m<1>(h, sl, update, alpha, j, p0, p1, p2);
    case 2: return gram_pass_m<2>(h, sl, update, alpha, j, p0, p1, p2);
    case 3: return gram_pass_m<3>(h, sl, update, alpha, j, p0, p1, p2);
    case 4: return gram_pass_m<4>(h, sl, update, alpha, j, p0, p1, p2);
    case 5: return gram_pass_m<5>(h, sl, update, alpha, j, p0, p1, p2);
    case 6: return gram_pass_m<6>(h, sl, update, alpha, j, p0, p1, p2);
    case 7: return gram_pass_m<7>(h, sl, update, alpha, j, p0, p1, p2);
    default: return gram_pass_m<8>(h, sl, update, alpha, j, p0, p1, p2);
    }
}

template <int M>
int32_t gram_form_m(sdplrp_handle *h, const Slice &sl, int jpre) {
    const LbPtrs P = gram_ptrs(h, sl.off);
    const double *coef = h->lb_small + kGramNB * kGramNB + 3 * kGramNB;
    if (sl.vec == 2)
        k_gram_form<2, M><<<sl.grid, TPB, 0, h->stream>>>(sl.nu, P, h->G + sl.off, coef, h->D + sl.off, h->Yh[jpre] + sl.off, jpre,
                                                          h->partials, h->ticket, h->dscal + SC_DESCENT);
    else
        k_gram_form<1, M><<<sl.grid, TPB, 0, h->stream>>>(sl.nu, P, h->G + sl.off, coef, h->D + sl.off, h->Yh[jpre] + sl.off, jpre,
                                                          h->partials, h->ticket, h->dscal + SC_DESCENT);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
int32_t gram_form(sdplrp_handle *h, const Slice &sl, int jpre) {
    switch (h->hist) {
    case 1: return gram_form_m<1>(h, sl, jpre);
    case 2: return gram_form_m<2>(h, sl, jpre);
    case 3: return gram_form_m<3>(h, sl, jpre);
    case 4: return gram_form_m<4>(h, sl, jpre);
    case 5: return gram_form_m<5>(h, sl, jpre);
    case 6: return gram_form_m<6>(h, sl, jpre);
    case 7: return gram_form_m<7>(h, sl, jpre);
    default: return gram_form_m<8>(h, sl, jpre);
    }
}

// bring the Gram matrix in line with the vectors in memory (only after uploads / out-of-order calls)
int32_t gram_refresh(sdplrp_handle *h, const Slice &sl) {
    const int M = h->hist, NB = 2 * M + 1;
    if (!h->gram_pairs_valid) {
        for (int p = 0; p < NB - 1; p += 3)
            SDP_CHECK(gram_pass(h, sl, false, 0.0, -1, p, p + 1 < NB - 1 ? p + 1 : -1, p + 2 < NB - 1 ? p + 2 : -1));
        h->gram_pairs_valid = true;
    }
    if (!h->gram_g_valid) {
        SDP_CHECK(gram_pass(h, sl, false, 0.0, -1, 2 * M, -1, -1));
        h->gram_g_valid = true;
    }
    return SDPLRP_OK;
}

int32_t lb_dir_gram(sdplrp_handle *h) {
    const Slice sl = owned(h);
    const int M = h->hist;
    SDP_CHECK(gram_refresh(h, sl));
    double *coef = h->lb_small + kGramNB * kGramNB + 3 * kGramNB;
    k_gram_coef<<<1, 32, 0, h->stream>>>(M, h->latest, h->lb_small, h->dscal + SC_RHO, coef);
    KLAUNCH(h);
    const int jpre = h->latest % M;
    SDP_CHECK(gram_form(h, sl, jpre));
    SDP_CHECK(comm_reduce_ptr(h, h->dscal + SC_DESCENT, 1));
    h->gram_pairs_valid = false;  // slot jpre now holds y = -g: its rows are rebuilt by the update pass (or a refresh)
    h->gram_prestored = jpre;
    return SDPLRP_OK;
}

int32_t lb_update_gram(sdplrp_handle *h, double alpha) {
    const Slice sl = owned(h);
    const int M = h->hist;
    const int j = h->latest % M;
    // everything except slot j must be current: true when the only stale rows are those of the pre-stored slot j
    if (!h->gram_pairs_valid && h->gram_prestored != j) {
        // unusual call order (e.g. uploads): rebuild the rows of every other slot from memory
        for (int p = 0; p < 2 * M; p++) {
            if (p == j || p == M + j) continue;
            SDP_CHECK(gram_pass(h, sl, false, 0.0, -1, p, -1, -1));
        }
    }
    SDP_CHECK(gram_pass(h, sl, true, alpha, j, j, M + j, 2 * M));
    h->gram_pairs_valid = true; h->gram_g_valid = true; h->gram_prestored = -1;
    h->latest = j + 1;
    return SDPLRP_OK;
}

}  // namespace

// lbfgs_dir!(dirt, his, Gt; negate=true) followed by descent = dot(dirt, Gt)
int32_t lb_dir(sdplrp_handle *h) {
    const Slice sl = owned(h);
    const int m = h->hist;
    const double *grad = h->G + sl.off;
    double *dir = h->D + sl.off;
    double *descent = h->dscal + SC_DESCENT;
    if (m == 0) {
        DISPATCH_VEC(sl, k_copy_dir, sl.nu, grad, dir, h->partials, h->ticket, descent);
        CUDA_TRY(h, cudaGetLastError());
        return comm_reduce_ptr(h, descent, 1);
    }
    if (gram_enabled(h)) return lb_dir_gram(h);
    // slot order: newest -> oldest (1-based j as in the reference)
    int order[kMaxHist];
    {
        int j = h->latest;
        for (int q = 0; q < m; q++) { order[q] = j - 1; j -= 1; if (j == 0) j = m; }
    }
    double *dotA = h->dscal + SC_DOT, *dotB = h->dscal + SC_DOT + 6;  // ping-pong
    // dot(s_newest, grad)
    DISPATCH_VEC(sl, k_dot, sl.nu, h->Sh[order[0]] + sl.off, grad, h->partials, h->ticket, dotA);
    SDP_CHECK(comm_reduce_ptr(h, dotA, 1));
    double *din = dotA, *dout = dotB;
    // first loop: newest -> oldest
    for (int q = 0; q < m; q++) {
        const int j = order[q];
        const double *in = (q == 0) ? grad : dir;
        const double *w = h->Yh[j] + sl.off;
        const double *z = (q + 1 < m) ? h->Sh[order[q + 1]] + sl.off : h->Yh[order[m - 1]] + sl.off;
        launch_two_loop<1>(h, sl, in, w, z, dir, h->dscal + SC_RHO + j, h->dscal + SC_A + j, din, dout, nullptr, nullptr);
        SDP_CHECK(comm_reduce_ptr(h, dout, 1));
        std::swap(din, dout);
    }
    // second loop: oldest -> newest
    const int jpre = h->latest % m;  // 0-based slot mod(latest,h)+1 that receives y = -grad
    for (int q = m - 1; q >= 0; q--) {
        const int j = order[q];
        const double *w = h->Sh[j] + sl.off;
        if (q > 0) {
            const double *z = h->Yh[order[q - 1]] + sl.off;
            launch_two_loop<2>(h, sl, dir, w, z, dir, h->dscal + SC_RHO + j, h->dscal + SC_A + j, din, dout, nullptr, nullptr);
            SDP_CHECK(comm_reduce_ptr(h, dout, 1));
            std::swap(din, dout);
        } else {
            launch_two_loop<3>(h, sl, dir, w, nullptr, dir, h->dscal + SC_RHO + j, h->dscal + SC_A + j, din, descent, grad,
                               h->Yh[jpre] + sl.off);
            SDP_CHECK(comm_reduce_ptr(h, descent, 1));
        }
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t lb_update(sdplrp_handle *h, double alpha) {
    const int m = h->hist;
    if (m == 0) return SDPLRP_OK;
    if (gram_enabled(h)) return lb_update_gram(h, alpha);
    const Slice sl = owned(h);
    const int j = h->latest % m;  // 0-based mod(latest,h)+1
    DISPATCH_VEC(sl, k_update, sl.nu, alpha, h->D + sl.off, h->G + sl.off, h->Sh[j] + sl.off, h->Yh[j] + sl.off, h->partials,
                 h->ticket, h->dscal + SC_RHO + j);
    SDP_CHECK(comm_reduce_ptr(h, h->dscal + SC_RHO + j, 1));
    k_invert<<<1, 1, 0, h->stream>>>(h->dscal + SC_RHO + j);
    KLAUNCH(h);
    h->latest = j + 1;
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

// lbfgs_clear!: zero s, y, rho, a; `latest` is kept (src/lbfgs.jl:52-59)
int32_t lb_clear(sdplrp_handle *h) {
    const size_t bytes = (size_t)h->n * h->r * sizeof(double);
    for (int j = 0; j < h->hist; j++) {
        CUDA_TRY(h, cudaMemsetAsync(h->Sh[j], 0, bytes, h->stream));
        CUDA_TRY(h, cudaMemsetAsync(h->Yh[j], 0, bytes, h->stream));
    }
    CUDA_TRY(h, cudaMemsetAsync(h->dscal + SC_RHO, 0, kMaxHist * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->dscal + SC_A, 0, kMaxHist * sizeof(double), h->stream));
    if (h->lb_small) CUDA_TRY(h, cudaMemsetAsync(h->lb_small, 0, (size_t)kLbSmallLen * sizeof(double), h->stream));
    h->gram_pairs_valid = true;   // all-zero history: every pair dot is zero
    h->gram_g_valid = true;       // ... and so is every <g, s_j>, <g, y_j>; <g, g> is not used by the recursion
    h->gram_prestored = -1;
    return SDPLRP_OK;
}

int32_t lb_axpy(sdplrp_handle *h, double alpha, const double *x, double *y) {
    const Slice sl = owned(h);
    DISPATCH_VEC(sl, k_axpy, sl.nu, alpha, x + sl.off, y + sl.off);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t lb_axpy2(sdplrp_handle *h, double alpha, const double *x1, double *y1, const double *x2, double *y2) {
    const Slice sl = owned(h);
    DISPATCH_VEC(sl, k_axpy2, sl.nu, alpha, x1 + sl.off, y1 + sl.off, x2 + sl.off, y2 + sl.off);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t lb_neg_copy(sdplrp_handle *h) {
    const Slice sl = owned(h);
    DISPATCH_VEC(sl, k_neg_copy, sl.nu, h->G + sl.off, h->D + sl.off);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

// dscal[slot] = <x, y> over the owned rows (the caller all-reduces)
int32_t lb_dot(sdplrp_handle *h, const double *x, const double *y, int slot) {
    const Slice sl = owned(h);
    DISPATCH_VEC(sl, k_dot2, sl.nu, x + sl.off, y + sl.off, h->partials, h->ticket, h->dscal + slot);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t lb_norm2(sdplrp_handle *h, const double *x, int slot) {
    const Slice sl = owned(h);
    DISPATCH_VEC(sl, k_norm2, sl.nu, x + sl.off, h->partials, h->ticket, h->dscal + slot);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
