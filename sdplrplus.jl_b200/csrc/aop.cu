// aop.cu -- constraint evaluation A(UU'), A((UV'+VU')/2) and the fused
// line-search pass {A(RD'+DR'), A(DD')} as sampled-dot kernels over the n x r
// factors with a segmented reduction into the (m+1)-vector.
//
// Reference: src/coreop.jl:36-203 (A!, A_sparse!, A_sparse_formUUt!/UVt!, mydot,
// A_symlowrank!, tr_UtAU, tr_UtAV) and src/linesearch.jl:10-16.
//
// Design (not a translation): the reference first materialises one dot per
// triu slot (nnzT doubles) and then runs a sparse transposed mat-vec over the
// per-matrix entry list.  Here every entry computes its own sampled dot from
// the two factor rows (gathered as 128-bit loads by a sub-warp group of lanes,
// one group per entry) and accumulates val*dot in registers, so the nnzT
// scratch round trip disappears and the RD and DD passes share every load.
//   * "short" matrices (<= kLongMatThreshold triu entries: the m diagonal /
//     edge constraints): one lane group per matrix, result stored directly.
//   * "long" matrices (C, identity, D ...): cut into chunks of kChunkEntries
//     entries, one CTA per chunk, partial sums combined per matrix in a fixed
//     order by a second tiny kernel (deterministic, no atomics).
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int TPB = 256;

template <int VEC>
struct Ld;
template <>
struct Ld<1> {
    typedef double T;
    static __device__ __forceinline__ T ld(const double *p) { return __ldg(p); }
    static __device__ __forceinline__ double dot(T a, T b) { return a * b; }
};
template <>
struct Ld<2> {
    typedef double2 T;
    static __device__ __forceinline__ T ld(const double *p) { return ldg2(p); }
    static __device__ __forceinline__ double dot(T a, T b) { return a.x * b.x + a.y * b.y; }
};

// this lane's share of the sampled dots of one entry (row <= col).
// MODE 0: a1 += v2 * <U_row, U_col>
// MODE 1: a1 += v2 * (<U_col,V_row> + <V_col,U_row>) / 2
// MODE 2: a1 += v2 * (<U_col,V_row> + <V_col,U_row>)     (A_RD, already x2)
//         a2 += v2 * <V_row, V_col>                      (A_DD)
template <int MODE, int VEC>
__device__ __forceinline__ void entry_accum(const double *__restrict__ U, const double *__restrict__ V, int r, int nv,
                                            int row, int col, int lg, int G, double v2, double &a1, double &a2,
                                            int own_lo, int own_hi) {
    typedef Ld<VEC> L;
    if (col < own_lo || col >= own_hi) return;  // multi-GPU: the owner of the column computes the entry
    const double *ur = U + (size_t)row * r, *uc = U + (size_t)col * r;
    const double *vr = (MODE == 0) ? nullptr : V + (size_t)row * r;
    const double *vc = (MODE == 0) ? nullptr : V + (size_t)col * r;
    double d1 = 0.0, d2 = 0.0;
    if (row == col) {
        for (int c = lg; c < nv; c += G) {
            typename L::T u = L::ld(ur + c * VEC);
            if (MODE == 0) {
                d1 += L::dot(u, u);
            } else {
                typename L::T v = L::ld(vr + c * VEC);
                d1 += 2.0 * L::dot(u, v);
                if (MODE == 2) d2 += L::dot(v, v);
            }
        }
    } else {
        for (int c = lg; c < nv; c += G) {
            typename L::T a = L::ld(ur + c * VEC), b = L::ld(uc + c * VEC);
            if (MODE == 0) {
                d1 += L::dot(a, b);
            } else {
                typename L::T p = L::ld(vr + c * VEC), q = L::ld(vc + c * VEC);
                d1 += L::dot(b, p) + L::dot(q, a);
                if (MODE == 2) d2 += L::dot(p, q);
            }
        }
    }
    if (MODE == 1) d1 *= 0.5;
    a1 += v2 * d1;
    if (MODE == 2) a2 += v2 * d2;
}

__device__ __forceinline__ double group_sum(double v, int G) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one lane group per short matrix
template <int MODE, int VEC>
__global__ void __launch_bounds__(TPB) k_A_short(int nA, const int *__restrict__ matptr, const int *__restrict__ mat_gid,
                                                 const int *__restrict__ ent_row, const int *__restrict__ ent_col,
                                                 const double *__restrict__ ent_two, const double *__restrict__ U,
                                                 const double *__restrict__ V, int r, int G, double *__restrict__ out1,
                                                 double *__restrict__ out2, int own_lo, int own_hi, int skip_mat,
                                                 const unsigned char *__restrict__ sd_flag) {
    const int nv = r / VEC;
    const int lg = threadIdx.x & (G - 1);
    const int gpw = 32 / G;  // groups per warp
    const long long warp_global = (long long)blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (TPB / 32);
    const int g_in_warp = (threadIdx.x & 31) / G;
    for (long long base = warp_global * gpw; base < nA; base += n_warps * gpw) {  // warp-uniform trip count
        const long long a = base + g_in_warp;
        double a1 = 0.0, a2 = 0.0;
        bool store = false;
        int gid = 0;
        if (a < nA) {
            const int beg = matptr[a], end = matptr[a + 1];
            if (end - beg <= kLongMatThreshold && a != skip_mat && !(sd_flag && sd_flag[a])) {
                store = true;
                gid = mat_gid[a];
                for (int k = beg; k < end; k++)
                    entry_accum<MODE, VEC>(U, V, r, nv, ent_row[k], ent_col[k], lg, G, ent_two[k], a1, a2, own_lo, own_hi);
            }
        }
        a1 = group_sum(a1, G);
        if (MODE == 2) a2 = group_sum(a2, G);
        if (store && lg == 0) {
            out1[gid] = a1;
            if (MODE == 2) out2[gid] = a2;
        }
    }
}

// single-diagonal-entry constraints (Diag(X) = 1 ...): a streaming pass over the factor rows.  Flat mapping
// (thread = one piece of one row, consecutive threads = consecutive addresses), a CTA takes tiles of TPB/nv whole
// rows, the per-row dots <U_i,U_i> / <U_i,V_i> / <V_i,V_i> are combined through shared memory in a fixed order and
// serve every constraint listed for the row (constraint k of row i is internal slot k: coalesced stores).
template <int MODE, int VEC>
__global__ void __launch_bounds__(TPB) k_A_rowc(long long lo, long long hi, const int *__restrict__ rowc_ptr,
                                                const double *__restrict__ rowc_val,
                                                const double *__restrict__ U, const double *__restrict__ V, int r, int G,
                                                double *__restrict__ out1, double *__restrict__ out2) {
    typedef Ld<VEC> L;
    __shared__ double sh1[TPB], sh2[TPB];
    const int nv = r / VEC;
    const int rpt = TPB / nv;                      // whole rows per tile
    const int rl = threadIdx.x / nv, c = threadIdx.x - rl * nv;
    const long long n_rows = hi - lo;
    for (long long t0 = (long long)blockIdx.x * rpt; t0 < n_rows; t0 += (long long)gridDim.x * rpt) {  // CTA-uniform
        const long long i = lo + t0 + rl;
        const bool live = rl < rpt && i < hi;
        double d1 = 0.0, d2 = 0.0;
        int beg = 0, end = 0;
        if (live && c == 0) { beg = rowc_ptr[i]; end = rowc_ptr[i + 1]; }  // issued with the factor loads, not after the barrier
        if (live) {
            typename L::T a = L::ld(U + (size_t)i * r + c * VEC);
            if (MODE == 0) {
                d1 = L::dot(a, a);
            } else {
                typename L::T b = L::ld(V + (size_t)i * r + c * VEC);
                d1 = L::dot(a, b);
                if (MODE == 2) d2 = L::dot(b, b);
            }
        }
        sh1[threadIdx.x] = d1;
        if (MODE == 2) sh2[threadIdx.x] = d2;
        __syncthreads();
        if (live && c == 0) {
            if (end > beg) {
                double s1 = 0.0, s2 = 0.0;
                for (int k = 0; k < nv; k++) { s1 += sh1[threadIdx.x + k]; if (MODE == 2) s2 += sh2[threadIdx.x + k]; }
                for (int k = beg; k < end; k++) {
                    const double val = rowc_val[k];
                    out1[k] = (MODE == 2 ? 2.0 : 1.0) * val * s1;  // MODE 2: A_RD is kept already doubled
                    if (MODE == 2) out2[k] = val * s2;
                }
            }
        }
        __syncthreads();
    }
}

// The same pass without a CTA barrier (nv <= 32; default): a warp takes 32/nv whole rows per step (6 at r = 10: 30 lanes,
// 480 contiguous bytes per factor), kRowcUnroll steps per trip with every factor load of the trip issued before the first
// use; the pieces of a row are combined by shuffles in the order of the shared-memory kernel above (piece 0, 1, 2, ...:
// the two kernels give the same bits), and the lane of piece 0 serves the row's constraints.  The tile kernel above
// serialises factor loads -> barrier -> (ptr -> val -> stores) -> barrier per CTA (0.68 of the copy peak on C5); here the
// warps of an SM are at different stages, so the dependent ptr -> val -> store tail of one warp runs under the loads of
// the others.  Minimum 4 CTAs per SM = a 64-register budget: with __launch_bounds__(TPB) alone ptxas targets 48 registers
// and interleaves the load pairs with their DMUL/DFMA (one round trip per step); at 64 all 2 x kRowcUnroll 128-bit loads of
// a trip are issued back to back (checked in the SASS).
constexpr int kRowcUnroll = 4;
template <int MODE, int VEC>
__global__ void __launch_bounds__(TPB, 4) k_A_rowc_warp(long long lo, long long hi, const int *__restrict__ rowc_ptr,
                                                     const double *__restrict__ rowc_val,
                                                     const double *__restrict__ U, const double *__restrict__ V, int r,
                                                     double *__restrict__ out1, double *__restrict__ out2) {
    typedef Ld<VEC> L;
    const int nv = r / VEC;
    const int rpw = 32 / nv;                       // whole rows per warp step
    const int lane = threadIdx.x & 31;
    const int rl = lane / nv, c = lane - rl * nv;
    const bool act = rl < rpw;
    const long long n_rows = hi - lo;
    const long long per_trip = (long long)rpw * kRowcUnroll;
    const long long warp = (long long)blockIdx.x * (TPB / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (TPB / 32);
    for (long long t0 = warp * per_trip; t0 < n_rows; t0 += n_warps * per_trip) {   // warp-uniform
        double d1[kRowcUnroll], d2[kRowcUnroll];
        int beg[kRowcUnroll], end[kRowcUnroll];
        typename L::T a[kRowcUnroll], b[kRowcUnroll];
        // every load of the trip before the first use: lanes without a live row read (and discard) piece c of the last
        // row, so that no load sits behind a branch
#pragma unroll
        for (int u = 0; u < kRowcUnroll; u++) {
            const long long i = lo + t0 + (long long)u * rpw + rl;
            const bool live = act && i < hi;
            const long long is = live ? i : hi - 1;
            beg[u] = 0; end[u] = 0;
            if (live && c == 0) { beg[u] = rowc_ptr[i]; end[u] = rowc_ptr[i + 1]; }
            a[u] = L::ld(U + (size_t)is * r + c * VEC);
            if (MODE != 0) b[u] = L::ld(V + (size_t)is * r + c * VEC);
        }
#pragma unroll
        for (int u = 0; u < kRowcUnroll; u++) {
            d2[u] = 0.0;
            if (MODE == 0) {
                d1[u] = L::dot(a[u], a[u]);
            } else {
                d1[u] = L::dot(a[u], b[u]);
                if (MODE == 2) d2[u] = L::dot(b[u], b[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kRowcUnroll; u++) {
            double s1 = 0.0 + d1[u], s2 = 0.0 + d2[u];   // (0 + piece 0) + piece 1 + ...: the order of k_A_rowc
            for (int k = 1; k < nv; k++) {               // lanes of a row are lane .. lane + nv - 1 < 32 for c == 0
                s1 += __shfl_down_sync(0xffffffffu, d1[u], k);
                if (MODE == 2) s2 += __shfl_down_sync(0xffffffffu, d2[u], k);
            }
            for (int k = beg[u]; k < end[u]; k++) {      // empty unless this lane holds piece 0 of a live row
                const double val = rowc_val[k];
                out1[k] = (MODE == 2 ? 2.0 : 1.0) * val * s1;  // MODE 2: A_RD is kept already doubled
                if (MODE == 2) out2[k] = val * s2;
            }
        }
    }
}

// one CTA per chunk of a long matrix
template <int MODE, int VEC>
__global__ void __launch_bounds__(TPB) k_A_long(const int *__restrict__ chunk_mat, const int *__restrict__ long_mat,
                                                const int *__restrict__ long_chunk_ptr, const int *__restrict__ matptr,
                                                const int *__restrict__ ent_row, const int *__restrict__ ent_col,
                                                const double *__restrict__ ent_two, const double *__restrict__ U,
                                                const double *__restrict__ V, int r, int G, double *__restrict__ chunk_part,
                                                int own_lo, int own_hi, int skip_mat) {
    const int c = blockIdx.x;
    const int l = chunk_mat[c];
    const int a = long_mat[l];
    if (a == skip_mat) return;  // the objective is handled by the CD = C*D pass
    const int beg = matptr[a] + (c - long_chunk_ptr[l]) * kChunkEntries;
    const int end = min(beg + kChunkEntries, matptr[a + 1]);
    const int nv = r / VEC;
    const int lg = threadIdx.x & (G - 1);
    const int gpb = TPB / G;
    double acc[2] = {0.0, 0.0};
    for (int k = beg + threadIdx.x / G; k < end; k += gpb)
        entry_accum<MODE, VEC>(U, V, r, nv, ent_row[k], ent_col[k], lg, G, ent_two[k], acc[0], acc[1], own_lo, own_hi);
    block_sum<2>(acc);
    if (threadIdx.x == 0) {
        chunk_part[2 * (size_t)c] = acc[0];
        chunk_part[2 * (size_t)c + 1] = acc[1];
    }
}

// fixed-order combination of the chunk partials of each long matrix
template <int MODE>
__global__ void __launch_bounds__(TPB) k_A_long_combine(const int *__restrict__ long_mat, const int *__restrict__ long_chunk_ptr,
                                                        const int *__restrict__ mat_gid, const double *__restrict__ chunk_part,
                                                        double *__restrict__ out1, double *__restrict__ out2, int skip_mat) {
    const int l = blockIdx.x;
    if (long_mat[l] == skip_mat) return;
    double acc[2] = {0.0, 0.0};
    for (int c = long_chunk_ptr[l] + threadIdx.x; c < long_chunk_ptr[l + 1]; c += TPB) {
        acc[0] += chunk_part[2 * (size_t)c];
        acc[1] += chunk_part[2 * (size_t)c + 1];
    }
    block_sum<2>(acc);
    if (threadIdx.x == 0) {
        const int gid = mat_gid[long_mat[l]];
        out1[gid] = acc[0];
        if (MODE == 2) out2[gid] = acc[1];
    }
}

// ---- low-rank matrices: X'B projections (tall-skinny, memory-bound) --------
// part[block][k*r+i] = sum over this block's rows j of X[j*r+i] * B[j + k*n], k < s (s <= 8 per launch)
constexpr int kLrS = 8;
__global__ void __launch_bounds__(TPB) k_lr_proj(i64 n, int r, int s, const double *__restrict__ X,
                                                 const double *__restrict__ B, i64 ldb, double *__restrict__ part,
                                                 unsigned *__restrict__ ticket, double *__restrict__ out) {
    __shared__ double sm[TPB];
    const int rpb = TPB / r;  // rows per block step
    const int i = threadIdx.x % r, jo = threadIdx.x / r;
    const bool active = jo < rpb;
    double acc[kLrS];
#pragma unroll
    for (int k = 0; k < kLrS; k++) acc[k] = 0.0;
    if (active) {
        for (i64 j = (i64)blockIdx.x * rpb + jo; j < n; j += (i64)gridDim.x * rpb) {
            const double x = X[j * r + i];
#pragma unroll
            for (int k = 0; k < kLrS; k++)
                if (k < s) acc[k] += x * __ldg(&B[j + k * ldb]);
        }
    }
    for (int k = 0; k < s; k++) {
        __syncthreads();
        sm[threadIdx.x] = active ? acc[k] : 0.0;
        __syncthreads();
        if (threadIdx.x < r) {
            double t = 0.0;
            for (int q = 0; q < rpb; q++) t += sm[q * r + threadIdx.x];
            part[(size_t)blockIdx.x * (r * s) + k * r + threadIdx.x] = t;
        }
    }
    __shared__ bool is_last;
    __threadfence();  // every writer publishes its partials before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        for (int e = threadIdx.x; e < r * s; e += TPB) {
            double t = 0.0;
            for (unsigned bI = 0; bI < gridDim.x; bI++) t += __ldcg(&part[(size_t)bI * (r * s) + e]);
            out[e] = t;
        }
        if (threadIdx.x == 0) *ticket = 0u;
    }
}

// out[gid] = sum_{i,k} UB[i,k]*VB[i,k]*D[k] * mult   (tr_UtAU / tr_UtAV)
__global__ void k_lr_trace(int r, int s, const double *__restrict__ UB, const double *__restrict__ VB,
                           const double *__restrict__ Dg, double mult, double *__restrict__ out, int gid) {
    double acc[1] = {0.0};
    for (int e = threadIdx.x; e < r * s; e += blockDim.x) acc[0] += UB[e] * VB[e] * Dg[e / r];
    block_sum<1>(acc);
    if (threadIdx.x == 0) out[gid] = acc[0] * mult;
}

int pick_group(int nv) {
    int G = 1;
    while (G < nv && G < 32) G <<= 1;
    return G;
}

template <int MODE>
int32_t run_sparse(sdplrp_handle *h, const double *U, const double *V, double *out1, double *out2, int skip_mat = -1) {
    if (h->nA <= 0) return SDPLRP_OK;
    cudaStream_t st = h->stream;
    const int r = h->r;
    const bool vec2 = (r % 2 == 0);
    const int nv = vec2 ? r / 2 : r;
    const int G = pick_group(nv);
    const int gpb = TPB / G;
    const int grid_short = grid_for(h->nA, gpb, 16 * kNumSM);
    const int lo = (int)h->row_lo, hi = (int)h->row_hi;
    const unsigned char *sdf = h->n_sd > 0 ? h->sd_flag : nullptr;
    if (h->nA - h->n_sd - h->n_long > 0) {  // short matrices that are not single diagonal entries
        if (vec2) k_A_short<MODE, 2><<<grid_short, TPB, 0, st>>>((int)h->nA, h->matptr, h->mat_gid, h->ent_row, h->ent_col, h->ent_two, U, V, r, G, out1, out2, lo, hi, skip_mat, sdf);
        else k_A_short<MODE, 1><<<grid_short, TPB, 0, st>>>((int)h->nA, h->matptr, h->mat_gid, h->ent_row, h->ent_col, h->ent_two, U, V, r, G, out1, out2, lo, hi, skip_mat, sdf);
        KLAUNCH(h);
    }
    // several GPUs keep the tile kernel: it is what the multi-GPU parity test saw on 2 / 4 / 8 ranks, and no multi-GPU box was
    // available after the warp kernel was written (same bits on one GPU: test_rowc_kernels_give_the_same_bits)
    if (h->n_sd > 0 && nv <= 32 && h->rowc_kernel == 1 && h->world == 1) {   // barrier-free warp-per-rows pass
        const int grid_rows = grid_for(h->row_hi - h->row_lo, (TPB / 32) * (32 / nv) * kRowcUnroll, 16 * kNumSM);
        if (vec2) k_A_rowc_warp<MODE, 2><<<grid_rows, TPB, 0, st>>>(h->row_lo, h->row_hi, h->rowc_ptr, h->rowc_val, U, V, r, out1, out2);
        else k_A_rowc_warp<MODE, 1><<<grid_rows, TPB, 0, st>>>(h->row_lo, h->row_hi, h->rowc_ptr, h->rowc_val, U, V, r, out1, out2);
        KLAUNCH(h);
    } else if (h->n_sd > 0) {                               // rows of more than 32 pieces (r > 64 even, r > 32 odd)
        const int grid_rows = grid_for(h->row_hi - h->row_lo, TPB / nv, 16 * kNumSM);
        if (vec2) k_A_rowc<MODE, 2><<<grid_rows, TPB, 0, st>>>(h->row_lo, h->row_hi, h->rowc_ptr, h->rowc_val, U, V, r, G, out1, out2);
        else k_A_rowc<MODE, 1><<<grid_rows, TPB, 0, st>>>(h->row_lo, h->row_hi, h->rowc_ptr, h->rowc_val, U, V, r, G, out1, out2);
        KLAUNCH(h);
    }
    if (h->n_chunks > 0) {
        if (vec2) k_A_long<MODE, 2><<<(int)h->n_chunks, TPB, 0, st>>>(h->chunk_mat, h->long_mat, h->long_chunk_ptr, h->matptr, h->ent_row, h->ent_col, h->ent_two, U, V, r, G, h->chunk_part, lo, hi, skip_mat);
        else k_A_long<MODE, 1><<<(int)h->n_chunks, TPB, 0, st>>>(h->chunk_mat, h->long_mat, h->long_chunk_ptr, h->matptr, h->ent_row, h->ent_col, h->ent_two, U, V, r, G, h->chunk_part, lo, hi, skip_mat);
        KLAUNCH(h);
        k_A_long_combine<MODE><<<(int)h->n_long, TPB, 0, st>>>(h->long_mat, h->long_chunk_ptr, h->mat_gid, h->chunk_part, out1, out2, skip_mat);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

}  // namespace

// X'B for one low-rank matrix into dst (r*s doubles, device), k-major: dst[k*r+i]
int32_t lr_project(sdplrp_handle *h, const LowRank &L, const double *X, double *dst) {
    const int r = h->r;
    if (r > TPB) return fail(h, SDPLRP_ERR_ARG, "low-rank path supports r <= 256");
    // each rank projects its own rows; the r x s partial products are summed over the ranks (no factor rows cross NVLink)
    const i64 lo = h->row_lo, nloc = h->row_hi - h->row_lo;
    for (i64 k0 = 0; k0 < L.s; k0 += kLrS) {
        const int sc = (int)std::min<i64>(kLrS, L.s - k0);
        i64 blocks = std::min<i64>(kRedBlocks, std::max<i64>(1, nloc / (TPB / r)));
        blocks = std::max<i64>(1, std::min<i64>(blocks, kPartialsLen / ((i64)r * sc)));
        k_lr_proj<<<(int)blocks, TPB, 0, h->stream>>>(nloc, r, sc, X + lo * r, L.dB + k0 * h->n + lo, h->n, h->partials, h->ticket, dst + k0 * r);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return comm_reduce_ptr(h, dst, (int)(L.s * r));
}

int32_t lr_scratch(sdplrp_handle *h) {
    i64 smax = 1;
    for (const LowRank &L : h->lr) smax = std::max(smax, L.s);
    i64 need = 2 * smax * (i64)h->r;
    if (h->lr_tmp_len < need) {
        SDP_CHECK(dev_alloc(h, &h->lr_tmp, need));
        h->lr_tmp_len = need;
    }
    return SDPLRP_OK;
}

// mode 0: A(UU'), mode 1: A((UV'+VU')/2), mode 2: line search (U=R, V=D)
static int32_t run_lowrank(sdplrp_handle *h, int mode, const double *U, const double *V, double *out1, double *out2) {
    if (h->lr.empty()) return SDPLRP_OK;
    SDP_CHECK(lr_scratch(h));
    const int r = h->r;
    const double once = (h->world == 1 || h->rank == 0) ? 1.0 : 0.0;  // every rank forms the same trace; the slot is all-reduced later
    for (const LowRank &L : h->lr) {
        double *ub = h->lr_tmp, *vb = h->lr_tmp + L.s * r;
        SDP_CHECK(lr_project(h, L, U, ub));
        if (mode != 0) SDP_CHECK(lr_project(h, L, V, vb));
        if (mode == 0) {
            k_lr_trace<<<1, 128, 0, h->stream>>>(r, (int)L.s, ub, ub, L.dD, once, out1, (int)L.gid);
        } else if (mode == 1) {
            k_lr_trace<<<1, 128, 0, h->stream>>>(r, (int)L.s, ub, vb, L.dD, once, out1, (int)L.gid);
        } else {
            k_lr_trace<<<1, 128, 0, h->stream>>>(r, (int)L.s, ub, vb, L.dD, 2.0 * once, out1, (int)L.gid);
            KLAUNCH(h);
            k_lr_trace<<<1, 128, 0, h->stream>>>(r, (int)L.s, vb, vb, L.dD, once, out2, (int)L.gid);
        }
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

int32_t aop_uu_skip(sdplrp_handle *h, const double *U, double *out, bool skip_objective) {
    CUDA_TRY(h, cudaMemsetAsync(out, 0, (size_t)(h->m + 1) * sizeof(double), h->stream));
    SDP_CHECK(run_sparse<0>(h, U, nullptr, out, nullptr, skip_objective ? h->obj_mat : -1));
    return run_lowrank(h, 0, U, nullptr, out, nullptr);
}

int32_t aop_uu(sdplrp_handle *h, const double *U, double *out) { return aop_uu_skip(h, U, out, false); }

int32_t aop_uv(sdplrp_handle *h, const double *U, const double *V, double *out) {
    CUDA_TRY(h, cudaMemsetAsync(out, 0, (size_t)(h->m + 1) * sizeof(double), h->stream));
    SDP_CHECK(run_sparse<1>(h, U, V, out, nullptr));
    return run_lowrank(h, 1, U, V, out, nullptr);
}

int32_t aop_linesearch(sdplrp_handle *h, bool skip_objective) {
    // One GPU: the row-list pass stores every slot [0, n_sd) on every call (each of those constraints belongs to exactly one
    // row), so only the slots after them start from zero (2 x 80 MB of stores per iteration saved on C5).  Several GPUs: a
    // rank stores the slots of its own rows only; everything is zeroed as before.
    const i64 z0 = h->world == 1 ? h->n_sd : 0;
    CUDA_TRY(h, cudaMemsetAsync(h->A_RD + z0, 0, (size_t)(h->m + 1 - z0) * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->A_DD + z0, 0, (size_t)(h->m + 1 - z0) * sizeof(double), h->stream));
    SDP_CHECK(run_sparse<2>(h, h->R, h->D, h->A_RD, h->A_DD, skip_objective ? h->obj_mat : -1));
    return run_lowrank(h, 2, h->R, h->D, h->A_RD, h->A_DD);
}
