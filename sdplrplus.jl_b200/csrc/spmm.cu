// spmm.cu -- the sparse x dense-factor product of the hot loop (CD = C*D, the
// ONE gather pass of an inner iteration; also G's off-diagonal constraint part
// and the seam-level At!(Y, X)) as an asynchronous-copy tile-stream kernel.
//
// Reference: src/coreop.jl:260-279 (At!(Y, X, aux, var): dense r x n times sparse
// n x n, column loop of SparseArrays) -- same arithmetic, different machine.
//
// Why this shape (measured on B200, profiles/r1_*): the product gathers one
// r-vector (80 B at r = 10) per nonzero from a factor that is far larger than L2,
// so it is bound by the latency of random 64 B DRAM granules, not by arithmetic.
// A register-staged kernel (gradient.cu, kept as the fallback for odd shapes)
// sustains ~3.7 TB/s of granule traffic because every lane can only hold a few
// gathers in flight.  Here a warp streams its rows in batches of <= kTileNnz
// nonzeros: it reads the batch's column indices coalesced, issues one cp.async
// (LDGSTS, L2-only, no register staging) per 16 B piece of every gathered row
// straight into shared memory, and only then consumes the PREVIOUS batch from
// shared memory -- two batches (128 row gathers) are in flight per warp, ~1000
// per SM.  The rows the epilogue needs (X_i, Z_i / ADD_i) ride on the same
// async-copy group, so the consume phase never waits on global memory.
//   * lanes form teams of r/2 (double2 pieces of one factor row): a team owns a
//     whole short row; for batches with fewer rows than teams the teams split
//     each row's nonzeros and are summed in a fixed order (deterministic).
//   * rows longer than kTileNnz are cut into chunks (one batch each) whose
//     partial sums go to a scratch array and are combined in chunk order by a
//     second small kernel that applies the same epilogue.
//   * L2 policy: the index/value streams and the tail gathers are evict_first,
//     gathers of the leading `hot_rows` (the hubs after the hub-first internal
//     relabeling, preprocess.cu) are evict_last so they stay L2-resident.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int kTileNnz = 64;    // nonzeros per batch buffer
constexpr int kTileRows = 32;   // rows per batch (one lane each for the row-pointer scan)
constexpr int kTaskRows = 128;  // rows per warp task (tasks are strided over all warps of the grid)

__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
template <int VEC>
__device__ __forceinline__ void cp_async_piece(double *dst_smem, const double *src, unsigned long long pol) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    if (VEC == 2) asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "l"(pol) : "memory");
    else asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ int ld_stream_i32(const int *p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double *p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

template <int VEC>
struct Piece;
template <>
struct Piece<1> {
    double v;
    __device__ __forceinline__ void zero() { v = 0.0; }
    __device__ __forceinline__ void fma(double s, const double *p) { v += s * p[0]; }
    __device__ __forceinline__ void add_shfl(const Piece &o, int src_lane) { v += __shfl_sync(0xffffffffu, o.v, src_lane); }
    __device__ __forceinline__ void add(const double *p) { v += p[0]; }
    __device__ __forceinline__ void scale_add(double sc, double a, const double *p) { v = sc * (v + a * p[0]); }
    __device__ __forceinline__ void scale(double sc) { v *= sc; }
    __device__ __forceinline__ double dot(const double *p) const { return v * p[0]; }
    __device__ __forceinline__ double norm2() const { return v * v; }
    __device__ __forceinline__ void store(double *p) const { p[0] = v; }
    __device__ __forceinline__ void store_cs(double *p) const { __stcs(p, v); }
};
template <>
struct Piece<2> {
    double2 v;
    __device__ __forceinline__ void zero() { v.x = v.y = 0.0; }
    __device__ __forceinline__ void fma(double s, const double *p) {
        const double2 x = *reinterpret_cast<const double2 *>(p);
        v.x += s * x.x; v.y += s * x.y;
    }
    __device__ __forceinline__ void add_shfl(const Piece &o, int src_lane) {  // += the (unmodified) partial of another lane
        v.x += __shfl_sync(0xffffffffu, o.v.x, src_lane);
        v.y += __shfl_sync(0xffffffffu, o.v.y, src_lane);
    }
    __device__ __forceinline__ void add(const double *p) {
        const double2 x = *reinterpret_cast<const double2 *>(p);
        v.x += x.x; v.y += x.y;
    }
    __device__ __forceinline__ void scale_add(double sc, double a, const double *p) {
        const double2 x = *reinterpret_cast<const double2 *>(p);
        v.x = sc * (v.x + a * x.x); v.y = sc * (v.y + a * x.y);
    }
    __device__ __forceinline__ void scale(double sc) { v.x *= sc; v.y *= sc; }
    __device__ __forceinline__ double dot(const double *p) const {
        const double2 x = *reinterpret_cast<const double2 *>(p);
        return v.x * x.x + v.y * x.y;
    }
    __device__ __forceinline__ double norm2() const { return v.x * v.x + v.y * v.y; }
    __device__ __forceinline__ void store(double *p) const { *reinterpret_cast<double2 *>(p) = v; }
    __device__ __forceinline__ void store_cs(double *p) const { __stcs(reinterpret_cast<double2 *>(p), v); }
};

// EPI 0: Y_i = scale * acc                                   (seam-level At!)
// EPI 1: Y_i = scale * (acc + yobj*ADD_i); sum0 += |Y_i|^2    (gradient)
// EPI 2: Y_i = acc; sum0 += <X_i, acc>; sum1 += <X_i, Z_i>    (CD = C*D with the line-search dots; CR = C*R with obj)
struct TileArgs {
    const int *ptr;           // row pointer (direct pass)
    const int *chunk_start, *chunk_end, *chunk_row;  // chunk pass
    i64 v_begin, v_end;       // rows (direct) or chunks (partial) of this launch
    const int *idx;
    const double *val;
    const int *src;           // IND: value = val[src[k]]
    const double *X;          // gathered factor, n x r
    const double *E0, *E1;    // epilogue row streams: EPI 1: E0 = ADD; EPI 2: E0 = X (same array), E1 = Z (may be null)
    double *Y;
    double *scratch;          // chunk pass: partial sums, n_chunks x r
    int r, ts, tpw;           // team size (pieces per row) and teams per warp
    int hot_rows;
    double scale, yobj;
    i64 own_lo, own_hi;
    double *partials;
    unsigned *ticket;
    double *out;              // EPI != 0: out[0], out[1]
};

struct Batch {
    i64 v;       // first row / the chunk
    int nb;      // rows in the batch
    int k0;      // first nonzero
    int cnt;     // nonzeros
    int myend;   // lane j < nb: end (exclusive, absolute) of row j's nonzeros
    bool skip;   // direct pass: the single row is long (handled by the chunk pass) or not owned
};

template <int VEC, int EPI>
__device__ __forceinline__ void row_epilogue(const TileArgs &a, i64 i, int piece, Piece<VEC> &acc, const double *e0, const double *e1,
                                             double &s0, double &s1) {
    const size_t off = (size_t)i * a.r + piece * VEC;
    if (EPI == 0) {
        acc.scale(a.scale);
    } else if (EPI == 1) {
        if (e0) acc.scale_add(a.scale, a.yobj, e0); else acc.scale(a.scale);
        s0 += acc.norm2();
    } else {
        s0 += acc.dot(e0);
        if (e1) {
            Piece<VEC> t;
            t.zero(); t.fma(1.0, e0);
            s1 += t.dot(e1);
        }
    }
    acc.store_cs(a.Y + off);
}

// per-warp shared memory, in doubles (kStages ring buffers + two task-descriptor buffers)
constexpr int kStages = 3;        // ring of batch buffers: one being consumed, kStages-1 in flight
constexpr int kInFlight = kStages - 1;
constexpr int kDescInts = 136;    // direct: ptr[kTaskRows + 1]; chunk pass: start/end/row of kTaskChunks chunks
constexpr int kTaskChunks = 32;
__host__ __device__ constexpr int tile_e_rows(int epi, bool partial) { return (partial || epi == 0) ? 0 : kTileRows; }
__host__ __device__ inline int tile_per_warp_doubles(int r, int epi, bool partial) {
    const int e0 = tile_e_rows(epi, partial), e1 = (epi == 2 && !partial) ? kTileRows : 0;
    return kStages * (kTileNnz * r + e0 * r + e1 * r + kTileNnz) + kDescInts;  // kDescInts ints x 2 buffers = kDescInts doubles
}

__device__ __forceinline__ void cp_async_b32(void *dst_smem, const void *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_b64(void *dst_smem, const void *src, unsigned long long pol) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "l"(pol) : "memory");
}

// One warp = one software-pipelined stream of batches; see the file header.
//   iteration k:  issue the gathers of batch k+2 (its column indices were prefetched into registers one
//   iteration ago) -> commit -> form batch k+3 from the task's row pointers (shared memory) and prefetch
//   its indices -> wait for batch k -> consume batch k from shared memory.
// No step of the steady state waits on a dependent global load.
template <int VEC, int EPI, bool IND, bool PARTIAL>
__global__ void __launch_bounds__(128) k_tile_stream(TileArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int r = a.r, ts = a.ts, tpw = a.tpw;
    const int team = lane / ts, piece = lane - team * ts;
    const bool team_ok = team < tpw;
    constexpr int E0R = tile_e_rows(EPI, PARTIAL), E1R = (EPI == 2 && !PARTIAL) ? kTileRows : 0;
    double *xs = smem + (size_t)wib * tile_per_warp_doubles(r, EPI, PARTIAL);
    double *e0s = xs + kStages * kTileNnz * r;
    double *e1s = e0s + kStages * E0R * r;
    double *vs = e1s + kStages * E1R * r;
    int *desc = reinterpret_cast<int *>(vs + kStages * kTileNnz);  // [2][kDescInts]
    const unsigned long long pol_hot = policy_evict_last(), pol_stream = policy_evict_first();

    const i64 n_items = a.v_end - a.v_begin;
    const i64 task_sz = PARTIAL ? kTaskChunks : kTaskRows;
    const i64 n_tasks = (n_items + task_sz - 1) / task_sz;
    const i64 n_warps = (i64)gridDim.x * wpb;
    i64 task = (i64)blockIdx.x * wpb + wib;
    int tbuf = 0, pos = 0, tlen = 0;
    int commits = 0, desc_commit = 0;
    bool started = false;

    auto task_len = [&](i64 t) -> int { return (int)min(task_sz, n_items - t * task_sz); };
    auto load_desc = [&](i64 t, int bufi) {  // async copy of one task's descriptors
        int *dst = desc + bufi * kDescInts;
        const int len = task_len(t);
        const i64 first = a.v_begin + t * task_sz;
        if (PARTIAL) {
            if (lane < len) {
                cp_async_b32(dst + lane, a.chunk_start + first + lane);
                cp_async_b32(dst + kTaskChunks + lane, a.chunk_end + first + lane);
                cp_async_b32(dst + 2 * kTaskChunks + lane, a.chunk_row + first + lane);
            }
        } else {
            for (int j = lane; j <= len; j += 32) cp_async_b32(dst + j, a.ptr + first + j);
        }
    };

    auto next_batch = [&](Batch &b) -> bool {
        if (!started || pos >= tlen) {
            if (started) { task += n_warps; tbuf ^= 1; }
            if (task >= n_tasks) return false;
            if (!started) {
                load_desc(task, tbuf);
                cp_async_commit(); commits++;
                cp_async_wait<0>();
                __syncwarp();
                started = true;
            } else if (commits - desc_commit < kInFlight + 2) {  // short task: its descriptors may still be in flight
                cp_async_wait<0>();
                __syncwarp();
            }
            pos = 0; tlen = task_len(task);
            if (task + n_warps < n_tasks) { load_desc(task + n_warps, tbuf ^ 1); desc_commit = commits; }
        }
        const int *sp = desc + tbuf * kDescInts;
        const i64 first = a.v_begin + task * task_sz;
        if (PARTIAL) {
            b.v = first + pos; b.nb = 1;
            b.k0 = sp[pos];
            b.myend = sp[kTaskChunks + pos];
            b.cnt = b.myend - b.k0;
            const i64 row = sp[2 * kTaskChunks + pos];
            b.skip = row < a.own_lo || row >= a.own_hi;
            if (b.skip) b.cnt = 0;
            pos += 1;
            return true;
        }
        const int base = sp[pos];
        const bool in = pos + lane < tlen;
        const int e = in ? sp[pos + lane + 1] : 0x7fffffff;
        const unsigned m = __ballot_sync(0xffffffffu, in && (e - base <= kTileNnz));
        const int nb = __popc(m);  // m is a run of low bits because ptr is monotone
        b.v = first + pos; b.k0 = base; b.skip = false;
        if (nb == 0) {  // the first row is longer than a tile: it belongs to the chunk pass
            b.nb = 1; b.cnt = 0; b.myend = base; b.skip = true;
            pos += 1;
            return true;
        }
        b.nb = nb; b.myend = e;
        b.cnt = __shfl_sync(0xffffffffu, e, nb - 1) - base;
        pos += nb;
        return true;
    };

    // column indices (and value sources) of a batch into registers: nonzero e = lane and lane + 32
    auto prefetch = [&](const Batch &b, int &ia, int &ib, int &sa, int &sb) {
        ia = ib = 0; sa = sb = 0;
        if (lane < b.cnt) {
            ia = ld_stream_i32(a.idx + b.k0 + lane, pol_stream);
            if (IND) sa = ld_stream_i32(a.src + b.k0 + lane, pol_stream);
        }
        if (lane + 32 < b.cnt) {
            ib = ld_stream_i32(a.idx + b.k0 + lane + 32, pol_stream);
            if (IND) sb = ld_stream_i32(a.src + b.k0 + lane + 32, pol_stream);
        }
    };

    auto issue = [&](const Batch &b, int stage, int ia, int ib, int sa, int sb) {
        double *xb = xs + (size_t)stage * kTileNnz * r;
        double *vb = vs + stage * kTileNnz;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int t = 32 * half;
            const int lim = min(32, b.cnt - t);
            if (lim <= 0) break;
            const int c = half ? ib : ia;
            if (lane < lim) {
                if (IND) cp_async_b64(vb + t + lane, a.val + (half ? sb : sa), pol_stream);
                else cp_async_b64(vb + t + lane, a.val + b.k0 + t + lane, pol_stream);
            }
            for (int q = 0; q < lim; q += tpw) {
                const int el = q + team;
                const int col = __shfl_sync(0xffffffffu, c, el & 31);
                if (team_ok && el < lim)
                    cp_async_piece<VEC>(xb + (size_t)(t + el) * r + piece * VEC, a.X + (size_t)col * r + piece * VEC,
                                        col < a.hot_rows ? pol_hot : pol_stream);
            }
        }
        if (E0R > 0 && !b.skip) {  // the epilogue's row streams: rows [b.v, b.v + nb) are contiguous
            double *e0b = e0s + (size_t)stage * E0R * r, *e1b = e1s + (size_t)stage * E1R * r;
            const int pieces = b.nb * ts;
            for (int p = lane; p < pieces; p += 32) {
                const size_t go = (size_t)b.v * r + (size_t)p * VEC;
                if (a.E0) cp_async_piece<VEC>(e0b + p * VEC, a.E0 + go, pol_stream);
                if (E1R > 0 && a.E1) cp_async_piece<VEC>(e1b + p * VEC, a.E1 + go, pol_stream);
            }
        }
    };

    double s0 = 0.0, s1 = 0.0;
    auto consume = [&](const Batch &b, int stage) {
        if (b.skip) return;
        const double *xb = xs + (size_t)stage * kTileNnz * r;
        const double *vb = vs + stage * kTileNnz;
        const double *e0b = e0s + (size_t)stage * E0R * r, *e1b = e1s + (size_t)stage * E1R * r;
        if (!PARTIAL && b.nb >= tpw) {
            // one team per row
            for (int jj = 0; jj < b.nb; jj += tpw) {
                const int j = jj + team;
                const bool ok = team_ok && j < b.nb;
                const int jc = ok ? j : 0;
                const int end = __shfl_sync(0xffffffffu, b.myend, jc) - b.k0;
                const int prev = __shfl_sync(0xffffffffu, b.myend, (jc + 31) & 31) - b.k0;  // every lane takes part in both shuffles
                const int beg = (jc == 0) ? 0 : prev;
                if (!ok) continue;
                Piece<VEC> acc;
                acc.zero();
                for (int e = beg; e < end; e++) acc.fma(vb[e], xb + (size_t)e * r + piece * VEC);
                const i64 i = b.v + j;
                if (i < a.own_lo || i >= a.own_hi) continue;
                row_epilogue<VEC, EPI>(a, i, piece, acc, (E0R > 0 && a.E0) ? e0b + (size_t)j * r + piece * VEC : nullptr,
                                       (E1R > 0 && a.E1) ? e1b + (size_t)j * r + piece * VEC : nullptr, s0, s1);
            }
        } else {
            // few rows: the teams split each row's nonzeros, fixed-order sum over the teams
            for (int j = 0; j < b.nb; j++) {
                const int end = __shfl_sync(0xffffffffu, b.myend, j) - b.k0;
                const int prev = __shfl_sync(0xffffffffu, b.myend, (j + 31) & 31) - b.k0;
                const int beg = (j == 0) ? 0 : prev;
                Piece<VEC> acc;
                acc.zero();
                if (team_ok)
                    for (int e = beg + team; e < end; e += tpw) acc.fma(vb[e], xb + (size_t)e * r + piece * VEC);
                Piece<VEC> tot = acc;
                for (int t = 1; t < tpw; t++) tot.add_shfl(acc, (piece + t * ts) & 31);  // lanes of team 0 end up with the total
                if (team != 0) continue;
                if (PARTIAL) {
                    tot.store(a.scratch + (size_t)b.v * r + piece * VEC);
                } else {
                    const i64 i = b.v + j;
                    if (i < a.own_lo || i >= a.own_hi) continue;
                    row_epilogue<VEC, EPI>(a, i, piece, tot, (E0R > 0 && a.E0) ? e0b + (size_t)j * r + piece * VEC : nullptr,
                                           (E1R > 0 && a.E1) ? e1b + (size_t)j * r + piece * VEC : nullptr, s0, s1);
                }
            }
        }
    };

    // ---- pipeline (kStages == 3: b0 is consumed, b1 in flight, b2 issued this iteration, b3 prefetched)
    static_assert(kStages == 3, "the rotation below is written for three stages");
    Batch b0, b1, b2, b3;
    int ia = 0, ib = 0, sa = 0, sb = 0;
    bool h0 = next_batch(b0);
    if (h0) { prefetch(b0, ia, ib, sa, sb); issue(b0, 0, ia, ib, sa, sb); }
    cp_async_commit(); commits++;
    bool h1 = h0 && next_batch(b1);
    if (h1) { prefetch(b1, ia, ib, sa, sb); issue(b1, 1, ia, ib, sa, sb); }
    cp_async_commit(); commits++;
    bool h2 = h1 && next_batch(b2);
    if (h2) prefetch(b2, ia, ib, sa, sb);
    int s_cur = 0;
    while (h0) {
        int s_new = s_cur + 2; if (s_new >= kStages) s_new -= kStages;
        if (h2) issue(b2, s_new, ia, ib, sa, sb);
        cp_async_commit(); commits++;
        const bool h3 = h2 && next_batch(b3);
        if (h3) prefetch(b3, ia, ib, sa, sb);
        cp_async_wait<kInFlight>();
        __syncwarp();
        consume(b0, s_cur);
        __syncwarp();
        b0 = b1; h0 = h1; b1 = b2; h1 = h2; b2 = b3; h2 = h3;
        s_cur = (s_cur + 1 == kStages) ? 0 : s_cur + 1;
    }
    cp_async_wait<0>();
    if (!PARTIAL && EPI != 0) {
        double v[2] = {s0, s1};
        double *out = a.out;
        grid_sum_finalize<2>(v, a.partials, a.ticket, [&](double (&s)[2]) { out[0] += s[0]; out[1] += s[1]; });
    }
}

// long rows: sum the chunk partials in chunk order, then the row epilogue
template <int VEC, int EPI>
__global__ void __launch_bounds__(128) k_tile_combine(TileArgs a, i64 n_long, const int *__restrict__ long_rows,
                                                      const int *__restrict__ long_cptr) {
    const int ts = a.ts;
    double s0 = 0.0, s1 = 0.0;
    for (i64 gt = (i64)blockIdx.x * blockDim.x + threadIdx.x; gt < n_long * ts; gt += (i64)gridDim.x * blockDim.x) {
        const i64 l = gt / ts;
        const int piece = (int)(gt - l * ts);
        const i64 i = long_rows[l];
        if (i < a.own_lo || i >= a.own_hi) continue;
        Piece<VEC> acc;
        acc.zero();
        for (int c = long_cptr[l]; c < long_cptr[l + 1]; c++) acc.add(a.scratch + (size_t)c * a.r + piece * VEC);
        const size_t off = (size_t)i * a.r + piece * VEC;
        row_epilogue<VEC, EPI>(a, i, piece, acc, a.E0 ? a.E0 + off : nullptr, (EPI == 2 && a.E1) ? a.E1 + off : nullptr, s0, s1);
    }
    if (EPI != 0) {
        double v[2] = {s0, s1};
        double *out = a.out;
        grid_sum_finalize<2>(v, a.partials, a.ticket, [&](double (&s)[2]) { out[0] += s[0]; out[1] += s[1]; });
    }
}

template <int VEC, int EPI, bool IND>
int32_t launch_tile(sdplrp_handle *h, TileArgs a, const TileLayout &lay, i64 n_rows) {
    cudaStream_t st = h->stream;
    const int r = a.r;
    // warps per CTA from the shared-memory footprint: two or three CTAs of <= ~100 KB per SM
    auto shape = [&](int epi, bool partial, int &wpb, size_t &smem, int &ctas_per_sm) -> bool {
        const size_t per_warp = (size_t)tile_per_warp_doubles(r, epi, partial) * sizeof(double);
        wpb = (int)std::min<size_t>(4, (100 * 1024) / per_warp);
        if (wpb < 1) wpb = (int)std::min<size_t>(4, (200 * 1024) / per_warp);
        if (wpb < 1) return false;
        smem = per_warp * wpb;
        ctas_per_sm = std::max(1, (int)((210 * 1024) / (smem + 1024)));
        return true;
    };
    static bool attr_set[2][3][2][2] = {};
    auto set_attr = [&](auto kern, bool &done) -> int32_t {
        if (!done) {
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done = true;
        }
        return SDPLRP_OK;
    };
    SDP_CHECK(set_attr(k_tile_stream<VEC, EPI, IND, false>, attr_set[VEC - 1][EPI][IND][0]));
    SDP_CHECK(set_attr(k_tile_stream<VEC, 0, IND, true>, attr_set[VEC - 1][0][IND][1]));
    if (EPI != 0) CUDA_TRY(h, cudaMemsetAsync(a.out, 0, 2 * sizeof(double), st));
    int wpb = 0, cps = 0;
    size_t smem = 0;
    // chunk pass + combine for the long rows
    if (lay.n_chunks > 0) {
        if (!shape(0, true, wpb, smem, cps)) return fail(h, SDPLRP_ERR_ARG, "rank too large for the tile-stream kernel");
        const i64 need = lay.n_chunks * (i64)r;
        if (h->tile_scratch_len < need) {
            SDP_CHECK(dev_alloc(h, &h->tile_scratch, need));
            h->tile_scratch_len = need;
        }
        TileArgs c = a;
        c.chunk_start = lay.chunk_start; c.chunk_end = lay.chunk_end; c.chunk_row = lay.chunk_row;
        c.scratch = h->tile_scratch;
        c.v_begin = 0; c.v_end = lay.n_chunks;
        const i64 n_tasks = (lay.n_chunks + kTaskChunks - 1) / kTaskChunks;
        const int grid = (int)std::min<i64>((i64)kNumSM * cps, (n_tasks + wpb - 1) / wpb);
        k_tile_stream<VEC, 0, IND, true><<<grid, wpb * 32, smem, st>>>(c);
        KLAUNCH(h);
        const i64 threads = lay.n_long * a.ts;
        k_tile_combine<VEC, EPI><<<grid_for(threads, 128, 4 * kRedBlocks), 128, 0, st>>>(c, lay.n_long, lay.long_rows, lay.long_cptr);
        KLAUNCH(h);
    }
    // direct pass over every row (long rows are skipped there)
    if (!shape(EPI, false, wpb, smem, cps)) return fail(h, SDPLRP_ERR_ARG, "rank too large for the tile-stream kernel");
    a.v_begin = 0; a.v_end = n_rows;
    if (h->world > 1) { a.v_begin = h->row_lo; a.v_end = h->row_hi; }
    const i64 n_tasks = (a.v_end - a.v_begin + kTaskRows - 1) / kTaskRows;
    if (n_tasks > 0) {
        const int grid = (int)std::min<i64>((i64)kNumSM * cps, (n_tasks + wpb - 1) / wpb);
        k_tile_stream<VEC, EPI, IND, false><<<grid, wpb * 32, smem, st>>>(a);
        KLAUNCH(h);
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

}  // namespace

bool tile_supported(const sdplrp_handle *h) {
    const int r = h->r;
    const int ts = (r % 2 == 0) ? r / 2 : r;
    return h->spmm_kernel == 1 && ts <= 32 && (size_t)tile_per_warp_doubles(r, 2, false) * sizeof(double) <= 200 * 1024;
}

i64 tile_hot_rows(const sdplrp_handle *h) {
    if (h->hot_rows >= 0) return std::min<i64>(h->hot_rows, h->n);
    if (!h->relabeled) return 0;
    if (h->dealt) return h->n;  // multi-GPU deal: hubs are spread over the rank blocks; one policy for every gather, streams evict_first
    // default: ~48 MB of leading factor rows (well inside the 126 MB L2 next to the streams)
    return std::min<i64>(h->n, (i64)(48.0 * 1024 * 1024) / (8 * (i64)std::max(1, h->r)));
}

// Y = epilogue(pattern * X) over the owned rows.  epi 0/1/2 as above; `ind`: values are val[src[k]].
// sums2 (device, 2 doubles) receives the fused sums for epi != 0.
int32_t tile_spmm(sdplrp_handle *h, const TileLayout &lay, const int *ptr, const int *idx, const double *val, const int *src,
                  const double *X, double *Y, int epi, double scale, double yobj, const double *E0, const double *E1,
                  double *sums2) {
    TileArgs a = {};
    a.ptr = ptr; a.idx = idx; a.val = val; a.src = src; a.X = X; a.Y = Y;
    a.E0 = E0; a.E1 = E1; a.scale = scale; a.yobj = yobj;
    a.r = h->r;
    const bool vec2 = (h->r % 2 == 0);
    a.ts = vec2 ? h->r / 2 : h->r;
    a.tpw = 32 / a.ts;
    a.hot_rows = (int)tile_hot_rows(h);
    a.own_lo = h->row_lo; a.own_hi = h->row_hi;
    a.partials = h->partials; a.ticket = h->ticket; a.out = sums2;
    const bool ind = src != nullptr;
#define TILE_CASE(V, E, I) return launch_tile<V, E, I>(h, a, lay, h->n)
    if (vec2) {
        if (epi == 0) { if (ind) TILE_CASE(2, 0, true); else TILE_CASE(2, 0, false); }
        if (epi == 1) { if (ind) TILE_CASE(2, 1, true); else TILE_CASE(2, 1, false); }
        if (ind) TILE_CASE(2, 2, true); else TILE_CASE(2, 2, false);
    } else {
        if (epi == 0) { if (ind) TILE_CASE(1, 0, true); else TILE_CASE(1, 0, false); }
        if (epi == 1) { if (ind) TILE_CASE(1, 1, true); else TILE_CASE(1, 1, false); }
        if (ind) TILE_CASE(1, 2, true); else TILE_CASE(1, 2, false);
    }
#undef TILE_CASE
}
