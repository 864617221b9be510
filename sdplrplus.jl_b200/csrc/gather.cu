// gather.cu -- the gather pass  Y = C*X  (one sparse x dense-factor product per inner iteration, the SpMM of
// src/coreop.jl:260-279 restated as CD = C*D, see gradient.cu) as an asynchronous tile pipeline for sm_100a.
//
// Why (profiles/r1_gather_size_sweep.md): the register kernels of gradient.cu pay four to five DEPENDENT
// memory round trips per row (row list -> ptr pair -> idx/val -> gathers -> epilogue operands) and ptxas splits the gathers
// of a block into dependent batches; the pass cost the same whether the gathered factor was L2-resident or not.  Here no
// load of the pass sits on a register dependency chain:
//   * the CSR of the owned rows is cut ONCE into TILES of at most T consecutive nonzeros made of whole rows (rows longer
//     than T become chunk tiles whose partial sums are combined per row, in chunk order, by k_tile_combine);
//   * every warp runs its own software pipeline over its tiles (tile gw, gw + nw, ...):
//       stage A  the tile's ptr / idx / val spans arrive by cp.async.bulk (UBLKCP, coalesced, 16-byte aligned spans) on an
//                mbarrier, NIV tiles ahead;
//       stage B  once the indices of a tile have landed, its factor rows are GATHERED ASYNCHRONOUSLY into shared memory:
//                one cp.async.bulk of 8r bytes per nonzero (MODE 1) or 16-byte cp.async pieces (MODE 2), completion counted
//                by the stage's mbarrier (complete_tx / cp.async.mbarrier.arrive) -- no registers held, NRS tiles in flight;
//       stage C  the FMAs run out of shared memory: one lane group of r/2 lanes per row (rows of a tile are neighbours in the
//                degree-sorted order, so they have the same length), or several lane groups per row when the tile holds few
//                rows, then the row epilogue (fused line-search dots) with coalesced 16-byte stores.
//   * summation order: the nonzeros of a row in stored order per lane group, lane groups of a split row in group order,
//     chunks in chunk order, CTA sums through the ticketed grid reduction: deterministic, independent of timing.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void cpasync16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cpasync_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

struct TileArgs {
    const int4 *tiles;      // {row0, nrows (< 0: chunk of a long row, slot = ~nrows), nz0, nnz}
    i64 ntiles;
    const int *ptr, *idx;
    const double *val;
    const double *Xg;       // gathered operand, n x r row-major
    const double *X, *Z;    // epilogue operands (rows of the output rows)
    double *Y;
    double *scratch;        // chunk partial sums, n_chunks x r
    int r, nv;              // nv = r / 2 sixteen-byte pieces per row
    int T, niv, nrs;        // nonzeros per tile, stages of the ptr/idx/val ring and of the gathered-rows ring
    int hot_rows;           // gathers of rows < hot_rows carry the L2 evict_last policy, the others evict_first (0: no hints)
    double scale;
    double *partials;
    unsigned *ticket;
    double *out;            // EPI 2 / 4: out[0] = sum <X_i, Y_i>, out[1] = sum <X_i, Z_i>
};

// EPI 0: Y_i = scale * acc      EPI 2: Y_i = acc + the two sums      EPI 4: Y_i += acc, the sums on the total
template <int MODE, int EPI>
__global__ void __launch_bounds__(512) k_tile_gather(TileArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int T = a.T, TC = T + 8, NIV = a.niv, NRS = a.nrs, r = a.r, nv = a.nv;
    const size_t rows_bytes = (size_t)NRS * T * r * 8, val_bytes = (size_t)NIV * TC * 8, i_bytes = (size_t)NIV * TC * 4;
    const size_t per_warp = (rows_bytes + val_bytes + 2 * i_bytes + (size_t)(NIV + 1) * 16 + (size_t)(NIV + NRS) * 8 + 127) & ~(size_t)127;
    unsigned char *base = smem_raw + (size_t)warp * per_warp;
    double *rows_s = reinterpret_cast<double *>(base);
    double *val_s = reinterpret_cast<double *>(base + rows_bytes);
    int *idx_s = reinterpret_cast<int *>(base + rows_bytes + val_bytes);
    int *ptr_s = reinterpret_cast<int *>(base + rows_bytes + val_bytes + i_bytes);
    int4 *desc_s = reinterpret_cast<int4 *>(base + rows_bytes + val_bytes + 2 * i_bytes);
    uint64_t *bar_iv = reinterpret_cast<uint64_t *>(base + rows_bytes + val_bytes + 2 * i_bytes + (size_t)(NIV + 1) * 16);
    uint64_t *bar_g = bar_iv + NIV;
    if (lane == 0) {
        for (int s = 0; s < NIV; s++) mbar_init(bar_iv + s, 1);
        for (int s = 0; s < NRS; s++) mbar_init(bar_g + s, MODE == 2 ? 32 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const i64 gw = (i64)blockIdx.x * nwarp + warp, nw = (i64)gridDim.x * nwarp;
    const i64 nk = gw < a.ntiles ? (a.ntiles - gw + nw - 1) / nw : 0;   // tiles of this warp: gw, gw + nw, ...
    const unsigned long long p_str = policy_evict_first(), p_hot = policy_evict_last();

    // lane 0: descriptor of the next tile to stage, loaded one iteration before it is needed
    int4 dnext = make_int4(0, 0, 0, 0);
    if (lane == 0 && nk > 0) dnext = __ldg(a.tiles + gw);
    auto issue_iv = [&](i64 k) {   // lane 0 only
        const int4 t = dnext;
        if (k + 1 < nk) dnext = __ldg(a.tiles + gw + (k + 1) * nw);
        desc_s[k % (NIV + 1)] = t;
        const int s = (int)(k % NIV);
        const int a0 = t.z & ~3, cnt = ((t.z + t.w - a0) + 3) & ~3;
        const int nrows = t.y < 0 ? 0 : t.y;
        const int p0 = t.x & ~3, pcnt = t.y < 0 ? 0 : ((t.x + nrows + 1 - p0) + 3) & ~3;
        fence_proxy_async();
        mbar_expect_tx(bar_iv + s, (uint32_t)cnt * 12u + (uint32_t)pcnt * 4u);
        if (cnt > 0) {
            bulk_g2s_hint(idx_s + s * TC, a.idx + a0, (uint32_t)cnt * 4u, bar_iv + s, p_str);
            bulk_g2s_hint(val_s + s * TC, a.val + a0, (uint32_t)cnt * 8u, bar_iv + s, p_str);
        }
        if (pcnt > 0) bulk_g2s(ptr_s + s * TC, a.ptr + p0, (uint32_t)pcnt * 4u, bar_iv + s);
    };
    auto issue_gathers = [&](i64 k) {   // whole warp
        const int s = (int)(k % NIV), sr = (int)(k % NRS);
        mbar_wait(bar_iv + s, (uint32_t)((k / NIV) & 1));
        const int4 t = desc_s[k % (NIV + 1)];
        const int *is = idx_s + s * TC + (t.z & 3);
        double *rs = rows_s + (size_t)sr * T * r;
        if (MODE == 1) {
            fence_proxy_async();
            if (lane == 0) mbar_expect_tx(bar_g + sr, (uint32_t)t.w * (uint32_t)(r * 8));
            __syncwarp();
            if (a.hot_rows > 0) {
                for (int j = lane; j < t.w; j += 32) {
                    const int c = is[j];
                    bulk_g2s_hint(rs + (size_t)j * r, a.Xg + (size_t)c * r, (uint32_t)(r * 8), bar_g + sr, c < a.hot_rows ? p_hot : p_str);
                }
            } else {
                for (int j = lane; j < t.w; j += 32) bulk_g2s(rs + (size_t)j * r, a.Xg + (size_t)is[j] * r, (uint32_t)(r * 8), bar_g + sr);
            }
        } else {
            const int pieces = t.w * nv;
            for (int e = lane; e < pieces; e += 32) {
                const int j = e / nv, p = e - j * nv;
                cpasync16(reinterpret_cast<double2 *>(rs) + e, reinterpret_cast<const double2 *>(a.Xg + (size_t)is[j] * r) + p);
            }
            cpasync_arrive_noinc(bar_g + sr);
        }
    };

    const int NG = 32 / nv;                 // lane groups per warp
    const int g = lane / nv, lg = lane - g * nv;
    const bool lane_ok = g < NG;
    double s0 = 0.0, s1 = 0.0;
    auto consume = [&](i64 k) {
        const int s = (int)(k % NIV), sr = (int)(k % NRS);
        const int4 t = desc_s[k % (NIV + 1)];
        const double *vs = val_s + s * TC + (t.z & 3);
        const int *ps = ptr_s + s * TC + (t.x & 3);
        const double2 *rs = reinterpret_cast<const double2 *>(rows_s + (size_t)sr * T * r);
        const bool chunk = t.y < 0;
        const int nrows = chunk ? 1 : t.y;
        const int gpr = nrows >= NG ? 1 : NG / nrows;   // lane groups per row
        const int rpp = NG / gpr;                       // rows per pass
        const int qi = g / gpr, sub = g - qi * gpr;
        bool waited = false;
        for (int q0 = 0; q0 < nrows; q0 += rpp) {
            const int q = q0 + qi;
            const bool act = lane_ok && qi < rpp && q < nrows;
            const size_t off = (size_t)(t.x + q) * r + (size_t)lg * 2;
            // epilogue operands first: they travel while the gathers of the tile land
            double2 xv = make_double2(0, 0), zv = make_double2(0, 0), yv = make_double2(0, 0);
            if (act && !chunk && sub == 0) {
                if (EPI == 2 || EPI == 4) {
                    xv = __ldg(reinterpret_cast<const double2 *>(a.X + off));
                    if (a.Z) zv = __ldg(reinterpret_cast<const double2 *>(a.Z + off));
                }
                if (EPI == 4) yv = *reinterpret_cast<const double2 *>(a.Y + off);
            }
            if (!waited) { mbar_wait(bar_g + sr, (uint32_t)((k / NRS) & 1)); waited = true; }
            int b = 0, e = 0;
            if (act) {
                if (chunk) { b = 0; e = t.w; }
                else { b = ps[q] - t.z; e = ps[q + 1] - t.z; }
            }
            double2 acc = make_double2(0, 0);
            int j = b + sub;
            for (; j + 3 * gpr < e; j += 4 * gpr) {
                const double v0 = vs[j], v1 = vs[j + gpr], v2 = vs[j + 2 * gpr], v3 = vs[j + 3 * gpr];
                const double2 x0 = rs[(size_t)j * nv + lg], x1 = rs[(size_t)(j + gpr) * nv + lg];
                const double2 x2 = rs[(size_t)(j + 2 * gpr) * nv + lg], x3 = rs[(size_t)(j + 3 * gpr) * nv + lg];
                acc.x += v0 * x0.x; acc.y += v0 * x0.y;
                acc.x += v1 * x1.x; acc.y += v1 * x1.y;
                acc.x += v2 * x2.x; acc.y += v2 * x2.y;
                acc.x += v3 * x3.x; acc.y += v3 * x3.y;
            }
            for (; j < e; j += gpr) {
                const double v0 = vs[j];
                const double2 x0 = rs[(size_t)j * nv + lg];
                acc.x += v0 * x0.x; acc.y += v0 * x0.y;
            }
            // lane groups of a split row: sub 0 adds the others in group order (warp-uniform trip count)
            for (int o = 1; o < gpr; o++) {
                const int srcl = (qi * gpr + o) * nv + lg;
                const double ox = __shfl_sync(0xffffffffu, acc.x, srcl & 31), oy = __shfl_sync(0xffffffffu, acc.y, srcl & 31);
                if (sub == 0) { acc.x += ox; acc.y += oy; }
            }
            if (act && sub == 0) {
                if (chunk) {
                    *reinterpret_cast<double2 *>(a.scratch + (size_t)(~t.y) * r + (size_t)lg * 2) = acc;
                } else {
                    if (EPI == 0) { acc.x *= a.scale; acc.y *= a.scale; }
                    if (EPI == 4) { acc.x = yv.x + 1.0 * acc.x; acc.y = yv.y + 1.0 * acc.y; }
                    if (EPI == 2 || EPI == 4) {
                        s0 += acc.x * xv.x + acc.y * xv.y;
                        if (a.Z) s1 += xv.x * zv.x + xv.y * zv.y;
                    }
                    *reinterpret_cast<double2 *>(a.Y + off) = acc;
                }
            }
        }
        if (!waited) mbar_wait(bar_g + sr, (uint32_t)((k / NRS) & 1));   // a tile without rows still consumes its phase
        __syncwarp();
    };

    if (lane == 0) for (i64 k = 0; k < NIV && k < nk; k++) issue_iv(k);
    __syncwarp();
    for (i64 k = 0; k < NRS - 1 && k < nk; k++) issue_gathers(k);
    for (i64 k = 0; k < nk; k++) {
        if (k + NRS - 1 < nk) issue_gathers(k + NRS - 1);   // rows stage (k-1) % NRS was released by consume(k-1)
        consume(k);
        if (lane == 0 && k + NIV < nk) issue_iv(k + NIV);   // ptr/idx/val stage k % NIV is free now
        __syncwarp();
    }
    if (EPI == 2 || EPI == 4) {
        double v[2] = {s0, s1};
        double *out = a.out;
        grid_sum_finalize<2>(v, a.partials, a.ticket, [&](double (&s)[2]) { out[0] = s[0]; out[1] = s[1]; });
    }
}

// long rows: the chunk partials are added in chunk order, then the row epilogue
struct CombineArgs {
    i64 n_long;
    const int *long_rows, *long_cptr;
    const double *scratch, *X, *Z;
    double *Y;
    int r, nv;
    double scale;
    double *partials;
    unsigned *ticket;
    double *out;
};
template <int EPI>
__global__ void __launch_bounds__(256) k_tile_combine(CombineArgs a) {
    const int G = 32;   // one warp per long row, lanes < nv active
    const int lane = threadIdx.x & 31;
    const i64 warp = ((i64)blockIdx.x * blockDim.x + threadIdx.x) / G, nw = (i64)gridDim.x * blockDim.x / G;
    double s0 = 0.0, s1 = 0.0;
    for (i64 q = warp; q < a.n_long; q += nw) {
        if (lane >= a.nv) continue;
        const size_t off = (size_t)a.long_rows[q] * a.r + (size_t)lane * 2;
        double2 acc = make_double2(0, 0);
        for (int c = a.long_cptr[q]; c < a.long_cptr[q + 1]; c++) {
            const double2 p = *reinterpret_cast<const double2 *>(a.scratch + (size_t)c * a.r + (size_t)lane * 2);
            acc.x += p.x; acc.y += p.y;
        }
        if (EPI == 0) { acc.x *= a.scale; acc.y *= a.scale; }
        if (EPI == 4) {
            const double2 yv = *reinterpret_cast<const double2 *>(a.Y + off);
            acc.x = yv.x + 1.0 * acc.x; acc.y = yv.y + 1.0 * acc.y;
        }
        if (EPI == 2 || EPI == 4) {
            const double2 xv = __ldg(reinterpret_cast<const double2 *>(a.X + off));
            s0 += acc.x * xv.x + acc.y * xv.y;
            if (a.Z) {
                const double2 zv = __ldg(reinterpret_cast<const double2 *>(a.Z + off));
                s1 += xv.x * zv.x + xv.y * zv.y;
            }
        }
        *reinterpret_cast<double2 *>(a.Y + off) = acc;
    }
    if (EPI == 2 || EPI == 4) {
        double v[2] = {s0, s1};
        double *out = a.out;
        grid_sum_finalize<2>(v, a.partials, a.ticket, [&](double (&s)[2]) { out[0] = s[0]; out[1] = s[1]; });
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// tile plan of a CSR pattern over a row range (built once per pattern / rank / tile size; a greedy walk over ptr on the host:
// 4(n+1) bytes down, 16 bytes per tile up -- one-time preprocessing, milliseconds at n = 10^7)
void gather_plan_free(GatherPlan &p) {
    dev_free(&p.tiles); dev_free(&p.long_rows); dev_free(&p.long_cptr);
    p = GatherPlan();
}

int32_t gather_plan_build(sdplrp_handle *h, GatherPlan &p, const int *ptr_dev, i64 row_lo, i64 row_hi, int T) {
    if (p.tiles && p.ptr_key == ptr_dev && p.row_lo == row_lo && p.row_hi == row_hi && p.T == T) return SDPLRP_OK;
    gather_plan_free(p);
    const i64 nrow = row_hi - row_lo;
    std::vector<int> ptr((size_t)nrow + 1);
    if (nrow > 0) {
        CUDA_TRY(h, cudaMemcpyAsync(ptr.data(), ptr_dev + row_lo, (size_t)(nrow + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    }
    std::vector<int4> tiles;
    std::vector<int> long_rows, long_cptr;
    tiles.reserve((size_t)((nrow > 0 ? ptr[nrow] - ptr[0] : 0) / std::max(1, T - 16) + 16));
    int n_chunks = 0;
    i64 i = 0;
    while (i < nrow) {
        const int len = ptr[i + 1] - ptr[i];
        if (len > T) {   // long row: chunk tiles
            long_rows.push_back((int)(row_lo + i));
            long_cptr.push_back(n_chunks);
            for (int k0 = ptr[i]; k0 < ptr[i + 1]; k0 += T) {
                tiles.push_back(make_int4((int)(row_lo + i), ~n_chunks, k0, std::min(T, ptr[i + 1] - k0)));
                n_chunks++;
            }
            i++;
            continue;
        }
        i64 j = i;
        int nnz = 0;
        while (j < nrow && (j - i) < T && ptr[j + 1] - ptr[j] <= T && nnz + (ptr[j + 1] - ptr[j]) <= T) { nnz += ptr[j + 1] - ptr[j]; j++; }
        tiles.push_back(make_int4((int)(row_lo + i), (int)(j - i), ptr[i], nnz));
        i = j;
    }
    long_cptr.push_back(n_chunks);
    p.ntiles = (i64)tiles.size();
    p.n_long = (i64)long_rows.size();
    p.n_chunks = n_chunks;
    SDP_CHECK(dev_alloc(h, &p.tiles, std::max<i64>(1, p.ntiles)));
    SDP_CHECK(dev_alloc(h, &p.long_rows, std::max<i64>(1, p.n_long)));
    SDP_CHECK(dev_alloc(h, &p.long_cptr, p.n_long + 1));
    if (p.ntiles) CUDA_TRY(h, cudaMemcpyAsync(p.tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    if (p.n_long) CUDA_TRY(h, cudaMemcpyAsync(p.long_rows, long_rows.data(), long_rows.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(p.long_cptr, long_cptr.data(), long_cptr.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    p.ptr_key = ptr_dev; p.row_lo = row_lo; p.row_hi = row_hi; p.T = T;
    return SDPLRP_OK;
}

// tile size / ring depths / warps per CTA for rank r: ~10 KB of gathered rows per stage, whole CTA within the 227 KB of an SM
static void gather_geometry(const sdplrp_handle *h, int r, int *T, int *niv, int *nrs, int *warps, size_t *smem) {
    int t = h->gather_tile > 0 ? h->gather_tile : 10240 / (8 * r);
    t = std::max(16, std::min(256, t)) & ~7;
    const int NIV = 4, NRS = h->gather_stages > 0 ? h->gather_stages : 2;
    const size_t per_warp = ((size_t)NRS * t * r * 8 + (size_t)NIV * (t + 8) * 16 + (size_t)(NIV + 1) * 16 + (size_t)(NIV + NRS) * 8 + 127) & ~(size_t)127;
    int w = (int)((size_t)(216 * 1024) / per_warp);
    if (h->gather_warps > 0) w = std::min(w, h->gather_warps);
    w = std::max(1, std::min(16, w));
    *T = t; *niv = NIV; *nrs = NRS; *warps = w; *smem = per_warp * w;
}

bool gather_supported(const sdplrp_handle *h) {
    return h->gather_mode > 0 && h->r % 2 == 0 && h->r / 2 <= 32 && h->r >= 2;
}

int gather_tile_size(const sdplrp_handle *h) {
    int T, niv, nrs, w; size_t smem;
    gather_geometry(h, h->r, &T, &niv, &nrs, &w, &smem);
    return T;
}

// Y = C*X over the rows of `plan` (EPI 0: Y = scale*..., 2: + sums out2[0..1] (tiles), out2[2..3] (long rows), 4: accumulate)
int32_t gather_spmm(sdplrp_handle *h, const GatherPlan &plan, const int *ptr, const int *idx, const double *val, const double *Xg,
                    const double *X, const double *Z, double *Y, int epi, double scale, double *sums4) {
    int T, niv, nrs, warps; size_t smem;
    gather_geometry(h, h->r, &T, &niv, &nrs, &warps, &smem);
    if (T != plan.T) return fail(h, SDPLRP_ERR_STATE, "gather plan was built for another tile size");
    cudaStream_t st = h->stream;
    if (plan.n_chunks > 0) {
        const i64 need = plan.n_chunks * (i64)h->r;
        if (h->tile_scratch_len < need) {
            SDP_CHECK(dev_alloc(h, &h->tile_scratch, need));
            h->tile_scratch_len = need;
        }
    }
    TileArgs a = {};
    a.tiles = plan.tiles; a.ntiles = plan.ntiles; a.ptr = ptr; a.idx = idx; a.val = val; a.Xg = Xg; a.X = X; a.Z = Z; a.Y = Y;
    a.scratch = h->tile_scratch; a.r = h->r; a.nv = h->r / 2; a.T = T; a.niv = niv; a.nrs = nrs;
    a.hot_rows = h->gather_hints ? (int)std::min<i64>(tile_hot_rows(h), 0x7fffffff) : 0;
    a.scale = scale; a.partials = h->partials; a.ticket = h->ticket; a.out = sums4;
    const int mode = h->gather_mode;
#define TILE_LAUNCH(MODE, EPI)                                                                                           \
    do {                                                                                                                 \
        auto kern = k_tile_gather<MODE, EPI>;                                                                            \
        if (h->gather_attr_smem[MODE - 1][EPI] < (i64)smem) {                                                            \
            CUDA_TRY(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
            h->gather_attr_smem[MODE - 1][EPI] = (i64)smem;                                                              \
        }                                                                                                                \
        const i64 want = (plan.ntiles + warps - 1) / warps;                                                              \
        kern<<<(int)std::max<i64>(1, std::min<i64>(want, kNumSM)), warps * 32, smem, st>>>(a);                           \
    } while (0)
    if (plan.ntiles > 0) {
        if (mode == 1) { if (epi == 0) TILE_LAUNCH(1, 0); else if (epi == 2) TILE_LAUNCH(1, 2); else TILE_LAUNCH(1, 4); }
        else { if (epi == 0) TILE_LAUNCH(2, 0); else if (epi == 2) TILE_LAUNCH(2, 2); else TILE_LAUNCH(2, 4); }
        KLAUNCH(h);
    } else if (sums4 && epi != 0) {
        CUDA_TRY(h, cudaMemsetAsync(sums4, 0, 2 * sizeof(double), st));
    }
#undef TILE_LAUNCH
    if (plan.n_long > 0) {
        CombineArgs c = {};
        c.n_long = plan.n_long; c.long_rows = plan.long_rows; c.long_cptr = plan.long_cptr; c.scratch = h->tile_scratch;
        c.X = X; c.Z = Z; c.Y = Y; c.r = h->r; c.nv = h->r / 2; c.scale = scale;
        c.partials = h->partials; c.ticket = h->ticket; c.out = sums4 ? sums4 + 2 : nullptr;
        const int grid = grid_for(plan.n_long, 256 / 32, 4 * kNumSM);
        if (epi == 0) k_tile_combine<0><<<grid, 256, 0, st>>>(c);
        else if (epi == 2) k_tile_combine<2><<<grid, 256, 0, st>>>(c);
        else k_tile_combine<4><<<grid, 256, 0, st>>>(c);
        KLAUNCH(h);
    } else if (sums4 && epi != 0) {
        CUDA_TRY(h, cudaMemsetAsync(sums4 + 2, 0, 2 * sizeof(double), st));
    }
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}
