// perm.cu -- permuting copies between the reference vertex order (everything that
// crosses the ABI) and the internal hub-first order the kernels work in
// (preprocess.cu, "internal vertex order").  With relabeling off they are plain copies.
#include "common.cuh"

namespace {
constexpr int TPB = 256;

// element (i, c) of an n x ncols array lives at i*rs + c*cs
// FWD: dst[perm[i]] = src[i] (reference -> internal);  !FWD: dst[i] = src[perm[i]] (internal -> reference)
template <bool FWD>
__global__ void k_perm_copy(i64 n, i64 ncols, i64 rs, i64 cs, const int *__restrict__ perm, const double *__restrict__ src,
                            double *__restrict__ dst) {
    const i64 total = n * ncols;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        i64 i, c;
        if (rs >= cs) { i = e / ncols; c = e - i * ncols; }  // row-major: consecutive threads walk a row
        else { c = e / n; i = e - c * n; }                   // column-major: consecutive threads walk a column
        const i64 p = perm[i];
        if (FWD) dst[p * rs + c * cs] = src[i * rs + c * cs];
        else dst[i * rs + c * cs] = src[p * rs + c * cs];
    }
}

// stage[(i - lo)*ncols + c] = src[perm[i]*ncols + c] for the reference rows i in [lo, hi)
__global__ void k_perm_rows_range(i64 lo, i64 hi, i64 ncols, const int *__restrict__ perm, const double *__restrict__ src,
                                  double *__restrict__ dst) {
    const i64 total = (hi - lo) * ncols;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < total; e += (i64)gridDim.x * blockDim.x) {
        const i64 i = e / ncols, c = e - i * ncols;
        dst[e] = src[(size_t)(perm ? perm[lo + i] : lo + i) * ncols + c];
    }
}

__global__ void k_perm_slots(i64 len, const int *__restrict__ map, const double *__restrict__ src, double *__restrict__ dst,
                             bool scatter) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < len; k += (i64)gridDim.x * blockDim.x) {
        if (scatter) dst[map[k]] = src[k]; else dst[k] = src[map[k]];
    }
}
}  // namespace

int32_t perm_stage(sdplrp_handle *h, i64 len) {
    if (h->stage_len >= len) return SDPLRP_OK;
    SDP_CHECK(dev_alloc(h, &h->stage, len));
    h->stage_len = len;
    return SDPLRP_OK;
}

// host (reference order) -> device array in internal order
int32_t perm_upload(sdplrp_handle *h, double *dst_dev, const double *src_host, i64 ncols, bool row_major) {
    const i64 n = h->n, len = n * ncols;
    if (len <= 0) return SDPLRP_OK;
    if (!h->relabeled) {
        CUDA_TRY(h, cudaMemcpyAsync(dst_dev, src_host, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return SDPLRP_OK;
    }
    SDP_CHECK(perm_stage(h, len));
    CUDA_TRY(h, cudaMemcpyAsync(h->stage, src_host, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
    const i64 rs = row_major ? ncols : 1, cs = row_major ? 1 : n;
    k_perm_copy<true><<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(n, ncols, rs, cs, h->perm, h->stage, dst_dev);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

// device array in internal order -> host (reference order)
int32_t perm_download(sdplrp_handle *h, const double *src_dev, double *dst_host, i64 ncols, bool row_major) {
    const i64 n = h->n, len = n * ncols;
    if (len <= 0) return SDPLRP_OK;
    const double *from = src_dev;
    if (h->relabeled) {
        SDP_CHECK(perm_stage(h, len));
        const i64 rs = row_major ? ncols : 1, cs = row_major ? 1 : n;
        k_perm_copy<false><<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(n, ncols, rs, cs, h->perm, src_dev, h->stage);
        KLAUNCH(h);
        CUDA_TRY(h, cudaGetLastError());
        from = h->stage;
    }
    CUDA_TRY(h, cudaMemcpyAsync(dst_host, from, (size_t)len * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

// device -> device, reference order <-> internal order (n x ncols)
int32_t perm_device(sdplrp_handle *h, double *dst, const double *src, i64 ncols, bool row_major, bool to_internal) {
    const i64 n = h->n, len = n * ncols;
    if (len <= 0) return SDPLRP_OK;
    if (!h->relabeled) {
        CUDA_TRY(h, cudaMemcpyAsync(dst, src, (size_t)len * 8, cudaMemcpyDeviceToDevice, h->stream));
        return SDPLRP_OK;
    }
    const i64 rs = row_major ? ncols : 1, cs = row_major ? 1 : n;
    if (to_internal) k_perm_copy<true><<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(n, ncols, rs, cs, h->perm, src, dst);
    else k_perm_copy<false><<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(n, ncols, rs, cs, h->perm, src, dst);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    return SDPLRP_OK;
}

// sparse_S.nzval crosses the ABI in reference slot order; the device keeps internal slot order
int32_t perm_slots_upload(sdplrp_handle *h, double *dst_dev, const double *src_host, i64 len) {
    if (len <= 0) return SDPLRP_OK;
    if (!h->relabeled) {
        CUDA_TRY(h, cudaMemcpyAsync(dst_dev, src_host, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        return SDPLRP_OK;
    }
    SDP_CHECK(perm_stage(h, len));
    CUDA_TRY(h, cudaMemcpyAsync(h->stage, src_host, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
    k_perm_slots<<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(len, h->r2i, h->stage, dst_dev, true);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

int32_t perm_slots_download(sdplrp_handle *h, const double *src_dev, double *dst_host, i64 len) {
    if (len <= 0) return SDPLRP_OK;
    const double *from = src_dev;
    if (h->relabeled) {
        SDP_CHECK(perm_stage(h, len));
        k_perm_slots<<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(len, h->r2i, src_dev, h->stage, false);
        KLAUNCH(h);
        CUDA_TRY(h, cudaGetLastError());
        from = h->stage;
    }
    CUDA_TRY(h, cudaMemcpyAsync(dst_host, from, (size_t)len * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

// ---- m-vectors: reference constraint order (ABI) <-> internal constraint order (cperm) -------------------
int32_t perm_cvec_upload(sdplrp_handle *h, double *dst_dev, const double *src_host, i64 len) {
    if (len <= 0) return SDPLRP_OK;
    SDP_CHECK(perm_stage(h, len));
    CUDA_TRY(h, cudaMemcpyAsync(h->stage, src_host, (size_t)len * 8, cudaMemcpyHostToDevice, h->stream));
    k_perm_slots<<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(len, h->cperm, h->stage, dst_dev, true);  // dst[cperm[g]] = src[g]
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

int32_t perm_cvec_download(sdplrp_handle *h, const double *src_dev, double *dst_host, i64 len) {
    if (len <= 0) return SDPLRP_OK;
    SDP_CHECK(perm_stage(h, len));
    k_perm_slots<<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(len, h->cperm, src_dev, h->stage, false);  // dst[g] = src[cperm[g]]
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(dst_host, h->stage, (size_t)len * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

// ---- several GPUs: contiguous slices of the caller's matrix ------------------------------------------------------------------
// Rank q moves rows [q*S, (q+1)*S) of the caller's (reference-order) matrix over PCIe, S = ceil(n / world); the slices are
// all-gathered over NVLink into the staging array and permuted on the device (upload), or the matrix is completed over
// NVLink and the rank's slice permuted out (download).  n*r/world doubles per rank cross PCIe, contiguous on both sides.
static void slice_range(const sdplrp_handle *h, i64 *S, i64 *lo, i64 *hi) {
    *S = (h->n + h->world - 1) / h->world;
    *lo = std::min<i64>(h->n, (i64)h->rank * *S);
    *hi = std::min<i64>(h->n, *lo + *S);
}

int32_t perm_upload_slice(sdplrp_handle *h, double *dst_dev, const double *src_host, i64 ncols) {
    i64 S, lo, hi;
    slice_range(h, &S, &lo, &hi);
    SDP_CHECK(perm_stage(h, (i64)h->world * S * ncols));
    if (hi > lo) CUDA_TRY(h, cudaMemcpyAsync(h->stage + (size_t)lo * ncols, src_host + (size_t)lo * ncols, (size_t)((hi - lo) * ncols) * 8, cudaMemcpyHostToDevice, h->stream));
    SDP_CHECK(comm_allgather_inplace(h, h->stage, (size_t)(S * ncols)));
    const i64 len = h->n * ncols;
    if (h->relabeled) {
        k_perm_copy<true><<<grid_for(len, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(h->n, ncols, ncols, 1, h->perm, h->stage, dst_dev);
        KLAUNCH(h);
        CUDA_TRY(h, cudaGetLastError());
    } else {
        CUDA_TRY(h, cudaMemcpyAsync(dst_dev, h->stage, (size_t)len * 8, cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}

// src_dev must hold every row (the caller completes it over NVLink first)
int32_t perm_download_slice(sdplrp_handle *h, const double *src_dev, double *dst_host, i64 ncols) {
    i64 S, lo, hi;
    slice_range(h, &S, &lo, &hi);
    if (hi <= lo) return SDPLRP_OK;
    SDP_CHECK(perm_stage(h, (i64)h->world * S * ncols));
    k_perm_rows_range<<<grid_for((hi - lo) * ncols, TPB, kRedBlocks * 4), TPB, 0, h->stream>>>(lo, hi, ncols, h->relabeled ? h->perm : nullptr, src_dev, h->stage);
    KLAUNCH(h);
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaMemcpyAsync(dst_host + (size_t)lo * ncols, h->stage, (size_t)((hi - lo) * ncols) * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SDPLRP_OK;
}
