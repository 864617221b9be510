// blocks.cu -- structured problem blocks: the families of constraint matrices that the reference's problem
// constructors build one `SparseMatrixCOO` at a time (test/problem.jl:16-30 MaxCut / :50-62 Lovasz theta / :80-92 bisection /
// :100-110 cut-norm = exps/problems.jl) are described by a few numbers or short arrays and EXPANDED ON THE DEVICE into the
// concatenated `findnz`-order triplets that sdplrp_preprocess takes.  A 10M-vertex MaxCut then hands over its objective as
// CSC arrays (or builds it on the device) and its 10^7 one-entry constraints as ONE descriptor instead of 10^7 COO objects
// and a 4.7 GB triplet upload.  The expansion only produces input: the maps are built by the same preprocessing as for
// triplets, so they are bit-exact by construction (tests/test_gpu_blocks.py compares them with the triplet path anyway).
#include <vector>
#include "common.cuh"

int32_t sdplrp_preprocess_device(sdplrp_handle *h, int64_t n, int64_t m, int64_t nA, const int64_t *mat_off, const int64_t *d_I,
                                 const int64_t *d_J, const double *d_V, const int64_t *gids);

namespace {
constexpr int TPB = 256;

// matrix k of a DIAG block: one entry (p_k, p_k) = v_k           (super_sparse([i], [i], [1]), test/problem.jl:24)
__global__ void k_expand_diag(i64 count, const int64_t *__restrict__ pos, const double *__restrict__ val, int64_t *__restrict__ I,
                              int64_t *__restrict__ J, double *__restrict__ V) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < count; k += (i64)gridDim.x * blockDim.x) {
        const int64_t p = pos ? pos[k] : k + 1;
        I[k] = p; J[k] = p; V[k] = val ? val[k] : 1.0;
    }
}
// matrix k of an EDGES block: stored entries (u_k, v_k), (v_k, u_k), both = w_k   (super_sparse([i, j], [j, i], [1, 1]), :53-55)
__global__ void k_expand_edges(i64 count, const int64_t *__restrict__ u, const int64_t *__restrict__ v, const double *__restrict__ w,
                               int64_t *__restrict__ I, int64_t *__restrict__ J, double *__restrict__ V) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < count; k += (i64)gridDim.x * blockDim.x) {
        const double x = w ? w[k] : 1.0;
        I[2 * k] = u[k]; J[2 * k] = v[k]; V[2 * k] = x;
        I[2 * k + 1] = v[k]; J[2 * k + 1] = u[k]; V[2 * k + 1] = x;
    }
}
// CSC -> column-major triplets (findnz of a SparseMatrixCSC); colptr / rowval 1-based as Julia holds them
__global__ void k_expand_csc(i64 ncol, const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, const double *__restrict__ nzval,
                             int64_t *__restrict__ I, int64_t *__restrict__ J, double *__restrict__ V) {
    for (i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x; c < ncol; c += (i64)gridDim.x * blockDim.x) {
        for (int64_t k = colptr[c] - 1; k < colptr[c + 1] - 1; k++) { I[k] = rowval[k]; J[k] = c + 1; V[k] = nzval[k]; }
    }
}
__global__ void k_copy_triplets(i64 nnz, const int64_t *__restrict__ i, const int64_t *__restrict__ j, const double *__restrict__ v,
                                int64_t *__restrict__ I, int64_t *__restrict__ J, double *__restrict__ V) {
    for (i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x; k < nnz; k += (i64)gridDim.x * blockDim.x) { I[k] = i[k]; J[k] = j[k]; V[k] = v[k]; }
}

struct Staged {   // a block input made visible to the device (uploaded when it is host memory)
    std::vector<void *> owned;
    ~Staged() { for (void *p : owned) cudaFree(p); }
    template <typename T>
    const T *get(sdplrp_handle *h, const T *p, i64 count, bool on_device, int32_t *rc) {
        if (!p || on_device || count <= 0) return p;
        T *d = nullptr;
        if (cudaMalloc((void **)&d, (size_t)count * sizeof(T)) != cudaSuccess) { h->err = "preprocess_blocks: out of device memory"; *rc = SDPLRP_ERR_CUDA; return nullptr; }
        owned.push_back(d);
        if (cudaMemcpyAsync(d, p, (size_t)count * sizeof(T), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { h->err = "preprocess_blocks: upload failed"; *rc = SDPLRP_ERR_CUDA; }
        return d;
    }
};
}  // namespace

int32_t sdplrp_preprocess_blocks(sdplrp_handle *h, int64_t n, int64_t m, int64_t nblocks, const sdplrp_block *blocks) {
    if (!h) return SDPLRP_ERR_ARG;
    if (nblocks < 0 || (nblocks > 0 && !blocks)) return fail(h, SDPLRP_ERR_ARG, "preprocess_blocks: null block list");
    CUDA_TRY(h, cudaSetDevice(h->device));
    // sizes: matrices and stored entries per block
    i64 nA = 0, nnz = 0;
    for (int64_t b = 0; b < nblocks; b++) {
        const sdplrp_block &B = blocks[b];
        switch (B.kind) {
        case SDPLRP_BLOCK_TRIPLETS: case SDPLRP_BLOCK_CSC:
            if (B.nnz < 0 || (B.nnz > 0 && (!B.I || !B.J || !B.V))) return fail(h, SDPLRP_ERR_ARG, "preprocess_blocks: matrix block without arrays");
            nA += 1; nnz += B.nnz; break;
        case SDPLRP_BLOCK_DIAG:
            if (B.count < 0 || (!B.I && B.count > n)) return fail(h, SDPLRP_ERR_ARG, "preprocess_blocks: DIAG block larger than n");
            nA += B.count; nnz += B.count; break;
        case SDPLRP_BLOCK_EDGES:
            if (B.count < 0 || (B.count > 0 && (!B.I || !B.J))) return fail(h, SDPLRP_ERR_ARG, "preprocess_blocks: EDGES block without endpoints");
            nA += B.count; nnz += 2 * B.count; break;
        case SDPLRP_BLOCK_IDENTITY:
            nA += 1; nnz += n; break;
        default:
            return fail(h, SDPLRP_ERR_ARG, "preprocess_blocks: unknown block kind");
        }
    }
    std::vector<int64_t> mat_off((size_t)nA + 1, 0), gids((size_t)std::max<i64>(nA, 1), 0);
    int64_t *I = nullptr, *J = nullptr;
    double *V = nullptr;
    SDP_CHECK(dev_alloc(h, &I, nnz)); 
    int32_t rc = dev_alloc(h, &J, nnz);
    if (rc == SDPLRP_OK) rc = dev_alloc(h, &V, nnz);
    Staged st;
    i64 a = 0, e = 0;
    const int GS = 8 * kNumSM;
    for (int64_t b = 0; b < nblocks && rc == SDPLRP_OK; b++) {
        const sdplrp_block &B = blocks[b];
        const bool dev = B.on_device != 0;
        if (B.kind == SDPLRP_BLOCK_TRIPLETS) {
            const int64_t *i = st.get(h, B.I, B.nnz, dev, &rc), *j = st.get(h, B.J, B.nnz, dev, &rc);
            const double *v = st.get(h, B.V, B.nnz, dev, &rc);
            if (rc == SDPLRP_OK && B.nnz > 0) { k_copy_triplets<<<grid_for(B.nnz, TPB, GS), TPB, 0, h->stream>>>(B.nnz, i, j, v, I + e, J + e, V + e); KLAUNCH(h); }
            gids[(size_t)a] = B.first_gid; mat_off[(size_t)a + 1] = e + B.nnz; a += 1; e += B.nnz;
        } else if (B.kind == SDPLRP_BLOCK_CSC) {   // I = rowval (nnz), J = colptr (n+1)
            const int64_t *rv = st.get(h, B.I, B.nnz, dev, &rc), *cp = st.get(h, B.J, n + 1, dev, &rc);
            const double *v = st.get(h, B.V, B.nnz, dev, &rc);
            if (rc == SDPLRP_OK && B.nnz > 0) { k_expand_csc<<<grid_for(n, TPB, GS), TPB, 0, h->stream>>>(n, cp, rv, v, I + e, J + e, V + e); KLAUNCH(h); }
            gids[(size_t)a] = B.first_gid; mat_off[(size_t)a + 1] = e + B.nnz; a += 1; e += B.nnz;
        } else if (B.kind == SDPLRP_BLOCK_DIAG) {
            const int64_t *pos = st.get(h, B.I, B.count, dev, &rc);
            const double *v = st.get(h, B.V, B.count, dev, &rc);
            if (rc == SDPLRP_OK && B.count > 0) { k_expand_diag<<<grid_for(B.count, TPB, GS), TPB, 0, h->stream>>>(B.count, pos, v, I + e, J + e, V + e); KLAUNCH(h); }
            for (i64 k = 0; k < B.count; k++) { gids[(size_t)(a + k)] = B.first_gid + k; mat_off[(size_t)(a + k) + 1] = e + k + 1; }
            a += B.count; e += B.count;
        } else if (B.kind == SDPLRP_BLOCK_EDGES) {
            const int64_t *u = st.get(h, B.I, B.count, dev, &rc), *v = st.get(h, B.J, B.count, dev, &rc);
            const double *w = st.get(h, B.V, B.count, dev, &rc);
            if (rc == SDPLRP_OK && B.count > 0) { k_expand_edges<<<grid_for(B.count, TPB, GS), TPB, 0, h->stream>>>(B.count, u, v, w, I + e, J + e, V + e); KLAUNCH(h); }
            for (i64 k = 0; k < B.count; k++) { gids[(size_t)(a + k)] = B.first_gid + k; mat_off[(size_t)(a + k) + 1] = e + 2 * (k + 1); }
            a += B.count; e += 2 * B.count;
        } else {   // IDENTITY: sparse(1.0I, n, n)
            if (n > 0) { k_expand_diag<<<grid_for(n, TPB, GS), TPB, 0, h->stream>>>(n, nullptr, nullptr, I + e, J + e, V + e); KLAUNCH(h); }
            gids[(size_t)a] = B.first_gid; mat_off[(size_t)a + 1] = e + n; a += 1; e += n;
        }
    }
    if (rc == SDPLRP_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) { h->err = "preprocess_blocks: expansion failed"; rc = SDPLRP_ERR_CUDA; }
    if (rc == SDPLRP_OK) rc = sdplrp_preprocess_device(h, n, m, nA, mat_off.data(), I, J, V, gids.data());
    cudaFree(I); cudaFree(J); cudaFree(V);
    return rc;
}
