// Host execution of the gather kernels of csrc/gradient.cu (source extracted verbatim by tests/test_kernel_emulation.py into
// extracted_common.inc / extracted_gradient.inc) under tests/emu/cuda_emu.h.  TEST INFRASTRUCTURE ONLY.
// Every kernel variant must reproduce the default kernels' rows bit for bit (same per-row summation order) and their fused
// sums to rounding; the default kernels are checked against a plain CSR product.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include "cuda_emu.h"

namespace {
constexpr int TPB = 256;
constexpr int kRowGroupMax = 32;
constexpr int kRowWarpMax = 512;
#include "extracted_common.inc"
#include "extracted_gradient.inc"

int pick_group(int nv) { int G = 1; while (G < nv && G < 32) G <<= 1; return G; }

struct Problem {
    int n, r;
    std::vector<int> ptr, idx;
    std::vector<double> val, X, Z;
    std::vector<int> cls[3];
    std::vector<int> chunk_start, chunk_end, chunk_row, long_rows, long_cptr;
};

Problem make(int n, int r, bool sorted, unsigned seed) {
    std::mt19937_64 g(seed);
    Problem P; P.n = n; P.r = r;
    std::vector<int> len(n);
    for (int i = 0; i < n; i++) {
        const unsigned u = g() % 100;
        len[i] = u < 70 ? (int)(g() % 33) : (u < 97 ? 33 + (int)(g() % 200) : 0);
    }
    len[g() % n] = 600; len[g() % n] = 1300; len[g() % n] = 512; len[g() % n] = 513; len[g() % n] = 32; len[g() % n] = 33;
    if (sorted) std::sort(len.begin(), len.end(), std::greater<int>());   // hub-first order: the classes are contiguous ranges
    P.ptr.assign(n + 1, 0);
    for (int i = 0; i < n; i++) P.ptr[i + 1] = P.ptr[i] + len[i];
    const int nnz = P.ptr[n];
    P.idx.resize(nnz); P.val.resize(nnz);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    for (int k = 0; k < nnz; k++) { P.idx[k] = (int)(g() % n); P.val[k] = U(g); }
    P.X.resize((size_t)n * r); P.Z.resize((size_t)n * r);
    for (auto &x : P.X) x = U(g);
    for (auto &x : P.Z) x = U(g);
    for (int i = 0; i < n; i++) P.cls[len[i] <= kRowGroupMax ? 0 : (len[i] <= kRowWarpMax ? 1 : 2)].push_back(i);
    P.long_cptr.push_back(0);
    for (int i : P.cls[2]) {
        P.long_rows.push_back(i);
        for (int b = P.ptr[i]; b < P.ptr[i + 1]; b += kRowWarpMax) {
            P.chunk_start.push_back(b); P.chunk_end.push_back(std::min(b + kRowWarpMax, P.ptr[i + 1])); P.chunk_row.push_back(i);
        }
        P.long_cptr.push_back((int)P.chunk_row.size());
    }
    return P;
}

struct Out { std::vector<double> Y; double sums[6]; };

int g_fail = 0;
void check(bool ok, const char *what, int n, int r, int variant) {
    if (!ok) { std::printf("FAIL %s (n=%d r=%d variant=%d)\n", what, n, r, variant); g_fail++; }
}

// variant 0: default kernels; 1: pipelined (group_pf + warp_pf); 2: bundle (+ warp_pf); 3: batched (group_b + warp_pf)
template <int VEC>
Out run(const Problem &P, int variant, int NB, i64 own_lo, i64 own_hi, bool pad, bool with_z) {
    const int n = P.n, r = P.r, nv = r / VEC;
    Out o; o.Y.assign((size_t)n * r, std::nan(""));
    std::vector<double> partials(1 << 16, 0.0), scratch(std::max<size_t>(1, P.chunk_row.size()) * r, std::nan(""));
    unsigned ticket[4] = {0, 0, 0, 0};
    for (double &s : o.sums) s = 0.0;
    const int ld = pad ? 16 : r;
    std::vector<double> Xp;
    if (pad) { Xp.assign((size_t)n * ld, 0.0); for (int i = 0; i < n; i++) for (int c = 0; c < r; c++) Xp[(size_t)i * ld + c] = P.X[(size_t)i * r + c]; }
    RowArgs a = {};
    a.ptr = P.ptr.data(); a.idx = P.idx.data(); a.val = P.val.data(); a.src = nullptr;
    a.X = P.X.data(); a.Xg = pad ? Xp.data() : P.X.data(); a.ldx = ld;
    a.Y = o.Y.data(); a.Z = with_z ? P.Z.data() : nullptr; a.scale = 1.0;
    a.r = r; a.G = pick_group(nv); a.G0 = nv; a.hot_rows = n / 3;
    a.partials = partials.data(); a.ticket = ticket; a.own_lo = own_lo; a.own_hi = own_hi;
    const bool contig = !P.cls[0].empty() && P.cls[0].back() - P.cls[0].front() + 1 == (int)P.cls[0].size();
    for (int c = 0; c < 3; c++) {
        a.out = o.sums + 2 * c;
        a.rows = P.cls[c].data(); a.n_rows = (i64)P.cls[c].size();
        if (a.n_rows == 0) continue;
        if (c == 0) {
            if (variant == 0) { if (NB == 8) emu::launch(k_rows_group<VEC, 1, false, 2, 8>, 3, TPB, a); else emu::launch(k_rows_group<VEC, 1, false, 2, 4>, 3, TPB, a); }
            else if (variant == 1) { if (NB == 8) emu::launch(k_rows_group_pf<VEC, 8, 2>, 3, TPB, a); else emu::launch(k_rows_group_pf<VEC, 4, 2>, 3, TPB, a); }
            else if (variant == 2 && contig && 32 / a.G0 <= kBundleRows) {
                a.c0_first = P.cls[0].front();
                if (NB == 8) emu::launch(k_rows_bundle<VEC, 8>, 3, TPB_B, a); else emu::launch(k_rows_bundle<VEC, 4>, 3, TPB_B, a);
            } else if (variant == 2) { emu::launch(k_rows_group_pf<VEC, 8, 2>, 2, TPB, a); }
            else { if (NB == 8) emu::launch(k_rows_group_b<VEC, 8>, 3, TPB, a); else emu::launch(k_rows_group_b<VEC, 4>, 3, TPB, a); }
        } else if (c == 1) {
            if (variant == 0) emu::launch(k_rows_warp<VEC, 1, false, 2, false>, 2, TPB, a);
            else emu::launch(k_rows_warp_pf<VEC, false, 2>, 2, TPB, a);
        } else {
            RowArgs b = a;
            b.chunk_start = P.chunk_start.data(); b.chunk_end = P.chunk_end.data(); b.chunk_row = P.chunk_row.data();
            b.long_rows = P.long_rows.data(); b.long_cptr = P.long_cptr.data(); b.scratch = scratch.data();
            b.n_rows = (i64)P.chunk_row.size();
            if (variant == 0) emu::launch(k_rows_warp<VEC, 1, false, 2, true>, 1, TPB, b);
            else emu::launch(k_rows_warp_pf<VEC, true, 2>, 1, TPB, b);
            b.n_rows = (i64)P.long_rows.size();
            emu::launch(k_rows_combine<VEC, 1, 2>, 1, TPB, b);
        }
    }
    check(ticket[0] == 0, "ticket reset", n, r, variant);
    return o;
}

// The two-phase pass (hub | tail columns; grad_obj_spmm with "spmm_phases"): phase one takes the columns < hub of every
// row with a plain store (EPI 0; the long rows whole, chunked), phase two the rest on top with the fused sums (EPI 4; the
// long rows only their epilogue).  pipelined = the k_rows_*_pf kernels, else the default ones.
template <int VEC>
Out run_phases(const Problem &P, bool pipelined, int NB, i64 own_lo, i64 own_hi, int hub) {
    const int n = P.n, r = P.r, nv = r / VEC;
    Out o; o.Y.assign((size_t)n * r, std::nan(""));
    std::vector<double> partials(1 << 16, 0.0), scratch(std::max<size_t>(1, P.chunk_row.size()) * r, std::nan(""));
    unsigned ticket[4] = {0, 0, 0, 0};
    for (double &s : o.sums) s = 0.0;
    // columns ascending inside a row (the library's patterns are; make() draws them at random, so sort a copy)
    std::vector<int> idx = P.idx; std::vector<double> val = P.val; std::vector<int> mid(n), ptr1(P.ptr.begin() + 1, P.ptr.end());
    for (int i = 0; i < n; i++) {
        std::vector<std::pair<int, double>> e;
        for (int k = P.ptr[i]; k < P.ptr[i + 1]; k++) e.push_back({idx[k], val[k]});
        std::stable_sort(e.begin(), e.end(), [](auto &x, auto &y) { return x.first < y.first; });
        int m = P.ptr[i];
        for (int k = P.ptr[i]; k < P.ptr[i + 1]; k++) { idx[k] = e[k - P.ptr[i]].first; val[k] = e[k - P.ptr[i]].second; if (idx[k] < hub) m = k + 1; }
        mid[i] = m;
    }
    RowArgs a = {};
    a.ptr = P.ptr.data(); a.idx = idx.data(); a.val = val.data();
    a.X = P.X.data(); a.Xg = P.X.data(); a.ldx = r; a.Y = o.Y.data(); a.Z = P.Z.data(); a.scale = 1.0;
    a.r = r; a.G = pick_group(nv); a.G0 = nv; a.hot_rows = hub;
    a.partials = partials.data(); a.ticket = ticket; a.own_lo = own_lo; a.own_hi = own_hi;
    for (int phase = 0; phase < 2; phase++) {
        RowArgs p = a;
        if (phase == 0) p.end_arr = mid.data(); else p.beg_arr = mid.data();
        for (int c = 0; c < 3; c++) {
            p.out = o.sums + 2 * c;
            p.rows = P.cls[c].data(); p.n_rows = (i64)P.cls[c].size();
            if (p.n_rows == 0) continue;
            if (c == 0) {
                if (phase == 0) {
                    if (pipelined) { if (NB == 8) emu::launch(k_rows_group_pf<VEC, 8, 0>, 3, TPB, p); else emu::launch(k_rows_group_pf<VEC, 4, 0>, 3, TPB, p); }
                    else emu::launch(k_rows_group<VEC, 1, false, 0, 8>, 3, TPB, p);
                } else {
                    if (pipelined) { if (NB == 8) emu::launch(k_rows_group_pf<VEC, 8, 4>, 3, TPB, p); else emu::launch(k_rows_group_pf<VEC, 4, 4>, 3, TPB, p); }
                    else emu::launch(k_rows_group<VEC, 1, false, 4, 8>, 3, TPB, p);
                }
            } else if (c == 1) {
                if (phase == 0) { if (pipelined) emu::launch(k_rows_warp_pf<VEC, false, 0>, 2, TPB, p); else emu::launch(k_rows_warp<VEC, 1, false, 0, false>, 2, TPB, p); }
                else { if (pipelined) emu::launch(k_rows_warp_pf<VEC, false, 4>, 2, TPB, p); else emu::launch(k_rows_warp<VEC, 1, false, 4, false>, 2, TPB, p); }
            } else if (phase == 0) {   // long rows whole, chunked, plain store
                RowArgs b = p;
                b.beg_arr = nullptr; b.end_arr = nullptr;
                b.chunk_start = P.chunk_start.data(); b.chunk_end = P.chunk_end.data(); b.chunk_row = P.chunk_row.data();
                b.long_rows = P.long_rows.data(); b.long_cptr = P.long_cptr.data(); b.scratch = scratch.data();
                b.n_rows = (i64)P.chunk_row.size();
                if (pipelined) emu::launch(k_rows_warp_pf<VEC, true, 0>, 1, TPB, b); else emu::launch(k_rows_warp<VEC, 1, false, 0, true>, 1, TPB, b);
                b.n_rows = (i64)P.long_rows.size();
                emu::launch(k_rows_combine<VEC, 1, 0>, 1, TPB, b);
            } else {                   // long_empty: only the epilogue with the sums
                RowArgs b = p;
                b.beg_arr = ptr1.data(); b.end_arr = ptr1.data();
                if (pipelined) emu::launch(k_rows_warp_pf<VEC, false, 4>, 1, TPB, b); else emu::launch(k_rows_warp<VEC, 1, false, 4, false>, 1, TPB, b);
            }
        }
    }
    return o;
}

template <int VEC>
void suite_phases(int n, int r, unsigned seed) {
    const Problem P = make(n, r, true, seed);
    const i64 ranges[2][2] = {{0, n}, {n / 4, 3 * n / 4}};
    for (int rg = 0; rg < 2; rg++) {
        const i64 lo = ranges[rg][0], hi = ranges[rg][1];
        const Out ref = run_phases<VEC>(P, false, 8, lo, hi, n / 5);
        double s0 = 0.0, s1 = 0.0;
        for (int i = 0; i < n; i++)
            for (int c = 0; c < r; c++) {
                const double got = ref.Y[(size_t)i * r + c];
                if (i < lo || i >= hi) { check(std::isnan(got), "phases: row outside the owned range written", n, r, 0); continue; }
                double acc = 0.0;
                for (int k = P.ptr[i]; k < P.ptr[i + 1]; k++) acc += P.val[k] * P.X[(size_t)P.idx[k] * r + c];
                check(std::fabs(got - acc) <= 1e-10 * (1.0 + std::fabs(acc)), "phases: default kernels vs CSR product", n, r, 0);
                s0 += acc * P.X[(size_t)i * r + c]; s1 += P.X[(size_t)i * r + c] * P.Z[(size_t)i * r + c];
            }
        const double t0 = ref.sums[0] + ref.sums[2] + ref.sums[4], t1 = ref.sums[1] + ref.sums[3] + ref.sums[5];
        check(std::fabs(t0 - s0) <= 1e-9 * (1.0 + std::fabs(s0)) && std::fabs(t1 - s1) <= 1e-9 * (1.0 + std::fabs(s1)), "phases: default sums", n, r, 0);
        for (int NB : {8, 4}) {
            const Out got = run_phases<VEC>(P, true, NB, lo, hi, n / 5);
            bool same = true;
            for (size_t e = 0; e < ref.Y.size(); e++) {
                const double x = ref.Y[e], y = got.Y[e];
                if (!((std::isnan(x) && std::isnan(y)) || x == y)) same = false;
            }
            check(same, "phases: pipelined rows differ from the default kernels (bitwise)", n, r, NB);
            const double u0 = got.sums[0] + got.sums[2] + got.sums[4], u1 = got.sums[1] + got.sums[3] + got.sums[5];
            check(std::fabs(u0 - t0) <= 1e-10 * (1.0 + std::fabs(t0)) && std::fabs(u1 - t1) <= 1e-10 * (1.0 + std::fabs(t1)), "phases: fused sums differ", n, r, NB);
        }
    }
}

template <int VEC>
void suite(int n, int r, bool sorted, unsigned seed) {
    const Problem P = make(n, r, sorted, seed);
    const i64 ranges[2][2] = {{0, n}, {n / 3, 2 * n / 3 + 1}};
    for (int rg = 0; rg < 2; rg++) {
        const i64 lo = ranges[rg][0], hi = ranges[rg][1];
        for (int with_z = 1; with_z >= 0; with_z--) {
            const Out ref = run<VEC>(P, 0, 8, lo, hi, false, with_z);
            // default kernels against the plain CSR product
            double s0 = 0.0, s1 = 0.0;
            for (int i = 0; i < n; i++) {
                const bool own = i >= lo && i < hi;
                for (int c = 0; c < r; c++) {
                    const double got = ref.Y[(size_t)i * r + c];
                    if (!own) { check(std::isnan(got), "row outside the owned range written", n, r, 0); continue; }
                    double acc = 0.0;
                    for (int k = P.ptr[i]; k < P.ptr[i + 1]; k++) acc += P.val[k] * P.X[(size_t)P.idx[k] * r + c];
                    check(std::fabs(got - acc) <= 1e-11 * (1.0 + std::fabs(acc)), "default kernel vs CSR product", n, r, 0);
                    s0 += acc * P.X[(size_t)i * r + c];
                    if (with_z) s1 += P.X[(size_t)i * r + c] * P.Z[(size_t)i * r + c];
                }
            }
            const double t0 = ref.sums[0] + ref.sums[2] + ref.sums[4], t1 = ref.sums[1] + ref.sums[3] + ref.sums[5];
            check(std::fabs(t0 - s0) <= 1e-10 * (1.0 + std::fabs(s0)) && std::fabs(t1 - s1) <= 1e-10 * (1.0 + std::fabs(s1)), "default sums", n, r, 0);
            for (int variant = 0; variant <= 3; variant++)
                for (int NB : {8, 4})
                    for (int pad = 0; pad <= ((variant == 1 || variant == 2) && r == 10 ? 1 : 0); pad++) {
                        if (variant == 0 && NB == 8) continue;
                        const Out got = run<VEC>(P, variant, NB, lo, hi, pad != 0, with_z);
                        bool same = true;
                        for (size_t e = 0; e < ref.Y.size(); e++) {
                            const double x = ref.Y[e], y = got.Y[e];
                            if (!((std::isnan(x) && std::isnan(y)) || x == y)) same = false;
                        }
                        check(same, "rows differ from the default kernels (bitwise)", n, r, variant * 10 + NB + 100 * pad);
                        const double u0 = got.sums[0] + got.sums[2] + got.sums[4], u1 = got.sums[1] + got.sums[3] + got.sums[5];
                        check(std::fabs(u0 - t0) <= 1e-10 * (1.0 + std::fabs(t0)) && std::fabs(u1 - t1) <= 1e-10 * (1.0 + std::fabs(t1)),
                              "fused sums differ", n, r, variant * 10 + NB + 100 * pad);
                    }
        }
    }
}
}  // namespace

int main() {
    suite<2>(700, 10, true, 1);    // the C5 shape: r = 10, hub-first order (bundles apply)
    suite<2>(500, 10, false, 2);   // natural order: class 0 is not a contiguous range (bundle falls back)
    suite<2>(400, 8, true, 3);     // 4 lanes per row, 8 rows per bundle
    suite<2>(300, 6, true, 4);     // 3 lanes per row: 10 rows per warp (> kBundleRows: fallback)
    suite<1>(300, 5, true, 5);     // odd rank: scalar pieces
    suite_phases<2>(600, 10, 6);   // hub | tail two-phase pass: EPI 0 with row sub-ranges, then EPI 4 on top
    suite_phases<1>(300, 5, 7);
    std::printf(g_fail ? "emulation: %d FAILED checks\n" : "emulation: all kernel variants agree with the default kernels\n", g_fail);
    return g_fail ? 1 : 0;
}
