// Host execution of the Lanczos kernels of csrc/lanczos.cu (extracted verbatim into extracted_lanczos.inc) under
// tests/emu/cuda_emu.h.  TEST INFRASTRUCTURE ONLY.  Checks the kernels written without GPU access against the ones the GPU
// parity tests cover: the bundle SpMV against k_lz_spmv (short rows, whole range and owned sub-ranges), the class-range
// search, and the partitioned update (k_lz_update_part on P row blocks + k_lz_beta_finish) against k_lz_update.
#include <algorithm>
#include <cstdio>
#include <random>
#include "cuda_emu.h"

namespace {
constexpr int TPB = 256;
constexpr int LZ_L = 8;
#include "extracted_common.inc"
#include "extracted_lanczos.inc"

int g_fail = 0;
void check(bool ok, const char *what, int a = 0, int b = 0) {
    if (!ok) { std::printf("FAIL %s (%d, %d)\n", what, a, b); g_fail++; }
}
bool close(double x, double y, double tol = 1e-13) { return std::fabs(x - y) <= tol * (1.0 + std::fabs(y)); }

void spmv_suite(int n, unsigned seed) {
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    // hub-first order: long rows first, the short rows (<= 32 nonzeros) are a contiguous tail
    std::vector<int> len(n);
    for (int i = 0; i < n; i++) len[i] = (g() % 10 == 0) ? 33 + (int)(g() % 100) : (int)(g() % 33);
    std::sort(len.begin(), len.end(), std::greater<int>());
    std::vector<int> ptr(n + 1, 0);
    for (int i = 0; i < n; i++) ptr[i + 1] = ptr[i] + len[i];
    const int nnz = ptr[n];
    std::vector<int> idx(nnz);
    std::vector<double> S(nnz), v(n);
    for (int k = 0; k < nnz; k++) { idx[k] = (int)(g() % n); S[k] = U(g); }
    for (double &x : v) x = U(g);
    std::vector<int> cls0;
    for (int i = 0; i < n; i++) if (len[i] <= 32) cls0.push_back(i);
    const int c0 = cls0.front(), n0 = (int)cls0.size();
    check(cls0.back() - c0 + 1 == n0, "test pattern: class 0 not contiguous");
    std::vector<double> partials(1 << 16, 0.0);
    unsigned ticket[4] = {0, 0, 0, 0};
    double stop = 0.0;
    const int ranges[3][2] = {{0, n}, {n / 3, 2 * n / 3}, {c0 + 3, c0 + 4}};
    for (auto &rg : ranges) {
        const int lo = std::max(rg[0], c0), hi = std::max(lo, std::min(rg[1], c0 + n0));
        std::vector<double> w_ref(n, std::nan("")), w_got(n, std::nan(""));
        double a_ref = 0.0, a_got = 0.0;
        // reference: k_lz_spmv over the owned part of the class list (list pointer offset, as lz_run_dist does)
        const int q_lo = lo - c0, nr = hi - lo;
        if (nr > 0) emu::launch(k_lz_spmv<4, true>, 2, TPB, (const int *)(cls0.data() + q_lo), (i64)nr, (const int *)ptr.data(), (const int *)idx.data(),
                                (const double *)S.data(), (const double *)v.data(), w_ref.data(), (const double *)&stop, partials.data(), ticket, &a_ref, (i64)0);
        emu::launch(k_lz_spmv_bundle<true>, 2, TPB, (i64)lo, (i64)hi, (const int *)ptr.data(), (const int *)idx.data(), (const double *)S.data(),
                    (const double *)v.data(), w_got.data(), (const double *)&stop, partials.data(), ticket, &a_got);
        for (int i = 0; i < n; i++) {
            if (i < lo || i >= hi) { check(std::isnan(w_got[i]), "bundle SpMV wrote a row outside its range", i); continue; }
            double acc = 0.0;
            for (int k = ptr[i]; k < ptr[i + 1]; k++) acc += S[k] * v[idx[k]];
            check(close(w_ref[i], acc, 1e-12), "k_lz_spmv vs plain product", i);
            check(close(w_got[i], w_ref[i]), "bundle SpMV vs k_lz_spmv", i);
        }
        check(close(a_got, a_ref, 1e-12), "bundle SpMV alpha share");
        // identity list with a row offset (k_lz_spmv's row_off, used when every row is short)
        if (nr > 0) {
            std::vector<double> w_off(n, std::nan(""));
            double a_off = 0.0;
            emu::launch(k_lz_spmv<4, true>, 1, TPB, (const int *)nullptr, (i64)nr, (const int *)ptr.data(), (const int *)idx.data(), (const double *)S.data(),
                        (const double *)v.data(), w_off.data(), (const double *)&stop, partials.data(), ticket, &a_off, (i64)lo);
            bool same = true;
            for (int i = lo; i < hi; i++) same = same && w_off[i] == w_ref[i];
            check(same && close(a_off, a_ref, 1e-12), "k_lz_spmv with row_off");
        }
    }
    // class-range search against a direct count
    std::vector<int> l1, l2;
    for (int i = 0; i < n; i++) if (len[i] > 32) (len[i] > 80 ? l2 : l1).push_back(i);
    for (auto &rg : ranges) {
        long long out[6] = {-1, -1, -1, -1, -1, -1};
        emu::launch(k_lz_class_ranges, 1, 32, (const int *)cls0.data(), (i64)cls0.size(), (const int *)l1.data(), (i64)l1.size(), (const int *)l2.data(),
                    (i64)l2.size(), (i64)rg[0], (i64)rg[1], out);
        const std::vector<int> *L[3] = {&cls0, &l1, &l2};
        for (int c = 0; c < 3; c++) {
            long long lo = 0, hi = 0;
            for (int x : *L[c]) { if (x < rg[0]) lo++; if (x < rg[1]) hi++; }
            check(out[2 * c] == lo && out[2 * c + 1] == hi, "k_lz_class_ranges", c);
        }
        long long idn[6];
        emu::launch(k_lz_class_ranges, 1, 32, (const int *)nullptr, (i64)n, (const int *)nullptr, (i64)0, (const int *)nullptr, (i64)0, (i64)rg[0], (i64)rg[1], idn);
        check(idn[0] == rg[0] && idn[1] == rg[1] && idn[2] == 0 && idn[3] == 0, "k_lz_class_ranges identity list");
    }
    // partitioned update: P row blocks + finish == k_lz_update
    const i64 q = 5;
    for (int step : {0, 2}) {
        std::vector<double> ab(2 * q, 0.0), vp(n), w0(n);
        for (double &x : vp) x = U(g);
        for (double &x : w0) x = U(g);
        ab[step] = 0.37; if (step > 0) ab[q + step - 1] = 1.21;
        std::vector<double> ab_ref = ab, w_ref = w0;
        double stop_ref = 0.0;
        emu::launch(k_lz_update, 3, TPB, (i64)n, step, (const double *)v.data(), (const double *)vp.data(), w_ref.data(), ab_ref.data(), q, &stop_ref, partials.data(), ticket);
        std::vector<double> w_got = w0, ab_got = ab;
        double total = 0.0, stop_got = 0.0;
        const int P = 3;
        for (int rk = 0; rk < P; rk++) {
            const i64 lo = (i64)n * rk / P, hi = (i64)n * (rk + 1) / P;
            double part = 0.0;
            emu::launch(k_lz_update_part, 2, TPB, hi - lo, step, (const double *)v.data() + lo, (const double *)vp.data() + lo, w_got.data() + lo,
                        (const double *)ab_got.data(), q, (const double *)&stop_got, partials.data(), ticket, &part);
            total += part;   // the all-reduce
        }
        emu::launch(k_lz_beta_finish, 1, 1, (i64)n, step, (const double *)&total, ab_got.data(), q, &stop_got);
        bool same = true;
        for (int i = 0; i < n; i++) same = same && w_got[i] == w_ref[i];
        check(same, "partitioned update: w", step);
        check(close(ab_got[q + step], ab_ref[q + step], 1e-13) && stop_got == stop_ref, "partitioned update: beta", step);
    }
}
}  // namespace

int main() {
    spmv_suite(900, 1);
    spmv_suite(257, 2);
    spmv_suite(64, 3);
    std::printf(g_fail ? "emulation: %d FAILED checks\n" : "emulation: the Lanczos kernels agree\n", g_fail);
    return g_fail ? 1 : 0;
}
