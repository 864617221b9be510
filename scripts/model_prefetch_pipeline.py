"""Host model of the control flow of the software-pipelined gather kernels (gradient.cu: k_rows_group_pf, k_rows_warp_pf,
owned_q_range).  The kernels were written without GPU access; this transliteration of their stage / rotation logic is run
against a plain CSR product on random patterns (empty rows, rows longer than one block, row lists, owned sub-ranges, more
groups than rows) to catch indexing mistakes before GPU time is spent.  `python scripts/model_prefetch_pipeline.py`."""
import numpy as np


def owned_q_range(lst, n_list, own_lo, own_hi):
    if lst is None:
        q_lo = 0 if own_lo < 0 else min(own_lo, n_list)
        q_hi = (max(own_hi, q_lo) if own_hi < n_list else n_list)
        return q_lo, q_hi
    lo, hi = 0, n_list
    if own_lo > 0:
        while lo < hi:
            mid = lo + ((hi - lo) >> 1)
            if lst[mid] < own_lo: lo = mid + 1
            else: hi = mid
    q_lo = lo
    hi = n_list
    if n_list > 0 and lst[n_list - 1] >= own_hi:
        while lo < hi:
            mid = lo + ((hi - lo) >> 1)
            if lst[mid] < own_hi: lo = mid + 1
            else: hi = mid
        hi = lo
    return q_lo, hi


def group_kernel(ptr, idx, val, X, rows, n_rows, own_lo, own_hi, n_groups, NB, Y, loads):
    """every lane group of k_rows_group_pf, one after the other"""
    q_lo, q_hi = owned_q_range(rows, n_rows, own_lo, own_hi)
    row_at = lambda q: ((rows[q] if rows is not None else q) if q < q_hi else -2)
    s0 = 0.0
    for group in range(n_groups):
        q = q_lo + group
        iC = row_at(q); k0 = endC = 0
        if iC >= 0: k0, endC = ptr[iC], ptr[iC + 1]
        q += n_groups
        iB = row_at(q); begB = endB = 0
        if iB >= 0: begB, endB = ptr[iB], ptr[iB + 1]
        q += n_groups
        iA = row_at(q)
        cc = [idx[k0 + j] if k0 + j < endC else 0 for j in range(NB)]
        xC = X[iC].copy() if iC >= 0 else None
        acc = np.zeros(X.shape[1])
        while iC >= 0:
            g = [X[cc[j]] if k0 + j < endC else None for j in range(NB)]
            vv = [val[k0 + j] if k0 + j < endC else 0.0 for j in range(NB)]
            loads[0] += sum(1 for j in range(NB) if k0 + j < endC)
            last = k0 + NB >= endC
            nk0 = begB if last else k0 + NB
            nend = endB if last else endC
            cn = [idx[nk0 + j] if nk0 + j < nend else 0 for j in range(NB)]
            iN, begA, endA, xB = -2, 0, 0, None
            if last:
                if iA >= 0: begA, endA = ptr[iA], ptr[iA + 1]
                q += n_groups
                iN = row_at(q)
                if iB >= 0: xB = X[iB].copy()
            for j in range(NB):
                if k0 + j < endC: acc += vv[j] * g[j]
            if last:
                s0 += float(acc @ xC)
                assert not np.any(np.isfinite(Y[iC])), "row written twice"
                Y[iC] = acc
                acc = np.zeros(X.shape[1])
                iC, k0, endC, xC = iB, begB, endB, xB
                iB, begB, endB = iA, begA, endA
                iA = iN
            else:
                k0 += NB
            cc = cn
    return s0


def warp_kernel(ptr, idx, val, X, rows, n_rows, own_lo, own_hi, n_warps, ng, Y, chunk=None):
    """every warp of k_rows_warp_pf; chunk = (chunk_row, chunk_start, chunk_end) for the CHUNK variant (Y = scratch per chunk)"""
    CH = chunk is not None
    lst = chunk[0] if CH else rows
    q_lo, q_hi = owned_q_range(lst, n_rows, own_lo, own_hi)
    item_row = lambda q: ((chunk[0][q] if CH else (rows[q] if rows is not None else q)) if q < q_hi else -2)
    item_beg = lambda q, i: 0 if i < 0 else (chunk[1][q] if CH else ptr[i])
    item_end = lambda q, i: 0 if i < 0 else (chunk[2][q] if CH else ptr[i + 1])
    step = ng * 4
    s0 = 0.0
    for warp in range(n_warps):
        q = q_lo + warp
        qC = q; iC = item_row(q); kb, endC = item_beg(q, iC), item_end(q, iC)
        q += n_warps
        qB = q; iB = item_row(q); begB, endB = item_beg(q, iB), item_end(q, iB)
        q += n_warps
        qA = q; iA = item_row(q)
        cc = [[idx[kb + grp * 4 + j] if kb + grp * 4 + j < endC else 0 for j in range(4)] for grp in range(ng)]
        acc = [np.zeros(X.shape[1]) for _ in range(ng)]
        while iC >= 0:
            for grp in range(ng):
                k0 = kb + grp * 4
                for j in range(4):
                    if k0 + j < endC: acc[grp] += val[k0 + j] * X[cc[grp][j]]
            last = kb + step >= endC
            nkb = begB if last else kb + step
            nend = endB if last else endC
            cn = [[idx[nkb + grp * 4 + j] if nkb + grp * 4 + j < nend else 0 for j in range(4)] for grp in range(ng)]
            iN, qN, begA, endA = -2, q, 0, 0
            if last:
                begA, endA = item_beg(qA, iA), item_end(qA, iA)
                q += n_warps
                qN = q
                iN = item_row(q)
                tot = sum(acc)
                if CH:
                    assert not np.any(np.isfinite(Y[qC]))
                    Y[qC] = tot
                else:
                    assert not np.any(np.isfinite(Y[iC]))
                    Y[iC] = tot
                    s0 += float(tot @ X[iC])
                acc = [np.zeros(X.shape[1]) for _ in range(ng)]
                qC, iC, kb, endC = qB, iB, begB, endB
                qB, iB, begB, endB = qA, iA, begA, endA
                qA, iA = qN, iN
            else:
                kb += step
            cc = cn
    return s0


def bundle_kernel(ptr, idx, val, X, c0_first, n_rows, own_lo, own_hi, n_warps, G, Y):
    """every warp of k_rows_bundle: lane l holds ptr[i0 + min(l, nrow)], the span is staged, group g walks [o, e)"""
    RPW = 32 // G
    r_lo, r_hi = max(c0_first, own_lo), min(c0_first + n_rows, own_hi)
    r_hi = max(r_hi, r_lo)
    n_bundles = (r_hi - r_lo + RPW - 1) // RPW
    s0 = 0.0

    def load_ptr(bb):
        if bb >= n_bundles: return [0] * 32
        i0 = r_lo + bb * RPW
        nrow = min(RPW, r_hi - i0)
        return [int(ptr[i0 + (l if l < nrow else nrow)]) for l in range(32)]
    for warp in range(n_warps):
        b = warp
        myp = load_ptr(b)
        while b < n_bundles:
            i0 = r_lo + b * RPW
            nrow = min(RPW, r_hi - i0)
            p0, pR = myp[0], myp[RPW]
            assert pR - p0 <= RPW * 32
            si = [None] * (RPW * 32); sv = [None] * (RPW * 32)
            for lane in range(32):
                for m in range(8):
                    k = p0 + lane + 32 * m
                    if k < pR: si[lane + 32 * m] = idx[k]; sv[lane + 32 * m] = val[k]
            mypN = load_ptr(b + n_warps)
            for g in range(32 // G + (1 if 32 % G else 0)):   # lane groups incl. the partial one that only stages
                grp_ok = g < RPW
                o = myp[g if grp_ok else RPW] - p0
                e = myp[g + 1 if grp_ok else RPW] - p0
                if not (grp_ok and g < nrow): continue
                i = i0 + g
                acc = np.zeros(X.shape[1])
                for k in range(o, e): acc += sv[k] * X[si[k]]
                s0 += float(acc @ X[i])
                assert not np.any(np.isfinite(Y[i])), "row written twice"
                Y[i] = acc
            myp = mypN
            b += n_warps
    return s0


def main():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(300):
        n = int(rng.integers(1, 60))
        r = int(rng.integers(1, 4))
        lens = rng.integers(0, 40, n)
        if trial % 7 == 0: lens[:] = 0
        ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        nnz = int(ptr[-1])
        idx = rng.integers(0, n, nnz)
        val = rng.standard_normal(nnz)
        X = rng.standard_normal((n, r))
        ref = np.zeros((n, r))
        for i in range(n):
            for k in range(ptr[i], ptr[i + 1]): ref[i] += val[k] * X[idx[k]]
        # a row list (a "class") or the identity; an owned range
        use_list = trial % 2 == 0
        rows = np.sort(rng.choice(n, int(rng.integers(0, n + 1)), replace=False)) if use_list else None
        n_rows = len(rows) if use_list else n
        own_lo = int(rng.integers(0, n + 1)) if trial % 3 == 0 else 0
        own_hi = int(rng.integers(own_lo, n + 1)) if trial % 3 == 0 else n
        want = [i for i in (rows if use_list else range(n)) if own_lo <= i < own_hi]
        for NB in (4, 8):
            for n_groups in (1, 3, 7, 100):
                Y = np.full((n, r), np.nan); loads = [0]
                s0 = group_kernel(ptr, idx, val, X, rows, n_rows, own_lo, own_hi, n_groups, NB, Y, loads)
                done = np.where(np.isfinite(Y).all(axis=1))[0].tolist()
                assert done == sorted(want), (trial, done, want)
                assert np.allclose(Y[want], ref[want], rtol=1e-12, atol=1e-12)
                assert np.isclose(s0, float(np.sum(ref[want] * X[want])), rtol=1e-10, atol=1e-10)
                assert loads[0] == sum(int(ptr[i + 1] - ptr[i]) for i in want)
                cases += 1
        for ng in (1, 4, 6):
            for n_warps in (1, 2, 5, 64):
                Y = np.full((n, r), np.nan)
                s0 = warp_kernel(ptr, idx, val, X, rows, n_rows, own_lo, own_hi, n_warps, ng, Y)
                done = np.where(np.isfinite(Y).all(axis=1))[0].tolist()
                assert done == sorted(want), (trial, done, want)
                assert np.allclose(Y[want], ref[want], rtol=1e-12, atol=1e-12)
                assert np.isclose(s0, float(np.sum(ref[want] * X[want])), rtol=1e-10, atol=1e-10)
                cases += 1
        # chunks of the long rows: rows with more than `cut` nonzeros, cut into pieces of `cut`
        cut = 8
        crow, cbeg, cend = [], [], []
        for i in range(n):
            if ptr[i + 1] - ptr[i] > cut:
                for b in range(ptr[i], ptr[i + 1], cut):
                    crow.append(i); cbeg.append(b); cend.append(min(b + cut, ptr[i + 1]))
        if crow:
            chunk = (np.array(crow), np.array(cbeg), np.array(cend))
            Ys = np.full((len(crow), r), np.nan)
            warp_kernel(ptr, idx, val, X, None, len(crow), own_lo, own_hi, 3, 4, Ys, chunk=chunk)
            for c, (i, b, e) in enumerate(zip(crow, cbeg, cend)):
                if own_lo <= i < own_hi:
                    exp = sum(val[k] * X[idx[k]] for k in range(b, e))
                    assert np.allclose(Ys[c], exp, rtol=1e-12, atol=1e-12)
                else:
                    assert not np.any(np.isfinite(Ys[c]))
            cases += 1
    # bundles: class 0 = a contiguous row range with rows of at most 32 nonzeros
    for trial in range(200):
        n = int(rng.integers(1, 80))
        r = int(rng.integers(1, 4))
        lens = rng.integers(0, 33, n)
        if trial % 5 == 0: lens[:] = 32
        ptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        nnz = int(ptr[-1])
        idx = rng.integers(0, n, nnz); val = rng.standard_normal(nnz); X = rng.standard_normal((n, r))
        ref = np.zeros((n, r))
        for i in range(n):
            for k in range(ptr[i], ptr[i + 1]): ref[i] += val[k] * X[idx[k]]
        c0_first = int(rng.integers(0, n))
        n_rows = int(rng.integers(0, n - c0_first + 1))
        own_lo = int(rng.integers(0, n + 1)) if trial % 3 == 0 else 0
        own_hi = int(rng.integers(own_lo, n + 1)) if trial % 3 == 0 else n
        want = [i for i in range(c0_first, c0_first + n_rows) if own_lo <= i < own_hi]
        for G in (4, 5, 6, 8, 10, 16):
            for n_warps in (1, 3, 50):
                Y = np.full((n, r), np.nan)
                s0 = bundle_kernel(ptr, idx, val, X, c0_first, n_rows, own_lo, own_hi, n_warps, G, Y)
                done = np.where(np.isfinite(Y).all(axis=1))[0].tolist()
                assert done == want, (trial, G, done, want)
                assert np.allclose(Y[want], ref[want], rtol=1e-12, atol=1e-12)
                assert np.isclose(s0, float(np.sum(ref[want] * X[want])), rtol=1e-10, atol=1e-10)
                cases += 1
    print("pipeline model: %d cases agree with the plain CSR product" % cases)


if __name__ == "__main__":
    main()
