"""Extracts the Gset G1 adjacency (n=800, 19,176 unit edges) from the reference's
MATLAB v7.3 fixture exps/data/MaxCut/G1.mat into tests/golden/g1_graph.npz.

The .mat file is HDF5 with uncompressed contiguous datasets, read at fixed
offsets because h5py is not installed.  Run in the build
container only (the GPU box has no /root/reference):
    python tests/golden/make_g1_fixture.py
"""
import os
import numpy as np

SRC = "/root/reference/exps/data/MaxCut/G1.mat"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "g1_graph.npz")

raw = open(SRC, "rb").read()
assert raw[:19] == b"MATLAB 7.3 MAT-file"
data = np.frombuffer(raw, dtype="<f8", count=38352, offset=2576)
ir = np.frombuffer(raw, dtype="<u8", count=38352, offset=311440)
jc = np.frombuffer(raw, dtype="<u8", count=801, offset=618256)
assert jc[0] == 0 and jc[-1] == 38352 and np.all(data == 1.0) and ir.max() == 799
cols = np.repeat(np.arange(800), np.diff(jc.astype(np.int64)))
rows = ir.astype(np.int64)
A = np.zeros((800, 800), dtype=bool)
A[rows, cols] = True
assert (A == A.T).all() and not A.diagonal().any() and A.sum() == 2 * 19176
np.savez_compressed(DST, indices=ir.astype(np.uint16), indptr=jc.astype(np.uint32), n=np.int64(800))
print("wrote", DST, os.path.getsize(DST), "bytes")
