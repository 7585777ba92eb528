#!/usr/bin/env python3
"""How cv::gemm rounds small CV_32F products (the `Rcw*x3Dw+tcw` of the reference's projection matchers, ORBmatcher.cc:1364).
Probed on cv2 4.13: for inner length 2..4 equal to the output width or height the products and the running sum are FLOAT
(left to right, addend last); every other shape accumulates in double and rounds once.  oracle/refshim/minicv.cpp
(gemm_eval) follows this table.  Usage: python tools/probe_cv_gemm.py"""
import numpy as np
import cv2

rng = np.random.default_rng(2)


def f32mm(A, B, C=None):
    m, k = A.shape; n = B.shape[1]; out = np.zeros((m, n), np.float32)
    for i in range(m):
        for j in range(n):
            s = np.float32(A[i, 0] * B[0, j])
            for q in range(1, k):
                s = np.float32(s + np.float32(A[i, q] * B[q, j]))
            if C is not None:
                s = np.float32(s + C[i, j])
            out[i, j] = s
    return out


def f64mm(A, B, C=None):
    m, k = A.shape; n = B.shape[1]; out = np.zeros((m, n))
    for i in range(m):
        for j in range(n):
            s = 0.0
            for q in range(k):
                s += float(A[i, q]) * float(B[q, j])
            if C is not None:
                s += float(C[i, j])
            out[i, j] = s
    return out.astype(np.float32)


if __name__ == "__main__":
    for (m, k, n) in [(3, 3, 1), (3, 3, 3), (4, 4, 4), (4, 4, 1), (2, 2, 2), (3, 4, 1), (5, 5, 5), (3, 3, 2), (1, 3, 1), (1, 3, 3), (3, 1, 3), (6, 6, 1)]:
        for useC in (False, True):
            c32 = c64 = 0; N = 1000
            for _ in range(N):
                A = rng.standard_normal((m, k)).astype(np.float32); B = rng.standard_normal((k, n)).astype(np.float32)
                Cm = rng.standard_normal((m, n)).astype(np.float32)
                out = cv2.gemm(A, B, 1.0, Cm if useC else None, 1.0 if useC else 0.0)
                c32 += np.array_equal(out, f32mm(A, B, Cm if useC else None)); c64 += np.array_equal(out, f64mm(A, B, Cm if useC else None))
            small = 2 <= k <= 4 and (k == n or k == m)
            print((m, k, n), "with C" if useC else "no C  ", f"float-path {c32}/{N}  double-path {c64}/{N}  -> model says {'float' if small else 'double'}")
