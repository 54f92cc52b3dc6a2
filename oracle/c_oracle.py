"""ctypes wrapper over oracle/_ref/libgrm_oracle.so (the plain-C restatement, oracle/grm_oracle.c).
TEST INFRASTRUCTURE ONLY — see the header of grm_oracle.c."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libgrm_oracle.so")


def available():
    return os.path.exists(LIB)


def _lib():
    lib = C.CDLL(LIB)
    lib.oracle_decode_stats.restype = C.c_int
    lib.oracle_standardise.restype = C.c_int
    lib.oracle_grm.restype = None
    return lib


def grm(packed, n_fam, keep_mask=None, flat_norm=False):
    lib = _lib()
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    m = packed.shape[0]
    keep = None if keep_mask is None else np.ascontiguousarray(keep_mask, dtype=np.int32)
    n = n_fam if keep is None else int((keep != 0).sum())
    codes = np.zeros((m, n), np.uint8)
    freq = np.zeros((m, 4), np.int32)
    nn = np.zeros(m, np.int32)
    p1, p2, sd = np.zeros(m), np.zeros(m), np.zeros(m)
    vp = C.c_void_p
    got = lib.oracle_decode_stats(vp(packed.ctypes.data), C.c_long(packed.strides[0]), C.c_int(m), C.c_int(n_fam),
                                  vp(keep.ctypes.data) if keep is not None else None, vp(codes.ctypes.data),
                                  vp(freq.ctypes.data), vp(nn.ctypes.data), vp(p1.ctypes.data), vp(p2.ctypes.data),
                                  vp(sd.ctypes.data))
    assert got == n
    z = np.zeros((m, n))
    bad = lib.oracle_standardise(vp(codes.ctypes.data), C.c_int(m), C.c_int(n), vp(p2.ctypes.data),
                                 vp(sd.ctypes.data), vp(z.ctypes.data))
    S = np.full((n, n), np.nan, order="F")
    N = np.full((n, n), np.nan, order="F")
    A = np.full((n, n), np.nan, order="F")
    lib.oracle_grm(vp(codes.ctypes.data), vp(z.ctypes.data), C.c_int(m), C.c_int(n), C.c_int(int(flat_norm)),
                   vp(S.ctypes.data), vp(N.ctypes.data), vp(A.ctypes.data))
    return {"S": S, "N": N, "A": A, "Z": z, "bad": bad,
            "stats": {"freq": freq, "nn": nn, "p1": p1, "p2": p2, "sd": sd}}
