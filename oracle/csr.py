"""ctypes wrapper of oracle/csr_kernels.c (oracle; test infrastructure and the timed
CPU baseline of bench.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcsr_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = C.CDLL(path)
        L.csr_num_threads.restype = C.c_int
        L.csr_set_threads.argtypes = [C.c_int]
        L.csr_spmv.argtypes = [C.c_int64] + [C.c_void_p] * 5
        L.csr_pcg.restype = C.c_int
        L.csr_pcg.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double),
                                                               C.POINTER(C.c_int)]
        L.csr_assemble.restype = C.c_int
        L.csr_assemble.argtypes = [C.c_int64, C.c_int64, C.c_int] + [C.c_void_p] * 7
        _LIB = L
    return _LIB


class CsrOperator:
    def __init__(self, A=None):
        if A is None:
            return
        A = A.tocsr()
        A.sort_indices()
        self.n = A.shape[0]
        self.rowptr = np.ascontiguousarray(A.indptr, dtype=np.int64)
        self.col = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(A.data, dtype=np.float64)
        self.nnz = len(self.val)

    @classmethod
    def from_element_matrices(cls, n, signed_elem_dofs, mat_id, elmat, ess=None):
        """blf->Assemble() + FormLinearSystem elimination in C (csr_assemble): signed_elem_dofs (ne, nd) in MFEM's
        -1-k encoding, mat_id (ne,) into elmat (nmat, nd, nd) local element matrices, ess = essential dof indices."""
        self = cls()
        sd = np.ascontiguousarray(signed_elem_dofs, dtype=np.int32)
        ne, nd = sd.shape
        mid = np.ascontiguousarray(mat_id, dtype=np.int32)
        em = np.ascontiguousarray(elmat, dtype=np.float64)
        assert em.shape[1:] == (nd, nd) and mid.max() < em.shape[0]
        flag = None
        if ess is not None:
            flag = np.zeros(n, np.uint8)
            flag[np.asarray(ess, dtype=np.int64)] = 1
        self.n = int(n)
        self.rowptr = np.zeros(n + 1, np.int64)
        p = self._p
        if lib().csr_assemble(n, ne, nd, p(sd), p(mid), p(em), p(flag), p(self.rowptr), None, None):
            raise MemoryError("csr_assemble (count)")
        self.nnz = int(self.rowptr[-1])
        self.col = np.empty(self.nnz, np.int32)
        self.val = np.empty(self.nnz, np.float64)
        if lib().csr_assemble(n, ne, nd, p(sd), p(mid), p(em), p(flag), p(self.rowptr), p(self.col), p(self.val)):
            raise MemoryError("csr_assemble (fill)")
        return self

    def diagonal(self):
        d = np.zeros(self.n)
        rows = np.repeat(np.arange(self.n), np.diff(self.rowptr))
        m = rows == self.col
        d[rows[m]] = self.val[m]
        return d

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def spmv(self, x, y):
        lib().csr_spmv(self.n, self._p(self.rowptr), self._p(self.col), self._p(self.val), self._p(x), self._p(y))
        return y

    def pcg(self, b, dinv=None, rtol=1e-9, atol=1e-16, max_iter=1000):
        x = np.zeros(self.n)
        rel, conv = C.c_double(0.0), C.c_int(0)
        it = lib().csr_pcg(self.n, self._p(self.rowptr), self._p(self.col), self._p(self.val), self._p(dinv), self._p(b),
                           self._p(x), rtol, atol, max_iter, C.byref(rel), C.byref(conv))
        return x, it, rel.value, bool(conv.value)
