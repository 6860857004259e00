"""oracle/pyoracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loaders for the two CPU checkers plus a NumPy closed-form third opinion:

* ``COracle``   -- oracle/libccn_oracle.so, the plain-C restatement (oracle/ccn_oracle.c).
* ``RefOracle`` -- oracle/_ref/libgfref_{f64,f32}.so, the UNMODIFIED reference headers compiled behind
                   oracle/ref_shim.cpp (present whenever /root/reference was available at build time).
* ``einsum18``  -- the 18 contractions as numpy einsums (the starred rows), fp64;
                   fast enough for full-size (N=32, C=64) instances.  Checked against both libraries in
                   tests/test_oracle_cpu.py before it is trusted anywhere else.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = os.path.join(HERE, "libccn_oracle.so")
_REF_LIB = {"f64": os.path.join(HERE, "_ref", "libgfref_f64.so"), "f32": os.path.join(HERE, "_ref", "libgfref_f32.so"),
            "f32_o3": os.path.join(HERE, "_ref", "libgfref_f32_o3.so")}

_DT = {"f64": (np.float64, ctypes.c_double), "f32": (np.float32, ctypes.c_float)}


def build(ref=True):
    """Compile the checkers (idempotent).  `make ref` is a no-op when /root/reference is absent."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def ref_available(prec="f64"):
    return os.path.exists(_REF_LIB[prec])


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


class COracle:
    """Plain-C restatement.  prec: 'f64' | 'f32'."""

    def __init__(self, prec="f64"):
        if not os.path.exists(_C_LIB):
            build(ref=False)
        self.lib = ctypes.CDLL(_C_LIB)
        self.prec = prec
        self.np_t, self.c_t = _DT[prec]

    def _fn(self, name):
        f = getattr(self.lib, "%s_%s" % (name, self.prec))
        f.restype = None
        return f

    def pattern_ref_lines(self):
        f = getattr(self.lib, "ccn_oracle_pattern_ref_line_%s" % self.prec)
        f.restype = ctypes.c_int
        return [f(k) for k in range(18)]

    def contract18_forward(self, T, adj, positive_part=True):
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        out = np.empty((N, N, 18 * C), self.np_t)
        self._fn("ccn_oracle_contract18_forward")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(out, self.c_t),
                                                  ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(int(positive_part)))
        return out

    def contract18_backward(self, gout, adj, gT_init=None, positive_part=True):
        N = gout.shape[0]
        C = gout.shape[2] // 18
        gout = np.ascontiguousarray(gout, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gT = np.zeros((N, N, N, C), self.np_t) if gT_init is None else np.array(gT_init, self.np_t, order="C")
        self._fn("ccn_oracle_contract18_backward")(_ptr(gout, self.c_t), _ptr(adj, self.c_t), _ptr(gT, self.c_t),
                                                   ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(int(positive_part)))
        return gT

    def contract50_forward(self, T, adj):
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        out = np.empty((N, N, 50 * C), self.np_t)
        self._fn("ccn_oracle_contract50_forward")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(out, self.c_t),
                                                  ctypes.c_int(N), ctypes.c_int(C))
        return out

    def contract50_backward(self, gout, adj, gT_init=None):
        N = gout.shape[0]
        C = gout.shape[2] // 50
        gout = np.ascontiguousarray(gout, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gT = np.zeros((N, N, N, C), self.np_t) if gT_init is None else np.array(gT_init, self.np_t, order="C")
        self._fn("ccn_oracle_contract50_backward")(_ptr(gout, self.c_t), _ptr(adj, self.c_t), _ptr(gT, self.c_t),
                                                   ctypes.c_int(N), ctypes.c_int(C))
        return gT

    def tensor_mul_forward(self, A, B):
        R, K, D = A.shape
        Cc = B.shape[1]
        A, B = np.ascontiguousarray(A, self.np_t), np.ascontiguousarray(B, self.np_t)
        out = np.empty((R, Cc, D), self.np_t)
        self._fn("ccn_oracle_tensor_mul_forward")(_ptr(A, self.c_t), _ptr(B, self.c_t), _ptr(out, self.c_t), ctypes.c_int(R),
                                                  ctypes.c_int(K), ctypes.c_int(Cc), ctypes.c_int(D))
        return out

    def tensor_mul_backward(self, A, B, g, gA_init=None, gB_init=None):
        R, K, D = A.shape
        Cc = B.shape[1]
        A, B, g = (np.ascontiguousarray(x, self.np_t) for x in (A, B, g))
        gA = np.zeros_like(A) if gA_init is None else np.array(gA_init, self.np_t, order="C")
        gB = np.zeros_like(B) if gB_init is None else np.array(gB_init, self.np_t, order="C")
        self._fn("ccn_oracle_tensor_mul_backward")(_ptr(A, self.c_t), _ptr(B, self.c_t), _ptr(g, self.c_t), _ptr(gA, self.c_t),
                                                   _ptr(gB, self.c_t), ctypes.c_int(R), ctypes.c_int(K), ctypes.c_int(Cc),
                                                   ctypes.c_int(D))
        return gA, gB

    def custom_matmul_tensor_forward(self, Kt, X):
        P, V = Kt.shape
        X2 = np.ascontiguousarray(X, self.np_t).reshape(-1, V)
        Kt = np.ascontiguousarray(Kt, self.np_t)
        Y = np.empty((X2.shape[0], P), self.np_t)
        self._fn("ccn_oracle_custom_matmul_tensor_forward")(_ptr(Kt, self.c_t), _ptr(X2, self.c_t), _ptr(Y, self.c_t),
                                                            ctypes.c_int64(X2.shape[0]), ctypes.c_int(V), ctypes.c_int(P))
        return Y.reshape(X.shape[:-1] + (P,))

    def custom_matmul_tensor_backward(self, Kt, X, gY, gKt_init=None, gX_init=None):
        P, V = Kt.shape
        X2 = np.ascontiguousarray(X, self.np_t).reshape(-1, V)
        Kt = np.ascontiguousarray(Kt, self.np_t)
        g2 = np.ascontiguousarray(gY, self.np_t).reshape(-1, P)
        gKt = np.zeros_like(Kt) if gKt_init is None else np.array(gKt_init, self.np_t, order="C")
        gX = np.zeros_like(X2) if gX_init is None else np.array(gX_init, self.np_t, order="C").reshape(-1, V)
        self._fn("ccn_oracle_custom_matmul_tensor_backward")(_ptr(Kt, self.c_t), _ptr(X2, self.c_t), _ptr(g2, self.c_t),
                                                             _ptr(gKt, self.c_t), _ptr(gX, self.c_t),
                                                             ctypes.c_int64(X2.shape[0]), ctypes.c_int(V), ctypes.c_int(P))
        return gKt, gX.reshape(X.shape)

    def promote_forward(self, f, pos):
        m, C = f.shape[0], f.shape[2]
        n = len(pos)
        f = np.ascontiguousarray(f, self.np_t)
        pos = np.ascontiguousarray(pos, np.int32)
        Q = np.empty((n, n, C), self.np_t)
        self._fn("ccn_oracle_promote")(_ptr(f, self.c_t), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _ptr(Q, self.c_t),
                                       ctypes.c_int(n), ctypes.c_int(m), ctypes.c_int(C), ctypes.c_int(0))
        return Q

    def promote_backward(self, gQ, pos, m, gf_init=None):
        n, C = gQ.shape[0], gQ.shape[2]
        gQ = np.ascontiguousarray(gQ, self.np_t)
        pos = np.ascontiguousarray(pos, np.int32)
        gf = np.zeros((m, m, C), self.np_t) if gf_init is None else np.array(gf_init, self.np_t, order="C")
        self._fn("ccn_oracle_promote")(_ptr(gf, self.c_t), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _ptr(gQ, self.c_t),
                                       ctypes.c_int(n), ctypes.c_int(m), ctypes.c_int(C), ctypes.c_int(1))
        return gf

    def matmul_forward(self, X, W):
        M, K = X.shape
        P = W.shape[1]
        X = np.ascontiguousarray(X, self.np_t)
        W = np.ascontiguousarray(W, self.np_t)
        Y = np.empty((M, P), self.np_t)
        self._fn("ccn_oracle_matmul_forward")(_ptr(X, self.c_t), _ptr(W, self.c_t), _ptr(Y, self.c_t),
                                              ctypes.c_int(M), ctypes.c_int(K), ctypes.c_int(P))
        return Y

    def matmul_backward(self, X, W, gY, gX_init=None, gW_init=None):
        M, K = X.shape
        P = W.shape[1]
        X = np.ascontiguousarray(X, self.np_t)
        W = np.ascontiguousarray(W, self.np_t)
        gY = np.ascontiguousarray(gY, self.np_t)
        gX = np.zeros((M, K), self.np_t) if gX_init is None else np.array(gX_init, self.np_t, order="C")
        gW = np.zeros((K, P), self.np_t) if gW_init is None else np.array(gW_init, self.np_t, order="C")
        self._fn("ccn_oracle_matmul_backward")(_ptr(X, self.c_t), _ptr(W, self.c_t), _ptr(gY, self.c_t),
                                               _ptr(gX, self.c_t), _ptr(gW, self.c_t), ctypes.c_int(M),
                                               ctypes.c_int(K), ctypes.c_int(P))
        return gX, gW

    def bias_lrelu_forward(self, Y, bias, alpha=0.01):
        Y = np.ascontiguousarray(Y, self.np_t)
        bias = np.ascontiguousarray(bias, self.np_t)
        Z = np.empty_like(Y)
        P = Y.shape[-1]
        self._fn("ccn_oracle_bias_lrelu_forward")(_ptr(Y, self.c_t), _ptr(bias, self.c_t), _ptr(Z, self.c_t),
                                                  ctypes.c_int64(Y.size // P), ctypes.c_int(P), self.c_t(alpha))
        return Z

    def bias_lrelu_backward(self, Y, bias, gZ, alpha=0.01):
        Y = np.ascontiguousarray(Y, self.np_t)
        bias = np.ascontiguousarray(bias, self.np_t)
        gZ = np.ascontiguousarray(gZ, self.np_t)
        gY = np.zeros_like(Y)
        P = Y.shape[-1]
        gb = np.zeros((P,), self.np_t)
        self._fn("ccn_oracle_bias_lrelu_backward")(_ptr(Y, self.c_t), _ptr(bias, self.c_t), _ptr(gZ, self.c_t),
                                                   _ptr(gY, self.c_t), _ptr(gb, self.c_t),
                                                   ctypes.c_int64(Y.size // P), ctypes.c_int(P), self.c_t(alpha))
        return gY, gb


class RefOracle:
    """The unmodified reference (serial RisiContraction_18 etc.) compiled behind oracle/ref_shim.cpp."""

    def __init__(self, prec="f64"):
        """prec: 'f64' | 'f32' | 'f32_o3' (the f32 tree built with -O3 -march=x86-64-v3, for timing only)."""
        if not ref_available(prec):
            raise FileNotFoundError("oracle/_ref not built (needs /root/reference at build time)")
        self.lib = ctypes.CDLL(_REF_LIB[prec])
        self.prec = prec = prec.split("_")[0]
        self.np_t, self.c_t = _DT[prec]

    def _fn(self, name, restype=None):
        f = getattr(self.lib, "%s_%s" % (name, self.prec))
        f.restype = restype
        return f

    def contract18_forward(self, T, adj, variant="serial"):
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        out = np.empty((N, N, 18 * C), self.np_t)
        name = {"serial": "gfref_contract18_forward", "definition": "gfref_contract18_forward_definition",
                "thread": "gfref_contract18_thread_forward"}[variant]
        self._fn(name)(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(out, self.c_t), ctypes.c_int(N), ctypes.c_int(C))
        return out

    def contract18_backward(self, gout, adj, gT_init=None):
        N = gout.shape[0]
        C = gout.shape[2] // 18
        gout = np.ascontiguousarray(gout, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gT = np.zeros((N, N, N, C), self.np_t) if gT_init is None else np.array(gT_init, self.np_t, order="C")
        self._fn("gfref_contract18_backward")(_ptr(gout, self.c_t), _ptr(adj, self.c_t), _ptr(gT, self.c_t),
                                              ctypes.c_int(N), ctypes.c_int(C))
        return gT

    def contract50_forward(self, T, adj):
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        out = np.empty((N, N, 50 * C), self.np_t)
        self._fn("gfref_contract50_forward")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(out, self.c_t),
                                             ctypes.c_int(N), ctypes.c_int(C))
        return out

    def contract50_backward(self, gout, adj, gT_init=None):
        N = gout.shape[0]
        C = gout.shape[2] // 50
        gout = np.ascontiguousarray(gout, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gT = np.zeros((N, N, N, C), self.np_t) if gT_init is None else np.array(gT_init, self.np_t, order="C")
        self._fn("gfref_contract50_backward")(_ptr(gout, self.c_t), _ptr(adj, self.c_t), _ptr(gT, self.c_t),
                                              ctypes.c_int(N), ctypes.c_int(C))
        return gT

    def contract4(self, T, gout):
        """RisiContraction_4 forward + backward into a fresh gradient -> (out [N,N,4C], gT [N,N,N,C])."""
        N, C = T.shape[0], T.shape[3]
        T, gout = np.ascontiguousarray(T, self.np_t), np.ascontiguousarray(gout, self.np_t)
        out, gT = np.empty((N, N, 4 * C), self.np_t), np.zeros((N, N, N, C), self.np_t)
        self._fn("gfref_contract4")(_ptr(T, self.c_t), _ptr(gout, self.c_t), _ptr(out, self.c_t), _ptr(gT, self.c_t),
                                    ctypes.c_int(N), ctypes.c_int(C))
        return out, gT

    def contract10(self, T, adj, gout):
        """RisiContraction_10 forward + backward into a fresh gradient -> (out [N,N,10C], gT)."""
        N, C = T.shape[0], T.shape[3]
        T, adj, gout = (np.ascontiguousarray(x, self.np_t) for x in (T, adj, gout))
        out, gT = np.empty((N, N, 10 * C), self.np_t), np.zeros((N, N, N, C), self.np_t)
        self._fn("gfref_contract10")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(gout, self.c_t), _ptr(out, self.c_t),
                                     _ptr(gT, self.c_t), ctypes.c_int(N), ctypes.c_int(C))
        return out, gT

    def contract18_dropout(self, T, adj, gout, n_kept, seed, train=True):
        """RisiContraction_18_dropout: srand(seed); forward() draws use[] with rand(); backward -> (out, gT, use[18])."""
        N, C = T.shape[0], T.shape[3]
        T, adj, gout = (np.ascontiguousarray(x, self.np_t) for x in (T, adj, gout))
        out, gT = np.empty((N, N, 18 * C), self.np_t), np.zeros((N, N, N, C), self.np_t)
        use = (ctypes.c_int * 18)()
        self._fn("gfref_contract18_dropout")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(gout, self.c_t), _ptr(out, self.c_t),
                                             _ptr(gT, self.c_t), ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(int(n_kept)),
                                             ctypes.c_uint(seed), ctypes.c_int(int(bool(train))), use)
        return out, gT, [bool(u) for u in use]

    def tensor_mul(self, A, B, gout=None, gA_init=None, gB_init=None):
        """TensorMul forward (+ backward when gout is given): returns out or (out, gA, gB)."""
        R, K, D = A.shape
        Cc = B.shape[1]
        A, B = np.ascontiguousarray(A, self.np_t), np.ascontiguousarray(B, self.np_t)
        out = np.empty((R, Cc, D), self.np_t)
        gA = np.zeros_like(A) if gA_init is None else np.array(gA_init, self.np_t, order="C")
        gB = np.zeros_like(B) if gB_init is None else np.array(gB_init, self.np_t, order="C")
        g = None if gout is None else np.ascontiguousarray(gout, self.np_t)
        self._fn("gfref_tensor_mul")(_ptr(A, self.c_t), _ptr(B, self.c_t), _ptr(out, self.c_t),
                                     None if g is None else _ptr(g, self.c_t), _ptr(gA, self.c_t), _ptr(gB, self.c_t),
                                     ctypes.c_int(R), ctypes.c_int(K), ctypes.c_int(Cc), ctypes.c_int(D))
        return out if gout is None else (out, gA, gB)

    def custom_matmul_tensor(self, Kt, X, gY=None, gKt_init=None, gX_init=None):
        P, V = Kt.shape
        R, Cc = X.shape[0], X.shape[1]
        Kt, X = np.ascontiguousarray(Kt, self.np_t), np.ascontiguousarray(X, self.np_t)
        Y = np.empty((R, Cc, P), self.np_t)
        gKt = np.zeros_like(Kt) if gKt_init is None else np.array(gKt_init, self.np_t, order="C")
        gX = np.zeros_like(X) if gX_init is None else np.array(gX_init, self.np_t, order="C")
        g = None if gY is None else np.ascontiguousarray(gY, self.np_t)
        self._fn("gfref_custom_matmul_tensor")(_ptr(Kt, self.c_t), _ptr(X, self.c_t), _ptr(Y, self.c_t),
                                               None if g is None else _ptr(g, self.c_t), _ptr(gKt, self.c_t),
                                               _ptr(gX, self.c_t), ctypes.c_int(R), ctypes.c_int(Cc), ctypes.c_int(V),
                                               ctypes.c_int(P))
        return Y if gY is None else (Y, gKt, gX)

    def promote(self, f, pos, gQ=None, gf_init=None):
        m, C = f.shape[0], f.shape[2]
        n = len(pos)
        f = np.ascontiguousarray(f, self.np_t)
        pos = np.ascontiguousarray(pos, np.int32)
        Q = np.empty((n, n, C), self.np_t)
        gf = np.zeros_like(f) if gf_init is None else np.array(gf_init, self.np_t, order="C")
        g = None if gQ is None else np.ascontiguousarray(gQ, self.np_t)
        self._fn("gfref_promote")(_ptr(f, self.c_t), pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _ptr(Q, self.c_t),
                                  None if g is None else _ptr(g, self.c_t), _ptr(gf, self.c_t), ctypes.c_int(n),
                                  ctypes.c_int(m), ctypes.c_int(C))
        return Q if gQ is None else (Q, gf)

    def level_forward_backward(self, T, adj, K, bias, gZ=None):
        N, C = T.shape[0], T.shape[3]
        Cout = K.shape[1]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        K = np.ascontiguousarray(K, self.np_t)
        bias = np.ascontiguousarray(bias, self.np_t)
        contracted = np.empty((N, N, 18 * C), self.np_t)
        Z = np.empty((N, N, Cout), self.np_t)
        null = ctypes.POINTER(self.c_t)()
        if gZ is None:
            self._fn("gfref_level_forward_backward")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(K, self.c_t),
                                                     _ptr(bias, self.c_t), ctypes.c_int(N), ctypes.c_int(C),
                                                     ctypes.c_int(Cout), _ptr(contracted, self.c_t), _ptr(Z, self.c_t),
                                                     null, null, null, null)
            return contracted, Z
        gZ = np.ascontiguousarray(gZ, self.np_t)
        gT = np.zeros((N, N, N, C), self.np_t)
        gK = np.zeros((18 * C, Cout), self.np_t)
        gb = np.zeros((Cout,), self.np_t)
        self._fn("gfref_level_forward_backward")(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(K, self.c_t),
                                                 _ptr(bias, self.c_t), ctypes.c_int(N), ctypes.c_int(C),
                                                 ctypes.c_int(Cout), _ptr(contracted, self.c_t), _ptr(Z, self.c_t),
                                                 _ptr(gZ, self.c_t), _ptr(gT, self.c_t), _ptr(gK, self.c_t),
                                                 _ptr(gb, self.c_t))
        return contracted, Z, gT, gK, gb

    def matmul_forward(self, X, W):
        M, K = X.shape
        P = W.shape[1]
        X = np.ascontiguousarray(X, self.np_t)
        W = np.ascontiguousarray(W, self.np_t)
        Y = np.empty((M, P), self.np_t)
        self._fn("gfref_matmul_forward")(_ptr(X, self.c_t), _ptr(W, self.c_t), _ptr(Y, self.c_t), ctypes.c_int(M),
                                         ctypes.c_int(K), ctypes.c_int(P))
        return Y

    def matmul_backward(self, X, W, gY, gX_init=None, gW_init=None):
        M, K = X.shape
        P = W.shape[1]
        X = np.ascontiguousarray(X, self.np_t)
        W = np.ascontiguousarray(W, self.np_t)
        gY = np.ascontiguousarray(gY, self.np_t)
        gX = np.zeros((M, K), self.np_t) if gX_init is None else np.array(gX_init, self.np_t, order="C")
        gW = np.zeros((K, P), self.np_t) if gW_init is None else np.array(gW_init, self.np_t, order="C")
        self._fn("gfref_matmul_backward")(_ptr(X, self.c_t), _ptr(W, self.c_t), _ptr(gY, self.c_t), _ptr(gX, self.c_t),
                                          _ptr(gW, self.c_t), ctypes.c_int(M), ctypes.c_int(K), ctypes.c_int(P))
        return gX, gW

    def time_replicas(self, T, adj, gout, threads, reps=1):
        """Wall seconds for `threads` host threads x `reps` (forward+backward) on private replicas."""
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gout = np.ascontiguousarray(gout, self.np_t)
        f = self._fn("gfref_contract18_time_replicas", ctypes.c_double)
        return f(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(gout, self.c_t), ctypes.c_int(N), ctypes.c_int(C),
                 ctypes.c_int(threads), ctypes.c_int(reps))


    def time_thread_variant(self, T, adj, gout, reps=1):
        """Wall seconds for `reps` x (RisiContraction_18_thread forward + backward): 6 threads inside one op."""
        N, C = T.shape[0], T.shape[3]
        T = np.ascontiguousarray(T, self.np_t)
        adj = np.ascontiguousarray(adj, self.np_t)
        gout = np.ascontiguousarray(gout, self.np_t)
        f = self._fn("gfref_contract18_thread_time", ctypes.c_double)
        return f(_ptr(T, self.c_t), _ptr(adj, self.c_t), _ptr(gout, self.c_t), ctypes.c_int(N), ctypes.c_int(C), ctypes.c_int(reps))


_MODEL_LIB = os.path.join(HERE, "_ref", "libgfref_model_f64.so")


def model_available():
    return os.path.exists(_MODEL_LIB)


def smp_beta_num_params(L, C, F, n_depth):
    return C * F * (n_depth + 1) + L * (18 * C * C + C) + C


def omega_widths(L, C):
    """Channel widths of SMP_omega_physics per level 0..L (SMP_omega_physics.h:142-146)."""
    w = [C]
    for _ in range(L):
        w.append(max(1, w[-1] // 2))
    return w


def smp_omega_num_params(L, C, F):
    w = omega_widths(L, C)
    tot = sum(w)
    return C * F + sum(18 * w[l - 1] * w[l] + w[l] for l in range(1, L + 1)) + (tot // 2) * tot + tot // 2


def ref_checkpoint_roundtrip(Vmax, L, C, F, n_depth, params, save_path=None, load_path=None):
    """SMP_beta::save_model / load_model of the compiled reference (SMP_beta.h:980-1002): writes `params` to save_path,
    then loads load_path into the model and returns its parameters (or None)."""
    lib = ctypes.CDLL(_MODEL_LIB)
    params = np.ascontiguousarray(params, np.float64)
    loaded = np.zeros_like(params)
    fn = lib.gfref_smp_beta_checkpoint_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p]
    total = fn(Vmax, L, C, F, n_depth, params.ctypes.data, save_path.encode() if save_path else None,
               load_path.encode() if load_path else None, loaded.ctypes.data)
    assert total == params.size, (total, params.size)
    return loaded if load_path else None


def ref_smp_beta_batchlearn(graphs, targets, L, C, n_depth, params, epochs, lr):
    """SMP_beta::BatchLearn (SMP_beta.h:745-772) of the compiled reference, `epochs` times on the same batch.
    graphs: list of (adj [V,V], feat [V,F]).  Returns (losses [epochs, 2] = summed loss before/after, final params)."""
    lib = ctypes.CDLL(_MODEL_LIB)
    F = graphs[0][1].shape[1]
    V = np.array([a.shape[0] for a, _ in graphs], np.int32)
    adj = np.concatenate([np.asarray(a, np.int32).ravel() for a, _ in graphs])
    feat = np.concatenate([np.asarray(f, np.float64).ravel() for _, f in graphs])
    params = np.ascontiguousarray(params, np.float64)
    tg = np.ascontiguousarray(targets, np.float64)
    losses, out = np.zeros((epochs, 2)), np.zeros_like(params)
    fn = lib.gfref_smp_beta_batchlearn_f64
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]
    total = fn(len(graphs), V.ctypes.data, adj.ctypes.data, feat.ctypes.data, L, C, F, n_depth, params.ctypes.data, tg.ctypes.data,
               epochs, lr, losses.ctypes.data, out.ctypes.data)
    assert total == params.size, (total, params.size)
    return losses, out


def ref_optimizer(mode, values, grads, n0, alpha, n_batch=1):
    """The compiled reference's Adam / Momentum on two registered parameter vectors (sizes n0, len - n0).
    mode: "adam_batch" = Adam::Learn(alpha, nBatch), "adam" = Adam::Learn(alpha), "momentum" = Momentum::Learn(alpha, nBatch).
    grads: [steps, len].  Returns the updated values."""
    lib = ctypes.CDLL(_MODEL_LIB)
    values = np.array(values, np.float64)
    grads = np.ascontiguousarray(grads, np.float64)
    lib.gfref_optimizer_f64.restype = None
    lib.gfref_optimizer_f64(ctypes.c_int({"adam_batch": 0, "adam": 1, "momentum": 2}[mode]), _ptr(values, ctypes.c_double),
                            _ptr(grads, ctypes.c_double), ctypes.c_int(n0), ctypes.c_int(values.size - n0),
                            ctypes.c_int(grads.shape[0]), ctypes.c_double(alpha), ctypes.c_int(n_batch))
    return values


def adam_reference_restatement(values, grads, alpha, n_batch=None, beta1=0.9, beta2=0.999, eps=1e-8):
    """fp64 restatement of Adam::Learn (Adam.h:76-137) on a flat vector; n_batch=None is the `Learn(alpha)` overload."""
    p = np.array(values, np.float64)
    m, v = np.zeros_like(p), np.zeros_like(p)
    b1t = b2t = 1.0
    for g in np.asarray(grads, np.float64):
        if n_batch is None:
            b1t *= beta1
            b2t *= beta2
            m = beta1 * m + (1 - beta1) * g
            v = beta2 * v + (1 - beta2) * g * g
            p -= alpha * (m / (1 - b1t)) / (np.sqrt(v / (1 - b2t)) + eps)
        else:
            gg = g / n_batch
            m = beta1 * m + (1 - beta1) * gg
            v = beta2 * v + (1 - beta2) * gg * gg
            k = np.arange(1, p.size + 1, dtype=np.float64)
            c1 = b1t * beta1 ** k        # the powers advance once per element (Adam.h:123,127)
            c2 = b2t * beta2 ** k
            b1t, b2t = c1[-1], c2[-1]
            p -= alpha * (m / (1 - c1)) / (np.sqrt(v / (1 - c2)) + eps)
    return p


def ref_smp_omega_physics(adj, feat, max_field, L, C, params, target):
    """The unmodified SMP_omega_physics on one graph: dict(feature [Ctot], loss, grads, phi)."""
    lib = ctypes.CDLL(_MODEL_LIB)
    adj = np.ascontiguousarray(adj, np.int32)
    feat = np.ascontiguousarray(feat, np.float64)
    V, F = feat.shape
    params = np.ascontiguousarray(params, np.float64)
    assert params.size == smp_omega_num_params(L, C, F)
    tot = sum(omega_widths(L, C))
    gfeat, loss, grads = np.zeros(tot), np.zeros(1), np.zeros(params.size)
    phi = np.zeros((L + 1, V, V + 1), np.int32)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))     # noqa: E731
    fn = lib.gfref_smp_omega_physics_f64
    fn.restype = ctypes.c_int
    n = fn(V, ip(adj), dp(feat), max_field, L, C, F, tot, dp(params), ctypes.c_double(target), dp(gfeat), dp(loss), dp(grads), ip(phi))
    assert n == params.size, (n, params.size)
    fields = [[list(phi[l, v, 1:1 + phi[l, v, 0]]) for v in range(V)] for l in range(L + 1)]
    return {"feature": gfeat, "loss": float(loss[0]), "grads": grads, "phi": fields}


def smp_omega_pairgraphs_num_params(L, C, F1, F2):
    """SMP_omega_pairgraphs.h:290-335: H_1, H_2, (K1_l, b1_l, K2_l, b2_l) per level, W1, W2, W3."""
    w = omega_widths(L, C)
    tot = 2 * sum(w)
    h1 = max(tot // 2, 10)
    h2 = max(h1 // 2, 10)
    return C * F1 + C * F2 + 2 * sum(18 * w[l - 1] * w[l] + w[l] for l in range(1, L + 1)) + h1 * tot + h2 * h1 + h2


def ref_smp_omega_pairgraphs(adj1, feat1, adj2, feat2, max_field, L, C, params, target):
    """The unmodified SMP_omega_pairgraphs on one (graph, line graph) example: dict(feature [Ctot], loss, predict, grads)."""
    lib = ctypes.CDLL(_MODEL_LIB)
    adj1, adj2 = np.ascontiguousarray(adj1, np.int32), np.ascontiguousarray(adj2, np.int32)
    feat1, feat2 = np.ascontiguousarray(feat1, np.float64), np.ascontiguousarray(feat2, np.float64)
    (V1, F1), (V2, F2) = feat1.shape, feat2.shape
    params = np.ascontiguousarray(params, np.float64)
    assert params.size == smp_omega_pairgraphs_num_params(L, C, F1, F2)
    tot = 2 * sum(omega_widths(L, C))
    gfeat, loss, pred, grads = np.zeros(tot), np.zeros(1), np.zeros(1), np.zeros(params.size)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))     # noqa: E731
    fn = lib.gfref_smp_omega_pairgraphs_f64
    fn.restype = ctypes.c_int
    n = fn(V1, ip(adj1), dp(feat1), F1, V2, ip(adj2), dp(feat2), F2, max_field, L, C, tot, dp(params), ctypes.c_double(target),
           dp(gfeat), dp(loss), dp(pred), dp(grads))
    assert n == params.size, (n, params.size)
    return {"feature": gfeat, "loss": float(loss[0]), "predict": float(pred[0]), "grads": grads}


def ref_smp_omega(adj, feat, max_field, L, C, n_depth, params, target):
    """The unmodified SMP_omega (SMP_beta's wiring, receptive fields limited to max_field) on one graph."""
    lib = ctypes.CDLL(_MODEL_LIB)
    adj = np.ascontiguousarray(adj, np.int32)
    feat = np.ascontiguousarray(feat, np.float64)
    V, F = feat.shape
    params = np.ascontiguousarray(params, np.float64)
    assert params.size == smp_beta_num_params(L, C, F, n_depth)
    gfeat, loss, grads = np.zeros(C), np.zeros(1), np.zeros(params.size)
    phi = np.zeros((L + 1, V, V + 1), np.int32)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))     # noqa: E731
    fn = lib.gfref_smp_omega_f64
    fn.restype = ctypes.c_int
    n = fn(V, ip(adj), dp(feat), max_field, L, C, F, n_depth, dp(params), ctypes.c_double(target), dp(gfeat), dp(loss), dp(grads), ip(phi))
    assert n == params.size, (n, params.size)
    fields = [[list(phi[l, v, 1:1 + phi[l, v, 0]]) for v in range(V)] for l in range(L + 1)]
    return {"feature": gfeat, "loss": float(loss[0]), "grads": grads, "phi": fields}


def ref_smp_2d_ver8(adj, feat, L, C, n_depth, params, target):
    """SMP_2D_ver8 (K_l stored [C, 18 C]); same interface as ref_smp_beta."""
    return ref_smp_beta(adj, feat, L, C, n_depth, params, target, symbol="gfref_smp_2d_ver8_f64")


def ref_smp_beta(adj, feat, L, C, n_depth, params, target, symbol="gfref_smp_beta_f64"):
    """The UNMODIFIED reference model SMP_beta (double tree) on one graph with the given flat parameters (optimizer
    order): returns dict(feature [C], loss, grads [flat], phi = list per level of per-vertex receptive fields)."""
    lib = ctypes.CDLL(_MODEL_LIB)
    adj = np.ascontiguousarray(adj, np.int32)
    feat = np.ascontiguousarray(feat, np.float64)
    V, F = feat.shape
    params = np.ascontiguousarray(params, np.float64)
    assert params.size == smp_beta_num_params(L, C, F, n_depth)
    gfeat, loss, grads = np.zeros(C), np.zeros(1), np.zeros(params.size)
    phi = np.zeros((L + 1, V, V + 1), np.int32)
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))     # noqa: E731
    fn = getattr(lib, symbol)
    fn.restype = ctypes.c_int
    n = fn(V, ip(adj), dp(feat), L, C, F, n_depth, dp(params), ctypes.c_double(target), dp(gfeat),
                               dp(loss), dp(grads), ip(phi))
    assert n == params.size
    fields = [[list(phi[l, v, 1:1 + phi[l, v, 0]]) for v in range(V)] for l in range(L + 1)]
    return {"feature": gfeat, "loss": float(loss[0]), "grads": grads, "phi": fields}


# The 18 contractions in slab order as einsums over T[a,b,c,f] and A[d,e] (starred rows;
# source lines GraphFlow/RisiContraction_18.h:102-318).  A repeated letter takes a diagonal.
EINSUM18 = [
    "abcf,de->abf", "abcf,de->adf", "abcf,de->bcf", "abcf,de->bdf", "abcf,de->def", "abcf,ce->abf",
    "abcf,dd->abf", "abbf,de->adf", "abcf,db->adf", "abcf,ae->bcf", "abaf,de->bdf", "abcf,da->bdf",
    "abcf,dc->bdf", "aacf,de->def", "abbf,de->def", "abbf,db->adf", "abaf,da->bdf", "aaaf,de->def",
]

# All 50 contractions of RisiContraction_50 in slab order (source lines
# GraphFlow/RisiContraction_50.h:94-428).  RAW adjacency (RisiContraction_50.h:63-65).
EINSUM50 = [
    "abcf,de->abf", "abcf,de->acf", "abcf,de->adf", "abcf,de->aef", "abcf,de->bcf", "abcf,de->bdf",
    "abcf,de->bef", "abcf,de->cdf", "abcf,de->cef", "abcf,de->def", "abcf,ce->abf", "abcf,dc->abf",
    "abcf,dd->abf", "abcf,be->acf", "abcf,db->acf", "abcf,dd->acf", "abbf,de->adf", "abcf,db->adf",
    "abcf,dc->adf", "abbf,de->aef", "abcf,be->aef", "abcf,ce->aef", "abcf,ae->bcf", "abcf,da->bcf",
    "abcf,dd->bcf", "abaf,de->bdf", "abcf,da->bdf", "abcf,dc->bdf", "abaf,de->bef", "abcf,ae->bef",
    "abcf,ce->bef", "aacf,de->cdf", "abcf,da->cdf", "abcf,db->cdf", "aacf,de->cef", "abcf,ae->cef",
    "abcf,be->cef", "aacf,de->def", "abaf,de->def", "abbf,de->def", "abcf,cc->abf", "abcf,bb->acf",
    "abbf,db->adf", "abbf,be->aef", "abcf,aa->bcf", "abaf,da->bdf", "abaf,ae->bef", "aacf,da->cdf",
    "aacf,ae->cef", "aaaf,de->def",
]


def _split(spec):
    ins, out = spec.split("->")
    t_idx, a_idx = ins.split(",")
    return t_idx, a_idx, out


def _uniq(letters):
    seen = []
    for ch in letters:
        if ch not in seen:
            seen.append(ch)
    return "".join(seen)


def _einsum_pair(spec, T, A):
    """einsum(spec, T, A) with single-operand pre-reduction so no intermediate exceeds O(N^3 C)."""
    t_idx, a_idx, out = _split(spec)
    keep_t = "".join(ch for ch in _uniq(t_idx) if ch in out or ch in a_idx)
    keep_a = "".join(ch for ch in _uniq(a_idx) if ch in out or ch in t_idx)
    Tr = np.einsum("%s->%s" % (t_idx, keep_t), T)
    Ar = np.einsum("%s->%s" % (a_idx, keep_a), A)
    return np.einsum("%s,%s->%s" % (keep_t, keep_a, out), Tr, Ar, optimize=True)


def einsum_forward(specs, T, adj, positive_part):
    """fp64 closed form of a contraction family: out[x, y, k*C + f]."""
    T = np.asarray(T, np.float64)
    A = np.asarray(adj, np.float64)
    if positive_part:
        A = np.maximum(A, 0.0)
    N, C = T.shape[0], T.shape[3]
    out = np.empty((N, N, len(specs), C), np.float64)
    for k, spec in enumerate(specs):
        out[:, :, k, :] = _einsum_pair(spec, T, A)
    return out.reshape(N, N, len(specs) * C)


def einsum18_forward(T, adj, positive_part=True):
    return einsum_forward(EINSUM18, T, adj, positive_part)


def einsum50_forward(T, adj):
    return einsum_forward(EINSUM50, T, adj, False)


def einsum18_backward(gout, adj, positive_part=True):
    return einsum_backward(EINSUM18, gout, adj, positive_part)


def einsum50_backward(gout, adj):
    return einsum_backward(EINSUM50, gout, adj, False)


def einsum_backward(specs, gout, adj, positive_part):
    """fp64 transpose of einsum_forward with respect to T (fresh gradient)."""
    A = np.asarray(adj, np.float64)
    if positive_part:
        A = np.maximum(A, 0.0)
    N = gout.shape[0]
    S = len(specs)
    C = gout.shape[2] // S
    g = np.asarray(gout, np.float64).reshape(N, N, S, C)
    gT = np.zeros((N, N, N, C), np.float64)
    ones = np.ones(N)
    for k, spec in enumerate(specs):
        t_idx, a_idx, out = _split(spec)
        pos = t_idx[:3]          # letters at T's three positions (repeats = diagonal)
        u = _uniq(pos)           # distinct letters of the diagonal view Tv[u..., f]
        keep_a = "".join(ch for ch in _uniq(a_idx) if ch in out or ch in u)
        Ar = np.einsum("%s->%s" % (a_idx, keep_a), A)
        subs, ops = [out, keep_a], [g[:, :, k, :], Ar]
        for ch in u:             # letters of T that reach neither the output nor adj: gradient broadcasts
            if ch not in out and ch not in keep_a:
                subs.append(ch)
                ops.append(ones)
        gv = np.einsum(",".join(subs) + "->" + u + "f", *ops, optimize=True)
        grids = np.meshgrid(*([np.arange(N)] * len(u)), indexing="ij")
        idx = tuple(grids[u.index(ch)] for ch in pos)
        gT[idx] += gv
    return gT


# RisiContraction_10 = the first ten rows of the 50 (RisiContraction_10.h:91-140), raw adjacency.
EINSUM10 = EINSUM50[:10]


def einsum10_forward(T, adj):
    return einsum_forward(EINSUM10, T, adj, False)


def einsum10_backward(gout, adj):
    return einsum_backward(EINSUM10, gout, adj, False)


def einsum4_forward(T):
    """RisiContraction_4.h:79-118: sum_c T[a,b,c] -> (a,b); sum_a -> (b,c); T[a,a,c] -> (a,c); T[a,b,b] -> (a,b)."""
    T = np.asarray(T, np.float64)
    N, C = T.shape[0], T.shape[3]
    out = np.empty((N, N, 4, C), np.float64)
    out[:, :, 0] = np.einsum("abcf->abf", T)
    out[:, :, 1] = np.einsum("abcf->bcf", T)
    out[:, :, 2] = np.einsum("aacf->acf", T)
    out[:, :, 3] = np.einsum("abbf->abf", T)
    return out.reshape(N, N, 4 * C)


def einsum4_backward(gout):
    N = gout.shape[0]
    C = gout.shape[2] // 4
    g = np.asarray(gout, np.float64).reshape(N, N, 4, C)
    gT = g[:, :, None, 0, :] + g[None, :, :, 1, :]
    gT = np.array(np.broadcast_to(gT, (N, N, N, C)))
    i = np.arange(N)
    gT[i, i, :, :] += g[:, :, 2, :]          # [a, a, c] += g2[a, c]
    gT[:, i, i, :] += g[:, :, 3, :]          # [a, b, b] += g3[a, b]
    return gT


def einsum18_dropout_forward(T, adj, use):
    """RisiContraction_18_dropout.h:148-478: the 18 slabs with A+ (`adj_value > 0`), dropped slabs left at zero."""
    out = einsum18_forward(T, adj, True)
    N, C = T.shape[0], T.shape[3]
    out = out.reshape(N, N, 18, C)
    out[:, :, [k for k in range(18) if not use[k]], :] = 0.0
    return out.reshape(N, N, 18 * C)


def einsum18_dropout_backward(gout, adj, use):
    N = gout.shape[0]
    C = gout.shape[2] // 18
    g = np.array(gout, np.float64).reshape(N, N, 18, C)
    g[:, :, [k for k in range(18) if not use[k]], :] = 0.0
    return einsum18_backward(g.reshape(N, N, 18 * C), adj, True)


def slab_rel_err(x, ref, n_slabs=18):
    """The parity metric of SURVEY.md section 8(c): per slab k, max|x - ref| / max|ref_k|; returns the worst slab."""
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    if n_slabs > 1:
        C = ref.shape[-1] // n_slabs
        xs = x.reshape(-1, n_slabs, C)
        rs = ref.reshape(-1, n_slabs, C)
        worst = 0.0
        for k in range(n_slabs):
            den = np.abs(rs[:, k]).max()
            num = np.abs(xs[:, k] - rs[:, k]).max()
            worst = max(worst, num / den if den > 0 else num)
        return worst
    den = np.abs(ref).max()
    num = np.abs(x - ref).max()
    return num / den if den > 0 else num
