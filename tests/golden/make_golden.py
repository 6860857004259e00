"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference).  Run here, in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Vectors (all small, .npz):
  kat_c1_n8_c4.npz     the reference test's own recipe (tests/test_RisiContraction_18_gpu.cu:80-121, 201-204):
                       srand(123456789); N symmetric tensors of rand()%10; adj = I + symmetric rand()%2;
                       gout = rand()%100.  Integer-valued, so fp32 and fp64 results are exact and identical.
  real_n6_c8.npz       uniform real T, 0/1(+I) adjacency, fp64 reference, forward + backward (+= into a non-zero gT).
  signed_n5_c3.npz     uniform real T and *signed real* adjacency (exercises the adj<=0 skip, RisiContraction_18.h:90).
  level_n6_c4.npz      contraction -> Reshape2D -> MatMul(K) -> +bias -> LeakyReLU chain, forward + backward
                       (SMP_beta.h:596-616 wiring).
  r50_n5_c2.npz        RisiContraction_50 forward + backward (+= into a non-zero gT), signed real adjacency (raw).
  aux_ops.npz          TensorMul, CustomMatMulTensor and the promotion X f X^T (MatTensorMul + TensorMatMul with 0/1
                       selection matrices, SMP_beta.h:446-459, 588-594), forward + backward with non-zero initial gradients.
  smp_beta_model.npz   the whole reference model SMP_beta (L=2, C=4) on three graphs: graph feature, loss and every
                       parameter gradient after one forward/backward, for caller-supplied parameters.
  kat_r50_n10_c5.npz   the reference R50 test's own recipe (tests/test_RisiContraction_50.cpp:14-16,49-80): no srand (glibc
                       seed 1); N=10 symmetric tensors of rand()%100, 5 channels; symmetric rand()%2 adjacency with a
                       zero diagonal; plus gout = rand()%100 for the backward.  Integer-valued and below 2^24, so fp32
                       results are exact.  `python tests/golden/make_golden.py kat50` writes only this file.
  family_n5_c2.npz     RisiContraction_4, RisiContraction_10 and RisiContraction_18_dropout (train mode: srand(seed), the
                       mask the reference drew with rand() is stored; test mode: all slabs, scaled by nKept/18), forward +
                       backward.  `python tests/golden/make_golden.py family` writes only this file.
  matmul_20x36x5.npz   MatMul forward/backward with pre-loaded non-zero input gradients (tests/test_MatMul_gpu.cu:103-116).
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle  # noqa: E402


def glibc_rand_stream(seed):
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(ctypes.c_uint(seed))
    libc.rand.restype = ctypes.c_int
    return libc.rand


def kat_inputs(N, C, seed=123456789):
    rand = glibc_rand_stream(seed)
    T = np.zeros((N, N, N, C), np.float64)
    for i in range(N):
        for ch in range(C):
            for row in range(N):
                for col in range(row, N):
                    v = rand() % 10
                    T[i, row, col, ch] = v
                    T[i, col, row, ch] = v
    adj = np.zeros((N, N), np.float64)
    for i in range(N):
        adj[i, i] = 1
        for j in range(i + 1, N):
            v = rand() % 2
            adj[i, j] = v
            adj[j, i] = v
    gout = np.zeros((N, N, 18 * C), np.float64)
    flat = gout.reshape(-1)
    for i in range(flat.size):
        flat[i] = rand() % 100
    return T, adj, gout


def kat_r50_fixture(r64, r32):
    """tests/test_RisiContraction_50.cpp:49-80 (generation order: tensors, then the adjacency), then gout."""
    N, C = 10, 5
    rand = glibc_rand_stream(1)  # the test never calls srand
    T = np.zeros((N, N, N, C), np.float64)
    for i in range(N):
        for ch in range(C):
            for row in range(N):
                for col in range(row, N):
                    v = rand() % 100
                    T[i, row, col, ch] = v
                    T[i, col, row, ch] = v
    adj = np.zeros((N, N), np.float64)
    for i in range(N):
        for j in range(i + 1, N):
            v = rand() % 2
            adj[i, j] = v
            adj[j, i] = v
    gout = np.zeros((N, N, 50 * C), np.float64)
    flat = gout.reshape(-1)
    for i in range(flat.size):
        flat[i] = rand() % 100
    out, gT = r64.contract50_forward(T, adj), r64.contract50_backward(gout, adj)
    assert np.array_equal(out, r32.contract50_forward(T, adj).astype(np.float64)), "integer KAT must be exact in both trees"
    assert np.array_equal(gT, r32.contract50_backward(gout, adj).astype(np.float64))
    assert max(np.abs(out).max(), np.abs(gT).max()) < 2 ** 24
    # the same tensors are the input of tests/test_RisiContraction_4_thread.cpp:49-66 (same stream, no adjacency)
    g4 = np.ascontiguousarray(gout[:, :, :4 * C])
    out4, gT4 = r64.contract4(T, g4)
    np.savez_compressed(os.path.join(HERE, "kat_r50_n10_c5.npz"), T=T, adj=adj, gout=gout, out=out, gT=gT, g4=g4, out4=out4,
                        gT4=gT4)


def family_fixture(r64):
    """RisiContraction_4.h:68-180, RisiContraction_10.h:72-230, RisiContraction_18_dropout.h:104-797."""
    rng = np.random.default_rng(20261020)
    N, C = 5, 2
    T = rng.uniform(-1, 1, (N, N, N, C))
    adj = rng.uniform(-1, 1, (N, N))
    g4, g10, g18 = (rng.uniform(-1, 1, (N, N, k * C)) for k in (4, 10, 18))
    out4, gT4 = r64.contract4(T, g4)
    out10, gT10 = r64.contract10(T, adj, g10)
    d = {"T": T, "adj": adj, "g4": g4, "g10": g10, "g18": g18, "out4": out4, "gT4": gT4, "out10": out10, "gT10": gT10}
    for i, (kept, seed) in enumerate(((7, 11), (12, 5), (1, 3))):
        out, gT, use = r64.contract18_dropout(T, adj, g18, kept, seed, train=True)
        assert sum(use) == kept
        d.update({"drop%d_out" % i: out, "drop%d_gT" % i: gT, "drop%d_use" % i: np.array(use, np.int32)})
    out, _, use = r64.contract18_dropout(T, adj, g18, 7, 1, train=False)
    assert all(use)
    d.update({"test_out": out, "test_kept": np.int32(7)})
    np.savez_compressed(os.path.join(HERE, "family_n5_c2.npz"), **d)


def pairgraphs_inputs():
    """Two (graph, line graph) examples for SMP_omega_pairgraphs; deterministic."""
    from tests.util import line_graph, molecular_adjacency

    rng = np.random.default_rng(20261020)
    L, C, F, mf = 2, 8, 3, 4
    params = rng.uniform(-1, 1, pyoracle.smp_omega_pairgraphs_num_params(L, C, F, F)) * 0.2
    ex = []
    for V in (7, 5):
        adj = (molecular_adjacency(V, rng, self_loops=False) > 0).astype(np.int32)
        feat = rng.uniform(0, 1, (V, F))
        a2, f2 = line_graph(adj, feat)
        ex.append((adj, feat, a2, f2, float(V) / 2))
    return L, C, F, mf, params, ex


def pairgraphs_fixture():
    """SMP_omega_pairgraphs (SMP_omega_pairgraphs.h) on two (graph, line graph) examples, caller-supplied parameters."""
    L, C, F, mf, params, ex = pairgraphs_inputs()
    d = {"L": L, "C": C, "F": F, "max_field": mf, "params": params}
    for i, (adj, feat, a2, f2, target) in enumerate(ex):
        out = pyoracle.ref_smp_omega_pairgraphs(adj, feat, a2, f2, mf, L, C, params, target)
        d.update({"adj%d" % i: adj, "feat%d" % i: feat, "ladj%d" % i: a2, "lfeat%d" % i: f2, "target%d" % i: target,
                  "feature%d" % i: out["feature"], "loss%d" % i: out["loss"], "predict%d" % i: out["predict"], "grads%d" % i: out["grads"]})
    np.savez_compressed(os.path.join(HERE, "smp_omega_pairgraphs.npz"), **d)


def main():
    pyoracle.build(ref=True)
    r64 = pyoracle.RefOracle("f64")
    r32 = pyoracle.RefOracle("f32")
    if sys.argv[1:] == ["family"]:
        family_fixture(r64)
        return
    if sys.argv[1:] == ["kat50"]:
        kat_r50_fixture(r64, r32)
        return

    # --- c1 KAT -------------------------------------------------------------------------------------------------
    T, adj, gout = kat_inputs(8, 4)
    out64 = r64.contract18_forward(T, adj)
    out32 = r32.contract18_forward(T, adj)
    assert np.array_equal(out64, out32.astype(np.float64)), "integer KAT must be exact in both trees"
    gT64 = r64.contract18_backward(gout, adj)
    gT32 = r32.contract18_backward(gout, adj)
    assert np.array_equal(gT64, gT32.astype(np.float64))
    np.savez_compressed(os.path.join(HERE, "kat_c1_n8_c4.npz"), T=T.astype(np.float32), adj=adj.astype(np.float32),
                        gout=gout.astype(np.float32), out=out64.astype(np.float32), gT=gT64.astype(np.float32))

    # --- real-valued, 0/1 adjacency ------------------------------------------------------------------------------
    rng = np.random.default_rng(20261017)
    N, C = 6, 8
    T = rng.uniform(-1, 1, (N, N, N, C))
    up = np.triu((rng.uniform(size=(N, N)) < 0.35).astype(np.float64), 1)
    adj = up + up.T + np.eye(N)
    gout = rng.uniform(-1, 1, (N, N, 18 * C))
    gT0 = rng.uniform(-1, 1, (N, N, N, C))
    np.savez_compressed(os.path.join(HERE, "real_n6_c8.npz"), T=T, adj=adj, gout=gout, gT0=gT0,
                        out=r64.contract18_forward(T, adj), gT=r64.contract18_backward(gout, adj, gT0))

    # --- signed real adjacency -----------------------------------------------------------------------------------
    N, C = 5, 3
    T = rng.uniform(-1, 1, (N, N, N, C))
    adj = rng.uniform(-1, 1, (N, N))
    gout = rng.uniform(-1, 1, (N, N, 18 * C))
    np.savez_compressed(os.path.join(HERE, "signed_n5_c3.npz"), T=T, adj=adj, gout=gout,
                        out=r64.contract18_forward(T, adj), gT=r64.contract18_backward(gout, adj),
                        out_raw=r64.contract18_forward(T, adj, "definition"))

    # --- level chain ---------------------------------------------------------------------------------------------
    N, C, Cout = 6, 4, 4
    T = rng.uniform(-1, 1, (N, N, N, C))
    up = np.triu((rng.uniform(size=(N, N)) < 0.4).astype(np.float64), 1)
    adj = up + up.T + np.eye(N)
    K = rng.uniform(-0.3, 0.3, (18 * C, Cout))
    bias = rng.uniform(-0.5, 0.5, (Cout,))
    gZ = rng.uniform(-1, 1, (N, N, Cout))
    contracted, Z, gT, gK, gb = r64.level_forward_backward(T, adj, K, bias, gZ)
    np.savez_compressed(os.path.join(HERE, "level_n6_c4.npz"), T=T, adj=adj, K=K, bias=bias, gZ=gZ,
                        contracted=contracted, Z=Z, gT=gT, gK=gK, gb=gb)

    # --- MatMul --------------------------------------------------------------------------------------------------
    M, Kd, P = 20, 36, 5
    X = rng.integers(0, 10, (M, Kd)).astype(np.float64)
    W = rng.integers(0, 10, (Kd, P)).astype(np.float64)
    gY = rng.integers(0, 100, (M, P)).astype(np.float64)
    gX0 = rng.integers(0, 10, (M, Kd)).astype(np.float64)
    gW0 = rng.integers(0, 10, (Kd, P)).astype(np.float64)
    gX, gW = r64.matmul_backward(X, W, gY, gX0, gW0)
    np.savez_compressed(os.path.join(HERE, "matmul_20x36x5.npz"), X=X, W=W, gY=gY, gX0=gX0, gW0=gW0,
                        Y=r64.matmul_forward(X, W), gX=gX, gW=gW)
    # --- RisiContraction_50 (RisiContraction_50.h:73-802), own generator so the other fixtures stay bit-identical ----
    rng50 = np.random.default_rng(20261017)
    N, C = 5, 2
    T = rng50.uniform(-1, 1, (N, N, N, C))
    adj = rng50.uniform(-1, 1, (N, N))
    gout = rng50.uniform(-1, 1, (N, N, 50 * C))
    gT0 = rng50.uniform(-1, 1, (N, N, N, C))
    np.savez_compressed(os.path.join(HERE, "r50_n5_c2.npz"), T=T, adj=adj, gout=gout, gT0=gT0,
                        out=r64.contract50_forward(T, adj), gT=r64.contract50_backward(gout, adj, gT0))
    # --- TensorMul, CustomMatMulTensor, promotion (MatTensorMul + TensorMatMul with selection matrices) ------------
    rngx = np.random.default_rng(20261018)
    A, B = rngx.uniform(-1, 1, (5, 4, 3)), rngx.uniform(-1, 1, (4, 6, 3))
    g, gA0, gB0 = rngx.uniform(-1, 1, (5, 6, 3)), rngx.uniform(-1, 1, (5, 4, 3)), rngx.uniform(-1, 1, (4, 6, 3))
    out, gA, gB = r64.tensor_mul(A, B, g, gA0, gB0)
    Kt, X = rngx.uniform(-1, 1, (4, 8)), rngx.uniform(-1, 1, (3, 5, 8))
    gY, gKt0, gX0 = rngx.uniform(-1, 1, (3, 5, 4)), rngx.uniform(-1, 1, (4, 8)), rngx.uniform(-1, 1, (3, 5, 8))
    Y, gKt, gX = r64.custom_matmul_tensor(Kt, X, gY, gKt0, gX0)
    f, pos = rngx.uniform(-1, 1, (4, 4, 3)), np.array([2, -1, 0, 3, -1, 1], np.int32)
    gQ, gf0 = rngx.uniform(-1, 1, (6, 6, 3)), rngx.uniform(-1, 1, (4, 4, 3))
    Q, gf = r64.promote(f, pos, gQ, gf0)
    np.savez_compressed(os.path.join(HERE, "aux_ops.npz"), tm_A=A, tm_B=B, tm_g=g, tm_gA0=gA0, tm_gB0=gB0, tm_out=out, tm_gA=gA,
                        tm_gB=gB, cm_Kt=Kt, cm_X=X, cm_gY=gY, cm_gKt0=gKt0, cm_gX0=gX0, cm_Y=Y, cm_gKt=gKt, cm_gX=gX,
                        pr_f=f, pr_pos=pos, pr_gQ=gQ, pr_gf0=gf0, pr_Q=Q, pr_gf=gf)
    # --- whole model: SMP_beta (SMP_beta.h) on three small molecular-like graphs, caller-supplied parameters -----------
    from tests.util import molecular_adjacency

    rngm = np.random.default_rng(20261019)
    L, C, F, D = 2, 4, 3, 2
    nparams = pyoracle.smp_beta_num_params(L, C, F, D)
    params = rngm.uniform(-1, 1, nparams) * 0.08
    model = {"L": L, "C": C, "F": F, "D": D, "params": params}
    for gi, V in enumerate((6, 9, 4)):
        adj = (molecular_adjacency(V, rngm, self_loops=False) > 0).astype(np.int32)
        feat = np.eye(F)[rngm.integers(0, F, V)]
        target = float(V)
        out = pyoracle.ref_smp_beta(adj, feat, L, C, D, params, target)
        model.update({"adj%d" % gi: adj, "feat%d" % gi: feat, "target%d" % gi: target, "feature%d" % gi: out["feature"],
                      "loss%d" % gi: out["loss"], "grads%d" % gi: out["grads"],
                      "phi%d" % gi: np.array([len(f) for f in out["phi"][L]], np.int32)})
    np.savez_compressed(os.path.join(HERE, "smp_beta_model.npz"), **model)
    family_fixture(r64)
    kat_r50_fixture(r64, r32)
    pairgraphs_fixture()
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
