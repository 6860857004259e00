#!/usr/bin/env python
"""Generate graphflow_b200/csrc/r50_table.inc: the factorised evaluation plan of the 50 contractions of
RisiContraction_50 (GraphFlow/RisiContraction_50.h:94-428; einsum statement in SURVEY.md Appendix A).

Every case out[x,y] = sum T[..] * A[..] factorises (per channel) into one of four forms over a small set of
T-side intermediates:

  planes   PL[p,q]   0 Pab=sum_c T   1 Pac=sum_b T   2 Pbc=sum_a T
                     3+w Wab_w=sum_c T w[c]   6+w Wac_w=sum_b T w[b]   9+w Wbc_w=sum_a T w[a]    w: 0 r (row sums of A),
                     12 Daac=T[a,a,c]   13 Daba=T[a,b,a]   14 Dabb=T[a,b,b]                       1 cs (column sums), 2 dg (diagonal)
  vectors  V[i]      0 Sa  1 Sb  2 Sc  3 sum_b T[a,b,b]  4 sum_a T[a,b,a]  5 sum_a T[a,a,c]
  scalars  X         0 sum T   1 sum T[a,a,c]   2 sum T[a,b,a]   3 sum T[a,b,b]   4 sum T[a,a,a]

  form 0   out[x,y] = s * PL[id][x,y]                     s: 0 -> 1, 1 -> sA (sum of A), 2 -> tr (trace of A)
  form 1   out[x,y] = V[id][x] * w[aux][y]                aux: 0 r, 1 cs
  form 2   out[x,y] = sum_j PL[id][x,j] * A[y,j]          flag bit 0: plane read as [j,x]; bit 1: A read as [j,y]
  form 3   out[x,y] = X[id] * A[x,y]

Run:  python graphflow_b200/csrc/gen/gen_r50_table.py   (rewrites r50_table.inc; `--check` also validates the plan
against numpy einsum on random input).  The emitted file is committed; the library does not need Python to build.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

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

PLAIN = {"ab": 0, "ac": 1, "bc": 2}
DIAG_PLANE = {"aac": 12, "aba": 13, "abb": 14}
DIAG_VEC = {"abb": 3, "aba": 4, "aac": 5}       # free letter a / b / c respectively
DIAG_SCAL = {"abc": 0, "aac": 1, "aba": 2, "abb": 3, "aaa": 4}


def plan(spec):
    ins, out = spec.split("->")
    t, a = ins.split(",")
    t, out = t[:3], out[:2]
    tl = sorted(set(t))
    in_t = [ch for ch in out if ch in tl]
    in_a = [ch for ch in out if ch in a]
    assert len(in_t) + len(in_a) == 2
    if len(in_t) == 2:
        assert t == "abc"
        third = [ch for ch in "abc" if ch not in out][0]
        pair = "".join(out)
        if third not in a:                                   # A fully reduced: total or trace
            return (0, PLAIN[pair], 1 if a[0] != a[1] else 2, 0)
        w = 2 if a[0] == a[1] else (0 if a[0] == third else 1)  # dg / row sums (A[third, e]) / column sums (A[d, third])
        base = {"ab": 3, "ac": 6, "bc": 9}[pair]
        return (0, base + w, 0, 0)
    if len(in_a) == 2:
        return (3, DIAG_SCAL[t], 0, 0)
    x, y = in_t[0], in_a[0]
    other_a = a[1] if a[0] == y else a[0]
    a_is_row = a[0] == y                                      # y indexes the rows of A
    if other_a not in t:                                      # A reduces to a vector over y
        aux = 0 if a_is_row else 1
        if len(tl) == 3:
            return (1, "abc".index(x), aux, 0)
        return (1, DIAG_VEC[t], aux, 0)
    j = other_a                                               # contracted between T and A
    flag_a = 0 if a_is_row else 2                             # A[y,j] or A[j,y]
    if len(tl) == 3:                                          # third T index summed alone -> plain plane over (x, j)
        pair = "".join(sorted(x + j))
        flag_p = 0 if x < j else 1
        return (2, PLAIN[pair], 0, flag_p | flag_a)
    pid = DIAG_PLANE[t]                                       # both non-free T slots tied with A's index
    letters = {"aac": "ac", "aba": "ab", "abb": "ab"}[t]
    flag_p = 0 if letters[0] == x else 1
    return (2, pid, 0, flag_p | flag_a)


def emulate(specs, T, A):
    """numpy evaluation of the plan (what the CUDA kernels compute), for validation."""
    import numpy as np

    N = T.shape[0]
    r, cs, dg = A.sum(1), A.sum(0), np.diag(A).copy()
    w = [r, cs, dg]
    PL = [T.sum(2), T.sum(1), T.sum(0)]
    PL += [np.einsum("abcf,c->abf", T, v) for v in w]
    PL += [np.einsum("abcf,b->acf", T, v) for v in w]
    PL += [np.einsum("abcf,a->bcf", T, v) for v in w]
    PL += [np.einsum("aacf->acf", T), np.einsum("abaf->abf", T), np.einsum("abbf->abf", T)]
    V = [PL[0].sum(1), PL[0].sum(0), PL[1].sum(0), PL[14].sum(1), PL[13].sum(0), PL[12].sum(0)]
    X = [V[0].sum(0), V[5].sum(0), V[4].sum(0), V[3].sum(0), np.einsum("aaf->f", PL[12])]
    scal = [1.0, A.sum(), np.trace(A)]
    out = np.zeros((N, N, len(specs), T.shape[3]))
    for k, spec in enumerate(specs):
        form, idx, aux, flag = plan(spec)
        if form == 0:
            out[:, :, k] = scal[aux] * PL[idx]
        elif form == 1:
            out[:, :, k] = V[idx][:, None, :] * w[aux][None, :, None]
        elif form == 2:
            P = PL[idx].transpose(1, 0, 2) if flag & 1 else PL[idx]
            Am = A.T if flag & 2 else A
            out[:, :, k] = np.einsum("xjf,yj->xyf", P, Am)
        else:
            out[:, :, k] = X[idx][None, None, :] * A[:, :, None]
    return out.reshape(N, N, -1)


def main():
    rows = []
    for k, spec in enumerate(EINSUM50):
        form, idx, aux, flag = plan(spec)
        rows.append("    {%d, %2d, %d, %d},  /* %2d: %s */" % (form, idx, aux, flag, k + 1, spec))
    text = ("// r50_table.inc -- GENERATED by gen/gen_r50_table.py (do not edit): factorised plan of RisiContraction_50's 50 cases\n"
            "// {form, id, aux, flags}; see the generator's docstring and DESIGN.md section 4.5.\n" + "\n".join(rows) + "\n")
    with open(os.path.join(HERE, "..", "r50_table.inc"), "w") as fh:
        fh.write(text)
    if "--check" in sys.argv:
        import numpy as np

        rng = np.random.default_rng(0)
        T, A = rng.uniform(-1, 1, (6, 6, 6, 3)), rng.uniform(-1, 1, (6, 6))
        ref = np.stack([np.einsum(spec, T, A) for spec in EINSUM50], axis=2).reshape(6, 6, -1)
        err = np.abs(emulate(EINSUM50, T, A) - ref).max()
        print("plan vs einsum: max abs err %.2e" % err)
        assert err < 1e-12


if __name__ == "__main__":
    main()
