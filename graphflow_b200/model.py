"""SMP_beta (second-order CCN, GraphFlow/SMP_beta.h) forward + backward for a batch of graphs on the B200 path.

Drop-in for what SMP_beta::BatchLearn computes per example (SMP_beta.h:757-765: complete_computation_graph, forward,
backward, parameter-gradient sum) with every level running as ONE launch set over all vertices of all graphs of the
batch: promotion gather (ccn_promote_forward) -> 18-way contraction (ccn_contract18_forward, ragged receptive fields)
-> tensor-core feature mix with bias + leaky-ReLU (ccn_mix_forward), and the transposed chain backwards.  Level 0
(H . WL-features, SMP_beta.h:563-573) and the read-out head (ShrinkTensor -> LeakyReLU -> SumVectors -> InnerProduct
-> SquaredLoss, :623-639) are O(V C) and stay thin torch epilogues, as SURVEY.md section 2.2 scopes them.

Parameters are the reference's, in its registration order (SMP_beta.h:276-282): H [C, F (nDepth+1)], K_l [18 C, C],
b_l [C] for l = 1..L, W [C]."""
import numpy as np
import torch

from .graph import GraphTables
from .ops import Context

ALPHA = 0.01  # LeakyReLU.h:31, LeakyReLU3D.h:31


class BatchTables:
    """Device index tables of a batch of graphs (built once, reused every step).

    Receptive fields are ragged (in a 24-vertex molecule the level-3 fields have 3..23 members, mean 11), and the kernels
    pad every instance of a launch to that launch's n_max, so the instances of a level are sorted by size and cut into
    `n_buckets` size classes, each with its own n_max and its own launches.  The level's activations live in ONE flat
    buffer (bucket after bucket, instance stride n_max_k^2 C inside bucket k); `slot` maps (level, global vertex) to its
    element offset, which is what the next level's promotion table and the read-out use."""

    def __init__(self, graphs, n_levels, C, device, n_buckets=4):
        self.graphs = graphs
        self.n_levels = n_levels
        self.Vtot = sum(g.V for g in graphs)
        self.graph_of = np.concatenate([np.full(g.V, i, np.int64) for i, g in enumerate(graphs)])
        base = np.cumsum([0] + [g.V for g in graphs])[:-1]
        self.features = torch.from_numpy(np.concatenate([g.features for g in graphs]).astype(np.float32)).to(device)
        self.levels = []
        prev_off = np.arange(self.Vtot, dtype=np.int64) * C  # level 0: one [1, 1, C] tensor per vertex
        self.elems = [self.Vtot * C]                           # flat activation buffer sizes per level (0..L)
        for l in range(n_levels):
            items = [(base[gi] + v, gi, g.levels[l][v]) for gi, g in enumerate(graphs) for v in range(g.V)]
            items.sort(key=lambda t: t[2]["n"])
            sizes = np.array([it["n"] for _, _, it in items])
            cuts = sorted(set(int(c) for c in np.linspace(0, len(items), n_buckets + 1)))
            # do not split a run of equal sizes across buckets needlessly: snap cuts to size changes
            cuts = sorted(set([0, len(items)] + [int(np.searchsorted(sizes, sizes[c - 1], side="right")) for c in cuts[1:-1]]))
            buckets, off, cur_off = [], 0, np.zeros(self.Vtot, np.int64)
            rowsel = []
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                chunk = items[lo:hi]
                if not chunk:
                    continue
                n_max = max(it["n"] for _, _, it in chunk)
                B = len(chunk)
                n = np.array([it["n"] for _, _, it in chunk], np.int32)
                adj = np.zeros((B, n_max * n_max), np.float32)
                f_off = np.zeros((B, n_max), np.int64)
                m = np.ones((B, n_max), np.int32)
                pos = np.full((B, n_max, n_max), -1, np.int32)
                gv = np.array([g for g, _, _ in chunk], np.int64)
                for i, (g_v, gi, it) in enumerate(chunk):
                    k = it["n"]
                    adj[i, :k * k] = it["adj"].ravel()
                    f_off[i, :k] = prev_off[base[gi] + np.asarray(it["src"], np.int64)]
                    m[i, :k] = it["m"]
                    pos[i, :k, :k] = it["pos"]
                    cur_off[g_v] = off + i * n_max * n_max * C
                buckets.append({
                    "n_max": int(n_max), "B": B, "n_host": n, "offset": off, "vertex": torch.from_numpy(gv).to(device),
                    "n": torch.from_numpy(n).to(device), "adj": torch.from_numpy(adj).to(device),
                    "f_off": torch.from_numpy(f_off.ravel()).to(device), "m": torch.from_numpy(m.ravel()).to(device),
                    "pos": torch.from_numpy(pos.ravel()).to(device),
                    # rows of the padded [B, n_max^2] block that really exist (the dense n_i^2 prefix of every instance)
                    "rowmask": torch.from_numpy((np.arange(n_max * n_max)[None, :] < (n.astype(np.int64) ** 2)[:, None])).to(device),
                })
                off += B * n_max * n_max * C
            self.levels.append(buckets)
            self.elems.append(off)
            prev_off = cur_off
        self.contractions = sum(b["B"] for lv in self.levels for b in lv)
        self.padded_rows = [sum(b["B"] * b["n_max"] ** 2 for b in lv) for lv in self.levels]
        self.real_rows = [int(sum((b["n_host"].astype(np.int64) ** 2).sum() for b in lv)) for lv in self.levels]


class SMPBetaB200:
    """k_transposed=True gives SMP_2D_ver8 (SMP_2D_ver8.h): the same wiring with the feature mix done by
    CustomMatMulTensor, i.e. K_l stored [C, 18 C] (SMP_2D_ver8.h:130, 526-527)."""

    def __init__(self, n_levels, C, n_features, n_depth, device=0, ctx=None, k_transposed=False):
        self.L, self.C, self.F, self.D = n_levels, C, n_features, n_depth
        self.k_transposed = k_transposed
        self.device = torch.device("cuda", device)
        self.ctx = ctx if ctx is not None else Context(device)
        kshape = (C, 18 * C) if k_transposed else (18 * C, C)
        self.shapes = [(C, n_features * (n_depth + 1))] + [s for _ in range(n_levels) for s in (kshape, (C,))] + [(C,)]
        self.params = [torch.zeros(s, device=self.device) for s in self.shapes]

    # ---- parameters in the reference's flat order ------------------------------------------------------------------
    def num_params(self):
        return int(sum(np.prod(s) for s in self.shapes))

    def set_flat_params(self, flat):
        flat = np.asarray(flat, np.float32)
        off = 0
        for p, s in zip(self.params, self.shapes):
            k = int(np.prod(s))
            p.copy_(torch.from_numpy(flat[off:off + k].reshape(s)))
            off += k

    def tables(self, graphs):
        """graphs: list of (adj [V,V] int, feat [V,F]) -> BatchTables."""
        return BatchTables([GraphTables(a, f, self.L, self.D) for a, f in graphs], self.L, self.C, self.device)

    # ---- one forward (+ backward) over a batch ---------------------------------------------------------------------
    def forward_backward(self, tb, targets=None):
        """Returns (graph_feature [G, C], loss [G] or None, flat parameter-gradient SUM over the batch or None)."""
        ctx, C, L = self.ctx, self.C, self.L
        H, W = self.params[0], self.params[-1]
        pre0 = tb.features @ H.t()                                   # MatMul(H, feature[v]) (SMP_beta.h:565-566)
        f_prev = torch.where(pre0 > 0, pre0, ALPHA * pre0).reshape(-1).contiguous()  # LeakyReLU3D on [1,1,C] (:571-572)
        saved = []
        Ks = [self.params[1 + 2 * l].t().contiguous() if self.k_transposed else self.params[1 + 2 * l] for l in range(L)]
        for l in range(L):
            K, b = Ks[l], self.params[2 + 2 * l]
            f_cur = torch.empty(tb.elems[l + 1], device=self.device)
            per_bucket = []
            for bk in tb.levels[l]:
                nm, B = bk["n_max"], bk["B"]
                T = ctx.promote_forward(f_prev, bk["f_off"], bk["m"], bk["pos"], nm, C, n=bk["n"])
                X = torch.zeros((B, nm, nm, 18 * C), device=self.device)  # padding rows must be zero for the grad-W product
                ctx.contract18_forward(T, bk["adj"].reshape(B, nm, nm), out=X, n=bk["n"])
                del T
                rows = B * nm * nm
                Y = torch.empty((rows, C), device=self.device)
                Z = f_cur[bk["offset"]:bk["offset"] + rows * C].view(rows, C)
                ctx.lib.ccn_mix_forward(ctx.h, X.data_ptr(), K.data_ptr(), b.data_ptr(), Y.data_ptr(), Z.data_ptr(), rows, 18 * C, C,
                                        ALPHA, ctx._stream(None))
                per_bucket.append((X, Y))
            saved.append(per_bucket)
            f_prev = f_cur
        s = torch.zeros((tb.Vtot, C), device=self.device)
        for bk in tb.levels[-1]:
            rows = bk["B"] * bk["n_max"] ** 2
            Zb = f_prev[bk["offset"]:bk["offset"] + rows * C].view(bk["B"], bk["n_max"] ** 2, C)
            s[bk["vertex"]] = (Zb * bk["rowmask"][:, :, None]).sum(1)  # ShrinkTensor (:623-625)
        vf = torch.where(s > 0, s, ALPHA * s)                        # LeakyReLU (:626-627)
        G = len(tb.graphs)
        gidx = torch.from_numpy(tb.graph_of).to(self.device)
        gf = torch.zeros((G, C), device=self.device).index_add_(0, gidx, vf)  # SumVectors (:629, 632)
        if targets is None:
            return gf, None, None
        t = torch.as_tensor(targets, dtype=torch.float32, device=self.device)
        pred = gf @ W                                                # InnerProduct (:634-635)
        loss = 0.5 * (pred - t) ** 2                                 # SquaredLoss (SquaredLoss.h:50-58)
        # ---- backward ------------------------------------------------------------------------------------------------
        grads = [torch.zeros_like(p) for p in self.params]
        dpred = pred - t
        grads[-1] += (dpred[:, None] * gf).sum(0)
        dvf = (dpred[:, None] * W[None, :])[gidx]
        ds = torch.where(s > 0, dvf, ALPHA * dvf)
        g_cur = torch.zeros(tb.elems[L], device=self.device)         # gradient of the level-L activations
        for bk in tb.levels[-1]:
            rows = bk["B"] * bk["n_max"] ** 2
            g_cur[bk["offset"]:bk["offset"] + rows * C] = (ds[bk["vertex"]][:, None, :] * bk["rowmask"][:, :, None]).reshape(-1)
        gKs = [torch.zeros_like(k) for k in Ks]  # [18 C, C] whatever the storage order of K_l
        for l in reversed(range(L)):
            K, b = Ks[l], self.params[2 + 2 * l]
            g_prev = torch.zeros(tb.elems[l], device=self.device)
            for bk, (X, Y) in zip(tb.levels[l], saved[l]):
                nm, B = bk["n_max"], bk["B"]
                rows = B * nm * nm
                gZ = g_cur[bk["offset"]:bk["offset"] + rows * C].view(rows, C)
                gX = torch.empty_like(X)
                ctx.mix_backward(X.reshape(rows, 18 * C), K, gZ, bias=b, Y=Y, gX=gX.reshape(rows, 18 * C), gW=gKs[l],
                                 gbias=grads[2 + 2 * l])
                gT = ctx.contract18_backward(gX, bk["adj"].reshape(B, nm, nm), n=bk["n"])
                del gX
                ctx.promote_backward(gT, bk["f_off"], bk["m"], bk["pos"], g_prev, n=bk["n"])
                del gT
            saved[l] = None
            grads[1 + 2 * l] = gKs[l].t().contiguous() if self.k_transposed else gKs[l]
            g_cur = g_prev
        gz0 = g_cur.view(-1, C)
        dpre0 = torch.where(pre0 > 0, gz0, ALPHA * gz0)
        grads[0] += dpre0.t() @ tb.features
        return gf, loss, torch.cat([g.reshape(-1) for g in grads])
