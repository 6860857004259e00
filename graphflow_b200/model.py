"""The second-order CCN models (SMP_beta, SMP_2D_ver8, SMP_omega_physics) forward + backward for a batch of graphs on the
B200 path.

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

    def __init__(self, graphs, n_levels, C, device, n_buckets=4, widths=None):
        widths = widths if widths is not None else [C] * (n_levels + 1)  # channels of the level-l activations
        self.graphs = graphs
        self.n_levels = n_levels
        self.Vtot = sum(g.V for g in graphs)
        self.graph_of = np.concatenate([np.full(g.V, i, np.int64) for i, g in enumerate(graphs)])
        self.graph_of_dev = torch.from_numpy(self.graph_of).to(device)
        base = np.cumsum([0] + [g.V for g in graphs])[:-1]
        self.features = torch.from_numpy(np.concatenate([g.features for g in graphs]).astype(np.float32)).to(device)
        self.levels = []
        prev_off = np.arange(self.Vtot, dtype=np.int64) * widths[0]  # level 0: one [1, 1, C_0] tensor per vertex
        self.elems = [self.Vtot * widths[0]]                           # flat activation buffer sizes per level (0..L)
        for l in range(n_levels):
            C = widths[l + 1]  # channels of this level's output (the instance stride of its activation buffer)
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


class CCNModelB200:
    """kind = "beta"  : SMP_beta (SMP_beta.h)
       kind = "ver8"  : SMP_2D_ver8 (SMP_2D_ver8.h): the same wiring with the feature mix done by CustomMatMulTensor, i.e. K_l
                        stored [C, 18 C] (SMP_2D_ver8.h:130, 526-527)
       kind = "omega" : SMP_omega_physics (SMP_omega_physics.h): raw vertex features, receptive fields limited to `max_field`
                        members (:367-418), the channel width halves per level (:142-146), every level feeds the read-out,
                        which ends in a hidden layer (:560-595).
    Parameters are the reference's, in its optimizer registration order: H, (K_l, b_l) for l = 1..L, then W (beta / ver8)
    or W1 [Ctot/2, Ctot], W2 [Ctot/2] (omega)."""

    def __init__(self, kind, n_levels, C, n_features, n_depth=None, max_field=None, device=0, ctx=None):
        assert kind in ("beta", "ver8", "omega", "omega_wl")
        self.kind, self.L, self.C, self.F, self.D, self.max_field = kind, n_levels, C, n_features, n_depth, max_field
        self.k_transposed = kind == "ver8"
        self.cache_workspaces = True  # keep the zero-padded contraction outputs with the batch tables between steps
        self.device = torch.device("cuda", device)
        self.ctx = ctx if ctx is not None else Context(device)
        w = [C]
        for _ in range(n_levels):
            w.append(max(1, w[-1] // 2) if kind == "omega" else C)
        self.widths = w
        fin = n_features if kind == "omega" else n_features * (n_depth + 1)
        shapes = [(C, fin)]
        for l in range(1, n_levels + 1):
            shapes += [(w[l], 18 * w[l - 1]) if self.k_transposed else (18 * w[l - 1], w[l]), (w[l],)]
        if kind == "omega":
            tot = sum(w)
            shapes += [(tot // 2, tot), (tot // 2,)]
        else:
            shapes += [(C,)]
        self.shapes = shapes
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

    def get_flat_params(self):
        """The parameters as ONE flat device tensor in registration order (a copy)."""
        return torch.cat([p.reshape(-1) for p in self.params])

    def set_flat_params_device(self, flat):
        off = 0
        for p in self.params:
            p.copy_(flat[off:off + p.numel()].reshape(p.shape))
            off += p.numel()

    def save_model(self, path):
        """The reference's text checkpoint (SMP_beta.h:980-990): readable by the reference's `load_model`."""
        from . import checkpoint

        checkpoint.save_model(path, self.get_flat_params())

    def load_model(self, path):
        """Reads a checkpoint written by the reference's `save_model` (or ours) (SMP_beta.h:992-1002)."""
        from . import checkpoint

        self.set_flat_params(checkpoint.load_model(path, self.num_params()))

    # ---- whole-step CUDA graph ----------------------------------------------------------------------------------------
    def capture_step(self, tb, targets):
        """Captures forward_backward(tb, targets) -- some hundred kernel launches across the size buckets and levels -- in ONE
        CUDA graph.  Returns a callable; each call replays the step and returns the same (graph_feature, loss, grads)
        tensors, refreshed.  Parameters are read from `self.params` at replay time (update them in place, e.g. with
        `set_flat_params_device` or an optimizer step); `targets` is copied to a device tensor that the graph keeps reading
        (`step.targets.copy_(...)` to change it).  The tables `tb` must not change.

        The graph bakes the context's scratch pointers in, so the context is FROZEN after the warm-up (ccn_ctx_set_frozen):
        a later call that would have to grow a context buffer (a bigger batch, another operator) fails loudly with
        CCN_ERR_UNSUPPORTED instead of reallocating under the graph.  `step.release()` drops the graph and unfreezes."""
        tg = torch.as_tensor(targets, dtype=torch.float32, device=self.device).clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                                # warm-up: workspaces, cached buffers, first-use allocations
            for _ in range(2):
                self.forward_backward(tb, tg)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self.ctx.set_frozen(True)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward_backward(tb, tg)

        def step():
            if step.graph is None:
                raise RuntimeError("this captured step was released")
            step.graph.replay()
            return out

        def release():
            step.graph = None
            self.ctx.set_frozen(False)

        step.graph, step.targets, step.outputs, step.release = graph, tg, out, release
        return step

    # ---- the reference models' training API (SMP_beta.h:745-772, 871-879) -------------------------------------------
    def getLoss(self, graphs, targets, tb=None):
        """Summed loss of the batch at the current parameters (SMP_beta.h:640-649)."""
        tb = tb if tb is not None else self.tables(graphs)
        return self.forward_backward(tb, targets, need_grads=False)[1].sum().item()

    def BatchLearn(self, graphs, targets, learning_rate, tb=None):
        """One `BatchLearn(nBatch, molecule, target, learning_rate)`: the gradients summed over the batch, then the
        reference's `Adam::Learn(learning_rate, nBatch)` on the device.  Returns (loss before, loss after) like the
        reference; pass `tb` (from `tables`) to reuse the graph tables across epochs."""
        from . import optim

        tb = tb if tb is not None else self.tables(graphs)
        _, loss, g = self.forward_backward(tb, targets)
        before = loss.sum().item()
        if getattr(self, "_adam", None) is None:
            self._adam = optim.Adam(self.ctx, self.get_flat_params())
        else:
            self._adam.params.copy_(self.get_flat_params())          # set_flat_params / load_model may have intervened
        self._adam.learn(g.contiguous(), learning_rate, len(graphs))
        self.set_flat_params_device(self._adam.params)
        return before, self.getLoss(graphs, targets, tb)

    def Feature(self, graph):
        """`Feature(molecule)`: the graph-level feature vector of one graph (adj, feat), as a numpy array."""
        return self.forward_backward(self.tables([graph]), None)[0][0].cpu().numpy()

    def Predict(self, graph):
        """`Predict(molecule)`: the model output for one graph (adj, feat) (SMP_beta.h:871-879)."""
        self.forward_backward(self.tables([graph]), None)
        return float(self.last_pred[0].item())

    def tables(self, graphs):
        """graphs: list of (adj [V,V] int, feat [V,F]) -> BatchTables."""
        kind = self.kind if self.kind in ("omega", "omega_wl") else "beta"
        gts = [GraphTables(a, f, self.L, self.D, kind=kind, max_field=self.max_field) for a, f in graphs]
        return BatchTables(gts, self.L, self.C, self.device, widths=self.widths)

    def _shrink(self, tb, l, buf):
        """ShrinkTensor of every vertex's level-l activation: [Vtot, C_l] (sum over the n x n cells that really exist)."""
        C = self.widths[l]
        if l == 0:
            return buf.view(tb.Vtot, C)
        s = torch.zeros((tb.Vtot, C), device=self.device)
        for bk in tb.levels[l - 1]:
            rows = bk["B"] * bk["n_max"] ** 2
            Zb = buf[bk["offset"]:bk["offset"] + rows * C].view(bk["B"], bk["n_max"] ** 2, C)
            s[bk["vertex"]] = (Zb * bk["rowmask"][:, :, None]).sum(1)
        return s

    def _unshrink_add(self, tb, l, ds, g):
        """Transpose of _shrink: add ds [Vtot, C_l] to every existing cell of the level-l activation gradient g (flat)."""
        C = self.widths[l]
        if l == 0:
            g += ds.reshape(-1)
            return
        for bk in tb.levels[l - 1]:
            rows = bk["B"] * bk["n_max"] ** 2
            view = g[bk["offset"]:bk["offset"] + rows * C].view(bk["B"], bk["n_max"] ** 2, C)
            view += ds[bk["vertex"]][:, None, :] * bk["rowmask"][:, :, None]

    # ---- one forward (+ backward) over a batch ---------------------------------------------------------------------
    def forward_backward(self, tb, targets=None, need_grads=True):
        """Returns (graph_feature [G, Ctot], loss [G] or None, flat parameter-gradient SUM over the batch or None)."""
        st = self._trunk_forward(tb)
        gf = st["gf"]
        if self.kind == "omega":
            W1, W2 = self.params[-2], self.params[-1]
            hid = gf @ W1.t()                                        # MatVecMul (SMP_omega_physics.h:585-586)
            ha = torch.where(hid > 0, hid, ALPHA * hid)
            pred = ha @ W2
        else:
            pred = gf @ self.params[-1]                              # InnerProduct (SMP_beta.h:634-635)
        self.last_pred = pred
        if targets is None:
            return gf, None, None
        t = torch.as_tensor(targets, dtype=torch.float32, device=self.device)
        loss = 0.5 * (pred - t) ** 2                                 # SquaredLoss (SquaredLoss.h:50-58)
        if not need_grads:
            return gf, loss, None
        grads = [torch.zeros_like(p) for p in self.params]
        dpred = pred - t
        if self.kind == "omega":
            grads[-1] += (dpred[:, None] * ha).sum(0)
            dha = dpred[:, None] * W2[None, :]
            dh = torch.where(hid > 0, dha, ALPHA * dha)
            grads[-2] += dh.t() @ gf
            dgf = dh @ W1
        else:
            grads[-1] += (dpred[:, None] * gf).sum(0)
            dgf = dpred[:, None] * self.params[-1][None, :]
        self._trunk_backward(tb, st, dgf, grads)
        return gf, loss, torch.cat([g.reshape(-1) for g in grads])

    def _trunk_forward(self, tb):
        """Level 0, the L contraction levels and the per-level read-out (ShrinkTensor -> LeakyReLU -> SumVectors -> concatenation).
        Returns the state `_trunk_backward` needs; st["gf"] is the graph feature [G, Ctot]."""
        ctx, L, w = self.ctx, self.L, self.widths
        H = self.params[0]
        # level 0 on our own mix kernels: MatMul(H, feature[v]) (SMP_beta.h:565-566) for all vertices at once is
        # features [Vtot, F'] . H^T [F', C]; LeakyReLU3D on the [1,1,C] tensors (:571-572) is the fused epilogue with a zero bias
        Ht = H.t().contiguous()
        zero_b0 = torch.zeros(w[0], device=self.device)
        pre0, act0 = ctx.mix_forward(tb.features, Ht, zero_b0)
        acts = [act0.reshape(-1)]
        Ks = [self.params[1 + 2 * l].t().contiguous() if self.k_transposed else self.params[1 + 2 * l] for l in range(L)]
        saved = []
        for l in range(L):
            K, b, Ci, Co = Ks[l], self.params[2 + 2 * l], w[l], w[l + 1]
            f_cur = torch.empty(tb.elems[l + 1], device=self.device)
            per_bucket = []
            for bk in tb.levels[l]:
                nm, B = bk["n_max"], bk["B"]
                # promotion + stack + contraction in one launch where the fused kernels apply (no stacked T in HBM)
                # The contraction writes the n_i^2 real rows of every instance; the padding rows up to n_max^2 must read as
                # zero in the grad-W product.  The buffer is zeroed once and kept with the bucket: later steps only rewrite
                # the real rows (saves a memset of the whole [rows, 18 C] block per step: 15 GB per step at config 3).
                X = bk.get("X") if self.cache_workspaces else None
                if X is None:
                    X = torch.zeros((B, nm, nm, 18 * Ci), device=self.device)
                    if self.cache_workspaces:
                        bk["X"] = X
                ctx.gather_contract18_forward(acts[l], bk["f_off"], bk["m"], bk["pos"], bk["adj"].reshape(B, nm, nm), nm, Ci, out=X,
                                              n=bk["n"])
                rows = B * nm * nm
                Y = torch.empty((rows, Co), device=self.device)
                Z = f_cur[bk["offset"]:bk["offset"] + rows * Co].view(rows, Co)
                ctx._rc(ctx.lib.ccn_mix_forward(ctx.h, X.data_ptr(), K.data_ptr(), b.data_ptr(), Y.data_ptr(), Z.data_ptr(), rows,
                                                18 * Ci, Co, ALPHA, ctx._stream(None)))
                per_bucket.append((X, Y))
            saved.append(per_bucket)
            acts.append(f_cur)
        # ---- read-out ------------------------------------------------------------------------------------------------
        G = len(tb.graphs)
        gidx = tb.graph_of_dev
        levels_out = list(range(L + 1)) if self.kind == "omega" else [L]
        s = {l: self._shrink(tb, l, acts[l]) for l in levels_out}                          # ShrinkTensor
        vf = {l: torch.where(s[l] > 0, s[l], ALPHA * s[l]) for l in levels_out}            # LeakyReLU
        lf = [torch.zeros((G, w[l]), device=self.device).index_add_(0, gidx, vf[l]) for l in levels_out]  # SumVectors
        return {"gf": torch.cat(lf, 1), "lf": lf, "s": s, "levels_out": levels_out, "Ks": Ks, "saved": saved, "Ht": Ht,
                "zero_b0": zero_b0, "pre0": pre0}

    def _trunk_backward(self, tb, st, dgf, grads):
        """Transpose of `_trunk_forward` from dgf [G, Ctot]: adds the gradients of H and (K_l, b_l) into grads[0 .. 2L]."""
        ctx, L, w = self.ctx, self.L, self.widths
        s, levels_out, Ks, saved, Ht, zero_b0, pre0 = (st[k] for k in ("s", "levels_out", "Ks", "saved", "Ht", "zero_b0", "pre0"))
        gidx = tb.graph_of_dev
        ds, off = {}, 0
        for l in levels_out:
            dvf = dgf[:, off:off + w[l]][gidx]
            ds[l] = torch.where(s[l] > 0, dvf, ALPHA * dvf)
            off += w[l]
        g_cur = torch.zeros(tb.elems[L], device=self.device)         # gradient of the level-L activations
        self._unshrink_add(tb, L, ds[L], g_cur)
        gKs = [torch.zeros_like(k) for k in Ks]                      # [18 C_in, C_out] whatever the storage order of K_l
        for l in reversed(range(L)):
            K, b, Ci, Co = Ks[l], self.params[2 + 2 * l], w[l], w[l + 1]
            g_prev = torch.zeros(tb.elems[l], device=self.device)
            if l in ds:                                              # omega: level l also feeds the read-out directly
                self._unshrink_add(tb, l, ds[l], g_prev)
            for bk, (X, Y) in zip(tb.levels[l], saved[l]):
                nm, B = bk["n_max"], bk["B"]
                rows = B * nm * nm
                gZ = g_cur[bk["offset"]:bk["offset"] + rows * Co].view(rows, Co)
                gX = torch.empty_like(X)
                ctx.mix_backward(X.reshape(rows, 18 * Ci), K, gZ, bias=b, Y=Y, gX=gX.reshape(rows, 18 * Ci), gW=gKs[l],
                                 gbias=grads[2 + 2 * l])
                ctx.gather_contract18_backward(gX, bk["adj"].reshape(B, nm, nm), bk["f_off"], bk["m"], bk["pos"], g_prev, n=bk["n"])
                del gX
            saved[l] = None  # (cached X buffers stay referenced by their bucket)
            grads[1 + 2 * l] += gKs[l].t().contiguous() if self.k_transposed else gKs[l]
            g_cur = g_prev
        gHt = torch.zeros_like(Ht)
        ctx.mix_backward(tb.features, Ht, g_cur.view(-1, w[0]), bias=zero_b0, Y=pre0, gW=gHt, gbias=torch.zeros_like(zero_b0), need_gX=False)
        grads[0] += gHt.t()


class PairGraphsModelB200:
    """SMP_omega_pairgraphs (SMP_omega_pairgraphs.h): the path run twice -- one SMP_omega_physics-style trunk on the graph and one
    on its line graph (separate H, K_l, b_l; `computation_graph_` :147-281, `complete_computation_graph_` :565-655) -- the level
    features of both concatenated level by level (:705-710) and a two-hidden-layer head (MatVecMul + LeakyReLU twice, InnerProduct,
    SquaredLoss, :714-729).  Parameters in the reference's registration order (:365-377): H_1, H_2, then per level K1_l, b1_l,
    K2_l, b2_l, then W1 [max(Ctot/2, 10), Ctot], W2 [max(h1/2, 10), h1], W3 [h2]."""

    def __init__(self, n_levels, C, n_features_1, n_features_2, max_field, device=0, ctx=None):
        self.ctx = ctx if ctx is not None else Context(device)
        self.t1 = CCNModelB200("omega", n_levels, C, n_features_1, max_field=max_field, device=device, ctx=self.ctx)
        self.t2 = CCNModelB200("omega", n_levels, C, n_features_2, max_field=max_field, device=device, ctx=self.ctx)
        self.L, self.device = n_levels, self.t1.device
        self.widths = self.t1.widths
        tot = 2 * sum(self.widths)
        h1 = max(tot // 2, 10)
        h2 = max(h1 // 2, 10)
        self.head = [torch.zeros(s, device=self.device) for s in ((h1, tot), (h2, h1), (h2,))]
        # views of the trunks' parameters in registration order (the trunks' own head parameters are unused)
        self.params = [self.t1.params[0], self.t2.params[0]]
        for l in range(n_levels):
            self.params += [self.t1.params[1 + 2 * l], self.t1.params[2 + 2 * l], self.t2.params[1 + 2 * l], self.t2.params[2 + 2 * l]]
        self.params += self.head

    def num_params(self):
        return int(sum(p.numel() for p in self.params))

    def set_flat_params(self, flat):
        flat = np.asarray(flat, np.float32)
        off = 0
        for p in self.params:
            p.copy_(torch.from_numpy(flat[off:off + p.numel()].reshape(tuple(p.shape))))
            off += p.numel()

    def get_flat_params(self):
        return torch.cat([p.reshape(-1) for p in self.params])

    def tables(self, pairs):
        """pairs: list of ((adj_1, feat_1), (adj_2, feat_2)) -> (BatchTables of the graphs, BatchTables of the line graphs)."""
        return self.t1.tables([p[0] for p in pairs]), self.t2.tables([p[1] for p in pairs])

    def forward_backward(self, tbs, targets=None, need_grads=True):
        """Returns (graph_feature [G, Ctot], loss [G] or None, flat parameter-gradient SUM over the batch or None)."""
        L, w = self.L, self.widths
        s1, s2 = self.t1._trunk_forward(tbs[0]), self.t2._trunk_forward(tbs[1])
        gf = torch.cat([x for l in range(L + 1) for x in (s1["lf"][l], s2["lf"][l])], 1)   # ConcatVectors, level by level
        W1, W2, W3 = self.head
        lrelu = lambda x: torch.where(x > 0, x, ALPHA * x)  # noqa: E731
        h1 = gf @ W1.t()
        a1 = lrelu(h1)
        h2 = a1 @ W2.t()
        a2 = lrelu(h2)
        pred = a2 @ W3
        self.last_pred = pred
        if targets is None:
            return gf, None, None
        t = torch.as_tensor(targets, dtype=torch.float32, device=self.device)
        loss = 0.5 * (pred - t) ** 2
        if not need_grads:
            return gf, loss, None
        dpred = pred - t
        gW3 = (dpred[:, None] * a2).sum(0)
        da2 = dpred[:, None] * W3[None, :]
        dh2 = torch.where(h2 > 0, da2, ALPHA * da2)
        gW2 = dh2.t() @ a1
        da1 = dh2 @ W2
        dh1 = torch.where(h1 > 0, da1, ALPHA * da1)
        gW1 = dh1.t() @ gf
        dgf = dh1 @ W1
        # split dgf back into the two trunks' level-major layouts
        d1, d2, off = [], [], 0
        for l in range(L + 1):
            d1.append(dgf[:, off:off + w[l]])
            d2.append(dgf[:, off + w[l]:off + 2 * w[l]])
            off += 2 * w[l]
        g1 = [torch.zeros_like(p) for p in self.t1.params]
        g2 = [torch.zeros_like(p) for p in self.t2.params]
        self.t1._trunk_backward(tbs[0], s1, torch.cat(d1, 1).contiguous(), g1)
        self.t2._trunk_backward(tbs[1], s2, torch.cat(d2, 1).contiguous(), g2)
        grads = [g1[0], g2[0]]
        for l in range(L):
            grads += [g1[1 + 2 * l], g1[2 + 2 * l], g2[1 + 2 * l], g2[2 + 2 * l]]
        grads += [gW1, gW2, gW3]
        return gf, loss, torch.cat([g.reshape(-1) for g in grads])

    def getLoss(self, pairs, targets, tbs=None):
        tbs = tbs if tbs is not None else self.tables(pairs)
        return self.forward_backward(tbs, targets, need_grads=False)[1].sum().item()

    def Predict(self, pair):
        """`Predict(molecule_1, molecule_2)` (SMP_omega_pairgraphs.h:1051-1080)."""
        self.forward_backward(self.tables([pair]), None)
        return float(self.last_pred[0].item())

    def BatchLearn(self, pairs, targets, learning_rate, tbs=None):
        """`BatchLearn(nBatch, molecule_1, molecule_2, target, learning_rate)` (:865-895): summed gradients, Adam, (loss before, after)."""
        from . import optim

        tbs = tbs if tbs is not None else self.tables(pairs)
        _, loss, g = self.forward_backward(tbs, targets)
        before = loss.sum().item()
        if getattr(self, "_adam", None) is None:
            self._adam = optim.Adam(self.ctx, self.get_flat_params())
        else:
            self._adam.params.copy_(self.get_flat_params())
        self._adam.learn(g.contiguous(), learning_rate, len(pairs))
        off = 0
        for p in self.params:
            p.copy_(self._adam.params[off:off + p.numel()].reshape(p.shape))
            off += p.numel()
        return before, self.getLoss(pairs, targets, tbs)


class SMPBetaB200(CCNModelB200):
    """SMP_beta (k_transposed=False) or SMP_2D_ver8 (k_transposed=True); kept for the earlier call sites."""

    def __init__(self, n_levels, C, n_features, n_depth, device=0, ctx=None, k_transposed=False):
        super().__init__("ver8" if k_transposed else "beta", n_levels, C, n_features, n_depth=n_depth, device=device, ctx=ctx)
