"""Thin torch-facing wrapper over the C-ABI.  torch supplies device memory and streams only; every number is
produced by the CUDA kernels behind include/ccn_b200.h.  No fallback paths."""
import ctypes

import torch

from . import _lib
from ._lib import ADJ_POSITIVE_PART, ADJ_RAW, CCNError  # noqa: F401

NUM_CONTRACTIONS = 18  # RisiContraction_18_gpu::nContractions (RisiContraction_18_gpu.h:1749)


_EMPTY = {}  # per device: a small live allocation whose address stands in for empty tensors


def _ptr(t):
    """Device / host pointer of a tensor (None -> NULL).  An EMPTY tensor has no storage (data_ptr() == 0), but the C-ABI
    requires valid pointers even for an empty batch (it then returns before touching them), so empty tensors map to a
    small live buffer on the same device."""
    if t is None:
        return None
    if t.numel() == 0:
        key = str(t.device)
        if key not in _EMPTY:
            _EMPTY[key] = torch.zeros(64, dtype=torch.float32, device=t.device)
        return ctypes.c_void_p(_EMPTY[key].data_ptr())
    return ctypes.c_void_p(t.data_ptr())


def _check(t, name, device=None, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    if device is not None and t.device != device:
        raise ValueError("%s must live on %s, got %s" % (name, device, t.device))
    return t


def _mask_bits(keep_mask):
    if keep_mask is None:
        return (1 << 64) - 1
    if isinstance(keep_mask, int):
        return keep_mask & ((1 << 64) - 1)
    bits = 0
    for k, use in enumerate(keep_mask):
        if use:
            bits |= 1 << k
    return bits


class Context:
    """One ccn_ctx: owns the scratch workspace and staging buffers.  Use one per host thread (and per GPU)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise CCNError("no CUDA device: graphflow_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        h = ctypes.c_void_p()
        rc = self.lib.ccn_ctx_create(ctypes.byref(h), device)
        if rc != 0:
            raise CCNError("ccn_ctx_create failed: %s" % self.lib.ccn_status_string(rc).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.ccn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rc(self, rc):
        if rc != 0:
            raise CCNError("%s: %s" % (self.lib.ccn_status_string(rc).decode(), self.lib.ccn_last_error(self.h).decode()))

    def _stream(self, stream):
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    @property
    def kernel_launches(self):
        return int(self.lib.ccn_ctx_kernel_launches(self.h))

    def set_workspace_limit(self, nbytes):
        self._rc(self.lib.ccn_ctx_set_workspace_limit(self.h, int(nbytes)))

    def set_kernel_path(self, path):
        """_lib.PATH_AUTO (fused kernels when the shape allows) or PATH_GENERIC."""
        self._rc(self.lib.ccn_ctx_set_kernel_path(self.h, int(path)))
        self._kernel_path = int(path)

    def set_mix_path(self, path):
        """_lib.MIX_AUTO (tcgen05 3xTF32 when the shape allows), MIX_SIMT (fp32 CUDA cores) or MIX_TENSOR."""
        self._rc(self.lib.ccn_ctx_set_mix_path(self.h, int(path)))

    def set_phase_trace(self, trace):
        """trace: int64 cuda tensor [tiles, 8] (or None to switch off); see ccn_ctx_set_phase_trace."""
        if trace is None:
            self._rc(self.lib.ccn_ctx_set_phase_trace(self.h, None, 0))
        else:
            _check(trace, "trace", self.device, torch.int64)
            self._rc(self.lib.ccn_ctx_set_phase_trace(self.h, _ptr(trace), trace.numel() // 8))
        self._trace = trace

    def set_frozen(self, flag):
        """While frozen no context buffer may grow (calls that would need to fail with CCN_ERR_UNSUPPORTED): set while a
        CUDA graph that captured this context's launches is alive."""
        self._rc(self.lib.ccn_ctx_set_frozen(self.h, int(bool(flag))))

    def fused_error_flag(self):
        """Synchronises, returns the sticky sibling-timeout flag and clears it (0 = never happened)."""
        flag = ctypes.c_int()
        self._rc(self.lib.ccn_ctx_fused_error_flag(self.h, ctypes.byref(flag)))
        return flag.value

    def set_kernel_timing(self, flag):
        """Bracket every kernel launch with CUDA events on the launching stream (clears earlier totals)."""
        self._rc(self.lib.ccn_ctx_set_kernel_timing(self.h, int(bool(flag))))

    def kernel_timing(self):
        """{kernel name: (total device ms, launches)} since set_kernel_timing(True)."""
        res = {}
        for k in range(self.lib.ccn_num_kernels()):
            ms, cnt = ctypes.c_double(), ctypes.c_int64()
            self._rc(self.lib.ccn_ctx_get_kernel_timing(self.h, k, ctypes.byref(ms), ctypes.byref(cnt)))
            if cnt.value:
                res[self.lib.ccn_kernel_name(k).decode()] = (ms.value, cnt.value)
        return res

    # ---- StackTensor3D + RisiContraction_18 -----------------------------------------------------------------------
    def contract18_forward(self, T, adj, out=None, n=None, slabs=None, n_max=None, C=None, batch=None, strides=None,
                           adj_mode=ADJ_POSITIVE_PART, stream=None):
        """T: [B, N, N, N, C] (or any flat buffer with explicit n_max/C/batch/strides); adj: [B, N, N];
        n: optional int32 [B] per-instance sizes (ragged); slabs: optional int64 [B*n_max] table of slab pointers
        (then T must be None).  Returns out [B, N, N, 18*C]."""
        dev = self.device
        adj = _check(adj, "adj", dev)
        if T is not None:
            T = _check(T, "T", dev)
            if n_max is None:
                batch, n_max, C = T.shape[0], T.shape[1], T.shape[4]
        if slabs is not None:
            slabs = _check(slabs, "slabs", dev, torch.int64)
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        sT, sA, sO = strides if strides is not None else (n_max ** 3 * C, n_max * n_max, n_max * n_max * 18 * C)
        if out is None:
            out = torch.empty((batch, n_max, n_max, NUM_CONTRACTIONS * C), device=dev, dtype=torch.float32)
        _check(out, "out", dev)
        self._rc(self.lib.ccn_contract18_forward(self.h, _ptr(T), _ptr(slabs), _ptr(adj), _ptr(out), _ptr(n), n_max, C,
                                                 batch, sT, sA, sO, adj_mode, self._stream(stream)))
        return out

    def contract18_backward(self, gout, adj, gT=None, n=None, gslabs=None, n_max=None, C=None, batch=None,
                            strides=None, adj_mode=ADJ_POSITIVE_PART, beta=0.0, stream=None):
        """gout: [B, N, N, 18*C]; returns gT [B, N, N, N, C] = beta*gT + contraction^T(gout)."""
        dev = self.device
        gout = _check(gout, "gout", dev)
        adj = _check(adj, "adj", dev)
        if n_max is None:
            batch, n_max, C = gout.shape[0], gout.shape[1], gout.shape[3] // NUM_CONTRACTIONS
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if gslabs is not None:
            gslabs = _check(gslabs, "gslabs", dev, torch.int64)
        elif gT is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an existing gT")
            gT = torch.empty((batch, n_max, n_max, n_max, C), device=dev, dtype=torch.float32)
        if gT is not None:
            _check(gT, "gT", dev)
        sG, sA, sT = strides if strides is not None else (n_max * n_max * 18 * C, n_max * n_max, n_max ** 3 * C)
        self._rc(self.lib.ccn_contract18_backward(self.h, _ptr(gout), _ptr(adj), _ptr(gT), _ptr(gslabs), _ptr(n), n_max,
                                                  C, batch, sG, sA, sT, adj_mode, beta, self._stream(stream)))
        return gT

    # ---- StackTensor3D + RisiContraction_50 -----------------------------------------------------------------------
    def contract50_forward(self, T, adj, out=None, n=None, adj_mode=ADJ_RAW, stream=None):
        """T: [B, N, N, N, C]; adj: [B, N, N] -> out [B, N, N, 50*C] (RisiContraction_50, raw adjacency)."""
        dev = self.device
        T, adj = _check(T, "T", dev), _check(adj, "adj", dev)
        B, N, C = T.shape[0], T.shape[1], T.shape[4]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if out is None:
            out = torch.empty((B, N, N, 50 * C), device=dev, dtype=torch.float32)
        _check(out, "out", dev)
        self._rc(self.lib.ccn_contract50_forward(self.h, _ptr(T), None, _ptr(adj), _ptr(out), _ptr(n), N, C, B,
                                                 N ** 3 * C, N * N, N * N * 50 * C, adj_mode, self._stream(stream)))
        return out

    def contract50_backward(self, gout, adj, gT=None, n=None, adj_mode=ADJ_RAW, beta=0.0, stream=None):
        """gout: [B, N, N, 50*C] -> gT [B, N, N, N, C] = beta*gT + contraction^T(gout)."""
        dev = self.device
        gout, adj = _check(gout, "gout", dev), _check(adj, "adj", dev)
        B, N, C = gout.shape[0], gout.shape[1], gout.shape[3] // 50
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if gT is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an existing gT")
            gT = torch.empty((B, N, N, N, C), device=dev, dtype=torch.float32)
        _check(gT, "gT", dev)
        self._rc(self.lib.ccn_contract50_backward(self.h, _ptr(gout), _ptr(adj), _ptr(gT), None, _ptr(n), N, C, B,
                                                  N * N * 50 * C, N * N, N ** 3 * C, adj_mode, beta,
                                                  self._stream(stream)))
        return gT

    # ---- the other members of the contraction family (RisiContraction_4 / _10 / _18_dropout) ------------------------
    @staticmethod
    def _family_adj_mode(variant, adj_mode):
        if adj_mode is not None:
            return adj_mode
        return ADJ_POSITIVE_PART if variant == 18 else ADJ_RAW  # `adj_value > 0` guard only in the 18-way operators

    def contract_family_forward(self, variant, T, adj=None, keep_mask=None, out=None, n=None, adj_mode=None, out_scale=1.0,
                                stream=None):
        """variant 4 / 10 / 18 / 50 -> out [B, N, N, variant*C].  keep_mask: iterable of bools (use[] of
        RisiContraction_18_dropout) or an int bit mask; dropped slabs are written as zeros.  adj may be None for 4.
        out_scale: nKept/18 for the dropout operator's test mode (all slabs kept)."""
        dev = self.device
        T = _check(T, "T", dev)
        if adj is not None:
            adj = _check(adj, "adj", dev)
        B, N, C = T.shape[0], T.shape[1], T.shape[4]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if out is None:
            out = torch.empty((B, N, N, variant * C), device=dev, dtype=torch.float32)
        _check(out, "out", dev)
        self._rc(self.lib.ccn_contract_family_forward(self.h, variant, _mask_bits(keep_mask), _ptr(T), None, _ptr(adj), _ptr(out),
                                                      _ptr(n), N, C, B, N ** 3 * C, N * N, N * N * variant * C,
                                                      self._family_adj_mode(variant, adj_mode), out_scale,
                                                      self._stream(stream)))
        return out

    def contract_family_backward(self, variant, gout, adj=None, keep_mask=None, gT=None, n=None, adj_mode=None, beta=0.0,
                                 stream=None):
        """gout: [B, N, N, variant*C] -> gT [B, N, N, N, C] = beta*gT + contraction^T(gout) (dropped slabs ignored)."""
        dev = self.device
        gout = _check(gout, "gout", dev)
        if adj is not None:
            adj = _check(adj, "adj", dev)
        B, N, C = gout.shape[0], gout.shape[1], gout.shape[3] // variant
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if gT is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an existing gT")
            gT = torch.empty((B, N, N, N, C), device=dev, dtype=torch.float32)
        _check(gT, "gT", dev)
        self._rc(self.lib.ccn_contract_family_backward(self.h, variant, _mask_bits(keep_mask), _ptr(gout), _ptr(adj), _ptr(gT), None,
                                                       _ptr(n), N, C, B, N * N * variant * C, N * N, N ** 3 * C,
                                                       self._family_adj_mode(variant, adj_mode), beta, self._stream(stream)))
        return gT

    # ---- parameter updates (the reference's Adam / Momentum / SGD on the flat parameter vector) ----------------------
    def adam_step(self, params, grads, m, v, alpha, n_batch=1, updates_before=0, per_element_bias=True, beta1=0.9,
                  beta2=0.999, epsilon=1e-8, stream=None):
        dev = self.device
        params, grads, m, v = (_check(t, nm, dev) for t, nm in ((params, "params"), (grads, "grads"), (m, "m"), (v, "v")))
        count = params.numel()
        if not (grads.numel() == m.numel() == v.numel() == count):
            raise ValueError("params, grads, m and v must have the same number of elements")
        self._rc(self.lib.ccn_adam_step(self.h, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), count, alpha, beta1, beta2,
                                        epsilon, int(n_batch), int(updates_before), int(bool(per_element_bias)),
                                        self._stream(stream)))
        return params

    def momentum_step(self, params, grads, moments, learning_rate, gamma=0.9, n_batch=1, stream=None):
        dev = self.device
        params, grads, moments = (_check(t, nm, dev) for t, nm in ((params, "params"), (grads, "grads"), (moments, "moments")))
        count = params.numel()
        if not (grads.numel() == moments.numel() == count):
            raise ValueError("params, grads and moments must have the same number of elements")
        self._rc(self.lib.ccn_momentum_step(self.h, _ptr(params), _ptr(grads), _ptr(moments), count, learning_rate, gamma,
                                            int(n_batch), self._stream(stream)))
        return params

    # ---- host-buffer variants (what a reference op with host value[]/gradient[] arrays calls) -----------------------
    def contract18_forward_host(self, T, adj, out=None, adj_mode=ADJ_POSITIVE_PART):
        cpu = torch.device("cpu")
        T, adj = _check(T, "T", cpu), _check(adj, "adj", cpu)
        B, N, C = T.shape[0], T.shape[1], T.shape[4]
        if out is None:
            out = torch.empty((B, N, N, NUM_CONTRACTIONS * C), dtype=torch.float32, pin_memory=True)
        self._rc(self.lib.ccn_contract18_forward_host(self.h, _ptr(T), _ptr(adj), _ptr(_check(out, "out", cpu)), N, C,
                                                      B, adj_mode))
        return out

    def contract18_backward_host(self, gout, adj, gT=None, adj_mode=ADJ_POSITIVE_PART, beta=0.0):
        cpu = torch.device("cpu")
        gout, adj = _check(gout, "gout", cpu), _check(adj, "adj", cpu)
        B, N, C = gout.shape[0], gout.shape[1], gout.shape[3] // NUM_CONTRACTIONS
        if gT is None:
            gT = torch.empty((B, N, N, N, C), dtype=torch.float32, pin_memory=True)
        self._rc(self.lib.ccn_contract18_backward_host(self.h, _ptr(gout), _ptr(adj), _ptr(_check(gT, "gT", cpu)), N, C,
                                                       B, adj_mode, beta))
        return gT

    def contract18_forward_backward_host(self, T, adj, gout, out, gT, adj_mode=ADJ_POSITIVE_PART):
        cpu = torch.device("cpu")
        for t, nm in ((T, "T"), (adj, "adj"), (gout, "gout"), (out, "out"), (gT, "gT")):
            _check(t, nm, cpu)
        B, N, C = T.shape[0], T.shape[1], T.shape[4]
        self._rc(self.lib.ccn_contract18_forward_backward_host(self.h, _ptr(T), _ptr(adj), _ptr(gout), _ptr(out),
                                                               _ptr(gT), N, C, B, adj_mode))
        return out, gT

    # ---- promotion (MatTensorMul + TensorMatMul + StackTensor3D as a gather) ------------------------------------------
    def promote_forward(self, f, f_off, m, pos, n_max, C, T=None, n=None, stream=None):
        """f: flat float32 buffer holding the level l-1 tensors; f_off int64 [B*n_max], m int32 [B*n_max],
        pos int32 [B*n_max*n_max] (see include/ccn_b200.h).  Returns the stacked T [B, n_max, n_max, n_max, C]."""
        dev = self.device
        f = _check(f, "f", dev)
        f_off, m, pos = _check(f_off, "f_off", dev, torch.int64), _check(m, "m", dev, torch.int32), _check(pos, "pos", dev, torch.int32)
        B = m.numel() // n_max
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if T is None:
            T = torch.zeros((B, n_max, n_max, n_max, C), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_promote_forward(self.h, _ptr(f), _ptr(f_off), _ptr(m), _ptr(pos), _ptr(_check(T, "T", dev)), _ptr(n),
                                              n_max, C, B, n_max ** 3 * C, self._stream(stream)))
        return T

    def promote_backward(self, gT, f_off, m, pos, gf, n=None, stream=None):
        """gf (flat, same layout as f) += the promoted gradients of gT [B, n_max, n_max, n_max, C]."""
        dev = self.device
        gT, gf = _check(gT, "gT", dev), _check(gf, "gf", dev)
        B, n_max, C = gT.shape[0], gT.shape[1], gT.shape[4]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        self._rc(self.lib.ccn_promote_backward(self.h, _ptr(gT), _ptr(f_off), _ptr(m), _ptr(pos), _ptr(gf), _ptr(n), n_max, C, B,
                                               n_max ** 3 * C, self._stream(stream)))
        return gf

    # ---- chained entry points (one call per stage of a CCN level) ----------------------------------------------------
    def promotion_fuses(self, n_max, C, f):
        """True when ccn_gather_contract18_* read / scatter the promotion inside the fused contraction kernels (no stacked
        T / gT scratch needed): the shapes of the fused path, 16-byte aligned f, default kernel path."""
        return (n_max <= 32 and C in (8, 16, 32, 64, 128) and f.data_ptr() % 16 == 0 and getattr(self, "_kernel_path", 0) == 0)

    def gather_level_forward(self, f, f_off, m, pos, adj, K, bias, n_max, X=None, n=None, adj_mode=ADJ_POSITIVE_PART, alpha=0.01,
                             stream=None):
        """One CCN level from the level l-1 tensors (SMP_beta.h:588-616 for a batch of vertices): promotion + stack +
        contraction -> X [B, n_max^2, 18 C_in] -> Y = X K -> Z = lrelu(Y + bias).  Returns (X, Y, Z)."""
        dev = self.device
        f, adj, K, bias = _check(f, "f", dev), _check(adj, "adj", dev), _check(K, "K", dev), _check(bias, "bias", dev)
        f_off, m, pos = _check(f_off, "f_off", dev, torch.int64), _check(m, "m", dev, torch.int32), _check(pos, "pos", dev, torch.int32)
        B = m.numel() // n_max
        Ci, Co = K.shape[0] // NUM_CONTRACTIONS, K.shape[1]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        T_scratch = None if self.promotion_fuses(n_max, Ci, f) else torch.empty((B, n_max, n_max, n_max, Ci), device=dev)
        if X is None:
            X = torch.zeros((B, n_max * n_max, NUM_CONTRACTIONS * Ci), device=dev, dtype=torch.float32)
        Y = torch.empty((B * n_max * n_max, Co), device=dev, dtype=torch.float32)
        Z = torch.empty_like(Y)
        self._rc(self.lib.ccn_gather_level_forward(self.h, _ptr(f), _ptr(f_off), _ptr(m), _ptr(pos), _ptr(adj), _ptr(K), _ptr(bias),
                                                   _ptr(T_scratch), _ptr(_check(X, "X", dev)), _ptr(Y), _ptr(Z), _ptr(n), n_max, Ci, Co,
                                                   B, n_max * n_max, adj_mode, alpha, self._stream(stream)))
        return X, Y, Z

    def gather_level_backward(self, gZ, X, Y, K, bias, adj, f_off, m, pos, gf, n_max, gK=None, gbias=None, gX=None, n=None,
                              adj_mode=ADJ_POSITIVE_PART, alpha=0.01, stream=None):
        """gZ -> (gf += ..., gK += ..., gbias += ...).  Returns (gf, gK, gbias)."""
        dev = self.device
        gZ, X, Y, K, bias, adj, gf = (_check(t, nm, dev) for t, nm in ((gZ, "gZ"), (X, "X"), (Y, "Y"), (K, "K"), (bias, "bias"),
                                                                        (adj, "adj"), (gf, "gf")))
        B = adj.shape[0]
        Ci, Co = K.shape[0] // NUM_CONTRACTIONS, K.shape[1]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        gK = torch.zeros_like(K) if gK is None else gK
        gbias = torch.zeros_like(bias) if gbias is None else gbias
        gX = torch.empty_like(X) if gX is None else gX
        gT_scratch = None if self.promotion_fuses(n_max, Ci, gf) else torch.empty((B, n_max, n_max, n_max, Ci), device=dev)
        self._rc(self.lib.ccn_gather_level_backward(self.h, _ptr(gZ), _ptr(X), _ptr(Y), _ptr(K), _ptr(bias), _ptr(adj), _ptr(f_off),
                                                    _ptr(m), _ptr(pos), _ptr(gX), _ptr(gT_scratch), _ptr(gf), _ptr(gK), _ptr(gbias),
                                                    _ptr(n), n_max, Ci, Co, B, n_max * n_max, adj_mode, alpha, self._stream(stream)))
        return gf, gK, gbias

    def gather_level_forward_backward_host(self, f, f_group_ptr, inst_group_ptr, f_off, m, pos, adj, K, bias, gZ, Z, gf, gK, gbias,
                                           n, adj_mode=ADJ_POSITIVE_PART, alpha=0.01):
        """Host-array form (see include/ccn_b200.h): all arguments are CPU tensors (pinned for full PCIe speed); Z, gf, gK,
        gbias are overwritten."""
        cpu = torch.device("cpu")
        for t, nm in ((f, "f"), (adj, "adj"), (K, "K"), (bias, "bias"), (gZ, "gZ"), (Z, "Z"), (gf, "gf"), (gK, "gK"), (gbias, "gbias")):
            _check(t, nm, cpu)
        for t, nm in ((f_group_ptr, "f_group_ptr"), (inst_group_ptr, "inst_group_ptr"), (f_off, "f_off")):
            _check(t, nm, cpu, torch.int64)
        _check(m, "m", cpu, torch.int32)
        _check(pos, "pos", cpu, torch.int32)
        Ci, Co = K.shape[0] // NUM_CONTRACTIONS, K.shape[1]
        self._rc(self.lib.ccn_gather_level_forward_backward_host(
            self.h, _ptr(f), _ptr(f_group_ptr), _ptr(inst_group_ptr), f_group_ptr.numel() - 1, _ptr(f_off), _ptr(m), _ptr(pos),
            _ptr(adj), _ptr(K), _ptr(bias), _ptr(gZ), _ptr(Z), _ptr(gf), _ptr(gK), _ptr(gbias), n, Ci, Co, adj_mode, alpha))
        return Z, gf, gK, gbias

    def gather_levels_forward_backward_host(self, f, f_group_ptr, inst_group_ptr, f_off, m, pos, adj, K, bias, gZ, Z, gf, gK, gbias, n,
                                            adj_mode=ADJ_POSITIVE_PART, alpha=0.01):
        """A stack of L levels from host arrays, device resident between the levels (include/ccn_b200.h).  f_off, m, pos, adj, K,
        bias, gK, gbias are LISTS of L CPU tensors (level 1 first); f, gZ, Z, gf as in the single-level call."""
        cpu = torch.device("cpu")
        Lv = len(K)
        for t, nm in ((f, "f"), (gZ, "gZ"), (Z, "Z"), (gf, "gf")):
            _check(t, nm, cpu)
        for t, nm in ((f_group_ptr, "f_group_ptr"), (inst_group_ptr, "inst_group_ptr")):
            _check(t, nm, cpu, torch.int64)
        for l in range(Lv):
            _check(f_off[l], "f_off", cpu, torch.int64), _check(m[l], "m", cpu, torch.int32), _check(pos[l], "pos", cpu, torch.int32)
            for t, nm in ((adj[l], "adj"), (K[l], "K"), (bias[l], "bias"), (gK[l], "gK"), (gbias[l], "gbias")):
                _check(t, nm, cpu)
        arr = lambda ts: (ctypes.c_void_p * Lv)(*[t.data_ptr() for t in ts])  # noqa: E731
        Ci, Co = K[0].shape[0] // NUM_CONTRACTIONS, K[0].shape[1]
        self._rc(self.lib.ccn_gather_levels_forward_backward_host(
            self.h, Lv, _ptr(f), _ptr(f_group_ptr), _ptr(inst_group_ptr), f_group_ptr.numel() - 1, arr(f_off), arr(m), arr(pos), arr(adj),
            arr(K), arr(bias), _ptr(gZ), _ptr(Z), _ptr(gf), arr(gK), arr(gbias), n, Ci, Co, adj_mode, alpha))
        return Z, gf, gK, gbias

    def gather_levels_readout_forward_backward_host(self, f, f_group_ptr, inst_group_ptr, f_off, m, pos, adj, K, bias, W, target,
                                                    predict, loss, gf, gK, gbias, gW, n, adj_mode=ADJ_POSITIVE_PART, alpha=0.01):
        """The level stack + the read-out head + loss from host arrays (include/ccn_b200.h): a training step of the model minus
        level 0.  W [C], target / predict / loss [groups], gW [C]; the rest as gather_levels_forward_backward_host."""
        cpu = torch.device("cpu")
        Lv = len(K)
        for t, nm in ((f, "f"), (gf, "gf"), (W, "W"), (target, "target"), (predict, "predict"), (loss, "loss"), (gW, "gW")):
            _check(t, nm, cpu)
        for t, nm in ((f_group_ptr, "f_group_ptr"), (inst_group_ptr, "inst_group_ptr")):
            _check(t, nm, cpu, torch.int64)
        for l in range(Lv):
            _check(f_off[l], "f_off", cpu, torch.int64), _check(m[l], "m", cpu, torch.int32), _check(pos[l], "pos", cpu, torch.int32)
            for t, nm in ((adj[l], "adj"), (K[l], "K"), (bias[l], "bias"), (gK[l], "gK"), (gbias[l], "gbias")):
                _check(t, nm, cpu)
        arr = lambda ts: (ctypes.c_void_p * Lv)(*[t.data_ptr() for t in ts])  # noqa: E731
        C = K[0].shape[1]
        self._rc(self.lib.ccn_gather_levels_readout_forward_backward_host(
            self.h, Lv, _ptr(f), _ptr(f_group_ptr), _ptr(inst_group_ptr), f_group_ptr.numel() - 1, arr(f_off), arr(m), arr(pos), arr(adj),
            arr(K), arr(bias), _ptr(W), _ptr(target), _ptr(predict), _ptr(loss), _ptr(gf), arr(gK), arr(gbias), _ptr(gW), n, C,
            adj_mode, alpha))
        return predict, loss, gf, gK, gbias, gW

    def host_register(self, t):
        """Page-locks the storage of a CPU tensor / numpy-backed tensor in place (cudaHostRegister)."""
        self._rc(self.lib.ccn_host_register(self.h, ctypes.c_void_p(t.data_ptr()), t.numel() * t.element_size()))

    def host_unregister(self, t):
        self._rc(self.lib.ccn_host_unregister(self.h, ctypes.c_void_p(t.data_ptr())))

    def gather_contract18_forward(self, f, f_off, m, pos, adj, n_max, C, T_scratch=None, out=None, n=None,
                                  adj_mode=ADJ_POSITIVE_PART, stream=None):
        """promotion gather + stack + 18-way contraction: f (flat level l-1 tensors) -> out [B, n_max, n_max, 18 C]."""
        dev = self.device
        f, adj = _check(f, "f", dev), _check(adj, "adj", dev)
        f_off, m, pos = _check(f_off, "f_off", dev, torch.int64), _check(m, "m", dev, torch.int32), _check(pos, "pos", dev, torch.int32)
        B = m.numel() // n_max
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if T_scratch is None and not self.promotion_fuses(n_max, C, f):
            T_scratch = torch.empty((B, n_max, n_max, n_max, C), device=dev, dtype=torch.float32)
        if out is None:
            out = torch.zeros((B, n_max, n_max, NUM_CONTRACTIONS * C), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_gather_contract18_forward(self.h, _ptr(f), _ptr(f_off), _ptr(m), _ptr(pos), _ptr(adj),
                                                        _ptr(T_scratch), _ptr(_check(out, "out", dev)),
                                                        _ptr(n), n_max, C, B, n_max * n_max, n_max * n_max * NUM_CONTRACTIONS * C,
                                                        adj_mode, self._stream(stream)))
        return out

    def gather_contract18_backward(self, gout, adj, f_off, m, pos, gf, gT_scratch=None, n=None, adj_mode=ADJ_POSITIVE_PART,
                                   stream=None):
        """gf (flat, layout of f) += promotion^T(contraction^T(gout)); gout [B, n_max, n_max, 18 C]."""
        dev = self.device
        gout, adj, gf = _check(gout, "gout", dev), _check(adj, "adj", dev), _check(gf, "gf", dev)
        B, n_max, C = gout.shape[0], gout.shape[1], gout.shape[3] // NUM_CONTRACTIONS
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if gT_scratch is None and not self.promotion_fuses(n_max, C, gf):
            gT_scratch = torch.empty((B, n_max, n_max, n_max, C), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_gather_contract18_backward(self.h, _ptr(gout), _ptr(adj), _ptr(f_off), _ptr(m), _ptr(pos),
                                                         _ptr(gT_scratch), _ptr(gf), _ptr(n), n_max, C, B,
                                                         n_max * n_max * NUM_CONTRACTIONS * C, n_max * n_max, adj_mode,
                                                         self._stream(stream)))
        return gf

    def level_forward(self, T, adj, K, bias, X=None, n=None, adj_mode=ADJ_POSITIVE_PART, alpha=0.01, stream=None):
        """contraction + feature mix (+bias, leaky-ReLU): T [B, N, N, N, C_in] -> (X [B, N*N, 18 C_in], Y, Z [B*N*N, C_out])."""
        dev = self.device
        T, adj, K, bias = _check(T, "T", dev), _check(adj, "adj", dev), _check(K, "K", dev), _check(bias, "bias", dev)
        B, N, Ci = T.shape[0], T.shape[1], T.shape[4]
        Co = K.shape[1]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if X is None:
            X = torch.zeros((B, N * N, NUM_CONTRACTIONS * Ci), device=dev, dtype=torch.float32)
        Y = torch.empty((B * N * N, Co), device=dev, dtype=torch.float32)
        Z = torch.empty_like(Y)
        self._rc(self.lib.ccn_level_forward(self.h, _ptr(T), None, _ptr(adj), _ptr(K), _ptr(bias), _ptr(_check(X, "X", dev)), _ptr(Y),
                                            _ptr(Z), _ptr(n), N, Ci, Co, B, N ** 3 * Ci, N * N, adj_mode, alpha, self._stream(stream)))
        return X, Y, Z

    def level_backward(self, gZ, X, Y, K, bias, adj, gT=None, gK=None, gbias=None, n=None, adj_mode=ADJ_POSITIVE_PART, alpha=0.01,
                       beta=0.0, stream=None):
        """Returns (gT [B, N, N, N, C_in], gK, gbias); gK / gbias accumulate into the tensors passed in."""
        dev = self.device
        gZ, X, Y, K, bias, adj = (_check(t, nm, dev) for t, nm in ((gZ, "gZ"), (X, "X"), (Y, "Y"), (K, "K"), (bias, "bias"), (adj, "adj")))
        B, N = adj.shape[0], adj.shape[1]
        Ci, Co = X.shape[2] // NUM_CONTRACTIONS, K.shape[1]
        if n is not None:
            n = _check(n, "n", dev, torch.int32)
        if gT is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an existing gT")
            gT = torch.empty((B, N, N, N, Ci), device=dev, dtype=torch.float32)
        gK = torch.zeros_like(K) if gK is None else gK
        gbias = torch.zeros_like(bias) if gbias is None else gbias
        gX = torch.empty_like(X)
        self._rc(self.lib.ccn_level_backward(self.h, _ptr(gZ), _ptr(X), _ptr(Y), _ptr(K), _ptr(bias), _ptr(adj), _ptr(gX),
                                             _ptr(_check(gT, "gT", dev)), None, _ptr(gK), _ptr(gbias), _ptr(n), N, Ci, Co, B, N * N,
                                             N ** 3 * Ci, adj_mode, alpha, beta, self._stream(stream)))
        return gT, gK, gbias

    # ---- TensorMul / CustomMatMulTensor ----------------------------------------------------------------------------
    def tensor_mul_forward(self, A, B, stream=None):
        """A [Bt, R, K, D], B [Bt, K, Cc, D] -> [Bt, R, Cc, D] (TensorMul, per channel)."""
        dev = self.device
        A, B = _check(A, "A", dev), _check(B, "B", dev)
        Bt, R, K, D = A.shape
        Cc = B.shape[2]
        out = torch.empty((Bt, R, Cc, D), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_tensor_mul_forward(self.h, _ptr(A), _ptr(B), _ptr(out), R, K, Cc, D, Bt, self._stream(stream)))
        return out

    def tensor_mul_backward(self, A, B, gout, gA=None, gB=None, beta=0.0, stream=None):
        dev = self.device
        A, B, gout = _check(A, "A", dev), _check(B, "B", dev), _check(gout, "gout", dev)
        Bt, R, K, D = A.shape
        Cc = B.shape[2]
        if gA is None:
            gA = torch.zeros_like(A)
        if gB is None:
            gB = torch.zeros_like(B)
        self._rc(self.lib.ccn_tensor_mul_backward(self.h, _ptr(A), _ptr(B), _ptr(gout), _ptr(gA), _ptr(gB), R, K, Cc, D, Bt, beta,
                                                  self._stream(stream)))
        return gA, gB

    def custom_matmul_tensor_forward(self, Kt, X, stream=None):
        """Kt [P, V], X [..., V] -> Y [..., P] (CustomMatMulTensor)."""
        dev = self.device
        Kt, X = _check(Kt, "Kt", dev), _check(X, "X", dev)
        P, V = Kt.shape
        M = X.numel() // V
        Y = torch.empty(tuple(X.shape[:-1]) + (P,), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_custom_matmul_tensor_forward(self.h, _ptr(Kt), _ptr(X), _ptr(Y), M, V, P, self._stream(stream)))
        return Y

    def custom_matmul_tensor_backward(self, Kt, X, gY, gKt=None, gX=None, beta_x=0.0, stream=None):
        dev = self.device
        Kt, X, gY = _check(Kt, "Kt", dev), _check(X, "X", dev), _check(gY, "gY", dev)
        P, V = Kt.shape
        M = X.numel() // V
        if gKt is None:
            gKt = torch.zeros_like(Kt)
        if gX is None:
            gX = torch.zeros_like(X)
        self._rc(self.lib.ccn_custom_matmul_tensor_backward(self.h, _ptr(Kt), _ptr(X), _ptr(gY), _ptr(gKt), _ptr(gX), M, V, P, beta_x,
                                                            self._stream(stream)))
        return gKt, gX

    # ---- feature mix ------------------------------------------------------------------------------------------------
    def mix_forward(self, X, W, bias=None, want_Y=True, alpha=0.01, stream=None):
        """X: [M, K], W: [K, P] -> (Y [M, P] or None, Z = lrelu(Y + bias) or None)."""
        dev = self.device
        X, W = _check(X, "X", dev), _check(W, "W", dev)
        M, K = X.shape
        P = W.shape[1]
        if W.shape[0] != K:
            raise ValueError("X is [M, %d] but W is %s" % (K, tuple(W.shape)))  # MatMul.h:31 assert
        Y = torch.empty((M, P), device=dev, dtype=torch.float32) if (want_Y or bias is None) else None
        Z = torch.empty((M, P), device=dev, dtype=torch.float32) if bias is not None else None
        if bias is not None:
            _check(bias, "bias", dev)
        self._rc(self.lib.ccn_mix_forward(self.h, _ptr(X), _ptr(W), _ptr(bias), _ptr(Y), _ptr(Z), M, K, P, alpha,
                                          self._stream(stream)))
        return Y, Z

    def mix_backward(self, X, W, gZ, bias=None, Y=None, gX=None, gW=None, gbias=None, need_gX=True, alpha=0.01,
                     beta_x=0.0, stream=None):
        """Returns (gX, gW, gbias).  gW / gbias accumulate into the tensors passed in (fresh zeros otherwise)."""
        dev = self.device
        X, W, gZ = _check(X, "X", dev), _check(W, "W", dev), _check(gZ, "gZ", dev)
        M, K = X.shape
        P = W.shape[1]
        if need_gX and gX is None:
            if beta_x != 0.0:
                raise ValueError("beta_x != 0 needs an existing gX")
            gX = torch.empty((M, K), device=dev, dtype=torch.float32)
        if gW is None:
            gW = torch.zeros((K, P), device=dev, dtype=torch.float32)
        if bias is not None and gbias is None:
            gbias = torch.zeros((P,), device=dev, dtype=torch.float32)
        self._rc(self.lib.ccn_mix_backward(self.h, _ptr(X), _ptr(W), _ptr(bias), _ptr(Y), _ptr(gZ), _ptr(gX), _ptr(gW),
                                           _ptr(gbias), M, K, P, alpha, beta_x, self._stream(stream)))
        return gX, gW, gbias
