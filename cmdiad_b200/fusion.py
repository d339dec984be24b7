"""Late-fusion head on the device and the device-side result store (8f-3) -- thin marshalling only.

`LateFusion` drives cmdb_score_fused_batch*: B images are scored against 1..3 banks (one per modality), the lambda
scaling and the two linear One-Class-SVM heads of compute_s_s_map (multiple_features.py:986-994) run in one kernel on the
per-modality maps while they are still in HBM, and per image one float64 map + one float64 score come back (or stay on
the device for cmdb_eval_*).  torch is used for buffers and streams only."""
import ctypes

import numpy as np
import torch

from . import _lib as L
from .bank import _as_f32, _ptr


class FusedResult:
    """arrays of one fused batch: s [B] float64, s_map [B,hw,hw] float64 or None, s_modal [B,M] float32 (lambda * s per
    modality), min_val / min_idx: lists per modality ([B,P_m]) or None"""
    __slots__ = ("s", "s_map", "s_modal", "min_val", "min_idx")

    def __len__(self):
        return self.s.shape[0]


class LateFusion:
    def __init__(self, banks, s_lambda, smap_lambda, detect_coef, detect_offset, seg_coef, seg_offset):
        self._lib = L.load()
        self.banks = list(banks)
        M = len(self.banks)
        assert 1 <= M <= 3
        h = L.FusionHead()
        h.n_modal = M
        dc, sc = np.asarray(detect_coef, np.float64).reshape(-1), np.asarray(seg_coef, np.float64).reshape(-1)
        assert dc.shape[0] == M and sc.shape[0] == M, "one coefficient per modality"
        for m in range(M):
            h.s_lambda[m] = float(s_lambda[m])
            h.smap_lambda[m] = float(smap_lambda[m])
            h.detect_coef[m] = float(dc[m])
            h.seg_coef[m] = float(sc[m])
        h.detect_offset = float(np.asarray(detect_offset).reshape(-1)[0])
        h.seg_offset = float(np.asarray(seg_offset).reshape(-1)[0])
        self.head = h
        self._handles = (ctypes.c_void_p * M)(*[b._h for b in self.banks])

    @classmethod
    def from_sklearn(cls, banks, s_lambda, smap_lambda, detect_fuser, seg_fuser):
        """heads = fitted sklearn SGDOneClassSVM objects (features.py:114-115, 352-358): purely linear at predict time"""
        return cls(banks, s_lambda, smap_lambda, detect_fuser.coef_, detect_fuser.offset_, seg_fuser.coef_, seg_fuser.offset_)

    def max_batch(self):
        return min(32, *[b.max_shard_batch() for b in self.banks])

    # ---- result store on banks[0]'s GPU ---------------------------------------------------------------------------
    def eval_reserve(self, n_images, out_hw=224):
        L.check(self._lib.cmdb_eval_reserve(self.banks[0]._h, int(n_images), int(out_hw)))
        self._eval_hw = int(out_hw)

    def eval_reset(self):
        L.check(self._lib.cmdb_eval_reset(self.banks[0]._h))

    def eval_count(self):
        n = ctypes.c_int64()
        L.check(self._lib.cmdb_eval_count(self.banks[0]._h, ctypes.byref(n)))
        return n.value

    def eval_read(self, first=0, n=None, maps=True):
        n = self.eval_count() - first if n is None else n
        hw = self._eval_hw
        m = np.empty((n, hw, hw), np.float64) if maps else None
        s = np.empty(n, np.float64)
        L.check(self._lib.cmdb_eval_read(self.banks[0]._h, int(first), int(n), _ptr(m), _ptr(s)))
        return m, s

    # ---- scoring ----------------------------------------------------------------------------------------------------
    def _args(self, patches, dims):
        M = len(self.banks)
        assert len(patches) == M and len(dims) == M
        patches = [_as_f32(p) for p in patches]
        B = patches[0].shape[0]
        is_cuda = patches[0].is_cuda
        for m, (p, b) in enumerate(zip(patches, self.banks)):
            assert p.dim() == 3 and p.shape[0] == B and p.shape[2] == b.dim, f"modality {m}: expected [B,P,{b.dim}], got {tuple(p.shape)}"
            assert p.is_cuda == is_cuda, "all modalities on the host or all on the device"
            b._order_after_producer(p)
        ptrs = (ctypes.c_void_p * M)(*[p.data_ptr() for p in patches])
        P = (ctypes.c_int * M)(*[int(p.shape[1]) for p in patches])
        fh = (ctypes.c_int * M)(*[int(d[0]) for d in dims])
        fw = (ctypes.c_int * M)(*[int(d[1]) for d in dims])
        return patches, ptrs, P, fh, fw, B, int(is_cuda)

    @staticmethod
    def _alloc(B, M, Ps, out_hw, host_maps, want_patch):
        r = FusedResult()
        r.s = np.empty(B, np.float64)
        r.s_map = np.empty((B, out_hw, out_hw), np.float64) if host_maps else None
        r.s_modal = np.empty((B, M), np.float32)
        r.min_val = [np.empty((B, p), np.float32) for p in Ps] if want_patch else None
        r.min_idx = [np.empty((B, p), np.int64) for p in Ps] if want_patch else None
        outs = (L.FusedOut * B)()
        for i in range(B):
            o = outs[i]
            o.s = ctypes.cast(r.s.ctypes.data + 8 * i, L.c_f64_p)
            if host_maps:
                o.s_map = ctypes.cast(r.s_map.ctypes.data + r.s_map.strides[0] * i, L.c_f64_p)
            o.s_modal = ctypes.cast(r.s_modal.ctypes.data + r.s_modal.strides[0] * i, L.c_f32_p)
            if want_patch:
                for m in range(M):
                    o.min_val[m] = ctypes.cast(r.min_val[m].ctypes.data + r.min_val[m].strides[0] * i, L.c_f32_p)
                    o.min_idx[m] = ctypes.cast(r.min_idx[m].ctypes.data + r.min_idx[m].strides[0] * i, L.c_i64_p)
        return r, outs

    def score_batch(self, patches, dims, out_hw=224, keep_on_device=False, host_maps=True, want_patch=False):
        """patches: per modality a float32 [B,P_m,D_m] tensor (host or device; RAW when the banks have a query norm set);
        dims: per modality (fh, fw).  Any B (internal sub-batches, pipelined)."""
        patches, ptrs, P, fh, fw, B, is_cuda = self._args(patches, dims)
        flags = (L.FUSED_KEEP_ON_DEVICE if keep_on_device else 0) | (0 if host_maps else L.FUSED_NO_HOST_MAPS)
        r, outs = self._alloc(B, len(self.banks), [int(x) for x in P], out_hw, host_maps, want_patch)
        L.check(self._lib.cmdb_score_fused_batch(self._handles, ptrs, P, fh, fw, B, int(out_hw), is_cuda,
                                                 ctypes.byref(self.head), flags, outs))
        return r

    def score_batch_async(self, patches, dims, out_hw=224, keep_on_device=False, host_maps=True, want_patch=False):
        """submit form (B <= max_batch()); returns a ticket, .wait() -> FusedResult.  Two tickets may be outstanding."""
        patches, ptrs, P, fh, fw, B, is_cuda = self._args(patches, dims)
        flags = (L.FUSED_KEEP_ON_DEVICE if keep_on_device else 0) | (0 if host_maps else L.FUSED_NO_HOST_MAPS)
        t = ctypes.c_int64()
        L.check(self._lib.cmdb_score_fused_batch_submit(self._handles, ptrs, P, fh, fw, B, int(out_hw), is_cuda,
                                                        ctypes.byref(self.head), flags, ctypes.byref(t)))
        return _FusedTicket(self, t.value, patches, (B, len(self.banks), [int(x) for x in P], out_hw, host_maps, want_patch))


class _FusedTicket:
    def __init__(self, fusion, ticket, patches, shape):
        self._f, self._ticket, self._patches, self._shape = fusion, ticket, patches, shape
        self._result = None

    def wait(self):
        if self._result is None:
            r, outs = LateFusion._alloc(*self._shape)
            L.check(self._f._lib.cmdb_score_fused_batch_wait(self._f.banks[0]._h, self._ticket, outs))
            self._result, self._patches = r, None
        return self._result
