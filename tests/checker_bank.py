"""TEST DOUBLE (CPU): the interface of cmdiad_b200.Bank that cmdiad_b200.dropin uses, answered by the oracle.

The build container has no GPU and the GPU box has no reference tree, so the patched reference can never meet the real
library in one process.  tests/test_integration_patch.py therefore checks the two halves separately:
  * here (CPU): the PATCHED reference, driven through its real public API, with this checker standing in for the device
    bank, must reproduce the UNPATCHED reference bit for bit -- that pins the wiring of the seams;
  * on the GPU: tests/test_gpu_dropin.py drives the same seam objects with the real Bank.
Never imported by the product."""
import types

import numpy as np
import torch

from oracle import restate as O


class CheckerBank:
    def __init__(self, dim, capacity_rows, device=0, row_offset=0):
        self.dim, self.capacity = int(dim), int(capacity_rows)
        self._rows = []
        self._data = None
        self.calls = []

    def _cat(self):
        if self._data is None:
            self._data = torch.cat(self._rows, 0)
        return self._data

    @property
    def rows(self):
        return self._cat().shape[0] if (self._rows or self._data is not None) else 0

    def append(self, rows):
        rows = torch.as_tensor(rows, dtype=torch.float32)
        assert rows.shape[1] == self.dim and self.rows + rows.shape[0] <= self.capacity
        if self._data is not None:
            self._rows, self._data = [self._data], None
        self._rows.append(rows.clone())
        self.calls.append("append")

    def stats(self):
        t = self._cat()
        m, s = torch.mean(t), torch.std(t)   # the reference's own float32 reductions
        self.calls.append("stats")
        return float(m), float(s), 0.0, 0.0

    def normalize(self, mean, std):
        self._data = (self._cat() - torch.tensor(np.float32(mean))) / torch.tensor(np.float32(std))
        self._rows = []
        self.calls.append("normalize")

    def gather(self, idx):
        self._data = self._cat()[torch.as_tensor(np.asarray(idx, dtype=np.int64))]
        self._rows = []
        self.calls.append("gather")

    def read(self, row0=0, n_rows=None):
        t = self._cat()
        return t[row0:] if n_rows is None else t[row0:row0 + n_rows]

    def finalize(self):
        self.calls.append("finalize")

    def build_knn(self):
        self.calls.append("build_knn")

    def coreset_select(self, n_select, csr, dtype_mode=0):
        lib = self._cat().numpy()
        if csr is None:
            z = lib.astype(np.float64)
        else:
            indptr, indices, data, d_proj = csr
            z = O.project_restated(lib, np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32),
                                   np.ascontiguousarray(data, np.float64), int(d_proj))
        self.calls.append("coreset_select")
        return O.coreset_restated(z, int(n_select), "FP16" if dtype_mode == 0 else "TF32")

    def score(self, patch, feature_map_dims, out_hw=224, full=False):
        ref = O.score_restated(torch.as_tensor(patch), self._cat(), tuple(feature_map_dims), out_hw)
        self.calls.append("score")
        return types.SimpleNamespace(s=np.array([ref["s"]], np.float32), s_map=ref["s_map"], min_val=ref["min_val"],
                                     min_idx=ref["min_idx"])

    def close(self):
        pass
