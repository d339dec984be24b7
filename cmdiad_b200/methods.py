"""Host-side mirror of the reference's memory-bank method classes, driving the B200 library.

Mirrors `Features` (feature_extractors/features.py:21-425) and its six subclasses
(feature_extractors/multiple_features.py:28-1015) for the hot path only: same method names, argument meaning, attribute
names and error behaviour for add_sample_to_mem_bank / run_coreset / add_sample_to_late_fusion_mem_bank /
run_late_fusion / predict / compute_s_s_map / get_coreset_idx_randomp / calculate_dist / compute_single_s_s_map.
What differs is where the work happens:
  * the banks live pre-allocated in HBM (cmdiad_b200.Bank) instead of Python lists + torch.cat on the host;
  * calculate_dist returns a lazy handle and compute_single_s_s_map runs the fused CUDA path, so the [P,R] distance
    matrix of features.py:190 never exists;
  * get_coreset_idx_randomp runs the projection and the whole greedy loop in one persistent kernel.
Backbones (DINO / Point-MAE, features.py:123-184) and the hallucination networks are out of scope: a sample is a dict
of already extracted float32 patches {"rgb": [P,D], "xyz": [P,D], "fusion": [P,D]} or whatever `feature_fn(sample)`
returns in that form.  The late-fusion One-Class-SVM head and the AUROC metrics stay sklearn calls exactly as in the
reference (features.py:114-115, 321-322, 352-358).
"""
import math
from argparse import Namespace

import numpy as np
import torch
from sklearn import linear_model, random_projection
from sklearn.metrics import roc_auc_score

from . import _lib as L
from .bank import Bank

_MODALS = ("xyz", "rgb", "fusion")


def default_args(**over):
    """The hot-path knobs of main.py:85-189 with their defaults."""
    a = dict(dist_method_s="l2", dist_method_coreset="l2", coreset_dtype="FP16", f_coreset=0.1, coreset_eps=0.9,
             random_state=None, rgb_s_lambda=0.1, rgb_smap_lambda=0.1, xyz_s_lambda=1.0, xyz_smap_lambda=1.0,
             fusion_s_lambda=1.0, fusion_smap_lambda=1.0, main_modality="rgb", gt_size=224, max_sample=500,
             ocsvm_nu=0.5, ocsvm_maxiter=1000, save_seg_results=False, save_raw_results=False)
    a.update(over)
    return Namespace(**a)


class FusedDist:
    """What calculate_dist returns instead of the [P,R] matrix: the operands of the fused kernel."""
    __slots__ = ("patch", "lib")

    def __init__(self, patch, lib):
        self.patch, self.lib = patch, lib


class _PendingLib(list):
    """`self.patch_*_lib` before run_coreset: list semantics of the reference, rows go straight to the device bank."""

    def __init__(self, owner, modal):
        super().__init__()
        self._owner, self._modal = owner, modal

    def append(self, patch):
        self._owner._append(self._modal, patch)
        super().append(tuple(patch.shape))  # shapes only; the data is in HBM


class DeviceLib:
    """`self.patch_*_lib` after run_coreset: tensor-like view of the device bank (shape, indexing read rows back)."""

    def __init__(self, bank):
        self.bank = bank
        self._host = None

    @property
    def shape(self):
        return torch.Size((self.bank.rows, self.bank.dim))

    def cpu(self):
        if self._host is None:
            self._host = self.bank.read()
        return self._host

    def __getitem__(self, idx):
        return self.cpu()[idx]

    def __len__(self):
        return self.bank.rows


def coreset_idx_randomp(bank, n, eps=0.90, coreset_dtype="FP16", random_state=None, dist_method="l2", verbose=False):
    """Features.get_coreset_idx_randomp (features.py:360-425) on a device bank: sklearn draws the sparse projection
    matrix on the host exactly as the reference does, the projection and the whole greedy loop run on the GPU.
    Returns a CPU LongTensor [n] of LOCAL rows with idx[0] == 0."""
    if dist_method != "l2":
        raise NotImplementedError  # l1 / dot / cos_dist branches of features.py:379-384 are out of scope
    if coreset_dtype == "FP16":
        mode = L.CORESET_FP16
    elif coreset_dtype == "TF32":
        mode = L.CORESET_FP64  # the reference only flips a matmul flag; its data stays float64
    else:
        raise NotImplementedError  # features.py:394-395
    if verbose:
        print(f"   Fitting random projections. Start dim = {(bank.rows, bank.dim)}.")
    csr = None
    try:
        transformer = random_projection.SparseRandomProjection(eps=eps, random_state=random_state)
        # fit() reads only X.shape / X.dtype; the reference's torch float32 input is validated to float64, so the
        # components stay float64 with unsorted indices.  A zero-stride float64 dummy reproduces that (and consumes
        # numpy's global RNG identically when random_state is None) without touching the bank.
        transformer.fit(np.broadcast_to(np.zeros((1, 1)), (bank.rows, bank.dim)))
        c = transformer.components_
        csr = (c.indptr, c.indices, c.data, c.shape[0])
        if verbose:
            print(f"   DONE.                 Transformed dim = ({bank.rows}, {c.shape[0]}).")
    except ValueError:
        print("   Error: could not project vectors. Please increase `eps`.")
    # n = int(f_coreset * rows) may be 0 for tiny banks: the reference's `range(n - 1)` loop is then empty and it returns
    # [0] (features.py:372, 401, 425)
    idx = bank.coreset_select(max(1, int(n)), csr, mode)
    return torch.from_numpy(idx)


class Features(torch.nn.Module):
    """Base class (features.py:21-121, hot-path attributes only)."""
    # per-class wiring, see the subclasses
    bank_modals = ()
    mean_from = {}
    std_from = {}

    def __init__(self, args=None, device=0, feature_fn=None, parity_stats=False, bank_capacity_rows=None, verbose=False,
                 device_head=True):
        super().__init__()
        self.device_head = device_head
        self.args = args if args is not None else default_args()
        self.cuda_device = int(device)
        self.feature_fn = feature_fn
        # False: statistics on the device (float64 accumulation).  True: keep host copies and call torch.mean / torch.std
        # exactly like the reference.  dict {modal: (mean|None, std|None)}: use these scalars (bit-parity experiments)
        self.parity_stats = parity_stats
        self.bank_capacity_rows = bank_capacity_rows
        self.verbose = verbose
        self.class_name = None
        self.gt_size = self.args.gt_size
        self.f_coreset = self.args.f_coreset
        self.coreset_eps = self.args.coreset_eps
        self.coreset_dtype = self.args.coreset_dtype
        self.random_state = self.args.random_state
        self.n_reweight = 3
        np.random.seed(0)  # set_seeds(0) in the reference ctor (features.py:48): the projection draws from this RNG
        self._banks = {}
        self._host_copies = {m: [] for m in _MODALS}
        self.patch_xyz_lib = _PendingLib(self, "xyz")
        self.patch_rgb_lib = _PendingLib(self, "rgb")
        self.patch_fusion_lib = _PendingLib(self, "fusion")
        for m in _MODALS:
            setattr(self, f"{m}_mean", 0)
            setattr(self, f"{m}_std", 0)
        self.image_preds, self.image_labels, self.pixel_preds, self.pixel_labels = [], [], [], []
        self.gts, self.predictions, self.img_name = [], [], []
        self.image_rocauc = self.pixel_rocauc = self.au_pro = self.au_pro_001 = 0
        self.detect_fuser = linear_model.SGDOneClassSVM(random_state=42, nu=self.args.ocsvm_nu,
                                                        max_iter=self.args.ocsvm_maxiter)
        self.seg_fuser = linear_model.SGDOneClassSVM(random_state=42, nu=self.args.ocsvm_nu,
                                                     max_iter=self.args.ocsvm_maxiter)
        self.s_lib, self.s_map_lib = [], []
        self.coreset_idx = None
        self._fusion = None
        self._pinned = {}

    # ---- storage ---------------------------------------------------------------------------------------------
    def _patches(self, sample):
        out = self.feature_fn(sample) if self.feature_fn is not None else sample
        return {k: torch.as_tensor(v, dtype=torch.float32) for k, v in out.items() if v is not None}

    def _append(self, modal, patch):
        patch = torch.as_tensor(patch, dtype=torch.float32)
        if modal not in self._banks:
            # the reference's fit loop admits max_sample + 1 samples (cmdiad_runner.py:46-52: `flag += 1; if flag >
            # self.count: break` after the append)
            cap = self.bank_capacity_rows or (int(self.args.max_sample) + 1) * patch.shape[0]
            self._banks[modal] = Bank(patch.shape[1], cap, device=self.cuda_device)
        self._banks[modal].append(patch)
        if self.parity_stats is True:
            self._host_copies[modal].append(patch.cpu())

    def _lib(self, modal):
        return getattr(self, f"patch_{modal}_lib")

    # ---- features.py:186-205 -----------------------------------------------------------------------------------
    def calculate_dist(self, single_patch, patch_lib):
        assert len(single_patch.shape) == 2
        assert len(patch_lib.shape) == 2
        if self.args.dist_method_s != "l2":
            raise NotImplementedError  # l1 / cos_dist go through cupy in the reference and are out of scope
        return FusedDist(single_patch, patch_lib)

    # ---- features.py:225-297 -----------------------------------------------------------------------------------
    def compute_single_s_s_map(self, patch, dist, feature_map_dims, modal="xyz"):
        lib = self._lib(modal)
        assert isinstance(dist, FusedDist) and dist.lib is lib, "dist must come from calculate_dist on this modal's bank"
        r = lib.bank.score(dist.patch, feature_map_dims, out_hw=self.gt_size)
        self.last_score = r
        s = torch.tensor(r.s[0])
        s_map = torch.from_numpy(r.s_map).view(1, self.gt_size, self.gt_size)
        return s, s_map

    # ---- features.py:352-358 -----------------------------------------------------------------------------------
    def run_late_fusion(self):
        self.s_lib = torch.cat(self.s_lib, 0)
        self.s_map_lib = torch.cat(self.s_map_lib, 0)
        self.detect_fuser.fit(self.s_lib)
        self.seg_fuser.fit(self.s_map_lib)
        self._fusion = None  # the device-side head is rebuilt from the new coef_ / offset_ on first use

    # ---- features.py:360-425 -----------------------------------------------------------------------------------
    def get_coreset_idx_randomp(self, z_lib, n=1000, eps=0.90, coreset_dtype="FP16", force_cpu=False, lib=""):
        """z_lib: DeviceLib (the normalised bank in HBM).  Returns a CPU LongTensor [n] with idx[0] == 0."""
        return coreset_idx_randomp(z_lib.bank, n, eps, coreset_dtype, self.random_state, self.args.dist_method_coreset,
                                   self.verbose)

    # ---- run_coreset: one implementation for the six variants ----------------------------------------------------
    def _normalised_modals(self):
        return self.bank_modals

    def _coreset_modals(self):
        return self.bank_modals

    def run_coreset(self):
        stats = {}
        for m in set(self.mean_from.values()) | set(self.std_from.values()):
            if self.parity_stats is True:
                cat = torch.cat(self._host_copies[m], 0)
                stats[m] = (torch.mean(cat), torch.std(cat))
            else:
                mean, std, _, _ = self._banks[m].stats()
                stats[m] = [torch.tensor(mean, dtype=torch.float32), torch.tensor(std, dtype=torch.float32)]
                if isinstance(self.parity_stats, dict) and m in self.parity_stats:
                    for i, v in enumerate(self.parity_stats[m]):
                        if v is not None:
                            stats[m][i] = torch.tensor(np.float32(v))
        for m in self.bank_modals:
            setattr(self, f"{m}_mean", stats[self.mean_from[m]][0])
            setattr(self, f"{m}_std", stats[self.std_from[m]][1])
            setattr(self, f"patch_{m}_lib", DeviceLib(self._banks[m]))
        for m in self._normalised_modals():
            self._banks[m].normalize(float(getattr(self, f"{m}_mean")), float(getattr(self, f"{m}_std")))
        if self.f_coreset < 1:
            for m in self._coreset_modals():
                lib = self._lib(m)
                self.coreset_idx = self.get_coreset_idx_randomp(lib, n=int(self.f_coreset * lib.shape[0]),
                                                                eps=self.coreset_eps, lib=f"patch_{m}_lib",
                                                                coreset_dtype=self.coreset_dtype)
                self._banks[m].gather(self.coreset_idx.numpy())
                setattr(self, f"patch_{m}_lib", DeviceLib(self._banks[m]))
        for m in self._score_modals():
            self._banks[m].finalize()
            # the w_dist top-3 of features.py:239-254 only depends on the bank -> precomputed per bank row
            self._banks[m].build_knn()
        self._host_copies = {m: [] for m in _MODALS}
        self._fusion = None

    # ---- scoring of one sample: shared by late fusion and predict ------------------------------------------------
    score_modals = ()  # order of the columns of s / s_map

    def _score_modals(self):
        return self.score_modals

    def _score_sample(self, patches):
        s_cols, map_cols = [], []
        for m in self._score_modals():
            patch = (patches[m] - getattr(self, f"{m}_mean")) / getattr(self, f"{m}_std")
            dist = self.calculate_dist(patch, self._lib(m))
            side = int(math.sqrt(patch.shape[0]))
            s_m, s_map_m = self.compute_single_s_s_map(patch, dist, (side, side), modal=m)
            s_cols.append(getattr(self.args, f"{m}_s_lambda") * s_m)
            map_cols.append(getattr(self.args, f"{m}_smap_lambda") * s_map_m)
        s = torch.tensor([s_cols])
        s_map = torch.cat(map_cols, dim=0).squeeze().reshape(len(map_cols), -1).permute(1, 0)
        return s, s_map

    def add_sample_to_mem_bank(self, sample, class_name=None):
        self.class_name = class_name
        patches = self._patches(sample)
        for m in self.bank_modals:
            self._lib(m).append(patches[m])

    def add_sample_to_late_fusion_mem_bank(self, sample):
        s, s_map = self._score_sample(self._patches(sample))
        self.s_lib.append(s)
        self.s_map_lib.append(s_map)

    def predict(self, sample, mask, label, rgb_path):
        """one test image (cmdiad_runner.py:84).  device_head (default): normalisation, scoring, lambda scaling and the two
        linear heads all run behind the ABI (predict_batch with one image); otherwise the mirror of the reference's
        compute_s_s_map with sklearn's score_samples on the host."""
        if self.device_head:
            self.predict_batch([sample], [mask], [label], [rgb_path])
        else:
            self.compute_s_s_map(self._patches(sample), mask, label, rgb_path)

    # ---- batch forms: same per-image results, one device call per batch ---------------------------------
    def _stack(self, patch_dicts, m, slot=0):
        """[B,P,D] batch of RAW patches of modality m: device samples are stacked on the device, host samples are gathered
        into a reused pinned block (two alternating blocks, so that a pipelined caller may fill one while the other is
        still being copied)"""
        first = patch_dicts[0][m]
        if first.is_cuda:
            return torch.stack([pd[m] for pd in patch_dicts])
        key = (m, tuple(first.shape), slot)
        buf = self._pinned.get(key)
        if buf is None or buf.shape[0] < len(patch_dicts):
            # pinned allocations cost tens of milliseconds: one block per (modality, slot), sized for the largest batch
            cap = max(len(patch_dicts), 32)
            buf = self._pinned[key] = torch.empty((cap,) + tuple(first.shape), dtype=torch.float32).pin_memory()
        for i, pd in enumerate(patch_dicts):
            buf[i].copy_(pd[m])
        return buf[:len(patch_dicts)]

    def _set_query_norm(self, enabled):
        for m in self._score_modals():
            self._banks[m].set_query_norm(float(getattr(self, f"{m}_mean")), float(getattr(self, f"{m}_std")), enabled)

    def _score_samples(self, patch_dicts):
        """[(s [1,m], s_map [gt*gt,m])] for a list of samples: the reference's per-image loop (cmdiad_runner.py:58-65,
        80-85) collapsed into one batched device call per modality; (patch - mean) / std runs on the device"""
        cols = []
        self._set_query_norm(True)
        try:
            for m in self._score_modals():
                batch = self._stack(patch_dicts, m)
                side = int(math.sqrt(batch.shape[1]))
                res = self._lib(m).bank.score_batch(batch, (side, side), out_hw=self.gt_size)
                cols.append((getattr(self.args, f"{m}_s_lambda"), getattr(self.args, f"{m}_smap_lambda"), res))
        finally:
            self._set_query_norm(False)
        out = []
        for i in range(len(patch_dicts)):
            s = torch.tensor([[lam_s * torch.tensor(res[i].s[0]) for lam_s, _, res in cols]])
            maps = [lam_m * torch.from_numpy(res[i].s_map).view(1, self.gt_size, self.gt_size) for _, lam_m, res in cols]
            out.append((s, torch.cat(maps, dim=0).squeeze().reshape(len(maps), -1).permute(1, 0)))
        return out

    def add_samples_to_late_fusion_mem_bank(self, samples):
        for s, s_map in self._score_samples([self._patches(x) for x in samples]):
            self.s_lib.append(s)
            self.s_map_lib.append(s_map)

    def fusion(self):
        """the fitted late-fusion head on the device: built once after run_late_fusion"""
        if getattr(self, "_fusion", None) is None:
            from .fusion import LateFusion
            mods = self._score_modals()
            self._fusion = LateFusion.from_sklearn([self._banks[m] for m in mods],
                                                   [getattr(self.args, f"{m}_s_lambda") for m in mods],
                                                   [getattr(self.args, f"{m}_smap_lambda") for m in mods],
                                                   self.detect_fuser, self.seg_fuser)
        return self._fusion

    def predict_batch(self, samples, masks, labels, rgb_paths, keep_on_device=False):
        """predict() for a list of samples with everything after the backbones on the device: normalisation, distance
        GEMM, re-weighting, maps, lambda scaling and the two linear One-Class-SVM heads (multiple_features.py:976-994).
        Batches larger than the per-call limit are pipelined (batch k + 1 is staged and submitted before the host waits
        for batch k).  keep_on_device: the fused maps are also appended to the device-side result store (reserve it with
        self.fusion().eval_reserve(n_images) first) for calculate_metrics_device."""
        fus = self.fusion()
        mods = self._score_modals()
        step = fus.max_batch()
        pds = [self._patches(x) for x in samples]
        self._set_query_norm(True)
        try:
            pending = None
            for k, b0 in enumerate(range(0, len(pds), step)):
                chunk = pds[b0:b0 + step]
                batches = [self._stack(chunk, m, slot=k & 1) for m in mods]
                dims = [(int(math.sqrt(b.shape[1])),) * 2 for b in batches]
                t = fus.score_batch_async(batches, dims, out_hw=self.gt_size, keep_on_device=keep_on_device)
                if pending is not None:
                    self._record_fused(pending[0].wait(), *pending[1])
                pending = (t, (masks[b0:b0 + step], labels[b0:b0 + step], rgb_paths[b0:b0 + step]))
            if pending is not None:
                self._record_fused(pending[0].wait(), *pending[1])
        finally:
            self._set_query_norm(False)

    def _record_fused(self, res, masks, labels, rgb_paths):
        """result bookkeeping of multiple_features.py:996-1003 for a fused batch (per-image arrays, not scalar lists)"""
        self.last_fused = res
        for i in range(len(res)):
            mask = torch.as_tensor(masks[i])
            self.image_preds.append(res.s[i:i + 1].copy())
            self.image_labels.append(labels[i])
            self.pixel_preds.append(res.s_map[i].reshape(-1))
            self.pixel_labels.append(mask.flatten().numpy())
            self.predictions.append(res.s_map[i])
            self.gts.append(mask.detach().cpu().squeeze().numpy())
            self.img_name.append(rgb_paths[i])

    def compute_s_s_map(self, patches, mask, label, rgb_path=None):
        s, s_map = self._score_sample(patches)
        self._record(s, s_map, mask, label, rgb_path)

    def _record(self, s, s_map, mask, label, rgb_path):
        """late-fusion head + result bookkeeping (multiple_features.py:986-1003)"""
        s = torch.tensor(self.detect_fuser.score_samples(s))
        s_map = torch.tensor(self.seg_fuser.score_samples(s_map))
        s_map = s_map.view(1, self.gt_size, self.gt_size)
        mask = torch.as_tensor(mask)
        self.image_preds.append(s.numpy())
        self.image_labels.append(label)
        # the reference extends two Python lists by 50 176 scalars per image (multiple_features.py:998-999); the same
        # values are kept as one array per image here and concatenated once in calculate_metrics
        self.pixel_preds.append(s_map.flatten().numpy())
        self.pixel_labels.append(mask.flatten().numpy())
        self.predictions.append(s_map.detach().cpu().squeeze().numpy())
        self.gts.append(mask.detach().cpu().squeeze().numpy())
        self.img_name.append(rgb_path)

    # ---- features.py:302-324 (vectorised on the host, cmdiad_b200/metrics.py) ------------------------
    def calculate_metrics(self):
        from . import metrics
        self.image_preds = np.stack(self.image_preds)
        self.image_labels = np.stack(self.image_labels)
        self.img_name = np.stack(self.img_name)
        if self.args.save_raw_results:  # features.py:316-318
            import os
            d = f"./visualization/{getattr(self.args, 'experiment_note', '')}"
            os.makedirs(d, exist_ok=True)
            txt_to_save = np.concatenate((self.image_preds, self.image_labels, self.img_name), axis=1)
            np.savetxt(f"{d}/{self.class_name}_raw_results.csv", txt_to_save, delimiter=",", fmt="%s")
        self.pixel_preds = np.concatenate(self.pixel_preds) if len(self.pixel_preds) else np.zeros(0)
        self.pixel_labels = np.concatenate(self.pixel_labels) if len(self.pixel_labels) else np.zeros(0)
        self.image_rocauc = roc_auc_score(self.image_labels, self.image_preds)
        self.pixel_rocauc = roc_auc_score(self.pixel_labels, self.pixel_preds)
        self.au_pro, _ = metrics.au_pro(self.gts, self.predictions)            # features.py:323
        self.au_pro_001, _ = metrics.au_pro(self.gts, self.predictions, 0.01)  # features.py:324

    def calculate_metrics_device(self):
        """calculate_metrics (features.py:302-324) with the pixel-level part on the GPU: the fused maps were kept in the
        device-side result store by predict_batch(..., keep_on_device=True); one radix sort of all (score, label) pairs
        yields the pixel AUROC and the counts behind both AU-PRO values (cmdb_eval_pixel_metrics).  au_pro / au_pro_001
        are bit-identical to the host path, pixel_rocauc agrees to float64 rounding (exact U statistic vs sklearn's
        trapezoids)."""
        from . import metrics
        self.image_preds = np.stack(self.image_preds)
        self.image_labels = np.stack(self.image_labels)
        self.image_rocauc = roc_auc_score(self.image_labels, self.image_preds)
        r = metrics.device_pixel_metrics(self.fusion(), self.gts)
        self.pixel_rocauc = r["pixel_rocauc"]
        self.au_pro, self.au_pro_001 = r["au_pro"][0.3], r["au_pro"][0.01]
        return r

    # ---- persistence of the fitted state (banks after run_coreset, statistics, late-fusion head) ----------------------
    def save_state(self, directory):
        """after run_coreset (and optionally run_late_fusion): one .npz per bank + the scalars + the parameters of the two
        linear One-Class-SVM heads; load_state restores a ready-to-predict object without re-running the coreset selection.
        Plain arrays only -- nothing in a state directory is unpickled."""
        import os
        os.makedirs(directory, exist_ok=True)
        for m in self.bank_modals:
            self._banks[m].save(os.path.join(directory, f"bank_{m}.npz"), mean=np.float32(getattr(self, f"{m}_mean")),
                                std=np.float32(getattr(self, f"{m}_std")))
        heads = {}
        for name, f in (("detect", self.detect_fuser), ("seg", self.seg_fuser)):
            if hasattr(f, "coef_"):
                heads[f"{name}_coef"] = np.asarray(f.coef_, np.float64)
                heads[f"{name}_offset"] = np.asarray(f.offset_, np.float64)
        if self.coreset_idx is not None:
            heads["coreset_idx"] = np.asarray(self.coreset_idx, np.int64)
        np.savez(os.path.join(directory, "heads.npz"), **heads)

    def load_state(self, directory):
        import os
        for m in self.bank_modals:
            bank, meta = Bank.load(os.path.join(directory, f"bank_{m}.npz"), device=self.cuda_device,
                                   finalize=m in self._score_modals())
            self._banks[m] = bank
            if m in self._score_modals():
                bank.build_knn()   # not stored: rebuilding the neighbour table takes milliseconds
            setattr(self, f"{m}_mean", torch.tensor(np.float32(meta["mean"])))
            setattr(self, f"{m}_std", torch.tensor(np.float32(meta["std"])))
            setattr(self, f"patch_{m}_lib", DeviceLib(bank))
        with np.load(os.path.join(directory, "heads.npz"), allow_pickle=False) as d:
            for name, f in (("detect", self.detect_fuser), ("seg", self.seg_fuser)):
                if f"{name}_coef" in d.files:  # score_samples is linear: coef_ / offset_ are the whole fitted state
                    f.coef_ = d[f"{name}_coef"].copy()
                    f.offset_ = d[f"{name}_offset"].copy()
                    f.n_features_in_ = int(f.coef_.reshape(-1).shape[0])
            self.coreset_idx = torch.from_numpy(d["coreset_idx"].copy()) if "coreset_idx" in d.files else None
        self._fusion = None

    def close(self):
        for b in self._banks.values():
            b.close()
        self._banks = {}


class RGBFeatures(Features):
    """multiple_features.py:28-121"""
    bank_modals = ("rgb",)
    score_modals = ("rgb",)
    mean_from = {"rgb": "rgb"}
    std_from = {"rgb": "rgb"}


class DepthFeatures(RGBFeatures):
    """multiple_features.py:124-204: identical bank logic, the depth image feeds the RGB backbone upstream."""


class PointFeatures(Features):
    """multiple_features.py:207-309"""
    bank_modals = ("xyz",)
    score_modals = ("xyz",)
    mean_from = {"xyz": "xyz"}
    std_from = {"xyz": "xyz"}


class DoubleRGBPointFeatures(Features):
    """multiple_features.py:800-1015.  The statistics are cross-wired in the reference (:877-880) and kept so:
    xyz_mean = rgb_mean = mean(xyz lib); xyz_std = rgb_std = std(rgb lib)."""
    bank_modals = ("xyz", "rgb")
    score_modals = ("xyz", "rgb")
    mean_from = {"xyz": "xyz", "rgb": "xyz"}
    std_from = {"xyz": "rgb", "rgb": "rgb"}


class RGBorXYZWithOneHallucination(Features):
    """multiple_features.py:312-573: main modality bank + hallucinated ("fusion") bank.  All three libraries are
    filled, only main + fusion are normalised / subsampled / scored (:379-402, :533-547); statistics cross-wired as in
    :372-377."""
    bank_modals = ("xyz", "rgb", "fusion")
    mean_from = {"xyz": "xyz", "rgb": "xyz", "fusion": "xyz"}
    std_from = {"xyz": "rgb", "rgb": "rgb", "fusion": "rgb"}

    def _main(self):
        if self.args.main_modality not in ("rgb", "xyz"):
            raise ValueError(f"main_modality must be 'rgb' or 'xyz', got {self.args.main_modality!r}")
        return self.args.main_modality

    def _normalised_modals(self):
        return (self._main(), "fusion")

    def _coreset_modals(self):
        return (self._main(), "fusion")

    def _score_modals(self):
        return (self._main(), "fusion")


class RGBorXYZWithOneHallucinationFromFeature(RGBorXYZWithOneHallucination):
    """multiple_features.py:576-797: same bank logic, the fusion patches come from a feature-level network."""


METHODS = {  # cmdiad_runner.py:16-31
    "DINO": RGBFeatures,
    "Point_MAE": PointFeatures,
    "DINO+Point_MAE": DoubleRGBPointFeatures,
    "WithHallucination": RGBorXYZWithOneHallucination,
    "WithHallucinationFromFeature": RGBorXYZWithOneHallucinationFromFeature,
}
