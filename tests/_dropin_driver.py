"""Subprocess driver of tests/test_integration_patch.py:  python tests/_dropin_driver.py <reference root> <out.npz> <cls>

Imports the reference found at <reference root> (patched or not) through oracle/ref_loader.py, replaces only the
backbone (`Model`, which needs timm + checkpoints) by a deterministic synthetic feature generator, constructs the method
class through its REAL __init__ and drives the REAL public API the runner calls (cmdiad_runner.py:44-92):
add_sample_to_mem_bank -> run_coreset -> add_sample_to_late_fusion_mem_bank -> run_late_fusion -> predict ->
calculate_metrics.  With the patched tree, cmdiad_b200.dropin is active and tests.checker_bank.CheckerBank (CPU) stands
in for the device bank."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ref_root, out_path, cls_name = sys.argv[1], sys.argv[2], sys.argv[3]
os.environ["CMDIAD_REFERENCE_ROOT"] = ref_root

from cmdiad_b200 import synth  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

F, MF = R.load_reference()
patched = getattr(F, "_b200", None) is not None
if patched:
    from cmdiad_b200 import dropin
    from tests.checker_bank import CheckerBank
    dropin.BankClass = CheckerBank

N_GROUP, D = 16, 768


class FakeModel(torch.nn.Module):
    """stands in for models.models.Model: features are a deterministic function of the image's first pixel (= sample id)"""

    def __init__(self, **kw):
        super().__init__()

    def forward(self, rgb=None, xyz=None, out_type="rgb+xyz"):
        sid = int(round(float(rgb[0, 0, 0, 0]) * 1000))
        k = 48
        rgb_f = torch.from_numpy(synth.patches(784, D, 100 + sid, k=k, anomalous_frac=0.02 if sid >= 50 else 0.0)).T.reshape(1, D, 28, 28).contiguous()
        xyz_f = torch.from_numpy(synth.patches(N_GROUP, D, 200 + sid, k=k)).T.reshape(1, D, N_GROUP).contiguous()
        g = np.random.Generator(np.random.PCG64(300 + sid))
        n_pts = xyz.shape[2]
        center_idx = torch.from_numpy(g.choice(n_pts, N_GROUP, replace=False)).view(1, N_GROUP, 1)
        center = xyz[:, :, center_idx[0, :, 0]].permute(0, 2, 1).contiguous()
        ori_idx = torch.zeros(1, N_GROUP, 4, 1, dtype=torch.long)
        return rgb_f, xyz_f, center, ori_idx, center_idx


F.Model = FakeModel


def sample(sid):
    img = torch.full((1, 3, 224, 224), sid / 1000.0)
    g = np.random.Generator(np.random.PCG64(400 + sid))
    pc = torch.from_numpy(g.random((1, 3, 224, 224), dtype=np.float32) + 0.1)
    pc[:, :, :20, :] = 0  # background points are dropped (multiple_features.py:10-25)
    return [img, pc, pc.clone()]


args = R.default_args(coreset_dtype="TF32", random_state=0, f_coreset=0.1, main_modality="rgb")
for k, v in dict(rgb_backbone_name="vit_base_patch8_224_dino", xyz_backbone_name="Point_MAE", group_size=128, num_group=N_GROUP,
                 rgb_size=224, xyz_size=224, use_hn=False, use_hn_conv=False, use_hn_from_rgb_mlp=False,
                 use_hn_from_rgb_conv=False, use_hrnet=False, fusion_module_path="", max_sample=4, experiment_note="t").items():
    setattr(args, k, v)
torch.manual_seed(0)
with R.cuda_to_cpu_if_needed():
    m = getattr(MF, cls_name)(args)                     # the real constructor (patched: ends with dropin.attach)
    train = [sample(i) for i in range(4)]
    for s in train:
        m.add_sample_to_mem_bank(s, class_name="synthetic")       # cmdiad_runner.py:46
    m.run_coreset()                                                # :54
    for s in train:
        m.add_sample_to_late_fusion_mem_bank(s)                    # :61
    m.run_late_fusion()                                            # :69
    for i, lab in ((50, 1), (51, 0)):
        mask = torch.zeros(1, 224, 224)
        if lab:
            mask[0, 100:120, 90:130] = 1
        m.predict(sample(i), mask, lab, [f"img{i}.png"])          # :84
    m.calculate_metrics()                                          # :88
out = dict(patched=np.array(patched), coreset_idx=m.coreset_idx.numpy(), image_preds=np.asarray(m.image_preds),
           predictions=np.stack(m.predictions), pixel_rocauc=np.float64(m.pixel_rocauc), au_pro=np.float64(m.au_pro),
           rgb_mean=np.float32(m.rgb_mean), rgb_std=np.float32(m.rgb_std))
for name in ("rgb", "xyz"):
    lib = getattr(m, f"patch_{name}_lib")
    if hasattr(lib, "shape") and len(lib) > 0:
        out[f"lib_{name}_shape"] = np.array(tuple(lib.shape))
        rows = lib.cpu() if patched and hasattr(lib, "store") else lib
        out[f"lib_{name}_head"] = np.asarray(rows[:5])
        if patched and hasattr(lib, "store"):
            out[f"calls_{name}"] = np.array(",".join(lib.store.bank.calls))
np.savez(out_path, **out)
print("driver done", cls_name, "patched" if patched else "stock", out["coreset_idx"][:5])
