"""GPU tests of the device-side late-fusion head (cmdb_score_fused_batch*, multiple_features.py:986-994),
the device-side query normalisation, the result store, and the stream-ordering contract of device inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(built):
    from cmdiad_b200 import synth
    assert torch.cuda.is_available()
    return dict(synth=synth)


def _fit(env, cls, args, samples, cap):
    m = cls(args, bank_capacity_rows=cap)
    for x in samples:
        m.add_sample_to_mem_bank(x)
    m.run_coreset()
    m.add_samples_to_late_fusion_mem_bank(samples)
    m.run_late_fusion()
    return m


def _ulp_diff(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 1e-300)


def test_fused_head_matches_sklearn_single_modality(env):
    """RGBFeatures.predict with the head on the device == the reference's host sequence (lambda * map -> float64 ->
    SGDOneClassSVM.score_samples): one column, so X @ coef_ is a single product and the result must be bit-identical"""
    from cmdiad_b200 import RGBFeatures, default_args
    sy = env["synth"]
    train = [{"rgb": x} for x in sy.image_bank(4, 784, 768, 71, k=64)]
    m = _fit(env, RGBFeatures, default_args(coreset_dtype="TF32", random_state=0), train, 4 * 784)
    tests = [{"rgb": sy.patches(784, 768, 900 + i, anomalous_frac=0.02, k=64)} for i in range(3)]
    masks = [torch.zeros(1, 224, 224) for _ in tests]
    m.device_head = True
    for t, k in zip(tests, masks):
        m.predict(t, k, 0, ["a.png"])
    dev_s, dev_map = [x.copy() for x in m.image_preds], [x.copy() for x in m.predictions]
    m.image_preds, m.predictions, m.pixel_preds, m.pixel_labels, m.image_labels, m.gts, m.img_name = [], [], [], [], [], [], []
    m.device_head = False
    for t, k in zip(tests, masks):
        m.predict(t, k, 0, ["a.png"])
    for i in range(3):
        assert (dev_s[i] == m.image_preds[i]).all(), (dev_s[i], m.image_preds[i])
        assert dev_map[i].dtype == np.float64 and (dev_map[i] == m.predictions[i]).all()
    # batch of 3 == one by one; query normalisation on the device == the host expression
    m.image_preds, m.predictions = [], []
    m.device_head = True
    m.predict_batch(tests, masks, [0, 0, 0], [["a.png"]] * 3)
    for i in range(3):
        assert (m.image_preds[i] == dev_s[i]).all() and (m.predictions[i] == dev_map[i]).all()
    m.close()


def test_fused_head_two_modalities_and_store(env):
    """DoubleRGBPointFeatures: two banks with different P, [npix, 2] @ coef_ in float64.  sklearn's BLAS order is
    fma(x0, c0, x1 * c1) on the build machine; the device uses that order, so the maps are normally bit-identical and in
    any case within 2 ulp of float64.  Also: results kept on the device == results returned to the host."""
    from cmdiad_b200 import DoubleRGBPointFeatures, default_args
    sy = env["synth"]
    rgb = sy.image_bank(3, 784, 768, 81, k=64)
    xyz = [x * 1.5 + 0.25 for x in sy.image_bank(3, 3136, 768, 82, k=64)]
    train = [{"rgb": r, "xyz": x} for r, x in zip(rgb, xyz)]
    m = _fit(env, DoubleRGBPointFeatures, default_args(coreset_dtype="TF32", random_state=0), train, 3 * 3136)
    tests = [{"rgb": sy.patches(784, 768, 910 + i, anomalous_frac=0.02, k=64),
              "xyz": sy.patches(3136, 768, 920 + i, anomalous_frac=0.02, k=64) * 1.5 + 0.25} for i in range(4)]
    masks = [torch.zeros(1, 224, 224) for _ in tests]
    m.fusion().eval_reserve(8)
    m.predict_batch(tests, masks, [0, 1, 0, 1], [["a.png"]] * 4, keep_on_device=True)
    dev_s, dev_map = np.stack(m.image_preds), np.stack(m.predictions)
    kept_maps, kept_s = m.fusion().eval_read()
    assert m.fusion().eval_count() == 4 and (kept_maps == dev_map).all() and (kept_s == dev_s[:, 0]).all()
    m.image_preds, m.predictions = [], []
    m.device_head = False
    for t, k in zip(tests, masks):
        m.predict(t, k, 0, ["a.png"])
    host_s, host_map = np.stack(m.image_preds), np.stack(m.predictions)
    assert _ulp_diff(dev_s, host_s).max() <= 2 and _ulp_diff(dev_map, host_map).max() <= 2
    frac = float((dev_map == host_map).mean())
    print(f"fused two-modality maps bit-identical to sklearn on {frac:.4%} of pixels; image scores equal: {(dev_s == host_s).all()}")
    assert frac > 0.5
    m.close()


def test_device_inputs_are_ordered_after_their_producer(env):
    """a CUDA tensor that is still being written on torch's stream when score_batch / append is called.  The
    wrapper makes the handle's stream wait for the producer stream; without that the kernels would read zeros."""
    from cmdiad_b200 import Bank
    sy = env["synth"]
    lib = sy.patches(3000, 768, 5, k=64)
    patches = torch.from_numpy(np.stack([sy.patches(784, 768, 930 + i, anomalous_frac=0.01, k=64) for i in range(2)]))
    bank = Bank(768, 3000)
    bank.append(lib)
    bank.finalize()
    want = bank.score_batch(patches, (28, 28), 224)
    src = patches.cuda()
    torch.cuda.synchronize()
    late = torch.zeros_like(src)
    torch.cuda._sleep(400_000_000)   # ~0.2 s of busy waiting on torch's current stream ...
    late.copy_(src)                  # ... before the data lands
    got = bank.score_batch(late, (28, 28), 224)
    for i in range(2):
        assert (got[i].min_idx == want[i].min_idx).all() and (got[i].s_map == want[i].s_map).all()
    # same for append
    rows = torch.zeros(3000, 768, device="cuda")
    torch.cuda._sleep(400_000_000)
    rows.copy_(torch.from_numpy(lib).cuda())
    b2 = Bank(768, 3000)
    b2.append(rows)
    assert (b2.read().numpy() == lib).all()
    bank.close()
    b2.close()


def test_bank_capacity_admits_max_sample_plus_one(env):
    """the reference's fit loop appends max_sample + 1 samples (cmdiad_runner.py:46-52)"""
    from cmdiad_b200 import RGBFeatures, default_args
    m = RGBFeatures(default_args(max_sample=2))
    for i in range(3):
        m.add_sample_to_mem_bank({"rgb": env["synth"].patches(196, 768, 940 + i, k=64)})
    assert m._banks["rgb"].rows == 3 * 196
    m.close()


def test_device_side_metrics_match_the_host_path(env):
    """fused maps kept on the device, pixel AUROC + AU-PRO from one radix sort of (score, label) pairs
    (cmdb_eval_pixel_metrics) against calculate_metrics on the host (sklearn roc_auc_score + the vectorised rewrite of
    utils/au_pro_util.py, itself bit-identical to the reference module -- tests/test_metrics.py)."""
    from cmdiad_b200 import RGBFeatures, default_args
    sy = env["synth"]
    train = [{"rgb": x} for x in sy.image_bank(4, 196, 768, 91, k=32)]
    m = _fit(env, RGBFeatures, default_args(coreset_dtype="TF32", random_state=0, f_coreset=0.25), train, 4 * 196)
    n_test = 12
    g = np.random.Generator(np.random.PCG64(5))
    tests, masks, labels = [], [], []
    for i in range(n_test):
        tests.append({"rgb": sy.patches(196, 768, 950 + i, anomalous_frac=0.05 if i % 2 else 0.0, k=32)})
        mask = np.zeros((224, 224), np.float32)
        for _ in range(i % 3):
            cy, cx, r = g.integers(20, 200), g.integers(20, 200), g.integers(3, 15)
            yy, xx = np.ogrid[:224, :224]
            mask[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1
        if i == 4:
            mask[5, 5] = mask[6, 6] = 1   # two pixels touching diagonally: one component under 8-connectivity
        masks.append(torch.from_numpy(mask).view(1, 224, 224))
        labels.append(int(mask.any()))
    m.fusion().eval_reserve(n_test)
    m.predict_batch(tests, masks, labels, [["t.png"]] * n_test, keep_on_device=True)
    dev = m.calculate_metrics_device()
    img_auc_dev = m.image_rocauc
    au_dev, au001_dev, pix_dev = m.au_pro, m.au_pro_001, m.pixel_rocauc
    m.calculate_metrics()   # host path over the same per-image arrays
    assert m.au_pro == au_dev and m.au_pro_001 == au001_dev, (m.au_pro, au_dev, m.au_pro_001, au001_dev)
    assert abs(m.pixel_rocauc - pix_dev) <= 1e-12 and m.image_rocauc == img_auc_dev
    assert dev["n_pos"] == int(sum(float(k.sum()) for k in masks)) and dev["n_pos"] + dev["n_neg"] == n_test * 224 * 224
    m.close()
