"""Host-side mirror of the reference's entry points (config.json, image2label.train/evaluate,
NiftiDataset3D patch interface, checkpoints), driven through the CPU-emulated engine."""
import json
import os

import numpy as np
import pytest

from oracle import ref_vnet as R
from vnet_tensorflow_b200 import checkpoint, config as config_mod, nifti
from vnet_tensorflow_b200.model import image2label
from vnet_tensorflow_b200.pipeline import NiftiDataset3D


def _config(tmp, **over):
    cfg = {
        "TrainingSetting": {
            "Data": {"TrainingDataDirectory": str(tmp / "train"), "TestingDataDirectory": str(tmp / "test"),
                     "ImageFilenames": ["image.nii"], "LabelFilename": "label.nii"},
            "Restore": False, "SegmentationClasses": [0, 1], "LogDir": str(tmp / "log"), "CheckpointDir": str(tmp / "ckpt"),
            "BatchSize": 2, "PatchShape": [8, 8, 8], "ImageLog": False, "Testing": True, "TestStep": 2, "Epoches": 2,
            "MaxIterations": 100, "LogInterval": 2, "Precision": "fp32", "Synthetic": True,
            "Networks": {"Name": "VNet", "Dropout": 0.01, "NumChannel": 4, "NumLevels": 2, "NumCovolutions": [1, 2],
                         "BottomConvolutions": 1},
            "Loss": {"Name": "weighted_sorensen", "Weights": [0.1, 1], "Alpha": 1},
            "Optimizer": {"Name": "Adam", "InitialLearningRate": 1e-2, "Decay": {"Factor": 0.99, "Steps": 100}},
            "Spacing": [1, 1, 1], "DropRatio": 0.01, "MinPixel": 1,
        },
        "EvaluationSetting": {
            "Data": {"EvaluateDataDirectory": str(tmp / "eval"), "ImageFilenames": ["image.nii"],
                     "LabelFilename": "label_out.nii.gz", "ProbabilityFilename": "prob_out.nii.gz"},
            "CheckpointPath": str(tmp / "ckpt" / "checkpoint-8"), "Stride": [4, 4, 4], "BatchSize": 2,
            "ProbabilityOutput": True,
        },
    }
    cfg["TrainingSetting"].update(over)
    return cfg


def test_reference_config_files_load_with_tolerant_reader():
    """the shipped JSONs use `NumCovolutions` and lack keys the reference code reads."""
    shipped = {
        "TrainingSetting": {"Data": {"ImageFilenames": ["image.nii"], "LabelFilename": "label.nii"},
                            "SegmentationClasses": [0, 1, 2], "BatchSize": 32, "PatchShape": [64, 64, 64],
                            "Networks": {"Name": "VNet", "Dropout": 0.01, "NumChannel": 16, "NumLevels": 4,
                                         "NumCovolutions": [1, 2, 3, 3], "BottomConvolutions": 3},
                            "Loss": {"Name": "weighted_sorensen", "Weights": [0.01, 0.1, 1], "Alpha": 1},
                            "Optimizer": {"Name": "Adam", "InitialLearningRate": 1e-2, "Decay": {"Factor": 0.99, "Steps": 100}}},
        "EvaluationSetting": {"Data": {}, "CheckpointPath": "./tmp/ckpt/checkpoint-23125", "Stride": [64, 64, 64], "BatchSize": 10},
    }
    c = config_mod.from_dict(shipped)
    assert c.num_convolutions == (1, 2, 3, 3) and c.output_channel_num == 3 and c.evaluate_lcc is False
    shipped["TrainingSetting"]["Networks"]["Name"] = "UNet"
    with pytest.raises(SystemExit):
        config_mod.from_dict(shipped)


def _check_event_files(log_dir, losses):
    """LogDir/{train,test} hold TensorBoard event files (tf.summary.FileWriter, model.py:705-709) that TensorBoard's
    own reader accepts: version record first, then one scalar event per step under the reference's tags."""
    loader_mod = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")
    files = [f for f in os.listdir(log_dir / "train") if f.startswith("events.out.tfevents.")]
    assert len(files) == 1 and any(f.startswith("events.out.tfevents.") for f in os.listdir(log_dir / "test"))
    evs = list(loader_mod.LegacyEventFileLoader(str(log_dir / "train" / files[0])).Load())
    assert evs[0].file_version == "brain.Event:2" and len(evs) == 1 + len(losses)
    for k, ev in enumerate(evs[1:]):
        vals = {v.tag: v.simple_value for v in ev.summary.value}
        assert ev.step == k + 1 and ev.wall_time > 1e9
        assert vals["loss/0.total_loss"] == np.float32(losses[k])
        assert vals["learning_rate"] == np.float32(1e-2 * 0.99 ** (k / 100))   # exponential_decay at the step's start


def test_train_checkpoint_restore_evaluate_roundtrip(emul_lib, tmp_path):
    cfg = _config(tmp_path)
    m = image2label(None, cfg, library=emul_lib)
    m.train()
    assert m.engine.global_step == 8  # 2 epochs x 4 batches (synthetic set = 4*BatchSize patches)
    assert os.path.exists(tmp_path / "ckpt" / "checkpoint-latest")
    assert checkpoint.latest(str(tmp_path / "ckpt")).endswith("checkpoint-8")
    losses = [json.loads(l)["total_loss"] for l in open(tmp_path / "log" / "train" / "scalars.jsonl")]
    assert len(losses) == 8 and all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert os.path.exists(tmp_path / "log" / "test" / "scalars.jsonl")
    # the tf.metrics block of summary_op (model.py:586-626) is logged with every step under the reference's tags
    rows = [json.loads(l) for l in open(tmp_path / "log" / "train" / "scalars.jsonl")]
    for tag in ("metrics/accuracy", "metrics/sensitivity_1", "metrics/specificity_1", "metrics/dice_1", "metrics/auc_1"):
        assert all(tag in r for r in rows), tag
        assert all(r[tag] is None or 0.0 <= r[tag] <= 1.0 + 1e-6 for r in rows), tag
    _check_event_files(tmp_path / "log", losses)
    with np.load(str(tmp_path / "ckpt" / "checkpoint-8.npz")) as z:  # TF variable names + Adam slots
        assert "vnet/encoder/level_1/conv_1/weights" in z and "vnet/encoder/level_1/conv_1/weights/Adam_1" in z
        assert int(z["global_step"]) == 8
    # evaluation on a 12x10x8 volume: windows (stride 4, last clamped), softmax accumulation, argmax
    case = tmp_path / "eval" / "case0"
    os.makedirs(case)
    rng = np.random.default_rng(0)
    vol = rng.uniform(0, 255, (12, 10, 8)).astype(np.float32)
    nifti.write(str(case / "image.nii"), nifti.Image(vol, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)))
    m2 = image2label(None, cfg, library=emul_lib)
    m2.evaluate()
    out = nifti.read(str(case / "label_out.nii.gz"))
    assert out.array.shape == (12, 10, 8) and set(np.unique(out.array)) <= {0, 1}
    prob = nifti.read(str(case / "prob_out_1.nii.gz")).array
    assert prob.shape == (12, 10, 8) and prob.min() >= 0 and prob.max() <= 1.0 + 1e-5
    # the evaluator's window grid equals the reference arithmetic (model.py:866-892)
    assert R.window_starts(12, 8, 4) == [0, 4] and R.window_starts(10, 8, 4) == [0, 2]
    lab, sm, w = m2.evaluate_single_3D(vol[..., None])
    # overlap counts of the 2x2x1 windows, with the reference's quirk: the last batch is on the work list twice
    # (model.py:898-904), so its windows count double
    from oracle import ref_eval as _re
    w_plain = _re.evaluate_volume(vol[..., None], m2.patch_shape, m2.evaluate_stride, m2.evaluate_batch, m2.output_channel_num,
                                  lambda x: np.zeros(x.shape[:4] + (m2.output_channel_num,), np.float32), replay_last_batch=False)[2]
    assert w_plain.max() == 4 and w_plain.min() == 1
    assert w.min() >= 1 and w.max() > w_plain.max() - 1 and (w - w_plain).max() >= 1 and w.sum() > w_plain.sum()
    assert np.array_equal(lab, np.argmax(sm, -1))
    # the device window loop (vnb_evaluate_volume) against the oracle restatement of model.py:866-937 fed by
    # vnb_forward: same windows, same batches, same order of additions -> bit-exact sums, weights and labels
    from oracle import ref_eval
    P, S, B = m2.patch_shape, m2.evaluate_stride, m2.evaluate_batch
    assert ref_eval.window_starts((12, 10, 8), P, S)[-1] == (12 - P[0], 10 - P[1], 8 - P[2])
    lab_o, sm_o, w_o = ref_eval.evaluate_volume(vol[..., None], P, S, B, m2.output_channel_num,
                                                lambda x: m2.engine.forward(x, want_logits=False, want_argmax=False)[1])
    assert np.array_equal(sm, sm_o) and np.array_equal(w, w_o) and np.array_equal(lab, lab_o)


def test_restore_continues_from_latest_checkpoint(emul_lib, tmp_path):
    cfg = _config(tmp_path, Epoches=1)
    image2label(None, cfg, library=emul_lib).train()
    cfg2 = _config(tmp_path, Epoches=2, Restore=True)
    m = image2label(None, cfg2, library=emul_lib)
    m.train()
    assert m.engine.global_step == 8  # resumed at step 4 / epoch 1, ran one more epoch


def test_nifti_dataset_patch_contract(tmp_path):
    """NiftiDataset3D.NiftiDataset(...).get_dataset(): (float32 [X,Y,Z,M], int32 [X,Y,Z]) with labels remapped."""
    case = tmp_path / "0"
    os.makedirs(case)
    rng = np.random.default_rng(1)
    img = rng.normal(100, 20, (20, 18, 16)).astype(np.float32)
    lab = np.zeros((20, 18, 16), np.int16)
    lab[5:12, 4:10, 3:9] = 7
    nifti.write(str(case / "a.nii"), nifti.Image(img))
    nifti.write(str(case / "b.nii.gz"), nifti.Image(img * 2))
    nifti.write(str(case / "label.nii"), nifti.Image(lab))
    tfm = [NiftiDataset3D.StatisticalNormalization(2.5), NiftiDataset3D.Padding((24, 24, 24)),
           NiftiDataset3D.ConfidenceCrop2((16, 16, 16), rand_range=2, probability=1.0), NiftiDataset3D.RandomNoise(0.1)]
    ds = NiftiDataset3D.NiftiDataset(str(tmp_path), ["a.nii", "b.nii.gz"], "label.nii", tfm, train=True, labels=[0, 7]).get_dataset()
    image, label = next(iter(ds))
    assert image.shape == (16, 16, 16, 2) and image.dtype == np.float32
    assert label.shape == (16, 16, 16) and label.dtype == np.int32 and set(np.unique(label)) <= {0, 1}
    assert label.sum() > 0  # ConfidenceCrop2 centred on the labelled component
    assert -1 <= image.min() and image.max() <= 256


def test_parallel_patch_pipeline_keeps_order_and_propagates_errors(tmp_path):
    """get_dataset(num_parallel_calls=N): same patches in the same order as the sequential map; prefetch_iter hands
    items through unchanged and re-raises producer errors in the consumer."""
    rng = np.random.default_rng(2)
    for c in range(5):
        os.makedirs(tmp_path / str(c))
        img = rng.normal(50 + 10 * c, 5, (12, 10, 9)).astype(np.float32)
        lab = np.zeros((12, 10, 9), np.int16)
        lab[2:6, 3:7, 1:5] = 1
        nifti.write(str(tmp_path / str(c) / "image.nii"), nifti.Image(img))
        nifti.write(str(tmp_path / str(c) / "label.nii"), nifti.Image(lab))
    tfm = [NiftiDataset3D.ManualNormalization(0, 100), NiftiDataset3D.Padding((16, 16, 16))]
    mk = lambda: NiftiDataset3D.NiftiDataset(str(tmp_path), ["image.nii"], "label.nii", tfm, train=True, labels=[0, 1])
    seq = list(mk().get_dataset())
    par = list(mk().get_dataset(num_parallel_calls=3, prefetch=4))
    assert len(seq) == len(par) == 5
    for (a, la), (b, lb) in zip(seq, par):
        assert np.array_equal(a, b) and np.array_equal(la, lb)
    assert list(NiftiDataset3D.prefetch_iter(iter(range(7)), depth=2)) == list(range(7))

    def boom():
        yield 1
        raise ValueError("broken case")
    it = NiftiDataset3D.prefetch_iter(boom(), depth=2)
    assert next(it) == 1
    with pytest.raises(ValueError):
        next(it)
    with pytest.raises(ZeroDivisionError):
        list(NiftiDataset3D.parallel_map(lambda x: 1 // x, [1, 0, 2], workers=2))


def test_native_inference_driver_matches_the_python_evaluator(emul_lib, tmp_path):
    """cxx/vnb_infer.cpp (C++ over the C ABI, the counterpart of the reference's cxx/ program) on the emulated library:
    same label volume as image2label.evaluate_single_3D for the same weights, case, stride and batch."""
    from tests.conftest import EMUL_LIB
    _native_driver_case(emul_lib, EMUL_LIB, "fp32", tmp_path)


@pytest.mark.gpu
def test_native_inference_driver_on_the_gpu(gpu_lib, tmp_path):
    _native_driver_case(gpu_lib, gpu_lib.path, "bf16x3", tmp_path)


def _native_driver_case(lib, lib_path, precision, tmp_path):
    import subprocess
    from tests.conftest import ROOT
    from vnet_tensorflow_b200.engine import VNetEngine
    exe = tmp_path / "vnb_infer"
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "cxx", "vnb_infer.cpp"), "-ldl"], check=True)
    P, B, K = (8, 8, 8), 2, 3
    eng = VNetEngine(num_classes=K, in_channels=2, patch_shape=P, max_batch=B, num_channels=16, num_levels=2,
                     num_convolutions=(1, 1), bottom_convolutions=1, precision=precision, library=lib)
    spec = R.VNetSpec(num_classes=K, in_channels=2, num_channels=16, num_levels=2, num_convolutions=(1, 1), bottom_convolutions=1)
    eng.set_params(R.init_params(spec, 11))
    checkpoint.export_binary(eng, str(tmp_path / "w.vnbw"))
    rng = np.random.default_rng(5)
    a = rng.uniform(-1200, 1500, (11, 6, 9)).astype(np.float32)     # smaller than the patch along y: padded
    b = rng.integers(0, 255, (11, 6, 9)).astype(np.int16)
    nifti.write(str(tmp_path / "a.nii"), nifti.Image(a, (0.5, 0.5, 2.0), (1.0, 2.0, 3.0)))
    nifti.write(str(tmp_path / "b.nii"), nifti.Image(b))
    subprocess.run([str(exe), "--lib", lib_path, "--weights", str(tmp_path / "w.vnbw"), "--image", str(tmp_path / "a.nii"),
                    "--image", str(tmp_path / "b.nii"), "--out", str(tmp_path / "label.nii"), "--patch", "8", "8", "8",
                    "--stride", "3", "4", "8", "--batch", str(B), "--classes", str(K), "--labels", "0", "5", "9",
                    "--precision", precision, "--channels", "16", "--levels", "2", "--convs", "1", "1", "--bottom", "1"], check=True)
    out = nifti.read(str(tmp_path / "label.nii"))
    vol = np.stack([a, b.astype(np.float32)], -1)
    vol = np.pad(vol, [(0, 0), (0, 2), (0, 0), (0, 0)])
    lab, _, _ = eng.evaluate_volume(vol, (3, 4, 8), B)
    want = np.asarray([0, 5, 9], np.int32)[lab[:, :6, :]]
    assert out.array.shape == (11, 6, 9) and out.array.dtype == np.int32
    assert np.array_equal(out.array, want)
    assert out.spacing == (0.5, 0.5, 2.0) and out.origin == (1.0, 2.0, 3.0)
    eng.close()


def test_remaining_reference_transforms(tmp_path):
    """ExtremumNormalization, Reorient, ConfidenceCrop, BSplineDeformation (NiftiDataset3D.py:256-283,310-328,574-659,
    795-835) under the names the pipeline YAML resolves."""
    import random
    random.seed(3)
    np.random.seed(3)
    rng = np.random.default_rng(0)
    vol = rng.normal(100, 30, (40, 36, 32)).astype(np.float32)
    lab = np.zeros((40, 36, 32), np.int16)
    lab[10:14, 20:24, 5:9] = 3
    lab[30:33, 4:8, 20:26] = 1
    sample = {'image': [nifti.Image(vol, (1.0, 2.0, 0.5), (5.0, 6.0, 7.0))], 'label': nifti.Image(lab, (1.0, 2.0, 0.5), (5.0, 6.0, 7.0))}

    out = NiftiDataset3D.ExtremumNormalization(0.05)(sample)['image'][0].array
    lo, hi = vol.min() + 0.05 * (vol.max() - vol.min()), vol.min() + 0.95 * (vol.max() - vol.min())
    assert out.min() == 0 and out.max() == 255 and out.dtype == np.float32
    inside = (vol > lo) & (vol < hi)
    assert np.allclose(out[inside], (vol[inside] - lo) / (hi - lo) * 255, atol=1e-3)
    assert np.all(out[vol <= lo] == 0) and np.all(out[vol >= hi] == 255)

    out = NiftiDataset3D.Reorient([2, 0, 1])(sample)
    assert out['image'][0].array.shape == (32, 40, 36) and out['image'][0].spacing == (0.5, 1.0, 2.0)
    assert out['image'][0].origin == (7.0, 5.0, 6.0)
    assert out['image'][0].array[3, 1, 2] == vol[1, 2, 3] and out['label'].array[6, 11, 21] == 3

    crop = NiftiDataset3D.ConfidenceCrop((16, 12, 8), 0.25)
    hits = 0
    for _ in range(20):
        o = crop(sample)
        assert o['image'][0].array.shape == (16, 12, 8) and o['label'].array.shape == (16, 12, 8)
        hits += bool(o['label'].array.any())
    assert hits >= 15                                   # small sigma: the crop stays on the chosen component
    empty = {'image': sample['image'], 'label': nifti.Image(np.zeros_like(lab))}
    assert NiftiDataset3D.ConfidenceCrop(16, 0.1)(empty)['image'][0].array.shape == (16, 16, 16)
    assert NiftiDataset3D.ConfidenceCrop((40, 36, 32))(sample)['image'][0].array.shape == (40, 36, 32)  # only offset 0 fits
    with pytest.raises(ValueError, match="smaller than the crop"):
        NiftiDataset3D.ConfidenceCrop(48)(sample)

    bs = NiftiDataset3D.BSplineDeformation(4)
    n = bs.MESH + bs.ORDER
    # cubic B-splines sum to one and reproduce linear functions: constant coefficients = a rigid shift of c / spacing
    # voxels, coefficients equal to the control index = the continuous grid coordinate itself
    const = np.stack([np.full((n, n, n), c) for c in (2.0, 3.0, 1.0)])
    d = bs.displacement(vol.shape, (1.0, 2.0, 0.5), const)
    assert np.allclose(d[0], 2.0) and np.allclose(d[1], 1.5) and np.allclose(d[2], 2.0)
    ramp = np.stack([np.broadcast_to(np.arange(n, dtype=np.float64).reshape([-1 if a == c else 1 for a in range(3)]), (n, n, n))
                     for c in range(3)])
    d = bs.displacement(vol.shape, (1.0, 1.0, 1.0), ramp)
    assert np.allclose(d[0][:, 0, 0], np.arange(40) * 10 / 40 + 1) and np.allclose(d[2][0, 0, :], np.arange(32) * 10 / 32 + 1)
    o = bs(sample)
    assert o['image'][0].array.shape == vol.shape and o['image'][0].array.dtype == np.float32
    assert set(np.unique(o['label'].array)) <= {0, 1, 3} and o['label'].array.dtype == lab.dtype
    assert 0 < (o['label'].array > 0).sum() < 2 * (lab > 0).sum()
    assert not np.array_equal(o['image'][0].array, vol)
    with pytest.raises(RuntimeError):
        NiftiDataset3D.BSplineDeformation(0)
    # the YAML resolver finds every transform class the reference defines
    for name in ("Normalization", "RandomFlip", "StatisticalNormalization", "ExtremumNormalization", "ManualNormalization",
                 "Reorient", "Invert", "Resample", "Padding", "RandomCrop", "RandomNoise", "ConfidenceCrop",
                 "ConfidenceCrop2", "BSplineDeformation"):
        assert callable(getattr(NiftiDataset3D, name))


def test_resample_image_follows_itk_index_arithmetic():
    """sitk.ResampleImageFilter with the identity transform, as NiftiDataset3D.py:378-396 and model.py:957-975 set it."""
    rng = np.random.default_rng(0)
    a = rng.normal(size=(10, 12, 8)).astype(np.float32)
    im = nifti.Image(a, (1.0, 2.0, 0.5), (3.0, 4.0, 5.0))
    res = NiftiDataset3D.resample_image
    assert np.array_equal(res(im, im.spacing, im.GetSize(), im.origin, 1).array, a)
    assert np.array_equal(res(im, im.spacing, im.GetSize(), im.origin, 0).array, a)
    up = res(im, (0.5, 2.0, 0.5), (20, 12, 8), im.origin, 1).array
    assert np.array_equal(up[::2], a) and np.allclose(up[1:-1:2], (a[:-1] + a[1:]) / 2, atol=1e-6)
    assert not up[-1].any()                                 # continuous index 9.5 is outside [-0.5, 9.5)
    near = res(im, (0.5, 2.0, 0.5), (20, 12, 8), im.origin, 0).array
    assert np.array_equal(near[0::2], a) and np.array_equal(near[1:-1:2], a[1:])   # halves round up
    lab = nifti.Image((a > 0).astype(np.int16), im.spacing, im.origin)
    assert res(lab, (0.5, 2.0, 0.5), (20, 12, 8), im.origin, 0).array.dtype == np.int16
    # a shifted, larger output grid: zeros outside the input, input values where the grids coincide
    big = res(im, im.spacing, (14, 12, 8), (1.0, 4.0, 5.0), 1).array
    assert not big[:2].any() and np.array_equal(big[2:12], a) and not big[12:].any()
    # the Resample transform: grid size ceil(extent / voxel), same origin; Padding appends at the far end only
    s = NiftiDataset3D.Resample((2.0, 2.0, 2.0))({'image': [im], 'label': lab})
    assert s['image'][0].array.shape == (5, 12, 2) and s['label'].array.shape == (5, 12, 2) and s['image'][0].origin == im.origin
    assert np.array_equal(s['image'][0].array[:, :, 0], a[::2, :, 0])
    p = NiftiDataset3D.Padding((12, 12, 16))({'image': [im], 'label': lab})
    assert p['image'][0].array.shape == (12, 12, 16) and np.array_equal(p['image'][0].array[:10, :, :8], a)
    assert not p['image'][0].array[10:].any() and not p['image'][0].array[:, :, 8:].any() and p['image'][0].origin == im.origin
    c = NiftiDataset3D.RandomCrop((4, 4, 4))._crop(im, [2, 3, 1])
    assert c.origin == (5.0, 10.0, 5.5) and np.array_equal(c.array, a[2:6, 3:7, 1:5])


def test_label_postprocessing_matches_the_reference_filters():
    from vnet_tensorflow_b200.model import postprocess_label
    lab = np.zeros((12, 12, 12), np.int32)
    lab[1:4, 1:4, 1:4] = 2            # 27 voxels
    lab[6:10, 6:10, 6:10] = 1         # 64 voxels, touching the next block only by an edge (not face-connected)
    lab[10:12, 10:12, 5] = 3          # 4 voxels
    assert postprocess_label(lab, (1, 1, 1)) is lab
    lcc = postprocess_label(lab, (1, 1, 1), lcc=True)
    assert lcc.dtype == np.uint8 and lcc.sum() == 64 and lcc[7, 7, 7] == 1      # binary mask of the largest component
    vt = postprocess_label(lab, (1.0, 1.0, 0.5), volume_threshold=13.5)          # physical size, strict '>'
    assert vt.sum() == 64 and set(np.unique(vt)) == {0, 1}                        # 27 * 0.5 = 13.5 is not kept
    assert postprocess_label(lab, (1.0, 1.0, 0.5), volume_threshold=13.4).sum() == 64 + 27
    both = postprocess_label(lab, (1, 1, 1), lcc=True, volume_threshold=100.0)
    assert both.shape == lab.shape and not both.any()
    assert not postprocess_label(np.zeros((4, 4, 4), np.int32), (1, 1, 1), lcc=True).any()


def test_evaluate_with_a_pipeline_returns_to_the_input_grid(emul_lib, tmp_path):
    """model.py:817-977,1196-1243: evaluate transforms (Resample, Padding) run before the window loop; the label comes
    back on the input image's grid by nearest neighbour, the probabilities linearly; the label holds class indices."""
    import yaml
    pipe = {"preprocess": {"evaluate": {"3D": [{"name": "StatisticalNormalization", "variables": {"sigma": 2.5}},
                                                {"name": "Resample", "variables": {"voxel_size": [2.0, 2.0, 2.0]}},
                                                {"name": "Padding", "variables": {"output_size": [8, 8, 8]}}]}}}
    with open(tmp_path / "pipe.yaml", "w") as f:
        yaml.safe_dump(pipe, f)
    cfg = _config(tmp_path, Epoches=1, Testing=False, SegmentationClasses=[0, 4])
    cfg["EvaluationSetting"]["Pipeline"] = str(tmp_path / "pipe.yaml")
    cfg["EvaluationSetting"]["CheckpointPath"] = str(tmp_path / "ckpt" / "checkpoint-4")
    image2label(None, cfg, library=emul_lib).train()
    case = tmp_path / "eval" / "case0"
    os.makedirs(case)
    rng = np.random.default_rng(0)
    vol = rng.uniform(0, 255, (20, 14, 9)).astype(np.float32)          # 1 mm voxels -> 10 x 7 x 5 at 2 mm -> padded to 10 x 8 x 8
    nifti.write(str(case / "image.nii"), nifti.Image(vol, (1.0, 1.0, 1.0), (-3.0, 2.0, 9.0)))
    m = image2label(None, cfg, library=emul_lib)
    m.evaluate()
    out = nifti.read(str(case / "label_out.nii.gz"))
    assert out.array.shape == (20, 14, 9) and out.spacing == (1.0, 1.0, 1.0) and out.origin == (-3.0, 2.0, 9.0)
    assert set(np.unique(out.array)) <= {0, 1}                          # class indices, as the reference writes them
    prob = [nifti.read(str(case / ("prob_out_%d.nii.gz" % v))).array for v in (0, 4)]
    assert prob[0].shape == (20, 14, 9) and np.allclose((prob[0] + prob[1])[:19, :13, :9], 1.0, atol=1e-5)
    # the same numbers by hand: transforms, window loop, resample back
    sample = {'image': [nifti.Image(vol, (1.0, 1.0, 1.0), (-3.0, 2.0, 9.0))], 'label': nifti.Image(np.zeros(vol.shape, np.int32))}
    for t in m._transforms(str(tmp_path / "pipe.yaml"), "evaluate"):
        sample = t(sample)
    assert sample['image'][0].array.shape == (10, 8, 8)
    lab, sm, w = m.evaluate_single_3D(sample['image'][0].array[..., None])
    ref = sample['image'][0]
    back = NiftiDataset3D.resample_image(nifti.Image(lab.astype(np.int32), ref.spacing, ref.origin), (1.0, 1.0, 1.0), (20, 14, 9), (-3.0, 2.0, 9.0), 0)
    assert np.array_equal(back.array, out.array)
    assert np.array_equal(out.array[::2, ::2, ::2], lab[:10, :7, :5])   # even voxels coincide with the 2 mm grid
    # MapLabelValues (extension): SegmentationClasses values instead of indices
    cfg["EvaluationSetting"]["MapLabelValues"] = True
    image2label(None, cfg, library=emul_lib).evaluate()
    mapped = nifti.read(str(case / "label_out.nii.gz")).array
    assert np.array_equal(mapped, out.array * 4)


def test_transform_semantics_follow_the_simpleitk_filters():
    """Details of NiftiDataset3D.py the YAML pipelines rely on: constructor argument names, pixel-type casts of the
    intensity filters, the N-1 standard deviation, one flip coin, the crop index arithmetic."""
    import random
    rng = np.random.default_rng(0)
    ct = rng.normal(40, 300, (24, 20, 16)).astype(np.int16)
    lab = np.zeros(ct.shape, np.uint8)
    lab[4:8, 10:15, 6:9] = 1
    sample = {'image': [nifti.Image(ct)], 'label': nifti.Image(lab)}

    # StatisticalNormalization on an integer image: window bounds and outputs truncated to the pixel type
    out = NiftiDataset3D.StatisticalNormalization(2.5)(sample)['image'][0].array
    d = ct.astype(np.float64)
    lo, hi = int(d.mean() - 2.5 * d.std(ddof=1)), int(d.mean() + 2.5 * d.std(ddof=1))
    want = np.where(ct < lo, 0, np.where(ct > hi, 255, np.trunc((d - lo) * (255.0 / (hi - lo)))))
    assert out.dtype == np.int16 and np.abs(out - want).max() <= 1 and (out == want).mean() > 0.999
    assert out.min() == 0 and out.max() == 255
    f32 = NiftiDataset3D.StatisticalNormalization(2.5)({'image': [nifti.Image(ct.astype(np.float32))], 'label': nifti.Image(lab)})['image'][0].array
    assert f32.dtype == np.float32 and not np.array_equal(f32, np.trunc(f32)) and np.abs(f32 - out).max() < 1.5
    pre = NiftiDataset3D.StatisticalNormalization(2.5, pre_norm=True)(sample)['image'][0].array
    assert pre.dtype == np.float32 and np.abs(pre - f32).max() < 1e-2
    with pytest.raises(AssertionError):
        NiftiDataset3D.StatisticalNormalization(2)                       # the reference insists on a float sigma
    man = NiftiDataset3D.ManualNormalization(-100, 155)(sample)['image'][0].array
    assert man.dtype == np.int16 and man[ct <= -100].max() == 0 and man[ct >= 155].min() == 255
    k = (ct > -100) & (ct < 155)
    assert np.array_equal(man[k], ct[k] + 100)                           # scale 1: exact
    resc = NiftiDataset3D.Normalization()(sample)['image'][0].array
    assert resc.dtype == np.int16 and resc.min() == 0 and resc.max() == 255
    inv = NiftiDataset3D.Invert()({'image': [nifti.Image(man)], 'label': nifti.Image(lab)})['image'][0].array
    assert np.array_equal(inv, 255 - man)

    # RandomNoise(sigma=5): the YAML keyword is `sigma`
    np.random.seed(0)
    noisy = NiftiDataset3D.RandomNoise(sigma=5)({'image': [nifti.Image(ct.astype(np.float32))], 'label': nifti.Image(lab)})['image'][0].array
    assert 4.5 < (noisy - ct).std() < 5.5 and NiftiDataset3D.RandomNoise().sigma == 5
    assert NiftiDataset3D.RandomNoise(5)(sample)['image'][0].array.dtype == np.int16

    # RandomFlip: one coin for all marked axes
    np.random.seed(1)
    seen = set()
    for _ in range(16):
        o = NiftiDataset3D.RandomFlip([True, False, True])(sample)
        flipped = np.array_equal(o['image'][0].array, ct[::-1, :, ::-1]) and np.array_equal(o['label'].array, lab[::-1, :, ::-1])
        assert flipped or o['image'][0].array is ct
        seen.add(flipped)
    assert seen == {True, False}

    # RandomCrop: redraws until min_pixel foreground voxels are inside; never draws the last admissible start
    np.random.seed(2)
    random.seed(2)
    crop = NiftiDataset3D.RandomCrop((8, 8, 8), drop_ratio=0, min_pixel=20)
    starts = set()
    for _ in range(40):
        o = crop(sample)
        assert o['label'].array.shape == (8, 8, 8) and o['label'].array.sum() >= 20
        starts.add(tuple(int(v) for v in o['label'].origin))
    assert max(s[0] for s in starts) <= 24 - 8 - 1 and len(starts) > 3
    with pytest.raises(RuntimeError):
        NiftiDataset3D.RandomCrop(8, drop_ratio=1.5)

    # ConfidenceCrop2: bounding-box centre + integer offset, clamped with the reference's "- 1"
    c2 = NiftiDataset3D.ConfidenceCrop2((8, 8, 8), rand_range=0, probability=1.0)
    o = c2(sample)
    assert tuple(int(v) for v in o['label'].origin) == (4 + 2 - 4, 10 + 2 - 4, 6 + 1 - 4)  # start + int(extent/2) - int(size/2)
    edge = np.zeros(ct.shape, np.uint8)
    edge[22:24, 0:2, 14:16] = 1
    o = c2({'image': [nifti.Image(ct)], 'label': nifti.Image(edge)})
    assert tuple(int(v) for v in o['label'].origin) == (24 - 8 - 1, 0, 16 - 8 - 1)
    neg = NiftiDataset3D.ConfidenceCrop2((8, 8, 8), rand_range=(1, 2, 3), probability=0.0, random_empty_region=True)
    for _ in range(10):
        assert not neg(sample)['label'].array.any()
    tight = NiftiDataset3D.ConfidenceCrop2((24, 19, 16), probability=0.0)(sample)    # margins 0, 1, 0 -> start 0
    assert tight['label'].array.shape == (24, 19, 16)
    assert NiftiDataset3D.ConfidenceCrop2(8, probability=0.75).probability == 0.75


def test_labels_are_remapped_before_the_transforms(tmp_path):
    """NiftiDataset3D.py:119-137 runs before the transform loop: a crop never centres on a label value outside `labels`."""
    case = tmp_path / "0"
    os.makedirs(case)
    img = np.zeros((20, 20, 20), np.float32)
    lab = np.zeros((20, 20, 20), np.int16)
    lab[1:4, 1:4, 1:4] = 9            # not a segmentation class: background for the pipeline
    lab[14:18, 14:18, 14:18] = 7
    nifti.write(str(case / "image.nii"), nifti.Image(img))
    nifti.write(str(case / "label.nii"), nifti.Image(lab))
    nifti.write(str(tmp_path / ".DS_Store"), nifti.Image(img))          # skipped like NiftiDataset3D.py:41-45
    tfm = [NiftiDataset3D.ConfidenceCrop2((8, 8, 8), rand_range=0, probability=1.0)]
    ds = NiftiDataset3D.NiftiDataset(str(tmp_path), ["image.nii"], "label.nii", tfm, train=True, labels=[0, 7])
    assert len(ds.get_dataset()) == 1
    for _ in range(5):
        image, label = next(iter(ds.get_dataset()))
        assert label.dtype == np.int32 and label.sum() == 64 and set(np.unique(label)) == {0, 1}
    nifti.write(str(case / "label.nii"), nifti.Image(lab, (2.0, 1.0, 1.0)))
    with pytest.raises(Exception, match="Header info inconsistent"):
        next(iter(ds.get_dataset()))


def test_shipped_config_and_pipeline_produce_training_patches(tmp_path):
    """configs/config_b200_64.json + configs/pipeline3D_b200.yaml: every key loads, every transform resolves and the
    training pipeline turns a NIfTI case into a 64^3 patch with the (float32 [X,Y,Z,M], int32 [X,Y,Z]) contract."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = config_mod.load(os.path.join(root, "configs", "config_b200_64.json"))
    assert cfg.patch_shape == (64, 64, 64) and cfg.num_convolutions == (1, 2, 3, 3) and cfg.loss_weights == (0.1, 1.0)
    assert cfg.evaluate_stride == (32, 32, 32) and cfg.checkpoint_format == "npz" and cfg.data_workers == 4
    m = image2label.__new__(image2label)
    pipe = os.path.join(root, cfg.training_pipeline)
    names = {ph: [t.name for t in m._transforms(pipe, ph)] for ph in ("train", "test", "evaluate")}
    assert names["train"] == ['StatisticalNormalization', 'Resample', 'Padding', 'Confidence Crop 2', 'Flip', 'Random Noise']
    assert names["test"] == names["train"][:4] and names["evaluate"] == names["train"][:3]
    case = tmp_path / "case"
    os.makedirs(case)
    rng = np.random.default_rng(0)
    vol = rng.normal(30, 200, (50, 44, 30)).astype(np.int16)
    lab = np.zeros(vol.shape, np.uint8)
    lab[20:30, 10:20, 8:16] = 1
    nifti.write(str(case / "image.nii"), nifti.Image(vol, (1.0, 1.0, 1.5)))
    nifti.write(str(case / "label.nii"), nifti.Image(lab, (1.0, 1.0, 1.5)))
    ds = NiftiDataset3D.NiftiDataset(str(tmp_path), cfg.image_filenames, cfg.label_filename, m._transforms(pipe, "train"),
                                     train=True, labels=cfg.label_classes).get_dataset()
    image, label = next(iter(ds))
    assert image.shape == (64, 64, 64, 1) and image.dtype == np.float32 and label.shape == (64, 64, 64) and label.dtype == np.int32
    assert set(np.unique(label)) <= {0, 1} and -30 <= image.min() and image.max() <= 285


def test_last_evaluation_batch_is_replayed_like_the_reference():
    """model.py:898-904: the work list holds the last batch twice, so its windows enter the sums and the weights twice."""
    from oracle import ref_eval
    vol = np.arange(6 * 4 * 4, dtype=np.float32).reshape(6, 4, 4, 1)
    calls = []

    def softmax_fn(x):
        calls.append(x.shape[0])
        return np.ones(x.shape[:4] + (2,), np.float32) * np.array([0.25, 0.75], np.float32)
    # windows along x at 0, 1, 2 (patch 4, stride 1), batches of two: [0, 1], [2], and [2] again
    lab, sums, w = ref_eval.evaluate_volume(vol, (4, 4, 4), (1, 4, 4), 2, 2, softmax_fn)
    assert calls == [2, 1, 1]
    assert list(w[:, 0, 0]) == [1, 2, 4, 4, 3, 2]
    assert np.allclose(sums[..., 1], 0.75 * w) and np.array_equal(lab, np.ones((6, 4, 4), np.int64))
    _, _, w1 = ref_eval.evaluate_volume(vol, (4, 4, 4), (1, 4, 4), 2, 2, softmax_fn, replay_last_batch=False)
    assert list(w1[:, 0, 0]) == [1, 2, 3, 3, 2, 1]


def test_nifti_orientation_round_trip_and_itk_convention(tmp_path):
    """qform (quaternion + qfac) and sform are read into Image.direction / origin in ITK's LPS convention
    and written back, so an output volume keeps the input scan's orientation."""
    import struct
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    for flip in (False, True):
        R = q.copy()
        if (np.linalg.det(R) < 0) != flip:
            R[:, 2] = -R[:, 2]
        img = nifti.Image(rng.integers(0, 5, (4, 5, 6)).astype(np.int16), (0.5, 1.5, 2.0), (10.0, -20.0, 30.0), tuple(R.reshape(-1)))
        p = str(tmp_path / ("a%d.nii.gz" % flip))
        nifti.write(p, img)
        back = nifti.read(p)
        assert np.abs(np.array(back.direction) - np.array(img.direction)).max() < 1e-6
        assert np.allclose(back.origin, img.origin) and np.allclose(back.spacing, img.spacing) and np.array_equal(back.array, img.array)
        # sform-only file (qform_code = 0): same geometry
        raw = bytearray(open(p, "rb").read() if not p.endswith(".gz") else __import__("gzip").open(p, "rb").read())
        struct.pack_into("<h", raw, 252, 0)
        p2 = str(tmp_path / ("s%d.nii" % flip))
        open(p2, "wb").write(bytes(raw))
        s_only = nifti.read(p2)
        assert np.abs(np.array(s_only.direction) - np.array(img.direction)).max() < 1e-6 and np.allclose(s_only.origin, img.origin)
        assert np.allclose(s_only.spacing, img.spacing)
    # a plain RAS-identity NIfTI (what dcm2niix writes for an axial LAS->RAS scan) is ITK direction diag(-1, -1, 1)
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into("<h", hdr, 70, 2)
    struct.pack_into("<h", hdr, 72, 8)
    struct.pack_into("<8f", hdr, 76, 1.0, 1.0, 1.0, 1.0, 1, 1, 1, 1)
    struct.pack_into("<f", hdr, 108, 352.0)
    struct.pack_into("<h", hdr, 252, 1)
    struct.pack_into("<3f", hdr, 268, 5.0, 6.0, 7.0)
    hdr[344:348] = b"n+1\0"
    p3 = str(tmp_path / "ras.nii")
    open(p3, "wb").write(bytes(hdr) + bytes(8))
    ras = nifti.read(p3)
    assert ras.direction == (-1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 1.0) and ras.origin == (-5.0, -6.0, 7.0)
    # the header check of NiftiDataset now sees real directions
    other = nifti.Image(ras.array, ras.spacing, ras.origin)
    assert NiftiDataset3D.NiftiDataset._same_header(ras, other)[2] is False


def test_checkpoint_retention_follows_tf_saver(emul_lib, tmp_path):
    """tf.train.Saver(keep_checkpoint_every_n_hours=5) with the default max_to_keep=5 (model.py:676): older
    checkpoints are deleted, no temporary files stay behind, the state file lists what is kept."""
    from tests.helpers import engine_for
    from oracle import ref_vnet as R
    from vnet_tensorflow_b200 import checkpoint
    spec = R.VNetSpec(num_classes=2, in_channels=1, num_channels=4, num_levels=1, num_convolutions=(1,), bottom_convolutions=1)
    eng = engine_for(spec, 8, 1, "sorensen", (), emul_lib)
    d = str(tmp_path / "ckpt")
    for step in range(1, 9):
        checkpoint.save(eng, d, step, 0, "both" if step % 2 else "npz")
    files = sorted(os.listdir(d))
    assert not [f for f in files if ".tmp" in f]
    kept = sorted(int(f.split("-")[1].split(".")[0]) for f in files if f.endswith(".npz"))
    assert kept == [4, 5, 6, 7, 8]
    assert not os.path.exists(os.path.join(d, "checkpoint-3.index")) and os.path.exists(os.path.join(d, "checkpoint-5.index"))
    assert checkpoint.latest(d).endswith("checkpoint-8")
    state = open(os.path.join(d, checkpoint.LATEST)).read()
    assert state.count("all_model_checkpoint_paths") == 5
    # a checkpoint older than keep_every_n_hours relative to the last preserved one survives the rotation
    d2 = str(tmp_path / "ckpt2")
    checkpoint.save(eng, d2, 1, 0, "npz", max_to_keep=1, keep_every_n_hours=1e9)      # the clock starts here
    import time
    time.sleep(0.05)
    checkpoint.save(eng, d2, 2, 0, "npz", max_to_keep=1, keep_every_n_hours=1e9)      # rotates 1 out: deleted
    checkpoint.save(eng, d2, 3, 0, "npz", max_to_keep=1, keep_every_n_hours=1e-9)     # rotates 2 out: old enough, preserved
    assert not os.path.exists(os.path.join(d2, "checkpoint-1.npz")) and os.path.exists(os.path.join(d2, "checkpoint-2.npz"))
    step, _ = checkpoint.restore(eng, checkpoint.latest(d2))
    assert step == 3
    eng.close()


def test_missing_pipeline_or_unknown_transform_is_an_error(emul_lib, tmp_path):
    """the reference opens the pipeline YAML unconditionally and getattr()s every transform by name
    (model.py:341-356); silently running without preprocessing is not an option."""
    cfg = _config(tmp_path)
    cfg["TrainingSetting"]["Synthetic"] = False
    m = image2label(None, cfg, library=emul_lib)
    m.read_config()
    with pytest.raises(FileNotFoundError):
        m._transforms(str(tmp_path / "nope.yaml"), "train")
    y = tmp_path / "p.yaml"
    y.write_text("preprocess:\n  train:\n    3D:\n      - name: NoSuchTransform\n        variables: {}\n")
    with pytest.raises(AttributeError):
        m._transforms(str(y), "train")
    y.write_text("preprocess:\n  train:\n    3D:\n      - name: ManualNormalization\n        variables: {windowMin: 0, windowMax: 100}\n")
    assert len(m._transforms(str(y), "train")) == 1
