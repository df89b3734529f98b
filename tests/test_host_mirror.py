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
    assert w.max() == 4 and w.min() == 1  # overlap counts of 2x2x1 windows
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
