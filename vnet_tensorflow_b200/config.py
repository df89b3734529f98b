"""Tolerant reader of the reference's config.json (model.py:185-245).

None of the JSON files shipped with the reference loads cleanly with the reference's own code
(`NumCovolutions` vs `NumConvolutions`, missing `MaxIterations` / `TestStep` / `Pipeline`,
`LargestConnectedComponent`, `VolumeThreshold`, ...).  This loader accepts both spellings and applies
the documented defaults (networks.py:213-216), and adds optional keys for the B200 engine:
`TrainingSetting.Precision` ("fp32" | "bf16x3" | "bf16"), `TrainingSetting.Synthetic` and
`TrainingSetting.DataWorkers` (patch-pipeline threads; the reference maps with num_parallel_calls=1).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional, Sequence


def _get(d: dict, *names, default=None, required=False):
    for n in names:
        if n in d:
            return d[n]
    if required:
        raise KeyError(names[0])
    return default


@dataclass
class Config:
    # training (model.py:189-224)
    input_channel_num: int = 1
    output_channel_num: int = 2
    label_classes: List[int] = field(default_factory=lambda: [0, 1])
    train_data_dir: str = "./data/training"
    test_data_dir: str = "./data/testing"
    image_filenames: List[str] = field(default_factory=lambda: ["image.nii"])
    label_filename: str = "label.nii"
    batch_size: int = 1
    patch_shape: Sequence[int] = (64, 64, 64)
    image_log: bool = False
    testing: bool = True
    test_step: int = 30
    restore_training: bool = True
    log_dir: str = "./tmp/log"
    ckpt_dir: str = "./tmp/ckpt"
    epoches: int = 99999
    max_itr: int = 10 ** 9
    log_interval: int = 50
    network_name: str = "VNet"
    dropout_rate: float = 0.01
    num_channel: int = 16
    num_levels: int = 4
    num_convolutions: Sequence[int] = (1, 2, 3, 3)
    bottom_convolutions: int = 3
    optimizer_name: str = "Adam"
    initial_learning_rate: float = 1e-2
    decay_factor: float = 0.99
    decay_steps: float = 100
    momentum: float = 0.9
    spacing: Sequence[float] = (1.0, 1.0, 1.0)
    drop_ratio: float = 0.01
    min_pixel: int = 30
    loss_name: str = "weighted_sorensen"
    loss_weights: Sequence[float] = (1.0, 1.0)
    loss_alpha: float = 1.0
    training_pipeline: Optional[str] = None
    precision: str = "bf16x3"
    synthetic: bool = False
    data_workers: int = 4
    checkpoint_format: str = "npz"
    # evaluation (model.py:227-241)
    checkpoint_path: str = "./tmp/ckpt/checkpoint-latest"
    evaluate_data_dir: str = "./data/evaluate"
    evaluate_image_filenames: List[str] = field(default_factory=lambda: ["image.nii"])
    evaluate_label_filename: str = "label_tf.nii.gz"
    evaluate_probability_filename: str = "probability_tf.nii.gz"
    evaluate_stride: Sequence[int] = (64, 64, 64)
    evaluate_batch: int = 1
    evaluate_probability_output: bool = True
    evaluate_lcc: bool = False
    evaluate_volume_threshold: float = 0.0
    evaluate_pipeline: Optional[str] = None
    evaluate_map_label_values: bool = False

    @property
    def dimension(self) -> int:
        return len(self.patch_shape)


def from_dict(cfg: dict) -> Config:
    t = cfg["TrainingSetting"]
    e = cfg.get("EvaluationSetting", {})
    data, net = t.get("Data", {}), t.get("Networks", {})
    opt, loss = t.get("Optimizer", {}), t.get("Loss", {})
    c = Config()
    c.image_filenames = list(_get(data, "ImageFilenames", default=c.image_filenames))
    c.input_channel_num = len(c.image_filenames)
    c.label_classes = list(_get(t, "SegmentationClasses", default=c.label_classes))
    c.output_channel_num = len(c.label_classes)
    c.train_data_dir = _get(data, "TrainingDataDirectory", default=c.train_data_dir)
    c.test_data_dir = _get(data, "TestingDataDirectory", default=c.test_data_dir)
    c.label_filename = _get(data, "LabelFilename", default=c.label_filename)
    c.batch_size = int(_get(t, "BatchSize", default=c.batch_size))
    c.patch_shape = tuple(int(p) for p in _get(t, "PatchShape", default=c.patch_shape))
    c.image_log = bool(_get(t, "ImageLog", default=False))
    c.testing = bool(_get(t, "Testing", default=True))
    c.test_step = int(_get(t, "TestStep", default=c.test_step))
    c.restore_training = bool(_get(t, "Restore", default=True))
    c.log_dir = _get(t, "LogDir", default=c.log_dir)
    c.ckpt_dir = _get(t, "CheckpointDir", default=c.ckpt_dir)
    c.epoches = int(_get(t, "Epoches", "Epochs", default=c.epoches))
    c.max_itr = int(_get(t, "MaxIterations", default=c.max_itr))
    c.log_interval = int(_get(t, "LogInterval", default=c.log_interval))
    c.network_name = _get(net, "Name", default="VNet")
    c.dropout_rate = float(_get(net, "Dropout", default=c.dropout_rate))
    c.num_channel = int(_get(net, "NumChannel", default=16))
    c.num_levels = int(_get(net, "NumLevels", default=4))
    c.num_convolutions = tuple(int(v) for v in _get(net, "NumConvolutions", "NumCovolutions", default=(1, 2, 3, 3)))
    c.bottom_convolutions = int(_get(net, "BottomConvolutions", default=3))
    c.optimizer_name = _get(opt, "Name", default="Adam")
    c.initial_learning_rate = float(_get(opt, "InitialLearningRate", default=1e-2))
    dec = opt.get("Decay", {})
    c.decay_factor = float(_get(dec, "Factor", default=0.99))
    c.decay_steps = float(_get(dec, "Steps", default=100))
    c.momentum = float(_get(opt, "Momentum", default=0.9))  # the reference reads self.momentum, never set (R6)
    c.spacing = tuple(_get(t, "Spacing", default=c.spacing))
    c.drop_ratio = float(_get(t, "DropRatio", default=0.01))
    c.min_pixel = int(_get(t, "MinPixel", default=30))
    c.loss_name = _get(loss, "Name", default="weighted_sorensen")
    c.loss_weights = tuple(float(w) for w in _get(loss, "Weights", default=[1.0] * c.output_channel_num))
    c.loss_alpha = float(_get(loss, "Alpha", default=1.0))
    c.training_pipeline = _get(t, "Pipeline", default=None)
    c.precision = _get(t, "Precision", default="bf16x3")
    c.synthetic = bool(_get(t, "Synthetic", default=False))
    c.data_workers = max(1, int(_get(t, "DataWorkers", default=4)))
    c.checkpoint_format = _get(t, "CheckpointFormat", default="npz")  # "npz" | "tf" | "both" (checkpoint.py)
    ed = e.get("Data", {})
    c.checkpoint_path = _get(e, "CheckpointPath", default=c.checkpoint_path)
    c.evaluate_data_dir = _get(ed, "EvaluateDataDirectory", default=c.evaluate_data_dir)
    c.evaluate_image_filenames = list(_get(ed, "ImageFilenames", default=c.image_filenames))
    c.evaluate_label_filename = _get(ed, "LabelFilename", default=c.evaluate_label_filename)
    c.evaluate_probability_filename = _get(ed, "ProbabilityFilename", default=c.evaluate_probability_filename)
    c.evaluate_stride = tuple(int(s) for s in _get(e, "Stride", default=c.patch_shape))
    c.evaluate_batch = int(_get(e, "BatchSize", default=1))
    c.evaluate_probability_output = bool(_get(e, "ProbabilityOutput", default=True))
    c.evaluate_lcc = bool(_get(e, "LargestConnectedComponent", default=False))
    c.evaluate_volume_threshold = float(_get(e, "VolumeThreshold", default=0))  # physical size, model.py:125
    c.evaluate_pipeline = _get(e, "Pipeline", default=None)
    c.evaluate_map_label_values = bool(_get(e, "MapLabelValues", default=False))  # extension: class index -> SegmentationClasses value
    if c.network_name != "VNet":
        raise SystemExit("Invalid Network")  # model.py:439-440 (UNet / Dense are outside the accelerated path)
    if len(c.num_convolutions) != c.num_levels:
        raise AssertionError("num_levels == len(num_convolutions)")
    return c


def load(path: str) -> Config:
    with open(path) as f:
        return from_dict(json.load(f))
