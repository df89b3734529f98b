"""TensorBoard event files for the scalars of the training loop, written without TensorFlow.

The reference logs through `tf.summary.FileWriter(LogDir + '/train' | '/test')` (model.py:705-709, written at
:750-756,790-794).  Its image and confusion-metric summaries are out of scope; the scalars the
hot path itself produces - `loss/0.total_loss` (model.py:562) and `learning_rate` (model.py:644) - are kept so that
`tensorboard --logdir LogDir` keeps working on runs of this engine.

File format (tensorflow/core/lib/io/record_writer.cc, tensorflow/core/util/event.proto): a sequence of records
`length:u64 | masked_crc32c(length):u32 | data | masked_crc32c(data):u32`, each `data` an Event proto
{wall_time = 1 (double), step = 2 (int64), file_version = 3 (string) | summary = 5 {repeated value = 1 {tag = 1,
simple_value = 2 (float)}}}; the first record carries file_version "brain.Event:2".
"""
from __future__ import annotations

import os
import socket
import struct
import time
from typing import Dict

from .tf_bundle import _varint, crc32c, crc_mask


def _record(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", crc_mask(crc32c(head))) + data + struct.pack("<I", crc_mask(crc32c(data)))


def _event(wall_time: float, step: int, payload: bytes) -> bytes:
    out = b"\x09" + struct.pack("<d", wall_time)
    if step:
        out += b"\x10" + _varint(step)
    return out + payload


def _scalar_summary(scalars: Dict[str, float]) -> bytes:
    body = b""
    for tag, value in scalars.items():
        t = tag.encode()
        v = b"\x0a" + _varint(len(t)) + t + b"\x15" + struct.pack("<f", float(value))
        body += b"\x0a" + _varint(len(v)) + v
    return b"\x2a" + _varint(len(body)) + body


class EventFileWriter:
    """`events.out.tfevents.<seconds>.<host>` under `logdir`, one flushed record per `add_scalars` call."""

    def __init__(self, logdir: str):
        os.makedirs(logdir, exist_ok=True)
        now = time.time()
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s.%d" % (int(now), socket.gethostname(), os.getpid()))
        self._f = open(self.path, "ab")
        version = b"brain.Event:2"
        self._f.write(_record(_event(now, 0, b"\x1a" + _varint(len(version)) + version)))
        self._f.flush()

    def add_scalars(self, step: int, scalars: Dict[str, float]):
        self._f.write(_record(_event(time.time(), int(step), _scalar_summary(scalars))))
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.close()
