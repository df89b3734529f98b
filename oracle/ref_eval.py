"""CPU restatement of the reference's sliding-window evaluation loop (TEST INFRASTRUCTURE ONLY -- never imported
by the product path).  Follows model.py:866-937 (`evaluate_single_3D`): window grid `inum = ceil((dim - patch) /
stride) + 1` per axis with the last window clamped to `dim - patch` (model.py:866-903), batches of
`EvaluationSetting.BatchSize` consecutive windows (model.py:895-903), per batch one network run, then
`softmax_np[c][window] += softmax[j, ..., c]; weight_np[window] += 1` in window order (model.py:919-929) and
`label = argmax` over the accumulated, un-normalised sums (model.py:934).

Reference quirk kept on purpose: every batch's index list is put on the work list when its first window is created
(model.py:898-899) and "for last batch" the current list is appended once more after the loops (model.py:903-904) --
the same list object that is already there.  The last batch of a case is therefore run and accumulated TWICE (its
windows count 2 in `weight_np`, its softmax enters the sums twice), which changes the argmax where that batch overlaps
earlier windows.  `replay_last_batch=False` gives the loop without the duplicate.

Parity unpinned: the reference has no test for this loop; the restatement is checked by its own
known-answer tests in tests/test_host_mirror.py."""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np


def window_starts(vol: Sequence[int], patch: Sequence[int], stride: Sequence[int]) -> List[Tuple[int, int, int]]:
    """model.py:866-892: (i, j, k)-ordered window origins, last window of each axis clamped to the border."""
    num = [int(math.ceil((vol[a] - patch[a]) / float(stride[a]))) + 1 for a in range(3)]
    out = []
    for i in range(num[0]):
        for j in range(num[1]):
            for k in range(num[2]):
                st = []
                for a, idx in enumerate((i, j, k)):
                    s = idx * stride[a]
                    if s + patch[a] > vol[a]:
                        s = vol[a] - patch[a]
                    st.append(s)
                out.append(tuple(st))
    return out


def evaluate_volume(volume: np.ndarray, patch: Sequence[int], stride: Sequence[int], batch: int, num_classes: int,
                    softmax_fn: Callable[[np.ndarray], np.ndarray], replay_last_batch: bool = True):
    """model.py:895-937 with `softmax_fn(batch [B,X,Y,Z,M]) -> softmax [B,X,Y,Z,K]` standing in for
    sess.run('softmax:0').  Returns (label int64, softmax sums float32, weight float32)."""
    vol = volume.shape[:3]
    windows = window_starts(vol, patch, stride)
    sums = np.zeros(tuple(vol) + (num_classes,), np.float32)
    weight = np.zeros(vol, np.float32)
    starts = list(range(0, len(windows), batch))
    if replay_last_batch and starts:
        starts.append(starts[-1])          # model.py:903-904: the last batch is on the work list twice
    for b0 in starts:
        group = windows[b0:b0 + batch]
        x = np.stack([volume[s[0]:s[0] + patch[0], s[1]:s[1] + patch[1], s[2]:s[2] + patch[2], :] for s in group], 0)
        sm = softmax_fn(x)
        for j, s in enumerate(group):
            sl = (slice(s[0], s[0] + patch[0]), slice(s[1], s[1] + patch[1]), slice(s[2], s[2] + patch[2]))
            sums[sl] += sm[j]
            weight[sl] += 1.0
    return np.argmax(sums, axis=-1), sums, weight
