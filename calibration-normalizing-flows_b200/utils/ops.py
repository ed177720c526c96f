"""``onehot_encode`` -- drop-in for reference ``utils/ops.py:42-51`` (host side, used by
``Calibrator.__init__``).  The number of classes is ``max(label) + 1`` as in the reference."""
import numpy as np


def onehot_encode(target):
    labels = np.asarray(target).astype(np.int32)
    width = int(labels.max()) + 1
    return (labels[:, None] == np.arange(width, dtype=np.int32)[None, :]).astype(np.int32)
