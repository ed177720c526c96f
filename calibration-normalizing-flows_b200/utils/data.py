"""On-disk logit formats of the reference, loaded straight into PINNED host
memory so that the flow kernels can stream them over PCIe without a staging copy
(``CouplingStack.transform_host`` runs zero-copy on pinned buffers).

Same file names and return shapes as the reference:
  * ``load_toy_dataset``  utils/data.py:170-176  ``<dataset>_separable_{logits,target}.npy``
  * ``load_logits``       utils/data.py:184-210  ``<model>_<dataset>/<dataset>_<model>_logit_prediction_{train,valid,test}.npy``
                                                 and ``..._true_{train,valid,test}.npy``
  * ``load_pickled_logits``  the ``train_logits.pkl`` / ``test_logits.pkl`` files written by
                             scripts/compute_logits.py:70-74
The rest of the reference's utils/data.py (argparse config, CIFAR image loaders) is out of scope.
"""
import os
import pickle

import numpy as np
import torch


def _pinned(arr, dtype):
    """numpy array (possibly memory-mapped) -> pinned torch tensor of `dtype`, one copy."""
    arr = np.asarray(arr)
    out = torch.empty(arr.shape, dtype=dtype, pin_memory=torch.cuda.is_available())
    out.copy_(torch.from_numpy(np.array(arr, copy=True)))      # (a memory-mapped array is read-only: copy, then wrap)
    return out


def load_toy_dataset(data_path, dataset):
    """Returns (logits, target) numpy arrays, exactly as the reference does."""
    stem = os.path.join(data_path, dataset + '_separable')
    return np.load(stem + '_logits.npy'), np.load(stem + '_target.npy')


def load_logits(dataset, model, data_path='../data', pin=True):
    """Returns ((logits, labels) for train, validation, test) as float32 / int64 torch tensors
    (pinned host memory when `pin`), the reference's layout and dtypes."""
    folder = os.path.join(data_path, '_'.join([model, dataset]))
    stem = os.path.join(folder, '_'.join([dataset, model]))
    out = []
    for split in ('train', 'valid', 'test'):
        logits = np.load(stem + '_logit_prediction_%s.npy' % split, mmap_mode='r')
        labels = np.load(stem + '_true_%s.npy' % split, mmap_mode='r')
        if pin:
            out.append((_pinned(logits, torch.float32), _pinned(labels, torch.int64)))
        else:
            out.append((torch.as_tensor(np.array(logits), dtype=torch.float32),
                        torch.as_tensor(np.array(labels), dtype=torch.int64)))
    return tuple(out)


def load_pickled_logits(model_dir, pin=True):
    """(train_logits, test_logits) from scripts/compute_logits.py's pickles."""
    out = []
    for name in ('train_logits.pkl', 'test_logits.pkl'):
        with open(os.path.join(model_dir, name), 'rb') as f:
            arr = pickle.load(f)
        out.append(_pinned(arr, torch.float32) if pin else torch.as_tensor(np.asarray(arr), dtype=torch.float32))
    return tuple(out)
