"""SURVEY.md 8f rank 4 on the GPU: the reference's on-disk logit layout (utils/data.py:184-210) -> pinned host tensors
(cnf_b200.utils.data) -> training that streams the samples from host memory (HostStreamNLLTrainer): every step for
sets larger than HBM, only the first step (load hidden behind the first epoch) for sets that fit.  Both must follow
the resident trainer: same losses (1e-6 relative), same first gradient (1e-5 of its maximum)."""
import os

import numpy as np
import pytest

import flow_oracle as orc
from conftest import load_golden
from helpers import build_flow_from_golden, rel_err

pytestmark = pytest.mark.gpu


def _write_reference_layout(root, x, y):
    folder = os.path.join(root, 'resnet_cifar10')
    os.makedirs(folder)
    stem = os.path.join(folder, 'cifar10_resnet')
    cut = [0, len(x) - 2000, len(x) - 1000, len(x)]
    for i, split in enumerate(('train', 'valid', 'test')):
        np.save(stem + '_logit_prediction_%s.npy' % split, x[cut[i]:cut[i + 1]])
        np.save(stem + '_true_%s.npy' % split, y[cut[i]:cut[i + 1]])


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
@pytest.mark.parametrize('resident', [False, True])
def test_streamed_training_from_disk_layout_matches_resident(precision, resident, tmp_path, cuda_device):
    import torch
    import cnf_b200
    from cnf_b200.utils import data as D
    N, K = 302_000, 10
    x, y = orc.synth_logits(N, K, seed=12)
    _write_reference_layout(str(tmp_path), x, y)
    (xtr, ytr), (xva, yva), (xte, yte) = D.load_logits('cifar10', 'resnet', data_path=str(tmp_path))
    assert xtr.is_pinned() and ytr.is_pinned() and xtr.dtype == torch.float32 and ytr.dtype == torch.int64
    assert xtr.shape == (N - 2000, K) and xva.shape == (1000, K) and xte.shape == (1000, K)
    assert np.array_equal(xtr.numpy(), x[:N - 2000])
    g = load_golden('flow_c2_nvp_k10_init')
    f_res = build_flow_from_golden(g, cuda_device)
    f_str = build_flow_from_golden(g, cuda_device)
    tr_res = cnf_b200.FusedNLLTrainer(f_res.engine(), xtr.to(cuda_device), ytr.to(cuda_device), precision=precision)
    tr_str = cnf_b200.HostStreamNLLTrainer(f_str.engine(), xtr, ytr, cuda_device, chunk_rows=65536, resident=resident,
                                           precision=precision)
    assert tr_str.n_total == N - 2000
    losses = []
    for i in range(3):
        tr_res.step()
        tr_str.step()
        if i == 0:
            tol = 1e-5 if precision == 'fp32' else 2e-3
            assert rel_err(f_str.engine().flat_grad.cpu().numpy(), f_res.engine().flat_grad.cpu().numpy()) < tol
        losses.append((-float(tr_res.loss_acc[0]) / tr_res.n_total, -float(tr_str.loss_acc[0]) / tr_str.n_total))
    losses = np.array(losses)
    assert np.allclose(losses[:, 0], losses[:, 1], rtol=1e-6 if precision == 'fp32' else 1e-4)
    ev_res, ev_str = tr_res.evaluate().cpu().numpy(), tr_str.evaluate().cpu().numpy()
    assert np.allclose(ev_res[:3], ev_str[:3], rtol=1e-5 if precision == 'fp32' else 1e-3)
    assert tr_str._loaded == resident
    # held-out split through the fused evaluate of the calibrator API
    with pytest.raises(NotImplementedError):
        tr_str.fit_loop(1, 1000, lambda n: torch.arange(n))


def test_calibrator_fit_host_stream_switch(cuda_device):
    """TorchFlowCalibrator(host_stream=True): same history as the resident fit."""
    import cnf_b200
    g = load_golden('calibrator_cal_nvp_k10')
    hist = {}
    for hs in (False, True):
        import torch
        torch.manual_seed(3)
        cal = cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, g['x'], g['y'], layers=3, hidden_size=[16], epochs=5,
                                           dev=cuda_device, host_stream=hs)
        assert isinstance(cal.trainer, cnf_b200.HostStreamNLLTrainer) == hs
        hist[hs] = np.array([float(v) for v in cal.history['loss']])
    assert np.allclose(hist[False], hist[True], rtol=1e-5)
