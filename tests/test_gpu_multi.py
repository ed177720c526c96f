"""Multi-GPU data-parallel training on real devices (skipped with fewer than 2 GPUs): the NCCL
all-reduced gradient and the resulting parameters equal the single-GPU full-batch run."""
import os
import socket

import numpy as np
import pytest

import flow_oracle as orc

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir, precision='fp32'):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from cnf_b200.calibrators import FusedNLLTrainer, shard_bounds
    from conftest import load_golden
    from helpers import build_flow_from_golden
    g = load_golden('flow_c2_nvp_k10_init')
    N = 200_003
    x, y = orc.synth_logits(N, 10, seed=21)
    lo, hi = shard_bounds(N, rank, world)
    flow = build_flow_from_golden(g, dev)
    tr = FusedNLLTrainer(flow.engine(), torch.from_numpy(x[lo:hi]).to(dev), torch.from_numpy(y[lo:hi]).to(dev),
                         precision=precision)
    assert tr.n_total == N
    losses = []
    for _ in range(4):
        tr.step()
        losses.append(-float(tr.loss_acc[0]) / N)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), flat=flow.engine().flat.cpu().numpy(),
             losses=np.array(losses), grad=flow.engine().flat_grad.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_two_gpu_dp_matches_single_gpu(precision, tmp_path, cuda_device):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from cnf_b200.calibrators import FusedNLLTrainer
    from conftest import load_golden
    from helpers import build_flow_from_golden, rel_err
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), precision), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), 'rank%d.npz' % i))) for i in range(world)]
    assert np.array_equal(r[0]['flat'], r[1]['flat'])           # identical update on every rank
    g = load_golden('flow_c2_nvp_k10_init')
    N = 200_003
    x, y = orc.synth_logits(N, 10, seed=21)
    flow = build_flow_from_golden(g, cuda_device)
    tr = FusedNLLTrainer(flow.engine(), torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device),
                         precision=precision)
    losses = []
    for _ in range(4):
        tr.step()
        losses.append(-float(tr.loss_acc[0]) / N)
    assert np.allclose(losses, r[0]['losses'], rtol=1e-6, atol=1e-7)
    # bf16: the two runs cut the samples into different tiles and accumulate in a different order
    assert rel_err(r[0]['grad'], flow.engine().flat_grad.cpu().numpy()) < (1e-4 if precision == 'fp32' else 1e-3)
    disp = flow.engine().flat.cpu().numpy() - g['flat']
    disp_dp = r[0]['flat'] - g['flat']
    # Adam normalises every entry to ~lr per step, so an entry whose gradient is rounding noise can land +-lr either
    # way, and the two runs sum the samples in different orders (different shards, and at this size the single-rank
    # run takes the register-resident kernel while the half-size shards take the tile kernel): the entries whose
    # gradient is well above the noise (> 1e-3 of the largest) -- where the update is a smooth function of the
    # gradient -- must agree closely; of all entries at most 1 % may differ by more than 5 % of the largest
    # displacement
    scale = np.max(np.abs(disp))
    diff = np.abs(disp_dp - disp)
    gref = flow.engine().flat_grad.cpu().numpy()
    solid = np.abs(gref) > 1e-3 * np.max(np.abs(gref))
    tol = 5e-3 if precision == 'fp32' else 5e-2
    assert solid.any() and np.max(diff[solid]) < tol * scale, np.max(diff[solid]) / scale
    assert np.mean(diff > 0.05 * scale) < 0.01, np.mean(diff > 0.05 * scale)
    assert np.all(disp_dp[gref == 0] == 0)
    # ... and the entries whose gradient is well above the noise agree tightly
    gref = flow.engine().flat_grad.cpu().numpy()
    solid = np.abs(gref) > 1e-3 * np.max(np.abs(gref))
    err = np.max(np.abs((r[0]['flat'] - g['flat']) - disp)[solid]) / np.max(np.abs(disp))
    assert err < (5e-3 if precision == 'fp32' else 3e-2), err


def _cal_worker(rank, world, port, out_dir, batch_size, graph):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import cnf_b200
    from cnf_b200.calibrators import shard_bounds
    N, K = 1001, 10                       # N % world != 0: the shards differ by one row
    x, y = orc.synth_logits(N, K, seed=5)
    x = x + 0.25
    torch.manual_seed(1000 + rank)        # every rank builds DIFFERENT initial weights: fit must take rank 0's
    cal = cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, x, y, layers=3, hidden_size=[16], epochs=4,
                                       batch_size=batch_size, dev=dev, cuda_graph=graph)
    hist = np.array([[float(v) for v in cal.history[k]] for k in ('loss', 'ce', 'log_det')])
    flat = np.concatenate([p.detach().cpu().numpy().reshape(-1) for lay in cal.flow.layers for p in lay.canonical_parameters()])
    # sharded evaluation: every rank passes its own rows, the 48 statistics are all-reduced
    lo, hi = shard_bounds(N, rank, world)
    m = cal.evaluate(x[lo:hi], y[lo:hi], bins=15)
    m_local = cal.evaluate(x[lo:hi], y[lo:hi], bins=15, reduce=False)
    pred = cal.predict(x[:50])
    np.savez(os.path.join(out_dir, 'cal%d.npz' % rank), hist=hist, flat=flat, pred=pred,
             m=np.array([m['ece'], m['nll'], m['accuracy'], m['n']]),
             m_local=np.array([m_local['ece'], m_local['nll'], m_local['accuracy'], m_local['n']]))
    dist.destroy_process_group()


@pytest.mark.parametrize('batch_size,graph', [(1001, False), (400, False)])
def test_two_gpu_calibrator_fit_uneven_shards_and_sharded_metrics(batch_size, graph, tmp_path, cuda_device):
    """TorchFlowCalibrator.fit under NCCL with N % world != 0: same collective sequence on every rank,
    rank 0's start state everywhere, full-batch history equal to the single-GPU fit; evaluate() all-reduces the
    3*bins+3 statistics so every rank reports the metrics of the whole set."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import cnf_b200
    world = 2
    mp.spawn(_cal_worker, args=(world, _free_port(), str(tmp_path), batch_size, graph), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), 'cal%d.npz' % i))) for i in range(world)]
    assert np.array_equal(r[0]['flat'], r[1]['flat']) and np.array_equal(r[0]['hist'], r[1]['hist'])
    assert np.array_equal(r[0]['pred'], r[1]['pred'])
    assert np.array_equal(r[0]['m'], r[1]['m']) and int(r[0]['m'][3]) == 1001
    assert int(r[0]['m_local'][3]) + int(r[1]['m_local'][3]) == 1001
    assert np.isfinite(r[0]['hist']).all()
    # single-GPU run from rank 0's seed
    N, K = 1001, 10
    x, y = orc.synth_logits(N, K, seed=5)
    x = x + 0.25
    torch.manual_seed(1000)
    cal = cnf_b200.TorchFlowCalibrator(cnf_b200.RealNvpFlow, x, y, layers=3, hidden_size=[16], epochs=4,
                                       batch_size=batch_size, dev=cuda_device)
    if batch_size >= N:
        hist = np.array([[float(v) for v in cal.history[k]] for k in ('loss', 'ce', 'log_det')])
        assert np.allclose(hist, r[0]['hist'], rtol=1e-5, atol=1e-7)
        assert np.max(np.abs(cal.predict(x[:50]) - r[0]['pred'])) < 1e-5
        m1 = cal.evaluate(x, y, bins=15)
        assert abs(m1['ece'] - r[0]['m'][0]) < 1e-5 and abs(m1['nll'] - r[0]['m'][1]) < 1e-5
