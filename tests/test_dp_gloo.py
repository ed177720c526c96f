"""Data-parallel host logic on CPU: world_size-2 gloo run of FusedNLLTrainer with the numpy oracle
standing in for the CUDA engine (test-only injection; the product engine is CUDA-only).
Checks: shards are contiguous and cover every sample once; the all-reduced gradient equals the
single-process full-batch gradient; every rank applies the same update (bit-identical params);
the reduced loss statistics equal the full-batch ones."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flow_oracle as orc


class OracleEngine:
    """Duck-typed stand-in for cnf_b200._engine.StackEngine on CPU tensors."""

    def __init__(self, params, K=5):
        self.like = params
        self.K = K
        self.tables = torch.arange(4)          # the trainer compares the index tables across ranks
        self.adam_m = self.adam_v = None       # optimiser state the trainer would broadcast from rank 0
        self.adam_t = 0
        self.flat = torch.from_numpy(orc.flatten(params).astype(np.float64))
        self.n_flat = self.flat.numel()
        self.flat_grad = torch.zeros_like(self.flat)
        self.m = np.zeros(self.flat.numel())
        self.v = np.zeros(self.flat.numel())
        self.t = 0

    def ensure(self, device):
        pass

    def pack(self, tc=False, fp32=True):
        pass

    def nll_step(self, x, y, loss_acc, eps=1e-7, gamma=1.0, n_total=None, with_grad=True, precision="fp32"):
        p = orc.unflatten(self.flat.numpy(), self.like)
        xn, yn = x.numpy().astype(np.float64), y.numpy()
        zs, ld = orc.flow_forward(p, xn)
        loss, ce, ldm, gz, gld = orc.nll_head(zs[-1], ld, yn, eps, gamma, n_total)
        n = n_total
        loss_acc += torch.tensor([-(loss) * n, -(ce) * n, ldm * n, 0.0], dtype=torch.float64)
        if with_grad:
            grads, _ = orc.flow_backward(p, xn, gz, gld)
            self.flat_grad.copy_(torch.from_numpy(orc.flatten(grads)))

    def adam(self, lr, betas, eps, wd):
        self.t += 1
        pnew, self.m, self.v = orc.adam_step(self.flat.numpy(), self.flat_grad.numpy(), self.m, self.v, self.t,
                                             lr, betas[0], betas[1], eps, wd)
        self.flat.copy_(torch.from_numpy(pnew))

    def sgd(self, lr, wd):
        self.flat.copy_(torch.from_numpy(orc.sgd_step(self.flat.numpy(), self.flat_grad.numpy(), lr, wd)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cnf_b200.calibrators import FusedNLLTrainer, shard_bounds
    K, L, H, N = 5, 3, [8], 101
    params = orc.init_params(K, L, H, rng=np.random.default_rng(1), wscale=0.3, dtype=np.float64)
    x, y = orc.synth_logits(N, K, seed=2)
    lo, hi = shard_bounds(N, rank, world)
    eng = OracleEngine(params)
    tr = FusedNLLTrainer(eng, torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]))
    assert tr.n_total == N
    losses = []
    for _ in range(3):
        tr.step()
        losses.append(-float(tr.loss_acc[0]) / N)
    ev = tr.evaluate()
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), flat=eng.flat.numpy(), losses=np.array(losses),
             ev=ev.numpy(), lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), 'rank%d.npz' % i))) for i in range(world)]
    # shards: contiguous, disjoint, complete
    assert int(r[0]['lo']) == 0 and int(r[0]['hi']) == int(r[1]['lo']) and int(r[1]['hi']) == 101
    # bit-identical parameters on both ranks (same reduced gradient, same update)
    assert np.array_equal(r[0]['flat'], r[1]['flat'])
    assert np.array_equal(r[0]['losses'], r[1]['losses'])
    # single-process reference
    K, L, H, N = 5, 3, [8], 101
    params = orc.init_params(K, L, H, rng=np.random.default_rng(1), wscale=0.3, dtype=np.float64)
    x, y = orc.synth_logits(N, K, seed=2)
    flat = orc.flatten(params)
    m = np.zeros_like(flat)
    v = np.zeros_like(flat)
    losses = []
    for t in range(1, 4):
        p = orc.unflatten(flat, params)
        loss, _, _, grads, _ = orc.train_step_grads(p, x.astype(np.float64), y)
        losses.append(loss)
        flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
    assert np.allclose(r[0]['losses'], losses, rtol=1e-10, atol=1e-12)
    assert np.allclose(r[0]['flat'], flat, rtol=1e-9, atol=1e-12)
    p = orc.unflatten(flat, params)
    zs, ld = orc.flow_forward(p, x.astype(np.float64))
    loss, ce, ldm, _, _ = orc.nll_head(zs[-1], ld, y)
    assert np.allclose(r[0]['ev'][:3], [-loss * N, -ce * N, ldm * N], rtol=1e-9)


def test_shard_bounds_cover_everything():
    from cnf_b200.calibrators import shard_bounds
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------------
# fit() schedule under data parallelism with N % world != 0 (ranks must not disagree on
# full-batch / number of steps, and the gradient of an unequal tail batch must be scaled by its true size)
# ---------------------------------------------------------------------------------------------------
def test_plan_fit_is_rank_invariant_and_complete():
    from cnf_b200.calibrators import plan_fit, shard_bounds
    for n in (1, 2, 7, 101, 1000, 4097):
        for world in (1, 2, 3, 8):
            for bs in (1, 3, 16, 40, n - 1, n, n + 5):
                if bs < 1:
                    continue
                full, local_bs, steps, totals = plan_fit(n, world, bs)
                assert full == (bs >= n)
                assert len(totals) == steps and sum(totals) == n
                if not full:
                    # every rank's shard is consumed in exactly `steps` slices of local_bs rows
                    for r in range(world):
                        lo, hi = shard_bounds(n, r, world)
                        assert -(-(hi - lo) // local_bs) <= steps
                    if world == 1:
                        assert steps == -(-n // bs)      # the reference DataLoader's batch count


def _perm_for(rank, call, n):
    return torch.from_numpy(np.random.default_rng(1000 * rank + call).permutation(n))


def _fit_worker(rank, world, port, out_dir, batch_size):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cnf_b200.calibrators import FusedNLLTrainer, shard_bounds
    K, L, H, N = 5, 2, [6], 101
    # every rank starts from DIFFERENT parameters: the trainer must take rank 0's
    params = orc.init_params(K, L, H, rng=np.random.default_rng(7 + rank), wscale=0.3, dtype=np.float64)
    x, y = orc.synth_logits(N, K, seed=3)
    lo, hi = shard_bounds(N, rank, world)
    eng = OracleEngine(params, K)
    tr = FusedNLLTrainer(eng, torch.from_numpy(x[lo:hi]), torch.from_numpy(y[lo:hi]), n_total=N)
    calls = [0]

    def perm_fn(n):
        calls[0] += 1
        return _perm_for(rank, calls[0], n)
    hist = tr.fit_loop(2, batch_size, perm_fn)
    np.savez(os.path.join(out_dir, 'fit%d.npz' % rank), flat=eng.flat.numpy(), hist=hist.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize('batch_size', [101, 40, 50])
def test_two_rank_fit_uneven_shards(tmp_path, batch_size):
    from cnf_b200.calibrators import plan_fit, shard_bounds
    world, N, K, L, H = 2, 101, 5, 2, [6]
    port = _free_port()
    mp.spawn(_fit_worker, args=(world, port, str(tmp_path), batch_size), nprocs=world, join=True)
    r = [dict(np.load(os.path.join(str(tmp_path), 'fit%d.npz' % i))) for i in range(world)]
    assert np.array_equal(r[0]['flat'], r[1]['flat'])       # same start (rank 0's), same updates
    assert np.array_equal(r[0]['hist'], r[1]['hist'])
    # single-process emulation of the same schedule from rank 0's initial parameters
    params = orc.init_params(K, L, H, rng=np.random.default_rng(7), wscale=0.3, dtype=np.float64)
    x, y = orc.synth_logits(N, K, seed=3)
    x = x.astype(np.float64)
    flat = orc.flatten(params)
    m, v, t = np.zeros_like(flat), np.zeros_like(flat), 0
    full, local_bs, steps, totals = plan_fit(N, world, batch_size)
    shards = [shard_bounds(N, q, world) for q in range(world)]
    calls = [0, 0]
    hist = []
    for epoch in range(2):
        if full:
            loss, _, _, grads, _ = orc.train_step_grads(orc.unflatten(flat, params), x, y)
            t += 1
            flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
            continue
        perms = []
        for q in range(world):
            calls[q] += 1
            perms.append(_perm_for(q, calls[q], shards[q][1] - shards[q][0]).numpy())
        for s in range(steps):
            idx = np.concatenate([shards[q][0] + perms[q][s * local_bs:(s + 1) * local_bs] for q in range(world)])
            assert len(idx) == totals[s]
            _, _, _, grads, _ = orc.train_step_grads(orc.unflatten(flat, params), x[idx], y[idx])
            t += 1
            flat, m, v = orc.adam_step(flat, orc.flatten(grads), m, v, t)
        perms = []
        for q in range(world):
            calls[q] += 1
            perms.append(_perm_for(q, calls[q], shards[q][1] - shards[q][0]).numpy())
        idx = np.concatenate([shards[q][0] + perms[q][(steps - 1) * local_bs:steps * local_bs] for q in range(world)])
        p = orc.unflatten(flat, params)
        zs, ld = orc.flow_forward(p, x[idx])
        loss, ce, ldm, _, _ = orc.nll_head(zs[-1], ld, y[idx])
        hist.append([-loss * len(idx), -ce * len(idx), ldm * len(idx)])
    assert np.allclose(r[0]['flat'], flat, rtol=1e-9, atol=1e-12)
    if not full:
        assert np.allclose(r[0]['hist'][:, :3], np.array(hist), rtol=1e-9)
