"""bf16 tensor-core training step at the C2 shape: samples per forward / backward kernel pair (StackEngine.TRAIN_CHUNK).
A 2^20-sample chunk writes and reads a 400 MB tape through HBM; smaller chunks keep it in the 126 MB L2 but pay more
launches and more persistent-kernel prologues."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
import cnf_b200  # noqa: E402,F401

dev = torch.device('cuda:0')
N = int(os.environ.get('N', 16 << 20))
x, y = bench.synth_dev(N, 7, dev)
for chunk in (1 << 20, 1 << 19, 1 << 18, 1 << 17, 1 << 21, 1 << 22):
    model = bench.make_model(seed=2, wmult=1.0).to(dev)
    eng = model.engine()
    eng.TRAIN_CHUNK = chunk
    tr = cnf_b200.FusedNLLTrainer(eng, x, y, precision='bf16')
    for _ in range(2):
        tr.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        tr.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('chunk 2^%d: %.2f ms per step  %.1f M samples/s  loss %.6f' % (chunk.bit_length() - 1, ms, N / ms / 1e3, -float(tr.loss_acc[0]) / N), flush=True)
    del tr, model, eng
