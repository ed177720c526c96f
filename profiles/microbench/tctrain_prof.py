"""One bf16 tensor-core training step at 2^21 samples (ncu target for flow_tcb_kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
n = 1 << 21
xt, yt = bench.synth_dev(n, 5000, dev)
m = bench.make_model(seed=2, wmult=1.0).to(dev)
tr = cnf_b200.FusedNLLTrainer(m.engine(), xt, yt, precision='bf16')
for _ in range(3): tr.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): tr.step()
e1.record(); torch.cuda.synchronize()
print('bf16 train ms/step @2M', e0.elapsed_time(e1)/5, 'samples/s', n / (e0.elapsed_time(e1) / 5 * 1e-3))
