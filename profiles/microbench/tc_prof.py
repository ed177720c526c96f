"""C2 forward on the tensor-core kernel, 1M samples x 20 launches (ncu target for flow_tc_kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import torch, cnf_b200, bench
dev = torch.device('cuda:0')
n = 1_000_000
xs = [bench.synth(n, 100 + i, dev)[0] for i in range(4)]
m = bench.make_weights(seed=1).to(dev)
m.flow.precision = 'bf16'
with torch.no_grad():
    for i in range(3): m(xs[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20): m(xs[i % 4])
    e1.record(); torch.cuda.synchronize()
print('fwd ms per 1M', e0.elapsed_time(e1) / 20, 'G samples/s', n / (e0.elapsed_time(e1) / 20) / 1e6)
