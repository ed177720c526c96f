import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle')); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch, cnf_b200
import flow_oracle as orc
from conftest import load_golden
from helpers import build_flow_from_golden
dev = torch.device('cuda:0')
for name in ('c2_nvp_k10', 'nvp_k7_randflip', 'scaleonly_k4_h888', 'nvp_k10_nohidden'):
    g = load_golden('flow_' + name)
    flow = build_flow_from_golden(g, dev)
    eng = flow.engine()
    x = torch.from_numpy(orc.synth_logits(777, int(g['K']), seed=1)[0]).to(dev)
    y = torch.from_numpy(orc.synth_logits(777, int(g['K']), seed=1)[1]).to(dev)
    z, ld, allz = eng.apply(x, want_all=True)
    eng.apply(z, inverse=True)
    acc = torch.zeros(4, dtype=torch.float64, device=dev)
    eng.nll_step(x, y, acc); eng.adam(); eng.pack()
    if eng.tc_bytes: eng.apply(x, precision='bf16'); eng.apply(z, inverse=True, precision='bf16')
    p = torch.softmax(z, 1)
    cnf_b200.expected_calibration_error(p, y); cnf_b200.neg_log_likelihood(p.double(), y)
m = cnf_b200.RealNvpFlow(100, layers=2, hidden_size=[512]).to(dev)
x = torch.randn(300, 100, device=dev)
m.flow.precision = 'bf16'
with torch.no_grad(): m(x)
torch.cuda.synchronize(); print('sanitizer workload done')
