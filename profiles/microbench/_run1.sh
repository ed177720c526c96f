set -x
python profiles/microbench/tc_fwd_speed.py 2>&1 | tail -1
CNF_B200_LIB=$PWD/profiles/microbench/libcnf_slots2.so python profiles/microbench/tc_fwd_speed.py 2>&1 | tail -1
python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench, cnf_b200
dev = torch.device('cuda:0')
m = bench.make_model().to(dev)
N = 10_000_000
xh = torch.empty((N, 10), dtype=torch.float32, pin_memory=True); xh.copy_(bench.synth_dev(N, 3, dev)[0])
zh = torch.empty((N, 10), dtype=torch.float32, pin_memory=True); lh = torch.empty(N, dtype=torch.float32, pin_memory=True)
for prec in ('fp32', 'bf16'):
    m.flow.precision = prec
    for zc in ('1', ''):
        os.environ['CNF_LIVE_ENV'] = '1'
        if zc: os.environ['CNF_NO_ZEROCOPY'] = '1'
        else: os.environ.pop('CNF_NO_ZEROCOPY', None)
        for chunk in ((1 << 17, 1 << 19) if zc else (1 << 17,)):
            for _ in range(2): m.transform_host(xh, zh, lh, device=dev, chunk=chunk)
            t0 = time.perf_counter()
            for _ in range(5): m.transform_host(xh, zh, lh, device=dev, chunk=chunk)
            dt = (time.perf_counter() - t0) / 5
            print('e2e %s %s chunk %d: %.2f ms  %.3f G samples/s' % (prec, 'copy-engine pipeline' if zc else 'zero-copy', chunk, dt * 1e3, N / dt / 1e9), flush=True)
PY
python bench.py --steps 2 --warmup 1 > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo launches rc=$?
python profiles/microbench/head_prof.py > gpurun_out/plain_head.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'flow_tc_kernel|flow_reg10_kernel' -c 8 -o gpurun_out/prof_r02_head python profiles/microbench/head_prof.py > gpurun_out/ncu_head.log 2>&1
echo full rc=$?
cat gpurun_out/plain_head.log
