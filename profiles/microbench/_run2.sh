python profiles/microbench/tc_fwd_speed.py 2>&1 | tail -1
CNF_B200_LIB=$PWD/profiles/microbench/libcnf_slots2.so python profiles/microbench/tc_fwd_speed.py 2>&1 | tail -1
N=20000000 python profiles/microbench/head_prof.py > gpurun_out/plain_head.log 2>&1 && N=20000000 ncu --set full --clock-control none --import-source on -k regex:'flow_tc_kernel|flow_reg10_kernel' -c 8 -o gpurun_out/prof_r02_head python profiles/microbench/head_prof.py > gpurun_out/ncu_head.log 2>&1
echo full rc=$?
cat gpurun_out/plain_head.log
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "host_buffer" 2>&1 | tail -2
