"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv > profiles/r01_launches_fp32.txt
    python profiles/summarize.py full gpurun_out/prof_r01_fp32.ncu-rep > profiles/r01_ncu_full_fp32.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0]
        tot[name] += float(r[vi].replace(',', ''))
        cnt[name] += 1
    s = sum(tot.values())
    print('# per-kernel totals of gpu__time_duration.sum (%s); ncu launch list, cold-cache & serialised: compare SHARES' % rows[start][ui])
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print('%-100s n=%4d total=%14.1f mean=%12.1f share=%.4f' % (k[:100], cnt[k], v, v / cnt[k], v / s))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('kernel:', r[hdr.index('Kernel Name')][:120])
        for k in KEYS:
            if k in hdr:
                print('  %-85s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
