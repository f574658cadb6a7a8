"""Print the handful of ncu raw-page metrics used when tuning the DFT kernels.
usage: ncu_keys.py REPORT.ncu-rep"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, v = r[0], r[-1]
pat = re.compile(r'(gpu__time_duration.sum$|sm__warps_active.avg.pct|pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed|'
                 r'smsp__issue_active.avg.pct|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|smsp__inst_executed.sum$|'
                 r'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active|sm__cycles_elapsed.max$|l1tex__m_xbar2l1tex_read_bytes.sum$|'
                 r'sm__pipe_fmaheavy_cycles_active.avg.pct|sm__inst_executed_pipe_[a-z0-9_]+.sum$|launch__registers_per_thread|'
                 r'dram__bytes_(read|write).sum$|smsp__inst_executed_op_tma|sm__throughput.avg.pct)')
for k, x in zip(h, v):
    if pat.search(k):
        try:
            if float(x.replace(',', '')) == 0:
                continue
        except ValueError:
            pass
        print(k, x)
