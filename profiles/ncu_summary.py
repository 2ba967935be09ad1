"""Writes the judged subset of an ncu report as JSON.
usage: python profiles/ncu_summary.py report.ncu-rep out.json ["note"]"""
import csv
import json
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__cluster_size', 'launch__cluster_max_active',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__inst_executed.sum',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'lts__t_sector_hit_rate.pct',
]
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
out = []
for vals in rows[2:]:
  d = {}
  for h, u, v in zip(hdr, units, vals):
    if h == 'Kernel Name':
      d[h] = v
    elif h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h and
                       float(v or 0) >= 0.25):
      d[f'{h} [{u}]' if u else h] = v
  out.append(d)
res = out[0] if len(out) == 1 else {'launches': out}
if len(sys.argv) > 3:
  res['note'] = sys.argv[3]
json.dump(res, open(sys.argv[2], 'w'), indent=1)
print(json.dumps(res, indent=1)[:3000])
