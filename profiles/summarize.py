#!/usr/bin/env python
"""Turn Nsight Compute output into the short text summaries committed under profiles/.

    python profiles/summarize.py launches <launch-list.csv>          # --metrics gpu__time_duration.sum pass
    python profiles/summarize.py full <report.ncu-rep> [kernel-substr]   # --set full capture
    python profiles/summarize.py raw <raw-page.csv> [kernel-substr]      # `ncu -i rep --page raw --csv` of such a capture

The launch list gives each kernel's share of the profiled command; the full capture gives
the counters quoted in DESIGN.md / bench.py (FP64 pipe utilisation, DRAM bytes, occupancy,
stall reasons).
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = (
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'dram__bytes_write.sum.per_second', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__warps_active.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_warps', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
    'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg',
    'sm__cycles_elapsed.max', 'gpc__cycles_elapsed.max', 'lts__t_sector_hit_rate.pct',
)
STALL = 'smsp__average_warps_issue_stalled_'


def launches(path):
  rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
  hdr, rows = rows[0], rows[1:]
  k, v = hdr.index('Kernel Name'), hdr.index('Metric Value')
  agg = OrderedDict()
  for r in rows:
    name = r[k].split('(')[0][:90]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[v].replace(',', '')))
  total = sum(t for _, t in agg.values())
  print('launch list: %s  (%d launches, %.3f ms GPU time in total)' % (path, len(rows), total / 1e6))
  print('%-92s %6s %12s %7s' % ('kernel', 'count', 'total us', 'share'))
  for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-92s %6d %12.1f %6.1f%%' % (name, n, t / 1e3, 100 * t / total))


def full(path, substr='', is_csv=False, member_steps=0):
  out = open(path).read() if is_csv else subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                                                          capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(out)))
  hdr, units, rows = rows[0], rows[1], rows[2:]
  for r in rows:
    d = dict(zip(hdr, r))
    if substr and substr not in d['Kernel Name']:
      continue
    print('kernel: %s   grid %s x block %s' % (d['Kernel Name'], d.get('Grid Size'), d.get('Block Size')))
    for key in KEEP:
      if key in d:
        print('  %-70s %16s %s' % (key, d[key], units[hdr.index(key)]))
    stalls = sorted(((float(d[h]), h[len(STALL):].replace('_per_issue_active.ratio', '')) for h in hdr
                     if h.startswith(STALL) and h.endswith('_per_issue_active.ratio') and d[h]), reverse=True)
    print('  warp stall reasons (warps stalled per issue-active cycle):')
    for val, name in stalls[:8]:
      print('    %-40s %8.3f' % (name, val))
    # executed FP64 work: thread-level DADD / DMUL / DFMA counts (per elapsed cycle, summed over the SM sub-partitions)
    try:
      op = lambda n: float(d['smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed' % n])
      cyc = float(d['smsp__cycles_elapsed.avg'])
      dadd, dmul, dfma = op('dadd'), op('dmul'), op('dfma')
      flops = (dadd + dmul + 2 * dfma) * cyc
      print('  executed FP64: DADD %.3e  DMUL %.3e  DFMA %.3e thread instructions -> %.4e flops in this launch'
            % (dadd * cyc, dmul * cyc, dfma * cyc, flops))
      print('  executed FP64 flops per cycle %.0f of 18944 (148 SM x 64 DFMA x 2) = %.1f %%' % (flops / cyc, flops / cyc / 189.44))
      if member_steps:
        print('  executed FP64 flops per member-step %.0f (%d member-steps in this launch)' % (flops / member_steps, member_steps))
    except (KeyError, ValueError):
      pass


if __name__ == '__main__':
  if sys.argv[1] == 'launches':
    launches(sys.argv[2])
  else:
    # optional 4th argument: member-steps of the captured launch (members x steps), for the per-member-step flop count
    full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '', is_csv=sys.argv[1] == 'raw',
         member_steps=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
