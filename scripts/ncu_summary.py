#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key metrics per captured launch + top stall reasons + hottest source lines.
usage: python scripts/ncu_summary.py gpurun_out/prof_step.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_allocated', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    ki = hdr.index('Kernel Name')
    for r in rows:
        print('=== %s  (id %s)' % (r[ki][:70], r[0]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-70s %14s %s' % (k, r[i], units[i]))
        stalls = [(float(r[i].replace(',', '') or 0), h) for i, h in enumerate(hdr)
                  if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
        tot = sum(s for s, _ in stalls) or 1
        print('  stall samples: ' + ', '.join('%s %.0f%%' % (h.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * s / tot)
                                              for s, h in sorted(stalls, reverse=True)[:8]))
    if '--source' in sys.argv:
        n = int(sys.argv[sys.argv.index('--source') + 1])
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        # per-file blocks; rows whose first column is a line number carry the per-source-line aggregates
        cur = None; hdr = None; lines = []; nfunc = 0
        for r in rows:
            if len(r) >= 2 and r[0] == 'File Path':
                cur = r[1].split('/')[-1]; continue
            if len(r) >= 2 and r[0] == 'Function Name':
                continue
            if len(r) > 4 and r[0] == 'Line No':
                hdr = r
                if cur and cur.startswith('crowdsim_common') :
                    nfunc += 1
                continue
            if nfunc > 1:
                break          # first captured launch only
            if hdr and len(r) == len(hdr) and r[0].isdigit():
                g = lambda name: float((r[hdr.index(name)] or '0').replace(',', '')) if name in hdr and r[hdr.index(name)] not in ('-', '') else 0.0  # noqa: E731
                lines.append((g('# Samples'), g('Instructions Executed'), g('Avg. Threads Executed'), cur, int(r[0]), r[1].strip()[:110]))
        tot_s = sum(l[0] for l in lines) or 1; tot_i = sum(l[1] for l in lines) or 1
        print('--- hottest source lines of the first captured launch: samples  %%samples  warp-instr  %%instr  avg-threads  file:line')
        for smp, ins, thr, f, ln, src in sorted(lines, key=lambda l: -l[1])[:n]:
            print('  %6.0f %5.1f%% %10.0f %5.1f%% %5.1f  %s:%d  %s' % (smp, 100 * smp / tot_s, ins, 100 * ins / tot_i, thr, f, ln, src))


if __name__ == '__main__':
    main()
