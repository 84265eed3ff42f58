#!/usr/bin/env python
"""Summarise an .ncu-rep (run here, no GPU needed): headline metrics + per-basic-block SASS profile.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [out.md]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']


def ncu(rep, page, extra=()):
    return subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', *extra], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(ncu(rep, 'raw'))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu summary of {rep}\n", file=out)
    print(f"kernel: `{data[0][hdr.index('Kernel Name')]}`  ({len(data)} launch(es) captured)\n", file=out)
    print("| metric | unit | value(s) |\n|---|---|---|", file=out)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"| {w} | {units[i]} | {', '.join(r[i] for r in data)} |", file=out)
    rows = list(csv.reader(io.StringIO(ncu(rep, 'source', ['--print-source', 'sass']))))
    idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    hdr = rows[idx[0]]
    body = rows[idx[0] + 1: (idx[1] - 1) if len(idx) > 1 else None]
    H = {h: i for i, h in enumerate(hdr)}
    blocks, cur = [], None
    ti = tt = 0

    def f(r, k):
        try:
            return int(float(r[H[k]] or 0))
        except Exception:
            return 0
    for r in body:
        if len(r) < len(hdr):
            continue
        ie, te, sm = f(r, 'Instructions Executed'), f(r, 'Thread Instructions Executed'), f(r, '# Samples')
        op = (r[H['Source']].strip().split() or [''])[0]
        ti += ie; tt += te
        if cur and cur['ie'] == ie:
            cur['n'] += 1; cur['te'] += te; cur['s'] += sm; cur['ops'].append(op)
            for k in ('long_sb', 'short_sb', 'wait', 'branch_resolving', 'no_inst', 'not_selected'):
                cur[k] += f(r, 'stall_' + k)
        else:
            cur = {'addr': r[0], 'ie': ie, 'n': 1, 'te': te, 's': sm, 'ops': [op]}
            for k in ('long_sb', 'short_sb', 'wait', 'branch_resolving', 'no_inst', 'not_selected'):
                cur[k] = f(r, 'stall_' + k)
            blocks.append(cur)
    ts = sum(b['s'] for b in blocks) or 1
    print(f"\nSASS: {ti} warp instructions, {tt / max(ti, 1):.2f} active threads per instruction, {ts} stall samples\n", file=out)
    print("| addr | #instr | executions | share of warp-inst | avg threads | samples | long_sb | short_sb | wait | branch | first ops |\n|---|---|---|---|---|---|---|---|---|---|---|", file=out)
    for b in blocks:
        w = b['ie'] * b['n']
        if w / max(ti, 1) > 0.004 or b['s'] / ts > 0.006:
            print(f"| {b['addr'][-5:]} | {b['n']} | {b['ie']} | {w / ti:.3f} | {b['te'] / max(w, 1):.1f} | {b['s'] / ts:.3f} | {b['long_sb'] / ts:.3f} | "
                  f"{b['short_sb'] / ts:.3f} | {b['wait'] / ts:.3f} | {b['branch_resolving'] / ts:.3f} | {' '.join(b['ops'][:5])} |", file=out)


if __name__ == '__main__':
    main()
