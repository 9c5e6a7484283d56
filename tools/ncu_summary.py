#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): headline metrics and a per-source-line breakdown.
usage: python tools/ncu_summary.py <report.ncu-rep> [mangled-kernel-substring [demangled-name-substring]]"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__inst_executed.avg.per_cycle_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def run(*a):
    return subprocess.run(a, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(run('ncu', '-i', rep, '--page', 'raw', '--csv'))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print('kernel:', vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'  {w:80s} {vals[i]:>16s} {units[i]}')
    if len(sys.argv) < 3:
        return
    # per-source-line: SASS metrics from the report joined with nvdisasm -g line info of the in-tree .so
    sub = sys.argv[2]
    import glob, os, tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'smudgeplot_b200/lib/libhetmers_b200.so')],
                   cwd=tmp, capture_output=True)
    sass = ''
    for f in glob.glob(os.path.join(tmp, '*.cubin')):
        sass += run('nvdisasm', '-g', '-c', f)
    m = re.search(r'\.text\.(\S*' + re.escape(sub) + r'\S*):\n(.*?)\n\t\.section', sass, re.S)
    if not m:
        print('kernel', sub, 'not found in the in-tree .so')
        return
    cur, off2line = None, {}
    for l in m.group(2).splitlines():
        mm = re.search(r'//## File "([^"]*)", line (\d+)', l)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if mm:
            off2line[int(mm.group(1), 16)] = cur
    rows = list(csv.reader(io.StringIO(run('ncu', '-i', rep, '--page', 'source', '--csv'))))
    # one section per profiled launch ("Kernel Name" row, header row, instruction rows): take the first
    # section whose kernel name matches the demangled form of `sub` (or simply the first one)
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
    want = sys.argv[3] if len(sys.argv) > 3 else None
    sec = 0
    for k in range(len(starts) - 1):
        if want and want in rows[starts[k]][1]:
            sec = k
            break
    hdr, data = rows[starts[sec] + 1], rows[starts[sec] + 2:starts[sec + 1]]
    ia, ii, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    base = int(data[0][ia], 16)
    agg = defaultdict(lambda: [0, 0, 0])
    ti = ts = 0
    for r in data:
        ln = off2line.get(int(r[ia], 16) - base)
        agg[ln][0] += int(r[ii]); agg[ln][1] += int(r[isamp]); agg[ln][2] += 1
        ti += int(r[ii]); ts += int(r[isamp])
    srcs = {}
    print(f'  total warp instructions {ti}, stall samples {ts}, SASS instructions {len(data)}')
    for ln, (i, s, c) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        text = '?'
        if ln:
            if ln[0] not in srcs:
                for cand in glob.glob(os.path.join(root, 'smudgeplot_b200/csrc', ln[0])):
                    srcs[ln[0]] = open(cand).read().splitlines()
            text = srcs.get(ln[0], [''] * (ln[1] + 1))[ln[1] - 1].strip()[:88]
        print(f'  {(ln[0][:14] + ":" + str(ln[1])) if ln else "None":>20s}: sass {c:4d}  inst {100 * i / ti:5.1f}%  stalls {100 * s / ts:5.1f}%  {text}')


if __name__ == '__main__':
    main()
