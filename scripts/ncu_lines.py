#!/usr/bin/env python
"""usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [top]  -- per CUDA source line of every file the kernel's
SASS maps to: warp instructions, active lanes per instruction, stall samples (ncu source page, -lineinfo builds)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, lines = None, None, []
for r in rows:
    if len(r) == 2 and r[0] in ('File Name', 'File Path'):
        cur = r[1].split('/')[-1]; hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or not r[0].isdigit():
        continue
    d = dict(zip(hdr, r))
    try:
        inst = int(d.get('Instructions Executed') or 0); tinst = int(d.get('Thread Instructions Executed') or 0)
        smp = int(d.get('# Samples') or 0)
    except ValueError:
        continue
    if inst or smp:
        lines.append((cur, int(r[0]), r[1].strip()[:110], inst, tinst, smp,
                      {k: int(d[k] or 0) for k in d if k.startswith('stall_') and 'Not Issued' not in k and (d[k] or '0').isdigit()}))
ti = sum(l[3] for l in lines); ts = sum(l[5] for l in lines)
print('total warp inst %d  thread inst %d  lanes %.1f  samples %d' % (ti, sum(l[4] for l in lines), sum(l[4] for l in lines) / max(ti, 1), ts))
print('--- by instructions')
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    st = sorted(l[6].items(), key=lambda kv: -kv[1])[:2]
    print('%5.2f%% inst %5.2f%% smp %4.1f lanes  %s:%d  %s   %s' % (100 * l[3] / ti, 100 * l[5] / max(ts, 1), l[4] / max(l[3], 1), l[0], l[1], l[2],
                                                                  ' '.join('%s=%d' % (k[6:], v) for k, v in st if v)))
print('--- by samples')
for l in sorted(lines, key=lambda l: -l[5])[:top // 2]:
    st = sorted(l[6].items(), key=lambda kv: -kv[1])[:3]
    print('%5.2f%% inst %5.2f%% smp %4.1f lanes  %s:%d  %s   %s' % (100 * l[3] / ti, 100 * l[5] / max(ts, 1), l[4] / max(l[3], 1), l[0], l[1], l[2],
                                                                  ' '.join('%s=%d' % (k[6:], v) for k, v in st if v)))
