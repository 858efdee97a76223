#!/usr/bin/env python
"""Executed warp instructions and stall samples per SOURCE LINE for the first kernel of an .ncu-rep.

Joins the SASS view of the report (per-instruction execution counts) with `nvdisasm -g` line info of the same kernel
from the library the report was taken with (build with -lineinfo).

    python tools/ncu_lines.py prof.ncu-rep <mangled kernel name substring> [sites_per_launch] [lib.so]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, pat = sys.argv[1], sys.argv[2]
sites = float(sys.argv[3]) if len(sys.argv) > 3 else None
lib = sys.argv[4] if len(sys.argv) > 4 else 'supervillain_b200/libsvb200.so'

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
line_of = {}
matched = []
for f in os.listdir(tmp):
    if not f.endswith('sm_100a.cubin'):
        continue
    out = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fun, cur_line, active = None, None, False
    for l in out.split('\n'):
        m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
        if m:
            # ONE function only: a pattern that matches several instantiations of a kernel would mix their line tables
            # (addresses are offsets within a function) -- the first match is taken, the others are reported
            hit = pat in m.group(1)
            if hit and matched and m.group(1) != matched[0]:
                print(f'# note: pattern also matches {m.group(1)[:100]} (ignored: give a longer pattern to select it)')
                hit = False
            if hit and not matched:
                matched.append(m.group(1))
            active = hit
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            line_of[int(m.group(1), 16)] = cur_line

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
st = starts[0]
hdr = rows[st]
end = starts[1] - 1 if len(starts) > 1 else len(rows)
iA, iE, iN = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
base = None
per_line_instr, per_line_samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[st + 1:end]:
    try:
        addr = int(r[iA], 16); e = int(r[iE]); sm = int(r[iN] or 0)
    except (ValueError, IndexError):
        continue
    if base is None:
        base = addr
    key = line_of.get(addr - base, ('?', 0))
    per_line_instr[key] += e
    per_line_samp[key] += sm
    tot += e
files = {}
def srcline(key):
    fn, ln = key
    for root in ('supervillain_b200/csrc',):
        p = os.path.join(root, fn)
        if os.path.exists(p):
            if p not in files:
                files[p] = open(p).read().split('\n')
            return files[p][ln - 1].strip()[:90] if 0 < ln <= len(files[p]) else ''
    return ''
tots = sum(per_line_samp.values())
print(f'total warp instructions {tot}' + (f' = {32 * tot / sites:.1f} thread-instr per site-update' if sites else ''))
for key, c in per_line_instr.most_common(int(os.environ.get('NCU_LINES_TOP', 60))):
    per = f'{32 * c / sites:6.1f}/site' if sites else ''
    print(f'{key[0][:24]:24s}:{key[1]:4d} {100 * c / tot:5.1f}% {per} samp {100 * per_line_samp[key] / max(tots, 1):4.1f}%  {srcline(key)}')
