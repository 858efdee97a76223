#!/usr/bin/env python
"""Offline SASS statistics for one kernel of libsvb200.so: opcode histogram and register use.

    python tools/sass_stats.py <substring of the mangled or demangled kernel name> [--dump]
"""
import collections
import re
import subprocess
import sys

LIB = 'supervillain_b200/libsvb200.so'


def main():
    pat = sys.argv[1]
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    blocks = re.split(r'\n\s*Function : ', out)
    for b in blocks[1:]:
        name = b.split('\n', 1)[0].strip()
        dem = subprocess.run(['cu++filt', name], capture_output=True, text=True).stdout.strip()
        if pat not in name and pat not in dem:
            continue
        ops = collections.Counter()
        lines = re.findall(r'/\*[0-9a-f]{4}\*/\s+(.*?);', b)
        for l in lines:
            m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
            if m:
                ops[m.group(2)] += 1
        print(dem, '\n  instructions:', len(lines))
        print('  ' + ', '.join(f'{k}:{v}' for k, v in ops.most_common(30)))
        if '--dump' in sys.argv:
            print('\n'.join(lines))


if __name__ == '__main__':
    main()
