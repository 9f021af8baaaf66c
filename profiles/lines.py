#!/usr/bin/env python
"""Attribute an ncu source-page CSV (SASS view) to source lines with nvdisasm's line table.

    python profiles/lines.py <prof.source.csv[.gz]> <object.o> <kernel-substr> [top] [source-dir]

(the object must be the build that was profiled: `unmatched` counts opcode disagreements)

The SASS page of `ncu --page source --csv` carries per-instruction counters but no line numbers;
`nvdisasm -g` of the same build carries the line of every instruction offset.  Joined on the
offset (after checking that the opcodes agree), this prints instructions executed and stall
samples per source line and per function-sized bucket of pmoc_device.cuh / pmoc_model.cu.
"""
import csv, gzip, io, os, re, subprocess, sys, tempfile
from collections import defaultdict


def line_table(obj, kernel):
  tmp = tempfile.mkdtemp()
  subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=tmp, capture_output=True)
  cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
  txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
  table, cur, on = {}, None, False
  for ln in txt.splitlines():
    if ln.startswith('.text.'):
      on = kernel in ln
      continue
    if not on:
      continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
      cur = (os.path.basename(m.group(1)), int(m.group(2)))
      continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
      table[int(m.group(1), 16)] = (cur, m.group(2).strip())
  return table


def functions(path):
  """(start line, name) of the device functions / lambdas in a source file (crude, by regex)."""
  out = []
  for i, ln in enumerate(open(path), 1):
    m = re.match(r'\s*(?:PM_DEV|PM_COLD|PM_GLOBAL|template.*PM_DEV)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(', ln)
    if m and not ln.strip().startswith('//'):
      out.append((i, m.group(1)))
    m = re.match(r'\s*auto (\w+) = \[&\]', ln)
    if m:
      out.append((i, 'lambda ' + m.group(1)))
  return out


def main():
  src, obj, kernel = sys.argv[1:4]
  top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
  srcdir = sys.argv[5] if len(sys.argv) > 5 else None
  op = gzip.open if src.endswith('.gz') else open
  rows = list(csv.reader(io.StringIO(op(src, 'rt').read())))
  h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
  hdr, rows = rows[h], rows[h + 1:]
  ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
  table = line_table(obj, kernel)
  base = int(rows[0][ia], 16)
  per_line, per_line_s = defaultdict(int), defaultdict(int)
  tot = tots = bad = 0
  for r in rows:
    off = int(r[ia], 16) - base
    ent = table.get(off)
    if ent is None or ent[1].split()[0].lstrip('@!P0123456789T ') .split('.')[0] != r[isrc].split()[0].lstrip('@!P0123456789T ').split('.')[0]:
      bad += 1
    key = ent[0] if ent else ('?', 0)
    n, s = int(r[iex] or 0), int(r[ismp] or 0)
    per_line[key] += n
    per_line_s[key] += s
    tot += n
    tots += s
  print('%d SASS instructions, %d unmatched; %d warp-instructions executed, %d stall samples' % (len(rows), bad, tot, tots))
  root = srcdir or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pymoc_b200', 'csrc')
  buckets, buckets_s = defaultdict(int), defaultdict(int)
  fcache = {}
  for (f, l), n in per_line.items():
    if f not in fcache:
      p = os.path.join(root, f)
      fcache[f] = functions(p) if os.path.exists(p) else []
    name = '?'
    for start, fn in fcache[f]:
      if start <= l:
        name = fn
    buckets[(f, name)] += n
    buckets_s[(f, name)] += per_line_s[(f, l)]
  print('\nper function (attributed by line ranges; inlined code counts where it was written):')
  for k, n in sorted(buckets.items(), key=lambda kv: -kv[1])[:top]:
    print('  %-22s %-26s inst %5.1f%%   samples %5.1f%%' % (k[0], k[1], 100. * n / tot, 100. * buckets_s[k] / max(tots, 1)))
  print('\nper line:')
  for k, n in sorted(per_line.items(), key=lambda kv: -kv[1])[:top]:
    print('  %-22s %5d  inst %5.1f%%   samples %5.1f%%' % (k[0], k[1], 100. * n / tot, 100. * per_line_s[k] / max(tots, 1)))


if __name__ == '__main__':
  try:
    main()
  except BrokenPipeError:
    pass
