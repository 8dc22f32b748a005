"""Instruction count per source line of one function of a built object: python scripts/sass_lines.py <obj.o> <function-substring> [top]"""
import re, collections, subprocess, sys, tempfile, os
obj, fn_sub = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp(); subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(d, cub)], capture_output=True, text=True).stdout
cur = None; fn = None; cnt = collections.Counter(); fcnt = collections.Counter()
for l in txt.splitlines():
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'^(\S+):\s*$', l)
    if m and not l.startswith('.L'): fn = m.group(1)
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): cnt[(fn, cur)] += 1; fcnt[fn] += 1
for f, c in fcnt.most_common(20): print(c, f[-60:])
bl = collections.Counter()
for (f, c), v in cnt.items():
    if f and fn_sub in f: bl[c] += v
print()
for c, v in sorted(bl.items(), key=lambda x: -x[1])[:top]: print(v, c)
