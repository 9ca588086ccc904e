#!/usr/bin/env python3
"""Static SASS instruction count per source line of one kernel: sass_lines.py lib.so kernel_substring [top_n]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.startswith("bc7")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sec = cur = None
cnt = collections.Counter()
for l in dis.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
    if m: sec = m.group(1); cur = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if sec and kern in sec and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): cnt[cur] += 1
src = open(os.path.join(root, "fastc_b200/csrc/bc7.cu")).read().split('\n')
print("total", sum(cnt.values()))
for (f, ln), v in cnt.most_common(top):
    print(f"{v:4d} {f}:{ln:5d} {src[ln-1].strip()[:110] if f == 'bc7.cu' else ''}")
