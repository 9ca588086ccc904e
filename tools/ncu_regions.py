#!/usr/bin/env python3
"""Per-function-region summary of an ncu source page (cuda view): dynamic warp instructions, all stall
samples and the no-instruction / long-scoreboard / wait samples, by line range of bc7.cu.
usage: ncu_regions.py report.ncu-rep kernel_regex [launch_skip]"""
import bisect, collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{kern}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
lines = open('fastc_b200/csrc/bc7.cu').read().split('\n')
marks = ['struct QeEndpoints', '__device__ __noinline__ Chain decode_chain', '__device__ __forceinline__ float div_small',
         '__device__ __forceinline__ uint32_t qe_cluster(', '__device__ __forceinline__ void fit_pca', '// ---- k-means over the NB',
         '// ---- least squares endpoints', '__device__ __noinline__ void fit_finish(', '__device__ __forceinline__ int sort_key(',
         '__device__ __forceinline__ void setup_variant(', '    // scalar k-means over the alpha', '__device__ __forceinline__ void setup_chain(',
         '__device__ __forceinline__ int chain_pixels']
starts = [(1, 'top')]
for m in marks:
    for i, l in enumerate(lines):
        if m in l:
            starts.append((i + 1, m.strip()[:40])); break
starts.sort(); keys = [s[0] for s in starts]
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0]); hdr = None; fname = ''
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != '-': continue  # keep the per-source-line summary rows
    try: ln = int(r[0])
    except ValueError: continue
    name = starts[bisect.bisect_right(keys, ln) - 1][1] if fname == 'bc7.cu' else fname
    def f(k):
        try: return float(r[ix[k]] or 0)
        except ValueError: return 0.0
    v = agg[name]
    v[0] += f('Instructions Executed'); v[1] += f('# Samples'); v[2] += f('stall_no_inst'); v[3] += f('stall_long_sb'); v[4] += f('stall_wait'); v[5] += f('Thread Instructions Executed')
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f"warp inst {ti:.3e}  samples {ts:.0f}  no_inst {sum(v[2] for v in agg.values())/ts*100:.1f}%  long_sb {sum(v[3] for v in agg.values())/ts*100:.1f}%  wait {sum(v[4] for v in agg.values())/ts*100:.1f}%")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:42s} inst {v[0]/ti*100:5.1f}%  thr {v[5]/max(v[0],1):4.1f}  samples {v[1]/ts*100:5.1f}%  no_inst {v[2]/ts*100:5.1f}%  long_sb {v[3]/ts*100:5.1f}%  wait {v[4]/ts*100:5.1f}%")
