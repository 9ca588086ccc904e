#!/usr/bin/env python3
"""Builds profiles/r02_instr_counts.json -- what bench.py's `roofline` divides by the live step time --
from one ncu metrics pass over ONE BC7 -q 50 compression of the 8192^2 bench texture:

    ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,\
dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:bc7_ --csv \
        --log-file gpurun_out/r02_counts_8192.csv python tools/profile_run.py BPTC 8192 50 1
    python tools/collect_counts.py gpurun_out/r02_counts_8192.csv [counters.json]

Instruction counts of a fixed input are reproducible up to the scheduling of the persistent annealing
kernel (which lanes share a warp), a fraction of a percent.  `counters.json` (optional) holds the
FASTC_GPU_COUNTERS build's measured work: {"qe_calls": .., "pixel_bucket_evals": ..} (tools/measure_counters.py)."""
import csv
import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def kernel_source_sha():
    h = hashlib.sha256()
    for f in sorted((ROOT / "fastc_b200" / "csrc").glob("*")):  # the BC7 kernels' sources (the counts are theirs)
        if f.name in ("bc7.cu", "bc7_tables.cuh", "common.cuh"):
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


def main():
    src = Path(sys.argv[1])
    rows = list(csv.reader(l for l in src.read_text().splitlines() if l.startswith('"')))
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    per = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].split("::")[-1].split("<")[0]
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)  # -> ms
        per.setdefault(name, {}).setdefault(r[mi], 0.0)
        per[name][r[mi]] += v
    out = {
        "what": "per-launch counters of the BC7 kernels, one -q 50 compression of the synthetic 8192^2 texture (seed 1)",
        "command": "see tools/collect_counts.py", "source_csv": f"profiles/{src.name}",
        "kernel_source_sha": kernel_source_sha(),
        "thread_inst_per_launch_8192": {k: v.get("smsp__thread_inst_executed.sum", 0.0) for k, v in per.items()},
        "warp_inst_per_launch_8192": {k: v.get("smsp__inst_executed.sum", 0.0) for k, v in per.items()},
        "dram_bytes_per_launch_8192": {k: v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
                                       for k, v in per.items()},
        "ncu_ms_per_launch_8192": {k: v.get("gpu__time_duration.sum", 0.0) for k, v in per.items()},
    }
    out["dram_bytes_per_step_8192"] = sum(out["dram_bytes_per_launch_8192"].values())
    if len(sys.argv) > 2:
        out["work_counters"] = json.loads(Path(sys.argv[2]).read_text())
    dst = ROOT / "profiles" / "r02_instr_counts.json"
    dst.write_text(json.dumps(out, indent=1) + "\n")
    (ROOT / "profiles" / src.name).write_text(src.read_text())
    tot = sum(out["thread_inst_per_launch_8192"].values())
    print(f"{dst}: {tot:.4e} thread-instructions per step, {out['dram_bytes_per_step_8192'] / 1e9:.2f} GB DRAM traffic")
    for k in out["thread_inst_per_launch_8192"]:
        w = out["warp_inst_per_launch_8192"][k] or 1
        print(f"  {k:18s} thread-inst {out['thread_inst_per_launch_8192'][k]:.4e}  lanes/inst "
              f"{out['thread_inst_per_launch_8192'][k] / w:5.2f}  dram {out['dram_bytes_per_launch_8192'][k] / 1e9:6.2f} GB  "
              f"ncu {out['ncu_ms_per_launch_8192'][k]:8.3f} ms")


if __name__ == "__main__":
    main()
