#!/usr/bin/env python3
"""A/B timing of library builds in ONE GPU session: for each given libfastc_gpu.so variant, copy it over
fastc_b200/libfastc_gpu.so and run `bench.py --no-cpu-baseline [--size N]` in a fresh process, then print
value / e2e / stage times side by side and restore the original library.

Build variants on the CPU box first (they travel with the snapshot), e.g.
    make gpu EXTRA_NVCCFLAGS=-DSOME_EXPERIMENT && cp fastc_b200/libfastc_gpu.so variants/exp.so
    git stash; make gpu; cp fastc_b200/libfastc_gpu.so variants/base.so; git stash pop
    gpurun -- 'python tools/ab_bench.py variants/base.so variants/exp.so --size 8192 --repeat 2'
(keep `variants/` out of git: it is listed in .gitignore)."""
import argparse
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "fastc_b200" / "libfastc_gpu.so"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+", help="libfastc_gpu.so variants")
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    backup = LIB.with_suffix(".so.orig")
    shutil.copy2(LIB, backup)
    try:
        for rep in range(args.repeat):
            for lib in args.libs:
                shutil.copy2(lib, LIB)
                r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--no-cpu-baseline", "--no-other-configs", "--size", str(args.size),
                                    "--steps", str(args.steps)], capture_output=True, text=True)
                if r.returncode != 0:
                    print(f"{lib}: bench failed\n{r.stderr[-400:]}")
                    continue
                d = json.loads(r.stdout.strip().splitlines()[-1])
                st = {k: round(v, 2) for k, v in d["roofline"]["stages_ms"].items()}
                print(f"{Path(lib).name:24s} run {rep}: {d['value']:7.1f} Mpix/s  e2e {d['e2e']['value']:7.1f}  {st}", flush=True)
    finally:
        shutil.copy2(backup, LIB)
        backup.unlink()


if __name__ == "__main__":
    main()
