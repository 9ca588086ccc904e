#!/usr/bin/env python3
"""Headline benchmark: BC7 (BPTC) -q 50 on a synthetic 8192x8192 RGBA texture
(BASELINE.json configs[2]; it fits one B200), blocks sharded over N GPUs.

  python bench.py [--gpus N] [--steps K] [--warmup W]          # our CUDA path
  python bench.py --impl reference [...]                        # the reference's CPU path

One "step" = one pass of the hot path over the whole texture: every 4x4 block
encoded (classify, shape select, endpoint fits + annealing, pack) and, for N > 1,
the compressed slabs gathered to rank 0 over NCCL/NVLink.

`value`  = Mpix/s with the input slab already resident in HBM (CUDA events on the
           launching stream, max over ranks).
`e2e`    = the same metric through the C ABI's host-pointer entry
           (fastc_gpu_compress: what FasTC's CompressImageData binds to) with
           pinned HOST buffers -- H2D and D2H inside the timed region.
`roofline` describes the dominant kernel (bc7_anneal).  BC7 is an ALU-issue
           bound per-block search, so besides the HBM figure the schema asks for
           we report lane-instruction throughput against the chip's issue peak.
`cpu_baseline` = the UNMODIFIED reference (oracle/_ref/libfastc_ref.so, built from
           /root/reference by oracle/Makefile) on this box's host cores, on a bounded
           slab of the same texture.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WIDTH = HEIGHT = 8192
QUALITY = 50
SEED = 1
METRIC = "BC7 Mpix/s at -q 50 (1/2/4/8 B200) vs ref CPU all cores; PSNR delta vs ref"
WORKLOAD = "BPTC (BC7) -q 50, synthetic 8192x8192 RGBA (SURVEY 8d generator, seed 1), block rows sharded over N GPUs"

# op model of SURVEY.md 8(d): lane-ops per block at -q 50 on this generator, and the ALU issue peak
LANE_OPS_PER_BLOCK = 1.2e6
ALU_PEAK_LANE_OPS = 148 * 128 * 1.965e9
ALGO_BYTES_PER_BLOCK = 64 + 16  # read one 4x4 RGBA block, write one 128-bit BC7 block (5 B/px)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE bc7_anneal launch over the whole 8192^2 texture
# (ncu, profiles/r01_v14_dram_8192.csv): the sorted start states, the pixels of every chain, the results
ANNEAL_DRAM_BYTES_8192 = 7_862_972_416 + 1_578_225_920


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners to fd 1 (NCCL prints its
    version there), so fd 1 is pointed at stderr for the whole run and the line is written to the
    saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6537.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def shard_rows(rank: int, world: int):
    """Contiguous block-row slab of rank `rank` (fastc_b200/sharding.py)."""
    from fastc_b200.sharding import shard_block_rows
    return shard_block_rows(HEIGHT // 4, rank, world)


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "tests"))
    from _checkers import Reference
    from fastc_b200.synth import synth_rgba
    cores = os.cpu_count() or 1
    threads = min(cores, 256)  # ThreadGroup cap (reference Core/src/ThreadGroup.h:88)
    # bounded sample: the top 8192 x rows slab of the same texture (same generator coordinates);
    # blocks are independent, so Mpix/s on the slab is representative of the full image.
    rows = 64
    img = synth_rgba(WIDTH, rows, SEED, full_height=HEIGHT)
    ref = Reference()
    times = []
    for i in range(args.warmup + args.steps):
        _, ms = ref.compress("BPTC", img, quality=QUALITY, threads=threads, seed=None)
        if i >= args.warmup:
            times.append(ms)
    ms = sum(times) / len(times)
    val = WIDTH * rows / 1e6 / (ms / 1e3)
    sample = f"top {WIDTH}x{rows} slab of the 8192^2 texture per step, CompressImageData -t {threads} (static split)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist
    from fastc_b200 import ECompressionFormat as F, lib
    from fastc_b200.synth import synth_rgba_torch

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    g = lib()
    multi = world > 1

    # ---- this rank's slab of the texture, generated on the device, plus a pinned host copy
    r0, r1 = shard_rows(rank, world)
    rows = (r1 - r0) * 4
    nblk = (r1 - r0) * (WIDTH // 4)
    d_in = synth_rgba_torch(WIDTH, rows, SEED, y0=r0 * 4, full_height=HEIGHT, device=dev)
    d_out = torch.zeros(nblk * 16, dtype=torch.uint8, device=dev)
    h_in = torch.empty((rows, WIDTH, 4), dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_out = torch.empty(nblk * 16, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    # watermark chain across ranks (one integer per rank)
    from fastc_b200.sharding import gather_slabs, watermark_base
    my_solid = g.count_solid_device(d_in, width=WIDTH, height=rows)
    wm_base = watermark_base(my_solid, rank, world, device=dev)
    base_idx = r0 * (WIDTH // 4)

    sizes = [(shard_rows(r, world)[1] - shard_rows(r, world)[0]) * (WIDTH // 4) * 16 for r in range(world)]
    gather_list = None
    if multi and rank == 0:
        gather_list = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    launches = 0

    def step_device():
        nonlocal launches
        n = g.compress_device(F.BPTC, d_in, d_out, width=WIDTH, height=rows, quality=QUALITY, seed=SEED,
                              wm_base=wm_base, block_index_base=base_idx)
        launches += n
        if multi:
            # exchange step named by north_star: compressed slabs gathered to rank 0 over NVLink
            # (send/recv of exact sizes: slabs can differ by one block row)
            gather_slabs(d_out, rank, world, sizes, gather_list)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also arms the per-stage timing of the dominant kernel)
    g.bc7_stage_ms(enable=True, read=False)
    for _ in range(args.warmup):
        step_device()
    barrier()

    # ---- timed: K steps, each bracketed by events on the launching (torch current) stream;
    # L2 flushed between steps (outside the event pairs)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    chains_ms, stage_tot = [], []
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        step_device()
        ev[k][1].record()
        st = g.bc7_stage_ms(enable=True, read=True)  # syncs on this step's stage events
        chains_ms.append(st["anneal"]); stage_tot.append(st)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    device_launches = launches

    # ---- end to end through the C ABI host entry, pinned host buffers, H2D + D2H timed
    h_in_np, h_out_np = h_in.numpy(), h_out.numpy()
    e2e_launches = 0
    for _ in range(args.warmup):
        g.compress(F.BPTC, h_in_np, h_out_np, quality=QUALITY, seed=SEED)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_steps_ms = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        _, tm = g.compress(F.BPTC, h_in_np, h_out_np, quality=QUALITY, seed=SEED)
        e2e_steps_ms.append((time.perf_counter() - ts) * 1e3)
        h2d, d2h = tm["h2d_bytes"], tm["d2h_bytes"]
        e2e_launches += tm["kernel_launches"]
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        tb = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
        dist.all_reduce(tb)
        h2d, d2h = int(tb[0]), int(tb[1])
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        return

    mpix = WIDTH * HEIGHT / 1e6
    ms_per_step = total_ms / args.steps
    value = mpix / (ms_per_step / 1e3)
    e2e_value = mpix / (e2e_s / args.steps)

    # ---- roofline of the dominant kernel (bc7_anneal), rank 0's slab
    hbm_peak, peak_src = load_peaks()
    k_ms = sum(chains_ms) / len(chains_ms)
    share = k_ms / (sum(s["total"] for s in stage_tot) / len(stage_tot))
    achieved_gbs = nblk * ALGO_BYTES_PER_BLOCK / (k_ms / 1e3) / 1e9
    lane_ops = nblk * LANE_OPS_PER_BLOCK / (k_ms / 1e3)
    roofline = {
        "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved_gbs / hbm_peak,
        # measured for the full texture; a rank's slab moves its share of it
        "traffic": (ANNEAL_DRAM_BYTES_8192 * nblk / ((WIDTH // 4) * (HEIGHT // 4))) if WIDTH == 8192 else None,
        "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r01_v14_dram_8192.csv)",
        "algorithmic_bytes": nblk * ALGO_BYTES_PER_BLOCK, "kernel": "bc7_anneal",
        "kernel_ms": k_ms, "kernel_share_of_step": share, "peak_source": peak_src,
        "note": "BC7 is ALU-issue bound, not HBM bound (SURVEY 8d): see alu_issue",
        "alu_issue": {"achieved_lane_ops_per_s": lane_ops, "peak_lane_ops_per_s": ALU_PEAK_LANE_OPS,
                      "frac": lane_ops / ALU_PEAK_LANE_OPS,
                      "model": "1.2e6 lane-ops per block at -q 50 (SURVEY 8d op model of the REFERENCE's instruction "
                               "mix) / bc7_anneal time; peak = 148 SM x 128 lanes x 1.965 GHz.  The GPU formulation needs "
                               "fewer instructions than the model, so this overstates utilisation: the measured figure is "
                               "ncu's issue-slot utilisation in profiles/ (86.6 % with 28.3 of 32 lanes active)"},
        "stages_ms": {k: sum(s[k] for s in stage_tot) / len(stage_tot) for k in stage_tot[0]},
    }

    # ---- CPU baseline: the reference itself on this box's host cores, bounded slab (N = 1 only)
    cpu_baseline = None
    quality_check = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "tests"))
        from _checkers import Reference, Oracle
        cores = os.cpu_count() or 1
        rows_s = 128
        img_s = np.ascontiguousarray(h_in_np[:rows_s])
        if Reference.available():
            ref = Reference()
            threads = min(cores, 256)
            ref_cmp, ms = ref.compress("BPTC", img_s, quality=QUALITY, threads=threads, seed=None)
            kind = "reference"
            how = f"CompressImageData -t {threads}"
            # the metric's "PSNR delta vs ref": both outputs of that slab decoded by the reference's own
            # decompressor, PSNR by the reference's formula (Base/src/Image.cpp:205-255)
            ours_cmp = h_out_np[: ref_cmp.size]  # rank 0's slab starts at block row 0
            p_ref = ref.psnr(img_s, ref.decode("BPTC", ref_cmp, WIDTH, rows_s))
            p_gpu = ref.psnr(img_s, ref.decode("BPTC", np.ascontiguousarray(ours_cmp), WIDTH, rows_s))
            same = float((ours_cmp.reshape(-1, 16) == ref_cmp.reshape(-1, 16)).all(1).mean())
            quality_check = {"psnr_gpu_db": p_gpu, "psnr_ref_db": p_ref, "delta_db": p_gpu - p_ref,
                             "bit_identical_block_fraction": same, "tolerance_db": 0.05,
                             "sample": f"top {WIDTH}x{rows_s} slab, decoded by the reference's decompressor"}
        else:  # reference .so did not travel: time the oracle port (single thread)
            t0 = time.perf_counter()
            Oracle().compress("BPTC", img_s, quality=QUALITY, rng_mode=0)
            ms = (time.perf_counter() - t0) * 1e3
            kind, threads, how = "port", 1, "oracle restatement, 1 thread"
        cpu_baseline = {"value": WIDTH * rows_s / 1e6 / (ms / 1e3), "unit": "Mpix/s", "cores": threads,
                        "kind": kind, "sample": f"top {WIDTH}x{rows_s} slab of the same texture, one pass, {how}",
                        "ms": ms}

    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "quality": QUALITY, "width": WIDTH, "height": HEIGHT,
                   "parallelism": f"block-row slabs x{world}" + (", compressed output gathered to rank 0 (NCCL send/recv)" if multi else ""),
                   "l2": "256 MiB flush between timed steps (input slab is also > L2 at N=1)",
                   "timing": "per-step CUDA events on the launching stream, summed; max over ranks"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3, "rank0_ms_steps": e2e_steps_ms,
                "path": "fastc_gpu_compress (C ABI, pinned host in/out) per rank"},
        "gpu_launches": device_launches + e2e_launches,
        "roofline": roofline,
        "wall_s_timed_region": t_wall,
    }
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if quality_check:
        line["psnr_vs_reference"] = quality_check
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--size", type=int, default=8192,
                    help="profiling only: square texture size (the reported benchmark is the default 8192)")
    args = ap.parse_args()
    global WIDTH, HEIGHT
    WIDTH = HEIGHT = args.size
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
