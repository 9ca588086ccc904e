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
           pinned HOST buffers -- H2D and D2H inside the timed region; one process
           per GPU, each on its slab.  `e2e.pageable` repeats it with the pageable
           (new[] / malloc) memory FasTC really passes.
`e2e_single_call` (N > 1) = ONE fastc_gpu_compress(num_gpus = N) call by rank 0 on the
           whole texture in one caller buffer (what SCompressionSettings::iNumGPUs /
           `tc -g N` runs), its bytes checked against the NCCL-gathered slabs.
`roofline` = the bound that binds.  BC7 is a per-block search bound by ALU
           instruction issue: achieved = thread-instructions executed by our kernels in
           one step (ncu smsp__thread_inst_executed.sum, profiles/r02_instr_counts.json)
           / the live step time; peak = 148 SMs x 128 lanes x the SM clock sampled
           under load.  The HBM figure the base schema asks for is kept under `hbm`.
`other_configs` (N = 1) = BASELINE configs 2, 4 and 5 measured in the same run, each
           with a bit-exactness check of a sample against the CPU oracle.
`cpu_baseline` = the UNMODIFIED reference (oracle/_ref/libfastc_ref.so, built from
           /root/reference by oracle/Makefile) on this box's host cores, on a bounded
           slab of the same texture.
"""
from __future__ import annotations

import argparse
import hashlib
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
REF_SAMPLE_ROWS = 128  # reference arm / cpu_baseline: top 8192 x 128 slab (1/64 of the texture) per step

ALGO_BYTES_PER_BLOCK = 64 + 16  # read one 4x4 RGBA block, write one 128-bit BC7 block (5 B/px)
SM_COUNT, LANES_PER_SM = 148, 128
INSTR_COUNTS = ROOT / "profiles" / "r02_instr_counts.json"


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
        return float(d.get("hbm_gbs", 6537.0)), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    h = hashlib.sha256()
    for f in sorted((ROOT / "fastc_b200" / "csrc").glob("*")):  # the BC7 kernels' sources (the counts are theirs)
        if f.name in ("bc7.cu", "bc7_tables.cuh", "common.cuh"):
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


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
    rows = REF_SAMPLE_ROWS
    img = synth_rgba(WIDTH, rows, SEED, full_height=HEIGHT)
    ref = Reference()
    times = []
    for i in range(args.warmup + args.steps):
        _, ms = ref.compress("BPTC", img, quality=QUALITY, threads=threads, seed=None)
        if i >= args.warmup:
            times.append(ms)
    ms = sum(times) / len(times)
    val = WIDTH * rows / 1e6 / (ms / 1e3)
    sample = (f"top {WIDTH}x{rows} slab ({100.0 * rows / HEIGHT:.1f} % of the 8192^2 texture) per step, "
              f"CompressImageData -t {threads} (static split); Mpix/s extrapolated from the slab "
              f"(blocks are independent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "quality": QUALITY, "width": WIDTH, "height": HEIGHT},
        "sample": sample,
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- helpers of our arm
def _pin(t):
    import torch
    p = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    p.copy_(t)
    torch.cuda.synchronize()
    return p


def _time_device(fn, reps, flush=None):
    """min / mean CUDA-event time of fn() over reps launches on torch's current stream."""
    import torch
    ms = []
    for k in range(reps):
        if flush is not None:
            flush.fill_(k & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return min(ms), sum(ms) / len(ms)


def _time_host(fn, reps):
    ms = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ms.append((time.perf_counter() - t0) * 1e3)
    return min(ms), sum(ms) / len(ms)


def other_configs(g, dev, flush, hbm_peak, reps=3):
    """BASELINE configs 2, 4 and 5 on one GPU: kernel-only (device-resident, CUDA events), end to end
    through the C ABI with pinned and with pageable host memory, and a bit-exactness check of a
    sample of the output against the CPU oracle."""
    import numpy as np
    import torch
    from fastc_b200 import ECompressionFormat as F
    from fastc_b200.synth import synth_rgba_torch
    sys.path.insert(0, str(ROOT / "tests"))
    from _checkers import Oracle
    orc = Oracle()
    out = {}

    # ---- config 2: BPTC -q 50, 2048^2
    size = 2048
    d_in = synth_rgba_torch(size, size, SEED, device=dev)
    nblk = (size // 4) ** 2
    d_out = torch.zeros(nblk * 16, dtype=torch.uint8, device=dev)
    run = lambda: g.compress_device(F.BPTC, d_in, d_out, width=size, height=size, quality=QUALITY, seed=SEED)
    run(); torch.cuda.synchronize()
    k_min, k_mean = _time_device(run, reps, flush)
    h_pin, h_out = _pin(d_in).numpy(), _pin(d_out).numpy()
    h_page, o_page = np.array(h_pin), np.empty_like(h_out)
    g.compress(F.BPTC, h_pin, h_out, quality=QUALITY, seed=SEED)
    g.compress(F.BPTC, h_page, o_page, quality=QUALITY, seed=SEED)
    e_pin, _ = _time_host(lambda: g.compress(F.BPTC, h_pin, h_out, quality=QUALITY, seed=SEED), reps)
    e_page, _ = _time_host(lambda: g.compress(F.BPTC, h_page, o_page, quality=QUALITY, seed=SEED), reps)
    # sample: 192 blocks across an alpha tile boundary, keyed RNG streams + solid count make a range well defined
    blocks = h_pin.reshape(size // 4, 4, size // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(nblk, 16, 4)
    solid = (blocks == blocks[:, :1]).all((1, 2))
    first = int(np.flatnonzero(solid)[solid.sum() // 2]) - 100
    want, _ = orc.compress("BPTC", h_pin, quality=QUALITY, first_block=first, num_blocks=192, rng_mode=1, seed=SEED,
                           wm_base=int(solid[:first].sum()))
    exact = bool((want.reshape(-1, 16)[first:first + 192] == h_out.reshape(-1, 16)[first:first + 192]).all()
                 and (o_page == h_out).all())
    mp = size * size / 1e6
    out["config2_bptc_q50_2048"] = {
        "workload": "BPTC -q 50, synthetic 2048x2048 RGBA, 1 GPU", "kernel_ms": k_min, "kernel_mpix_s": mp / k_min * 1e3,
        "e2e_pinned_ms": e_pin, "e2e_pinned_mpix_s": mp / e_pin * 1e3, "e2e_pageable_ms": e_page,
        "e2e_pageable_mpix_s": mp / e_page * 1e3, "bit_exact_vs_oracle_sample": exact,
        "sample": f"blocks [{first}, {first + 192}) vs the oracle on keyed RNG streams; pageable == pinned bytes"}
    del d_in, d_out, h_pin, h_out, h_page, o_page

    # ---- config 4: DXT1 and DXT5, batch of 256 textures of 1024^2
    ntex, size = 256, 1024
    texs = torch.empty((ntex, size, size, 4), dtype=torch.uint8, device=dev)
    for k in range(ntex):
        texs[k] = synth_rgba_torch(size, size, k + 1, device=dev)
    h_pin = _pin(texs).numpy()
    h_page = np.array(h_pin)
    nblk = (size // 4) ** 2
    for name, fmt, bsz in (("dxt1", F.DXT1, 8), ("dxt5", F.DXT5, 16)):
        d_out = torch.zeros((ntex, nblk * bsz), dtype=torch.uint8, device=dev)

        def run():
            for k in range(ntex):
                g.compress_device(fmt, texs[k], d_out[k], width=size, height=size)
        run(); torch.cuda.synchronize()
        k_min, _ = _time_device(run, reps, flush)
        o_pin = _pin(d_out).numpy()
        # the same bytes as ONE launch: the batch's textures are contiguous here, and 256 stacked 1024-row
        # textures are one 1024 x 262144 image with the same block order (256 launches of 12 us each are
        # launch-bound; this is the kernel's streaming rate)
        d_one = torch.zeros_like(d_out)
        run1 = lambda: g.compress_device(fmt, texs, d_one, width=size, height=ntex * size)
        run1(); torch.cuda.synchronize()
        k1_min, _ = _time_device(run1, reps, flush)
        stacked_same = bool((d_one == d_out).all().item())
        del d_one
        o_page = np.empty_like(o_pin)
        ims_pin, ims_page = [h_pin[k] for k in range(ntex)], [h_page[k] for k in range(ntex)]
        outs_pin, outs_page = [o_pin[k] for k in range(ntex)], [o_page[k] for k in range(ntex)]
        g.compress_batch(fmt, ims_pin, outs=outs_pin)
        g.compress_batch(fmt, ims_page, outs=outs_page)
        e_pin, _ = _time_host(lambda: g.compress_batch(fmt, ims_pin, outs=outs_pin), reps)
        e_page, _ = _time_host(lambda: g.compress_batch(fmt, ims_page, outs=outs_page), reps)
        exact = bool((o_page == o_pin).all())
        for k in (0, 77, 255):
            want, _ = orc.compress(name.upper(), h_pin[k])
            exact = exact and bool((want == o_pin[k]).all())
        algo = ntex * nblk * (64 + bsz)
        in_out = ntex * (size * size * 4 + nblk * bsz)
        out[f"config4_{name}_batch256x1024"] = {
            "workload": f"{name.upper()}, one batch submission of 256 synthetic 1024x1024 textures (seeds 1..256), 1 GPU",
            "kernel_ms": k_min, "kernel_gpix_s": ntex * size * size / k_min / 1e6,
            "kernel_hbm_gb_s": algo / k_min / 1e6, "kernel_hbm_frac": algo / k_min / 1e6 / hbm_peak,
            "kernel_launches": ntex,
            "stacked_single_launch": {"kernel_ms": k1_min, "kernel_gpix_s": ntex * size * size / k1_min / 1e6,
                                      "kernel_hbm_gb_s": algo / k1_min / 1e6,
                                      "kernel_hbm_frac": algo / k1_min / 1e6 / hbm_peak,
                                      "bytes_equal_batch": stacked_same},
            "e2e_pinned_ms": e_pin, "e2e_pinned_pcie_gb_s": in_out / e_pin / 1e6,
            "e2e_pageable_ms": e_page, "e2e_pageable_pcie_gb_s": in_out / e_page / 1e6,
            "bit_exact_vs_oracle_sample": exact,
            "sample": "textures 0, 77 and 255 whole vs the oracle; pageable batch == pinned batch bytes"}
        del d_out, o_pin, o_page
    del texs, h_pin, h_page

    # ---- config 5: ETC1, 4096^2 RGB (A = 255)
    size = 4096
    d_in = synth_rgba_torch(size, size, SEED, opaque=True, device=dev)
    nblk = (size // 4) ** 2
    h_pin = _pin(d_in).numpy()
    h_page = np.array(h_pin)
    for q, qname in ((0, "low"), (2, "high")):  # (low: the dynamically scheduled kernel; high: the quad kernel)
        d_out = torch.zeros(nblk * 8, dtype=torch.uint8, device=dev)
        run = lambda: g.compress_device(F.ETC1, d_in, d_out, width=size, height=size, etc1_quality=q)
        try:
            run(); torch.cuda.synchronize()
        except Exception as e:  # quality level not built
            out[f"config5_etc1_{qname}_4096"] = {"unavailable": str(e)}
            continue
        k_min, _ = _time_device(run, reps, flush)
        h_out = _pin(d_out).numpy()
        o_page = np.empty_like(h_out)
        g.compress(F.ETC1, h_pin, h_out, etc1_quality=q)
        g.compress(F.ETC1, h_page, o_page, etc1_quality=q)
        e_pin, _ = _time_host(lambda: g.compress(F.ETC1, h_pin, h_out, etc1_quality=q), reps)
        e_page, _ = _time_host(lambda: g.compress(F.ETC1, h_page, o_page, etc1_quality=q), reps)
        n_s = 4096 if q == 0 else 512
        first = 517 * (size // 4) + 300
        want, _ = orc.compress("ETC1", h_pin, first_block=first, num_blocks=n_s, etc1_quality=q)
        exact = bool((want.reshape(-1, 8)[first:first + n_s] == h_out.reshape(-1, 8)[first:first + n_s]).all()
                     and (o_page == h_out).all())
        algo = nblk * 72
        out[f"config5_etc1_{qname}_4096"] = {
            "workload": f"ETC1 (rg_etc1 {qname} quality), synthetic 4096x4096 RGB (A = 255), 1 GPU",
            "kernel_ms": k_min, "kernel_gpix_s": size * size / k_min / 1e6, "kernel_hbm_gb_s": algo / k_min / 1e6,
            "kernel_hbm_frac": algo / k_min / 1e6 / hbm_peak, "e2e_pinned_ms": e_pin,
            "e2e_pinned_gpix_s": size * size / e_pin / 1e6, "e2e_pageable_ms": e_page,
            "e2e_pageable_gpix_s": size * size / e_page / 1e6, "bit_exact_vs_oracle_sample": exact,
            "sample": f"blocks [{first}, {first + n_s}) vs the oracle; pageable == pinned bytes"}
        del d_out
    del d_in

    # ---- PVRTC 4bpp, 1024^2 (image-level encoder; the whole texture is compared with the
    # compiled reference, which is also the CPU time next to it: PVRTCC::Compress is single-threaded)
    try:
        from _checkers import Reference
        size = 1024
        d_in = synth_rgba_torch(size, size, SEED, device=dev)
        nblk = (size // 4) ** 2
        d_out = torch.zeros(nblk * 8, dtype=torch.uint8, device=dev)
        run = lambda: g.compress_device(F.PVRTC4, d_in, d_out, width=size, height=size)
        run(); torch.cuda.synchronize()
        k_min, _ = _time_device(run, reps, flush)
        h_pin = _pin(d_in).numpy()
        h_out = _pin(d_out).numpy()
        g.compress(F.PVRTC4, h_pin, h_out)
        e_pin, _ = _time_host(lambda: g.compress(F.PVRTC4, h_pin, h_out), reps)
        entry = {"workload": "PVRTC 4bpp, synthetic 1024x1024 RGBA, 1 GPU", "kernel_ms": k_min,
                 "kernel_mpix_s": size * size / k_min / 1e3, "e2e_pinned_ms": e_pin,
                 "e2e_pinned_mpix_s": size * size / e_pin / 1e3,
                 "note": "the reference's backward labelling scan is one serial chain per texture; here rows stay "
                         "sequential and a row is a scan (pvrtc.cu): the encoder scales over the textures of a batch"}
        if Reference.available():
            want, ref_ms = Reference().compress("PVRTC4", h_pin, seed=None)
            entry.update({"bit_exact_vs_reference": bool((want == h_out).all()), "reference_cpu_ms": ref_ms,
                          "reference_cpu_mpix_s": size * size / ref_ms / 1e3, "reference_threads": 1})
        # a batch: the textures are encoded side by side (grid.y = texture of a run of equal sizes)
        nb_, bs_ = 64, 512
        bimgs = [_pin(synth_rgba_torch(bs_, bs_, k + 1, device=dev)).numpy() for k in range(nb_)]
        g.compress_batch(F.PVRTC4, bimgs[:2])
        b_ms, _ = _time_host(lambda: g.compress_batch(F.PVRTC4, bimgs), reps)
        entry["batch_64x512"] = {"e2e_pinned_ms": b_ms, "e2e_mpix_s": nb_ * bs_ * bs_ / b_ms / 1e3}
        if Reference.available():
            outs, _ = g.compress_batch(F.PVRTC4, bimgs)
            wantb, refb_ms = Reference().compress("PVRTC4", bimgs[-1], seed=None)
            entry["batch_64x512"].update({"bit_exact_vs_reference_last": bool((outs[-1] == wantb).all()),
                                          "reference_cpu_ms_per_texture": refb_ms,
                                          "reference_cpu_mpix_s_one_core": bs_ * bs_ / refb_ms / 1e3})
        out["n4_pvrtc4_1024"] = entry
        del d_in, d_out
    except Exception as e:  # noqa: BLE001
        out["n4_pvrtc4_1024"] = {"unavailable": str(e)[:200]}
    return out


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
    stage_tot = []
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        step_device()
        ev[k][1].record()
        stage_tot.append(g.bc7_stage_ms(enable=True, read=True))  # syncs on this step's stage events
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    device_launches = launches
    g.bc7_stage_ms(enable=False, read=False)

    # ---- end to end through the C ABI host entry, H2D + D2H timed: pinned, then pageable buffers
    h_in_np, h_out_np = h_in.numpy(), h_out.numpy()

    def e2e_run(src, dst, steps, warmup):
        n_l = 0
        for _ in range(warmup):
            g.compress(F.BPTC, src, dst, quality=QUALITY, seed=SEED)
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        per = []
        for _ in range(steps):
            ts = time.perf_counter()
            _, tm = g.compress(F.BPTC, src, dst, quality=QUALITY, seed=SEED)
            per.append((time.perf_counter() - ts) * 1e3)
            h2d, d2h = tm["h2d_bytes"], tm["d2h_bytes"]
            n_l += tm["kernel_launches"]
        barrier()
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
            tb = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
            dist.all_reduce(tb)
            h2d, d2h = int(tb[0]), int(tb[1])
        return float(secs.item()), h2d, d2h, per, n_l

    e2e_s, h2d, d2h, e2e_steps_ms, e2e_launches = e2e_run(h_in_np, h_out_np, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    page_steps = max(1, min(args.steps, 5))
    page_in, page_out = np.array(h_in_np), np.empty_like(h_out_np)
    page_s, _, _, _, page_launches = e2e_run(page_in, page_out, page_steps, 2)
    pageable_same = bool((page_out == h_out_np).all())
    del page_in, page_out

    # ---- N > 1: the in-process N-GPU API path -- ONE fastc_gpu_compress(num_gpus = N) call by rank 0 on
    # the whole texture in one caller buffer, the other ranks' processes idle on a host-side wait
    single = None
    if multi:
        store = dist.distributed_c10d._get_default_store()
        barrier()
        if rank == 0:
            gathered = torch.cat(gather_list).cpu().numpy()
            full_dev = synth_rgba_torch(WIDTH, HEIGHT, SEED, device=dev)
            full_pin = _pin(full_dev)
            del full_dev
            out_pin = torch.empty((WIDTH // 4) * (HEIGHT // 4) * 16, dtype=torch.uint8, pin_memory=True)
            single = {}
            single_launches = 0
            for kind, src, dst in (("pinned", full_pin.numpy(), out_pin.numpy()),
                                   ("pageable", np.array(full_pin.numpy()), np.empty(out_pin.numel(), dtype=np.uint8))):
                for _ in range(3):
                    g.compress(F.BPTC, src, dst, quality=QUALITY, seed=SEED, num_gpus=world)
                per = []
                for _ in range(page_steps):
                    ts = time.perf_counter()
                    _, tm = g.compress(F.BPTC, src, dst, quality=QUALITY, seed=SEED, num_gpus=world)
                    per.append((time.perf_counter() - ts) * 1e3)
                    single_launches += tm["kernel_launches"]
                ms = sum(per) / len(per)
                single[kind] = {"value": WIDTH * HEIGHT / 1e6 / (ms / 1e3), "unit": "Mpix/s", "ms_per_step": ms,
                                "ms_steps": per, "host_memory": kind,
                                "bytes_equal_nccl_gathered_slabs": bool((dst == gathered).all())}
            single["path"] = (f"one fastc_gpu_compress(num_gpus={world}) call by rank 0: whole texture in one caller "
                              f"buffer, block-row slabs over {world} GPUs inside the library, each GPU copying into "
                              f"its slice of the caller's output")
            single["steps"] = page_steps
            single["gpu_launches"] = single_launches
            store.set("single_call_done", "1")
        else:
            store.wait(["single_call_done"])
        barrier()

    if rank != 0:
        return

    mpix = WIDTH * HEIGHT / 1e6
    ms_per_step = total_ms / args.steps
    value = mpix / (ms_per_step / 1e3)
    e2e_value = mpix / (e2e_s / args.steps)

    # ---- roofline: ALU instruction issue, from measured thread-instruction counts
    hbm_peak, sm_max_mhz, peak_src = load_peaks()
    stages = {k: sum(s[k] for s in stage_tot) / len(stage_tot) for k in stage_tot[0]}
    k_ms = stages["anneal"]
    sm_mhz = (clocks or {}).get("sm_mhz") or sm_max_mhz
    issue_peak = SM_COUNT * LANES_PER_SM * sm_mhz * 1e6  # thread-instructions / s at the clock under load
    frac_blocks = nblk / ((WIDTH // 4) * (HEIGHT // 4))
    roofline = {"bound": "alu_issue", "unit": "thread-instr/s", "peak": issue_peak,
                "peak_source": f"{SM_COUNT} SMs x {LANES_PER_SM} lanes x {sm_mhz:.0f} MHz (SM clock sampled under load; "
                               f"max {sm_max_mhz:.0f} MHz, {peak_src})",
                "kernel": "bc7_anneal", "kernel_ms": k_ms, "kernel_share_of_step": k_ms / stages["total"],
                "stages_ms": stages}
    if INSTR_COUNTS.exists() and WIDTH == 8192:
        ic = json.loads(INSTR_COUNTS.read_text())
        per_kernel = ic["thread_inst_per_launch_8192"]
        step_inst = sum(per_kernel.values()) * frac_blocks
        stage_of = {"bc7_classify": "classify", "bc7_wm_scan": "classify", "bc7_select": "select", "bc7_setup": "setup",
                    "bc7_bin_offsets": "setup", "bc7_scatter": "setup", "bc7_anneal": "anneal", "bc7_pack": "pack"}
        by_stage = {}
        for kname, cnt in per_kernel.items():
            st = stage_of.get(kname)
            if st:
                by_stage[st] = by_stage.get(st, 0.0) + cnt * frac_blocks
        roofline.update({
            "achieved": step_inst / (stages["total"] / 1e3), "frac": step_inst / (stages["total"] / 1e3) / issue_peak,
            "achieved_definition": "thread-instructions executed by all BC7 kernels of one step (ncu "
                                   "smsp__thread_inst_executed.sum per launch at 8192^2, scaled to this rank's blocks) / "
                                   "the live CUDA-event time of the step's kernels",
            "per_stage_frac": {st: by_stage[st] / (stages[st] / 1e3) / issue_peak for st in by_stage if stages.get(st)},
            "thread_inst_per_step": step_inst, "counts_file": str(INSTR_COUNTS.relative_to(ROOT)),
            "counts_kernel_sha": ic.get("kernel_source_sha"), "counts_stale": ic.get("kernel_source_sha") != kernel_source_sha(),
            "warp_inst_per_step": sum(ic.get("warp_inst_per_launch_8192", {}).values()) * frac_blocks or None,
        })
        traffic = ic.get("dram_bytes_per_step_8192")
        if ic.get("work_counters"):
            roofline["work_counters"] = ic["work_counters"]
    else:
        roofline.update({"achieved": None, "frac": None, "note": "no committed instruction counts for this size"})
        traffic = None
    algo = nblk * ALGO_BYTES_PER_BLOCK
    roofline["traffic"] = traffic * frac_blocks if traffic else None
    roofline["traffic_unit"] = ("bytes per step, all BC7 kernels (ncu dram__bytes_read.sum + dram__bytes_write.sum), "
                                "scaled to this rank's blocks")
    roofline["hbm"] = {"achieved": algo / (stages["total"] / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                       "frac": algo / (stages["total"] / 1e3) / 1e9 / hbm_peak, "algorithmic_bytes": algo,
                       "note": "80 B per block over the step's kernel time: BC7 is nowhere near the HBM roofline"}

    # ---- CPU baseline: the reference itself on this box's host cores, bounded slab (N = 1 only)
    cpu_baseline = None
    quality_check = None
    others = None
    if world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, str(ROOT / "tests"))
        from _checkers import Reference, Oracle
        cores = os.cpu_count() or 1
        rows_s = REF_SAMPLE_ROWS
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
    if world == 1 and not args.no_other_configs:
        del d_in, d_out, h_in, h_out
        try:
            others = other_configs(g, dev, flush, hbm_peak)
        except Exception as e:  # the headline line must survive a failure of the side measurements
            others = {"error": f"{type(e).__name__}: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "quality": QUALITY, "width": WIDTH, "height": HEIGHT},
        "parallelism": f"block-row slabs x{world}" + (", compressed output gathered to rank 0 (NCCL send/recv)" if multi else ""),
        "l2": "256 MiB flush between timed steps (input slab is also > L2 at N=1)",
        "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / args.steps * 1e3, "rank0_ms_steps": e2e_steps_ms, "host_memory": "pinned",
                "path": "fastc_gpu_compress (C ABI, pinned host in/out), one process per GPU on its slab",
                "pageable": {"value": mpix / (page_s / page_steps), "unit": "Mpix/s",
                             "ms_per_step": page_s / page_steps * 1e3, "steps": page_steps,
                             "host_memory": "pageable (what FasTC passes: new[])", "bytes_equal_pinned": pageable_same}},
        "gpu_launches": device_launches + e2e_launches + page_launches + ((single or {}).get("gpu_launches") or 0),
        "roofline": roofline,
        "wall_s_timed_region": t_wall,
    }
    if single:
        line["e2e_single_call"] = single
    if cpu_baseline:
        line["cpu_baseline"] = cpu_baseline
    if quality_check:
        line["psnr_vs_reference"] = quality_check
    if others:
        line["other_configs"] = others
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
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
