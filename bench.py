#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense-matching hot path (BASELINE.json metric: MDE/s, ms/frame).

Workload (config.workload): BASELINE config 5 -- a batch of 1920x1080 synthetic stereo pairs, 256 disparities,
guided-filter ASW (GuidedF_2, r = 9, eps = 1e-4) for the left AND right view + LR check + weighted-median
refinement; pairs are sharded over ranks with no collective (weak scaling: every rank owns --pairs pairs).
One "step" = one pass of that pipeline over the rank's batch.

  value : whole-job MDE/s with the inputs already resident in HBM (CUDA events on the library's stream,
          barrier + synchronize on both sides, max over ranks)
  e2e   : the same metric through the C-ABI batch calls with HOST (pinned) buffers: per step every pair is
          copied host->device, processed, and its refined map copied device->host inside the timed region
  roofline / cpu_baseline : see DESIGN.md section "Measurement"

Scaling: the named config is ONE batch of 64 pairs, so by default the 64 pairs are divided over the ranks
("scaling": "strong", --pairs-total 64); `--pairs P` instead gives every rank P pairs ("weak").  At N > 1 the
strong line also carries the weak figure ("weak": {...}) from one extra pass.

`--impl reference` times the reference's CPU behaviour (oracle/_ref for the methods it covers; GuidedF_2 needs
OpenCV's filters, so this path runs the oracle port, all host threads) on the SAME config: each step is one
full-size pair of the batch (1920x1080, D = 256, both views + LR + refine), a bounded sample of the 64-pair step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, D = 1080, 1920, 256
WIN, EPS = 9, 1e-4
VIEWS = 2
# algorithmic bytes per disparity evaluation (SURVEY 8d, 3-pass model = 52 B/DE: write p 4 + read L,R 6; read p 4 +
# I 3, write a,b 16; read a,b 16 + I 3).  The streaming kernel gfs_filter runs all three model passes on chip, so
# it is charged the whole 52 B/DE; the older tiled pair splits it 33 (gf_ab) + 19 (gf_q).
ALG_BYTES = {"gfs_filter": 52.0, "gf_ab": 33.0, "gf_q": 19.0}
# FP32-issue view of the same kernel (~80 lane-instructions per DE for the guided filter)
ALG_LANE_INSTR = {"gfs_filter": 80.0}
WORKLOAD = ("cfg5: batch of 1920x1080 synthetic pairs, 256 disparities, GuidedF_2 (r=9, eps=1e-4) left+right view "
            "+ LR check + weighted-median refine, pair-sharded")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        v = float(json.load(open(p))["hbm_gbs"])
        return v, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(sample_hw=(1080, 1920), sample_d=256, threads=None):
    """oracle (CPU port of the reference) on a bounded sample of the workload: one pair, both views + LR + refine"""
    from aswstereomatch_b200.synth import make_pair
    from oracle import orc
    if threads:
        orc.set_num_threads(threads)
    h, w = sample_hw
    L, R, _ = make_pair(h, w, sample_d, 1000)
    t0 = time.perf_counter()
    out, _ = orc.guidedf2_lr_refine(L, R, EPS, WIN, 0, sample_d)
    dt = time.perf_counter() - t0
    mde = h * w * sample_d * VIEWS / 1e6
    return {"value": mde / dt, "unit": "MDE/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"1 pair {w}x{h}, D={sample_d}, r={WIN}, eps={EPS}, 2 views + LR + refine ({dt:.1f} s; "
                      f"oracle/asw_oracle.c, OpenMP over slices/rows)"}, out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation (oracle port; the reference itself needs OpenCV
    4.1.0 C++ and cannot be built here) on a bounded sample per step."""
    if rank != 0:
        return
    from aswstereomatch_b200.synth import make_pair
    from oracle import orc
    orc.build()
    orc.set_num_threads(os.cpu_count() or 1)       # all host threads (torchrun presets OMP_NUM_THREADS=1)
    h, w, d = H, W, D                              # same config as the b200 arm: one full-size pair per step
    L, R, _ = make_pair(h, w, d, 1000)
    for _ in range(args.warmup):
        orc.guidedf2_lr_refine(L, R, EPS, WIN, 0, d)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.guidedf2_lr_refine(L, R, EPS, WIN, 0, d)
    dt = time.perf_counter() - t0
    mde_step = h * w * d * VIEWS / 1e6
    val = mde_step * args.steps / dt
    sample = (f"per step: 1 pair {w}x{h}, D={d}, r={WIN}, eps={EPS}, 2 views + LR + refine (1 of the 64 pairs of a "
              f"cfg5 step; same image size and disparity range as the b200 arm)")
    line = {"impl": "reference", "metric": "MDE/s", "value": val, "unit": "MDE/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample, "image": [W, H], "disparities": D, "views": VIEWS},
            "cpu_baseline": {"value": val, "unit": "MDE/s", "cores": orc.num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "MDE/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs-total", type=int, default=64, help="pairs per step over ALL ranks (cfg5: 64; strong scaling)")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (weak scaling; overrides --pairs-total)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 1-4 section")
    ap.add_argument("--no-split", action="store_true", help="skip the disparity-split section")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import aswstereomatch_b200 as asw
    from aswstereomatch_b200 import sharding
    from aswstereomatch_b200.synth import make_batch
    ctx = asw.Context(local_rank)          # raises without the CUDA library / a device: no CPU fallback
    weak = args.pairs > 0
    PW = args.pairs if weak else 64                                           # pairs per GPU of the weak leg
    mine = list(range(PW)) if weak else sharding.shard_pairs(args.pairs_total, rank, world)   # pair i -> rank i % N
    P = len(mine)
    total_pairs = PW * world if weak else args.pairs_total
    n_buf = max(P, PW) if (world > 1 and not weak) else P                     # the extra weak pass needs PW pairs
    # global pair i has seed 1000 + i (weak: rank r owns 1000 + 10000 r + i); pair 0 of rank 0 is always seed 1000
    Ls, Rs = make_batch(n_buf, H, W, D, seed0=1000 + 10000 * rank, distinct=min(n_buf, 4))
    # pinned host buffers (inputs and results) for the end-to-end leg
    hL = asw.pinned_empty((n_buf, H, W, 3), np.uint8)
    hR = asw.pinned_empty((n_buf, H, W, 3), np.uint8)
    hD = asw.pinned_empty((n_buf, H, W), np.float32)
    for i in range(n_buf):
        hL[i] = Ls[i]; hR[i] = Rs[i]
    del Ls, Rs
    batch = asw.Batch(ctx, n_buf, H, W)

    def barrier():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    def step_resident(n=P):
        batch.run_guidedf2_lr_refine(EPS, WIN, 0, D, n_pairs=n)

    def step_e2e(n=P):
        for i in range(n):
            batch.upload(i, hL[i], hR[i])
        batch.run_guidedf2_lr_refine(EPS, WIN, 0, D, n_pairs=n)
        for i in range(n):
            batch.download(i, hD[i], sync=False)

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        for _ in range(steps):
            fn()
        ms = ctx.timer_stop()          # records the stop event and synchronises the stream
        barrier()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # inputs resident in HBM before the timed region of the `value` leg
    for i in range(n_buf):
        batch.upload(i, hL[i], hR[i])
    for _ in range(args.warmup):
        step_resident()
    ctx.sync()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    n0 = ctx.launch_count()
    ms_res = timed(step_resident, args.steps)
    launches = ctx.launch_count() - n0
    clk = clocks.stop() if rank == 0 else None

    # per-kernel launch durations, measured live with CUDA events on the launching stream (one extra step)
    ctx.profile_enable(True); ctx.profile_reset()
    step_resident(); ctx.sync()
    prof = ctx.profile()
    ctx.profile_enable(False)

    # end-to-end leg (host buffers; H2D + D2H inside the timed region)
    step_e2e(); ctx.sync()
    ms_e2e = timed(step_e2e, args.steps)
    map0 = np.array(hD[0])                  # refined map of pair 0 (rank 0: seed 1000) from the last e2e step

    weak_leg = None
    if world > 1 and not weak:
        # the same pipeline with 64 pairs on EVERY rank (weak scaling), one warm-up + the timed steps
        step_resident(PW); ctx.sync()
        ms_w = timed(lambda: step_resident(PW), args.steps)
        weak_leg = {"pairs_per_gpu": PW, "ms_per_step": ms_w / args.steps,
                    "value": H * W * D * VIEWS * PW * world / 1e6 / (ms_w / args.steps) * 1e3, "unit": "MDE/s"}

    split = None
    if not args.no_split:
        split = bench_split(ctx, asw, dist, torch, rank, world, local_rank, args)
    cfgs = methods = None
    if rank == 0 and world == 1 and not args.no_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        cfgs = bench_configs.run_configs(ctx, reps=3, warmup=3, brief=True)
        methods = bench_configs.run_driver_methods(ctx)

    # sanity: the last e2e result is a plausible disparity map (guards against timing a no-op)
    ok = bool(np.isfinite(map0).all() and map0.max() <= D - 1 and map0.std() > 0)

    if rank == 0:
        mde_step = H * W * D * VIEWS * total_pairs / 1e6
        value = mde_step / (ms_res / args.steps) * 1e3
        e2e = mde_step / (ms_e2e / args.steps) * 1e3
        peak, peak_src = measured_peak()
        dom = max(prof.items(), key=lambda kv: kv[1][0])
        name, (tot_ms, n_l) = dom
        de_per_launch = H * W * D * VIEWS * P / n_l            # one launch of the dominant kernel covers one view of one pair
        alg = ALG_BYTES.get(name)
        achieved = alg * de_per_launch / (tot_ms / n_l * 1e-3) / 1e9 if alg else None
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr[name]["dram_bytes_per_de"] * de_per_launch
        except Exception:
            pass
        total_prof = sum(v[0] for v in prof.values())
        line = {
            "metric": "MDE/s", "value": value, "unit": "MDE/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "ms_per_frame": ms_res / args.steps / P, "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_total": total_pairs, "pairs_per_gpu": P, "image": [W, H],
                       "disparities": D, "views": VIEWS,
                       "l2": "inputs larger than L2: every step streams 12.4 MB of images and, per view, a 2.1 GB "
                             "filtered-cost volume per pair through the 126 MB L2", "result_check": ok},
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "MDE/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(hL[:P].nbytes + hR[:P].nbytes), "d2h_bytes_per_step": int(hD[:P].nbytes)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "avg_launch_ms": tot_ms / n_l, "share_of_step": tot_ms / total_prof,
                         "alg_bytes_per_de": alg, "de_per_launch": de_per_launch,
                         "frac_of_traffic": (traffic / (tot_ms / n_l * 1e-3) / 1e9 / peak) if traffic else None},
            "fp32_issue": None,
            "kernels_ms_per_step": {k: round(v[0], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:6]},
        }
        if weak_leg:
            line["weak"] = weak_leg
        if split:
            line["split"] = split
        if cfgs:
            line["configs"] = cfgs
        if methods:
            # every dispatcher value of the hot path with the driver's own literals (640x360, win 15, D 64), kernel-only ms
            line["driver_methods"] = methods
        if name in ALG_LANE_INSTR:
            # secondary view: the kernel moves far fewer HBM bytes than the 3-pass model, its own bound is the FP32
            # issue rate (148 SMs x 128 lanes x SM clock)
            sm_clk = (clk or {}).get("sm_mhz") or 1965.0
            peak_li = 148 * 128 * sm_clk * 1e6
            ach_li = ALG_LANE_INSTR[name] * de_per_launch / (tot_ms / n_l * 1e-3)
            line["fp32_issue"] = {"alg_lane_instr_per_de": ALG_LANE_INSTR[name], "achieved": ach_li, "peak": peak_li,
                                  "unit": "lane-instr/s", "frac": ach_li / peak_li}
        if world == 1 and not args.no_cpu_baseline:
            # rank 0 at N = 1 only: the oracle on the pair the GPU holds at index 0 (seed 1000) -- timed as the CPU
            # baseline AND compared with the GPU's refined map of that pair (parity of the headline run itself)
            try:
                from oracle import orc
                orc.build()
                line["cpu_baseline"], ref_map = cpu_baseline(threads=os.cpu_count() or 1)
                n_diff = int((ref_map != map0).sum())
                agree = 1.0 - n_diff / ref_map.size
                line["parity"] = {"what": "refined map of pair 0 (e2e leg, through the C-ABI) vs the oracle on the same pair",
                                  "agree": agree, "n_diff": n_diff, "pixels": int(ref_map.size)}
                line["config"]["result_check"] = bool(ok and agree >= 0.995)
            except Exception as e:          # the baseline is reported, never required for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "MDE/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
        if not line["config"]["result_check"]:
            sys.exit("bench: result check failed (GPU map implausible or below 99.5 % agreement with the oracle)")
    batch.close()
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def bench_split(ctx, asw, dist, torch, rank, world, local_rank, args):
    """disparity-range split of ONE 1080p x 256 GuidedF_2 pair (the dispatcher call, left view) over the ranks
    (SURVEY 8e-2): every rank evaluates D / N slices, the u64 keys are MIN-all-reduced in device memory over NCCL, every
    rank converts them to the map.  Timed through the C ABI with host buffers (upload of the pair, kernels, collective,
    download of the map), max over ranks; at N = 1 the same call unsplit."""
    from aswstereomatch_b200 import sharding
    from aswstereomatch_b200.synth import make_pair
    L0, R0, _ = make_pair(H, W, D, 1000)
    L = asw.pinned_empty(L0.shape, np.uint8); R = asw.pinned_empty(R0.shape, np.uint8)     # pinned, like the e2e leg's buffers
    L[...] = L0; R[...] = R0
    alg = asw.ADAPTIVE_WEIGHT_GUIDED_FILTER_2
    dev = torch.device("cuda", local_rank) if world > 1 else None

    def one():
        if world == 1:
            return ctx.stereoMatching(L, R, 0, alg, WIN, 0, D, strict=True)
        return sharding.split_stereo_matching(ctx, L, R, alg, 0, WIN, 0, D, rank, world, device=dev)

    def sync_all():
        ctx.sync()
        if dist is not None:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()

    out = None
    for _ in range(max(args.warmup, 2)):
        out = one()
    times, ar = [], []
    for _ in range(max(args.steps, 3)):
        sync_all()
        t0 = time.perf_counter()
        out = one()
        ctx.sync()
        dt = (time.perf_counter() - t0) * 1e3
        if dist is not None:
            t = torch.tensor([dt, sharding.split_stereo_matching.last_allreduce_ms or 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt, a = float(t[0].item()), float(t[1].item())
            ar.append(a)
        times.append(dt)
    res = {"what": f"one {W}x{H} pair, D={D}, GuidedF_2 (dispatcher, left view) disparity-split over {world} rank(s), "
                   "pinned host images in, host map out, max over ranks", "ms": float(np.median(times)),
           "allreduce_ms": float(np.median(ar)) if ar else 0.0, "slices_per_rank": -(-D // world),
           "mde_s": H * W * D / 1e6 / (float(np.median(times)) * 1e-3)}
    if world > 1:
        # every rank holds the same merged map; rank 0 also checks it against its own unsplit run
        if rank == 0:
            res["bit_identical_to_unsplit"] = bool(np.array_equal(out, ctx.stereoMatching(L, R, 0, alg, WIN, 0, D, strict=True)))
        sync_all()
    return res


if __name__ == "__main__":
    main()
