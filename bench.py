#!/usr/bin/env python
"""Benchmark of the camera-candidate sweep (BASELINE.json metric: camera candidates scored/s at a
512^3 grid) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload, defined in bench_workload.py): the deterministic synthetic 512^3 semantic monument, all 9
parts scored (21.9 M points), one 1024x1024 ground-truth label image rendered through a hidden camera, 65 536 candidate
cameras (base + U(-1,1) * reference step sizes, seed 20240607).  One "step" = every rank scores its block of
`--cands-per-step` candidates and the ranks agree on the best (score, index) with one 16-byte all-gather.  Candidates are
sharded across ranks, so per-GPU work is fixed (weak scaling).

value       : candidates/s with grid points, ground truth and candidates resident in HBM.
e2e         : the same through the public API with HOST candidate arrays in and host scores/counts out.
roofline    : the splat kernel, algorithmic bytes = cameras/launch * (G*1 B + 9 B*H*W)  (SURVEY 8d),
              duration from CUDA events recorded around every splat launch inside the timed region.
parity      : counts and scores of candidates sampled from every z-buffer batch and both double-buffer slots of one timed
              step, re-computed by the reference's own functions on the host (baseline/_ref; the NumPy port under
              oracle/ only if that is absent); at N > 1 every rank checks candidates of its own block and the NCCL
              winner is compared with the arg-max over an all-gather of the full score vectors.  Any mismatch: exit 1.
sweep_total : strong scaling of the literal configuration -- wall time from a device-resident RGB grid + the host
              (65536, 9) candidate array to the global best on every rank: scorer set-up (labels, compaction, segments,
              ground-truth labels) + this rank's contiguous shard in blocks + one all-gather.
cpu_baseline: the reference path on this box's host cores (same sampled candidates as the parity gate).
"""
import argparse
import json
import os
import sys

# the reference arm's worker processes run one BLAS thread each: set before NumPy loads OpenBLAS
if "reference" in sys.argv[1:]:
    for _k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = "1"

import importlib
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import bench_workload as bw                                  # noqa: E402  (NumPy only; shared by both arms)

PKG = "part-based-3d-reconstruction_b200"
METRIC = "camera candidates scored/s at 512^3 grid"
UNIT = "candidates/s"
TOTAL_CANDIDATES = bw.TOTAL_CANDIDATES
HIDDEN_DELTA = bw.HIDDEN_DELTA
PARITY_INDICES = [0, 127, 128, 300, 1023, 1024, 1500, 2047]   # of one step's block: every 128-camera batch boundary, both slots


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--mask", type=int, default=1024)
    ap.add_argument("--cands-per-step", type=int, default=2048, help="candidates per rank per step")
    ap.add_argument("--parts", default="all", choices=["all", "minarets"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="candidates in the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU sample (this also skips the parity gate)")
    ap.add_argument("--no-carve", action="store_true", help="skip the secondary carving measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements of configs 2, 3 and 5")
    ap.add_argument("--no-sweep-total", action="store_true", help="skip the whole-sweep strong-scaling measurement")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8 or not (t0 - 0.2 <= t <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# parity gate helpers
# ------------------------------------------------------------------------------------------------
def compare_with_cpu(cpu, cand_rows, gpu_counts, gpu_scores, processes):
    """Score `cand_rows` with the CPU arm and compare integer counts (exactly) and scores (1e-5 relative; in fact they
    are identical).  Returns (seconds, n_checked, list of mismatch descriptions)."""
    dt, counts, scores = cpu.run(cand_rows, processes=processes)
    bad = []
    for k in range(len(cand_rows)):
        if not np.array_equal(counts[k], gpu_counts[k]):
            bad.append(f"counts[{k}]: gpu {gpu_counts[k].tolist()} vs cpu {counts[k].tolist()}")
        elif not abs(scores[k] - gpu_scores[k]) <= 1e-5 * max(abs(scores[k]), 1e-300):
            bad.append(f"score[{k}]: gpu {gpu_scores[k]!r} vs cpu {scores[k]!r}")
    return dt, len(cand_rows), bad, bool(np.array_equal(scores, gpu_scores))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    syn = importlib.import_module(PKG + ".synthetic")
    ce = importlib.import_module(PKG + ".utils.camera_estimation")
    cfg = importlib.import_module(PKG + ".utils.config")
    nv = importlib.import_module(PKG + ".utils._native")
    sweep_mod = importlib.import_module(PKG + ".utils.sweep")

    N, H, W, B = args.grid, args.mask, args.mask, args.cands_per_step
    parts = syn.PART_NAMES if args.parts == "all" else ["front_minarets", "back_minarets"]
    lut = torch.from_numpy(syn.label_lut()).to(dev)
    rgb = lut[syn.monument_labels(N, dev).long()]
    base = syn.base_camera(N, H, W, "front")
    hidden = base + HIDDEN_DELTA
    full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
    gt = torch.from_numpy(full.render(ce.row_to_params(hidden))).to(dev)
    del full
    torch.cuda.synchronize()
    t_setup = time.perf_counter()
    scorer = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, parts)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup                  # first construction (includes allocator warm-up)
    n_points = scorer.n_points
    P = scorer.P

    cand_all = syn.candidates(base, TOTAL_CANDIDATES)                       # (65536, 9) float64, host
    steps_total = args.warmup + args.steps

    def step_block(s):
        """Global candidate indices of step s for this rank (contiguous shard of the step's chunk)."""
        g0 = (s * B * world) % TOTAL_CANDIDATES
        idx = (g0 + rank * B + np.arange(B)) % TOTAL_CANDIDATES
        return idx

    host_blocks = [torch.from_numpy(np.ascontiguousarray(cand_all[step_block(s)])).pin_memory() for s in range(steps_total)]
    offsets = [int(step_block(s)[0]) for s in range(steps_total)]
    dev_blocks = [b.to(dev) for b in host_blocks]
    reducer = sweep_mod.BestReducer(dev, world)

    def device_step(s):
        counts, scores, best = scorer.score_device(dev_blocks[s])
        return reducer.reduce(scores, best, offsets[s]), counts, scores

    # ---- value: device-resident inputs ------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                       # before the barrier, so that no rank enters the timed region late
    for s in range(args.warmup):
        device_step(s)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = nv.launch_count
    scorer.workspace.timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for s in range(args.warmup, steps_total):
        result, last_counts, last_scores = device_step(s)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall1 = time.time()
    ms = e0.elapsed_time(e1)
    splat_ms_v, splat_launches_v = scorer.workspace.timing_read()
    scorer.workspace.timing(False)
    launches = nv.launch_count - launches0 + args.steps * reducer.launches_per_reduce
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = B * world * args.steps / (ms * 1e-3)
    # what the last timed step returned (kept for the parity gate before anything overwrites the output tensors)
    s_last = steps_total - 1
    gate_counts = last_counts.cpu().numpy()
    gate_scores = last_scores.cpu().numpy()
    gate_best = (float(result[0]), int(result[1]))

    # ---- e2e: host candidates in, host scores/counts out, through the public API ------------------
    for s in range(min(2, args.warmup)):
        scorer.score(host_blocks[s].numpy())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for s in range(args.warmup, steps_total):
        sc, cn, bi = scorer.score(host_blocks[s].numpy())
    e3.record()
    torch.cuda.synchronize()
    t2 = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / (float(t2.item()) * 1e-3)
    h2d = B * 9 * 8
    d2h = B * 8 + B * cn.shape[1] * 2 * 8 + 8
    e2e_same = bool(np.array_equal(cn, gate_counts[:, scorer._cols, :]) and np.array_equal(sc, gate_scores))

    # ---- sweep_total: the literal configuration, strong scaling incl. set-up -----------------------
    sweep_total = None
    if not args.no_sweep_total:
        sweep_total = whole_sweep(ce, cfg, sweep_mod, nv, rgb, gt, parts, cand_all, B, dev, world, rank, dist)

    # ---- parity gate --------------------------------------------------------------------------------
    parity = None
    cpu = None
    if not args.no_cpu_baseline:
        parity, cpu = parity_gate(args, scorer, gt, parts, cand_all, step_block(s_last), gate_counts, gate_scores, gate_best,
                                  offsets[s_last], B, H, W, dev, world, rank, dist, sweep_mod)

    # ---- N > 1: BASELINE.json configs[4] across the box: the 1024^3 / 2048^2 sweep, candidates sharded -------------
    config5 = None
    if world > 1 and not args.no_extra:
        try:
            del scorer
            torch.cuda.empty_cache()
            config5 = config5_sweep(dev, world, rank, dist, not args.no_cpu_baseline)
        except Exception as exc:
            config5 = {"error": repr(exc)}

    # ---- N > 1: carving of a 1024^3 grid sharded by x-slab, max over ranks --------------------------
    carve_multi = None
    if world > 1 and not args.no_carve:
        try:
            carve_multi = carve_sharded_bench(1024, dev, world, rank, dist)      # configs[4]: 1024^3 across the box
        except Exception as exc:
            carve_multi = {"error": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if parity is not None and parity.get("ok") is False:
            sys.exit(1)
        return

    # ---- roofline of the dominant kernel (splat) --------------------------------------------------
    peak, peak_src = peaks()
    G = N ** 3
    batch = -(-B * args.steps // max(1, splat_launches_v))       # cameras per splat launch, from the launch count
    alg_per_launch = batch * (G * 1 + 9 * H * W)
    avg_launch_s = (splat_ms_v / max(1, splat_launches_v)) * 1e-3
    achieved = alg_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    traffic = None
    try:        # DRAM bytes per launch of this kernel from the committed ncu --set full capture (same batch size only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_splat_traffic.json")))
        if int(tj["cameras_per_launch"]) == batch and N == 512 and H == 1024 and args.parts == "all":
            traffic = int(tj["traffic_bytes_per_launch"])
    except Exception:
        pass
    seg_on = n_points >= int(os.environ.get("P3D_SEG_MIN_POINTS", "6000000")) and os.environ.get("P3D_SPLAT_POINTS") != "1"
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic,
                "kernel": ("splat_seg_kernel<double, joint-packed> (x-run segments, FP32 filter + exact FP64 queue)" if seg_on
                           else "splat_filtered_kernel<double, joint-packed> (FP32 filter + exact FP64 queue)"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_per_launch,
                "cameras_per_launch": batch, "avg_launch_ms": round(avg_launch_s * 1e3, 4),
                "splat_share_of_step": round(splat_ms_v / ms, 4),
                "step_achieved": round(value / max(world, 1) * (G * 1 + 9 * H * W) / 1e9, 1),
                "step_frac": round(value / max(world, 1) * (G * 1 + 9 * H * W) / 1e9 / peak, 4),
                "point_candidates_per_s": round(n_points * batch / avg_launch_s, 1) if avg_launch_s > 0 else None,
                "note": "effective GB/s under the streaming model of SURVEY 8(d) (dense u8 grid once per camera + 9 B/pixel); "
                        "the kernel batches cameras per point pass and is bound on-chip (issue slots / early-out load "
                        "latency), see DESIGN.md 4.1; avg_launch_ms is measured inside the step, where the score pass of "
                        "the previous batch runs beside this kernel on a helper stream; "
                        "step_achieved / step_frac = the same model over the whole step (per GPU)"}

    config = bw.bench_config(N, H, W, parts, n_points, B)
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config,
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "identical_to_device_path": e2e_same},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "best": {"index": gate_best[1], "score": gate_best[0]},
        "sharding": f"{TOTAL_CANDIDATES} candidates, {B} per GPU per step over {world} GPU(s)",
        "setup_s_first_scorer": round(t_setup, 3),
    }
    if parity is not None:
        out["parity"] = parity
    if sweep_total is not None:
        out["sweep_total"] = sweep_total
    if cpu is not None:
        out["cpu_baseline"] = cpu
    if world == 1 and not args.no_extra:
        out["other_configs"] = extra_configs(dev, not args.no_cpu_baseline)
        p5 = out["other_configs"].get("synthetic_1024_2048mask", {}).get("parity")
        if p5 is not None and p5.get("ok") is False:
            out.setdefault("parity", {})["config5_ok"] = False
        for key in ("taj_front_minarets", "taj_drone_minarets"):          # the real-data checks fail the run as well
            pt = out["other_configs"].get(key, {}).get("parity")
            if pt is not None and pt.get("ok") is False:
                out.setdefault("parity", {})["taj_ok"] = False
    if carve_multi is not None:
        out["carve"] = carve_multi
    if config5 is not None:
        out["config5_sweep"] = config5
        if isinstance(config5.get("parity"), dict) and config5["parity"].get("ok") is False:
            out.setdefault("parity", {})["config5_ok"] = False
    if world == 1 and not args.no_carve:
        del scorer
        torch.cuda.empty_cache()
        out["carve"] = carve_bench(N, dev, peak)
        if not args.no_cpu_baseline:
            try:
                out["carve"]["cpu_baseline"] = carve_cpu_baseline()
                if out["carve"]["cpu_baseline"]["parity"]["ok"] is False:
                    out.setdefault("parity", {})["carve_ok"] = False
            except Exception as exc:
                out["carve"]["cpu_baseline"] = {"error": repr(exc)}
    _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    failed = ((parity is not None and parity.get("ok") is False) or out.get("parity", {}).get("config5_ok") is False
              or out.get("parity", {}).get("carve_ok") is False or out.get("parity", {}).get("taj_ok") is False)
    if failed:
        print("PARITY GATE FAILED: " + json.dumps(out.get("parity")), file=sys.stderr)
        sys.exit(1)


def config5_sweep(dev, world, rank, dist, with_cpu):
    """BASELINE.json configs[4] across the GPUs of the box: the 1024^3 synthetic monument (174.8 M points) against a
    2048x2048 mask, 128 candidates per GPU per pass (weak scaling, contiguous shards of 128 * world candidates), with the
    16-byte all-gather of the best candidate.  CUDA events, max over ranks; per-GPU fraction of the streaming roofline
    (G*1 B + 9 B*H*W per candidate); rank 0 checks one candidate of its shard against the reference's functions."""
    import torch
    syn = importlib.import_module(PKG + ".synthetic")
    ce = importlib.import_module(PKG + ".utils.camera_estimation")
    cfg = importlib.import_module(PKG + ".utils.config")
    sweep_mod = importlib.import_module(PKG + ".utils.sweep")
    N, H, W, B = 1024, 2048, 2048, 128
    rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
    base = syn.base_camera(N, H, W, "front")
    full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
    gt = full.render(ce.row_to_params(base + HIDDEN_DELTA))
    del full
    sc = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
    del rgb
    cand = syn.candidates(base, B * world)
    mine = torch.from_numpy(np.ascontiguousarray(cand[rank * B:(rank + 1) * B])).to(dev)
    reducer = sweep_mod.BestReducer(dev, world)

    def step():
        counts, scores, best = sc.score_device(mine)
        return reducer.reduce(scores, best, rank * B), counts, scores
    step()
    torch.cuda.synchronize()
    dist.barrier()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res, counts, scores = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / reps
    peak, _ = peaks()
    model = peak * 1e9 / (N ** 3 + 9 * H * W)
    value = B * world / (ms * 1e-3)
    out = {"value": round(value, 1), "unit": UNIT, "n_gpus": world, "per_gpu": round(value / world, 1),
           "grid": N, "mask": [H, W], "points": sc.n_points, "candidates_per_gpu": B, "ms_per_pass": round(ms, 3),
           "scaling": "weak", "best": {"index": int(res[1]), "score": float(res[0])},
           "roofline": {"bound": "hbm", "model_candidates_per_s_per_gpu": round(model, 1), "peak": peak, "unit": "GB/s",
                        "achieved": round(value / world * (N ** 3 + 9 * H * W) / 1e9, 1), "frac": round(value / world / model, 4),
                        "note": "whole sweep call per GPU (splat + score + clear + best reduction) over the streaming model "
                                "G*1 B + 9 B*H*W per candidate, max over ranks"}}
    if with_cpu and rank == 0:
        try:
            import cpu_arm
            pts = sc.pts.cpu().numpy()
            lut = np.zeros((256, 3), np.uint8)
            lut[1:1 + len(sc.colours)] = np.array(sc.colours, np.uint8)
            cols = lut[sc.pt_label.cpu().numpy()]
            cpu = cpu_arm.CpuScorer(pts, cols, gt, syn.PART_NAMES, part_colors=cfg.PART_COLORS)
            pick = [B - 1]
            gc = counts.cpu().numpy()[pick][:, sc._cols, :]
            dt, n, bad, identical = compare_with_cpu(cpu, cand[pick], gc, scores.cpu().numpy()[pick], 1)
            out["parity"] = {"checked": n, "ok": not bad, "kind": cpu.kind, "scores_bit_identical": identical, "mismatches": bad[:2]}
            del pts, cols, cpu
        except MemoryError as exc:
            out["parity"] = {"checked": 0, "ok": None, "error": repr(exc)}
    # the second view of the "multi-view" sweep: the same point list against the aerial mask (CandidateScorer.set_image)
    try:
        base_a = syn.base_camera(N, H, W, "aerial")
        sc.set_image(sc.render(ce.row_to_params(base_a + HIDDEN_DELTA)))
        cand_a = syn.candidates(base_a, B * world)
        mine = torch.from_numpy(np.ascontiguousarray(cand_a[rank * B:(rank + 1) * B])).to(dev)
        step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res_a, _, _ = step()
        e1.record()
        torch.cuda.synchronize()
        ta = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        va = B * world / (float(ta.item()) / reps * 1e-3)
        out["aerial"] = {"value": round(va, 1), "per_gpu": round(va / world, 1), "frac": round(va / world / model, 4),
                         "best": {"index": int(res_a[1]), "score": float(res_a[0])}}
    except Exception as exc:
        out["aerial"] = {"error": repr(exc)}
    del sc
    torch.cuda.empty_cache()
    return out


def whole_sweep(ce, cfg, sweep_mod, nv, rgb, gt, parts, cand_all, B, dev, world, rank, dist):
    """BASELINE.json configs[3] taken literally: all 65 536 candidates sharded contiguously over the ranks, timed from a
    device-resident RGB grid + the host candidate array to the global best on every rank, set-up included.  Wall clock
    between synchronised barriers, max over ranks (run twice, the second run is reported: the first one pays the
    allocator's growth for the second scorer)."""
    import torch
    K = len(cand_all)
    lo, hi = sweep_mod.shard_range(K, world, rank)
    pinned = torch.from_numpy(np.ascontiguousarray(cand_all[lo:hi])).pin_memory()
    res = None
    for rep in range(2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        scorer = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, parts)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        n_blocks = -(-(hi - lo) // B)
        pairs = torch.empty((max(n_blocks, 1), 2), dtype=torch.int64, device=dev)
        pairs[:, 0] = torch.tensor(np.float64(-1.0).view(np.int64).item(), dtype=torch.int64, device=dev)
        pairs[:, 1] = -1
        for b in range(n_blocks):
            blk = pinned[b * B:(b + 1) * B].to(dev, non_blocking=True)
            counts, scores, best = scorer.score_device(blk)
            nv.check(nv.lib.p3d_best_pack(nv.ptr(scores), nv.ptr(best), lo + b * B, nv.ptr(pairs[b]), nv.stream_ptr()))
        mine = torch.empty(2, dtype=torch.int64, device=dev)
        nv.check(nv.lib.p3d_best_select(nv.ptr(pairs), max(n_blocks, 1), nv.ptr(mine), nv.stream_ptr()))
        if world > 1:
            gathered = torch.empty((world, 2), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(gathered.view(-1), mine)
            out = torch.empty(2, dtype=torch.int64, device=dev)
            nv.check(nv.lib.p3d_best_select(nv.ptr(gathered), world, nv.ptr(out), nv.stream_ptr()))
        else:
            out = mine
        host = out.cpu().numpy()                              # the global best on this rank (synchronises)
        t2 = time.perf_counter()
        tt = torch.tensor([t2 - t0, t1 - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res = {"seconds": round(float(tt[0].item()), 4), "setup_seconds": round(float(tt[1].item()), 4),
               "candidates": K, "candidates_per_s": round(K / float(tt[0].item()), 1), "n_gpus": world,
               "shard_of_rank0": [int(lo), int(hi)], "blocks_per_rank": n_blocks, "scaling": "strong",
               "best": {"index": int(host[1]), "score": float(host[:1].view(np.float64)[0])},
               "note": "wall clock, max over ranks: CandidateScorer construction (rgb -> labels, ordered compaction, x-run "
                       "segments, ground-truth labels) + this rank's contiguous shard in blocks of candidates_per_gpu_per_step "
                       "(pinned host -> device copies inside) + local arg-max + one 16-byte all-gather + host read of the winner"}
        del scorer
    return res


def parity_gate(args, scorer, gt, parts, cand_all, block_idx, gate_counts, gate_scores, gate_best, offset, B, H, W, dev,
                world, rank, dist, sweep_mod):
    """See the module docstring.  Returns (parity object on every rank, cpu_baseline object on rank 0 at N = 1)."""
    import torch
    import cpu_arm
    cfgm = importlib.import_module(PKG + ".utils.config")
    pts = scorer.pts.cpu().numpy()
    lut = np.zeros((256, 3), np.uint8)
    lut[1:1 + len(scorer.colours)] = np.array(scorer.colours, np.uint8)
    cols = lut[scorer.pt_label.cpu().numpy()]
    gt_np = gt.cpu().numpy() if isinstance(gt, torch.Tensor) else gt
    cpu = cpu_arm.CpuScorer(pts, cols, gt_np, parts, part_colors=cfgm.PART_COLORS)
    threads = cpu_arm.host_threads()
    if world == 1:
        procs = min(threads, 16)
        extra = args.cpu_sample if args.cpu_sample > 0 else max(0, 2 * procs - len(PARITY_INDICES))
        idx = [i for i in PARITY_INDICES if i < B]
        rng = np.random.default_rng(7)
        idx += [int(i) for i in rng.choice(B, size=min(extra, B), replace=False) if int(i) not in idx]
    else:                                                     # every rank: two candidates of its own block, two processes
        procs = max(1, min(2, threads // max(world, 1)))
        idx = [i for i in (127, 1024) if i < B] or [0]
    rows = cand_all[block_idx[idx]]
    g_counts = gate_counts[idx][:, scorer._cols, :]
    dt, n, bad, identical = compare_with_cpu(cpu, rows, g_counts, gate_scores[idx], procs)
    ok = not bad
    nccl = None
    if world > 1:
        # the NCCL winner of that step against the arg-max over an all-gather of the full score vectors
        sc = torch.from_numpy(gate_scores).to(dev)
        allsc = torch.empty((world, B), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allsc.view(-1), sc)
        allsc = allsc.cpu().numpy()
        offs = torch.tensor([offset], dtype=torch.int64, device=dev)
        alloff = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(alloff, offs)
        alloff = alloff.cpu().numpy()
        pairs = [(float(allsc[r, i]), int((alloff[r] + i) % TOTAL_CANDIDATES)) for r in range(world)
                 for i in np.flatnonzero(allsc[r] == allsc[r].max())]
        want = sweep_mod.select_best(pairs)
        nccl = {"winner": [gate_best[0], gate_best[1]], "argmax_of_gathered_scores": [want[0], want[1]],
                "ok": bool(want[0] == gate_best[0] and want[1] == gate_best[1])}
        ok = ok and nccl["ok"]
        # this rank re-scores the winner locally through the public API: same score on every rank
        w_row = cand_all[gate_best[1]][None]
        w_s, _, _ = scorer.score(w_row)
        ok = ok and bool(w_s[0] == gate_best[0])
        flag = torch.tensor([1 if ok else 0, n], dtype=torch.int64, device=dev)
        tot = flag.clone()
        dist.all_reduce(flag[:1], op=dist.ReduceOp.MIN)
        dist.all_reduce(tot[1:], op=dist.ReduceOp.SUM)
        ok_all, n_all = bool(flag[0].item()), int(tot[1].item())
    else:
        ok_all, n_all = ok, n
    parity = {"checked": n_all, "ok": ok_all, "kind": cpu.kind, "what": "per-part (inter, union) counts exactly, scores within 1e-5 "
              "relative, against the reference's project_colored_voxels + compute_partwise_iou on the host",
              "scores_bit_identical": identical, "block_indices_rank0": [int(i) for i in idx][:16],
              "step": "last timed step", "mismatches": bad[:4]}
    if nccl is not None:
        parity["nccl_best"] = nccl
        parity["winner_rescored_on_every_rank"] = True
    cpu_obj = None
    if world == 1:
        cpu_obj = {"value": round(n / dt, 4), "unit": UNIT, "cores": procs, "kind": cpu.kind,
                   "sample": f"{n} candidates of the last timed step (the parity-gate sample) on the same grid/mask, "
                             f"candidate-parallel over {procs} processes with one BLAS thread each, {dt:.1f} s",
                   "host_threads": threads, "blas_env": cpu_arm.blas_env()}
    return parity, cpu_obj


def carve_sharded_bench(N, dev, world, rank, dist):
    """global_carve of the synthetic monument's front silhouette at N^3, one x-slab per rank (utils.sweep.carve_sharded:
    masks replicated, no data-path collective).  Strong scaling: the grid is fixed, the slab shrinks with N."""
    import torch
    syn = importlib.import_module(PKG + ".synthetic")
    vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
    cfg = importlib.import_module(PKG + ".utils.config")
    sw = importlib.import_module(PKG + ".utils.sweep")
    lab = syn.monument_labels(N, dev)
    front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
    del lab
    lut = syn.label_lut()
    lut[0] = cfg.PART_COLORS["background"]
    ext = torch.from_numpy(lut[front]).to(dev)
    binm = (front > 0).astype(np.uint8)
    binm_d = torch.from_numpy(binm).to(dev)                  # masks resident on the device, as for a caller chaining stages
    carve = lambda a, b: vc.global_carve(binm_d, ext, 90, return_tensor=True, x_range=(a, b))
    for _ in range(5):
        slab, span = sw.carve_sharded(carve, N)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    reps = 40        # (round 2 timed 5 calls: the empty-queue start-up of the window was a third of the figure at N = 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        slab, span = sw.carve_sharded(carve, N)
    e1.record()
    torch.cuda.synchronize()
    # kernel-only: this rank's slab through the C ABI, graph-replayed (the Python call is launch-overhead bound)
    kms = 0.0
    try:
        nv = importlib.import_module(PKG + ".utils._native")
        M, off = vc._pass_transform((N, N, N), 90)
        table, foldable = vc._fold_table(N, N, M, off, dev)
        bits = vc._fold_bits(table, N, N, (N, N, M.tobytes(), off.tobytes(), str(dev))) if foldable else None
        if bits is not None:
            wpr = (N + 31) // 32 + 2
            m_hw = torch.from_numpy(binm).to(dev)
            mbits = torch.empty((N, wpr), dtype=torch.int32, device=dev)
            nv.check(nv.lib.p3d_pack_mask_bits(nv.ptr(m_hw), N, N, nv.ptr(mbits), wpr, nv.stream_ptr()))
            x0, x1 = span
            kout = torch.empty_like(slab)
            kms, _ = time_launches(lambda: nv.check(nv.lib.p3d_global_carve_fold_bits(
                N, N, N, x0, x1 - x0, nv.ptr(bits[0]), bits[1], nv.ptr(mbits), wpr, nv.ptr(ext), 1, nv.ptr(kout), nv.stream_ptr())))
            assert torch.equal(kout, slab)
    except Exception as exc:
        print("sharded carve kernel timing failed:", repr(exc), file=sys.stderr)
    # part_carve of the same grid by output x-slab from the replicated input (no exchange): the four slab kernels
    pms = 0.0
    try:
        gfull = vc.global_carve(binm, ext, 90, return_tensor=True)
        jobs90 = [([n], 90) for n in ("full_building", "chhatris", "plinth", "front_minarets", "small_minarets", "dome")]
        pslab = vc.part_carve(gfull, ext, jobs90, x_range=span)
        launch_pc, kslab = part_carve_launcher(gfull, ext, jobs90, dev, x_range=span)
        if launch_pc is not None:
            pms, _ = time_launches(launch_pc, reps=10, rounds=3)
            assert torch.equal(kslab, pslab)
        del gfull, pslab, kslab
    except Exception as exc:
        print("sharded part_carve timing failed:", repr(exc), file=sys.stderr)
    # part_carve with the INPUT sharded as well (rank r holds only its x rows): pass A -> exchange of occupancy bits ->
    # pass B; 6 B per slab voxel + the exchange.  Three exchange forms (utils.sweep.part_carve_sharded), whole calls
    # through the public function, CUDA events, max over ranks.
    sharded_ms = {}
    try:
        gfull = vc.global_carve(binm, ext, 90, return_tensor=True)
        x0, x1 = span
        slab_in = gfull[x0:x1].contiguous()
        want = vc.part_carve(gfull, ext, jobs90, x_range=span)
        del gfull
        for mode in ("alltoall", "peer", "allgather"):
            try:
                got, _ = sw.part_carve_sharded(slab_in, ext, jobs90, N, exchange=mode)
                assert torch.equal(got, want), mode
                for _ in range(3):
                    sw.part_carve_sharded(slab_in, ext, jobs90, N, exchange=mode)
                torch.cuda.synchronize()
                dist.barrier()
                torch.cuda.synchronize()
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                for _ in range(20):
                    sw.part_carve_sharded(slab_in, ext, jobs90, N, exchange=mode)
                eb.record()
                torch.cuda.synchronize()
                sharded_ms[mode] = ea.elapsed_time(eb) / 20
            except Exception as exc:
                print(f"sharded-input part_carve ({mode}) failed:", repr(exc), file=sys.stderr)
                sharded_ms[mode] = 0.0
        # the same peer form with the PartCarveSlab object (its tables, output and symmetric workspace) reused: what is
        # left is pass A, a tiny all-reduce, pass B reading the peers over NVLink, a tiny all-reduce
        try:
            nbytes = vc.PartCarveSlab.workspace_bytes(N, N, N, len(jobs90))
            slots = []                      # two slab objects on two symmetric workspaces used in turn: one barrier per call
            for slot in (0, 1):
                buf, hdl, ptrs = sw.symmetric_workspace(nbytes, dev, slot=slot)
                slots.append((vc.PartCarveSlab(slab_in, ext, jobs90, N, span, workspace=buf), hdl, ptrs))
            turn = [0]

            def steady():
                job, hdl, ptrs = slots[turn[0] & 1]
                turn[0] += 1
                job.begin()
                sw.peer_barrier(hdl)
                return job.finish(peers=ptrs, n_ranks=world)
            assert torch.equal(steady(), want)
            for _ in range(3):
                steady()
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(20):
                steady()
            eb.record()
            torch.cuda.synchronize()
            sharded_ms["peer_steady"] = ea.elapsed_time(eb) / 20
            del slots
        except Exception as exc:
            print("sharded-input part_carve (peer, steady) failed:", repr(exc), file=sys.stderr)
        del want, slab_in
    except Exception as exc:
        print("sharded-input part_carve timing failed:", repr(exc), file=sys.stderr)
    sms, xms = sharded_ms.get("alltoall", 0.0), sharded_ms.get("peer", 0.0)
    gms_ag = sharded_ms.get("allgather", 0.0)
    pst = sharded_ms.get("peer_steady", 0.0)
    t = torch.tensor([e0.elapsed_time(e1) / reps, kms, pms, sms, xms, gms_ag, pst], dtype=torch.float64, device=dev)
    occ = torch.count_nonzero(slab.view(-1, 3).any(dim=1)).to(torch.float64).reshape(1)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(occ, op=dist.ReduceOp.SUM)
    ms, kms, pms, sms, xms, gms_ag, pst = (float(v) for v in t.tolist())
    return {"global_carve_sharded_gvoxel_s": round(N ** 3 / (ms * 1e-3) / 1e9, 2), "grid": N, "ms_per_call": round(ms, 4),
            "kernel_ms_max_over_ranks": round(kms, 4),
            "kernel_gvoxel_s": round(N ** 3 / (kms * 1e-3) / 1e9, 2) if kms > 0 else None,
            "part_carve_slab_kernel_ms_max_over_ranks": round(pms, 4),
            "part_carve_gvoxel_s": round(N ** 3 / (pms * 1e-3) / 1e9, 2) if pms > 0 else None,
            "part_carve_sharded_input": {
                "alltoall_ms": round(sms, 4), "peer_ms": round(xms, 4), "allgather_ms": round(gms_ag, 4),
                "peer_steady_ms": round(pst, 4),
                "peer_steady_gvoxel_s": round(N ** 3 / (pst * 1e-3) / 1e9, 2) if pst > 0 else None,
                "best_gvoxel_s": round(N ** 3 / (min(v for v in (sms, xms, gms_ag) if v > 0) * 1e-3) / 1e9, 2) if max(sms, xms, gms_ag) > 0 else None,
                "note": "whole part_carve_sharded call per rank (PartCarveSlab set-up, pass A, exchange, pass B), max over "
                        "ranks: alltoall = only the z-bit words each slab reads (W*H*D/(8 world) bytes received per rank), "
                        "peer = pass B reads the other ranks' rows in NVLink peer-mapped symmetric memory (no exchange "
                        "buffer), allgather = the whole bit array (W*H*D/8 bytes) on every rank"},
            "n_gpus": world, "scaling": "strong", "occupied": int(occ.item()), "slab_of_rank0": list(span),
            "note": "whole Python call per rank (device-resident masks, table lookup, slab kernel), x-slab per rank, max over ranks; "
                    "no collective on the data path"}


def part_carve_launcher(grid, ext, jobs, dev, x_range=None):
    """The kernels of part_carve's all-90-degree bit path as a re-issuable closure (bench scaffolding: inputs prepared
    with the package's own helpers, launches through the C ABI), for graph-replayed kernel-only timing.  Returns
    (launch, output tensor) or (None, None) when the grid does not take that path."""
    import torch
    vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
    nv = importlib.import_module(PKG + ".utils._native")
    W, H, D, _ = grid.shape
    if D != W or D % 32:
        return None, None
    M, off = vc._pass_transform((W, H, D), 90)
    table, foldable = vc._fold_table(W, D, M, off, dev)
    bits = vc._fold_bits(table, W, D, (W, D, M.tobytes(), off.tobytes(), str(dev))) if foldable else None
    if bits is None or bits[2] is None:
        return None, None
    gm = vc._group_image_device(vc._PackedMask(ext), [[vc.PART_COLORS[n] for n in names] for names, _ in jobs], dev)
    ws_bytes = int(nv.lib.p3d_part_carve_bits_workspace_bytes(W, H, D, len(jobs)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if x_range is None:
        out = torch.empty_like(grid)

        def launch():
            nv.check(nv.lib.p3d_part_carve_fold_bits(nv.ptr(grid), W, H, D, nv.ptr(bits[0]), bits[1], bits[2], nv.ptr(gm), len(jobs),
                                                     nv.ptr(out), nv.ptr(ws), ws_bytes, nv.stream_ptr()))
    else:
        x0, x1 = x_range
        out = torch.empty((x1 - x0, H, D, 3), dtype=torch.uint8, device=dev)

        def launch():
            nv.check(nv.lib.p3d_part_carve_fold_bits_slab(nv.ptr(grid), W, H, D, x0, x1 - x0, nv.ptr(bits[0]), bits[1], bits[2],
                                                          nv.ptr(gm), len(jobs), nv.ptr(out), nv.ptr(ws), ws_bytes, nv.stream_ptr()))
    return launch, out


def time_launches(launch, reps=20, rounds=5, use_graph=True):
    """Average device time of one `launch()` (a library call that only enqueues kernels on the current stream): `reps`
    launches captured in a CUDA graph and replayed, so host-side launch cost (ctypes, GIL contention with the clock
    sampler thread) cannot bound kernels that take tens of microseconds; best of `rounds` replays.  Falls back to a
    plain launch loop if capture is refused."""
    import torch
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    graph = None
    try:
        if not use_graph:                     # launches that allocate or copy from the host cannot be captured
            raise RuntimeError("plain loop requested")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) stay free
            for _ in range(reps):
                launch()
        graph = g
    except Exception:
        graph = None
        torch.cuda.synchronize()
    best = None
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for _ in range(reps):
                launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        best = ms if best is None else min(best, ms)
    return best, graph is not None


def carve_kernels_at(N, dev, peak):
    """Kernel-only carve timings at N^3 (BASELINE.json configs[4] carves 1024^3): global_carve's fused fold+colour kernel
    (3 B/voxel) and the three part_carve kernels (6 B/voxel), graph-replayed like the 512^3 figures."""
    import torch
    syn = importlib.import_module(PKG + ".synthetic")
    vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
    cfg = importlib.import_module(PKG + ".utils.config")
    lab = syn.monument_labels(N, dev)
    front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
    del lab
    lut = syn.label_lut()
    lut[0] = cfg.PART_COLORS["background"]
    ext = torch.from_numpy(lut[front]).to(dev)
    binm = (front > 0).astype(np.uint8)
    binm_d = torch.from_numpy(binm).to(dev)
    out = vc.global_carve(binm_d, ext, 90, return_tensor=True)
    gms, _ = time_launches(lambda: vc.global_carve(binm_d, ext, 90, return_tensor=True), reps=5, rounds=3, use_graph=False)
    jobs90 = [([n], 90) for n in ("full_building", "chhatris", "plinth", "front_minarets", "small_minarets", "dome")]
    pc = vc.part_carve(out, ext, jobs90)
    res = {"grid": N, "voxels": N ** 3, "global_carve_call_ms": round(gms, 4),
           "global_carve_gvoxel_s": round(N ** 3 / (gms * 1e-3) / 1e9, 2),
           "global_carve_frac_of_peak": round(3 * N ** 3 / (gms * 1e-3) / 1e9 / peak, 4)}
    launch_pc, kout = part_carve_launcher(out, ext, jobs90, dev)
    if launch_pc is not None:
        pms, _ = time_launches(launch_pc, reps=10, rounds=3)
        assert torch.equal(kout, pc)
        res.update({"part_carve_kernel_ms": round(pms, 4), "part_carve_gvoxel_s": round(N ** 3 / (pms * 1e-3) / 1e9, 2),
                    "part_carve_frac_of_peak": round(6 * N ** 3 / (pms * 1e-3) / 1e9 / peak, 4)})
    res["note"] = ("global_carve: whole Python call on the device (mask upload, cached tables, one kernel), device tensor "
                   "out, 3 B/voxel; part_carve: the three kernels, 6 B/voxel")
    del out, pc
    torch.cuda.empty_cache()
    return res


def carve_bench(N, dev, peak):
    """Second metric of BASELINE.json: carving Gvoxel/s vs the HBM roofline.  global_carve of the synthetic monument's
    front silhouette at N^3 (3 B per output voxel), timed with CUDA events over the fused kernel; plus the
    whole Bibi@256 pipeline (global_carve + partwise_carve, notebook-1 jobs) on the real mask as wall time."""
    import contextlib
    import io
    import torch
    syn = importlib.import_module(PKG + ".synthetic")
    vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
    cfg = importlib.import_module(PKG + ".utils.config")
    mu = importlib.import_module(PKG + ".utils.mask_utils")
    lab = syn.monument_labels(N, dev)
    front = torch.flip(lab.max(dim=0).values, dims=[0]).cpu().numpy()
    lut = syn.label_lut()
    lut[0] = cfg.PART_COLORS["background"]
    ext = torch.from_numpy(lut[front]).to(dev)
    binm = (front > 0).astype(np.uint8)
    del lab
    binm_d = torch.from_numpy(binm).to(dev)                  # device-resident masks (a caller chaining stages on the device)
    for _ in range(3):
        out = vc.global_carve(binm_d, ext, 90, return_tensor=True)
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = vc.global_carve(binm_d, ext, 90, return_tensor=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gvox = N ** 3 / (ms * 1e-3) / 1e9
    # kernel-only: the fused fold+colour kernel through the C ABI on prepared device inputs
    nv = importlib.import_module(PKG + ".utils._native")
    M, off = vc._pass_transform((N, N, N), 90)
    table, foldable = vc._fold_table(N, N, M, off, dev)
    m_hw = torch.from_numpy(binm).to(dev)
    kout = torch.empty((N, N, N, 3), dtype=torch.uint8, device=dev)
    kms = None
    bits = vc._fold_bits(table, N, N, (N, N, M.tobytes(), off.tobytes(), str(dev))) if foldable else None
    if bits is not None:
        wpr = (N + 31) // 32 + 2
        mbits = torch.empty((N, wpr), dtype=torch.int32, device=dev)
        nv.check(nv.lib.p3d_pack_mask_bits(nv.ptr(m_hw), N, N, nv.ptr(mbits), wpr, nv.stream_ptr()))

        def launch():
            nv.check(nv.lib.p3d_global_carve_fold_bits(N, N, N, 0, N, nv.ptr(bits[0]), bits[1], nv.ptr(mbits), wpr, nv.ptr(ext), 1,
                                                       nv.ptr(kout), nv.stream_ptr()))
        kms, graphed = time_launches(launch)
        assert torch.equal(kout, out)
    kgvox = N ** 3 / (kms * 1e-3) / 1e9 if kms else None
    res = {"global_carve_gvoxel_s": round(gvox, 2), "grid": N, "ms_per_call": round(ms, 4),
           "occupied": int(torch.count_nonzero(out.view(-1, 3).any(dim=1)).item()),
           "kernel_gvoxel_s": round(kgvox, 2) if kgvox else None, "kernel_ms": round(kms, 4) if kms else None,
           "roofline": {"bound": "hbm", "achieved": round(3 * kgvox, 1) if kgvox else None, "peak": peak, "unit": "GB/s",
                        "frac": round(3 * kgvox / peak, 4) if kgvox else None, "kernel": "global_fold_bits_kernel<RGB>",
                        "note": "3 B per output voxel (RGB grid written once, SURVEY 8d), kernel-only (20 launches replayed "
                                "from a CUDA graph, best of 5); "
                                "global_carve_gvoxel_s is the whole Python call (device-resident masks, table lookup, two launches)"}}
    # part_carve (all six notebook groups at 90 degrees) on that grid: 6 B per voxel (read RGB + write RGB)
    jobs90 = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
              (["small_minarets"], 90), (["dome"], 90)]
    pc = vc.part_carve(out, ext, jobs90)                      # first call: allocator growth, cached tables
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        pc = vc.part_carve(out, ext, jobs90)
    torch.cuda.synchronize()
    pc_call_ms = (time.perf_counter() - t0) * 1e3 / 5
    launch_pc, kpc = part_carve_launcher(out, ext, jobs90, dev)
    if launch_pc is not None:
        pms, _ = time_launches(launch_pc)
        assert torch.equal(kpc, pc)
        res["part_carve"] = {"kernel_ms": round(pms, 4), "kernel_gvoxel_s": round(N ** 3 / (pms * 1e-3) / 1e9, 2),
                             "call_ms": round(pc_call_ms, 3), "occupied": int(torch.count_nonzero(pc.view(-1, 3).any(dim=1)).item()),
                             "roofline": {"bound": "hbm", "achieved": round(6 * N ** 3 / (pms * 1e-3) / 1e9, 1), "peak": peak,
                                          "unit": "GB/s", "frac": round(6 * N ** 3 / (pms * 1e-3) / 1e9 / peak, 4),
                                          "kernel": "pack_group_bits + part_copy_bits + part_clear",
                                          "note": "6 B per voxel (SURVEY 8d: read RGB + write RGB); copy-then-clear: pass A "
                                                  "writes the output from voxel-local terms and packs occupancy/alive bits "
                                                  "(~0.25 B/voxel), pass B reads bits only and rewrites the runs whose "
                                                  "rotated source is empty (none for this 4-way-symmetric grid); call_ms = "
                                                  "the whole Python call (device tensors in and out), mean of 5 after a warm-up"}}
        # the hard input: the same grid with 2 % of its occupied voxels knocked out at random -- no longer 4-way symmetric,
        # so the clear pass has real work
        try:
            gen = torch.Generator(device=dev)
            gen.manual_seed(3)
            hole = torch.rand(out.shape[:3], device=dev, generator=gen) < 0.02
            asym = out.clone()
            asym[hole] = 0
            del hole
            want = vc.part_carve(asym, ext, jobs90)
            launch_as, kas = part_carve_launcher(asym, ext, jobs90, dev)
            ams, _ = time_launches(launch_as)
            assert torch.equal(kas, want)
            res["part_carve"]["asymmetric_input"] = {
                "kernel_ms": round(ams, 4), "frac": round(6 * N ** 3 / (ams * 1e-3) / 1e9 / peak, 4),
                "cleared_voxels": int((torch.count_nonzero(asym.view(-1, 3).any(dim=1)) - torch.count_nonzero(want.view(-1, 3).any(dim=1))).item()),
                "note": "global_carve output with 2 % of the occupied voxels removed at random: every removed voxel empties up "
                        "to three rotated partners, the clear pass rewrites those runs"}
            del asym, want, kas
        except Exception as exc:
            res["part_carve"]["asymmetric_input"] = {"error": repr(exc)}
        del kpc
    del out, kout, pc
    torch.cuda.empty_cache()
    if N != 1024:
        try:
            res["synthetic_1024"] = carve_kernels_at(1024, dev, peak)
        except Exception as exc:
            res["synthetic_1024"] = {"error": repr(exc)}
    data = os.path.join(ROOT, "tests", "golden", "data")
    try:
        sem, sem_ext, binary = mu.load_and_prepare_masks(data, "Bibi", "front", 256, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
        jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
                (["small_minarets"], 90), (["dome"], 90)]
        sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
        ext_d = {"main_door": 20, "windows": 10}
        best = None
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g = vc.global_carve(binary, sem_ext, 90)
                final = vc.partwise_carve(g, sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
        dbest = None                   # same chain with the grid kept on the device between the two calls
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                gd = vc.global_carve(binary, sem_ext, 90, return_tensor=True)
                fd = vc.partwise_carve(gd, sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                dbest = dt if dbest is None else min(dbest, dt)
        same = bool(torch.equal(fd, torch.from_numpy(final).to(fd.device))) if isinstance(fd, torch.Tensor) else None
        res["bibi256_pipeline"] = {"wall_ms": round(best * 1e3, 2), "device_chain_wall_ms": round(dbest * 1e3, 2),
                                   "device_chain_identical": same,
                                   "voxels": int(g.shape[0] * g.shape[1] * g.shape[2]),
                                   "gvoxel_s": round(g.shape[0] * g.shape[1] * g.shape[2] / best / 1e9, 4),
                                   "note": "real Bibi front mask, load_and_prepare_masks(max_dim=256) -> global_carve -> "
                                           "partwise_carve, NumPy in / NumPy out (host copies included), best of 3; device_chain_wall_ms: the "
                                           "same two calls with device tensors in between (return_tensor=True)"}
    except Exception as exc:      # the real mask is a test asset; never fail the headline bench because of it
        res["bibi256_pipeline"] = {"error": repr(exc)}
    return res


def extra_configs(dev, with_cpu):
    """BASELINE.json configs 2 and 3 as side measurements (not the headline): the real Taj grid + front mask with the
    stored final camera (minaret parts as in notebook 2, and all parts), and the synthetic 256^3 / 4096-candidate case."""
    import torch
    syn = importlib.import_module(PKG + ".synthetic")
    ce = importlib.import_module(PKG + ".utils.camera_estimation")
    cfg = importlib.import_module(PKG + ".utils.config")
    mu = importlib.import_module(PKG + ".utils.mask_utils")
    out = {}

    def rate(scorer, cand, reps=3):
        cd = torch.from_numpy(cand).to(dev)
        scorer.score_device(cd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            scorer.score_device(cd)
        e1.record()
        torch.cuda.synchronize()
        return len(cand) * reps / (e0.elapsed_time(e1) * 1e-3)

    data = os.path.join(ROOT, "tests", "golden", "data")
    try:
        grid = np.load(os.path.join(data, "results", "1.Orthographic_Voxel_Carving", "Taj_voxel_grid.npz"))["voxel_grid"]
        cams = json.load(open(os.path.join(data, "results", "2.Perspective_Camera_Estimation", "Taj_camera_params_final.json")))
        front = mu.load_mask(data, "Taj", "front", int(max(grid.shape)))
        c = cams["front"]
        base = np.array([*c["cam_pos"], *c["target"], c["f"], c["cx"], c["cy"]])
        cand = syn.candidates(base, 4096)
        gdev = torch.from_numpy(grid).to(dev)
        for tag, parts in (("minarets", ["front_minarets", "back_minarets"]), ("all_parts", syn.PART_NAMES)):
            sc = ce.CandidateScorer(gdev, front, cfg.PART_COLORS, parts)
            entry = {"points": sc.n_points, "mask": list(front.shape[:2]), "candidates": len(cand),
                     "value": round(rate(sc, cand), 1), "unit": UNIT}
            if with_cpu and tag == "minarets":             # real data: 64 candidates through the reference, counts compared
                import cpu_arm
                pts = sc.pts.cpu().numpy()
                lut = np.zeros((256, 3), np.uint8)
                lut[1:1 + len(sc.colours)] = np.array(sc.colours, np.uint8)
                cols = lut[sc.pt_label.cpu().numpy()]
                cpu = cpu_arm.CpuScorer(pts, cols, front, parts, part_colors=cfg.PART_COLORS)
                g_scores, g_counts, _ = sc.score(cand[:64])
                dt, n, bad, identical = compare_with_cpu(cpu, cand[:64], g_counts, g_scores, 1)
                entry["cpu_baseline"] = {"value": round(64 / dt, 2), "unit": UNIT, "cores": 1, "kind": cpu.kind,
                                         "sample": "first 64 candidates, 1 process, 1 BLAS thread"}
                entry["parity"] = {"checked": n, "ok": not bad, "scores_bit_identical": identical, "mismatches": bad[:2]}
            out["taj_front_" + tag] = entry
        # BASELINE.json configs[1] names front + aerial masks: the same sweep against the drone view's mask and stored camera
        try:
            drone = mu.load_mask(data, "Taj", "drone", int(max(grid.shape)))
            cd = cams["drone"]
            base_d = np.array([*cd["cam_pos"], *cd["target"], cd["f"], cd["cx"], cd["cy"]])
            cand_d = syn.candidates(base_d, 4096)
            for tag, parts in (("minarets", ["front_minarets", "back_minarets"]), ("all_parts", syn.PART_NAMES)):
                sc = ce.CandidateScorer(gdev, drone, cfg.PART_COLORS, parts)
                entry = {"points": sc.n_points, "mask": list(drone.shape[:2]), "candidates": len(cand_d),
                         "value": round(rate(sc, cand_d), 1), "unit": UNIT}
                if with_cpu and tag == "minarets":
                    import cpu_arm
                    lut = np.zeros((256, 3), np.uint8)
                    lut[1:1 + len(sc.colours)] = np.array(sc.colours, np.uint8)
                    cpu = cpu_arm.CpuScorer(sc.pts.cpu().numpy(), lut[sc.pt_label.cpu().numpy()], drone, parts,
                                            part_colors=cfg.PART_COLORS)
                    g_scores, g_counts, _ = sc.score(cand_d[:32])
                    dt, n, bad, identical = compare_with_cpu(cpu, cand_d[:32], g_counts, g_scores, 1)
                    entry["parity"] = {"checked": n, "ok": not bad, "scores_bit_identical": identical, "mismatches": bad[:2]}
                out["taj_drone_" + tag] = entry
        except Exception as exc:
            out["taj_drone_error"] = repr(exc)
        # notebook 3's part-wise deformation sweep on the same grid: dome, fixed final camera
        try:
            de = importlib.import_module(PKG + ".utils.deformation_estimation")
            import contextlib
            import io
            cam = {"cam_pos": np.array(c["cam_pos"]), "target": np.array(c["target"]), "f": c["f"], "cx": c["cx"], "cy": c["cy"]}
            labels = {k: v for k, v in cfg.PART_COLORS.items() if k != "background"}
            with contextlib.redirect_stdout(io.StringIO()):
                viewer = de.DeformViewer(gdev, labels, front, cam, ["dome"])
            rng = np.random.default_rng(7)
            D = 4096
            rows = np.column_stack([rng.uniform(0.8, 1.2, D), rng.uniform(-60, 60, D), rng.uniform(0.8, 1.2, D), rng.uniform(-60, 60, D)])
            viewer.score("dome", rows[:64])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ious, _, _ = viewer.score("dome", rows)
            e1.record()
            torch.cuda.synchronize()
            secs = e0.elapsed_time(e1) * 1e-3
            n_dome = viewer.part_points("dome").n
            entry = {"part": "dome", "points": n_dome, "jitters": 7, "deformations": D, "value": round(D / secs, 1),
                     "unit": "deformations/s", "point_jitter_deformations_per_s": round(D * 7 * n_dome / secs, 1),
                     "best_iou": float(ious.max()),
                     "note": "deform (FP64, reference op order) + bounds + projection through the fixed final camera + coverage "
                             "bitmap + IoU per deformation; host rows in, host IoUs out"}
            if with_cpu:
                from oracle import oracle as orc
                t0 = time.perf_counter()
                ref_iou, _ = orc.deform_part_iou(grid, labels, front, cam, "dome", de.row_to_deform(rows[0]))
                dt = time.perf_counter() - t0
                entry["cpu_baseline"] = {"value": round(1 / dt, 4), "unit": "deformations/s", "cores": 1, "kind": "port",
                                         "sample": "1 deformation through oracle/ (NumPy restatement incl. np.unique)",
                                         "iou_matches": bool(ref_iou == ious[0])}
            out["taj_front_deform_dome"] = entry
        except Exception as exc:
            out["taj_front_deform_dome"] = {"error": repr(exc)}
        # notebook 4's visibility evaluator: global min-Z depth buffer of all 12 M voxels + one part's visibility
        try:
            eh = importlib.import_module(PKG + ".utils.eval_helpers_intra")
            cam32 = {"cam_pos": np.array(c["cam_pos"], np.float32), "target": np.array(c["target"], np.float32),
                     "f": float(c["f"]), "cx": float(c["cx"]), "cy": float(c["cy"])}       # load_camera_json dtype
            Hm, Wm = front.shape[:2]
            best_z = best_v = None
            for _ in range(3):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                zb = eh.compute_global_depth_buffer(gdev, cam32, Hm, Wm, return_tensor=True)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                vis = eh.project_part_visible(viewer.part_points("dome").pts, cam32, zb, Hm, Wm, return_tensor=True)
                torch.cuda.synchronize(); t2 = time.perf_counter()
                best_z = t1 - t0 if best_z is None else min(best_z, t1 - t0)
                best_v = t2 - t1 if best_v is None else min(best_v, t2 - t1)
            entry = {"occupied_voxels": int(torch.count_nonzero(gdev.view(-1, 3).any(dim=1)).item()), "mask": [Hm, Wm],
                     "depth_buffer_ms": round(best_z * 1e3, 3), "part_visible_ms": round(best_v * 1e3, 3),
                     "visible_pixels": int(vis.sum().item()), "dtype": "f32",
                     "note": "compute_global_depth_buffer (occupancy + compaction + 32-bit atomicMin on depth bits) and "
                             "project_part_visible for the dome; device grid in, device tensors out, best of 3"}
            if with_cpu:
                from oracle import oracle as orc
                t0 = time.perf_counter()
                z_ref = orc.compute_global_depth_buffer(grid, cam32, Hm, Wm)
                dt = time.perf_counter() - t0
                entry["cpu_baseline"] = {"depth_buffer_ms": round(dt * 1e3, 1), "cores": 1, "kind": "port",
                                         "sample": "oracle/ C restatement (the reference itself loops over every voxel in Python)",
                                         "matches": bool(np.array_equal(z_ref, zb.cpu().numpy()))}
            out["taj_front_depth_visibility"] = entry
        except Exception as exc:
            out["taj_front_depth_visibility"] = {"error": repr(exc)}
        del gdev
    except Exception as exc:
        out["taj"] = {"error": repr(exc)}
    try:
        N, H, W = 256, 1024, 1024
        rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
        base = syn.base_camera(N, H, W)
        full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
        gt = full.render(ce.row_to_params(base + HIDDEN_DELTA))
        sc = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
        v256 = rate(sc, syn.candidates(base, 4096), reps=2)
        peak256, _ = peaks()
        model256 = peak256 * 1e9 / (N ** 3 + 9 * H * W)
        out["synthetic_256_4096cand"] = {"points": sc.n_points, "mask": [H, W], "candidates": 4096,
                                         "value": round(v256, 1), "unit": UNIT,
                                         "roofline": {"bound": "hbm", "model_candidates_per_s": round(model256, 1), "peak": peak256,
                                                      "unit": "GB/s", "achieved": round(v256 * (N ** 3 + 9 * H * W) / 1e9, 1),
                                                      "frac": round(v256 / model256, 4),
                                                      "note": "BASELINE.json configs[2]; whole sweep call over the streaming model "
                                                              "G*1 B + 9 B*H*W per candidate; 2.7 M points take the per-point splat "
                                                              "(voxel pitch 3.4 px: every lane of an early-out load in its own sector)"}}
    except Exception as exc:
        out["synthetic_256_4096cand"] = {"error": repr(exc)}
    try:        # BASELINE.json configs[4] on one GPU: 1024^3 grid, 2048x2048 masks (front + aerial views)
        N, H, W = 1024, 2048, 2048
        del rgb, full, sc
        torch.cuda.empty_cache()
        rgb = torch.from_numpy(syn.label_lut()).to(dev)[syn.monument_labels(N, dev).long()]
        entry = {"mask": [H, W], "candidates": 256, "unit": UNIT}
        peak, _ = peaks()
        model = peak * 1e9 / (N ** 3 + 9 * H * W)              # candidates/s at the HBM roofline of the streaming model
        for view in ("front", "aerial"):
            base = syn.base_camera(N, H, W, view)
            full = ce.CandidateScorer(rgb, torch.zeros((H, W, 3), dtype=torch.uint8, device=dev), cfg.PART_COLORS, syn.PART_NAMES)
            gt = full.render(ce.row_to_params(base + HIDDEN_DELTA))
            del full
            sc = ce.CandidateScorer(rgb, gt, cfg.PART_COLORS, syn.PART_NAMES)
            entry["points"] = sc.n_points
            cand = syn.candidates(base, 256)
            entry[view] = round(rate(sc, cand, reps=2), 1)
            entry[view + "_frac_of_streaming_roofline"] = round(entry[view] / model, 4)
            if with_cpu and view == "front":                    # parity at this size: 2 candidates through the reference
                try:
                    import cpu_arm
                    pts = sc.pts.cpu().numpy()
                    lut = np.zeros((256, 3), np.uint8)
                    lut[1:1 + len(sc.colours)] = np.array(sc.colours, np.uint8)
                    cols = lut[sc.pt_label.cpu().numpy()]
                    cpu = cpu_arm.CpuScorer(pts, cols, gt, syn.PART_NAMES, part_colors=cfg.PART_COLORS)
                    pick = [1, 200]                              # different z-buffer batches of the 256-candidate sweep
                    g_scores, g_counts, _ = sc.score(cand)
                    dt, n, bad, identical = compare_with_cpu(cpu, cand[pick], g_counts[pick], g_scores[pick], 2)
                    entry["parity"] = {"checked": n, "ok": not bad, "kind": cpu.kind, "scores_bit_identical": identical,
                                       "mismatches": bad[:2], "cpu_seconds": round(dt, 1)}
                    del pts, cols, cpu
                except MemoryError as exc:
                    entry["parity"] = {"checked": 0, "ok": None, "error": repr(exc)}
            del sc
        entry["roofline"] = {"bound": "hbm", "model_candidates_per_s": round(model, 1), "peak": peak, "unit": "GB/s",
                             "achieved": round(entry["front"] * (N ** 3 + 9 * H * W) / 1e9, 1),
                             "frac": entry["front_frac_of_streaming_roofline"],
                             "note": "whole sweep call (splat + score + clear) over the streaming model G*1 B + 9 B*H*W per "
                                     "candidate, front view; 174.8 M points exceed the 27-bit packed keys, so the score pass "
                                     "gathers labels"}
        out["synthetic_1024_2048mask"] = entry
    except Exception as exc:
        out["synthetic_1024_2048mask"] = {"error": repr(exc)}
    return out


def carve_cpu_baseline():
    """The Bibi@256 pipeline (BASELINE.json configs[0]) through the reference's OWN global_carve + partwise_carve
    (baseline/_ref, single process: the carving path has no cheap CPU-parallel form), with the
    bytes of its result compared with the GPU pipeline's; the C/NumPy restatement under oracle/ only if baseline/_ref is
    absent."""
    import contextlib
    import io
    import harness
    cfg = importlib.import_module(PKG + ".utils.config")
    mu = importlib.import_module(PKG + ".utils.mask_utils")
    vc = importlib.import_module(PKG + ".utils.voxel_carving_utils")
    data = os.path.join(ROOT, "tests", "golden", "data")
    sem, sem_ext, binary = mu.load_and_prepare_masks(data, "Bibi", "front", 256, cfg.PART_COLORS_NP, cfg.INTERIOR_PARTS)
    jobs = [(["full_building"], 90), (["chhatris"], 90), (["plinth"], 90), (["front_minarets"], 90),
            (["small_minarets"], 90), (["dome"], 90)]
    sym = {"dome": 5, "chhatris": 45, "front_minarets": 5, "small_minarets": 5}
    ext_d = {"main_door": 20, "windows": 10}
    buf_ours = io.StringIO()
    with contextlib.redirect_stdout(buf_ours):
        ours = vc.partwise_carve(vc.global_carve(binary, sem_ext, 90), sem_ext, sem, cfg.PART_COLORS_NP, jobs, sym, ext_d)
    buf = io.StringIO()
    if harness.available():
        ref = harness.Reference()
        kind, rvc, colors = "reference", ref.voxel_carving_utils, ref.config.PART_COLORS_NP
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            g = rvc.global_carve(binary, sem_ext, 90)
            out = rvc.partwise_carve(g, sem_ext, sem, colors, jobs, sym, ext_d)
        dt = time.perf_counter() - t0
        sample = "Bibi@256 global_carve + partwise_carve through the reference's own functions (baseline/_ref), 1 process"
    else:
        from oracle import oracle as orc
        kind = "port"
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            g = orc.global_carve(binary, sem_ext, 90)
            out = orc.partwise_carve(g, sem_ext, sem, orc.PART_COLORS_NP, jobs, sym, ext_d)
        dt = time.perf_counter() - t0
        sample = "Bibi@256 global_carve + partwise_carve through oracle/ (C restatement of scipy affine/label + NumPy)"
    same = bool(np.array_equal(np.asarray(out), ours))
    res = {"wall_ms": round(dt * 1e3, 1), "cores": 1, "kind": kind, "sample": sample,
           "parity": {"ok": same, "what": "bytes of the carved (D,H,W,3) grid against the GPU pipeline's",
                      "log_identical": (buf.getvalue() == buf_ours.getvalue()) if kind == "reference" else None}}
    return res


# ------------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference functions (baseline/_ref, installed by tools/install_ref.py) on the host cores,
# candidate-parallel over all of them with one BLAS thread per process, bounded sample per step.  Imports nothing of the
# product package.
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import cpu_arm
    N, H, W, B = args.grid, args.mask, args.mask, args.cands_per_step
    parts = bw.PART_NAMES if args.parts == "all" else ["front_minarets", "back_minarets"]
    labels = bw.monument_labels(N)
    base = bw.base_camera(N, H, W, "front")
    hidden = base + HIDDEN_DELTA
    pts_all, cols_all = bw.points_of(labels, bw.PART_NAMES)
    probe = cpu_arm.CpuScorer(pts_all, cols_all, np.zeros((H, W, 3), np.uint8), bw.PART_NAMES, part_colors=bw.PART_COLORS)
    gt = probe.render(pts_all, cols_all, hidden)              # ground truth = the reference's own render of the hidden camera
    pts, cols = (pts_all, cols_all) if args.parts == "all" else bw.points_of(labels, parts)
    del labels
    cpu = cpu_arm.CpuScorer(pts, cols, gt, parts, part_colors=bw.PART_COLORS)
    cand_all = bw.candidates(base, TOTAL_CANDIDATES)
    procs = cpu_arm.host_threads()
    per_step = args.cpu_sample if args.cpu_sample > 0 else max(procs, int(round(procs * 2.2e7 * 1.5 / max(len(pts), 1))))
    for s in range(args.warmup):
        cpu.run(cand_all[:procs], processes=procs)
    total = 0.0
    score0 = None
    for s in range(args.steps):
        blk = cand_all[(s * per_step) % TOTAL_CANDIDATES:][:per_step]
        dt, _, scores = cpu.run(blk, processes=procs)
        total += dt
        score0 = float(scores[0]) if score0 is None else score0
    value = per_step * args.steps / total
    sample = (f"{per_step} candidates per step (bounded sample of the {TOTAL_CANDIDATES}-candidate sweep) on the same "
              f"{N}^3 grid / {H}x{W} mask, candidate-parallel over {procs} host processes, one BLAS thread each")
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bw.bench_config(N, H, W, parts, len(pts), B),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": procs, "kind": cpu.kind, "sample": sample,
                         "blas_env": cpu_arm.blas_env()},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "score0": score0,
        "product_library_loaded": any("libp3d_b200" in l for l in open("/proc/self/maps")) if os.path.exists("/proc/self/maps") else None,
    }))


def _emit(line):
    """The one JSON line goes to the real stdout; everything else a library prints there (NCCL's version banner, ...)
    was re-routed to stderr by _quiet_stdout()."""
    _REAL_STDOUT.write(line + "\n")
    _REAL_STDOUT.flush()


def _quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


if __name__ == "__main__":
    _quiet_stdout()
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
