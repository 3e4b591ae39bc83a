#!/usr/bin/env python
"""bench.py -- ray-bounces/s of the compute_paths() hot path on B200.

Workload (BASELINE.json configs[3], the one the metric is quoted on):
scenes/simple_street_canyon_with_cars.hrt, 4 TX / 64 RX, 1e8 rays (2.5e7 per
TX), 5 bounces, 3.5 GHz; geometry of SURVEY section 8d.  One "step" = the whole
job: ray generation, 5 wavefront depths (closest hit + Fresnel + reflection,
then 64 shadow queries + scattering per hit), per-(rx,tx,bounce) reductions
(summary mode -- the dense [R][T][B][P] output of this config would be 6.7 TB).

  value : ray-bounces/s, scene + BVH already resident in HBM (CUDA events on the
          launching stream, barrier + synchronize on both sides, max over ranks)
  e2e   : same metric through the C ABI with HOST buffers: host Scene struct ->
          hrt_scene_upload (H2D + GPU BVH build) -> hrt_run -> summaries D2H,
          every step, wall clock
  N > 1 : the path range is dealt in 2^20-path blocks to the ranks (strong
          scaling: the job stays 1e8 rays), scene replicated, per-rank summary
          tables all-gathered over NCCL and reduced on rank 0.

`--impl reference` times the UNMODIFIED reference (oracle/_ref, compiled from
/root/reference/src by oracle/Makefile) on the host CPU cores on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SCENE = os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt")
F_GHZ, NUM_TX, NUM_RX, BOUNCES = 3.5, 4, 64, 5
TOTAL_RAYS = int(float(os.environ.get("HRT_BENCH_RAYS", "1e8")))
SHARD_BLOCK = int(os.environ.get("HRT_SHARD_BLOCK", 1 << 16))   # paths per block dealt round-robin to the ranks
METRIC = "ray-bounces/s on street_canyon_with_cars at 1/2/4/8 B200 vs host-CPU C path"

# algorithmic flops per unit of work (edges are stored, so
# stage A is 14): ray-box test 22; Moeller-Trumbore stage A 14, B +9, C +16, D +6
FLOPS_BOX, FLOPS_STAGE = 22, (14, 9, 16, 6)


def c4_positions():
    tx = [[-45.0 + 30.0 * i, 0.0, 10.0] for i in range(NUM_TX)]
    rx = [[-60.0 + 8.0 * j, y, 1.5] for y in (-3.0, -1.0, 1.0, 3.0) for j in range(16)]
    return np.asarray(rx, np.float32), np.asarray(tx, np.float32)


def work_flops(w):
    return FLOPS_BOX * w[0] + sum(f * n for f, n in zip(FLOPS_STAGE, w[1:5]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm

def ref_harness_path():
    return os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def run_reference_sample(paths_per_tx, parallel):
    """The unmodified reference on a bounded sample of the workload: one
    single-TX process per transmitter (TXs are independent; the reference is
    single-threaded), run in parallel or one after the other.
    Returns (ray_bounces, seconds, cores)."""
    rx, tx = c4_positions()
    exe = ref_harness_path()
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], capture_output=True)
    if not os.path.exists(exe):
        return None
    cmds = []
    # all host cores: every TX's rays go to `split` processes (the reference has no
    # first-ray argument, so each process traces its own Fibonacci lattice of
    # paths_per_tx / split rays: same scene, TX, RX and bounce statistics)
    split = max(1, min((os.cpu_count() or 1) // NUM_TX, 32)) if parallel else 1
    split = int(os.environ.get("HRT_REF_SPLIT", split))
    for t in range(NUM_TX):
        for _ in range(split):
            cmds.append([exe, SCENE, str(F_GHZ), str(max(paths_per_tx // split, 1)), str(BOUNCES), str(NUM_RX), "1"]
                        + [repr(float(v)) for v in rx.reshape(-1)] + [repr(float(v)) for v in tx[t]])
    t0 = time.perf_counter()
    outs = []
    if parallel:
        procs = [subprocess.Popen(c, stdout=subprocess.PIPE, text=True) for c in cmds]
        outs = [p.communicate()[0] for p in procs]
    else:
        outs = [subprocess.run(c, capture_output=True, text=True).stdout for c in cmds]
    wall = time.perf_counter() - t0
    recs = [json.loads(o) for o in outs]
    rb = sum(r["ray_bounces"] for r in recs)
    cpu_s = sum(r["seconds"] for r in recs)
    if parallel:
        return rb, max(r["seconds"] for r in recs), min(len(cmds), os.cpu_count() or 1), wall
    return rb, cpu_s, 1, wall


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    P = int(os.environ.get("HRT_REF_PATHS", str(2500 * max(1, min((os.cpu_count() or 1) // NUM_TX, 32)))))
    for _ in range(args.warmup):
        run_reference_sample(max(P // 10, 100), True)
    tot_rb, tot_s = 0, 0.0
    for _ in range(args.steps):
        r = run_reference_sample(P, True)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness missing"}))
            return 0
        tot_rb += r[0]; tot_s += r[1]
        cores = r[2]
    v = tot_rb / tot_s
    sample = (f"canyon 4 TX / 64 RX / 5 bounces, {P} rays per TX per step, {cores} single-TX processes of the "
              f"unmodified (single-threaded) reference in parallel, one per host core, each tracing its share of a TX's rays")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "ray-bounces/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": v, "unit": "ray-bounces/s", "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": v, "unit": "ray-bounces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def workload_config():
    return {"workload": "BASELINE configs[3]: scenes/simple_street_canyon_with_cars.hrt (234 triangles), "
                        f"{NUM_TX} TX / {NUM_RX} RX, {TOTAL_RAYS:.0e} rays ({TOTAL_RAYS // NUM_TX} per TX), "
                        f"{BOUNCES} bounces, {F_GHZ} GHz, summary (streaming) outputs",
            "rays": TOTAL_RAYS, "num_tx": NUM_TX, "num_rx": NUM_RX, "bounces": BOUNCES,
            "shard_block": SHARD_BLOCK,
            "l2": "inputs larger than L2: per-ray state of a step is %.1f GB" % (TOTAL_RAYS * 148 / 1e9)}


# ------------------------------------------------------------------ GPU arm

def main_gpu(args):
    import torch
    import torch.distributed as dist
    import hrt_b200 as hrt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON); anything libraries print (NCCL's
    # version banner ...) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rx, tx = c4_positions()
    zr, zt = np.zeros_like(rx), np.zeros_like(tx)
    P = TOTAL_RAYS // NUM_TX
    ctx = hrt.Context(local)
    ctx.load_scene(SCENE)
    R, T, B = NUM_RX, NUM_TX, BOUNCES

    # per-rank summary tables live in torch device memory so that NCCL can gather them
    pair_dev = torch.zeros(R * T * B * 6, dtype=torch.int64, device="cuda")
    bounce_dev = torch.zeros(T * B * 4, dtype=torch.int64, device="cuda")
    gathered = torch.zeros(world * pair_dev.numel(), dtype=torch.int64, device="cuda") if world > 1 else None
    gathered_b = torch.zeros(world * bounce_dev.numel(), dtype=torch.int64, device="cuda") if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        pair_dev.zero_(); bounce_dev.zero_()
        r = ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=(rank == 0),
                    shard=(rank, world), shard_block=SHARD_BLOCK,
                    summary_dev_ptrs=(pair_dev.data_ptr(), bounce_dev.data_ptr()), stream=stream)
        if world > 1:
            dist.all_gather_into_tensor(gathered, pair_dev)
            dist.all_gather_into_tensor(gathered_b, bounce_dev)
        return r["stats"]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    stats = [step() for _ in range(args.steps)]
    e1.record()
    sync()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    rb_local = torch.tensor([float(sum(s["ray_bounces"] for s in stats)),
                             float(sum(s["shadow_queries"] for s in stats)),
                             float(sum(s["kernel_launches"] for s in stats))],
                            device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rb_local, op=dist.ReduceOp.SUM)
    ms_total = float(ms.item())
    rb_total, shadow_total, launches = (float(x) for x in rb_local.tolist())

    # result sanity: reduce the gathered tables, cross-check against the counters
    if world > 1:
        pairs = gathered.view(world, R * T * B, 6).cpu().numpy()
        n_valid = int(pairs[:, :, 0].sum())
    else:
        n_valid = int(pair_dev.view(R * T * B, 6)[:, 0].sum().item())

    # ---- e2e: host buffers through the C ABI, scene upload + BVH build + D2H in the loop
    scene_bytes = os.path.getsize(SCENE)
    sync()
    t0 = time.perf_counter()
    rb_e2e = 0
    for _ in range(args.steps):
        ctx.load_scene(SCENE)                         # host Scene -> H2D -> GPU BVH build
        r = ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=(rank == 0),
                    shard=(rank, world), shard_block=SHARD_BLOCK)      # host summary arrays (D2H)
        rb_e2e += r["stats"]["ray_bounces"]
    sync()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    rb_e = torch.tensor([float(rb_e2e)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(rb_e, op=dist.ReduceOp.SUM)
    h2d = scene_bytes + (R + T) * 24
    d2h = R * T * B * 48 + T * B * 32 + (B + 1) * T * 4

    # ---- per-path results: every rank lists the valid paths of its shard of a
    # smaller job on the GPU (ballot/prefix compaction), then one NCCL all-gather of
    # the counts and one of the record buffers over NVLink: every GPU ends up with
    # every path of the job
    PL = int(os.environ.get("HRT_BENCH_LIST_PATHS", "40000"))          # rays per TX of the listed job
    cap = int(PL * T * B * R * 0.6 / world) + 4096                      # records per rank (C4: ~0.42 valid per slot)
    lbuf = torch.empty(cap * 48, dtype=torch.uint8, device="cuda")
    g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    sync()
    g0.record()
    lr = ctx.run(rx, tx, zr, zt, F_GHZ, PL, B, los=False, shard=(rank, world), shard_block=4096,
                 path_list_dev=(lbuf.data_ptr(), cap), stream=stream)
    g1.record()
    allrec, counts = hrt.gather_path_lists(lbuf, min(lr["paths_found"], cap), cap)
    g2.record()
    sync()
    gather = {"job": f"{PL} rays per TX, same scene/TX/RX/bounces", "records": int(allrec.shape[0]),
              "record_bytes": 48, "bytes_gathered_per_rank": int(world * cap * 48),
              "overflow": bool(lr["paths_found"] > cap), "trace_and_list_ms": g0.elapsed_time(g1),
              "all_gather_ms": g1.elapsed_time(g2), "backend": "nccl" if world > 1 else "single rank"}

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (k_scatter): counted work / live launch time
        cnt = ctx.run(rx, tx, zr, zt, F_GHZ, min(P, 1 << 20), B, summary=True, los=False,
                      count_work=True)["stats"]
        flops_per_shadow = work_flops(cnt["work_scatter"]) / max(cnt["shadow_queries"], 1)
        flops_per_primary = work_flops(cnt["work_bounce"]) / max(cnt["ray_bounces"], 1)
        s0 = stats[0]
        ms_scatter = sum(s["ms_scatter"] for s in stats)
        n_scatter = sum(s["n_scatter_launches"] for s in stats)
        shadow_rank0 = sum(s["shadow_queries"] for s in stats)
        peak_unfused, peak_fma = ctx.fp32_peak()
        achieved = flops_per_shadow * shadow_rank0 / (ms_scatter * 1e-3) / 1e12 if ms_scatter else None
        nominal = 148 * 128 * 1.965e9 / 1e12
        traffic, traffic_note = None, "no ncu capture committed under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k_scatter_traffic.json")))
            l0 = tj["launches"][0]
            traffic = l0["dram_bytes_read"] + l0["dram_bytes_write"]
            traffic_note = ("DRAM bytes of ONE k_scatter launch under ncu (%s, %.1f ms launch): the kernel is fp32-issue "
                            "bound; its DRAM traffic is the gather of per-hit state" % (tj["source"], l0["duration_ms"]))
        except Exception:
            pass
        isect_share, isect_note = None, None
        try:
            sj = json.load(open(os.path.join(ROOT, "profiles", "k_scatter_shares.json")))
            isect_share = sj["intersection_share_of_issue_slots"]
            isect_note = ("k_scatter fuses the reference's per-path scattering math (0 counted flops) with the closest-hit "
                          "query; by ncu source-level instruction counts (%s) %.0f %% of its issue slots are the query. "
                          "frac_of_intersection_slots = frac / that share" % (sj["source"], 100 * isect_share))
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "kernel": "k_scatter", "achieved": achieved, "peak": peak_unfused,
            "unit": "TFLOP/s", "frac": (achieved / peak_unfused) if achieved and peak_unfused else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "intersection_share_of_issue_slots": isect_share,
            "frac_of_intersection_slots": (achieved / peak_unfused / isect_share) if achieved and peak_unfused and isect_share else None,
            "intersection_note": isect_note,
            "peak_source": "measured in this run with separately rounded FMUL+FADD chains "
                           "(hrt_fp32_peak); MEASURED_PEAKS.json has no fp32 figure. nominal "
                           f"unfused {nominal:.1f}, measured FFMA {peak_fma:.1f} TFLOP/s",
            "flops_per_shadow_query": flops_per_shadow, "flops_per_primary_query": flops_per_primary,
            "launches": n_scatter, "avg_launch_ms": ms_scatter / n_scatter if n_scatter else None,
            "kernel_share_of_step": ms_scatter / (ms_total if world == 1 else sum(s["ms_total"] for s in stats)),
            "box_tests_per_shadow_query": cnt["work_scatter"][0] / max(cnt["shadow_queries"], 1),
            "tri_tests_per_shadow_query": cnt["work_scatter"][1] / max(cnt["shadow_queries"], 1),
        }
        # ---- CPU baseline: the unmodified reference on a bounded sample, one core
        cpu = None
        if world == 1:
            r = run_reference_sample(int(os.environ.get("HRT_REF_PATHS", "1000")), False)
            if r:
                cpu = {"value": r[0] / r[1], "unit": "ray-bounces/s", "cores": 1, "kind": "reference",
                       "sample": f"same scene/TX/RX/bounces, {os.environ.get('HRT_REF_PATHS', '1000')} rays per TX "
                                 f"({r[0]} ray-bounces, {r[1]:.1f} s of compute_paths())"}
        value = rb_total / (ms_total * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": "ray-bounces/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(),
            "clocks": clocks,
            "e2e": {"value": float(rb_e.item()) / float(e2e_s.item()), "unit": "ray-bounces/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu,
            "path_list_gather": gather,
            "closest_hit_queries_per_s": (rb_total + shadow_total) / (ms_total * 1e-3),
            "shadow_queries_per_step": shadow_total / args.steps,
            "ray_bounces_per_step": rb_total / args.steps,
            "valid_paths_last_step": n_valid,
            "ambiguous_dirs_per_step": s0["ambiguous_dirs"],
            "bvh": {"triangles": s0["num_tris"], "nodes": s0["num_nodes"], "builder": "binned SAH" if s0["bvh_sah"] else "LBVH",
                    "build_ms": s0["bvh_build_ms"], "scene_in_smem": bool(s0["scene_in_smem"]), "box_pad_m": s0["box_pad"]},
            "ms_breakdown_rank0_last_step": {"scatter": stats[-1]["ms_scatter"], "bounce": stats[-1]["ms_bounce"],
                                             "hit_sort": stats[-1]["ms_sort"], "total": stats[-1]["ms_total"]},
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
