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
    # per step: 10,000 rays per process (~2 s of the reference each), one process per host core
    P = int(os.environ.get("HRT_REF_PATHS", str(10000 * max(1, min((os.cpu_count() or 1) // NUM_TX, 32)))))
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

def dense_e2e_block(hrt, abi, tl):
    """The drop-in C symbol compute_paths() itself (host buffers in, host buffers
    out: scene upload, BVH build, trace, D2H of every dense array inside) on
    BASELINE configs[1] and configs[2] at their stated sizes, each checked against
    the oracle in this run (style of the reference's test/test.c:62-72)."""
    out = []
    L = hrt.lib()
    for name, cfg, P, B in (("configs[1] box.hrt 1e6 rays x 3 bounces", "box_generic", 1_000_000, 3),
                            ("configs[2] 2cars.hrt 1e7 rays x 5 bounces, 70 GHz", "2cars_raised", 10_000_000, 5)):
        scene, rx, tx, f = tl.CONFIGS[cfg]
        zr, zt = [[0.0, 0.0, 0.0]], [[0.0, 0.0, 0.0]]
        sc = L.scene_load(tl.scene_path(scene).encode())
        o = abi.alloc_outputs(1, 1, P, B, 0)
        times = []
        for rep in range(3):
            t0 = time.perf_counter()
            abi.call_compute_paths(L, sc, rx, tx, zr, zt, f, P, B, out=o)
            times.append(time.perf_counter() - t0)
        abi.free_scene(sc)
        t_o = time.perf_counter()
        a, tr = tl.run_oracle(scene, rx, tx, zr, zt, f, P, B, fill=0)
        t_o = time.perf_counter() - t_o
        st = tr["slot_state"].reshape(-1)
        wr, va = st != 0, st == 1
        tau_ok = bool(np.array_equal(a.scat["tau"].reshape(-1).view(np.uint32)[wr], o.scat["tau"].reshape(-1).view(np.uint32)[wr]))
        dir_ok = bool(np.array_equal(a.scat["directions_rx"].reshape(-1, 3).view(np.uint32)[va],
                                     o.scat["directions_rx"].reshape(-1, 3).view(np.uint32)[va]))
        worst = 0.0
        for pol in ("te", "tm"):
            ar = a.scat[f"a_{pol}_re"].reshape(-1)[wr].astype(np.float64); ai = a.scat[f"a_{pol}_im"].reshape(-1)[wr].astype(np.float64)
            br = o.scat[f"a_{pol}_re"].reshape(-1)[wr].astype(np.float64); bi = o.scat[f"a_{pol}_im"].reshape(-1)[wr].astype(np.float64)
            mag = np.hypot(ar, ai); err = np.hypot(ar - br, ai - bi)
            nz = mag > 0
            if nz.any():
                worst = max(worst, float((err[nz] / mag[nz]).max()))
            if (~nz).any():
                worst = max(worst, float(err[~nz].max() > 0))
        rb = int((tr["hit_tri"] != tl.IDLE).sum())
        d2h = sum(v.nbytes for k, v in o.scat.items() if k != "directions_tx") + o.scat_rays.nbytes + o.scat_active.nbytes
        best = min(times[1:])
        out.append({"config": name, "entry": "compute_paths() (C symbol, host arrays)", "seconds": best, "first_call_seconds": times[0],
                    "ray_bounces": rb, "ray_bounces_per_s": rb / best, "d2h_bytes": int(d2h),
                    "oracle_check": {"slots_written": int(wr.sum()), "valid_paths": int(va.sum()), "tau_bit_exact": tau_ok,
                                     "directions_bit_exact": dir_ok, "worst_gain_rel_err": worst, "tolerance": tl.GAIN_RTOL,
                                     "oracle_seconds_1_core": t_o}})
        del a, tr, o
    return out


def c5_block(hrt, rank, world, local, peak_unfused):
    """BASELINE configs[4] (C5): the synthetic 64 x 64 tiled canyon (958,464
    triangles, mixed ITU materials), 16 TX / 1024 RX / 6 bounces.  The full job is
    3.4e12 closest-hit queries, so every rank traces ONE shard of 256 of it (shard
    `rank`: 65,536-path blocks dealt round-robin, i.e. the ray density and
    coherence of the full job) -- weak scaling in the number of GPUs."""
    from hrt_b200 import scenes
    meshes, pitch = scenes.tiled_canyon(SCENE, 64, 64)
    path = f"/tmp/c5_tiled_canyon_{rank}.hrt"
    scenes.write_hrt(path, meshes)
    rx, tx = scenes.c5_positions(pitch, 64, 64, n_tx=16, n_rx=1024)
    zr, zt = np.zeros_like(rx), np.zeros_like(tx)
    ctx = hrt.Context(local)
    t0 = time.perf_counter(); ctx.load_scene(path); t_up = time.perf_counter() - t0
    P, SH = 62_500_000, 256
    os.environ["HRT_NO_OVERLAP"] = "1"      # per-kernel intervals (launches of 100+ ms: the depth pipeline has nothing to hide here)
    try:
        ctx.run(rx, tx, zr, zt, 3.5, 200_000, 6, summary=True, los=False)          # warm-up (buffers, clocks)
        r = ctx.run(rx, tx, zr, zt, 3.5, P, 6, summary=True, shard=(rank % SH, SH), shard_block=1 << 16)
        s = r["stats"]
        c = ctx.run(rx, tx, zr, zt, 3.5, 20_000, 6, summary=True, count_work=True, los=False)["stats"]
    finally:
        del os.environ["HRT_NO_OVERLAP"]
    ctx.close()
    os.remove(path)
    fq = work_flops(c["work_scatter"]) / max(c["shadow_queries"], 1)
    return {"ms": s["ms_total"], "ms_scatter": s["ms_scatter"], "ray_bounces": s["ray_bounces"], "shadow_queries": s["shadow_queries"],
            "valid_paths": int(r["pair"]["n_valid"].sum()), "flops_per_shadow_query": fq,
            "box_tests_per_shadow_query": c["work_scatter"][0] / max(c["shadow_queries"], 1),
            "tri_tests_per_shadow_query": c["work_scatter"][1] / max(c["shadow_queries"], 1),
            "counted_on_rays_per_tx": 20_000, "triangles": s["num_tris"], "bvh_build_ms": s["bvh_build_ms"],
            "scene_load_upload_bvh_s": t_up, "scene_in_smem": bool(s["scene_in_smem"]),
            "achieved_tflops": fq * s["shadow_queries"] / (s["ms_scatter"] * 1e-3) / 1e12 if s["ms_scatter"] else None,
            "peak": peak_unfused}


def main_gpu(args):
    import torch
    import torch.distributed as dist
    import hrt_b200 as hrt
    from hrt_b200 import abi

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
    # host-side barrier for phases in which only rank 0 computes: an NCCL barrier would leave a
    # spinning kernel on the idle ranks' GPUs, time-sliced against rank 0's work on them
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    rx, tx = c4_positions()
    zr, zt = np.zeros_like(rx), np.zeros_like(tx)
    P = TOTAL_RAYS // NUM_TX
    ctx = hrt.Context(local)
    ctx.load_scene(SCENE)
    R, T, B = NUM_RX, NUM_TX, BOUNCES

    # ONE explicit stream carries everything of a step: the zeroing of the tables,
    # every kernel of hrt_run (p.stream), the NCCL all-gathers and the timing events
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        # per-rank summary tables live in torch device memory so that NCCL can gather them
        pair_dev = torch.zeros(R * T * B * 6, dtype=torch.int64, device="cuda")
        bounce_dev = torch.zeros(T * B * 4, dtype=torch.int64, device="cuda")
        gathered = torch.zeros(world * pair_dev.numel(), dtype=torch.int64, device="cuda") if world > 1 else None
        gathered_b = torch.zeros(world * bounce_dev.numel(), dtype=torch.int64, device="cuda") if world > 1 else None

    def step():
        with torch.cuda.stream(stream):
            pair_dev.zero_(); bounce_dev.zero_()
            r = ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=(rank == 0),
                        shard=(rank, world), shard_block=SHARD_BLOCK,
                        summary_dev_ptrs=(pair_dev.data_ptr(), bounce_dev.data_ptr()), stream=stream.cuda_stream)
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
    e0.record(stream)
    stats = [step() for _ in range(args.steps)]
    e1.record(stream)
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
    with torch.cuda.stream(stream):
        lbuf = torch.empty(cap * 48, dtype=torch.uint8, device="cuda")
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        sync()
        g0.record(stream)
        lr = ctx.run(rx, tx, zr, zt, F_GHZ, PL, B, los=False, shard=(rank, world), shard_block=4096,
                     path_list_dev=(lbuf.data_ptr(), cap), stream=stream.cuda_stream)
        g1.record(stream)
        allrec, counts = hrt.gather_path_lists(lbuf, min(lr["paths_found"], cap), cap)
        g2.record(stream)
    sync()
    gather = {"job": f"{PL} rays per TX, same scene/TX/RX/bounces", "records": int(allrec.shape[0]),
              "record_bytes": 48, "bytes_gathered_per_rank": int(world * cap * 48),
              "overflow": bool(lr["paths_found"] > cap), "trace_and_list_ms": g0.elapsed_time(g1),
              "all_gather_ms": g1.elapsed_time(g2), "backend": "nccl" if world > 1 else "single rank"}
    del lbuf, allrec

    peak_unfused, peak_fma = ctx.fp32_peak()

    # ---- BASELINE configs[4]: one shard of 256 of the C5 job per rank (weak scaling)
    c5 = None
    if not os.environ.get("HRT_BENCH_SKIP_C5"):
        mine = c5_block(hrt, rank, world, local, peak_unfused)
        vec = torch.tensor([mine["ms"], float(mine["ray_bounces"]), float(mine["shadow_queries"]), mine["ms_scatter"]],
                           device="cuda", dtype=torch.float64)
        mx, sm = vec.clone(), vec.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        t_s = float(mx[0]) * 1e-3
        c5 = {"workload": "BASELINE configs[4]: synthetic 64x64 tiled street canyon (958,464 triangles, mixed ITU materials), "
                          "16 TX / 1024 RX, 6 bounces, 6.25e7 rays per TX; every rank traces one shard of 256 of the job "
                          "(blocks of 65,536 paths dealt round-robin: the full job's ray density)",
              "scaling": "weak", "n_gpus": world, "shards_traced": world, "of_shards": 256,
              "seconds": t_s, "ray_bounces_per_s": float(sm[1]) / t_s,
              "closest_hit_queries_per_s": (float(sm[1]) + float(sm[2])) / t_s,
              "full_job_seconds_at_this_rate": 256.0 / world * t_s,
              "full_job_note": "a shard has 1/256 of the job's hit density, so this extrapolation is pessimistic: the full "
                               "1e9-ray job was run on 8 GPUs (scripts/run_c5_full.py): 14.65 s, profiles/r2_v3/c5_full.json",
              "target_rb_per_s_north_star": 1e10,
              "roofline": {"bound": "fp32", "kernel": "k_scatter (global-memory scene, 4-wide BVH)",
                           "achieved": mine["achieved_tflops"], "peak": peak_unfused, "unit": "TFLOP/s",
                           "frac": mine["achieved_tflops"] / peak_unfused if mine["achieved_tflops"] else None,
                           "flops_per_shadow_query": mine["flops_per_shadow_query"],
                           "box_tests_per_shadow_query": mine["box_tests_per_shadow_query"],
                           "tri_tests_per_shadow_query": mine["tri_tests_per_shadow_query"],
                           "counted_on_rays_per_tx": mine["counted_on_rays_per_tx"]},
              "rank0": {k: mine[k] for k in ("ms", "ms_scatter", "valid_paths", "triangles", "bvh_build_ms",
                                             "scene_load_upload_bvh_s", "scene_in_smem")}}

    # ---- the same job through ONE C call that shards it over all N GPUs inside the library
    # (hrt_multi_run / hrt_multi_run_gathered: per-device host threads, ncclCommInitAll +
    # ncclAllGather in this single process).  Rank 0 drives all devices; the other ranks idle.
    inlib = None
    if not os.environ.get("HRT_BENCH_SKIP_INLIB"):
        sync()
        if rank == 0:
            try:
                devs = list(range(world))
                with hrt.MultiContext(devs) as mc:
                    mc.load_scene(SCENE)
                    mc.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, shard_block=SHARD_BLOCK)        # warm-up
                    t0 = time.perf_counter()
                    n_in = max(1, min(args.steps, 3))
                    rb_in = 0
                    for _ in range(n_in):
                        rr = mc.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, shard_block=SHARD_BLOCK)
                        rb_in += rr["stats"]["ray_bounces"]
                    t_host = (time.perf_counter() - t0) / n_in
                    pair_t = [torch.zeros(R * T * B * 6, dtype=torch.int64, device=f"cuda:{d}") for d in devs]
                    bounce_t = [torch.zeros(T * B * 4, dtype=torch.int64, device=f"cuda:{d}") for d in devs]
                    mc.run_gathered(rx, tx, zr, zt, F_GHZ, P, B, [t.data_ptr() for t in pair_t], [t.data_ptr() for t in bounce_t],
                                    shard_block=SHARD_BLOCK)
                    t0 = time.perf_counter()
                    mc.run_gathered(rx, tx, zr, zt, F_GHZ, P, B, [t.data_ptr() for t in pair_t], [t.data_ptr() for t in bounce_t],
                                    shard_block=SHARD_BLOCK)
                    t_gath = time.perf_counter() - t0
                    nv = [int(t.view(R * T * B, 6)[:, 0].sum().item()) for t in pair_t]
                    inlib = {"entry": "hrt_multi_run (host summary tables) / hrt_multi_run_gathered (tables on every GPU via ncclAllGather)",
                             "devices": world, "seconds_per_step_host_results": t_host,
                             "ray_bounces_per_s_host_results": rb_in / n_in / t_host,
                             "seconds_per_step_gathered": t_gath, "ray_bounces_per_s_gathered": rb_in / n_in / t_gath,
                             "valid_paths_on_every_device": nv, "nccl_version": mc.nccl_version()}
                    del pair_t, bounce_t
            except Exception as e:                       # never lose the headline line to the extra block
                inlib = {"error": repr(e)[:300]}
        if world > 1:
            dist.barrier(group=cpu_group)                # the other ranks wait here on the CPU, their GPUs idle
        sync()

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (k_scatter): counted work / live launch time
        P_cnt = min(P, 1 << 20)
        cnt = ctx.run(rx, tx, zr, zt, F_GHZ, P_cnt, B, summary=True, los=False, count_work=True)["stats"]
        flops_per_shadow = work_flops(cnt["work_scatter"]) / max(cnt["shadow_queries"], 1)
        flops_per_primary = work_flops(cnt["work_bounce"]) / max(cnt["ray_bounces"], 1)
        s0 = stats[0]
        ms_scatter = sum(s["ms_scatter"] for s in stats)
        n_scatter = sum(s["n_scatter_launches"] for s in stats)
        shadow_rank0 = sum(s["shadow_queries"] for s in stats)
        achieved = flops_per_shadow * shadow_rank0 / (ms_scatter * 1e-3) / 1e12 if ms_scatter else None
        nominal = 148 * 128 * 1.965e9 / 1e12
        # the same step with the shadow queries walking the BVH instead of the receiver maps
        # (rank 0's shard): more counted work per query, more time
        bvh_mode = None
        if s0["rx_map"] and not os.environ.get("HRT_BENCH_SKIP_BVH_MODE"):
            os.environ["HRT_RXMAP"] = "0"
            try:
                ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=False, shard=(rank, world), shard_block=SHARD_BLOCK)
                sb = ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=False, shard=(rank, world), shard_block=SHARD_BLOCK)["stats"]
                cb = ctx.run(rx, tx, zr, zt, F_GHZ, P_cnt, B, summary=True, los=False, count_work=True)["stats"]
            finally:
                del os.environ["HRT_RXMAP"]
            fb = work_flops(cb["work_scatter"]) / max(cb["shadow_queries"], 1)
            ab = fb * sb["shadow_queries"] / (sb["ms_scatter"] * 1e-3) / 1e12
            bvh_mode = {"how": "HRT_RXMAP=0: the same step, shadow queries through the 4-wide BVH", "ms_scatter": sb["ms_scatter"],
                        "ms_total": sb["ms_total"], "ray_bounces_per_s": sb["ray_bounces"] / (sb["ms_total"] * 1e-3),
                        "flops_per_shadow_query": fb, "box_tests_per_shadow_query": cb["work_scatter"][0] / max(cb["shadow_queries"], 1),
                        "tri_tests_per_shadow_query": cb["work_scatter"][1] / max(cb["shadow_queries"], 1),
                        "achieved": ab, "frac": ab / peak_unfused if peak_unfused else None}
        # one more step with the depth pipeline off: clean per-kernel intervals
        os.environ["HRT_NO_OVERLAP"] = "1"
        try:
            ss = ctx.run(rx, tx, zr, zt, F_GHZ, P, B, summary=True, los=False, shard=(rank, world), shard_block=SHARD_BLOCK)["stats"]
        finally:
            del os.environ["HRT_NO_OVERLAP"]
        serial_breakdown = {"how": "HRT_NO_OVERLAP=1: every kernel of the step on one stream", "scatter": ss["ms_scatter"],
                            "bounce": ss["ms_bounce"], "hit_sort": ss["ms_sort"], "total": ss["ms_total"]}
        traffic, traffic_note = None, "no ncu capture committed under profiles/"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "k_scatter_traffic.json")))
            l0 = tj["launches"][0]
            traffic = l0["dram_bytes_read"] + l0["dram_bytes_write"]
            traffic_note = ("DRAM bytes of ONE k_scatter launch under ncu (%s, %.1f ms launch): the kernel is fp32-issue "
                            "bound; its DRAM traffic is the gather of per-hit state and the receiver-map cells" % (tj["source"], l0["duration_ms"]))
        except Exception:
            pass
        isect_share, isect_note, issue_pct, lanes = None, None, None, None
        try:
            sj = json.load(open(os.path.join(ROOT, "profiles", "k_scatter_shares.json")))
            isect_share = sj["intersection_share_of_issue_slots"]
            issue_pct, lanes = sj.get("issue_active_pct"), sj.get("active_threads_per_instruction")
            isect_note = ("k_scatter fuses the reference's per-path scattering math (0 counted flops) with the closest-hit "
                          "query; by ncu source-level instruction counts (%s) %.0f %% of its issue slots are the query. "
                          "frac_of_intersection_slots = frac / that share" % (sj["source"], 100 * isect_share))
        except Exception:
            pass
        roofline = {
            "bound": "fp32", "kernel": "k_scatter", "achieved": achieved, "peak": peak_unfused,
            "unit": "TFLOP/s", "frac": (achieved / peak_unfused) if achieved and peak_unfused else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "convention": "SURVEY 8(d): flops of the tests PERFORMED (22 per ray-box test; Moeller-Trumbore 14/9/16/6 by stage reached), "
                          "counted by the instrumented kernels, x shadow queries / k_scatter time.  With receiver maps a shadow "
                          "query performs no box test at all, so the counted work per query is a third of the BVH walk's while the "
                          "kernel finishes sooner: compare bvh_mode (same step, HRT_RXMAP=0)",
            "bvh_mode": bvh_mode,
            "reference_equivalent": {
                "note": "for context (SURVEY 8(d), last line): the reference tests every triangle for every query, "
                        "~29 flops x N triangles; the same queries at that cost per second",
                "flops_per_query": 29.0 * s0["num_tris"],
                "tflops": 29.0 * s0["num_tris"] * shadow_rank0 / (ms_scatter * 1e-3) / 1e12 if ms_scatter else None},
            "issue_slots_busy_pct_ncu": issue_pct, "active_lanes_per_instruction_ncu": lanes,
            "issue_note": "the kernel is bound by instruction issue, not by a memory level or the fp32 pipe alone: share of "
                          "issue slots in use and lanes active per issued instruction, from the committed ncu capture of this "
                          "kernel (profiles/k_scatter_shares.json -> source)",
            "intersection_share_of_issue_slots": isect_share,
            "frac_of_intersection_slots": (achieved / peak_unfused / isect_share) if achieved and peak_unfused and isect_share else None,
            "intersection_note": isect_note,
            "peak_source": "measured in this run with separately rounded FMUL+FADD chains "
                           "(hrt_fp32_peak); MEASURED_PEAKS.json has no fp32 figure. nominal "
                           f"unfused {nominal:.1f}, measured FFMA {peak_fma:.1f} TFLOP/s",
            "flops_per_shadow_query": flops_per_shadow, "flops_per_primary_query": flops_per_primary,
            "counted_on_rays_per_tx": P_cnt, "timed_rays_per_tx": P,
            "counted_note": "work per query is counted on a separate pass of the instrumented kernels at counted_on_rays_per_tx "
                            "rays per TX; per-query work does not depend on the ray count",
            "launches": n_scatter, "avg_launch_ms": ms_scatter / n_scatter if n_scatter else None,
            "kernel_share_of_step": ms_scatter / (ms_total if world == 1 else sum(s["ms_total"] for s in stats)),
            "box_tests_per_shadow_query": cnt["work_scatter"][0] / max(cnt["shadow_queries"], 1),
            "tri_tests_per_shadow_query": cnt["work_scatter"][1] / max(cnt["shadow_queries"], 1),
        }
        # ---- CPU baseline: the unmodified reference on a bounded sample, one core
        cpu = None
        if world == 1:
            r = run_reference_sample(int(os.environ.get("HRT_REF_PATHS", "12000")), False)     # ~10 s on one core
            if r:
                cpu = {"value": r[0] / r[1], "unit": "ray-bounces/s", "cores": 1, "kind": "reference",
                       "sample": f"same scene/TX/RX/bounces, {os.environ.get('HRT_REF_PATHS', '12000')} rays per TX "
                                 f"({r[0]} ray-bounces, {r[1]:.1f} s of compute_paths())"}
        # ---- the drop-in entry itself on configs[1] and configs[2] (N = 1 only: host-memory heavy)
        dense = None
        if world == 1 and not os.environ.get("HRT_BENCH_SKIP_DENSE"):
            ctx.close(); ctx = None
            import hrt_testlib as tl
            dense = dense_e2e_block(hrt, abi, tl)
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
            "dense_e2e": dense, "c5": c5, "in_library_multi_gpu": inlib,
            "path_list_gather": gather,
            "closest_hit_queries_per_s": (rb_total + shadow_total) / (ms_total * 1e-3),
            "shadow_queries_per_step": shadow_total / args.steps,
            "ray_bounces_per_step": rb_total / args.steps,
            "valid_paths_last_step": n_valid,
            "ambiguous_dirs_per_step": s0["ambiguous_dirs"],
            "bvh": {"triangles": s0["num_tris"], "nodes": s0["num_nodes"], "builder": "binned SAH" if s0["bvh_sah"] else "LBVH",
                    "build_ms": s0["bvh_build_ms"], "scene_in_smem": bool(s0["scene_in_smem"]), "box_pad_m": s0["box_pad"]},
            "receiver_maps": {"used": bool(s0["rx_map"]), "cells_per_face_edge": s0["rx_map_cells"],
                              "build_ms_first_step": stats[0]["rx_map_build_ms"]},
            "ms_breakdown_rank0_last_step": {"scatter": stats[-1]["ms_scatter"], "hit_sort": stats[-1]["ms_sort"],
                                             "total": stats[-1]["ms_total"],
                                             "note": "timed steps run the depth pipeline (k_scatter of depth b beside k_bounce and the "
                                                     "sort of depth b+1); event intervals of k_bounce then include waiting for SM slots"},
            "ms_breakdown_rank0_serial_step": serial_breakdown,
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.barrier()
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
