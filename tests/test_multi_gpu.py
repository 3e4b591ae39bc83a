"""Several GPUs behind one C call (hrt_multi_*, include/hrt_cuda.h; SURVEY
section 8 rows (b) and (e)): the job sharded inside the
library must give the single-device results bit for bit, dense outputs without
any collective, summaries / path lists through ncclAllGather for
device-resident consumers.  Tests that need two devices skip on a 1-GPU box."""
import json
import os
import subprocess

import numpy as np
import pytest

import hrt_testlib as tl
import hrt_b200 as hrt
from hrt_b200 import abi

pytestmark = pytest.mark.gpu


def _inputs():
    scene = "simple_street_canyon_with_cars"
    rx, tx = tl.canyon_c4_positions()
    rx, tx = rx[:12], tx[:2]
    rng = np.random.default_rng(3)
    return scene, rx, tx, rng.uniform(-3, 3, rx.shape), rng.uniform(-10, 10, tx.shape), 3.5


def _devices(n):
    if hrt.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")
    return list(range(n))


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_run_equals_single_device(ndev):
    """hrt_multi_run: dense arrays, RaysInfo, summaries, impulse response and path
    list of a job dealt over `ndev` devices == the same job on one device."""
    devs = _devices(ndev)
    scene, rx, tx, rxv, txv, f = _inputs()
    P, B = 300_000, 4
    with hrt.Context(0) as one:
        one.load_scene(tl.scene_path(scene))
        ref = one.run(rx, tx, rxv, txv, f, P, B, dense=True, raysinfo=True, summary=True, cir=(20e-9, 5e-9, 128))
        n_valid = int(ref["pair"]["n_valid"].sum())
        ref_list = one.run(rx, tx, rxv, txv, f, 40_000, B, path_list=4_000_000)
    with hrt.MultiContext(devs) as m:
        assert m.num_devices == ndev
        m.load_scene(tl.scene_path(scene))
        got = m.run(rx, tx, rxv, txv, f, P, B, dense=True, raysinfo=True, summary=True, cir=(20e-9, 5e-9, 128), shard_block=1 << 16)
        for k in abi.CHAN_FIELDS:
            if k == "directions_tx":
                continue
            assert np.array_equal(ref["out"].scat[k].view(np.uint32), got["out"].scat[k].view(np.uint32)), k
            assert np.array_equal(ref["out"].los[k].view(np.uint32), got["out"].los[k].view(np.uint32)), "los." + k
        assert np.array_equal(ref["out"].scat_rays.view(np.uint32), got["out"].scat_rays.view(np.uint32))
        assert np.array_equal(ref["out"].scat_active, got["out"].scat_active)
        for k in ("n_valid", "n_occluded", "hit_hash", "tau_bits"):
            assert np.array_equal(ref["pair"][k], got["pair"][k]), k
        for k in ("n_traced", "n_hit", "hit_hash", "t_bits"):
            assert np.array_equal(ref["bounce"][k], got["bounce"][k]), k
        np.testing.assert_allclose(got["pair"]["power_te"], ref["pair"]["power_te"], rtol=1e-9)
        np.testing.assert_allclose(got["cir"], ref["cir"], rtol=1e-4, atol=1e-12)
        assert got["stats"]["ray_bounces"] == ref["stats"]["ray_bounces"] and n_valid > 1_000_000
        lst = m.run(rx, tx, rxv, txv, f, 40_000, B, path_list=4_000_000, shard_block=4096)
        assert lst["paths_found"] == ref_list["paths_found"] == len(lst["paths"])
        key = lambda q: np.lexsort((q["path"], q["bounce"], q["tx"], q["rx"]))
        a, b = ref_list["paths"][key(ref_list["paths"])], lst["paths"][key(lst["paths"])]
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("ndev", [1, 2])
def test_gathered_results_on_every_device(ndev):
    """hrt_multi_run_gathered: after ncclAllGather + local reduction EVERY device
    holds the whole job's summary tables, and every device's path records."""
    import torch
    devs = _devices(ndev)
    scene, rx, tx, rxv, txv, f = _inputs()
    P, B = 60_000, 4
    R, T = len(rx), len(tx)
    with hrt.Context(0) as one:
        one.load_scene(tl.scene_path(scene))
        ref = one.run(rx, tx, rxv, txv, f, P, B, summary=True, los=False)
    n_valid = int(ref["pair"]["n_valid"].sum())
    cap = n_valid // ndev + 200_000
    pair = [torch.full((R * T * B * 6,), -1, dtype=torch.int64, device=f"cuda:{d}") for d in devs]
    bounce = [torch.full((T * B * 4,), -1, dtype=torch.int64, device=f"cuda:{d}") for d in devs]
    paths = [torch.zeros(ndev * cap * 48, dtype=torch.uint8, device=f"cuda:{d}") for d in devs]
    with hrt.MultiContext(devs) as m:
        m.load_scene(tl.scene_path(scene))
        counts = m.run_gathered(rx, tx, rxv, txv, f, P, B, [t.data_ptr() for t in pair], [t.data_ptr() for t in bounce],
                                [t.data_ptr() for t in paths], cap, shard_block=4096)
        assert m.nccl_version() > 20000
    assert sum(counts) == n_valid and all(c <= cap for c in counts)
    for i, d in enumerate(devs):
        torch.cuda.synchronize(d)
        pr = pair[i].cpu().numpy().view(hrt.PAIR_DTYPE).reshape(R, T, B)
        br = bounce[i].cpu().numpy().view(hrt.BOUNCE_DTYPE).reshape(T, B)
        for k in ("n_valid", "n_occluded", "hit_hash", "tau_bits"):
            assert np.array_equal(ref["pair"][k], pr[k]), (d, k)
        for k in ("n_traced", "n_hit", "hit_hash", "t_bits"):
            assert np.array_equal(ref["bounce"][k], br[k]), (d, k)
        np.testing.assert_allclose(pr["power_te"], ref["pair"]["power_te"], rtol=1e-9)
        # every device holds every device's records
        rec = paths[i].cpu().numpy().view(hrt.PATH_DTYPE).reshape(ndev, cap)
        allr = np.concatenate([rec[j, :counts[j]] for j in range(ndev)])
        assert len(allr) == n_valid
        keys = (allr["rx"].astype(np.int64) * T + allr["tx"]) * B * P + allr["bounce"].astype(np.int64) * P + allr["path"]
        assert np.unique(keys).size == n_valid
        for r in range(R):
            assert int((allr["rx"] == r).sum()) == int(ref["pair"]["n_valid"][r].sum())


@pytest.mark.parametrize("devices", ["0", "0,1"])
def test_c_caller_bit_identical_across_device_counts(tmp_path, devices):
    """tests/c_caller (a plain C program calling compute_paths()) with
    HRT_DEVICES=0 and HRT_DEVICES=0,1: identical hash over every dense output
    array, RaysInfo included; two calls on the same scene (cached upload) too."""
    if devices == "0,1":
        _devices(2)
    exe = tl.build_c_caller(str(tmp_path))
    base = None
    for env_dev, calls in (("0", "1"), (devices, "1"), (devices, "2")):
        env = dict(os.environ, HRT_DEVICES=env_dev)
        p = subprocess.run([exe, tl.scene_path("box"), "200000", "3", "3.0", calls], capture_output=True, text=True, timeout=300, env=env)
        assert p.returncode == 0, p.stderr
        got = json.loads(p.stdout.strip().splitlines()[-1])
        base = base or got
        assert got == base, (env_dev, calls)
    assert base["paths"] > 100_000
