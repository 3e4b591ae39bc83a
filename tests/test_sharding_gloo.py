"""Multi-rank host logic on CPU (gloo, world_size 2): the shard partition of
the path range and the reduction of per-rank summary tables, exactly as
bench.py does it on N GPUs (all_gather of the tables, sum on rank 0).  Per-rank
tables come from the oracle restricted to the rank's paths; their gathered sum
must equal the unsharded oracle summary bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import hrt_testlib as tl
import hrt_b200 as hrt

P, B, BLOCK = 4096, 3, 512
WORLD = 2


def _case():
    scene, rx, tx, f = tl.CONFIGS["box_generic"]
    rx = list(rx) + [[2.0, 2.0, 4.0]]
    return scene, rx, tx, [[0, 0, 0]] * 2, [[1.0, 0.0, 0.0]], f


def test_partition_covers_every_path_once():
    for P_, world, blk in [(4096, 2, 512), (5000, 3, 512), (100, 8, 32), (1 << 20, 8, 1 << 16), (33, 4, 32)]:
        seen = np.zeros(P_, np.int32)
        total = 0
        for r in range(world):
            g = hrt.shard_paths(P_, r, world, blk)
            assert g.size == hrt.lib().hrt_shard_count(P_, r, world, blk)
            seen[g.astype(np.int64)] += 1
            total += g.size
        assert total == P_ and (seen == 1).all()


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, rx, tx, rxv, txv, f = _case()
    o, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B)
    mine = hrt.shard_paths(P, rank, world, BLOCK).astype(np.int64)
    # restrict outputs + trace to this rank's paths
    sub = tl.abi.alloc_outputs(o.R, o.T, mine.size, B, 0)
    for k in ("tau", "a_te_re", "a_te_im", "a_tm_re", "a_tm_im"):
        sub.scat[k][...] = o.scat[k][..., mine]
    trs = {k: np.ascontiguousarray(v[..., mine]) for k, v in tr.items()}
    pair, bounce = tl.oracle_summaries(sub, trs, path_ids=mine)
    table = np.stack([pair[k].view(np.int64) if pair[k].dtype != np.float64 else pair[k].view(np.int64)
                      for k in ("n_valid", "n_occluded", "hit_hash", "tau_bits")], -1)
    t = torch.from_numpy(np.ascontiguousarray(table).reshape(-1))
    gathered = torch.zeros(world * t.numel(), dtype=torch.int64)
    dist.all_gather_into_tensor(gathered, t)
    power = torch.from_numpy(pair["power_te"].reshape(-1).copy())
    dist.all_reduce(power)
    if rank == 0:
        total = gathered.view(world, -1).numpy().view(np.uint64).sum(0, dtype=np.uint64)
        np.save(os.path.join(out_dir, "total.npy"), total.reshape(table.shape))
        np.save(os.path.join(out_dir, "power.npy"), power.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gathered_shard_summaries_equal_unsharded(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    scene, rx, tx, rxv, txv, f = _case()
    o, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B)
    pair, _ = tl.oracle_summaries(o, tr)
    total = np.load(tmp_path / "total.npy")
    for i, k in enumerate(("n_valid", "n_occluded", "hit_hash", "tau_bits")):
        assert np.array_equal(total[..., i], pair[k]), k
    np.testing.assert_allclose(np.load(tmp_path / "power.npy").reshape(pair["power_te"].shape),
                               pair["power_te"], rtol=1e-12)


def _list_from_oracle(o, tr, path_ids):
    """HrtPathRecord list (hrt_b200.PATH_DTYPE) of the valid slots of an oracle run
    restricted to `path_ids` (global path numbers of the columns)."""
    R, T = o.R, o.T
    st = tr["slot_state"]
    rr, tt, bb, pp = np.nonzero(st == 1)
    rec = np.zeros(rr.size, hrt.PATH_DTYPE)
    rec["rx"], rec["tx"], rec["bounce"], rec["path"] = rr, tt, bb, path_ids[pp]
    n = st.shape[-1]
    for k in ("a_te_re", "a_te_im", "a_tm_re", "a_tm_im", "tau", "freq_shift"):
        rec[k] = o.scat[k].reshape(R, T, B, n)[rr, tt, bb, pp]
    rec["direction_rx"] = o.scat["directions_rx"].reshape(R, T, B, n, 3)[rr, tt, bb, pp]
    return rec


def _list_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene, rx, tx, rxv, txv, f = _case()
    o, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B)
    mine = hrt.shard_paths(P, rank, world, BLOCK).astype(np.int64)
    sub = tl.abi.alloc_outputs(o.R, o.T, mine.size, B, 0)
    for k in ("tau", "a_te_re", "a_te_im", "a_tm_re", "a_tm_im", "freq_shift"):
        sub.scat[k][...] = o.scat[k][..., mine]
    sub.scat["directions_rx"][...] = o.scat["directions_rx"].reshape(o.R, o.T, B, P, 3)[:, :, :, mine].reshape(sub.scat["directions_rx"].shape)
    trs = {k: np.ascontiguousarray(v[..., mine]) for k, v in tr.items()}
    rec = _list_from_oracle(sub, trs, mine)
    cap = 2 * P * B
    buf = torch.zeros(cap * 48, dtype=torch.uint8)
    buf[: rec.size * 48] = torch.from_numpy(rec.view(np.uint8).copy())
    allrec, counts = hrt.gather_path_lists(buf, rec.size, cap)
    assert counts[rank] == rec.size
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), allrec.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_gathered_path_lists_equal_unsharded(tmp_path):
    """All-gather of per-rank compact path lists (hrt_b200.gather_path_lists, the
    call bench.py makes over NCCL) == the list of the unsharded run, as a set."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_list_worker, args=(WORLD, port, str(tmp_path)), nprocs=WORLD, join=True)
    scene, rx, tx, rxv, txv, f = _case()
    o, tr = tl.run_oracle(scene, rx, tx, rxv, txv, f, P, B)
    full = _list_from_oracle(o, tr, np.arange(P, dtype=np.int64))
    got = np.load(tmp_path / "gathered.npy").reshape(-1).view(hrt.PATH_DTYPE)
    assert got.size == full.size and full.size > 1000
    key = lambda q: np.lexsort((q["path"], q["bounce"], q["tx"], q["rx"]))
    a, b = full[key(full)], got[key(got)]
    assert a.tobytes() == b.tobytes()
