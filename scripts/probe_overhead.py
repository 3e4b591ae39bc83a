"""Where does the time of one bench step go outside hrt_run's own event window?
wall clock of ctx.run() vs its ms_total for host/device summaries x own/torch stream."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import hrt_b200 as hrt
import bench
rx, tx = bench.c4_positions(); zr, zt = np.zeros_like(rx), np.zeros_like(tx)
P = int(float(os.environ.get("RAYS", "1e8"))) // 4
ctx = hrt.Context(0); ctx.load_scene(bench.SCENE)
R, T, B = 64, 4, 5
st = torch.cuda.Stream()
pair_dev = torch.zeros(R * T * B * 6, dtype=torch.int64, device="cuda"); bounce_dev = torch.zeros(T * B * 4, dtype=torch.int64, device="cuda")
for name, kw in (("host summaries, ctx stream", {}),
                 ("dev summaries, ctx stream", dict(summary_dev_ptrs=(pair_dev.data_ptr(), bounce_dev.data_ptr()))),
                 ("dev summaries, torch stream", dict(summary_dev_ptrs=(pair_dev.data_ptr(), bounce_dev.data_ptr()), stream=st.cuda_stream)),
                 ("host summaries, torch stream", dict(stream=st.cuda_stream))):
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ctx.run(rx, tx, zr, zt, 3.5, P, B, summary=True, **kw)
        torch.cuda.synchronize()
        w = (time.perf_counter() - t0) * 1e3
        s = r["stats"]
        if rep:
            print(f"{name:32s} wall {w:8.2f} ms  ms_total {s['ms_total']:8.2f}  scatter {s['ms_scatter']:7.2f} bounce {s['ms_bounce']:6.2f} sort {s['ms_sort']:6.2f} other {s['ms_other']:6.2f} host_setup {s['host_ms_setup']:6.2f} host_total {s['host_ms_total']:7.2f}")
