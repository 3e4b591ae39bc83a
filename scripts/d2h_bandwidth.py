import torch, time, numpy as np, threading
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.empty(n, dtype=torch.uint8)              # pageable
hp.fill_(1)
hq = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, h in (("pageable", hp), ("pinned", hq)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); h.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"D2H {name}: {n/dt/1e9:.1f} GB/s")
# host memcpy pinned -> pageable, 1..8 threads
src = hq.numpy(); dst = hp.numpy()
for k in (1, 2, 4, 8):
    def work(i):
        a = n * i // k; b = n * (i + 1) // k
        dst[a:b] = src[a:b]
    for _ in range(2):
        ts = [threading.Thread(target=work, args=(i,)) for i in range(k)]
        t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
    print(f"host memcpy {k} threads: {n/dt/1e9:.1f} GB/s")
# fresh (untouched) pageable destination
for k in (1, 8):
    fresh = np.empty(n, np.uint8)
    def work(i):
        a = n * i // k; b = n * (i + 1) // k
        fresh[a:b] = src[a:b]
    ts = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    t0 = time.perf_counter(); [t.start() for t in ts]; [t.join() for t in ts]; dt = time.perf_counter() - t0
    print(f"host memcpy into FRESH pages, {k} threads: {n/dt/1e9:.1f} GB/s")
fresh = torch.empty(n, dtype=torch.uint8)
torch.cuda.synchronize(); t0 = time.perf_counter(); fresh.copy_(d); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H into FRESH pageable: {n/dt/1e9:.1f} GB/s")
