"""BVH builder laboratory (CPU, numpy): how many ray/box and ray/triangle tests
does a closest-hit shadow query need with trees from different builders?

Used to choose the GPU builder.  Counts follow the same
convention as the instrumented kernels (HRT_FLAG_COUNT): two box tests per
visited inner node, one triangle test per triangle of every visited leaf,
front-to-back order with culling against the running closest hit.

usage: bvh_lab.py [canyon|tiled:NX] [n_hits]
"""
import sys, os, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
from hrt_b200.scenes import read_hrt, tiled_canyon, c5_positions  # noqa: E402


def load(which):
    base = os.path.join(ROOT, "scenes", "simple_street_canyon_with_cars.hrt")
    if which == "canyon":
        meshes = read_hrt(base)
        rx = np.array([[-60 + 8 * j, y, 1.5] for j in range(16) for y in (-3, -1, 1, 3)], np.float64)
        tx = np.array([[-45 + 30 * i, 0, 10] for i in range(4)], np.float64)
    else:
        nx = int(which.split(":")[1])
        meshes, pitch = tiled_canyon(base, nx, nx)
        rx, tx = c5_positions(pitch, nx, nx)
        rx = rx.astype(np.float64); tx = tx.astype(np.float64)
    tris = np.concatenate([m["vs"][m["tris"]] for m in meshes]).astype(np.float64)
    return tris, rx, tx


def mt_all(o, d, tris):
    """closest hit of one ray against all triangles (float64) -> (t, idx)"""
    a = tris[:, 0]; ab = tris[:, 1] - a; ac = tris[:, 2] - a
    pv = np.cross(d, ac); det = (ab * pv).sum(1)
    ok = np.abs(det) > 1e-12
    det = np.where(ok, det, 1.0)
    s = o - a; u = (s * pv).sum(1) / det
    q = np.cross(s, ab); v = (q * d).sum(1) / det
    t = (ac * q).sum(1) / det
    ok &= (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-7)
    if not ok.any():
        return None
    t = np.where(ok, t, np.inf); i = int(np.argmin(t))
    return t[i], i


def fib_dirs(n, P):
    k = np.arange(n) * (P // n) + 0.5
    phi = np.arccos(1 - 2 * k / P); th = np.pi * (1 + 5 ** 0.5) * k
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)


def shadow_rays(tris, rx, tx, n_hits, max_rx=64, bounces=3):
    """(origins, directions) of shadow rays as compute_paths casts them"""
    nrm = np.cross(tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    O, D = [], []
    dirs = fib_dirs(n_hits, 1000003)
    rsel = rx if len(rx) <= max_rx else rx[np.linspace(0, len(rx) - 1, max_rx).astype(int)]
    for i, d in enumerate(dirs):
        o = tx[i % len(tx)].copy(); d = d.copy()
        for b in range(bounces):
            h = mt_all(o, d, tris)
            if h is None:
                break
            t, k = h
            o = o + d * t; d = d - 2 * (d @ nrm[k]) * nrm[k]; o = o + d * 1e-4
            for r in rsel:
                sd = r - o; sd /= np.linalg.norm(sd)
                O.append(o.copy()); D.append(sd)
    return np.array(O), np.array(D)


# ------------------------------------------------------------------ builders
# a tree is (left, right, lo, hi, leaf_first, leaf_count, order): arrays over nodes;
# leaves have left == -1 and cover order[leaf_first : leaf_first+leaf_count]

class Tree:
    def __init__(self):
        self.left = []; self.right = []; self.lo = []; self.hi = []; self.first = []; self.count = []
        self.order = []

    def add(self, lo, hi, left=-1, right=-1, first=0, count=0):
        self.left.append(left); self.right.append(right); self.lo.append(lo); self.hi.append(hi)
        self.first.append(first); self.count.append(count)
        return len(self.left) - 1


def area(lo, hi):
    e = np.maximum(hi - lo, 0)
    return 2 * (e[..., 0] * e[..., 1] + e[..., 1] * e[..., 2] + e[..., 2] * e[..., 0])


def build_sah(tlo, thi, leaf_max=2, bins=0):
    """top-down SAH, full sweep over the three axes (bins=0) -- the quality yardstick"""
    T = Tree(); cen = 0.5 * (tlo + thi)

    def rec(ids):
        lo = tlo[ids].min(0); hi = thi[ids].max(0)
        if len(ids) <= leaf_max:
            f = len(T.order); T.order.extend(ids.tolist())
            return T.add(lo, hi, first=f, count=len(ids))
        best = (np.inf, None, None)
        for ax in range(3):
            o = ids[np.argsort(cen[ids, ax], kind="stable")]
            l_lo = np.minimum.accumulate(tlo[o], 0); l_hi = np.maximum.accumulate(thi[o], 0)
            r_lo = np.minimum.accumulate(tlo[o][::-1], 0)[::-1]; r_hi = np.maximum.accumulate(thi[o][::-1], 0)[::-1]
            n = len(o); k = np.arange(1, n)
            cost = area(l_lo[:-1], l_hi[:-1]) * k + area(r_lo[1:], r_hi[1:]) * (n - k)
            j = int(np.argmin(cost))
            if cost[j] < best[0]:
                best = (cost[j], o, j + 1)
        _, o, j = best
        me = T.add(lo, hi)
        l = rec(o[:j]); r = rec(o[j:])
        T.left[me] = l; T.right[me] = r
        return me

    sys.setrecursionlimit(100000)
    rec(np.arange(len(tlo)))
    return T


def build_binned(tlo, thi, leaf_max=2, nbins=16, axes="all", bounds="centroid"):
    """top-down binned SAH as the GPU builder does it: nbins bins per axis over the
    node's centroid bounds (or its box), best plane over the tested axes; nodes
    whose centroids do not separate are halved by position"""
    T = Tree(); cen = 0.5 * (tlo + thi)

    def rec(ids):
        lo = tlo[ids].min(0); hi = thi[ids].max(0)
        if len(ids) <= leaf_max:
            f = len(T.order); T.order.extend(ids.tolist())
            return T.add(lo, hi, first=f, count=len(ids))
        if bounds == "centroid":
            blo = cen[ids].min(0); bhi = cen[ids].max(0)
        else:
            blo, bhi = lo, hi
        ext = bhi - blo
        ax_list = range(3) if axes == "all" else [int(np.argmax(ext))]
        best = (np.inf, None)
        for ax in ax_list:
            if ext[ax] <= 0:
                continue
            b = np.minimum(((cen[ids, ax] - blo[ax]) / ext[ax] * nbins).astype(int), nbins - 1)
            b = np.maximum(b, 0)
            cnt = np.bincount(b, minlength=nbins)
            blo_b = np.full((nbins, 3), np.inf); bhi_b = np.full((nbins, 3), -np.inf)
            np.minimum.at(blo_b, b, tlo[ids]); np.maximum.at(bhi_b, b, thi[ids])
            l_lo = np.minimum.accumulate(blo_b, 0); l_hi = np.maximum.accumulate(bhi_b, 0)
            r_lo = np.minimum.accumulate(blo_b[::-1], 0)[::-1]; r_hi = np.maximum.accumulate(bhi_b[::-1], 0)[::-1]
            cl = np.cumsum(cnt)
            for k in range(1, nbins):
                nl = cl[k - 1]; nr = len(ids) - nl
                if nl == 0 or nr == 0:
                    continue
                c = area(l_lo[k - 1], l_hi[k - 1]) * nl + area(r_lo[k], r_hi[k]) * nr
                if c < best[0]:
                    best = (c, b < k)
        if best[1] is None:
            m = np.zeros(len(ids), bool); m[:len(ids) // 2] = True
        else:
            m = best[1]
        me = T.add(lo, hi)
        l = rec(ids[m]); r = rec(ids[~m])
        T.left[me] = l; T.right[me] = r
        return me

    sys.setrecursionlimit(100000)
    rec(np.arange(len(tlo)))
    return T


def expand_bits(v, nbits):
    out = np.zeros_like(v, dtype=np.uint64)
    for b in range(nbits):
        out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out


def morton_order(tlo, thi, cube=False, nbits=10):
    cen = 0.5 * (tlo + thi); slo = tlo.min(0); ext = thi.max(0) - slo
    if cube:
        ext = np.full(3, ext.max())
    q = np.clip(((cen - slo) / np.maximum(ext, 1e-30) * (1 << nbits)).astype(np.int64), 0, (1 << nbits) - 1).astype(np.uint64)
    code = (expand_bits(q[:, 0], nbits) << np.uint64(2)) | (expand_bits(q[:, 1], nbits) << np.uint64(1)) | expand_bits(q[:, 2], nbits)
    key = (code << np.uint64(32 - 0)) if nbits <= 10 else code
    order = np.lexsort((np.arange(len(code)), code))
    return order, code[order]


def build_lbvh(tlo, thi, leaf_max=2, cube=False, nbits=10):
    """Karras-style radix tree over Morton codes (ties broken by index), subtrees of <= leaf_max cut into leaves"""
    order, code = morton_order(tlo, thi, cube, nbits)
    n = len(order)
    key = [(int(code[i]) << 32) | int(order[i]) for i in range(n)]
    T = Tree(); T.order = order.tolist()

    def rec(a, b):  # inclusive range
        ids = order[a:b + 1]
        lo = tlo[ids].min(0); hi = thi[ids].max(0)
        if b - a + 1 <= leaf_max:
            return T.add(lo, hi, first=a, count=b - a + 1)
        x = key[a] ^ key[b]
        hb = x.bit_length() - 1
        # split: last index whose bit hb equals that of key[a]
        lo_i, hi_i = a, b
        while lo_i < hi_i:   # largest m with bit equal to key[a]'s
            m = (lo_i + hi_i + 1) // 2
            if ((key[m] ^ key[a]) >> hb) == 0:
                lo_i = m
            else:
                hi_i = m - 1
        me = T.add(lo, hi)
        l = rec(a, lo_i); r = rec(lo_i + 1, b)
        T.left[me] = l; T.right[me] = r
        return me

    sys.setrecursionlimit(100000)
    rec(0, n - 1)
    return T


def build_ploc(tlo, thi, radius=16, leaf_max=2, cube=True, nbits=21, collapse=True):
    """PLOC (Meister & Bittner 2018): Morton-ordered clusters, each round every
    cluster picks the neighbour within +-radius that minimises the merged area;
    mutual pairs merge.  Then subtrees of <= leaf_max triangles become leaves."""
    order, _ = morton_order(tlo, thi, cube, nbits)
    n = len(order)
    # node arrays (2n-1): leaves first
    lo = np.zeros((2 * n - 1, 3)); hi = np.zeros((2 * n - 1, 3))
    left = -np.ones(2 * n - 1, np.int64); right = -np.ones(2 * n - 1, np.int64); cnt = np.ones(2 * n - 1, np.int64)
    lo[:n] = tlo[order]; hi[:n] = thi[order]
    C = np.arange(n); nxt = n
    rounds = 0
    while len(C) > 1:
        m = len(C); rounds += 1
        best = np.full(m, np.inf); nn = np.full(m, -1)
        for off in range(1, radius + 1):
            if off >= m:
                break
            a = area(np.minimum(lo[C[:-off]], lo[C[off:]]), np.maximum(hi[C[:-off]], hi[C[off:]]))
            # candidate for i (neighbour i+off) and for i+off (neighbour i)
            idx = np.arange(m - off)
            better = a < best[idx]
            best[idx] = np.where(better, a, best[idx]); nn[idx] = np.where(better, idx + off, nn[idx])
            better = a < best[idx + off]          # strict: ties keep the earlier (smaller-index) neighbour
            best[idx + off] = np.where(better, a, best[idx + off]); nn[idx + off] = np.where(better, idx, nn[idx + off])
        i = np.arange(m)
        mutual = (nn[nn] == i)
        lead = mutual & (i < nn)
        k = int(lead.sum())
        new_ids = nxt + np.arange(k)
        li = C[i[lead]]; ri = C[nn[lead]]
        left[new_ids] = li; right[new_ids] = ri
        lo[new_ids] = np.minimum(lo[li], lo[ri]); hi[new_ids] = np.maximum(hi[li], hi[ri]); cnt[new_ids] = cnt[li] + cnt[ri]
        nxt += k
        keep = ~(mutual & (i > nn))
        Cn = C.copy(); Cn[lead] = new_ids
        C = Cn[keep]
    root = C[0]
    # emit with leaf collapse, DFS order for leaf slots
    T = Tree()

    def tri_list(v, out):
        st = [v]
        while st:
            x = st.pop()
            if left[x] < 0:
                out.append(int(order[x]))
            else:
                st.append(right[x]); st.append(left[x])

    def rec(v):
        if cnt[v] <= leaf_max:
            f = len(T.order); tri_list(v, T.order)
            return T.add(lo[v], hi[v], first=f, count=int(cnt[v]))
        me = T.add(lo[v], hi[v])
        l = rec(left[v]); r = rec(right[v])
        T.left[me] = l; T.right[me] = r
        return me

    sys.setrecursionlimit(1000000)
    rec(root)
    T.rounds = rounds
    return T


def sah_cost(T):
    lo = np.array(T.lo); hi = np.array(T.hi); a = area(lo, hi); root = a[0]
    inner = np.array(T.left) >= 0
    return (a[inner].sum() * 1.0 + (a[~inner] * np.array(T.count)[~inner]).sum() * 1.0) / root


# ------------------------------------------------------------------ traversal count

def count_tests(T, tris, O, D):
    left = np.array(T.left); right = np.array(T.right); lo = np.array(T.lo); hi = np.array(T.hi)
    first = np.array(T.first); count = np.array(T.count); order = np.array(T.order)
    a = tris[:, 0]; ab = tris[:, 1] - a; ac = tris[:, 2] - a
    nb = nt = 0; depth_max = 0
    for o, d in zip(O, D):
        inv = 1.0 / np.where(np.abs(d) < 1e-20, 1e-20, d)
        tmax = np.inf

        def slab(k):
            t0 = (lo[k] - o) * inv; t1 = (hi[k] - o) * inv
            tn = max(np.minimum(t0, t1).max(), 0.0); tf = min(np.maximum(t0, t1).min(), tmax)
            return tn <= tf, tn

        if left[0] < 0:
            stack = []; cur = 0
        stack = []; cur = 0
        while True:
            while left[cur] >= 0:
                l, r = left[cur], right[cur]
                hl, tl = slab(l); hr, tr = slab(r); nb += 2
                if hl and hr:
                    if tl <= tr:
                        stack.append((r, tr)); cur = l
                    else:
                        stack.append((l, tl)); cur = r
                    depth_max = max(depth_max, len(stack))
                elif hl:
                    cur = l
                elif hr:
                    cur = r
                else:
                    cur = -1; break
            if cur >= 0:
                for s in range(first[cur], first[cur] + count[cur]):
                    k = order[s]; nt += 1
                    pv = np.cross(d, ac[k]); det = ab[k] @ pv
                    if abs(det) < 1e-12:
                        continue
                    sv = o - a[k]; u = (sv @ pv) / det
                    if u < 0 or u > 1:
                        continue
                    q = np.cross(sv, ab[k]); v = (d @ q) / det
                    if v < 0 or u + v > 1:
                        continue
                    t = (ac[k] @ q) / det
                    if t > 1e-7 and t < tmax:
                        tmax = t
            cur = -1
            while stack:
                c, tn = stack.pop()
                if tn <= tmax:
                    cur = c; break
            if cur < 0:
                break
    return nb / len(O), nt / len(O), depth_max


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "canyon"
    n_hits = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    tris, rx, tx = load(which)
    print(which, "triangles", len(tris), "rx", len(rx), "tx", len(tx))
    t0 = time.time()
    O, D = shadow_rays(tris, rx, tx, n_hits, max_rx=16 if len(tris) > 5000 else 64)
    print("shadow rays", len(O), "in %.1fs" % (time.time() - t0))
    tlo = tris.min(1); thi = tris.max(1)
    builders = [("lbvh10 (current)", lambda: build_lbvh(tlo, thi, 2, False, 10)),
                ("lbvh cube 21b", lambda: build_lbvh(tlo, thi, 2, True, 21)),
                ("ploc r=8", lambda: build_ploc(tlo, thi, 8)),
                ("ploc r=16", lambda: build_ploc(tlo, thi, 16)),
                ("ploc r=32", lambda: build_ploc(tlo, thi, 32))]
    builders += [("binned16 all cen", lambda: build_binned(tlo, thi, 2, 16, "all", "centroid")),
                 ("binned16 all aabb", lambda: build_binned(tlo, thi, 2, 16, "all", "aabb")),
                 ("binned16 longest cen", lambda: build_binned(tlo, thi, 2, 16, "longest", "centroid")),
                 ("binned8 all cen", lambda: build_binned(tlo, thi, 2, 8, "all", "centroid")),
                 ("binned32 all cen", lambda: build_binned(tlo, thi, 2, 32, "all", "centroid")),
                 ("binned16 all cen leaf1", lambda: build_binned(tlo, thi, 1, 16, "all", "centroid")),
                 ("binned16 all cen leaf4", lambda: build_binned(tlo, thi, 4, 16, "all", "centroid"))]
    if len(tris) <= 20000:
        builders.append(("sah sweep", lambda: build_sah(tlo, thi, 2)))
    sel = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    for name, fn in builders:
        if sel and not any(s in name for s in sel):
            continue
        t0 = time.time(); T = fn(); tb = time.time() - t0
        nb, nt, dm = count_tests(T, tris, O, D)
        print(f"{name:18s} nodes {len(T.left):8d} sah {sah_cost(T):8.2f} box/q {nb:7.2f} tri/q {nt:6.2f} stack {dm:3d} build {tb:.1f}s"
              + (f" rounds {T.rounds}" if hasattr(T, 'rounds') else ""))
