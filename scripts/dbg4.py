import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from tests.util import clustered
from cuvs_b200.neighbors import ivf_pq as m
ds, centers = clustered(40000, 128, 15, n_centers=64)
qs, _ = clustered(512, 128, 16, centers=centers)
index = m.build(m.IndexParams(n_lists=64, pq_dim=64, kmeans_n_iters=10), torch.from_numpy(ds).cuda())
os.environ["CUVS_B200_PQ_PATH"] = "tc"
os.environ["CUVS_B200_PQ_NO_BOUND"] = "1"
def run(g):
    os.environ["CUVS_B200_PQ_GROUP"] = g
    os.environ["CUVS_B200_PQ_DUMP"] = f"/tmp/dump_{g}.bin"
    m.search(m.SearchParams(n_probes=8, lut_dtype=np.float16), index, torch.from_numpy(qs).cuda(), 10)
    torch.cuda.synchronize()
    raw = open(f"/tmp/dump_{g}.bin", "rb").read()
    nq, npb, kcw, npairs = np.frombuffer(raw[:32], np.int64)
    o = 32
    slot = np.frombuffer(raw[o:o + nq * npb * 4], np.uint32).reshape(nq, npb); o += nq * npb * 4
    probes = np.frombuffer(raw[o:o + nq * npb * 4], np.uint32).reshape(nq, npb); o += nq * npb * 4
    cs = np.frombuffer(raw[o:o + npairs * kcw * 4], np.float32).reshape(npairs, kcw); o += npairs * kcw * 4
    cp = np.frombuffer(raw[o:o + npairs * kcw * 4], np.uint32).reshape(npairs, kcw)
    out = {}
    for q in range(nq):
        for p in range(npb):
            s = slot[q, p]
            if s != 0xffffffff:
                out[(q, int(probes[q, p]))] = sorted((float(a), int(b)) for a, b in zip(cs[s], cp[s]) if b != 0xffffffff)
    return out
a, b = run("32"), run("64")
sizes = index.list_sizes.cpu().numpy()
offs = np.concatenate([[0], np.cumsum((sizes.astype(np.int64) + 127) // 128 * 128)]).astype(np.int64)
bad = 0; kinds = {"missing_rows": 0, "score_diff": 0}; tiles = {}
for key in a:
    ca, cb = a[key], b.get(key, [])
    pa, pb_ = {p: s for s, p in ca}, {p: s for s, p in cb}
    if set(pa) != set(pb_) or any(abs(pa[p] - pb_[p]) > 1e-4 * max(1, abs(pa[p])) for p in pa if p in pb_):
        bad += 1
        for p in pa:
            if p not in pb_:
                kinds["missing_rows"] += 1
                t = (p - offs[key[1]]) // 128
                tiles[int(t)] = tiles.get(int(t), 0) + 1
            elif abs(pa[p] - pb_[p]) > 1e-4 * max(1, abs(pa[p])):
                kinds["score_diff"] += 1
        if bad <= 4:
            print("pair", key, "list rows", sizes[key[1]], "\n  g32:", ca[:6], "\n  g64:", cb[:6])
print("pairs", len(a), "differing", bad, kinds, "missing by tile index within list:", dict(sorted(tiles.items())))

# ---- which rows do the wrong scores belong to?  recompute s = |y|^2/2 - r.y on the host for the same tile slot of neighbouring tiles
pqc = index.pq_centers.cpu().numpy()            # [pq_dim, 2, 256]
crot = index.centers_rot.cpu().numpy()
rot = index.rotation_matrix.cpu().numpy()
def bf16(x):
    u = x.astype(np.float32).view(np.uint32)
    u = (u + 0x7fff + ((u >> 16) & 1)) & 0xffff0000
    return u.view(np.float32)
codes = {}
def decoded(l):
    if l not in codes:
        c = index.list_data(l).cpu().numpy()      # [n, 64]
        y = np.stack([bf16(pqc[j, t, c[:, j]]) for j in range(64) for t in range(2)], 1)   # [n, 128]
        codes[l] = y
    return codes[l]
shown = 0; explained = {}
for key in a:
    pa, pb_ = {p: s for s, p in a[key]}, {p: s for s, p in b.get(key, [])}
    wrong = [p for p in pb_ if (p not in pa and pb_[p] < max(pa.values())) or (p in pa and abs(pa[p] - pb_[p]) > 1e-3 * max(1, abs(pa[p])))]
    if not wrong:
        continue
    q, l = key
    r = bf16(qs[q] @ rot.T - crot[l])
    y = decoded(l)
    n = len(y)
    for p in wrong:
        i = int(p - offs[l])
        got = pb_[p]
        dts = [dt for dt in range(-4, 5) if 0 <= i + 128 * dt < n]
        P = {dt: np.array([-(r[16 * k:16 * k + 16] * y[i + 128 * dt, 16 * k:16 * k + 16]).sum() for k in range(8)]) for dt in dts}
        E = {dt: 0.5 * (y[i + 128 * dt] ** 2).sum() for dt in dts}
        true = E[0] + P[0].sum()
        hit = None
        for dA in dts:
            for dB in dts:
                for c in range(0, 9):
                    for dE in (dA, dB):
                        v = P[dA][:c].sum() + P[dB][c:].sum() + E[dE]
                        if abs(v - got) < 2e-3 * max(1, abs(got)):
                            hit = (dA, dB, c, "extA" if dE == dA else "extB"); break
                    if hit: break
                if hit: break
            if hit: break
        explained[hit] = explained.get(hit, 0) + 1
        if shown < 10:
            print("pair", key, "row", i, "tile", i // 128, "slot", i % 128, "warp", (i % 128) // 16, "reported", round(got, 3), "true", round(float(true), 3), "explained by (first k-steps from dt, rest from dt, cut, ext)", hit)
            shown += 1
print("explanations:", explained)
