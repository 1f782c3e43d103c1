"""Ad-hoc GPU debugging: step-by-step comparison of the CUDA path with numpy (not a test)."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from mbpls_b200 import MBPLS, engine as E, _cabi
from mbpls_b200.engine import call, ptr, stream_ptr
from oracle import OracleMBPLS, OracleScaler
from oracle.cases import latent_blocks
from helpers import rel_err, col_err, snapshot_model

warnings.simplefilter("ignore")
dev = torch.device("cuda:0")
np.set_printoptions(precision=6, linewidth=200)

def step_check(n, sizes, q):
    print(f"=== step check n={n} sizes={sizes} q={q}")
    X, Y = latent_blocks(n, sizes, q, 3, seed=1)
    Xs = [OracleScaler().fit_transform(x) for x in X]
    Ys = OracleScaler().fit_transform(Y)
    shard = E.ShardMap.build(sizes)
    Xt = E.ingest_blocks(X, n, shard, dev)
    print("ingest err", rel_err(Xt[:, :n].cpu().numpy(), np.hstack(X).T))
    st = E.standardize_fit(Xt, n)
    Xc = np.hstack(Xs)
    print("standardize err", rel_err(Xt[:, :n].cpu().numpy(), Xc.T))
    Yt = E.alloc_feature_major(q, n, dev); E.ingest_feature_major(Y, n, 0, q, Yt, dev); E.standardize_fit(Yt, n)
    print("Y std err", rel_err(Yt[:, :n].cpu().numpy(), Ys.T))
    p, ld = Xt.shape; B = len(sizes)
    boff = E._i32(shard.block_off, dev)
    u = Yt[0].clone(); uu = torch.zeros(8, dtype=torch.float64, device=dev)
    uu[0] = float((Ys[:, 0] ** 2).sum())
    w = torch.zeros(p, dtype=torch.float64, device=dev)
    nparts = call("mbpls_xtu_num_ctas", p)
    norm_part = torch.zeros(nparts * B, dtype=torch.float64, device=dev)
    call("mbpls_nipals_xtu_f64", ptr(Xt), ld, n, p, ptr(u), ptr(uu), ptr(boff), B, ptr(w), ptr(norm_part), 0, None, stream_ptr(dev))
    torch.cuda.synchronize()
    w_ref = Xc.T @ Ys[:, 0] / (Ys[:, 0] @ Ys[:, 0])
    print("xtu err", rel_err(w.cpu().numpy(), w_ref), "feats/cta", call("mbpls_xtu_feats_per_cta", p), "nparts", nparts)
    nb = norm_part.view(nparts, B).sum(0).cpu().numpy()
    bo = shard.block_off
    print("norm parts", nb, "ref", [float((w_ref[bo[b]:bo[b+1]] ** 2).sum()) for b in range(B)])
    f0, f1, bso = E.make_splits(shard.block_off, n, E.sm_count(dev))
    print("splits", len(f0), f0[:5], f1[:5], bso)
    ns = len(f0)
    Tnum = torch.zeros((ns, ld), dtype=torch.float64, device=dev)
    sf0, sf1, sb = E._i32(f0, dev), E._i32(f1, dev), E._i32(bso, dev)
    call("mbpls_nipals_xw_f64", ptr(Xt), ld, n, ptr(w), ptr(sf0), ptr(sf1), ns, ptr(Tnum), None, ld, 0, None, stream_ptr(dev))
    red = torch.zeros(B * ld + B, dtype=torch.float64, device=dev)
    call("mbpls_nipals_reduce_partials_f64", ptr(Tnum), None, ld, n, B, ptr(sb), ptr(norm_part), nparts, ptr(red), 0, None, stream_ptr(dev))
    torch.cuda.synchronize()
    for b in range(B):
        t_ref = Xc[:, bo[b]:bo[b+1]] @ w_ref[bo[b]:bo[b+1]]
        print(f"  xw block {b} err", rel_err(red[b*ld:b*ld+n].cpu().numpy(), t_ref))
    print("  red norms", red[B*ld:].cpu().numpy())

def fit_check(n, sizes, q, K=3, **kw):
    print(f"=== fit check n={n} sizes={sizes} q={q} {kw}")
    X, Y = latent_blocks(n, sizes, q, K, seed=2, nan_frac=0.1 if kw.get("sparse_data") else 0.0)
    Xte, Yte = latent_blocks(7, sizes, q, K, seed=3, nan_frac=0.1 if kw.get("sparse_data") else 0.0)
    o = OracleMBPLS(n_components=K, **kw).fit([x.copy() for x in X], Y.copy())
    m = MBPLS(n_components=K, **kw).fit([x.copy() for x in X], Y.copy())
    print("trips ours", m.n_iter_, "oracle", o.n_iter_)
    a, b = snapshot_model(m, Xte, Yte), snapshot_model(o, Xte, Yte)
    for k in sorted(b):
        if k not in a: print("  missing", k); continue
        r, v = b[k], np.asarray(a[k])
        if v.shape != r.shape: print(f"  {k}: shape {v.shape} vs {r.shape}"); continue
        if r.dtype.kind in "iub": print(f"  {k}: int equal={np.array_equal(v, r)}"); continue
        e = col_err(v, r) if (r.ndim == 2 and r.shape[1] > 0) else rel_err(v, r)
        flag = "" if e < 1e-8 else "   <<<<<<"
        print(f"  {k:28s} err {e:.3e}{flag}")

step_check(200, (300, 500), 2)
step_check(2500, (260, 400), 2)
fit_check(200, (300, 500), 2)
fit_check(200, (300, 500), 1)
fit_check(64, (40, 70, 25), 3, sparse_data=True)
