"""Worker for the multi-process tests: launched by torch.distributed.run (see test_gpu_multi.py /
test_dist_host_logic.py).  argv[1] = 'nccl' (one GPU per rank) or 'gloo' (CPU, host logic only)."""
import os
import sys
import warnings

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def gpu_checks(group, rank, world, dev):
    from helpers import assert_trips, compare, snapshot_model
    from mbpls_b200 import MBPLS
    from oracle import OracleMBPLS
    from oracle.cases import latent_blocks
    for nan_frac, n in ((0.0, 400), (0.08, 260), (0.0, 2300)):
        sizes = (300, 50, 450)  # uneven: shard boundaries cut through blocks, some ranks miss a block
        X, Y = latent_blocks(n, sizes, 3, 4, seed=5 + n, nan_frac=nan_frac)
        Xt, Yt = latent_blocks(13, sizes, 3, 4, seed=6, nan_frac=nan_frac)
        kw = dict(n_components=4, sparse_data=nan_frac > 0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = OracleMBPLS(**kw).fit([x.copy() for x in X], Y.copy())
            m = MBPLS(**kw).set_runtime(group=group, device=dev)
            m.fit([x.copy() for x in X], Y.copy())
        ref, ours = snapshot_model(o, Xt, Yt), snapshot_model(m, Xt, Yt)
        worst = compare(ours, ref, 1e-8, f"rank {rank} nan={nan_frac} n={n}")
        assert_trips(list(m.n_iter_), list(o.n_iter_), o.diff_trace_, 1e-14, f"rank {rank}")
        # pre-sharded input: every rank passes only its own column ranges
        from mbpls_b200.engine import ShardMap
        sh = ShardMap.build(sizes, rank, world)
        local = [X[b][:, c0:c1].copy() for b, (c0, c1) in enumerate(sh.local_ranges)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m2 = MBPLS(**kw).set_runtime(group=group, device=dev, global_sizes=list(sizes))
            m2.fit(local, Y.copy())
        # a rank that passed its own column ranges gets their per-feature attributes back ("local" gather: no all-gather of
        # p x K matrices, no redundant device->host copies); per-sample attributes are complete on every rank
        assert m2.beta_.shape[0] == sh.p_local and np.allclose(m2.beta_, m.beta_[sh.lo:sh.hi], rtol=1e-11, atol=1e-14)
        for b, (c0, c1) in enumerate(sh.local_ranges):
            assert m2.P_[b].shape == (c1 - c0, 4) and np.allclose(m2.P_[b], m.P_[b][c0:c1], rtol=1e-10, atol=1e-13)
            assert np.allclose(m2.x_scalers_[b].mean_, m.x_scalers_[b].mean_[c0:c1], rtol=1e-13, atol=1e-15)
        assert np.allclose(m2.Ts_, m.Ts_, rtol=1e-10, atol=1e-13)
        assert np.allclose(m2.predict([xt[:, c0:c1] for xt, (c0, c1) in zip(Xt, sh.local_ranges)]), m.predict(Xt), rtol=1e-9, atol=1e-12)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m3 = MBPLS(**kw).set_runtime(group=group, device=dev, global_sizes=list(sizes), gather="all")
            m3.fit(local, Y.copy())
        assert np.allclose(m3.beta_, m.beta_, rtol=1e-11, atol=1e-14), "pre-sharded fit with gather='all' differs"
        print(f"rank {rank}: world={world} nan={nan_frac} n={n} trips={m.n_iter_} worst={worst}", flush=True)
    # the other methods under the same feature sharding (KERNEL: the p > n branch; n >= p needs row sharding)
    # KERNEL n=1201 >= p=800 takes the row-sharded path (samples split over the ranks, p x p all-reduce)
    # UNIPALS n=1200 >= p=800 is row-sharded too: once in the reference's streaming form (per-component all-reduces of X'Y, X'ts,
    # Y'ts) and once through the Gram matrices (one all-reduce of X'X and X'Y)
    for method, n, route in (("SIMPLS", 400, None), ("UNIPALS", 400, None), ("UNIPALS", 1200, "stream"), ("UNIPALS", 1200, "gram"),
                             ("KERNEL", 400, None), ("KERNEL", 1201, None)):
        sizes = (300, 50, 450)
        X, Y = latent_blocks(n, sizes, 3, 4, seed=50 + n)
        Xt, Yt = latent_blocks(13, sizes, 3, 4, seed=6)
        kw = dict(n_components=4, method=method, full_svd=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            o = OracleMBPLS(**kw).fit([x.copy() for x in X], Y.copy())
            m = MBPLS(**kw).set_runtime(group=group, device=dev, unipals_route=route)
            m.fit([x.copy() for x in X], Y.copy())
        worst = compare(snapshot_model(m, Xt, Yt), snapshot_model(o, Xt, Yt), 1e-8, f"rank {rank} {method} n={n}")
        print(f"rank {rank}: world={world} {method} n={n} worst={worst}", flush=True)


def cv_checks(group, rank, world, dev):
    """Folds spread over the ranks: every rank returns the complete out-of-fold predictions of the single-GPU run."""
    from mbpls_b200 import MBPLS
    from mbpls_b200.model_selection import cross_val_predict
    from oracle.cases import latent_blocks
    X, Y = latent_blocks(42, (18, 11), 2, 3, seed=91)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for small in (True, False):
            one = cross_val_predict(MBPLS(n_components=3).set_runtime(device=dev, small_path=small), X, Y, cv=7, n_components_list=[1, 2, 3])
            many = cross_val_predict(MBPLS(n_components=3).set_runtime(device=dev, group=group, small_path=small), X, Y, cv=7,
                                     n_components_list=[1, 2, 3])
            for k in (1, 2, 3):
                assert many[k].shape == one[k].shape and np.allclose(many[k], one[k], rtol=1e-12, atol=1e-14), (small, k)
    print(f"rank {rank}: fold-parallel CV ok", flush=True)


def host_checks(group, rank, world):
    from mbpls_b200.engine import ShardMap
    from mbpls_b200.mbpls import MBPLS
    sizes = [7, 1, 12, 5]
    sh = ShardMap.build(sizes, rank, world)
    # every global feature is owned exactly once
    own = torch.zeros(sum(sizes), dtype=torch.int32)
    own[sh.lo:sh.hi] += 1
    dist.all_reduce(own, group=group)
    assert bool((own == 1).all())
    assert sh.block_off[-1] == sh.hi - sh.lo
    # gather of a K x p_local tensor reproduces the global matrix on every rank
    full = torch.arange(3 * sum(sizes), dtype=torch.float64).view(3, -1)
    m = MBPLS().set_runtime(group=group)
    got = m._gather_features(full[:, sh.lo:sh.hi].contiguous(), sh)
    assert np.array_equal(got, full.numpy())
    print(f"rank {rank}: host logic ok", flush=True)


def main():
    backend = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        gpu_checks(dist.group.WORLD, rank, world, dev)
        cv_checks(dist.group.WORLD, rank, world, dev)
    else:
        dist.init_process_group("gloo")
        host_checks(dist.group.WORLD, rank, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
