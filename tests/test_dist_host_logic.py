"""CPU: world_size-2 gloo run of the host-side sharding logic (ownership map, attribute gather)."""
import os
import socket
import subprocess
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_workers(backend, nproc, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(HERE, "dist_worker.py"), backend]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)


def test_sharding_host_logic_gloo_world2():
    r = run_workers("gloo", 2, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("host logic ok") == 2


def test_shardmap_and_splits_properties():
    from mbpls_b200.engine import ShardMap
    for sizes in ([100_000, 200_000, 300_000, 400_000], [20, 35, 60, 95, 140, 180, 220, 450], [5], [1, 1, 1]):
        for world in (1, 2, 3, 4, 8):
            seen = 0
            for rank in range(world):
                sh = ShardMap.build(sizes, rank, world)
                assert len(sh.block_off) == len(sizes) + 1 and sh.block_off[0] == 0
                assert sh.p_local == sh.hi - sh.lo == sum(c1 - c0 for c0, c1 in sh.local_ranges)
                seen += sh.p_local
            assert seen == sum(sizes)


def test_split_tables_cover_every_feature_inside_one_block():
    """Host logic shared by the sample-owning kernels and the one-pass workers: every split lies inside one block, the splits
    of a block tile it in order, and the table never exceeds one wave of workers."""
    from mbpls_b200.engine import make_splits
    for sizes, n, sms, per_sm, min_feats in (((100_000, 200_000, 300_000, 400_000), 1, 148, 1, 16),
                                             ((20_000, 35_000, 60_000, 95_000, 140_000, 180_000, 220_000, 450_000), 1, 148, 4, 16),
                                             ((1, 150, 2, 77, 1), 1, 148, 8, 16), ((5,), 1, 148, 8, 16),
                                             ((300, 50, 450), 400, 148, 4, 32), ((0, 7, 0, 9), 1, 148, 2, 16)):
        off = [0]
        for s in sizes:
            off.append(off[-1] + s)
        f0, f1, bso = make_splits(off, n, sms, ctas_per_sm=per_sm, min_feats=min_feats)
        assert len(bso) == len(sizes) + 1 and bso[0] == 0 and bso[-1] == len(f0)
        for b in range(len(sizes)):
            lo = off[b]
            for s in range(bso[b], bso[b + 1]):
                assert f0[s] == lo and f1[s] > f0[s] and f1[s] <= off[b + 1]
                lo = f1[s]
            assert lo == off[b + 1] or sizes[b] == 0
            assert (bso[b + 1] > bso[b]) == (sizes[b] > 0)
        row_chunks = max(1, -(-n // 512))
        assert len(f0) <= max(sms * per_sm // row_chunks, sum(1 for s in sizes if s > 0))


def test_one_pass_configuration_tables():
    """Pure host entry points of the C ABI: workers per CTA by feature length, NaN bit-matrix row length."""
    from mbpls_b200._cabi import call
    lds = (16, 1280, 1296, 2560, 2576, 5120, 5136, 10240, 10256, 20480, 20496)
    assert [call("mbpls_fused_workers_per_sm_pair", ld) for ld in lds] == [16, 16, 8, 8, 4, 4, 2, 2, 1, 1, 0]
    assert call("mbpls_fused_workers_per_sm_pair", 1000) == 0  # leading dimensions are multiples of 16
    assert [call("mbpls_nan_bitmask_ldw", n) for n in (1, 32, 33, 128, 129, 10_000)] == [4, 4, 4, 4, 8, 316]


def test_size_limits_are_checked_up_front():
    """K, q and B are capped at 64 by the single-CTA small-matrix kernels; the cap is reported by name before any
    data is uploaded, not as an opaque status after the NIPALS loop has run."""
    from mbpls_b200.mbpls import _check_limits, MAX_COMPONENTS
    assert _check_limits(3, 2, 4) == 3 and _check_limits(MAX_COMPONENTS, 64, 64) == 64
    for bad in ((65, 1, 1), (2, 65, 1), (2, 1, 65)):
        with pytest.raises(NotImplementedError, match="at most 64"):
            _check_limits(*bad)
    with pytest.raises(ValueError):
        _check_limits(0, 1, 1)
    with pytest.raises(ValueError):
        _check_limits("three", 1, 1)
