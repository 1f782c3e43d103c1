"""CPU: world_size-2 gloo run of the host-side sharding logic (ownership map, attribute gather)."""
import os
import socket
import subprocess
import sys

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
