"""CPU: bench.py's reference arm (the oracle port timed on the host cores) prints ONE JSON line with the keys the driver
reads, on a tiny bounded sample; and the canonical byte formula is the one of SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--n", "300",
           "--cpu-p", "400", "--components", "3"]
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must still use the host's cores
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    sys.path.insert(0, ROOT)
    from oracle import refshim
    # the unmodified reference when its tree is mounted (this container), the numpy port on the GPU box
    assert d["cpu_baseline"]["kind"] == ("reference" if refshim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= min(2, os.cpu_count()) and d["cpu_baseline"]["value"] == d["value"]
    assert len(d["linear_in_p"]) == 2 and all(v["gbs"] > 0 for v in d["linear_in_p"].values())
    assert d["nan_10pct"]["value"] > 0 and "nan_frac=0.1" in d["nan_10pct"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["trips_per_component"] == [2, 2, 2]  # PLS1: two trips per component, like the reference
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_canonical_byte_formula():
    sys.path.insert(0, ROOT)
    import bench
    # 16 n p (1 + K + sum of trips): standardise 1R+1W, per trip 2 reads, per component 1R+1W
    assert bench.fit_bytes(10_000, 1_000_000, 20, [2] * 20) == 16.0 * 10_000 * 1_000_000 * (1 + 20 + 40) == 9.76e12
