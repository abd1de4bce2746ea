"""bench.py's reference arm (the part of the bench contract that runs without a GPU): one JSON line
with the keys the driver reads, produced by the unmodified reference on the host cores."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, DABGPU_BENCH_REF_TFS="2")
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
    # the record describes what ran: 3 timed steps whose total fits inside the run, per-core figure, CPU model
    assert d["steps"] == 3 and d["warmup"] == 1
    assert d["steps"] * d["ms_per_step"] / 1e3 <= wall
    assert abs(d["run"]["timed_region_s"] - d["steps"] * d["ms_per_step"] / 1e3) < 1e-6
    assert d["cpu_baseline"]["per_core"] > 0 and d["cpu_baseline"]["cpu_model"]
    # both arms carry the same `config`
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.bench_config(1024, 128)


def test_reference_arm_cannot_map_the_product():
    """bench.py --impl reference sets DABGPU_FORBID_LOAD for itself and its workers: lib.load() refuses,
    while the synthetic transmitter (tables through the host-only libdabtables.so) keeps working."""
    code = ("import os; os.environ['DABGPU_FORBID_LOAD']='1'\n"
            "from dabtools_b200 import synth, lib\n"
            "synth.ModeITransmitter(synth.reference_ensemble()).generate(1, 1, seed=1)\n"
            "assert not any('libdabgpu' in l for l in open('/proc/self/maps'))\n"
            "try:\n    lib.load()\n    raise SystemExit(3)\nexcept lib.DabGpuError:\n    pass\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, env=env, timeout=120,
                       cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
