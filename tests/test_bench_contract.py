"""CPU-only: the reference arm of bench.py (`--impl reference`: the oracle's restatement of the reference's CPU path on
the host cores) prints ONE JSON line with the keys of the bench contract, reports the steps it actually ran, a
`config` identical to our arm's for the same workload, and figures that are consistent with each other."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra], capture_output=True,
                       text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_matmul_line():
    # launched like a torchrun worker: OMP_NUM_THREADS=1 in the environment must not throttle the CPU arm
    d = _run("--steps", "2", "--warmup", "0", env={"OMP_NUM_THREADS": "1"})
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "matmul_gflops" and d["unit"] == "GFLOP/s"
    assert d["steps"] == 2 and d["warmup"] == 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value and ms_per_step describe the same timed steps: rows x 2 N^2 flop per step
    sample = d["cpu_baseline"]["sample"]
    rows = 4096 if sample.startswith("the whole product") else int(sample.split()[1])
    assert abs(2.0 * rows * 4096 * 4096 / (d["ms_per_step"] * 1e-3) / 1e9 - d["value"]) / d["value"] < 1e-6
    # the same config dict as our arm prints for this workload (the driver compares them byte for byte)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.matmul_config(1)


def test_reference_arm_dense_line_is_rank0_only():
    d = _run("--workload", "dense", "--steps", "1", "--warmup", "0")
    assert d["metric"] == "dense_train_samples_per_s" and d["unit"] == "samples/s" and d["steps"] == 1
    assert abs(1024 / (d["ms_per_step"] * 1e-3) - d["value"]) / d["value"] < 1e-6
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.dense_config(1)
    # any other rank exits 0 without work and without output
    e = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=e, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
