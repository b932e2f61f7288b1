"""The bench contract, as far as it can be checked without a GPU: `bench.py --impl reference` (the CPU arm: the oracle port
timed on the host cores, the one place outside tests/ and smoke() where oracle/ may run) prints ONE contract-shaped JSON line,
and the last GPU line committed under profiles/ carries every key the contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches")


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    for k in BASE_KEYS:
        assert k in d, k
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["unit"] == "obs/s" and d["value"] > 0
    assert d["steps"] == 1 and d["warmup"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same metric as the GPU arm's line
    gpu = _json_lines(open(os.path.join(ROOT, "profiles", "bench_r2z.json")).read())[0]
    assert d["metric"] == gpu["metric"] and d["unit"] == gpu["unit"]


def test_committed_gpu_line_has_the_contract_keys():
    d = _json_lines(open(os.path.join(ROOT, "profiles", "bench_r2z.json")).read())[0]
    for k in BASE_KEYS + ("roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    # achieved = algorithmic bytes per launch / live kernel time (41 B per Bernoulli observation)
    n, ms = d["config"]["obs_per_gpu"], d["parts"]["ms_cavi"]
    assert abs(r["achieved"] - 41 * n / (ms * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # traffic (ncu dram bytes per launch) is not above the algorithmic bytes: no wasted re-reads
    assert r["traffic"] <= 41 * n * 1.02
    cb = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in cb, k
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] >= 2 * d["steps"]
    # the parts add up to the step
    p = d["parts"]
    assert abs(p["ms_cavi"] + p["ms_gibbs"] + p["ms_allreduce"] - d["ms_per_step"]) < 0.02 * d["ms_per_step"]


def test_reference_arm_under_torchrun_prints_from_rank_0_only():
    """N > 1: the driver launches the reference arm like the GPU arm; rank 0 alone runs and prints, the others exit 0."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
