"""The committed bench lines under profiles/ must be internally consistent: anybody recomputing a figure from the other
fields of the same line (as a reviewer does) has to land on the number that is printed.  CPU-only."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02_n*_cfg*.json")))


def _load(path):
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_recomputes(path):
    d = _load(path)
    c = d["config"]
    M, K, N, nnz, n = c["M"], c["K"], c["N"], c["nnz"], d["n_gpus"]
    assert d["verified"] is True and d["higher_is_better"] is True and d["dtype"] == "f32" and d["data"] == "synthetic"
    # metric: 2*M*nnz + M*N per call; `config` describes the WHOLE job (weak scaling: N and nnz grow with the GPU count)
    per_call = 2.0 * M * nnz + float(M) * N
    jobs = 1
    if d["scaling"] == "weak":
        assert N == 4096 * n
    assert d["value"] == pytest.approx(jobs * per_call / (d["ms_per_step"] * 1e-3) / 1e9, rel=2e-3)
    r = d["roofline"]
    assert r["bound"] == "fp32_add" and r["unit"] == "Tadd/s"
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
    assert r["kernel_ms"] <= d["ms_per_step"] * 1.001          # the kernel is part of the step
    assert r["peak"] == pytest.approx(148 * 128 * d["clocks"]["sm_max_mhz"] * 1e6 / 1e12, rel=1e-3)
    if n == 1:
        assert r["algorithmic_adds_per_launch"] == float(M) * nnz
        assert r["achieved"] == pytest.approx(r["algorithmic_adds_per_launch"] / (r["kernel_ms"] * 1e-3) / 1e12, rel=2e-3)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] == pytest.approx(jobs * per_call / (e["ms_per_step"] * 1e-3) / 1e9, rel=2e-3)
    assert e["value"] < d["value"]                              # copies inside the timed region can only cost time
    assert d["clocks"]["reasons"] == [] or set(d["clocks"]["reasons"]) <= {"sw_power_cap"}
    assert d["gpu_launches"] >= d["steps"]


def test_traffic_matches_the_kernel_sources():
    """profiles/traffic.json is only quoted by bench.py while it was captured from the kernel sources in the tree"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["cfg2"]
    if t["kernel_source_id"] != b.kernel_source_id():  # bench.py drops the figure from its line in that case; nothing is misquoted
        pytest.skip("kernel sources changed since the capture: re-run tools/profile_r02.sh and update profiles/traffic.json")
    assert t["dram_bytes_per_launch"] == t["dram_bytes_read"] + t["dram_bytes_write"]
    line = _load(os.path.join(ROOT, "profiles", "bench_r02_n1_cfg2.json"))
    assert line["roofline"]["traffic"] == t["dram_bytes_per_launch"]
