"""bench.py's JSON contract, checked on the CPU through the reference arm (the only arm that runs without a GPU)."""
import json
import os
import subprocess
import sys

import __graft_entry__ as ge


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ge.ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout  # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s-equiv" and d["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    if d["cpu_baseline"]["cores"] > 1:  # all host threads: only the reference's OpenMP form of the path can use them
        assert "omp" in d["details"]["function"] and "-fopenmp" in d["cpu_baseline"]["build"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["value"] > 0 and "workload" in d["config"]
    # `config` carries workload keys only, identical on both arms (bench.py workload_config): arm-specific facts live in `details`
    assert set(d["config"]) == {"workload", "M", "K", "N", "sparsity", "nnz", "alpha", "seeds"}


def test_reference_arm_single_thread_option():
    out = subprocess.run([sys.executable, os.path.join(ge.ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--steps", "1", "--warmup", "1",
                          "--ref-threads", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == 1 and "tcsc_sgemm_prelu_basic" in d["details"]["function"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ge.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
