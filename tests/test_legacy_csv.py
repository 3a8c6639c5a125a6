"""the result pipeline.  tools/legacy_csv.py turns the legacy lines of the reference's SparseGEMM.cpp
driver (here: the driver's own output on the B200 box, committed under profiles/) into the CSV the reference's plotter reads.
The check reads the CSV exactly the way performance.py:10-22 does and touches every column that script uses."""
import os
import subprocess
import sys

import numpy as np

import __graft_entry__ as ge

SRC = os.path.join(ge.ROOT, "profiles", "ref_sparsegemm_on_b200_r01.txt")
REF_HEADER = ("M,K,N,nonZero,cycles_GEMM,flops_GEMM,performance_GEMM,cycles_sGEMM,flops_sGEMM,performance_sGEMM,"
              "cycles_GEMM_PReLU,flops_GEMM_PReLU,performance_GEMM_PReLU,cycles_sGEMM_PReLU,flops_sGEMM_PReLU,performance_sGEMM_PReLU")  # parse-out2csv.sh:3


def _csv(tmp_path, *extra):
    out = subprocess.run([sys.executable, os.path.join(ge.ROOT, "tools", "legacy_csv.py"), SRC, *extra], capture_output=True, text=True, check=True).stdout
    p = tmp_path / "out.csv"
    p.write_text(out)
    return p, out


def test_header_and_reader_of_the_reference_plotter(tmp_path):
    p, out = _csv(tmp_path, "--tsc-ghz", "2.0")
    assert out.splitlines()[0].startswith(REF_HEADER + ",")  # the reference's sixteen columns first, ours behind them
    m = np.genfromtxt(p, delimiter=",", names=True)          # performance.py:10
    assert m.shape == (27,)                                   # SparseGEMM.cpp:74-80: 3 sparsities x 3 x 3 shapes
    nz = np.unique(m["nonZero"])                              # performance.py:12
    assert nz.tolist() == [2.0, 8.0, 16.0]
    for z in nz:
        d = np.sort(m[np.where(m["nonZero"] == z)], order=["flops_GEMM"])  # performance.py:18-20
        assert len(d) == 9 and np.all(np.diff(d["flops_GEMM"]) >= 0)
        _ = d[["M", "K", "N"]]
        for col in ("performance_GEMM", "performance_sGEMM", "performance_GEMM_PReLU", "performance_sGEMM_PReLU"):  # performance.py:27-44
            assert np.all(np.isfinite(d[col])) and np.all(d[col] > 0)
    # the FLOP model that replaces PAPI's zeros: main.cpp:47-51 / SparseGEMM.cpp's shapes
    r = m[0]
    nnz = r["K"] * (r["N"] // r["nonZero"])
    assert r["flops_sGEMM"] == 2 * r["M"] * nnz + r["M"] * r["N"] and r["flops_GEMM"] == 2 * r["M"] * r["K"] * r["N"] + r["M"] * r["N"]
    assert np.isclose(r["performance_sGEMM"], r["flops_sGEMM"] / r["cycles_sGEMM"], rtol=1e-3)
    # roofline columns
    assert np.all((m["roofline_frac"] > 0) & (m["roofline_frac"] < 1)) and np.all(m["roofline_us"] > 0)
    assert np.allclose(m["roofline_frac"], m["roofline_us"] / m["us_sGEMM_PReLU"], rtol=1e-2)


def test_raw_mode_keeps_the_driver_numbers(tmp_path):
    p, _ = _csv(tmp_path, "--raw")
    m = np.genfromtxt(p, delimiter=",", names=True)
    assert np.all(m["flops_sGEMM"] == 0) and np.all(m["performance_sGEMM"] == 0)  # -DDISABLE_PAPI
    assert np.all(np.isnan(m["roofline_frac"]))                                   # no TSC frequency given
