#!/usr/bin/env python
"""out.txt (the legacy lines the reference's SparseGEMM.cpp driver prints, SparseGEMM.cpp:91,182-198) -> CSV with the header of
the reference's parse-out2csv.sh:3, so that performance.py:10-44 reads it unchanged with np.genfromtxt(names=True), plus
columns of our own behind the reference's sixteen:

  speedup_sGEMM, speedup_sGEMM_PReLU     cycles of the CPU dense GEMM / cycles of the sparse call (same driver run)
  nnz_model                              K * (N // nonZero): what generateSparseMatrix emits (SparseGEMM.h:53-102)
  us_sGEMM_PReLU, gflops_equiv_sGEMM_PReLU, roofline_us, roofline_frac   (only with --tsc-ghz: cycles are host TSC cycles)
        roofline_us = max(algorithmic bytes / HBM GB/s, M*nnz adds / FP32-add peak)  (SURVEY.md 8d), frac = roofline_us / us

The reference's flops columns come from PAPI; built with -DDISABLE_PAPI (the only way it builds here) they are 0
and so is performance = flops/cycles.  --model-flops (default) fills them with the reference's own FLOP model instead
(main.cpp:47-51: 2*M*nnz + M*N for the sparse calls, 2*M*K*N + M*N for the dense ones) and recomputes performance;
--raw keeps the zeros.

    ./oracle/_ref/ref_sparsegemm_on_b200 | tee out.txt ; python tools/legacy_csv.py out.txt --tsc-ghz 2.0 > out.csv
"""
import argparse
import re
import sys

HEADER = ("M,K,N,nonZero,cycles_GEMM,flops_GEMM,performance_GEMM,cycles_sGEMM,flops_sGEMM,performance_sGEMM,"
          "cycles_GEMM_PReLU,flops_GEMM_PReLU,performance_GEMM_PReLU,cycles_sGEMM_PReLU,flops_sGEMM_PReLU,performance_sGEMM_PReLU")
KEYS = ["GEMM", "sGEMM", "GEMM_PReLU", "sGEMM_PReLU"]
EXTRA = ["speedup_sGEMM", "speedup_sGEMM_PReLU", "nnz_model", "us_sGEMM_PReLU", "gflops_equiv_sGEMM_PReLU", "roofline_us", "roofline_frac"]


def rows(path, model_flops=True, tsc_ghz=None, hbm_gbs=6534.8, fadd_tadds=37.22):
    shape = re.compile(r"M=(\d+),\s*K=(\d+),\s*N=(\d+),\s*nonZero=(\d+)")
    line = re.compile(r"^(s?GEMM(?:_PReLU)?)\s+cycles=([\d.]+),\s*flops=([\d.]+),\s*performance=([\d.]+)")
    cur, vals = None, {}
    for raw in open(path, errors="ignore"):
        raw = raw.strip()
        m = shape.search(raw)
        if m:
            cur, vals = m.groups(), {}
            continue
        m = line.match(raw)
        if not (m and cur):
            continue
        vals[m.group(1)] = list(m.groups()[1:])
        if len(vals) < 4:
            continue
        M, K, N, nz = (int(x) for x in cur)
        nnz = K * (N // nz)
        row = list(cur)
        for k in KEYS:
            cyc, fl, perf = vals[k]
            if model_flops and float(fl) == 0.0:
                f = (2 * M * nnz + M * N) if k.startswith("s") else (2 * M * K * N + M * N)
                fl, perf = str(f), f"{f / float(cyc):.4f}"
            row += [cyc, fl, perf]
        row.append(f"{float(vals['GEMM'][0]) / float(vals['sGEMM'][0]):.2f}")
        row.append(f"{float(vals['GEMM_PReLU'][0]) / float(vals['sGEMM_PReLU'][0]):.2f}")
        row.append(str(nnz))
        if tsc_ghz:
            us = float(vals["sGEMM_PReLU"][0]) / (tsc_ghz * 1e3)
            bytes_alg = 4.0 * M * K + 4.0 * M * N + 4.0 * nnz + 8.0 * (N + 1) + 4.0 * N
            roof_us = max(bytes_alg / (hbm_gbs * 1e3), M * nnz / (fadd_tadds * 1e6))
            row += [f"{us:.3f}", f"{(2.0 * M * nnz + M * N) / us / 1e3:.3f}", f"{roof_us:.4f}", f"{roof_us / us:.5f}"]
        else:
            row += ["nan"] * 4
        yield row
        cur = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path", nargs="?", default="out.txt")
    ap.add_argument("--raw", action="store_true", help="keep flops=0 / performance=0 as printed by a -DDISABLE_PAPI build")
    ap.add_argument("--tsc-ghz", type=float, default=None, help="host TSC frequency: turns cycles into time and fills the roofline columns")
    ap.add_argument("--hbm-gbs", type=float, default=6534.8)
    ap.add_argument("--fadd-tadds", type=float, default=37.22)
    a = ap.parse_args()
    print(HEADER + "," + ",".join(EXTRA))
    for r in rows(a.path, not a.raw, a.tsc_ghz, a.hbm_gbs, a.fadd_tadds):
        print(",".join(r))


if __name__ == "__main__":
    main()
