#!/usr/bin/env python
"""out.txt (the legacy lines the reference's SparseGEMM.cpp driver prints, SparseGEMM.cpp:91,182-198) -> CSV with the
header of the reference's parse-out2csv.sh:3 (so performance.py:10-44 can read it with np.genfromtxt(names=True)), plus
three columns of our own: speed-up of sGEMM over the CPU dense GEMM for both variants and host TSC GHz if given.

    ./oracle/_ref/ref_sparsegemm_on_b200 | tee out.txt ; python tools/legacy_csv.py out.txt > out.csv
"""
import re
import sys

HEADER = ("M,K,N,nonZero,cycles_GEMM,flops_GEMM,performance_GEMM,cycles_sGEMM,flops_sGEMM,performance_sGEMM,"
          "cycles_GEMM_PReLU,flops_GEMM_PReLU,performance_GEMM_PReLU,cycles_sGEMM_PReLU,flops_sGEMM_PReLU,performance_sGEMM_PReLU")
KEYS = ["GEMM", "sGEMM", "GEMM_PReLU", "sGEMM_PReLU"]


def main(path):
    shape = re.compile(r"M=(\d+),\s*K=(\d+),\s*N=(\d+),\s*nonZero=(\d+)")
    line = re.compile(r"^(s?GEMM(?:_PReLU)?)\s+cycles=([\d.]+),\s*flops=([\d.]+),\s*performance=([\d.]+)")
    print(HEADER + ",speedup_sGEMM,speedup_sGEMM_PReLU")
    cur, vals = None, {}
    for raw in open(path, errors="ignore"):
        raw = raw.strip()
        m = shape.search(raw)
        if m:
            cur, vals = m.groups(), {}
            continue
        m = line.match(raw)
        if m and cur:
            vals[m.group(1)] = m.groups()[1:]
            if len(vals) == 4:
                row = list(cur)
                for k in KEYS:
                    row += list(vals[k])
                row.append(f"{float(vals['GEMM'][0]) / float(vals['sGEMM'][0]):.2f}")
                row.append(f"{float(vals['GEMM_PReLU'][0]) / float(vals['sGEMM_PReLU'][0]):.2f}")
                print(",".join(row))
                cur = None


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "out.txt")
