# round-2 closing evidence run (one GPU), after the BCSR ring changes: tests, the default bench line, the BCSR sweep, one BCSR capture.
# (The TCSC captures of tools/profile_r02.sh stay valid: kernel_source_id unchanged.)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_r02.log
python bench.py > gpurun_out/bench_r02_n1_cfg2.json 2> gpurun_out/bench_r02_n1_cfg2.err
python tools/sweep.py --bcsr > gpurun_out/sweep_r02_cfg3.csv 2> gpurun_out/sweep_r02_cfg3_conversion.txt
python tools/ncu_target_bcsr.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:k_bcsr_gemm_ring -s 1 -c 1 -o gpurun_out/prof_r02_bcsr_ring_final python tools/ncu_target_bcsr.py > gpurun_out/ncu_bcsr.log 2>&1
tail -2 gpurun_out/pytest_gpu_r02.log; tail -c 300 gpurun_out/bench_r02_n1_cfg2.err; wc -c gpurun_out/bench_r02_n1_cfg2.json gpurun_out/sweep_r02_cfg3.csv; tail -2 gpurun_out/ncu_bcsr.log
