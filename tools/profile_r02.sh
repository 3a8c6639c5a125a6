# round-2 evidence run (one GPU): tests, the default bench line, the ncu launch list of a short bench, --set full captures, sweeps
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_r02.log
python bench.py > gpurun_out/bench_r02_n1_cfg2.json 2> gpurun_out/bench_r02_n1_cfg2.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r02_reference_cfg2.json 2> gpurun_out/bench_r02_reference_cfg2.err
python tools/sweep.py --bcsr > gpurun_out/sweep_r02_cfg3.csv 2> gpurun_out/sweep_r02_cfg3_conversion.txt
python tools/sweep.py --pattern window --Ms 1,32,256,4096 > gpurun_out/sweep_r02_cfg3_window.csv 2> gpurun_out/sweep_r02_cfg3_window.err
python tools/sweep.py --pattern skewed --Ms 1,32,256,4096 > gpurun_out/sweep_r02_cfg3_skewed.csv 2> gpurun_out/sweep_r02_cfg3_skewed.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r02_bench_cfg2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_target.py > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tcsc_gemm -s 2 -c 1 -o gpurun_out/prof_r02_cfg2 python tools/ncu_target.py > gpurun_out/ncu_full.log 2>&1
python tools/ncu_target_bcsr.py > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none -k regex:k_bcsr_gemm_ring -s 1 -c 1 -o gpurun_out/prof_r02_bcsr_ring python tools/ncu_target_bcsr.py > gpurun_out/ncu_bcsr.log 2>&1
tail -2 gpurun_out/pytest_gpu_r02.log; tail -c 300 gpurun_out/bench_r02_n1_cfg2.err; wc -c gpurun_out/bench_r02_n1_cfg2.json gpurun_out/sweep_r02_*.csv; tail -2 gpurun_out/ncu_full.log gpurun_out/ncu_bcsr.log
