for i in 1 2; do python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-secondary > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1b.json'))
print({k:d[k] for k in ('value','ms_per_step','verified')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['clocks'])"; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
