#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native sparse ternary GEMM  Y = PReLU(X*W + b)  (TCSC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4|cfg1|cfg5] [--dist-mode 0..5]
                    [--no-secondary] [--no-cpu-baseline]

Prints ONE JSON line on rank 0 (contract in the task statement).  Metric = BASELINE.json's "sparse GEMM GFLOP/s-equiv"
under the reference's own FLOP model 2*M*nnz + M*N (main.cpp:47-51).

N = 1   workload = BASELINE.json configs[1]: TCSC sparseGEMM+bias+PReLU, M=K=N=4096, 90 % sparsity, synthetic ternary W
        (P(+1)=P(-1)=5 %), X,b ~ U[-1,1), a = 0.2; inputs resident in HBM; four X/Y buffer sets (512 MiB) are rotated so
        that every step streams X and Y from HBM rather than from the 126 MB L2.
N > 1   the column-partitioned path of north_star (3), weak scaling: every GPU owns 4096 columns of a 4096 x (4096*N) W,
        X is broadcast from rank 0 and every rank ends with the full M x (4096*N) Y (--dist-mode: 0 ncclAllGather +
        re-layout, 1 per-lane peer stores, 2 copy engines gated by progress counters, 3/4 TMA bulk stores from staged output
        tiles, 5 (default) one multimem.st write per tile to the NVSwitch multicast mapping of Y; see include/tsgemm_b200.h).  After the timed region EVERY rank checks its full Y: every column slab bitwise
        against a single-GPU recompute of that slab, and a row slice against an fp64 dense evaluation ("verified"); a
        mismatch makes every rank exit non-zero.  At N = 1 this degenerates to the workload above.
`value`   whole-job throughput, device-timed (CUDA events), barrier + synchronize on both sides, max over ranks.
`e2e`     the same metric through the reference-named C entry point tcsc_sgemm_prelu_basic with HOST (pinned) buffers:
          host->device copy of X and b and device->host copy of Y inside the timed region, every step.
`roofline` dominant kernel k_tcsc_gemm: achieved = M*nnz gather-adds / its CUDA-event time measured inside the timed
          region, against the FP32-add peak (#SM x 128 lanes x max SM clock) -- the path is FP32-add/shared-memory bound
          at this shape, not HBM bound; the HBM view (algorithmic bytes / time vs the measured copy
          bandwidth of MEASURED_PEAKS.json) and the shared-memory gather ceiling are reported next to it.
`secondary` (N = 1) the other BASELINE.json configs and the sweep corners, each with its own device time, algorithmic
          units and roofline fraction, measured outside the headline timed region (tools/secondary.py).
`cpu_baseline` the unmodified reference (oracle/_ref, kind "reference"; else the oracle port) timed on this box's host
          cores on a bounded row sample of the same workload.
          `cores` = 1: tcsc_sgemm_prelu_basic has no threading and north_star names the single-threaded build; the
          reference's OpenMP form on all host threads is reported beside it as `all_threads`.
--impl reference   times the reference's own CPU implementation of the path with all the host threads it can use:
          sparseGEMM_PReLU<float> (SparseGEMM.h:151-168, omp parallel for over m) from the -fopenmp build of the
          unmodified reference (--ref-threads 1: tcsc_sgemm_prelu_basic on one pinned core), on bounded row samples of
          the same workload; prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (M, K, N, num, den, description); N is per GPU for "weak" workloads, total for "strong" ones
    "cfg2": (4096, 4096, 4096, 1, 10, "TCSC sparseGEMM+PReLU M=4096 K=4096 N=4096 90% sparsity (BASELINE.json configs[1])"),
    "cfg1": (64, 512, 512, 1, 2, "TCSC sparseGEMM+bias+PReLU M=64 K=512 N=512 50% sparsity (BASELINE.json configs[0])"),
    "cfg4": (8192, 4096, 14336, 1, 3, "ternary LLM-layer shape M=8192 K=4096 N=14336 66% sparsity (BASELINE.json configs[3])"),
    "cfg5": (16384, 16384, 16384, 1, 10, "large sweep M=K=N=16384 90% sparsity, dense->TCSC conversion included in every step (BASELINE.json configs[4])"),
}
STRONG = {"cfg4", "cfg5"}       # total N fixed, columns split across the GPUs
CONVERT_IN_STEP = {"cfg5"}      # the step re-converts the rank's dense slice (and rebuilds the gather stream) every time
ALPHA = 0.2  # main.cpp:268
SEED_W, SEED_X, SEED_B = 42, 43, 44
METRIC = "sparse GEMM GFLOP/s-equiv (2*M*nnz + M*N per call, TCSC+bias+PReLU fp32)"
UNIT = "GFLOP/s-equiv"


def workload_config(workload, gpus, nnz):
    """The `config` object: identical on both arms (same keys, same values) -- arm-specific facts go into `details`."""
    M, K, Ng, num, den, desc = WORKLOADS[workload]
    strong = workload in STRONG
    N = Ng if strong else Ng * max(1, gpus)
    if gpus > 1:
        desc += (f"; N={N} columns split over {gpus} GPUs" if strong else f"; per GPU: {gpus} x {Ng} = {N} columns") + \
                ", X broadcast from rank 0, Y all-gathered"
    return {"workload": desc, "M": M, "K": K, "N": N, "sparsity": 1 - num / den, "nnz": nnz, "alpha": ALPHA,
            "seeds": {"W": SEED_W, "X": SEED_X, "b": SEED_B}}


def flops_equiv(M, N, nnz):
    return 2.0 * M * nnz + 1.0 * M * N  # main.cpp:47-51


def algorithmic_bytes(M, K, N, nnz):
    return 4.0 * M * K + 4.0 * M * N + 4.0 * nnz + 8.0 * (N + 1) + 4.0 * N


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "MEASURED_PEAKS.json"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ---- clocks / throttle reasons during the timed region ------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self._stop = index, [], set(), threading.Event()
        self.max_mhz, self.thread, self.ok = None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((mhz, util))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.thread:
            self._stop.set()
            self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        busy = [m for m, u in self.samples if u >= 50] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# =====================================================================================================================
# reference arm / CPU baseline
# =====================================================================================================================
def cpu_reference_backend():
    from oracle import pyoracle
    if pyoracle.ref_available("native") and pyoracle.ref_native_runs():
        r = pyoracle.Ref("native")
        return r, "reference", r.build_flags
    if pyoracle.ref_available(""):
        r = pyoracle.Ref("")
        return r, "reference", r.build_flags + " (portable build: the -march=native objects do not run on this host)"
    return pyoracle.Port(), "port", "gcc -O2 -ffp-contract=off (oracle/tsg_oracle.c restatement; oracle/_ref absent)"


def pin_one_core():
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[-1]})
    except Exception:
        pass


def cpu_pick_rows(backend, W, M, K, N, seconds_target, port):
    """Rows of the workload one timed call should cover so that it lasts about `seconds_target` (probe: 8 rows)."""
    B = port.gen_uniform((N,), SEED_B)
    probe = min(M, 8)
    Xp = port.gen_uniform((probe, K), SEED_X)  # counter-based generator: rows 0..probe-1 of the workload's X
    t_probe = backend.time_prelu_basic(Xp, W, B, ALPHA, reps=2)
    rows = int(max(probe, min(M, seconds_target / max(t_probe / probe, 1e-9))))
    return min(rows, M)


def cpu_time_rows(backend, W, rows, K, N, port):
    """One call of tcsc_sgemm_prelu_basic (sparse/tcsc.c:143-165) on rows 0..rows-1 of the workload; seconds."""
    X = port.gen_uniform((rows, K), SEED_X)
    B = port.gen_uniform((N,), SEED_B)
    return backend.time_prelu_basic(X, W, B, ALPHA, reps=1)


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def omp_reference_runs():
    """The OpenMP build of the reference loads and runs on this host (probed in a subprocess, like the native build)."""
    from oracle import pyoracle
    if not pyoracle.ref_available("omp"):
        return False
    code = ("import sys; sys.path.insert(0, %r); import numpy as np; from oracle.pyoracle import Ref; r = Ref('omp');"
            "W = r.tcsc_from_dense(np.eye(64, dtype=np.float32)); r.time_sparse_gemm_prelu(np.ones((8, 64), np.float32), W, np.zeros(64, np.float32), 0.2)"
            ) % ROOT
    try:
        return subprocess.run([sys.executable, "-c", code], capture_output=True, timeout=120).returncode == 0
    except Exception:
        return False


def run_reference(args):
    """The reference's own CPU implementation of the path with all the host threads it can use: sparseGEMM_PReLU<float>
    (SparseGEMM.h:151-168, `#pragma omp parallel for` over m) from the -fopenmp build of the unmodified reference; where that
    build is absent, tcsc_sgemm_prelu_basic (sparse/tcsc.c:143-165), which has no threading, on one pinned core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pyoracle
    from oracle.pyoracle import Port
    port = Port()
    M, K, Ng, num, den, desc = WORKLOADS[args.workload]
    N = Ng if args.workload in STRONG else Ng * max(1, args.gpus)
    threads = 1 if args.ref_threads == 1 else host_threads()
    use_omp = threads > 1 and omp_reference_runs()
    if use_omp:
        os.environ["OMP_NUM_THREADS"] = str(threads)  # read by libgomp when the library is loaded below
        os.environ.setdefault("OMP_PROC_BIND", "true")
        backend = pyoracle.Ref("omp")
        kind, flags = "reference", backend.build_flags
        threads = backend.omp_max_threads()
        function = "sparseGEMM_PReLU<float> (SparseGEMM.h:151-168, omp parallel for over m)"
        timer = backend.time_sparse_gemm_prelu
    else:
        backend, kind, flags = cpu_reference_backend()
        threads = 1
        pin_one_core()  # one thread pinned to one core, as benchmark.sh:36 does
        function = "tcsc_sgemm_prelu_basic (sparse/tcsc.c:143-165)"
        timer = backend.time_prelu_basic
    Wd = port.gen_ternary(K, N, SEED_W, num, den)
    W = backend.tcsc_from_dense(Wd)
    nnz = W.nnz
    B = port.gen_uniform((N,), SEED_B)
    # bounded sample per step so that (steps + warmup) steps end within ~2 minutes: probe, then scale the row count
    per_step = max(0.5, min(6.0, 110.0 / max(1, args.steps + args.warmup)))
    probe = min(M, 8 * threads)
    t_probe = timer(port.gen_uniform((probe, K), SEED_X), W, B, ALPHA, reps=2)
    rows = min(M, int(max(probe, per_step / max(t_probe / probe, 1e-9))))
    X = port.gen_uniform((rows, K), SEED_X)  # counter-based generator: rows 0..rows-1 of the workload's X
    secs_list = []
    for i in range(args.warmup + args.steps):
        secs = timer(X, W, B, ALPHA, reps=1)
        if i >= args.warmup:
            secs_list.append(secs)
    mean_s = sum(secs_list) / len(secs_list)
    value = flops_equiv(rows, N, nnz) / mean_s / 1e9
    sample = (f"rows 0..{rows - 1} of {M} (all {N} columns, K={K}), one call per step, {threads} thread(s)"
              + ("" if use_omp else " pinned") + "; the m-outer loop nest is linear in M")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, nnz),
        "details": {"function": function},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample, "build": flags,
                         "host_cores_available": os.cpu_count(),
                         "seconds_for_full_M_extrapolated": mean_s * M / rows},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
def kernel_source_id():
    """sha256 over the sources of the dominant kernel: ties an ncu-derived number (profiles/traffic.json) to the code it was
    captured from, so that it is dropped -- not silently kept -- once the kernel changes"""
    import hashlib
    h = hashlib.sha256()
    for f in ("gemm_tcsc.cu", "ktformat.cu", "tsg_ptx.cuh", "tsg_internal.h"):
        with open(os.path.join(ROOT, "sparse-matrix-multiplication-benchmark_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def verify_full_y(t, torch, dist, D, world, rank, Y, X, B, M, K, N, num, den, mode):
    """After the timed region every rank checks the FULL Y it holds.  (1) bitwise: every rank's column slab against a
    single-GPU recompute of that slab on THIS GPU (same kernel, same order => same bits; catches wrong, torn or missing
    peer stores); (2) an fp64 dense evaluation of a row slice over all N columns (catches a wrong recompute too).
    Returns the dict that goes into the JSON line; `ok` is all-reduced so that every rank agrees."""
    torch.cuda.synchronize()
    bad_slabs, checked_cols = [], 0
    for p in range(world):
        c0, nc = t.partition(N, p, world) if world > 1 else (0, N)
        if nc == 0:
            continue
        Wp = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, SEED_W, num, den) if world > 1 else t.gen_ternary(K, N, SEED_W, num, den))
        Yp = torch.empty((M, nc), device="cuda")
        Wp.gemm(X, B[c0:c0 + nc].contiguous(), Yp, a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST)
        torch.cuda.synchronize()
        if not torch.equal(Yp, Y[:, c0:c0 + nc]):
            bad_slabs.append(p)
        checked_cols += nc
        Wp.destroy()
        del Yp
    f64_rows = min(M, 32)
    m0 = ((rank * 997) % max(1, M - f64_rows + 1))  # a different row slice on every rank
    Wfull = t.gen_ternary(K, N, SEED_W, num, den)
    rel, absd = t.verify_dense_f64(X, Wfull, B, Y, a=ALPHA, use_prelu=True, m0=m0, mrows=f64_rows)
    del Wfull
    ok_local = (not bad_slabs) and rel <= 1e-4  # the reference's own fp32 order is ~1.5e-5 from fp64 at K=4096
    flag = torch.tensor([0 if ok_local else 1], device="cuda", dtype=torch.int32)
    relt = torch.tensor([rel], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(flag)
        dist.all_reduce(relt, op=dist.ReduceOp.MAX)
    return {"verified": int(flag.item()) == 0, "verified_rows": M, "verified_cols": checked_cols, "ranks_checking": world,
            "method": "every rank: each column slab of its full Y bitwise == single-GPU recompute of that slab; fp64 dense check of a row slice",
            "f64_rows_per_rank": f64_rows, "f64_max_rel_err": float(relt.item()), "bad_slabs_on_rank0": bad_slabs, "dist_mode": mode}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != max(1, args.gpus):
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version there) go to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    t = ge.load()
    t.lib()  # raises if libtsgemm_b200.so is missing: there is no fallback
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    t.use_torch_stream()

    M, K, Ng, num, den, desc = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    N = Ng if strong else Ng * world
    hbm_peak, sm_max_mhz, peak_src = measured_peaks()

    # ---- W: this rank's column slice, generated and converted on the device ----
    D = None
    if world > 1:
        D = t.Dist(rank, world)
        col0, ncols = D.partition(N)
    else:
        col0, ncols = 0, N
    Wd = t.gen_ternary_slice(K, N, col0, ncols, SEED_W, num, den) if world > 1 else t.gen_ternary(K, N, SEED_W, num, den)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t.DeviceTcsc.from_dense(Wd).destroy()  # first call pays module load + pool growth; report the steady state
    torch.cuda.synchronize()
    ev[0].record()
    W = t.DeviceTcsc.from_dense(Wd)
    ev[1].record()
    info = W.stream_info()
    ev[2].record()
    torch.cuda.synchronize()
    convert_ms, stream_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    nnz_local = W.nnz
    nnz_t = torch.tensor([nnz_local], device="cuda", dtype=torch.int64)
    if world > 1:
        dist.all_reduce(nnz_t)
    nnz = int(nnz_t.item())

    # ---- inputs resident in HBM; rotate buffer sets so that each step streams from HBM, not L2 ----
    nsets = 4 if (M * K + M * N) * 4 * 4 <= (8 << 30) else (2 if (M * K + M * N) * 4 * 2 <= (24 << 30) else 1)
    Xs = [t.gen_uniform((M, K), SEED_X + 100 * i) for i in range(nsets)]
    B = t.gen_uniform((N,), SEED_B)
    if world > 1 and args.dist_mode >= 1:
        Ys = [D.alloc_y(M, N)]  # symmetric buffer for the fused all-gather
        if args.dist_mode == 5 and not D.has_multicast():  # the same answer on every rank (tsg_dist_alloc_y ends in a consensus)
            args.dist_mode = 3
    else:
        Ys = [torch.empty((M, N), device="cuda") for _ in range(nsets)]

    convert_in_step = args.workload in CONVERT_IN_STEP
    state = {"W": W}

    def step(i):
        X, Y = Xs[i % len(Xs)], Ys[i % len(Ys)]
        if convert_in_step:  # configs[4]: dense -> TCSC (+ private gather stream) is part of the measured step
            state["W"].destroy()
            state["W"] = t.DeviceTcsc.from_dense(Wd)
        W = state["W"]
        if world > 1:
            D.gemm(W, X, B, Y, N, a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST, root=0, mode=args.dist_mode)
        else:
            W.gemm(X, B, Y, a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t.profile_enable(True)
    t.profile_read()
    t.launch_count(reset=True)
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
    ev[1].record()
    barrier()
    launches = t.launch_count(reset=True)
    kern_ms_total, kern_launches = t.profile_read()
    t.profile_enable(False)
    elapsed_ms = ev[0].elapsed_time(ev[1])
    el = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(el.item())
    ms_per_step = elapsed_ms / args.steps
    value = flops_equiv(M, N, nnz) / (ms_per_step * 1e-3) / 1e9
    clocks = sampler.stop()

    # ---- the result of the LAST timed step is checked on every rank (world > 1: X was broadcast in place, so every rank
    #      holds rank 0's X of that step) ----
    last = args.warmup + args.steps - 1
    verification = verify_full_y(t, torch, dist, D, world, rank, Ys[last % len(Ys)], Xs[last % len(Xs)], B, M, K, N, num, den, args.dist_mode)

    # ---- roofline of the dominant kernel (this rank's launches) ----
    kern_ms = kern_ms_total / max(1, kern_launches)
    kern_ms_t = torch.tensor([kern_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(kern_ms_t, op=dist.ReduceOp.MAX)
    kern_ms_max = float(kern_ms_t.item())
    adds_per_launch = float(M) * nnz_local
    fadd_peak = 148 * 128 * sm_max_mhz * 1e6 / 1e12  # Tadd/s
    achieved_tadd = adds_per_launch / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
    bytes_per_launch = algorithmic_bytes(M, K, ncols, nnz_local)
    src_id = kernel_source_id()
    traffic, traffic_note = None, "no ncu capture on file for this kernel source"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            rec = json.load(open(tpath)).get(args.workload, {})
            if rec.get("kernel_source_id") == src_id:
                traffic, traffic_note = rec.get("dram_bytes_per_launch"), rec.get("source", "profiles/traffic.json")
            elif rec:
                traffic_note = f"profiles/traffic.json was captured from kernel source {rec.get('kernel_source_id')}, this build is {src_id}: dropped"
        except Exception:
            pass
    roofline = {
        "bound": "fp32_add", "kernel": "k_tcsc_gemm", "kernel_ms": kern_ms, "kernel_ms_max_over_ranks": kern_ms_max, "kernel_launches_timed": kern_launches,
        "achieved": achieved_tadd, "peak": fadd_peak, "unit": "Tadd/s", "frac": achieved_tadd / fadd_peak,
        "peak_source": f"#SM(148) x 128 FP32 lanes x {sm_max_mhz:.0f} MHz (max SM clock, {peak_src}); microbenchmark: 0.99 of it reached by a register-only FADD loop",
        "algorithmic_adds_per_launch": adds_per_launch, "traffic": traffic, "traffic_note": traffic_note, "kernel_source_id": src_id,
        "hbm": {"bound": "hbm", "achieved": bytes_per_launch / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0, "peak": hbm_peak, "unit": "GB/s",
                "frac": (bytes_per_launch / (kern_ms * 1e-3) / 1e9) / hbm_peak if kern_ms > 0 else 0.0, "algorithmic_bytes_per_launch": bytes_per_launch,
                "peak_source": peak_src},
        "smem_gather": {"achieved": achieved_tadd, "peak": fadd_peak / 4, "unit": "Tadd/s", "frac": achieved_tadd / (fadd_peak / 4),
                        "note": "one LDS word per add: 128 B/clk/SM shared-memory crossbar = 1/4 of the FP32-add peak; TMEM, L1 and "
                                "register-reuse alternatives measured in profiles/microbench/ubench2_r02_b200.jsonl"},
    }

    # ---- e2e: host buffers through the public entry points ----
    e2e = None
    cpu_baseline = None
    Wd_host = None
    if world == 1:
        Wd_host = Wd.cpu().numpy()
        Wh = t.tcsc_from_dense(Wd_host)  # host tcsc_t + cached device mirror, as a reference caller would hold it
        Xh = [torch.empty((M, K), dtype=torch.float32).pin_memory() for _ in range(2)]
        Yh = [torch.empty((M, N), dtype=torch.float32).pin_memory() for _ in range(2)]
        for i in range(2):
            Xh[i].copy_(Xs[i].cpu())
        Bh = B.cpu().numpy()
        e2e_steps = max(3, min(args.steps, 20))

        def timed(Xl, Yl, steps):
            for i in range(2):
                t.tcsc_sgemm_prelu_basic(Xl[i % 2], Wh, Bh, ALPHA, Y=Yl[i % 2])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(steps):
                t.tcsc_sgemm_prelu_basic(Xl[i % 2], Wh, Bh, ALPHA, Y=Yl[i % 2])  # returns with Y in host memory
            return (time.perf_counter() - t0) * 1e3 / steps

        e2e_ms = timed([x.numpy() for x in Xh], [y.numpy() for y in Yh], e2e_steps)
        W.gemm(Xs[0], B, Ys[0], a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST)
        same = bool(torch.equal(Yh[0], Ys[0].cpu()))  # host-pointer path and device-resident path: same bits
        # the reference's callers pass pageable posix_memalign memory (main.cpp:38-44): the same call, unpinned buffers
        Xp = [np.array(x.numpy(), copy=True) for x in Xh]
        Yp = [np.empty((M, N), np.float32) for _ in range(2)]
        pageable_ms = timed(Xp, Yp, max(3, e2e_steps // 2))
        same_pageable = bool(np.array_equal(Yp[0], Yh[0].numpy()))
        del Xp, Yp
        # fixed cost of a small host-pointer call: the reference driver's first shape (main.cpp:259: M=1, K=512, N=2048, 50 %)
        sm_, sk_, sn_ = 1, 512, 2048
        Wsm_d = t.gen_ternary(sk_, sn_, SEED_W, 1, 2).cpu().numpy()
        Wsm = t.tcsc_from_dense(Wsm_d)
        xs_, bs_, ys_ = Xh[0].numpy()[:sm_, :sk_].copy(), Bh[:sn_].copy(), np.empty((sm_, sn_), np.float32)
        for _ in range(20):
            t.tcsc_sgemm_prelu_basic(xs_, Wsm, bs_, ALPHA, Y=ys_)
        reps = 300
        fn, args_c = t.lib().tcsc_sgemm_prelu_basic, (xs_.ctypes.data, Wsm.handle, bs_.ctypes.data, ALPHA, ys_.ctypes.data, sm_, sn_, sk_)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn(*args_c)  # the C entry point itself: no Python wrapper work inside the loop
        small_us = (time.perf_counter() - t0) * 1e6 / reps
        Wsm.free()
        e2e = {"matches_device_resident_result": same, "value": flops_equiv(M, N, nnz) / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_bytes_per_step": 4 * M * K + 4 * N, "d2h_bytes_per_step": 4 * M * N,
               "api": "tcsc_sgemm_prelu_basic(X_host, W, B_host, a, Y_host, M, N, K): pinned host buffers, row slabs pipelined H2D/kernel/D2H",
               "pageable": {"value": flops_equiv(M, N, nnz) / (pageable_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": pageable_ms,
                            "matches_pinned_result": same_pageable, "note": "same call with unpinned (malloc) X and Y, as main.cpp:38-44 allocates them"},
               "small_call_us": small_us,
               "small_call": f"tcsc_sgemm_prelu_basic, host buffers, M={sm_} K={sk_} N={sn_} 50 % (main.cpp:259), mean of {reps} calls incl. ctypes dispatch"}
        Wh.free()
    else:
        # host X and host Y shared by all ranks (POSIX shared memory, pinned in every process): every rank moves its row
        # block over its own PCIe link (tsg_dist_gemm_host)
        from multiprocessing import shared_memory
        names = [None, None]
        shms = []
        shared = True
        if rank == 0:
            try:  # /dev/shm of a container can be tiny: a sparse segment would only fail (SIGBUS) when touched
                fs = os.statvfs("/dev/shm")
                shared = fs.f_bavail * fs.f_frsize > (M * K + M * N) * 4 + (64 << 20)
            except OSError:
                shared = False
            if shared:
                shms = [shared_memory.SharedMemory(create=True, size=M * K * 4), shared_memory.SharedMemory(create=True, size=M * N * 4)]
                names = [s_.name for s_ in shms]
        dist.broadcast_object_list(names, src=0)
        shared = names[0] is not None
        if shared:
            if rank != 0:
                shms = [shared_memory.SharedMemory(name=names[0]), shared_memory.SharedMemory(name=names[1])]
            Xh = np.ndarray((M, K), np.float32, buffer=shms[0].buf)
            Yh = np.ndarray((M, N), np.float32, buffer=shms[1].buf)
            if rank == 0:
                Xh[:] = Xs[0].cpu().numpy()
                Yh[:] = 0
        else:  # no room for a shared segment: rank-private host buffers (same bytes over the same links; Y_host is then per rank)
            Xh = Xs[0].cpu().numpy().copy()
            Yh = np.zeros((M, N), np.float32)
        barrier()
        t.host_register(Xh)
        t.host_register(Yh)
        ysym = Ys[0] if args.dist_mode >= 1 else D.alloc_y(M, N)
        e2e_steps = max(3, min(args.steps, 10))
        W = state["W"]
        for _ in range(2):
            D.gemm_host(W, Xh, B, Yh, N, a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST, mode=args.dist_mode)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            D.gemm_host(W, Xh, B, Yh, N, a=ALPHA, use_prelu=True, order=t.ORDER_BIAS_LAST, mode=args.dist_mode)
        barrier()  # Y_host is whole once every rank has returned
        t1 = time.perf_counter()
        e2e_ms_t = torch.tensor([(t1 - t0) * 1e3 / e2e_steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(e2e_ms_t, op=dist.ReduceOp.MAX)
        e2e_ms = float(e2e_ms_t.item())
        # the shared host Y must equal the device-resident result of the same X (checked by rank 0 on a row sample of every block)
        rows_per = (M + world - 1) // world
        sample_rows = sorted({min(M - 1, r * rows_per + o) for r in range(world) for o in (0, rows_per // 2, rows_per - 1)})
        idx = torch.tensor(sample_rows, device="cuda")
        same = bool(np.array_equal(ysym[idx].cpu().numpy(), Yh[sample_rows]))  # ysym: the symmetric device Y the call just filled
        same_t = torch.tensor([0 if same else 1], device="cuda", dtype=torch.int32)
        dist.all_reduce(same_t)
        e2e = {"value": flops_equiv(M, N, nnz) / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
               "h2d_bytes_per_step": 4 * M * K, "d2h_bytes_per_step": 4 * M * N,
               "host_y_matches_device_y_on_sampled_rows": int(same_t.item()) == 0,
               "host_buffers": "one POSIX shared-memory segment mapped and pinned by every rank" if shared else "rank-private pinned buffers (/dev/shm too small for a shared segment)",
               "api": "tsg_dist_gemm_host: X and Y in POSIX shared memory pinned by every rank; each rank copies its row block of X host->device over "
                      "its own PCIe link, ncclAllGather of the blocks, column-partitioned GEMM + Y exchange, each rank copies its row block of the full Y device->host"}
        barrier()
        t.host_unregister(Xh)
        t.host_unregister(Yh)
        del Xh, Yh
        for s_ in shms:
            s_.close()
        if rank == 0:
            for s_ in shms:
                s_.unlink()

    # ---- secondary measurements (N = 1): the other configs, sweep corners, conversion ----
    secondary = None
    if world == 1 and not args.no_secondary:
        for buf in (Xs, Ys):
            buf.clear()
        torch.cuda.empty_cache()
        from tools import secondary as sec
        try:
            secondary = sec.run(t, torch, hbm_peak, sm_max_mhz, quick=args.quick_secondary)
        except Exception as e:  # never lose the headline line to a secondary failure; the failure itself is reported
            secondary = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from oracle.pyoracle import Port
        backend, kind, flags = cpu_reference_backend()
        port = Port()
        # the reference's multi-threaded form (sparseGEMM_PReLU<float>, OpenMP over m) on all host threads: a separate
        # process (its own OpenMP runtime and CPU affinity), i.e. exactly what `--impl reference` prints
        all_threads = None
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
                                  "--warmup", "3"], capture_output=True, text=True, timeout=600)
            ref_line = json.loads(out.stdout.strip().splitlines()[-1])
            all_threads = {"value": ref_line["value"], "unit": UNIT, "cores": ref_line["cpu_baseline"]["cores"],
                           "function": ref_line["details"]["function"], "sample": ref_line["cpu_baseline"]["sample"]}
        except Exception as e:  # reported, never fatal: the single-thread figure below is the north-star baseline
            all_threads = {"unavailable": str(e)[:200]}
        pin_one_core()
        Wc = backend.tcsc_from_dense(Wd_host)
        nnz_c = Wc.nnz
        rows = cpu_pick_rows(backend, Wc, M, K, N, 12.0, port)
        secs = cpu_time_rows(backend, Wc, rows, K, N, port)
        cpu_baseline = {"value": flops_equiv(rows, N, nnz_c) / secs / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                        "sample": f"rows 0..{rows - 1} of {M} (all {N} columns), 1 call, {secs:.2f} s, 1 thread pinned (benchmark.sh:36)",
                        "build": flags, "host_cores_available": os.cpu_count(), "seconds_for_full_M_extrapolated": secs * M / rows,
                        "all_threads": all_threads}

    if rank == 0:
        modes = ["ncclAllGather + re-layout", "fused NVLink peer stores in the GEMM epilogue",
                 "copy-engine peer pushes gated by in-kernel progress counters, overlapped with the GEMM",
                 "fused: output tiles staged in shared memory, TMA bulk stores to the local Y and every peer",
                 "fused as mode 3 with a separate output tile (stores of one unit overlap the gathers of the next)",
                 "fused: output tiles staged in shared memory and written once with multimem.st to the NVSwitch multicast mapping of Y"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world, nnz),
            "details": {"order": "tcsc_sgemm_prelu_basic (0, +pos asc, -neg asc, +b, PReLU) -- bit-identical to the reference",
                        "l2": f"{nsets} X buffers{'' if world > 1 and args.dist_mode >= 1 else f' and {nsets} Y buffers'} rotated ({(nsets * M * K + (1 if world > 1 and args.dist_mode >= 1 else nsets) * M * N) * 4 >> 20} MiB > 126 MB L2)",
                        "y_exchange": modes[args.dist_mode] if world > 1 else None,
                        "gather_stream": info, "convert_dense_to_tcsc_ms": convert_ms, "build_gather_stream_ms": stream_ms},
            "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "gadd_per_s": float(M) * nnz / (ms_per_step * 1e-3) / 1e9,
            "dense_equiv_gflops": 2.0 * M * K * N / (ms_per_step * 1e-3) / 1e9,
        }
        line.update({k: verification[k] for k in ("verified", "verified_rows")})
        line["verification"] = verification
        if secondary is not None:
            line["secondary"] = secondary
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        if D is not None:
            torch.cuda.synchronize()
            D.destroy()
        dist.destroy_process_group()
    return 0 if verification["verified"] else 3


def main():
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # the multi-GPU path blocks copy streams on stream-wait ops
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="cfg2")
    ap.add_argument("--dist-mode", type=int, default=5, choices=[0, 1, 2, 3, 4, 5],
                    help="Y exchange of the multi-GPU path (include/tsgemm_b200.h); 5 = NVSwitch multicast, falls back to 3 where the box has none")
    ap.add_argument("--ref-threads", type=int, default=0, help="--impl reference: 0 = all host threads (OpenMP build of the reference), 1 = one pinned thread")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary measurements (other configs, sweep corners, conversion)")
    ap.add_argument("--quick-secondary", action="store_true", help="secondary measurements without the largest shapes")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
