"""The column-partitioned multi-GPU path.

CPU part (gloo, world_size 2, runs everywhere): the partition arithmetic, the rule that re-assembles the global
bit-exact TCSC from per-rank slices, and slab re-assembly of Y through a real all_gather -- with the
oracle standing in for the per-rank kernel, so only the host-side logic is under test.
GPU part (needs >= 2 GPUs): the real thing, both exchange modes, bit-exact against the oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import __graft_entry__ as ge

ROOT = ge.ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port_no, M, K, N, q):
    sys.path.insert(0, ROOT)
    from oracle.pyoracle import Port
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = ge.load()
    port = Port()
    c0, nc = t.partition(N, rank, world)
    Wd = port.gen_ternary(K, N, 42, 1, 4)
    X = port.gen_uniform((M, K), 43) if rank == 0 else np.zeros((M, K), np.float32)
    B = port.gen_uniform((N,), 44)
    xt = torch.from_numpy(X)
    dist.broadcast(xt, src=0)                                  # X broadcast
    wl = port.tcsc_from_dense(np.ascontiguousarray(Wd[:, c0:c0 + nc]))  # rank-local conversion of its column slice
    yl = port.tcsc_sgemm_prelu_basic(xt.numpy(), wl, B[c0:c0 + nc], 0.2)
    # Y all-gather: slabs padded to the widest rank, then re-laid out with the library's partition rule
    widths = [t.partition(N, r, world)[1] for r in range(world)]
    wmax = max(widths)
    slab = torch.zeros((M, wmax))
    slab[:, :nc] = torch.from_numpy(yl)
    out = [torch.zeros((M, wmax)) for _ in range(world)]
    dist.all_gather(out, slab)
    Y = np.concatenate([out[r][:, :widths[r]].numpy() for r in range(world)], axis=1)
    # global TCSC = concatenated row-index arrays + column pointers rebased by the exclusive prefix of (n_pos, n_neg)
    counts = torch.tensor([wl.n_elem_pos, wl.n_elem_neg])
    allc = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allc, counts)
    pre_p = sum(int(allc[r][0]) for r in range(rank))
    pre_n = sum(int(allc[r][1]) for r in range(rank))
    q.put((rank, Y, wl.col_start_pos[:-1] + pre_p, wl.col_start_neg[:-1] + pre_n, wl.row_index_pos, wl.row_index_neg))
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [96, 200])
def test_gloo_two_ranks_partition_and_reassembly(N):
    from oracle.pyoracle import Port
    world, M, K = 2, 5, 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port_no, M, K, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    port = Port()
    Wd = port.gen_ternary(K, N, 42, 1, 4)
    wg = port.tcsc_from_dense(Wd)
    Yref = port.tcsc_sgemm_prelu_basic(port.gen_uniform((M, K), 43), wg, port.gen_uniform((N,), 44), 0.2)
    for _, Y, *_ in res:
        assert np.array_equal(Y, Yref)                       # every rank holds the full, bit-identical Y
    csp = np.concatenate([r[2] for r in res] + [[wg.n_elem_pos]])
    csn = np.concatenate([r[3] for r in res] + [[wg.n_elem_neg]])
    assert np.array_equal(csp, wg.col_start_pos) and np.array_equal(csn, wg.col_start_neg)
    assert np.array_equal(np.concatenate([r[4] for r in res]), wg.row_index_pos)
    assert np.array_equal(np.concatenate([r[5] for r in res]), wg.row_index_neg)


# ---- real multi-GPU run ------------------------------------------------------------------------------------------------
MODES = {0: "nccl_allgather", 1: "fused_peer_stores", 2: "copy_engine_overlap", 3: "fused_tma_stores", 4: "fused_tma_separate_tile", 5: "multicast"}
SHAPES = [(256, 512, 1000), (130, 96, 40)]  # the second: fewer than 32 columns per rank, ragged (tsg_dist_partition's corner)


def _init_nccl(rank, world, port_no):
    import datetime
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank), timeout=datetime.timedelta(seconds=120))
    t = ge.load()
    t.lib()
    t.use_torch_stream()
    return t


def _gpu_worker(rank, world, port_no, q):
    """ONE process group for every (mode, shape): spawning and NCCL bootstrap dominate the cost of a 2-GPU test"""
    sys.path.insert(0, ROOT)
    t = _init_nccl(rank, world, port_no)
    D = t.Dist(rank, world)
    out = {}
    for (M, K, N) in SHAPES:
        c0, nc = D.partition(N)
        W = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, 42, 1, 4))
        B = t.gen_uniform((N,), 44)
        for mode in MODES:
            X = t.gen_uniform((M, K), 43) if rank == 0 else torch.zeros((M, K), device="cuda")
            Y = D.alloc_y(M, N) if mode >= 1 else torch.empty((M, N), device="cuda")
            if mode == 5 and not D.has_multicast():
                out[(mode, M, K, N)] = None
                continue
            for _ in range(2):  # twice: the second run overwrites a Y that peers have already read
                D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=mode)
            torch.cuda.synchronize()
            out[(mode, M, K, N)] = Y.cpu().numpy().copy()
            dist.barrier()
        if D.has_multicast():
            # root's X only float-aligned (and only root's): the choice of broadcast path must not depend on a rank's own pointer
            flat = torch.zeros(M * K + 1, device="cuda")
            X = flat[1:].view(M, K) if rank == 0 else flat[:M * K].view(M, K)
            if rank == 0:
                X.copy_(t.gen_uniform((M, K), 43))
            Y = D.alloc_y(M, N)
            D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=5)
            torch.cuda.synchronize()
            out[("unaligned_root_x", M, K, N)] = Y.cpu().numpy().copy()
            out[("unaligned_root_x_bcast", M, K, N)] = X.cpu().numpy().copy()
            dist.barrier()
            # X above the size up to which the switch broadcasts it: mode 5 with ncclBroadcast, then back
            os.environ["TSG_MC_BCAST_MAX_MB"] = "0"
            X = t.gen_uniform((M, K), 43) if rank == 0 else torch.zeros((M, K), device="cuda")
            for _ in range(2):
                D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=5)
            del os.environ["TSG_MC_BCAST_MAX_MB"]
            D.gemm(W, X, B, Y, N, a=0.2, use_prelu=True, root=0, mode=5)
            torch.cuda.synchronize()
            out[("nccl_bcast_mode5", M, K, N)] = Y.cpu().numpy().copy()
            dist.barrier()
        W.destroy()
    q.put((rank, out))
    dist.barrier()
    D.destroy()
    dist.destroy_process_group()


def _run_workers(target, world, extra, timeout):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port_no) + tuple(extra) + (q,), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted([q.get(timeout=timeout) for _ in range(world)], key=lambda x: x[0])
        for p in procs:
            p.join(timeout=60)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        return res
    finally:
        for p in procs:  # never leave a rank behind (it would keep the GPUs and the caller's pipes)
            if p.is_alive():
                p.kill()


@pytest.mark.gpu
def test_two_gpus_bit_exact_every_mode():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle.pyoracle import Port
    res = _run_workers(_gpu_worker, 2, (), 600)
    port = Port()
    skipped = []
    for (M, K, N) in SHAPES:
        wg = port.tcsc_from_dense(port.gen_ternary(K, N, 42, 1, 4))
        Yref = port.tcsc_sgemm_prelu_basic(port.gen_uniform((M, K), 43), wg, port.gen_uniform((N,), 44), 0.2)
        for mode, name in MODES.items():
            for rank, out in res:
                Y = out[(mode, M, K, N)]
                if Y is None:
                    skipped.append(name)
                    continue
                assert np.array_equal(Y, Yref), f"mode {mode} ({name}), shape {(M, K, N)}, rank {rank}"
        for rank, out in res:
            if ("unaligned_root_x", M, K, N) in out:
                assert np.array_equal(out[("unaligned_root_x", M, K, N)], Yref), f"unaligned root X, shape {(M, K, N)}, rank {rank}"
                assert np.array_equal(out[("unaligned_root_x_bcast", M, K, N)], port.gen_uniform((M, K), 43)), "X was not broadcast in place"
                assert np.array_equal(out[("nccl_bcast_mode5", M, K, N)], Yref), f"mode 5 with ncclBroadcast, shape {(M, K, N)}, rank {rank}"
    print("modes without hardware support on this box:", sorted(set(skipped)))


def _stress_worker(rank, world, port_no, iters, q):
    """random shapes, the fused exchange modes, every Y checked bitwise against a single-GPU recompute on this rank"""
    sys.path.insert(0, ROOT)
    t = _init_nccl(rank, world, port_no)
    D = t.Dist(rank, world)
    Mmax, Nmax = 1024, 2048
    Y = D.alloc_y(Mmax, Nmax)
    bad = {}
    for mode in (3, 5):
        if mode == 5 and not D.has_multicast():
            bad[mode] = -1
            continue
        rng = np.random.default_rng(1234 + mode)  # the same sequence on every rank
        nbad = 0
        for it in range(iters):
            M = int(rng.integers(32, Mmax + 1))
            K = int(rng.integers(1, 700))
            N = int(rng.integers(1, Nmax // 4 + 1)) * 4
            den = int(rng.choice([2, 4, 10, 50]))
            c0, nc = D.partition(N)
            W = t.DeviceTcsc.from_dense(t.gen_ternary_slice(K, N, c0, nc, 1000 + it, 1, den))
            X = t.gen_uniform((M, K), 2000 + it) if rank == 0 else torch.zeros((M, K), device="cuda")
            B = t.gen_uniform((N,), 3000 + it)
            Yv = Y.view(-1)[:M * N].view(M, N)
            D.gemm(W, X, B, Yv, N, a=0.2, use_prelu=True, root=0, mode=mode)
            torch.cuda.synchronize()
            Wf = t.DeviceTcsc.from_dense(t.gen_ternary(K, N, 1000 + it, 1, den))  # the whole W: one recompute covers every slab
            Yf = torch.empty((M, N), device="cuda")
            Wf.gemm(X, B, Yf, a=0.2, use_prelu=True)
            torch.cuda.synchronize()
            nbad += int(not torch.equal(Yf, Yv))
            Wf.destroy()
            W.destroy()
        bad[mode] = nbad
    q.put((rank, bad))
    dist.barrier()
    D.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpus_stress_random_shapes():
    """the cross-rank ordering of the fused modes (bulk / multicast stores, kernel retirement, 4-byte all-reduce) under many
    back-to-back calls with different shapes into the same symmetric buffer (TSG_STRESS_ITERS, default 150 per mode)"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    iters = int(os.environ.get("TSG_STRESS_ITERS", "150"))
    res = _run_workers(_stress_worker, 2, (iters,), 900)
    for rank, bad in res:
        for mode, n in bad.items():
            assert n <= 0, f"rank {rank}, mode {mode}: {n} of {iters} results differ from the single-GPU recompute"
    print("stress:", res)
