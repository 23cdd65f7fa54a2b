#!/usr/bin/env python
"""Sweep of the peer-pull exchange on N GPUs (torchrun): copy-engine vs SM-driven pulls, column-block grouping, lanes.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/peer_variants.py [--workload hubbard4x4]

Every variant is checked against the all-gather product and timed like bench.py (CUDA events on the compute stream,
barrier on both sides, max over ranks).  Rank 0 prints one JSON line per variant and a summary line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="hubbard4x4")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--variants", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import quantum_basis_b200 as qb
    from quantum_basis_b200 import dist as qd
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L = qb.lib()
    assert L.qbgpu_init(local_rank) == 0
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert L.qbgpu_set_stream(C.c_void_p(stream.cuda_stream)) == 0
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fam, p = bench.WORKLOADS[a.workload]
    n = L.qbgpu_dim_hubbard(p["Lx"] * p["Ly"], p["nup"], p["ndn"]) if fam == "hubbard" else L.qbgpu_dim_heisenberg(p["L"], p["L"] // 2)
    bounds, chunk = qd.equal_row_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    M = bench.build_matrix(qb, a.workload, row_range=(lo, hi))
    comm = qd.TorchComm()
    kern0 = qd.DeviceKernels(qb, M, real=False)
    opA = qd.ShardedOperator(kern0, n, rank, world, comm)
    full = qb.vec_randomize(n, 1)
    x_loc = kern0.alloc(chunk)
    x_loc[:2 * (hi - lo)] = torch.from_numpy(np.ascontiguousarray(full[lo:hi]).view(np.float64)).cuda()
    y_ref = kern0.alloc(chunk)
    msA = qd._timed(torch, dist, stream, lambda: opA.matvec(x_loc, y_ref), a.steps, a.warmup)
    opA.gather(x_loc)
    ms_local = qd._timed(torch, dist, stream, lambda: kern0.multmv(opA.x_full, y_ref), a.steps, a.warmup)
    opA.matvec(x_loc, y_ref)
    if rank == 0:
        print(json.dumps({"variant": "allgather", "ms": msA, "local_product_ms": ms_local, "n_gpus": world}), flush=True)

    # variant spec: g<group size>-<ce|sm<ctas>>-l<lanes>
    full_l = min(world - 1, 8)
    spec = a.variants.split(",") if a.variants else (
        [f"g1-ce-l{full_l}", f"g1-sm16-l{full_l}"] if world <= 2 else
        [f"g1-ce-l{full_l}", f"g2-ce-l{full_l}", f"g4-ce-l{full_l}", "g2-ce-l2", "g4-ce-l1", "g1-ce-l1", f"g1-sm16-l{full_l}", f"g2-sm16-l{full_l}"])
    variants = []
    for v in spec:
        if v.count("-") != 2:               # e.g. the "noring" switch
            continue
        gs, ms_, ls = v.split("-")
        sched = "matching" if ms_.endswith("m") else None      # e.g. g1-cem-l1: needed transfers only, in matching rounds
        ms_ = ms_.rstrip("m")
        variants.append((int(gs[1:]), ms_[:2], int(ms_[2:] or 0), int(ls[1:]), sched))
    results = []
    parts_cache = {}
    op = None
    for (g, mode, ctas, lanes, sched) in variants:
        groups = qd.peer_groups(world, rank, g)
        if g not in parts_cache:
            cb = [min(n, gr[0] * chunk) for gr in groups] + [n]
            t0 = time.time()
            parts_cache[g] = qd.DeviceKernels.split(qb, M, cb)
            torch.cuda.synchronize()
            split_s = time.time() - t0
        kern = qd.DeviceKernels(qb, M, real=False, parts=parts_cache[g])
        if op is None:                       # one set of exported ping-pong buffers serves every variant
            op = qd.PeerExchangeOperator(qb, kern, n, rank, world, comm, torch, lanes=lanes, mode=mode, ctas=max(1, ctas), groups=groups)
            op.own(0).upload(np.ascontiguousarray(full[lo:hi]))
        torch.cuda.synchronize(); dist.barrier()
        op.configure(kern, groups, mode=mode, ctas=max(1, ctas), lanes=lanes, schedule=sched or "ring")
        comm.all_reduce(op.token)
        y = kern.alloc(chunk)
        ms = qd._timed(torch, dist, stream, lambda: op.matvec(0, y), a.steps, a.warmup)
        err = float((y - y_ref).abs().max().item()) / max(1e-300, float(y_ref.abs().max().item()))
        ms_pull = qd._timed(torch, dist, stream, lambda: op.pull_only(0), a.steps, a.warmup)
        # products of all blocks with nothing to wait for (vector already in place)
        def blocks_only():
            kern.multmv_part(op.own_group, op.X[0], y, accumulate=False)
            for gi in op.group_order:
                kern.multmv_part(gi, op.X[0], y, accumulate=True)
        ms_blocks = qd._timed(torch, dist, stream, blocks_only, a.steps, a.warmup)
        errt = torch.tensor([err], dtype=torch.float64, device="cuda")
        dist.all_reduce(errt, op=dist.ReduceOp.MAX)
        rec = {"variant": f"g{g}-{mode}{ctas or ''}{'m' if sched else ''}-l{lanes}", "schedule": op.schedule, "group": g, "mode": mode, "ctas": ctas, "lanes": lanes, "ms": ms, "pull_only_ms": ms_pull,
               "blocks_only_ms": ms_blocks, "blocks": len(groups), "skipped_empty_blocks": op.skipped_blocks, "max_rel_err_vs_allgather": float(errt.item()),
               "pull_GBs_per_gpu": (n - (hi - lo)) * 16 / (ms_pull * 1e-3) / 1e9}
        results.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
        del y
    # ring-fused: one kernel over the unsplit shard (rows rotated into ring order) consuming the slices as they land
    if op is not None and "noring" not in a.variants:
        torch.cuda.synchronize(); dist.barrier()
        ringM = qd.PeerExchangeOperator.ring_view(qb, M, rank, world, chunk)
        kr = qd.DeviceKernels(qb, ringM, real=False)
        op.enable_ring(kr)
        op.configure(kern, groups, lanes=1, mode="ce", schedule="ring")
        y = kr.alloc(chunk)
        op.matvec_ring(0, y)
        torch.cuda.synchronize()
        bad = torch.tensor([1.0 if op.ring_timed_out() else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(bad)
        if float(bad.item()) == 0.0:
            for lanes in ((1, 2) if world > 2 else (1,)):
                op.configure(kern, groups, lanes=lanes, mode="ce", schedule="ring")
                ms = qd._timed(torch, dist, stream, lambda: op.matvec_ring(0, y), a.steps, a.warmup)
                err = float((y - y_ref).abs().max().item()) / max(1e-300, float(y_ref.abs().max().item()))
                errt = torch.tensor([err], dtype=torch.float64, device="cuda")
                dist.all_reduce(errt, op=dist.ReduceOp.MAX)
                rec = {"variant": f"ring-fused-l{lanes}", "group": 0, "mode": "ce", "lanes": lanes, "ms": ms, "blocks": 1,
                       "max_rel_err_vs_allgather": float(errt.item()), "timed_out": op.ring_timed_out()}
                results.append(rec)
                if rank == 0:
                    print(json.dumps(rec), flush=True)
            # the rotated shard behind a plain all-gather (ordinary kernel on the same arrays)
            ms_rot = qd._timed(torch, dist, stream, lambda: kern0.multmv(opA.x_full, y), a.steps, a.warmup)
            if rank == 0:
                print(json.dumps({"variant": "local_product_rotated_rows", "ms": ms_rot}), flush=True)
        elif rank == 0:
            print(json.dumps({"variant": "ring-fused", "error": "a waiter timed out: arrival flags were not set"}), flush=True)
    if rank == 0 and results:
        best = min(results, key=lambda r: r["ms"])
        print(json.dumps({"summary": True, "n_gpus": world, "allgather_ms": msA, "local_product_ms": ms_local, "best": best}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
