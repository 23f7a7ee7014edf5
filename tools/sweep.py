#!/usr/bin/env python
"""BASELINE configs[4]: single / pair / sponge(8) hashing sweep from 2^16 to 2^28 inputs on this rank's GPU(s), next to
the reference CPU path on the host cores.  Device-resident timing with CUDA events; one JSON document on stdout.

  python tools/sweep.py [--max-log2 28] [--cpu-sample 200000]
  torchrun --nproc-per-node N tools/sweep.py      (each rank hashes its contiguous slice; no collective)
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cuzk_b200 import api, lib as cl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--min-log2", type=int, default=16)
ap.add_argument("--max-log2", type=int, default=28)
ap.add_argument("--cpu-sample", type=int, default=200_000)
args = ap.parse_args()

world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
api.initialize(local)
L = cl.get_lib()
st = torch.cuda.current_stream(dev)
sp = st.cuda_stream


def timed(fn, reps):
    fn()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        fn()
    e1.record(st)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


rows = []
for lg in range(args.min_log2, args.max_log2 + 1):
    total = 1 << lg                      # inputs over the whole job
    lo, hi = total * rank // world, total * (rank + 1) // world
    m = hi - lo                          # this rank's inputs
    x = torch.empty((m, 4), dtype=torch.int64, device=dev)
    L.check(L.cuzk_synth_elements(x.data_ptr(), m, 5, lo, 1, sp), "synth")
    reps = max(1, min(20, (1 << 24) // m))
    row = {"log2_inputs": lg, "inputs": total}
    out = torch.empty((m, 4), dtype=torch.int64, device=dev)
    ms = timed(lambda: L.check(L.cuzk_poseidon_hash_single(x.data_ptr(), out.data_ptr(), m, 0, sp), "single"), reps)
    row["single"] = {"hashes": total, "ms": ms, "hashes_per_s": total / (ms * 1e-3)}
    half = m // 2                        # pair: inputs consumed two per hash
    ms = timed(lambda: L.check(L.cuzk_poseidon_hash_pairs(x.data_ptr(), x.data_ptr() + half * 32, out.data_ptr(), half, 0, sp), "pairs"), reps)
    row["pair"] = {"hashes": total // 2, "ms": ms, "hashes_per_s": (total // 2) / (ms * 1e-3), "inputs_per_s": total / (ms * 1e-3)}
    eighth = m // 8                      # sponge: hash_multiple of 8 inputs = 4 permutations
    ms = timed(lambda: L.check(L.cuzk_poseidon_sponge(x.data_ptr(), 8, 3, out.data_ptr(), eighth, 0, sp), "sponge"), reps)
    row["sponge8"] = {"hashes": total // 8, "ms": ms, "hashes_per_s": (total // 8) / (ms * 1e-3), "permutations_per_s": (total // 2) / (ms * 1e-3)}
    rows.append(row)
    del x, out

doc = {"n_gpus": world, "sweep": rows, "data": "synthetic splitmix64 stream (seed 5), canonical elements, device resident"}
if rank == 0:
    from oracle_lib import Oracle, Ref, have_ref, synth_elements

    impl = Ref() if have_ref() else Oracle()
    cores = len(os.sched_getaffinity(0))
    n = args.cpu_sample
    l, r = synth_elements(5, n), synth_elements(6, n)
    impl.hash_pairs_mt(l[:64], r[:64], cores)
    t0 = time.perf_counter()
    impl.hash_pairs_mt(l, r, cores)
    dt = time.perf_counter() - t0
    doc["cpu_baseline"] = {"kind": "reference" if have_ref() else "port", "cores": cores, "pair_hashes_per_s": n / dt,
                           "sample": f"{n} pair hashes on {cores} host threads"}
    print(json.dumps(doc))
if world > 1:
    dist.destroy_process_group()
