#!/usr/bin/env python
"""bench.py -- headline benchmark of the cuZK hot path on B200 (contract: see the task brief).

A *step* is one pass of the hot path over one batch of synthetic input: 1,000,000 Poseidon pair hashes
(BASELINE.json configs[0] workload, "Large Scale" of run_poseidon_benchmark.sh) per GPU, one launch through
the C ABI.  `value` = whole-job pair-hashes/s with inputs resident in HBM; `e2e` = the same through the
host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region).  The same JSON line carries
the Merkle-build leaves/s figures (configs[1..3]) under "merkle", the integer-multiply roofline under
"roofline", and the reference CPU path timed on this box's host cores under "cpu_baseline".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, NCCL)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAD_PER_PERM = 48_576          # 240 x 164 + 576 x 16 multiply-adds per permutation
IMAD_WIDE_PIPE_MODEL_PER_SM_CLK = 32.0   # 4 heavy-pipe cycles per warp IMAD.WIDE (profiles/r01_tuning_notes.md) = 32 lanes / SM / clock
# root of the 2^26-leaf 8-ary tree over cuzk_synth_u64_leaves(seed 4): the same on 1, 2, 4 and 8 GPUs, and equal to the
# single-GPU full build (tests/test_gpu_parity.py::test_config4_octary_2p26_sharded_equals_full_build)
EXPECTED_ROOTS = {(26, 8): "0af40a4830624406744ebf70622802c7811231c8e2ce8c441c20c1c714f9c075"}
# single-GPU time of the same 2^26-leaf 8-ary build (profiles/r02_bench_n1.json), the numerator of the strong-scaling efficiency
# the N > 1 lines report next to their own time
SINGLE_GPU_BUILD_MS = {(26, 8): 213.8}
BYTES_PER_PAIR_HASH = 96        # 2 x 32 B in + 32 B out
N_PAIRS = 1_000_000
REF_PUBLISHED_PAIR_HASHES_PER_S = 2_145_027   # BASELINE.md section 1 (reference README.md:134, A100)
METRIC = "poseidon_pair_hashes_per_s"
UNIT = "hashes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=N_PAIRS)
    ap.add_argument("--no-merkle", action="store_true", help="skip the Merkle sub-benchmarks")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the configs[4] hashing sweep")
    ap.add_argument("--sweep-max-log2", type=int, default=28)
    ap.add_argument("--merkle-log2", type=int, default=26, help="log2 leaves of the sharded 8-ary build")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"{args.pairs} Poseidon pair hashes per GPU per step (t=3, R_F=8, R_P=56, BN254 Fr; BASELINE configs[0] 'Large Scale' job on GPU); "
                    "the metric's second quantity, Merkle build leaves/s for configs[1..3], is under 'merkle'",
        "pairs_per_gpu": args.pairs,
        "sharding": f"independent slices x{world}, no collective",
        "l2": "4 rotating input/output sets of 96 MB each (384 MB > 126 MB L2)",
    }


# -------------------------------------------------------------------------------------------- CPU legs
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(n_hashes, threads):
    """Time the reference CPU implementation (oracle/_ref, else the oracle port) on `threads` host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Ref, have_ref, synth_elements

    kind = "reference" if have_ref() else "port"
    impl = Ref() if kind == "reference" else Oracle()
    l, r = synth_elements(1, n_hashes), synth_elements(2, n_hashes)
    impl.hash_pairs_mt(l[:64], r[:64], min(threads, 64))  # constants + thread warm-up
    t0 = time.perf_counter()
    impl.hash_pairs_mt(l, r, threads)
    dt = time.perf_counter() - t0
    return kind, n_hashes / dt, dt


def cpu_merkle_baseline(gpu_root):
    """the reference's own NaryMerkleTree on one host thread -- build over the same
    50 000 leaves (root compared with the GPU's), then a bounded sample of its generate_proof + verify_proof."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Ref, have_ref, synth_u64_leaves

    if not have_ref():
        return {"unavailable": "oracle/_ref/libcuzk_ref.so not built (needs the reference tree at build time)"}
    ref = Ref()
    leaves = synth_u64_leaves(3, 50_000)
    build_ms, root = ref.tree_build_ms(leaves, 2)
    tree = ref.tree(leaves, 2)
    sample = 400
    idx = [(i * 10) % 50_000 for i in range(sample)]
    t0 = time.perf_counter()
    proofs = [tree.prove(i) for i in idx]
    t1 = time.perf_counter()
    ok = all(tree.verify(leaves[i], sib, pos, root) for i, (sib, pos) in zip(idx, proofs))
    t2 = time.perf_counter()
    return {"kind": "reference", "cores": 1, "build_ms": build_ms, "leaves_per_s": 50_000 / (build_ms * 1e-3),
            "root_equals_gpu": bool((root == gpu_root).all()), "verify_proofs_per_s": sample / (t2 - t1),
            "prove_proofs_per_s": sample / (t1 - t0), "all_valid": bool(ok),
            "sample": f"NaryMerkleTree build of the same 50 000 leaves + {sample} of the 5 000 proofs, one host thread"}


def reference_arm(args):
    """--impl reference: the reference's own CPU path on all host cores, same metric and unit.  Each step is a BOUNDED SAMPLE
    of the workload (the CPU rate does not depend on the batch size); the line says how many hashes a step really ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    per_step = max(cores * 4000, 8000)  # ~0.4 s per step per core-set; whole run stays within minutes
    vals = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        kind, v, dt = cpu_reference_run(per_step, cores)
        if i >= args.warmup:
            vals.append((v, dt))
    total_h = per_step * len(vals)
    total_t = sum(dt for _, dt in vals)
    value = total_h / total_t
    # the reference's OWN benchmark loop, unmodified: one thread, Poseidon::benchmark_poseidon_pairs (poseidon.cpp:195-219,
    # called by src/poseidon/test/benchmark.cpp:162-163) -- BASELINE.md section 2 row 1a
    stock = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import Ref, have_ref

        if have_ref():
            n1 = 20_000
            stock = {"value": Ref().benchmark_poseidon_pairs(n1), "unit": UNIT, "cores": 1, "hashes": n1,
                     "what": "Poseidon::benchmark_poseidon_pairs (poseidon.cpp:195-219), the loop behind run_poseidon_benchmark.sh, one host thread"}
    except Exception as e:  # noqa: BLE001
        stock = {"unavailable": str(e)}
    cfg = workload_config(args, args.gpus)
    cfg["workload"] = (f"{per_step} Poseidon pair hashes per step on {cores} host threads: a bounded sample of the GPU arm's "
                       f"{args.pairs}-pair step (the CPU rate is independent of the batch size)")
    cfg["pairs_per_step"] = per_step
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total_t / len(vals),
        "hashes_per_step": per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u32x8 (256-bit integers)",
        "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{per_step} pair hashes per step on {cores} host threads (PoseidonHash::hash_pair), {len(vals)} steps"},
        "reference_single_thread": stock,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every few milliseconds from a host thread while the timed
    region runs (nvidia-smi -lms cannot sample a 50 ms region)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu, self.period, self.samples, self._stop, self.t, self.h = gpu_index, period_s, [], threading.Event(), None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                return i
        return i

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            except Exception as e:  # noqa: BLE001
                self.err = f"clock: {e}"
                time.sleep(self.period)
                continue
            rs = 0
            for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
                try:
                    rs = int(getattr(nv, fn)(self.h))
                    break
                except Exception as e:  # noqa: BLE001
                    self.err = f"{fn}: {e}"
            self.samples.append((time.perf_counter(), sm, rs))
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=1)

    def summary(self, t0, t1):
        sel = [(sm, rs) for t, sm, rs in self.samples if t0 <= t <= t1]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_sm", None), "reasons": [], "samples": 0,
                    "error": getattr(self, "err", None), "total_samples": len(self.samples)}
        reasons = sorted({name for _, rs in sel for name, bit in self.REASONS if rs & bit})
        return {"sm_mhz": float(np.median([sm for sm, _ in sel])), "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(sel)}


# -------------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return
    # keep stdout clean for the ONE JSON line: libraries (NCCL prints its version banner) write to fd 1 otherwise
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    from cuzk_b200 import api, lib as cl
    from cuzk_b200.distributed import shard_slice, sharded_all_valid

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api.initialize(local)
    L = cl.get_lib()
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.pairs
    nsets = 4
    # ---- inputs resident in HBM: seeded splitmix64 stream, distinct per rank and per set ----
    sets = []
    for s in range(nsets):
        l = torch.empty((n, 4), dtype=torch.int64, device=dev)
        r = torch.empty((n, 4), dtype=torch.int64, device=dev)
        o = torch.empty((n, 4), dtype=torch.int64, device=dev)
        start = (rank * nsets + s) * n
        L.check(L.cuzk_synth_elements(l.data_ptr(), n, 1, start, 1, sp), "synth")
        L.check(L.cuzk_synth_elements(r.data_ptr(), n, 2, start, 1, sp), "synth")
        sets.append((l, r, o))
    torch.cuda.synchronize(dev)

    def step_device(i):
        l, r, o = sets[i % nsets]
        L.check(L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), n, 0, sp), "hash_pairs")

    # ---- measured IMAD peak (roofline denominator), rank 0's GPU ----
    peak = {}
    for name, variant in (("imad_wide_reg", 0), ("imad_wide_imm", 10), ("imad_wide_constbank", 11), ("imad_wide_x_chain_reg", 3),
                          ("imad_wide_x_chain_imm", 12), ("imad_hi", 2), ("imad_lo", 1), ("iadd3_x_chain", 4), ("dfma", 9)):
        v = (cl.C.c_double)()
        L.check(L.cuzk_imad_peak(variant, 2000, cl.C.byref(v)), "imad_peak")
        peak[name] = v.value
    # the roofline denominator: the best rate any form of the 32x32->64 multiply-add reaches on this chip
    imad_peak = max(v for k, v in peak.items() if k.startswith("imad_wide"))
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing ----
    for i in range(args.warmup):
        step_device(i)
    barrier()
    launches0 = L.cuzk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = L.cuzk_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms_total / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host memory) ----
    hl = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hr = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    ho = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    hl.copy_(sets[0][0])
    hr.copy_(sets[0][1])
    torch.cuda.synchronize(dev)

    def step_host():
        L.check(L.cuzk_poseidon_hash_pairs(hl.data_ptr(), hr.data_ptr(), ho.data_ptr(), n, 1, sp), "hash_pairs(host)")

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()   # synchronous: returns after the D2H copy has landed
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * e2e_steps / e2e_s
    e2e_ok = bool((ho.view(torch.int64)[:256] == sets[0][2][:256].cpu()).all()) if args.warmup + args.steps > 0 else True

    # the same call on PAGEABLE host memory (numpy arrays = what a std::vector hands over; the library stages through its own
    # pinned bounce buffers)
    pl, pr = hl.numpy().copy(), hr.numpy().copy()
    po = np.empty_like(pl)

    def step_pageable():
        L.check(L.cuzk_poseidon_hash_pairs(pl.ctypes.data, pr.ctypes.data, po.ctypes.data, n, 1, sp), "hash_pairs(pageable)")

    step_pageable()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, e2e_steps // 2)):
        step_pageable()
    pageable_value = world * n * max(2, e2e_steps // 2) / max_over_ranks(time.perf_counter() - t0)
    pageable_ok = bool((po[:256] == ho.numpy()[:256]).all())

    # the reference harness shape: 245 synchronous host calls of <= 4096 pairs (poseidon_cuda_benchmarks.cpp:63-117), on pinned
    # and on pageable memory (the reference harness holds std::vectors)
    def harness(ptr_l, ptr_r, ptr_o):
        b = 4096
        t0 = time.perf_counter()
        done = 0
        while done < n:
            m = min(b, n - done)
            L.check(L.cuzk_poseidon_hash_pairs(ptr_l + done * 32, ptr_r + done * 32, ptr_o + done * 32, m, 1, sp), "hash_pairs(host)")
            done += m
        torch.cuda.synchronize(dev)
        return world * n / max_over_ranks(time.perf_counter() - t0)

    b = 4096
    harness(hl.data_ptr(), hr.data_ptr(), ho.data_ptr())
    b4096_value = harness(hl.data_ptr(), hr.data_ptr(), ho.data_ptr())
    b4096_pageable = harness(pl.ctypes.data, pr.ctypes.data, po.ctypes.data)
    harness_ok = bool((po[-256:] == sets[0][2][-256:].cpu().numpy()).all())

    # ---- latency of one launch narrower than the chip (device resident; what bounds upper tree levels and small batches) ----
    def launch_us(units, reps=20):
        l, r, o = sets[0]
        L.check(L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), units, 0, sp), "pairs")
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            L.check(L.cuzk_poseidon_hash_pairs(l.data_ptr(), r.data_ptr(), o.data_ptr(), units, 0, sp), "pairs")
        a1.record(stream)
        torch.cuda.synchronize(dev)
        return round(a0.elapsed_time(a1) / reps * 1e3, 1)

    latency = {"unit": "us per launch of N pair hashes, back to back on one stream",
               "cooperative": {str(u): launch_us(u) for u in (1, 1184, 2368, 4096)}}
    prev_coop = L.cuzk_debug_set_coop_max(0)
    latency["one_thread_per_hash"] = {str(u): launch_us(u) for u in (1, 4096)}
    L.cuzk_debug_set_coop_max(prev_coop)

    # ---- configs[4]: single / pair / sponge(8) hashing sweep over 2^16 .. 2^28 inputs (whole job; each rank its slice) ----
    sweep = None
    if not args.no_sweep:
        sweep = []
        for lg in (16, 20, 24, 28):
            if lg > args.sweep_max_log2:
                continue
            total = 1 << lg
            lo, hi = shard_slice(total, rank, world)
            m = hi - lo
            x = torch.empty((m, 4), dtype=torch.int64, device=dev)
            out = torch.empty((m, 4), dtype=torch.int64, device=dev)
            L.check(L.cuzk_synth_elements(x.data_ptr(), m, 5, lo, 1, sp), "synth")
            reps = max(1, min(10, (1 << 22) // max(m, 1)))

            def timed(fn):
                fn()
                barrier()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(reps):
                    fn()
                a1.record(stream)
                barrier()
                return max_over_ranks(a0.elapsed_time(a1) / reps)

            row = {"log2_inputs": lg}
            ms = timed(lambda: L.check(L.cuzk_poseidon_hash_single(x.data_ptr(), out.data_ptr(), m, 0, sp), "single"))
            row["single_hashes_per_s"] = total / (ms * 1e-3)
            ms = timed(lambda: L.check(L.cuzk_poseidon_hash_pairs(x.data_ptr(), x.data_ptr() + (m // 2) * 32, out.data_ptr(), m // 2, 0, sp), "pairs"))
            row["pair_hashes_per_s"] = (total // 2) / (ms * 1e-3)
            ms = timed(lambda: L.check(L.cuzk_poseidon_sponge(x.data_ptr(), 8, 3, out.data_ptr(), m // 8, 0, sp), "sponge"))
            row["sponge8_hashes_per_s"] = (total // 8) / (ms * 1e-3)
            row["sponge8_permutations_per_s"] = (total // 2) / (ms * 1e-3)
            sweep.append(row)
            del x, out

    # ---- Merkle sub-benchmarks ----
    merkle = {}
    if not args.no_merkle:

        def time_build(nleaves, arity, reps):
            leaves = torch.empty((nleaves, 4), dtype=torch.int64, device=dev)
            L.check(L.cuzk_synth_u64_leaves(leaves.data_ptr(), nleaves, 3, 0, sp), "synth")
            tot = api.total_nodes(nleaves, arity)
            levels = torch.empty((tot, 4), dtype=torch.int64, device=dev)
            L.check(L.cuzk_merkle_build(leaves.data_ptr(), nleaves, arity, levels.data_ptr(), 0, sp), "build")
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(reps):
                L.check(L.cuzk_merkle_build(leaves.data_ptr(), nleaves, arity, levels.data_ptr(), 0, sp), "build")
            a1.record(stream)
            torch.cuda.synchronize(dev)
            return a0.elapsed_time(a1) / reps, leaves, levels

        def time_build_e2e(leaves, arity, reps):
            """end to end through the device-resident tree handle: leaves start in pinned HOST memory, the root comes back to the
            host (cuzk_tree_build + cuzk_tree_root + cuzk_tree_free per step), wall clock"""
            host = leaves.cpu().pin_memory()
            root = np.empty(4, dtype=np.uint64)
            n_ = host.shape[0]

            def once():
                h = cl.C.c_void_p()
                L.check(L.cuzk_tree_build(host.data_ptr(), n_, arity, 1, sp, cl.C.byref(h)), "tree_build")
                L.check(L.cuzk_tree_root(h, root.ctypes.data, 1, sp), "tree_root")
                L.check(L.cuzk_tree_free(h), "tree_free")

            once()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(reps):
                once()
            dt = (time.perf_counter() - t0) / reps
            return {"e2e_build_ms": dt * 1e3, "e2e_leaves_per_s": n_ / dt, "h2d_bytes": 32 * n_, "d2h_bytes": 32,
                    "api": "cuzk_tree_build(mem=CUZK_MEM_HOST, pinned) + cuzk_tree_root"}

        if rank == 0:
            ms, leaves, levels = time_build(50_000, 2, 5)
            t = api.CudaNaryMerkleTree(arity=2)
            t.leaf_count, t.levels = 50_000, levels
            idx = (torch.arange(5000, device=dev, dtype=torch.int64) * 10) % 50_000
            pb = t.generate_batch_proofs(idx)
            lv = leaves[idx].contiguous()
            res = t.verify_batch_proofs(pb, lv)
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(5):
                res = t.verify_batch_proofs(pb, lv)
            a1.record(stream)
            torch.cuda.synchronize(dev)
            vms = a0.elapsed_time(a1) / 5
            merkle["binary_50k"] = {"leaves": 50_000, "arity": 2, "build_ms": ms, "leaves_per_s": 50_000 / (ms * 1e-3),
                                    "verify_5k_ms": vms, "proofs_per_s": 5000 / (vms * 1e-3), "all_valid": bool(res.all())}
            merkle["binary_50k"].update(time_build_e2e(leaves, 2, 5))
            if world == 1 and not args.no_cpu:
                merkle["binary_50k"]["cpu_reference"] = cpu_merkle_baseline(levels[-1].cpu().numpy().view(np.uint64).reshape(-1))
            ms, leaves, levels = time_build(1 << 20, 4, 3)
            t = api.CudaNaryMerkleTree(arity=4)
            t.leaf_count, t.levels = 1 << 20, levels
            idx = torch.arange(1 << 20, device=dev, dtype=torch.int64)
            pb = t.generate_batch_proofs(idx)
            torch.cuda.synchronize(dev)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            res = t.verify_batch_proofs(pb, leaves)
            a1.record(stream)
            torch.cuda.synchronize(dev)
            vms = a0.elapsed_time(a1)
            merkle["quaternary_2p20"] = {"leaves": 1 << 20, "arity": 4, "build_ms": ms, "leaves_per_s": (1 << 20) / (ms * 1e-3),
                                         "verify_full_batch_ms": vms, "proofs_per_s": (1 << 20) / (vms * 1e-3), "all_valid": bool(res.all())}
            merkle["quaternary_2p20"].update(time_build_e2e(leaves, 4, 3))
            del leaves, levels, pb, t
        # config 3's full-batch verification, sharded: every rank verifies its contiguous slice of the 2^20 proofs (no data-path
        # collective), the verdicts are ANDed with one all-reduce(MIN) of a byte
        nq = 1 << 20
        q0, q1 = shard_slice(nq, rank, world)
        leaves4 = torch.empty((nq, 4), dtype=torch.int64, device=dev)
        L.check(L.cuzk_synth_u64_leaves(leaves4.data_ptr(), nq, 3, 0, sp), "synth")
        t4 = api.DeviceMerkleTree(leaves4, arity=4)
        idx = torch.arange(q0, q1, device=dev, dtype=torch.int64)
        pb = t4.generate_batch_proofs(idx)
        lv = leaves4[q0:q1].contiguous()
        res = t4.verify_batch_proofs(pb, lv)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        res = t4.verify_batch_proofs(pb, lv)
        a1.record(stream)
        barrier()
        vms = max_over_ranks(a0.elapsed_time(a1))
        all_ok = sharded_all_valid(res)
        merkle["quaternary_2p20_sharded_verify"] = {"proofs": nq, "n_gpus": world, "verify_ms": vms, "proofs_per_s": nq / (vms * 1e-3),
                                                    "all_valid": all_ok, "scaling": "strong",
                                                    "collective": "none on the data path; verdicts: 1 x all_reduce(MIN) of one byte"}
        t4.close()
        del leaves4, pb, lv, res, idx

        # 8-ary 2^k-leaf tree sharded as subtrees across the ranks THROUGH THE LIBRARY (cuzk_mg_*): every rank keeps all levels of
        # its subtrees in HBM, one NCCL all-gather of 32-byte subtree roots issued by libcuzk_b200.so, top levels on every rank
        # (strong scaling).  torch.distributed only carried the 128-byte NCCL id.
        nleaves = 1 << args.merkle_log2
        mg = api.MultiGpu.from_torch_distributed(local)
        l0, cnt = mg.shard_leaves(nleaves, 8, rank)
        shard = torch.empty((max(cnt, 1), 4), dtype=torch.int64, device=dev)
        L.check(L.cuzk_synth_u64_leaves(shard.data_ptr(), cnt, 4, l0, sp), "synth")
        torch.cuda.synchronize(dev)
        mg_stream = torch.cuda.ExternalStream(mg.stream(0), device=dev)
        tree = mg.build_tree([shard], n=nleaves, arity=8)   # warm-up (also NCCL communicator bring-up)
        tree.close()
        barrier()
        reps = 2
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(mg_stream)
        for k in range(reps):
            tree = mg.build_tree([shard], n=nleaves, arity=8)
            if k + 1 < reps:
                tree.close()
        a1.record(mg_stream)
        barrier()
        ms = max_over_ranks(a0.elapsed_time(a1) / reps)
        root = tree.get_root_hash()
        root_hex = "%064x" % sum(int(v) << (64 * i) for i, v in enumerate(root))
        expected = EXPECTED_ROOTS.get((args.merkle_log2, 8))
        # proofs served from the levels the shards keep, for leaves of this rank, verified against the sharded root
        nproofs = 16384 // world
        rng = np.random.default_rng(1000 + rank)
        pidx = (l0 + rng.integers(0, max(cnt, 1), nproofs)).astype(np.uint64)
        t0 = time.perf_counter()
        proofs = tree.generate_batch_proofs(pidx)
        t1 = time.perf_counter()
        vals = shard[torch.from_numpy((pidx - np.uint64(l0)).astype(np.int64)).to(dev)].cpu().numpy().view(np.uint64)
        verdicts = tree.verify_batch_proofs(proofs, vals)
        proofs_ok = sharded_all_valid(torch.from_numpy(verdicts).to(dev)) and bool((proofs.positions != 0xFFFFFFFF).all())
        perms, width = 0, nleaves                                                  # real nodes x ceil(8 / 2) permutations
        while width > 1:
            width = -(-width // 8)
            perms += 4 * width
        merkle["octary_sharded"] = {"leaves": nleaves, "arity": 8, "n_gpus": world, "subtree_height": tree.subtree_height,
                                    "build_ms": ms, "leaves_per_s": nleaves / (ms * 1e-3),
                                    "scaling": "strong", "collective": "1 x ncclAllGather of 32 B subtree roots, issued by libcuzk_b200.so (cuzk_mg_tree_build)",
                                    "nccl_version": L.cuzk_mg_nccl_version(), "root": root_hex, "root_expected": expected,
                                    "root_ok": (root_hex == expected) if expected else None,
                                    "levels_stored": "every level of every subtree, on the GPU that owns it; top levels replicated",
                                    "proofs_from_shards": {"proofs": nproofs * world, "all_valid": proofs_ok, "prove_ms_rank0": 1e3 * (t1 - t0)},
                                    "roofline_frac": (perms / (ms * 1e-3)) * IMAD_PER_PERM / (imad_peak * world)}
        t1 = SINGLE_GPU_BUILD_MS.get((args.merkle_log2, 8))
        if t1:
            merkle["octary_sharded"]["strong_scaling_efficiency"] = {
                "value": t1 / (world * ms), "single_gpu_build_ms": t1,
                "note": "committed single-GPU time of the same build (profiles/r02_bench_n1.json) / (n_gpus x this build_ms)"}
        if expected:
            assert root_hex == expected, f"sharded root {root_hex} != expected {expected}"
        assert proofs_ok, "a proof served from the sharded levels did not verify"
        tree.close()
        mg.close()
        del shard
        # roofline fractions of the other Merkle configurations (SURVEY 8d convention: permutations x 48 576 multiply-adds)
        if rank == 0:
            def frac(perm_count, ms_):
                return (perm_count / (ms_ * 1e-3)) * IMAD_PER_PERM / imad_peak
            m2, m3 = merkle["binary_50k"], merkle["quaternary_2p20"]
            m2["build_roofline_frac"] = frac(50_006, m2["build_ms"])           # real nodes of the 50 000-leaf binary tree, 1 permutation each
            m2["verify_roofline_frac"] = frac(5000 * 16, m2["verify_5k_ms"])
            m3["build_roofline_frac"] = frac(2 * 349_525, m3["build_ms"])      # (4^10 - 1) / 3 nodes, 2 permutations each
            m3["verify_roofline_frac"] = frac((1 << 20) * 10 * 2, m3["verify_full_batch_ms"])

    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_wall0, t_wall1)

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = cpu_cores()
        sample = cores * 25_000
        kind, v, dt = cpu_reference_run(sample, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{sample} of the same seeded pair hashes on {cores} host threads in {dt:.1f} s (reference PoseidonHash::hash_pair)"}

    if rank == 0:
        achieved = value / world * IMAD_PER_PERM  # per GPU
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": b4096_pageable / REF_PUBLISHED_PAIR_HASHES_PER_S,
            "vs_baseline_note": "like for like: BASELINE.md's 2 145 027 pair hashes/s is the reference's CUDA path on an A100 through its 1 M-hash "
                                "batch-4096 host-vector harness (README.md:134); the numerator is the same harness shape here on pageable host "
                                "memory (e2e.reference_harness_batch4096.pageable), not the device-resident `value`",
            "dtype": "u32x8 (256-bit integers, IMAD.WIDE carry chains)",
            "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 64 * n, "d2h_bytes_per_step": 32 * n,
                    "api": "cuzk_poseidon_hash_pairs(mem=CUZK_MEM_HOST) on pinned host buffers", "steps": e2e_steps, "checked": e2e_ok,
                    "pageable": {"value": pageable_value, "unit": UNIT, "checked": pageable_ok,
                                 "api": "the same call on pageable memory (numpy / std::vector), staged by the library through pinned bounce buffers"},
                    "reference_harness_batch4096": {"value": b4096_value, "pageable": b4096_pageable, "unit": UNIT, "calls": -(-n // b),
                                                    "checked": harness_ok,
                                                    "note": "245 synchronous host calls of <= 4096 pairs; each call is one cooperative-kernel launch that reads and writes pinned host memory directly (pageable: through a pinned bounce copy)"}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "imad", "kernel": "hash_pairs_kernel", "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                         "unit": "T multiply-adds/s (32x32->64)", "frac": achieved / imad_peak, "traffic": ncu_traffic(),
                         "algorithmic_bytes_per_launch": BYTES_PER_PAIR_HASH * n, "algorithmic_imad_per_launch": IMAD_PER_PERM * n,
                         "kernel_ms_per_launch": ms_per_step,
                         "peak_source": "measured in this run by cuzk_imad_peak: best of the IMAD.WIDE.U32 microbenchmarks (register / immediate / constant-bank multiplier, free and carry-chained), whole chip; MEASURED_PEAKS.json has no integer peak",
                         "peak_pipe_model": {"imad_wide_per_sm_per_clk": IMAD_WIDE_PIPE_MODEL_PER_SM_CLK,
                                             "per_s": IMAD_WIDE_PIPE_MODEL_PER_SM_CLK * sm_count * 1.965e9,
                                             "frac_against_model": achieved / (IMAD_WIDE_PIPE_MODEL_PER_SM_CLK * sm_count * 1.965e9),
                                             "note": "4 heavy-pipe cycles per warp IMAD.WIDE = 32 per SM per clock at 1965 MHz; the microbenchmark reaches 28.3 (its own sm__pipe_fmaheavy_cycles_active is in profiles/r02_imad_peak_ncu.txt)"},
                         "note": "achieved counts the reference's algorithmic multiply-adds (48 576 per permutation); the kernel executes fewer (symmetric squarings, MDS layer on the FP64 pipe), so frac can exceed 1 -- the multiplier's measured busy share is in profiles/*_ncu_summary.txt (sm__pipe_fmaheavy_cycles_active)",
                         "imad_per_hash": IMAD_PER_PERM, "pipe_microbench_per_s": peak,
                         "hbm": {"achieved_gbs": value / world * BYTES_PER_PAIR_HASH / 1e9, "peak_gbs": measured_hbm(),
                                 "note": "supporting evidence only: the kernel is integer-pipe bound"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "merkle": merkle,
            "sweep": sweep,
            "latency": latency,
            "sweep_note": "configs[4]: inputs over the whole job, device resident, each rank hashes its contiguous slice; hashes/s aggregate",
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per hash_pairs_kernel launch (1M pairs), taken from the committed
    `ncu --set full` capture (profiles/hash_pairs_traffic.json, written by tools/ncu_summary.py); None when absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "hash_pairs_traffic.json")))
        return {"bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"], "source": t["source"],
                "ncu_pipe_busy_pct": {"fmaheavy (IMAD.WIDE)": t.get("fmaheavy_pct_of_peak"), "alu": t.get("alu_pct"), "fp64": t.get("fp64_pct")},
                "ncu_inst_per_hash": t.get("inst_per_hash"), "ncu_imad_wide_per_hash": t.get("imad_wide_per_hash")}
    except Exception:
        return None


def measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


if __name__ == "__main__":
    main()
