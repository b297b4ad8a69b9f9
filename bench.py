#!/usr/bin/env python3
"""Benchmark of the probability stage on the --zscore shuffle batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload mica_ompa|synthetic] [--num-shuffling 1000] [--scaling strong|weak]

Metric (BASELINE.json): shuffled pairs / second of the probability stage of
`ractip --zscore=12 --num-shuffling=1000` (per shuffled pair: 2 single-strand
McCaskill inside/outside + 2 unpaired-window passes + 1 two-strand McCaskill).
One "step" = one pass over the whole shuffle batch.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "shuffled pairs/sec, probability stage of --zscore=12 (McCaskill in/out x2 + unpaired windows x2 + two-strand McCaskill)"
UNIT = "pairs/s"


# --------------------------------------------------------------------------- workload
def make_workload(name: str, num: int, seed: int):
    from ractip_b200 import zscore_shuffles
    seqs = json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())["sequences"]
    if name == "mica_ompa":
        s1, s2 = seqs["MicA"], seqs["ompA"]
        desc = f"MicA(72) x ompA(137), --zscore=12 --num-shuffling={num} --seed={seed} (BASELINE configs[3])"
    elif name == "synthetic":
        rng = np.random.default_rng(20261018)
        s1 = "".join("ACGU"[x] for x in rng.integers(0, 4, 1000))
        s2 = "".join("ACGU"[x] for x in rng.integers(0, 4, 500))
        desc = f"synthetic 1000 x 500 nt, --zscore=12 --num-shuffling={num} --seed={seed} (BASELINE configs[4])"
    else:
        raise SystemExit("unknown workload " + name)
    r1, r2 = zscore_shuffles(s1, s2, num, seed, mode=12, k=2)
    return list(zip(r1, r2)), desc


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mxc = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for nm, val in zip(names, parts[3:7]):
                    if val == "Active":
                        reasons.add(nm)
        if not sm:
            sm = [float(l.split(",")[0]) for _, l in self.rows[-3:] if l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------- CPU oracle timing
def cpu_pairs_per_second(pairs, threads: int, model):
    """Times the CPU restatement (oracle/, kind 'port') on `pairs` with `threads` host threads."""
    from oracle.oracle import Oracle
    o = Oracle(model)
    todo = list(pairs)
    lock = threading.Lock()

    def work():
        while True:
            with lock:
                if not todo:
                    return
                a, b = todo.pop()
            o.rnafold(a, 15)
            o.rnafold(b, 15)
            o.rnaduplex(a, b, 0.1)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work) for _ in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return len(pairs) / dt, dt


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's probability stage cannot be built here (ViennaRNA absent,
    DESIGN.md), so the timed CPU implementation is the oracle port on all host threads."""
    if rank != 0:
        return
    from ractip_b200 import default_model
    pairs, desc = make_workload(args.workload, args.num_shuffling, args.seed)
    cores = os.cpu_count() or 1
    per_step = min(len(pairs), 500) if args.workload == "mica_ompa" else max(2, cores // 4)
    model = default_model()
    for _ in range(min(args.warmup, 1)):
        cpu_pairs_per_second(pairs[:cores], cores, model)
    tot_pairs, tot_t = 0, 0.0
    for s in range(args.steps):
        sample = [pairs[(s * per_step + k) % len(pairs)] for k in range(per_step)]
        _, dt = cpu_pairs_per_second(sample, cores, model)
        tot_pairs += len(sample)
        tot_t += dt
    v = tot_pairs / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample": f"{per_step} shuffled pairs per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{tot_pairs} shuffled pairs of the workload, {cores} threads over pairs "
                                   "(the reference itself is single-threaded, src/ractip.cpp:1494)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    from ractip_b200 import ProbabilityStage, default_model, default_opts
    from ractip_b200._lib import RpPair
    from ractip_b200.stage import REC_DTYPE

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the probability stage has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    all_pairs, desc = make_workload(args.workload, args.num_shuffling, args.seed)
    if args.scaling == "weak" and world > 1:
        # fixed per-GPU work: every rank gets its own full batch (different seed)
        mine, _ = make_workload(args.workload, args.num_shuffling, args.seed + rank)
        total_pairs = len(mine) * world
    else:
        mine = all_pairs[rank::world]   # interleaved shard of the one shuffle batch
        total_pairs = len(all_pairs)

    model = default_model()
    opts = default_opts()
    stage = ProbabilityStage(model, device=local_rank)
    lib = stage.lib
    # One explicit (non-default) stream carries the kernels, the NCCL collective and the
    # timing events.  (The legacy default stream has handle 0, which rp_set_stream reads as
    # "use the context's own stream": events recorded there would not see the kernels.)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    lib.rp_set_stream(stage.ctx, C.c_void_p(stream.cuda_stream))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident batch ("value")
    batch = stage.batch(mine, opts)
    gather_buf = gathered = None
    if world > 1:
        rec_b = batch.total_recs * 12
        up_b = batch.total_upf * 4
        cnt_b = batch.n * 16
        sizes = torch.tensor([rec_b, up_b, cnt_b], device=dev, dtype=torch.int64)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
        rec_b, up_b, cnt_b = [int(x) for x in sizes.tolist()]
        rec_b = (rec_b + 255) // 256 * 256
        up_b = (up_b + 255) // 256 * 256
        cnt_b = (cnt_b + 255) // 256 * 256
        gather_buf = torch.zeros(rec_b + up_b + cnt_b, dtype=torch.uint8, device=dev)
        gathered = torch.empty(world * gather_buf.numel(), dtype=torch.uint8, device=dev)

    def step_resident():
        batch.run()
        if world > 1:
            base = gather_buf.data_ptr()
            stage._check(lib.rp_batch_sparse_device(batch.handle, C.c_void_p(base), batch.total_recs,
                                                    C.c_void_p(base + rec_b), batch.total_upf,
                                                    C.c_void_p(base + rec_b + up_b)))
            dist.all_gather_into_tensor(gathered, gather_buf)   # the single collective of the path

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, launches = [], 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        if world == 1:
            t = stage.last_timing()      # CUDA events around the kernel on its stream
            kern_ms.append(t.ms_total)
            launches += t.kernel_launches
        else:
            launches += 3
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1)
    if world > 1:
        tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    ms_per_step = ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)
    if world > 1:
        # the timed path must have produced real data: every rank's section of the gathered
        # buffer carries counts, and this rank's records equal a fresh rp_run_sparse of its shard
        g = gathered.cpu().numpy().reshape(world, -1)
        for r in range(world):
            cnts = g[r][rec_b + up_b:rec_b + up_b + 16 * (len(all_pairs[r::world]) if args.scaling != "weak" else len(mine))]
            assert cnts.view(np.int32).reshape(-1, 4)[:, :3].sum() > 0, f"rank {r} gathered nothing"
        chk = stage.run_sparse(mine[:3], opts)
        mine_recs = g[rank][:rec_b // 12 * 12].view(REC_DTYPE)
        mine_cnt = g[rank][rec_b + up_b:rec_b + up_b + 16 * len(mine)].view(np.int32).reshape(-1, 4)
        for k, c in enumerate(chk):
            S = batch.slayout[k]
            assert mine_recs[S.x:S.x + int(mine_cnt[k][0])].tolist() == c.x.tolist(), "gathered records differ"
            assert mine_recs[S.z:S.z + int(mine_cnt[k][2])].tolist() == c.z.tolist(), "gathered records differ"

    # ---------------- end to end through the C ABI with host buffers
    n = len(mine)
    arr = (RpPair * max(n, 1))()
    keep = []
    for k, (a, b) in enumerate(mine):
        ba, bb = a.encode(), b.encode()
        keep.append((ba, bb))
        arr[k].s1, arr[k].n1, arr[k].s2, arr[k].n2 = ba, len(ba), bb, len(bb)
    out_bytes = batch.total_floats * 4
    pin = lib.rp_host_alloc(max(out_bytes, 4))
    pin_recs = lib.rp_host_alloc(max(batch.total_recs * 12, 4))
    pin_ups = lib.rp_host_alloc(max(batch.total_upf * 4, 4))
    pin_cnt = lib.rp_host_alloc(max(n * 16, 16))
    # encoded sequences + 3 problem descriptors (72 B) + 3 queue entries per pair
    h2d_bytes = sum(2 * (len(a) + len(b)) + 4 for a, b in mine) + 3 * n * 72 + 3 * n * 4

    def e2e_dense():
        stage._check(lib.rp_run_dense(stage.ctx, arr, n, C.byref(opts), C.c_void_p(pin), batch.total_floats))

    def e2e_sparse():
        stage._check(lib.rp_run_sparse(stage.ctx, arr, n, C.byref(opts), C.c_void_p(pin_recs), batch.total_recs,
                                       C.c_void_p(pin_ups), batch.total_upf, C.c_void_p(pin_cnt)))

    def time_e2e(fn):
        for _ in range(max(1, args.warmup - 1)):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            fn()
        b.record(stream)
        barrier()
        t = a.elapsed_time(b)
        if world > 1:
            tm = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            t = float(tm.item())
        return total_pairs / (t / args.steps * 1e-3)

    e2e_d = time_e2e(e2e_dense)
    e2e_s = time_e2e(e2e_sparse)
    launches_e2e = 0

    # ---------------- roofline of the dominant kernel + CPU baseline (rank 0, N=1)
    roofline, cpu = None, None
    if world == 1:
        fp64_tf, smem_gbs = stage.measure_peaks()
        t = stage.last_timing()
        batch.run()
        tt = stage.last_timing()
        alg = tt.alg_flops
        launch_ms = statistics.mean(kern_ms) if kern_ms else tt.ms_total
        achieved = alg / (launch_ms * 1e-3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get(args.workload)
            except Exception:
                traffic = None
        roofline = {
            "bound": "fp64_fma", "kernel": KERNELS[args.workload][0], "achieved": achieved, "peak": fp64_tf, "unit": "TFLOP/s",
            "frac": achieved / fp64_tf if fp64_tf else None, "traffic": traffic,
            "alg_flops_per_launch": alg, "launch_ms": launch_ms,
            "peak_source": "live fp64-FMA micro-benchmark on this GPU (rp_measure_peaks); "
                           "MEASURED_PEAKS.json holds only HBM/bf16 peaks, which do not bound this path",
            "smem_peak_gbs": smem_gbs,
            "smem_frac_at_16B_per_term": (alg / 2.0 * 16.0 / (launch_ms * 1e-3) / 1e9) / smem_gbs if smem_gbs else None,
        }
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            try:
                roofline["hbm_peak_gbs_measured"] = json.loads(peaks_file.read_text()).get("hbm_gbs")
            except Exception:
                pass
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            # bounded sample: ~5-30 s of CPU work on the box's cores
            sample_n = min(len(all_pairs), 1000 if args.workload == "mica_ompa" else max(2, cores // 4))
            v_all, dt_all = cpu_pairs_per_second(all_pairs[:sample_n], cores, model)
            n_1 = max(1, min(64, sample_n // 4)) if args.workload == "mica_ompa" else 1
            v_1, dt_1 = cpu_pairs_per_second(all_pairs[:n_1], 1, model)
            cpu = {"value": v_all, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {sample_n} shuffled pairs of the workload on {cores} threads ({dt_all:.1f} s)",
                   "single_thread_value": v_1,
                   "single_thread_sample": f"first {n_1} pairs on 1 thread ({dt_1:.1f} s); the reference runs single-threaded (src/ractip.cpp:1494)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "pairs_per_step": total_pairs, "sharding": f"shuffles r::{world}" if world > 1 else "none",
                       "collective": "one all_gather of sparse records" if world > 1 else "none",
                       "l2": "per-step working set (workspace slots of all resident CTAs, > 1 GB) exceeds the 126 MB L2; no flush needed",
                       "kernels": KERNELS[args.workload][1]},
            "clocks": clocks,
            "e2e": {"value": e2e_d, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                    "d2h_bytes_per_step": out_bytes * world, "api": "rp_run_dense (reference layouts, pinned host buffer)"},
            "e2e_sparse": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                           "d2h_bytes_per_step": (batch.total_recs * 12 + batch.total_upf * 4 + n * 16) * world,
                           "api": "rp_run_sparse (thresholded variable lists + up tables)"},
            "gpu_launches": launches,
        }
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)

    lib.rp_host_free(C.c_void_p(pin)); lib.rp_host_free(C.c_void_p(pin_recs))
    lib.rp_host_free(C.c_void_p(pin_ups)); lib.rp_host_free(C.c_void_p(pin_cnt))
    batch.close()
    stage.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dominant kernel (roofline) and routing of each workload (rp_kernel_plan / rp_batch_create)
KERNELS = {
    "mica_ompa": ("mcc_band_kernel (launch shapes <512,1> and <256,2>, timed together)",
                  "mcc_band_kernel<512,1> (n > ~95) + mcc_band_kernel<256,2> (shorter); general kernel for n > 215"),
    "synthetic": ("mcc_persistent<1,10> (general kernel, HBM tables, 128 registers, split sums in bands of 10 diagonals)",
                  "mcc_persistent<1,10>: one CTA per problem and per SM, 1500 / 1000 / 500-nt problems from one cost-ordered queue"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mica_ompa", choices=["mica_ompa", "synthetic"])
    ap.add_argument("--num-shuffling", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
