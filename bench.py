#!/usr/bin/env python3
"""bench.py -- Schnorr verifications per second on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload single|double|vargen] [--log2n 20]

A step is one pass of the verify path over one synthetic batch (default: BASELINE.json configs[1], 2^20 single
signatures with 10 % invalid items per GPU; weak scaling, every rank owns its own batch).
  value     whole-job verifications/s with the wire-format inputs already resident in HBM (device entry points,
            CUDA events on the launching stream, max over ranks)
  e2e       same metric through the host-buffer C ABI call a reference user would make (pinned host buffers,
            H2D + D2H inside the timed region)
  roofline  dominant kernel against the measured INT32-multiply peak (profiles/r01_microbench_int.json); this
            path is bound by the integer multiply pipe, not by HBM or tensor cores
  cpu_baseline  the oracle's C port of the reference algorithm on the host cores, bounded sample
--impl reference times that CPU port (the reference is Rust and cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "schnorr_verifications_per_second"
UNIT = "verifications/s"
# canonical algorithmic work, SURVEY.md section 8(d) / Appendix C: MAC32 = one 32x32->64 multiply(-accumulate)
MAC32_M, MAC32_S = 136, 108
MAC32_PER_ITEM = {"single": 4399 * MAC32_M + 2964 * MAC32_S, "double": 7932 * MAC32_M + 5728 * MAC32_S, "vargen": 5808 * MAC32_M + 4248 * MAC32_S}
# per-launch units of the kernels (canonical counts): one equation u*B + c*PK (table + Straus + compare), one subgroup check
MAC32_EQUATION = (1358 + 68 + 2) * MAC32_M + (1008 + 4) * MAC32_S
MAC32_EQUATION_VARGEN = (1512 + 2 * 68 + 2) * MAC32_M + (1008 + 8) * MAC32_S
MAC32_SUBGROUP = (1134 + 68) * MAC32_M + (1008 + 4) * MAC32_S
MAC32_DECODE = 53 * MAC32_M + 272 * MAC32_S
MAC32_PERMUTATION = 865 * MAC32_M + 200 * MAC32_S
BYTES_IN = {"single": 128, "double": 192, "vargen": 160}


def int32_peak():
    """Measured IMAD.WIDE carry-chain issue rate on this pool's B200 (tools/microbench_int.cu), T MAC32/s."""
    path = os.path.join(ROOT, "profiles", "r01_microbench_int.json")
    try:
        data = json.load(open(path))
        for r in data["results"]:
            if r["kernel"].startswith("imad_wide_carry_chain (") :
                return float(r["mult_Tops_per_s"]), "measured: profiles/r01_microbench_int.json (mad.lo.cc/madc.hi.cc chains -> IMAD.WIDE.U32)"
    except Exception:
        pass
    return 32 * 148 * 1.965e9 / 1e12, "nominal: 32 IMAD.WIDE lanes/clk/SM x 148 SMs x 1.965 GHz"


class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, rank, world):
    """CPU arm: the oracle's C port of the reference algorithm (kind 'port'), all host threads, bounded sample."""
    if rank != 0:
        return
    import numpy as np
    from oracle import c_oracle as co
    co.build()
    threads = host_threads()
    sample = int(args.cpu_sample or 1 << 13)
    gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[args.workload]
    ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[args.workload]
    pk, sig, msg = gen(0xB200, sample, threads=threads)
    k = max(1, int(round(0.10 * sample)))
    sig[:k, 0] ^= 1  # 10 % invalid (tampered u), same proportion as the GPU workload
    for _ in range(args.warmup):
        ver(pk, sig, msg, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, _ = ver(pk, sig, msg, threads=threads)
    dt = time.perf_counter() - t0
    assert int((st != 0).sum()) == k
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic", "config": workload_config(args, sample_items=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} {args.workload} signatures (10% invalid) per step, oracle/jjs_oracle.c (reference algorithm restated in C; the Rust crate cannot be built here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_items=None):
    n = 1 << args.log2n
    cfg = {"workload": f"2^{args.log2n} {args.workload} Schnorr signatures per GPU, 10% tampered/invalid (BASELINE.json configs[{ {'single': 1, 'double': 2, 'vargen': 3}[args.workload] }])",
           "items_per_gpu": n, "invalid_fraction": 0.10, "bytes_in_per_item": BYTES_IN[args.workload],
           "l2_policy": "inputs (>=128 MiB) plus >1 GiB of per-step scratch exceed the 126 MB L2; no explicit flush",
           "parallelism": f"{args.gpus} independent shard(s), no collective on the data path"}
    if sample_items is not None:
        cfg["cpu_sample_items"] = sample_items
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="single", choices=["single", "double", "vargen"])
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from jubjub_schnorr_b200 import BatchVerifier
    from jubjub_schnorr_b200 import workload as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout while the communicator is created; stdout must carry one JSON
        # line only, so fd 1 points at stderr until the first collective is through.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    # single-process launch with --gpus N > 1: one context over N devices (the library shards internally)
    devices = [local_rank] if world > 1 else list(range(args.gpus))
    n_local = (1 << args.log2n) * (1 if world > 1 else args.gpus)
    variant = {"single": wl.SINGLE, "double": wl.DOUBLE, "vargen": wl.VARGEN}[args.workload]
    bv = BatchVerifier(devices)
    pk, sig, msg, expected, _ = wl.make_batch(bv, variant, n_local, 0.10, seed=0xB200, rank=rank)
    n_dev = 1 << args.log2n  # items timed on this rank's first device in the device-resident measurement

    def barrier():
        if world > 1:
            dist.barrier()
        for d in devices:
            torch.cuda.synchronize(d)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement ("value") ------------------------------------------------------------
    # one resident shard of n_dev items per device of this process (one device under torchrun)
    shards = []
    for k, dnum in enumerate(devices):
        dev = torch.device("cuda", dnum)
        lo, hi = k * n_dev, (k + 1) * n_dev
        shards.append({
            "k": k, "dev": dev, "lo": lo, "hi": hi, "stream": torch.cuda.current_stream(dev),
            "pk": torch.from_numpy(pk[lo:hi]).to(dev), "sig": torch.from_numpy(sig[lo:hi]).to(dev), "msg": torch.from_numpy(msg[lo:hi]).to(dev),
            "status": torch.empty(n_dev, dtype=torch.uint8, device=dev), "c": torch.empty((n_dev, 32), dtype=torch.uint8, device=dev),
        })

    def step_device():
        for sh in shards:
            bv.verify_device(variant, sh["pk"].data_ptr(), sh["sig"].data_ptr(), sh["msg"].data_ptr(), n_dev, sh["status"].data_ptr(),
                             sh["c"].data_ptr(), stream=sh["stream"].cuda_stream, device_index=sh["k"])

    for _ in range(args.warmup):
        step_device()
    barrier()
    for sh in shards:
        assert np.array_equal(sh["status"].cpu().numpy(), expected[sh["lo"]:sh["hi"]]), "GPU statuses differ from the constructed expectation"
    sampler = ClockSampler(devices[0])
    sampler.start()
    bv.profile(True)
    launches0 = bv.launch_count
    events = []
    barrier()
    for sh in shards:
        with torch.cuda.device(sh["dev"]):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(sh["stream"])
            events.append((e0, e1))
    for _ in range(args.steps):
        step_device()
    for sh, (e0, e1) in zip(shards, events):
        with torch.cuda.device(sh["dev"]):
            e1.record(sh["stream"])
    barrier()
    ms_dev = max_over_ranks(max(e0.elapsed_time(e1) for e0, e1 in events))
    launches = bv.launch_count - launches0
    bv.profile(False)
    stages = bv.profile_collect()
    clocks = sampler.stop()
    n_ranks_devices = world if world > 1 else len(devices)
    value = n_dev * n_ranks_devices * args.steps / (ms_dev * 1e-3)

    # ---- end to end through the host-buffer C ABI ("e2e") ---------------------------------------------------
    h_pk, h_sig, h_msg = torch.from_numpy(pk).pin_memory(), torch.from_numpy(sig).pin_memory(), torch.from_numpy(msg).pin_memory()
    h_status = torch.empty(n_local, dtype=torch.uint8).pin_memory()

    def step_host():
        bv.verify_host_ptr(variant, h_pk.data_ptr(), h_sig.data_ptr(), h_msg.data_ptr(), n_local, h_status.data_ptr(), None)

    for _ in range(max(1, args.warmup - 1)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert np.array_equal(h_status.numpy(), expected), "host-path statuses differ from the constructed expectation"
    total_items = n_local * (world if world > 1 else 1)
    e2e_value = total_items * args.steps / e2e_s

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    peak, peak_src = int32_peak()
    dom = max(("subgroup", "equation"), key=lambda k: stages[k][0])
    slots = {"single": 2, "double": 4, "vargen": 3}[args.workload]
    neq = 2 if args.workload == "double" else 1
    units = {"subgroup": slots * n_dev, "equation": neq * n_dev}[dom]
    per_unit = {"subgroup": MAC32_SUBGROUP, "equation": MAC32_EQUATION_VARGEN if args.workload == "vargen" else MAC32_EQUATION}[dom]
    dom_ms = stages[dom][0] / max(1, args.steps) / len(devices)  # stage time per step per device (all of the stage's launches)
    achieved = units * per_unit / (dom_ms * 1e-3) / 1e12
    step_achieved = value / n_ranks_devices * MAC32_PER_ITEM[args.workload] / 1e12
    roofline = {"bound": "int32_mul", "kernel": f"k_{dom}", "achieved": achieved, "peak": peak, "unit": "TMAC32/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "ms_per_step": dom_ms,
                "algorithmic_mac32_per_unit": per_unit, "units_per_step": units,
                "step": {"achieved": step_achieved, "frac": step_achieved / peak, "mac32_per_item": MAC32_PER_ITEM[args.workload]},
                "stage_ms_per_step": {k: v[0] / max(1, args.steps) / len(devices) for k, v in stages.items()},
                "hbm_secondary": {"algorithmic_GBps": value / n_ranks_devices * (BYTES_IN[args.workload] + 33) / 1e9,
                                  "measured_peak_GBps": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as co
        co.build()
        threads = host_threads()
        sample = int(args.cpu_sample or 1 << 13)
        over = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[args.workload]
        over(pk[:256], sig[:256], msg[:256], threads=threads)
        t0 = time.perf_counter()
        st_o, _ = over(pk[:sample], sig[:sample], msg[:sample], threads=threads)
        dt = time.perf_counter() - t0
        assert np.array_equal(st_o, expected[:sample]) and np.array_equal(st_o, h_status.numpy()[:sample]), "oracle and GPU disagree on the sample"
        cpu = {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample} items of the same batch, oracle/jjs_oracle.c (C restatement of the reference algorithm, {threads} threads); statuses equal the GPU's"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 (8x32-bit Montgomery limbs, IMAD.WIDE)", "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": BYTES_IN[args.workload] * n_local, "d2h_bytes_per_step": n_local,
                        "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "impl": "b200"}
        print(json.dumps(line), flush=True)
    bv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
