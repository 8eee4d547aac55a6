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
    """nvidia-smi streaming every 100 ms during the timed region (B200_PROFILING.md, clocks line)."""
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        time.sleep(0.25)  # let the first samples land inside the timed region

    def stop(self):
        samples = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            samples = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.strip()]
        sm = sorted(int(s[0]) for s in samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in samples if len(s) >= 6 for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run_reference(args, rank, world):
    """CPU arm: the oracle's C port of the reference algorithm (kind 'port'), all host threads, bounded sample."""
    if rank != 0:
        return
    if not args.log2n:
        args.log2n = 24 if args.workload == "mixed5" else 20
    import numpy as np
    from oracle import c_oracle as co
    co.build()
    threads = host_threads()
    sample = int(args.cpu_sample or 1 << 13)
    kinds = {"single": ["single"], "double": ["double"], "vargen": ["vargen"], "aggregate": ["aggregate"], "mixed4": ["vargen", "aggregate"],
             "mixed5": ["single", "double"]}[args.workload]
    per = sample // len(kinds)
    jobs = []
    for kind in kinds:
        k = max(1, int(round((0.05 if "aggregate" in args.workload or args.workload == "mixed4" else 0.10) * per)))
        if kind == "aggregate":
            signers = np.random.default_rng(1).choice(np.array([2, 3, 4], dtype=np.uint32), size=per)
            pks, off, sig, msg = co.gen_aggregate(0xB200, signers, threads=threads)
            sig[:k, 0] ^= 1
            jobs.append((lambda pks=pks, off=off, sig=sig, msg=msg: co.verify_aggregate(pks, off, sig, msg, threads=threads)[0], k))
        else:
            gen = {"single": co.gen_single, "double": co.gen_double, "vargen": co.gen_vargen}[kind]
            ver = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[kind]
            pk, sig, msg = gen(0xB200, per, threads=threads)
            sig[:k, 0] ^= 1  # tampered u, same invalid proportion as the GPU workload
            jobs.append((lambda ver=ver, pk=pk, sig=sig, msg=msg: ver(pk, sig, msg, threads=threads)[0], k))
    sample = per * len(kinds)
    for _ in range(args.warmup):
        for fn, _ in jobs:
            fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        results = [fn() for fn, _ in jobs]
    dt = time.perf_counter() - t0
    assert all(int((st != 0).sum()) == k for st, (_, k) in zip(results, jobs))
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (4x64-bit Montgomery limbs)",
        "data": "synthetic", "config": workload_config(args, sample_items=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} {args.workload} items per step, oracle/jjs_oracle.c (reference algorithm restated in C; the Rust crate cannot be built here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


PERMS = {"single": 2, "double": 3, "vargen": 2}
SLOTS = {"single": 2, "double": 4, "vargen": 3}
NEQ = {"single": 1, "double": 2, "vargen": 1}
MAC32_POINT = (53 + 1134 + 68) * MAC32_M + (272 + 1008 + 4) * MAC32_S      # decompress + subgroup check of one point
CONFIG_INDEX = {"single": 1, "double": 2, "vargen": 3, "aggregate": 3, "mixed4": 3, "mixed5": 4}


def aggregate_mac32(counts):
    """Canonical MAC32 of aggregate-key items (aggregate(n)), split into the stages that do the work."""
    import numpy as np
    c = np.asarray(counts, dtype=np.int64)
    perms = c * ((2 + 2 * c + 3) // 4)
    decode = int((c * MAC32_DECODE).sum()) + len(c) * MAC32_POINT                      # n signer keys (decode only) + R
    agg = int((c * ((68 + 42 * 9) * MAC32_M + 4 * MAC32_S) + perms * MAC32_PERMUTATION).sum()) + len(c) * ((252 * 3 + 40) * MAC32_M + (252 * 4 + 255) * MAC32_S + MAC32_POINT - MAC32_DECODE)
    chall = len(c) * 2 * MAC32_PERMUTATION
    eq = len(c) * MAC32_EQUATION
    return {"decode": decode, "aggregate": agg, "challenge": chall, "equation": eq}


def plan_parts(args, n_gpus_total):
    """[(kind, items_per_gpu, invalid_fraction)] for one GPU."""
    n = 1 << args.log2n
    if args.workload in ("single", "double", "vargen", "aggregate"):
        return [(args.workload, n, 0.05 if args.workload == "aggregate" else 0.10)]
    if args.workload == "mixed4":   # BASELINE.json configs[3]: var-generator + aggregate-key halves, 5 % invalid
        return [("vargen", n // 2, 0.05), ("aggregate", n // 2, 0.05)]
    if args.workload == "mixed5":   # BASELINE.json configs[4]: 2^log2n items in total over all GPUs, half single half double
        per = max(2, n // n_gpus_total)
        return [("single", per // 2, 0.10), ("double", per // 2, 0.10)]
    raise SystemExit("unknown workload")


def workload_config(args, sample_items=None):
    n = 1 << args.log2n
    desc = {
        "single": f"2^{args.log2n} single Schnorr signatures per GPU, 10% tampered/invalid",
        "double": f"2^{args.log2n} double Schnorr signatures (JJSCHDBL transcript, G and G') per GPU, 10% tampered/invalid",
        "vargen": f"2^{args.log2n} var-generator signatures (distinct generator per item) per GPU, 10% tampered/invalid",
        "aggregate": f"2^{args.log2n} SpeedyMuSig aggregate-key verifications (2-4 signers) per GPU, 5% invalid",
        "mixed4": f"2^{args.log2n - 1} var-generator + 2^{args.log2n - 1} aggregate-key (2-4 signers) verifications per GPU, 5% invalid",
        "mixed5": f"2^{args.log2n} items in total, half single half double, 10% invalid, split contiguously over {args.gpus} GPU(s)",
    }[args.workload]
    cfg = {"workload": f"{desc} (BASELINE.json configs[{CONFIG_INDEX[args.workload]}])", "items_per_gpu": n if args.workload != "mixed5" else n // args.gpus,
           "l2_policy": "inputs (>=128 MiB) plus >1 GiB of per-step scratch exceed the 126 MB L2; no explicit flush",
           "parallelism": f"{args.gpus} independent shard(s), no collective on the data path"}
    if sample_items is not None:
        cfg["cpu_sample_items"] = sample_items
    return cfg


class Part:
    """One homogeneous sub-batch resident on one device."""

    def __init__(self, bv, wl, kind, n, frac, seed, rank, dev_index, torch):
        import numpy as np
        self.kind, self.n, self.bv, self.dev_index, self.torch = kind, n, bv, dev_index, torch
        self.variant = {"single": wl.SINGLE, "double": wl.DOUBLE, "vargen": wl.VARGEN, "aggregate": None}[kind]
        if kind == "aggregate":
            self.pk, self.off, self.sig, self.msg, self.expected, cls = wl.make_aggregate_batch(bv, n, frac, seed=seed, rank=rank)
            self.counts = np.diff(self.off.astype(np.int64))
            names = wl.AGG_CLASSES
        else:
            self.pk, self.sig, self.msg, self.expected, cls = wl.make_batch(bv, self.variant, n, frac, seed=seed, rank=rank)
            self.off = None
            names = wl.CLASSES
        # Items whose equations the equation kernel evaluates: everything decodes and the keys are valid, i.e. status
        # Ok / InvalidSignature, or InvalidPoint caused by the signature point alone (the library drops the rest from the
        # kernel's work list).  The kernel's roofline counts these units only.
        sig_point_only = np.array([nm.startswith("R_") and st == 2 for nm, st in names], dtype=bool)
        cls = np.asarray(cls)
        self.n_equation_items = int(((self.expected <= 1) | ((cls >= 0) & sig_point_only[np.clip(cls, 0, len(names) - 1)])).sum())
        self.h2d = self.pk.nbytes + self.sig.nbytes + self.msg.nbytes + (self.off.nbytes if self.off is not None else 0)

    def to_device(self, dev):
        t = self.torch
        self.dev = dev
        self.stream = t.cuda.current_stream(dev)
        self.d = [t.from_numpy(x).to(dev) for x in ((self.pk, self.sig, self.msg) if self.off is None else (self.pk, self.sig, self.msg, self.off))]
        self.d_status = t.empty(self.n, dtype=t.uint8, device=dev)
        self.d_c = t.empty((self.n, 32), dtype=t.uint8, device=dev)

    def pin(self):
        t = self.torch
        self.h = [t.from_numpy(x).pin_memory() for x in (self.pk, self.sig, self.msg)]
        self.h_status = t.empty(self.n, dtype=t.uint8).pin_memory()

    def step_device(self):
        if self.off is None:
            self.bv.verify_device(self.variant, self.d[0].data_ptr(), self.d[1].data_ptr(), self.d[2].data_ptr(), self.n, self.d_status.data_ptr(),
                                  self.d_c.data_ptr(), stream=self.stream.cuda_stream, device_index=self.dev_index)
        else:
            self.bv.verify_aggregate_device(self.d[0].data_ptr(), self.d[3].data_ptr(), self.off, self.d[1].data_ptr(), self.d[2].data_ptr(), self.n,
                                            self.d_status.data_ptr(), self.d_c.data_ptr(), None, stream=self.stream.cuda_stream, device_index=self.dev_index)

    def step_host(self, bv_host):
        if self.off is None:
            bv_host.verify_host_ptr(self.variant, self.h[0].data_ptr(), self.h[1].data_ptr(), self.h[2].data_ptr(), self.n, self.h_status.data_ptr(), None)
        else:
            st = bv_host.verify_aggregate(self.h[0].numpy(), self.off, self.h[1].numpy(), self.h[2].numpy())
            self.h_status.numpy()[:] = st

    def canonical_mac32(self):
        """{stage: canonical MAC32 of this part for one step}"""
        if self.kind == "aggregate":
            d = aggregate_mac32(self.counts)
            d["equation"] = self.n_equation_items * MAC32_EQUATION
            return d
        chall = PERMS[self.kind] * MAC32_PERMUTATION
        eq = NEQ[self.kind] * (MAC32_EQUATION_VARGEN if self.kind == "vargen" else MAC32_EQUATION)
        # the rest of SURVEY's per-item figure is point decoding and the subgroup checks it counts (those of the key points)
        return {"decode": self.n * (MAC32_PER_ITEM[self.kind] - chall - eq), "challenge": self.n * chall, "equation": self.n_equation_items * eq,
                "aggregate": 0}

    def cpu_check(self, co, sample, threads):
        import numpy as np
        s = min(sample, self.n)
        if self.kind == "aggregate":
            st, _, _ = co.verify_aggregate(self.pk[: self.off[s]], self.off[: s + 1], self.sig[:s], self.msg[:s], threads=threads)
        else:
            over = {"single": co.verify_single, "double": co.verify_double, "vargen": co.verify_vargen}[self.kind]
            st, _ = over(self.pk[:s], self.sig[:s], self.msg[:s], threads=threads)
        assert np.array_equal(st, self.expected[:s]), "oracle and constructed expectation disagree on the sample"
        return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="single", choices=["single", "double", "vargen", "aggregate", "mixed4", "mixed5"])
    ap.add_argument("--log2n", type=int, default=0, help="log2 of the items per GPU (mixed5: of the whole job); default 20 (mixed5: 24)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if not args.log2n:
        args.log2n = 24 if args.workload == "mixed5" else 20
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
    # under torchrun: one device per process; plain launch with --gpus N: one context over N devices
    devices = [local_rank] if world > 1 else list(range(args.gpus))
    n_gpus_total = world if world > 1 else len(devices)
    bv = BatchVerifier(devices)          # device entry points, one shard per device
    plan = plan_parts(args, n_gpus_total)
    shards = []                          # per device of this process: list of parts
    for k, dnum in enumerate(devices):
        dev = torch.device("cuda", dnum)
        parts = [Part(bv, wl, kind, n, frac, 0xB200 + 17 * j, rank * len(devices) + k, k, torch) for j, (kind, n, frac) in enumerate(plan)]
        for p in parts:
            p.to_device(dev)
        shards.append(parts)
    all_parts = [p for parts in shards for p in parts]
    items_per_device = sum(n for _, n, _ in plan)

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
    def step_device():
        for p in all_parts:
            p.step_device()

    for _ in range(args.warmup):
        step_device()
    barrier()
    for p in all_parts:
        assert np.array_equal(p.d_status.cpu().numpy(), p.expected), f"GPU statuses differ from the constructed expectation ({p.kind})"
    sampler = ClockSampler(devices[0])
    sampler.start()
    launches0 = bv.launch_count
    events = []
    barrier()
    for parts in shards:
        with torch.cuda.device(parts[0].dev):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(parts[0].stream)
            events.append((e0, e1))
    for _ in range(args.steps):
        step_device()
    for parts, (e0, e1) in zip(shards, events):
        with torch.cuda.device(parts[0].dev):
            e1.record(parts[0].stream)
    barrier()
    ms_dev = max_over_ranks(max(e0.elapsed_time(e1) for e0, e1 in events))
    launches = bv.launch_count - launches0
    clocks = sampler.stop()
    # Stage breakdown from a second, untimed pass of the same K steps: with the per-stage timers on, the library keeps
    # every launch on one stream (the timed pass above overlaps neighbouring sub-chunks on two streams, which per-stage
    # CUDA events cannot attribute), so the stage times add up to slightly more than ms_per_step.
    bv.profile(True)
    for _ in range(args.steps):
        step_device()
    barrier()
    bv.profile(False)
    stages = bv.profile_collect()
    value = items_per_device * n_gpus_total * args.steps / (ms_dev * 1e-3)

    # ---- end to end through the host-buffer C ABI ("e2e") ---------------------------------------------------
    # the host entry points shard one host batch over the context's devices themselves
    host_parts = shards[0] if len(devices) == 1 else [Part(bv, wl, kind, n * len(devices), frac, 0xE2E + 17 * j, rank, 0, torch) for j, (kind, n, frac) in enumerate(plan)]
    for p in host_parts:
        p.pin()

    def step_host():
        for p in host_parts:
            p.step_host(bv)

    for _ in range(max(1, args.warmup - 1)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    for p in host_parts:
        assert np.array_equal(p.h_status.numpy(), p.expected), f"host-path statuses differ from the constructed expectation ({p.kind})"
    e2e_value = items_per_device * n_gpus_total * args.steps / e2e_s
    h2d = sum(p.h2d for p in host_parts)
    d2h = sum(p.n for p in host_parts)

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------
    peak, peak_src = int32_peak()
    canon = {k: sum(p.canonical_mac32()[k] for p in shards[0]) for k in ("decode", "challenge", "aggregate", "equation")}
    stage_ms = {k: v[0] / max(1, args.steps) / len(devices) for k, v in stages.items()}
    dom = max(("decode", "challenge", "aggregate", "equation"), key=lambda k: stage_ms[k])
    achieved = canon[dom] / (stage_ms[dom] * 1e-3) / 1e12
    step_mac32 = sum(canon.values())
    step_achieved = step_mac32 * n_gpus_total / (ms_dev / args.steps * 1e-3) / 1e12 / n_gpus_total
    bytes_in = sum(p.h2d for p in shards[0])
    hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    traffic = None
    try:  # DRAM bytes of the dominant kernel per step, from the committed ncu capture (bytes per unit x units in this step)
        import glob
        traffic_file = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))[-1]   # newest capture
        tr = json.load(open(traffic_file))["kernels"]["k_" + dom]
        units = {"decode": sum(p.n * SLOTS.get(p.kind, 1) for p in shards[0]), "challenge": sum(p.n for p in shards[0]),
                 "equation": sum(p.n * NEQ.get(p.kind, 1) for p in shards[0]), "aggregate": sum(p.n for p in shards[0] if p.kind == "aggregate")}[dom]
        traffic = tr["dram_bytes_per_unit"] * units
    except Exception:
        pass
    roofline = {"bound": "int32_mul", "kernel": {"decode": "k_decode", "challenge": "k_challenge", "aggregate": "k_aggregate", "equation": "k_equation"}[dom],
                "achieved": achieved, "peak": peak, "unit": "TMAC32/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_unit": "DRAM bytes per step (ncu dram__bytes_read.sum + dram__bytes_write.sum, the newest profiles/r*_ncu_traffic.json); secondary: the bound is the multiply pipe",
                "peak_source": peak_src,
                "ms_per_step": stage_ms[dom], "algorithmic_mac32_per_step": canon[dom],
                "note": "achieved = canonical MAC32 (SURVEY 8(d) / Appendix C operation counts at 136/108 MAC32 per field mul/sqr) of the kernel's units / its measured time "
                        "(equation kernel: only the equations it evaluates -- items that fail to decode or have an invalid key never reach it); "
                        "the implementation executes fewer multiplies than the canonical algorithm (Tate subgroup test, integer MDS, half-size scalars), so fractions above 1 are possible",
                "step": {"achieved": step_achieved, "frac": step_achieved / peak, "mac32_per_step_per_gpu": step_mac32},
                "stage_ms_per_step": stage_ms,
                "stage_timing": "per-stage CUDA events on the launching stream in a second pass of the same steps with the launches serialised "
                                "(the timed pass overlaps neighbouring sub-chunks on two streams); ms_per_step of the kernel above is from that pass",
                "hbm_secondary": {"algorithmic_GBps": (bytes_in + items_per_device * 33) / (ms_dev / args.steps * 1e-3) / 1e9, "measured_peak_GBps": hbm_peak}}

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as co
        co.build()
        threads = host_threads()
        sample = int(args.cpu_sample or 1 << 13)
        shards[0][0].cpu_check(co, 256, threads)
        t0 = time.perf_counter()
        done = sum(p.cpu_check(co, max(1, sample * p.n // items_per_device), threads) for p in shards[0])
        dt = time.perf_counter() - t0
        cpu = {"value": done / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {done} items of the same batch(es), oracle/jjs_oracle.c (C restatement of the reference algorithm, {threads} threads); statuses equal the GPU's"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "mixed5" else "weak", "vs_baseline": None,
                "dtype": "u32 (8x32-bit Montgomery limbs, IMAD.WIDE)", "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "impl": "b200"}
        print(json.dumps(line), flush=True)
    bv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
