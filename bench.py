#!/usr/bin/env python
"""bench.py -- headline benchmark of the rtl-wmbus hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload t1x2|s1|both|d3|d3s]

One step = one pass of the hot path (cu8 -> ... -> datagram lines) over one synthetic capture:
BASELINE.json config 2 by default -- 1 GiB of 1.6 MS/s cu8 with two T1 emitters, T1+C1 chain
(`-p S`).  Printed JSON (one line, rank 0):
  value      IQ Msamples/s, whole job over all ranks, capture already resident in HBM
  e2e        the same through the C ABI with HOST (pinned) input: H2D copy, kernels, frame D2H
             and host framing all inside the timed region
  roofline   algorithmic bytes (2 + chains*(1 + 2/8)/d per input sample) divided by
             the CUDA-event time of the per-sample kernels (demod + bit-sync), vs measured HBM peak
  cpu_baseline  the reference's own -O3 build (oracle/_ref/rtl_wmbus) on a bounded prefix of the
             same capture, one process (the reference is single-threaded)

  strong     BASELINE config 4 in the same run: ONE 4 GiB 2.4 MS/s capture, -d 3, decoded in time chunks by the
             N ranks (strong scaling; `value`/`e2e` as above, for that capture)

--impl reference times the reference CPU implementation with every usable host CPU (one process per CPU,
each decoding its own copy of a bounded slice of the workload).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IQ Msamples/s (1.6MS/s cu8)"
UNIT = "Msamples/s"


def algorithmic_bytes_per_sample(chains, d):
    return 2.0 + chains * (1.0 + 2.0 / 8.0) / d


def workload_def(name, mib=1024):
    size = f"{mib / 1024:g} GiB"
    if name == "t1x2":
        return dict(emitters="t1x2", flags="-p S", chains=1, d=2, fs=1.6e6,
                    desc=f"{size} synthetic 1.6 MS/s cu8, two T1 emitters, T1+C1 chain (-p S)")
    if name == "s1":
        return dict(emitters="s1", flags="-p T", chains=1, d=2, fs=1.6e6,
                    desc=f"{size} synthetic 1.6 MS/s cu8, two S1 emitters, S1 chain (-p T)")
    if name == "both":
        return dict(emitters="mixed", flags="", chains=2, d=2, fs=1.6e6,
                    desc=f"{size} synthetic 1.6 MS/s cu8, T1/C1/S1 emitters, both chains (default flags)")
    if name == "d3":         # BASELINE config 4's signal: 2.4 MS/s, decimation 3 (general front end, no d = 2 fast path)
        return dict(emitters="mixed", flags="-d 3", chains=2, d=3, fs=2.4e6,
                    desc=f"{size} synthetic 2.4 MS/s cu8, T1/C1/S1 emitters, both chains, -d 3")
    if name == "d3s":        # ... its -s variant: capture centred on 868.625 MHz, T1/C1 at +325 kHz, S1 at -325 kHz
        return dict(emitters="mixed", flags="-d 3 -s", chains=2, d=3, fs=2.4e6, shift=325e3,
                    desc=f"{size} synthetic 2.4 MS/s cu8 centred on 868.625 MHz, T1/C1/S1 emitters, both chains, -d 3 -s")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU during the timed region, through NVML inside this process (a query
    costs microseconds; spawning nvidia-smi every 200 ms from every rank perturbed the other ranks' launches) with
    nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                return int(ids[index])
        return index

    def _sample_nvml(self):
        n = self.nvml
        mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        g = lambda name: getattr(n, name, 0)
        flags = [bool(r & g("nvmlClocksThrottleReasonHwSlowdown")), bool(r & g("nvmlClocksThrottleReasonHwThermalSlowdown")),
                 bool(r & g("nvmlClocksThrottleReasonSwThermalSlowdown")), bool(r & g("nvmlClocksThrottleReasonSwPowerCap"))]
        return [str(mhz), str(self.max_mhz)] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 6:
                        self.samples.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.02 if self.nvml else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "source": "nvml" if self.nvml else "nvidia-smi"}


def host_cpus():
    """CPUs this process may really use: min(cpu_count, scheduler affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    detail = {"cpu_count": n}
    try:
        a = len(os.sched_getaffinity(0))
        detail["affinity"] = a
        n = min(n, a)
    except (AttributeError, OSError):
        pass
    quota = None
    try:                                                        # cgroup v2
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except (OSError, ValueError):
        try:                                                    # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and p > 0:
                quota = q / p
        except (OSError, ValueError):
            pass
    if quota is not None:
        detail["cgroup_quota"] = round(quota, 2)
        n = min(n, max(1, int(quota + 0.5)))
    return max(1, n), detail


def config_dict(args, wl):
    """`config` of the JSON line -- the same dict in both arms (ours and --impl reference)."""
    return {"workload": wl["desc"], "flags": wl["flags"], "capture_mib_per_gpu": args.mib,
            "sharding": "one independent capture per GPU; NCCL all-reduce of packet counters only",
            "l2": "input (1 GiB) and intermediates are larger than L2; no flush needed",
            "device_batch_mib": min(args.batch_mib or args.mib, args.mib, 1024), "e2e_batch_mib": args.e2e_batch_mib}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_step():
    """What the committed ncu capture of one 1 GiB step says (profiles/ncu_step.json, written by profiles/ncu_extract.py
    from the `ncu --page raw --csv` dump -- nothing in it is typed in by hand): DRAM bytes per step over the captured
    kernels, and issue-active / DRAM percentages of the two largest kernels."""
    p = os.path.join(ROOT, "profiles", "ncu_step.json")
    if not os.path.exists(p):
        return None, None
    j = json.load(open(p))
    top = sorted(j.get("kernels", {}).items(), key=lambda kv: -kv[1].get("ms", 0.0))[:2]
    brief = {name: {k: v.get(k) for k in ("ms", "issue_active_pct", "dram_pct", "warps_active_pct", "threads_per_inst")}
             for name, v in top}
    brief["source"] = "profiles/ncu_step.json <- " + str(j.get("source"))
    return j.get("traffic_bytes_per_step"), brief


def ref_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "rtl_wmbus")
    if os.path.exists(p):
        return p, "reference"
    p = os.path.join(ROOT, "oracle", "_ref", "oracle_cli")
    if not os.path.exists(p):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    return p, "port"


def time_reference(path_to_capture, flags, procs):
    """Run `procs` copies of the reference CPU program on the file; returns wall seconds."""
    exe, _ = ref_binary()
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe] + flags.split(), stdin=open(path_to_capture, "rb"), stdout=subprocess.DEVNULL)
          for _ in range(procs)]
    for p in ps:
        p.wait()
    return time.perf_counter() - t0


def run_reference_arm(args, wl):
    """The reference's own CPU implementation of the path on the box's host cores: P = the CPUs this process may use
    (affinity and cgroup quota, not the machine's core count), one single-threaded reference process per CPU, each
    decoding its own copy of a slice of the workload.  The slice is 256 MiB (spawn cost < 0.1 %) unless
    (steps + warmup) of those would run longer than ~3 minutes, then it shrinks (not below 64 MiB)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    procs, cpu_detail = host_cpus()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(shm, f"wmbus_ref_slice_{os.getpid()}.cu8")
    full = 256 << 20
    buf, _ = synth.synth_capture(full, fs=wl["fs"], emitters=synth.default_emitters(wl["emitters"]),
                                 seed=0xB2000000 + 16 * 2, device="cpu", center_shift_hz=wl.get("shift", 0.0))
    buf.numpy().tofile(path)
    try:
        # calibrate: one process on the first 32 MiB
        cal = os.path.join(shm, f"wmbus_ref_cal_{os.getpid()}.cu8")
        buf[:32 << 20].numpy().tofile(cal)
        t_cal = min(time_reference(cal, wl["flags"], 1) for _ in range(2))
        os.unlink(cal)
        solo_rate = (32 << 20) / 2 / t_cal                                  # IQ samples/s of one process alone
        budget_s = 170.0
        slice_bytes = full
        while slice_bytes > (64 << 20) and (args.steps + args.warmup) * (slice_bytes / 2 / solo_rate) * 1.15 > budget_s:
            slice_bytes //= 2
        if slice_bytes != full:
            buf[:slice_bytes].numpy().tofile(path)
        for _ in range(args.warmup):
            time_reference(path, wl["flags"], procs)
        t = 0.0
        for _ in range(args.steps):
            t += time_reference(path, wl["flags"], procs)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    n_iq = slice_bytes // 2 * procs
    value = n_iq * args.steps / t / 1e6
    _, kind = ref_binary()
    emit(({
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, wl),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": procs, "kind": kind,
                         "sample": f"{procs} processes x {slice_bytes >> 20} MiB slice of the workload per step (the reference "
                                   f"is single-threaded: one process per usable host CPU; {cpu_detail})",
                         "per_process_msamples_s": round(value / procs, 3),
                         "one_process_alone_msamples_s": round(solo_rate / 1e6, 3)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_time_sharded(args, wl, pkg, shard, lib, cap, plan, rank, world, local, mib=None, steps=None, warmup=None):
    """Strong scaling: ONE capture, rank g decodes the time chunk [S_g, S_g+1) from a warm-up halo and proves the
    re-join with its left neighbour's boundary state (two digests per rank, all-gathered).  Returns the JSON object
    (rank 0) or None."""
    import torch
    import torch.distributed as dist
    mib = mib or args.mib
    steps = steps or args.steps
    warmup = args.warmup if warmup is None else warmup
    n_bytes = mib << 20
    n_iq = n_bytes // 2
    d = wl["d"]
    host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(cap)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    results = {}
    for leg, base, push_name in (("device", cap.data_ptr(), "push_device"), ("host", host.data_ptr(), "push")):
        ctx = pkg.WmbusB200(wl["flags"], device=local, lib=lib,
                            max_batch_mib=min(args.batch_mib or mib, mib, 1024) if leg == "device" else args.e2e_batch_mib)
        push = lambda lo, hi, c=ctx, b=base, f=push_name: getattr(c, f)(b + lo, hi - lo)
        lines, rounds = None, 0
        for _ in range(max(1, warmup)):
            lines, rounds = shard.decode_time_sharded(ctx, push, n_bytes, d)
        l0 = ctx.stats().kernel_launches
        sampler = ClockSampler(local) if leg == "device" else None
        if sampler: sampler.start()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            lines, rounds = shard.decode_time_sharded(ctx, push, n_bytes, d)
        barrier()
        t = max_over_ranks(time.perf_counter() - t0)
        results[leg] = dict(t=t, lines=lines, rounds=rounds, launches=ctx.stats().kernel_launches - l0,
                            clocks=sampler.stop() if sampler else None, st=ctx.stats())
        ctx.close()
    del host
    assert results["device"]["lines"] == results["host"]["lines"], "host-input and device-input legs disagree"
    totals = shard.reduce_counts(shard.count_lines(results["device"]["lines"]), device="cuda")
    if rank != 0:
        return None
    dv, hv = results["device"], results["host"]
    k = shard.chunk_bounds(n_bytes, d, world)
    return {
        "metric": METRIC, "value": round(n_iq * steps / dv["t"] / 1e6, 1), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * dv["t"] / steps, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "flags": wl["flags"], "capture_mib_total": mib,
                   "sharding": "time chunks of ONE capture: halo warm start, boundary states compared over "
                               "all_gather (2 x 32 B per rank), NCCL all-reduce of packet counters",
                   "chunk_bounds_iq": k, "halo_rounds": dv["rounds"],
                   "l2": "input and intermediates are larger than L2; no flush needed"},
        "clocks": dv["clocks"],
        "e2e": {"value": round(n_iq * steps / hv["t"] / 1e6, 1), "unit": UNIT,
                "h2d_bytes_per_step": int((hv["st"].h2d_bytes) // (steps + max(1, warmup))),
                "d2h_bytes_per_step": int((hv["st"].d2h_bytes) // (steps + max(1, warmup))),
                "ms_per_step": round(1e3 * hv["t"] / steps, 3)},
        "gpu_launches": int(dv["launches"]),
        "packets": dict(totals, planted=len(plan)),
    }


_JSON_OUT = None


def emit(obj):
    """the one JSON line, on the process's real stdout"""
    f = _JSON_OUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def main():
    # Libraries write to fd 1 behind Python's back (NCCL prints its version line there when the box sets NCCL_DEBUG):
    # everything but the result line goes to stderr, so that stdout holds exactly one JSON line.
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="t1x2")
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--batch-mib", type=int, default=0, help="device batch size (default: the whole capture, at most 1 GiB)")
    ap.add_argument("--e2e-batch-mib", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--warm", type=int, default=0, help="bit-sync warm-up samples (default: library)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (time chunks of one 2.4 MS/s -d 3 capture)")
    ap.add_argument("--strong-mib", type=int, default=4096)
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--sharding", default="captures", choices=["captures", "time"],
                    help="N>1: one independent capture per GPU (weak scaling, default) or time chunks of ONE capture "
                         "(strong scaling, SURVEY 8e / BASELINE config 4)")
    args = ap.parse_args()
    wl = workload_def(args.workload, args.mib)
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = importlib.import_module("rtl-wmbus_b200")
    synth = importlib.import_module("rtl-wmbus_b200.synth")
    shard = importlib.import_module("rtl-wmbus_b200.shard")
    lib = pkg.load_library()                 # no fallback: raises when the CUDA library is missing

    n_bytes = args.mib << 20
    n_iq = n_bytes // 2
    time_sharded = args.sharding == "time"
    # one independent capture per rank (weak scaling; BASELINE config 5's sharding rule), or the same capture on
    # every rank, of which each decodes its time chunk (strong scaling; config 4's rule)
    cap, plan = synth.synth_capture(n_bytes, fs=wl["fs"], emitters=synth.default_emitters(wl["emitters"]),
                                    seed=shard.capture_seed(4 if time_sharded else 2, 0 if time_sharded else rank),
                                    device="cuda", center_shift_hz=wl.get("shift", 0.0))
    torch.cuda.synchronize()
    if time_sharded:
        out = run_time_sharded(args, wl, pkg, shard, lib, cap, plan, rank, world, local)
        if rank == 0:
            emit(out)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    tune = dict(max_batch_mib=min(args.batch_mib or args.mib, args.mib, 1024))      # rings and candidate lists are sized for <= 1 GiB batches
    if args.chunk: tune["chunk_samples"] = args.chunk
    if args.warm: tune["warmup_samples"] = args.warm
    ctx = pkg.WmbusB200(wl["flags"], device=local, lib=lib, **tune)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident leg (value) ----------------
    lines = None
    for _ in range(args.warmup):
        ctx.reset()
        lines = ctx.process_device(cap.data_ptr(), n_bytes, flush=True, raw=True)
    launches0 = ctx.stats().kernel_launches
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    k1_ms = k2_ms = dev_ms = 0.0
    for _ in range(args.steps):
        ctx.reset()
        lines = ctx.process_device(cap.data_ptr(), n_bytes, flush=True, raw=True)     # the C ABI's text, as a user's C code gets it
        st = ctx.stats()
        k1_ms += st.demod_kernel_ms; k2_ms += st.bitsync_kernel_ms; dev_ms += st.batch_device_ms
    t_mine = time.perf_counter() - t0
    barrier()
    t_dev = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop()
    per_rank_ms = [round(1e3 * t_mine / args.steps, 3)]
    if world > 1:
        tt = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(tt, torch.tensor([1e3 * t_mine / args.steps], dtype=torch.float64, device="cuda"))
        per_rank_ms = [round(float(x.item()), 3) for x in tt]
    st = ctx.stats()
    launches = st.kernel_launches - launches0
    value = world * n_iq * args.steps / t_dev / 1e6

    # ---------------- host-input leg (e2e) ----------------
    host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
    host.copy_(cap)
    torch.cuda.synchronize()
    ctx_e = pkg.WmbusB200(wl["flags"], device=local, lib=lib, max_batch_mib=args.e2e_batch_mib,
                          **{k: v for k, v in tune.items() if k != "max_batch_mib"})
    e_lines = None
    for _ in range(max(1, args.warmup)):
        ctx_e.reset()
        e_lines = ctx_e.process(host.data_ptr(), n_bytes, flush=True, raw=True)
    d2h0 = ctx_e.stats().d2h_bytes
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx_e.reset()
        e_lines = ctx_e.process(host.data_ptr(), n_bytes, flush=True, raw=True)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    d2h = (ctx_e.stats().d2h_bytes - d2h0) // args.steps
    e2e_value = world * n_iq * args.steps / t_e2e / 1e6
    assert e_lines == lines, "host-input and device-input legs disagree"
    lines = pkg.WmbusB200.split_lines(lines)

    # ---------------- packet counters: the only collective on this path ----------------
    totals = shard.reduce_counts(shard.count_lines(lines), device="cuda")     # NCCL all-reduce over NVLink

    del host, ctx_e

    # ---------------- strong scaling over time chunks (BASELINE config 4), same run, reported under "strong" ----------------
    strong = None
    if not args.no_strong:
        wl4 = workload_def("d3", args.strong_mib)
        cap4, plan4 = synth.synth_capture(args.strong_mib << 20, fs=wl4["fs"], emitters=synth.default_emitters(wl4["emitters"]),
                                          seed=shard.capture_seed(4, 0), device="cuda")
        torch.cuda.synchronize()
        strong = run_time_sharded(args, wl4, pkg, shard, lib, cap4, plan4, rank, world, local,
                                  mib=args.strong_mib, steps=args.strong_steps, warmup=1)
        del cap4

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        abytes = algorithmic_bytes_per_sample(wl["chains"], wl["d"]) * n_iq
        # the per-sample path of one step on the device clock: first demod kernel -> last bit-sync kernel (the demod
        # kernel of one batch overlaps the bit-sync kernels of the batch before, so this is less than the sum of the two)
        kern_s = dev_ms / args.steps / 1e3
        achieved = abytes / kern_s / 1e9 if kern_s > 0 else None
        # the committed ncu capture is of the default workload (1 GiB, -p S): other workloads carry no traffic figure
        traffic, ncu_brief = ncu_step() if (args.workload == "t1x2" and args.mib == 1024) else (None, None)
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(1e3 * t_dev / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl),
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": n_bytes,
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": round(1e3 * t_e2e / args.steps, 3)},
            "gpu_launches": int(launches),
            "per_rank_ms_per_step": per_rank_ms,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1) if achieved else None, "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4) if achieved else None,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "k1_demod_kernel + bit-sync kernels: the whole per-sample path of a step, first demod kernel to last "
                                   "bit-sync kernel, CUDA events on the launching streams",
                         "algorithmic_bytes_per_step": int(abytes),
                         "k1_demod_ms": round(k1_ms / args.steps, 4), "k2_bitsync_ms": round(k2_ms / args.steps, 4),
                         "device_pass_ms": round(dev_ms / args.steps, 4),
                         # the same algorithmic bytes over the demod kernel alone (the largest kernel)
                         "k1_only": {"achieved": round(abytes / (k1_ms / args.steps / 1e3) / 1e9, 1) if k1_ms else None,
                                     "frac": round(abytes / (k1_ms / args.steps / 1e3) / 1e9 / peak, 4) if k1_ms else None},
                         "ncu": ncu_brief},
            "packets": dict(totals, planted_per_gpu=len(plan)),
            "lanes": {"run": int(st.lanes_run), "rerun": int(st.lanes_rerun), "rl_fallbacks": int(st.rl_fallbacks)},
            "host_ms_per_step": {"batch": round(st.host_batch_ms / (args.steps + args.warmup), 3),
                                 "gather": round(st.host_gather_ms / (args.steps + args.warmup), 3),
                                 "decode": round(st.host_decode_ms / (args.steps + args.warmup), 3)},
        }
        if strong:
            out["strong"] = {k: strong[k] for k in ("value", "unit", "n_gpus", "steps", "ms_per_step", "scaling", "config",
                                                   "e2e", "gpu_launches", "packets")}
        if not args.no_cpu_baseline:
            sample = 64 << 20
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
            path = os.path.join(shm, f"wmbus_cpu_sample_{os.getpid()}.cu8")
            cap[:sample].cpu().numpy().tofile(path)
            try:
                t_cpu = min(time_reference(path, wl["flags"], 1) for _ in range(2))
            finally:
                os.unlink(path)
            _, kind = ref_binary()
            out["cpu_baseline"] = {"value": round(sample / 2 / t_cpu / 1e6, 3), "unit": UNIT, "cores": 1, "kind": kind,
                                   "sample": f"first {sample >> 20} MiB of the same capture, one process, best of 2 "
                                             f"({host_cpus()[0]} usable host CPUs)"}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
