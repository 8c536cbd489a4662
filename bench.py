#!/usr/bin/env python
"""bench.py -- variants/second of the B200-native FamSeq posterior engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[2], the one quoted at 1/2/4/8 GPUs): synthetic trio pedigree,
Elston-Stewart peeling, 10 M variants PER GPU (weak scaling: every rank owns its own contiguous slice of
sites, no data-path collective).  A "step" is one pass of the kernel over the rank's resident batch.
  value    : whole-job variants/s, inputs and outputs resident in HBM, CUDA events, max over ranks
  e2e      : the same batch through the C ABI call fs_run() on HOST (pinned) buffers, H2D + D2H inside
  roofline : HBM, algorithmic bytes = 73*S+2 per variant (221 B for a trio), peak from MEASURED_PEAKS.json
  methods  : BN (3^14 exhaustive enumeration, ped14) and MCMC (ped40 with loops, 1 000 + 10 000 sweeps) timed
             the same way, FP64 roofline against the DFMA peak measured in this run
  cpu_baseline : the reference's own CPU engine (oracle/_ref/ref_harness, built from the unmodified
             reference sources) on all host cores, on a bounded sample of the same workload
`--impl reference` times that CPU engine as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261018
METHOD_ID = {"bn": 1, "es": 2, "mcmc": 3}


def algorithmic_bytes(S: int) -> int:
    """lk 24S + 1 flag byte in; post + single 48S, gt S, status 1 out."""
    return 73 * S + 2


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


# ----------------------------------------------------------------------------------------------------
# clocks during the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml = None
        self.samples = []  # (sm_mhz, max_mhz, reasons bitmask) polled through NVML every ~1 ms
        self._stop = False
        self._max = None
        self.error = None

    def sample_now(self):
        """One synchronous NVML sample (called by the timing loop while the kernels are in flight)."""
        if self.nvml is None:
            return
        import pynvml as nv

        h = self.nvml
        try:
            if self._max is None:
                self._max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception as ex:  # noqa: BLE001
            self.error = repr(ex)
            return
        try:
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:  # noqa: BLE001
            try:
                reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            except Exception as ex:  # noqa: BLE001
                self.error = repr(ex)
                reasons = 0
        self.samples.append((sm, self._max, reasons))

    def _poll(self):
        while not self._stop:
            self.sample_now()
            time.sleep(0.001)

    def __enter__(self):
        try:  # NVML polling catches regions of a few milliseconds that an nvidia-smi loop would miss
            import pynvml as nv

            nv.nvmlInit()
            self.nvml = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=2)
            return
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        if self.nvml is not None and self.samples:
            import pynvml as nv

            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            seen = 0
            for s in self.samples:
                seen |= s[2]
            return {"sm_mhz": int(statistics.median(s[0] for s in self.samples)), "sm_max_mhz": int(self.samples[0][1]),
                    "reasons": sorted(k for k, b in bits.items() if seen & b), "samples": len(self.samples), "source": "nvml"}
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx.append(int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": self.error}
        return {"sm_mhz": int(statistics.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm
# ----------------------------------------------------------------------------------------------------
def cpu_reference_rate(ped, method: str, n_variants_per_core: int, repeat: int, burn: int, rep: int, seed_offset=0):
    """Runs the reference CPU engine on every host core (one single-threaded process per core, disjoint slices,
    as BASELINE.md section 4 prescribes).  Returns (variants/s, cores, kind, sample text)."""
    from famseq_b200 import synth
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    cols = ped.sequenced_cols()
    mid = METHOD_ID[method]
    with tempfile.TemporaryDirectory() as td:
        pp = os.path.join(td, "p.ped")
        ped.write(pp)
        if O.have_ref():
            procs = []
            for c in range(cores):
                lk, fl = synth.synth_likelihoods(ped, n_variants_per_core, SEED + 1, v0=(seed_offset + c) * n_variants_per_core)
                fin = os.path.join(td, f"in{c}.bin")
                with open(fin, "wb") as fh:
                    fh.write(np.array([lk.shape[0], lk.shape[1]], np.int32).tobytes())
                    fh.write(fl.tobytes())
                    fh.write(np.ascontiguousarray(lk).tobytes())
                procs.append([O.REF_HARNESS, f"ped={pp}", f"method={mid}", f"burn={burn}", f"rep={rep}", "seed=1",
                              "cols=" + ",".join(str(x) for x in cols), f"repeat={repeat}", f"in={fin}"])
            running = [subprocess.Popen(a, stdout=subprocess.PIPE, text=True) for a in procs]
            outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in running]
            elapsed = max(o["elapsed_s"] for o in outs)
            total = sum(o["variants"] * o["repeat"] for o in outs)
            kind = "reference"
        else:  # the C restatement, one process per core
            import multiprocessing as mp

            with mp.Pool(cores) as pool:
                res = pool.starmap(_port_worker, [(ped, mid, n_variants_per_core, repeat, burn, rep, seed_offset + c) for c in range(cores)])
            elapsed = max(r[0] for r in res)
            total = sum(r[1] for r in res)
            kind = "port"
    sample = (f"{cores} processes x {n_variants_per_core} variants x {repeat} passes of the same synthetic workload "
              f"(engine only: set_LK + calPostProb + accessors, no file I/O)")
    return total / elapsed, cores, kind, sample


def _port_worker(ped, mid, n, repeat, burn, rep, c):
    from famseq_b200 import synth
    from oracle import oracle as O

    lk, fl = synth.synth_likelihoods(ped, n, SEED + 1, v0=c * n)
    t0 = time.perf_counter()
    for _ in range(repeat):
        O.run(ped, ped.sequenced_cols(), lk, fl, method=mid, burn=burn, rep=rep, rng=O.RNG_LIBC, seed=1)
    return time.perf_counter() - t0, n * repeat


CPU_SAMPLE = {  # (variants per core, passes): ~10-30 s of CPU work per core
    "es": (500_000, 8), "bn": (4, 4), "mcmc": (128, 2),
}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, name, ped, method, variants, burn=0, rep=0):
        self.name, self.ped, self.method, self.variants, self.burn, self.rep = name, ped, method, variants, burn, rep


def time_device_path(torch, dist, fs, eng, wl: Workload, rank, world, steps, warmup, device_index):
    """Kernel-only timing: inputs/outputs resident in HBM, CUDA events on the launch stream, max over ranks."""
    from famseq_b200 import synth

    V, S = wl.variants, len(wl.ped.sequenced_cols())
    lk, fl = synth.synth_likelihoods(wl.ped, V, SEED + METHOD_ID[wl.method], v0=rank * V)
    h_lk = torch.from_numpy(lk).pin_memory()
    h_fl = torch.from_numpy(fl).pin_memory()
    d_lk, d_fl = h_lk.cuda(non_blocking=True), h_fl.cuda(non_blocking=True)
    d_post = torch.empty((V, S, 3), dtype=torch.float64, device="cuda")
    d_single = torch.empty_like(d_post)
    d_gt = torch.empty((V, S), dtype=torch.uint8, device="cuda")
    d_st = torch.empty(V, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    mid = METHOD_ID[wl.method]

    def step():
        eng.run_device(mid, V, d_lk.data_ptr(), d_fl.data_ptr(), d_post.data_ptr(), d_single.data_ptr(), d_gt.data_ptr(),
                       d_st.data_ptr(), burn=wl.burn, rep=wl.rep, seed=SEED, v_offset=rank * V, stream=stream.cuda_stream)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = eng.info()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device_index) as clk:
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        clk.sample_now()  # the launches above are asynchronous: this sample is taken while they run
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.info()["kernel_launches"] - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    failed = int(d_st.sum().item())
    keep = dict(h_lk=h_lk, h_fl=h_fl, d_post=d_post, d_single=d_single, d_gt=d_gt, d_st=d_st)
    return ms / steps, launches, clk.summary(), failed, keep


def time_e2e_path(torch, dist, eng, wl: Workload, rank, world, steps, warmup, keep):
    """The reference-facing call: fs_run() on pinned HOST buffers, H2D and D2H inside the timed region."""
    V, S = wl.variants, len(wl.ped.sequenced_cols())
    h_lk, h_fl = keep["h_lk"], keep["h_fl"]
    h_post = torch.empty((V, S, 3), dtype=torch.float64).pin_memory()
    h_single = torch.empty((V, S, 3), dtype=torch.float64).pin_memory()
    h_gt = torch.empty((V, S), dtype=torch.uint8).pin_memory()
    h_st = torch.empty(V, dtype=torch.uint8).pin_memory()
    mid = METHOD_ID[wl.method]

    def step():
        eng.run_raw(mid, V, h_lk.data_ptr(), h_fl.data_ptr(), h_post.data_ptr(), h_single.data_ptr(), h_gt.data_ptr(),
                    h_st.data_ptr(), burn=wl.burn, rep=wl.rep, seed=SEED, v_offset=rank * V)

    for _ in range(max(1, min(warmup, 2))):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    # the device-resident and the host path must have produced the same bytes
    same = bool(torch.equal(keep["d_post"].cpu(), h_post) and torch.equal(keep["d_gt"].cpu(), h_gt))
    h2d = h_lk.numel() * 8 + h_fl.numel()
    d2h = (h_post.numel() + h_single.numel()) * 8 + h_gt.numel() + h_st.numel()
    return sec / steps, h2d, d2h, same


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variants", type=int, default=10_000_000, help="ES trio variants per GPU")
    ap.add_argument("--bn-variants", type=int, default=1_000_000)
    ap.add_argument("--mcmc-variants", type=int, default=1_000_000)
    ap.add_argument("--es14-variants", type=int, default=1_000_000)
    ap.add_argument("--methods", default="es,es14,bn,mcmc", help="which method lines to time (es is the headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from famseq_b200 import synth

    trio, ped14, ped40 = synth.trio(), synth.ped14(), synth.ped40()
    config = {"workload": "synthetic trio pedigree, ES peeling (-method 2), FP64 likelihood batches",
              "variants_per_gpu": args.variants, "pedigree_members": 3, "sequenced": 3,
              "layout": "lk[V][S][3] f64 + flags[V] u8 -> post,single[V][S][3] f64, gt[V][S] u8, status[V] u8",
              "l2_note": "per-step inputs+outputs (2.2 GB) exceed the 126 MB L2, no flush needed",
              "parallelism": f"variant-sharded x{world}, no collective"}

    # ------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        n, r = CPU_SAMPLE["es"]
        times = []
        for k in range(args.warmup + args.steps):
            v, cores, kind, sample = cpu_reference_rate(trio, "es", n // 4, max(1, r // 4), 0, 0, seed_offset=k)
            if k >= args.warmup:
                times.append(v)
        value = float(statistics.mean(times))
        print(json.dumps({
            "impl": "reference", "metric": "variants/sec (ES trio)", "value": value, "unit": "variants/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (n // 4) * max(1, r // 4) * cores / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": value, "unit": "variants/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    import famseq_b200 as fs

    if not torch.cuda.is_available() or fs.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    hbm_peak, peak_src, _ = measured_peaks()
    fp64_peak = fs.engine.measure_fp64_tflops(local_rank)
    methods = [m.strip() for m in args.methods.split(",") if m.strip()]

    def engine(ped):
        return fs.Engine(ped.ids, ped.mids, ped.fids, ped.genders, ped.sequenced_cols(), device=local_rank)

    out = {}
    # ---- headline: ES trio ---------------------------------------------------------------------------
    wl = Workload("es", trio, "es", args.variants)
    with engine(trio) as eng:
        ms_step, launches, clocks, failed, keep = time_device_path(torch, dist, fs, eng, wl, rank, world, args.steps, args.warmup, local_rank)
        e2e_s, h2d, d2h, same = time_e2e_path(torch, dist, eng, wl, rank, world, max(2, min(args.steps, 5)), args.warmup, keep)
    del keep
    torch.cuda.empty_cache()
    value = world * args.variants / (ms_step * 1e-3)
    achieved = algorithmic_bytes(3) * args.variants / (ms_step * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "es_trio_traffic.json")
    if os.path.exists(prof):
        # dram__bytes_read + dram__bytes_write of one `ncu --set full` capture, scaled to this launch's variant count
        traffic = json.load(open(prof)).get("dram_bytes_per_variant", 0.0) * args.variants or None
    out.update({
        "metric": "variants/sec (ES peeling, trio)", "value": value, "unit": "variants/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_variant": algorithmic_bytes(3),
                     "kernel": "es_nuclear_kernel<1,32> (one warp per block, TMA bulk load/store of a 32-variant tile, register-resident peeling)"},
        "e2e": {"value": world * args.variants / e2e_s, "unit": "variants/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "api": "fs_run() on pinned host buffers", "matches_device_path": same},
        "gpu_launches": launches, "clocks": clocks, "failed_variants": failed,
        "fp64_peak_tflops_measured": fp64_peak,
    })

    # ---- the other two methods ---------------------------------------------------------------------------
    sub = {}
    if "es14" in methods:  # ES on a pedigree that is not a nuclear family (any loop-free pedigree), HBM roofline
        wl = Workload("es14", ped14, "es", args.es14_variants)
        # the message program as generated straight-line code (csrc/cuda/es_jit.cu); compiled in the warm-up step here,
        # on a worker thread after 2e10 variants otherwise (FAMSEQ_ES_JIT=0: the interpreter of es_kernel.cu)
        os.environ.setdefault("FAMSEQ_ES_JIT", "1")
        with engine(ped14) as eng:
            ms_e, l_e, clk_e, failed_e, keep = time_device_path(torch, dist, fs, eng, wl, rank, world, 5, 2, local_rank)
            jit_e = eng.info()["jit_launches"]
        del keep
        torch.cuda.empty_cache()
        ach = algorithmic_bytes(14) * args.es14_variants / (ms_e * 1e-3) / 1e9
        sub["ES_ped14"] = {"workload": "synthetic 14-member 3-generation pedigree, ES peeling (compiled message program)",
                           "kernel": "famseq_es (generated for the pedigree, NVRTC)" if jit_e else "es_kernel (message-program interpreter)",
                           "variants_per_gpu": args.es14_variants, "value": world * args.es14_variants / (ms_e * 1e-3),
                           "unit": "variants/s", "ms_per_step": ms_e, "steps": 5, "warmup": 2, "gpu_launches": l_e, "clocks": clk_e,
                           "failed_variants": failed_e,
                           "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                        "algorithmic_bytes_per_variant": algorithmic_bytes(14)}}
    if "bn" in methods:
        N = 14
        wl = Workload("bn", ped14, "bn", args.bn_variants)
        with engine(ped14) as eng:
            ms_b, l_b, clk_b, failed_b, keep = time_device_path(torch, dist, fs, eng, wl, rank, world, 2, 1, local_rank)
        del keep
        torch.cuda.empty_cache()
        flops = 2.0 * N * 3 ** N  # N-1 multiplies + 1 scale per joint, N adds to scatter it
        ach = flops * args.bn_variants / (ms_b * 1e-3) / 1e12
        sub["BN"] = {"workload": "synthetic 14-member 3-generation pedigree, exhaustive 3^14 enumeration (-method 1)",
                     "variants_per_gpu": args.bn_variants, "value": world * args.bn_variants / (ms_b * 1e-3), "unit": "variants/s",
                     "ms_per_step": ms_b, "steps": 2, "warmup": 1, "gpu_launches": l_b, "clocks": clk_b, "failed_variants": failed_b,
                     "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                                  "algorithmic_flops_per_variant": flops,
                                  "note": "algorithmic flops are the reference's 2*N*3^N; the kernel carries the joint as a prefix "
                                          "product and executes ~2.5 FP64 instructions per configuration, so frac can exceed 1; ncu "
                                          "(profiles/r1q_bn_r1q.txt): FP64 pipe 71.8 % busy, top stall math-pipe throttle"}}
        # Not the contract number: the same kernel with the opt-in closed-form sum over the innermost block of childless
        # members (FAMSEQ_BN_FACTOR=1, bn_kernel.cu: bn_block_factored) -- 3^9 instead of 3^14 configurations visited.
        os.environ["FAMSEQ_BN_FACTOR"] = "1"
        try:
            with engine(ped14) as eng:
                ms_f, l_f, _, failed_f, keep = time_device_path(torch, dist, fs, eng, wl, rank, world, 2, 1, local_rank)
        finally:
            del os.environ["FAMSEQ_BN_FACTOR"]
        del keep
        torch.cuda.empty_cache()
        sub["BN_leaves_summed_analytically"] = {
            "workload": "same as BN; opt-in FAMSEQ_BN_FACTOR=1: the 3^5 configurations of the five innermost (childless) members "
                        "are summed in closed form, 3^9 configurations are enumerated; same marginals to 1e-9",
            "variants_per_gpu": args.bn_variants, "value": world * args.bn_variants / (ms_f * 1e-3), "unit": "variants/s",
            "ms_per_step": ms_f, "steps": 2, "warmup": 1, "gpu_launches": l_f, "failed_variants": failed_f}
    if "mcmc" in methods:
        N, founders, burn, rep = 40, 9, 1000, 10000
        wl = Workload("mcmc", ped40, "mcmc", args.mcmc_variants, burn, rep)
        # The Gibbs kernel of a large batch is generated for the pedigree and compiled at run time (csrc/cuda/gibbs_jit.cu).
        # By default that happens on a worker thread while the table-driven kernel carries on; here the compile is made
        # synchronous so that it falls into the warm-up step and every timed step runs the same kernel.
        os.environ.setdefault("FAMSEQ_MCMC_JIT", "1")
        with engine(ped40) as eng:
            ms_m, l_m, clk_m, failed_m, keep = time_device_path(torch, dist, fs, eng, wl, rank, world, 2, 1, local_rank)
            jit_m = eng.info()["jit_launches"]
        del keep
        torch.cuda.empty_cache()
        flops = (burn + rep) * (14.0 * N + 3.0 * 2 * (N - founders))
        ach = flops * args.mcmc_variants / (ms_m * 1e-3) / 1e12
        sub["MCMC"] = {"workload": "synthetic 40-member pedigree with loops, Gibbs 1000 burn-in + 10000 sweeps (-method 3)",
                       "variants_per_gpu": args.mcmc_variants, "value": world * args.mcmc_variants / (ms_m * 1e-3), "unit": "variants/s",
                       "ms_per_step": ms_m, "steps": 2, "warmup": 1, "gpu_launches": l_m, "clocks": clk_m, "failed_variants": failed_m,
                       "kernel": "famseq_gibbs (generated for the pedigree, NVRTC)" if jit_m else "mcmc_kernel (table-driven)",
                       "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                                    "algorithmic_flops_per_variant": flops,
                                    "note": "a Gibbs step has no FMAs, so the DFMA peak counts every FP64 instruction twice; ncu of the "
                                            "generated kernel (profiles/r1q_mcmc_r1q.txt): FP64 pipe 45 %, shared-memory pipe 75 % busy "
                                            "(transmission-table look-ups and chain state) -- the unit that bounds it"}}
    out["methods"] = sub

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n, r = CPU_SAMPLE["es"]
        v, cores, kind, sample = cpu_reference_rate(trio, "es", n, r, 0, 0)
        out["cpu_baseline"] = {"value": v, "unit": "variants/s", "cores": cores, "kind": kind, "sample": sample}
        if "bn" in methods:
            n, r = CPU_SAMPLE["bn"]
            v, cores, kind, sample = cpu_reference_rate(ped14, "bn", n, r, 0, 0)
            out["methods"]["BN"]["cpu_baseline"] = {"value": v, "unit": "variants/s", "cores": cores, "kind": kind, "sample": sample}
        if "mcmc" in methods:
            n, r = CPU_SAMPLE["mcmc"]
            v, cores, kind, sample = cpu_reference_rate(ped40, "mcmc", n, r, 1000, 10000)
            out["methods"]["MCMC"]["cpu_baseline"] = {"value": v, "unit": "variants/s", "cores": cores, "kind": kind, "sample": sample}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
