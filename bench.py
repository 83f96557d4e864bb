#!/usr/bin/env python
"""bench.py - nonbonded pair-interactions/sec (energy + forces) of the MMFF94 hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ions1m|water3k|solvated25k] [--impl reference]

A "step" is one full energy + gradient evaluation of the workload from fresh coordinates: AoS->SoA
marshalling, (cell build + sort on the cutoff path), nonbonded pair kernel, bonded kernel, final
reduction and - for N > 1 - the collectives of owner mode (every rank owns 1/N of the atoms: all-gather of the
coordinates, all-reduce of the energies, reduce-scatter of the gradient).  Prints ONE JSON line on rank 0
(contract in the task description).

Default workload: BASELINE.json config 5 ("1M-atom LJ+Coulomb box, atom-row sharded at 1/2/4/8 GPUs"), the
configuration the metric "pair-interactions/sec at 1/2/4/8 B200" is quoted on; it fits one GPU, so it is the
N = 1 workload as well (strong scaling: total work fixed as N grows).  --workload selects config 2
(water3k, all-pairs tiles) or config 3 (solvated25k, cutoff 10 A).

`value`   : device-resident throughput (coordinates already in HBM), CUDA events on the launching stream.
`e2e`     : the same metric through the host-buffer C ABI (jme_set_coords + jme_energy_gradient from/to pinned
            host memory), host wall clock between device synchronisations.  For N > 1 every rank copies the
            coordinate rows of its own atoms in and their gradient rows (+ the energies) out.
`roofline`: dominant kernel vs the measured FP64 FMA-pipe peak (this path is FP64-bound, not HBM- or
            tensor-bound - BASELINE.md section 3); `roofline_hbm` gives the same kernel against measured HBM.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 49.0            # energy + gradient
METRIC = "nonbonded pair-interactions/sec (energy+forces)"
UNIT = "pairs/s"

WORKLOADS = {
    # name: (builder kwargs, cutoff, path, description, algorithmic bytes)
    "ions1m": dict(cutoff=10.0, path="cells", desc="config 5: 10^6 Na+/Cl- ions, jittered 100^3 lattice, cutoff 10 A, cell-list path"),
    "water3k": dict(cutoff=0.0, path="allpairs", desc="config 2: 1000 waters (3000 atoms), no cutoff, all-pairs tiles, bonded terms included"),
    "solvated25k": dict(cutoff=10.0, path="cells", desc="config 3: ~25k atoms (chains + water), cutoff 10 A, cell-list path, exclusions + 1-4"),
    "opt5k": dict(cutoff=0.0, path="allpairs", desc="config 4: the reference's GradientDescentMinimizer loop (jme_gd_minimize) on 1667 waters "
                                                   "(5001 atoms): one coordinate move + one full ENERGY evaluation per call"),
}


def build_system(name: str, scale: float = 1.0):
    from jmolecularenergy_b200 import systems
    if name == "ions1m":
        return systems.ion_box(max(4, int(round(100 * scale ** (1.0 / 3.0)))))
    if name == "water3k":
        return systems.water_box(max(1, int(1000 * scale)))
    if name == "solvated25k":
        return systems.solvated_chains(max(300, int(24999 * scale)), exact_cutoff_pairs=100, cutoff=10.0)
    if name == "opt5k":
        return systems.water_box(max(1, int(1667 * scale)))
    raise SystemExit(f"unknown workload {name}")


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the
    CPU baseline: the thread count is passed to the oracle explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_baseline(workload: str, faithful: bool, budget_s: float = 20.0):
    """The oracle port timed on the host cores on a bounded sample of the workload (about 10-30 s of CPU
    work): OpenMP over all threads, type-pair table (`faithful`=False) or the reference's per-pair
    recomputation of R*, eps (`faithful`=True)."""
    from oracle import oracle
    threads = host_threads()
    cfg = WORKLOADS[workload]
    # size the sample from a quick probe
    scale = {"ions1m": 0.05, "water3k": 1.0, "solvated25k": 1.0}[workload]
    s = build_system(workload, scale)
    t0 = time.perf_counter()
    r = oracle.fast_evaluate(s, cutoff=cfg["cutoff"], gradient=True, threads=threads, faithful_constants=faithful)
    dt = time.perf_counter() - t0
    rate = r["pairs"] / dt
    if workload == "ions1m" and dt < budget_s / 8:
        scale = min(1.0, scale * min(8.0, (budget_s / 2) / max(dt, 1e-3)))
        s = build_system(workload, scale)
        t0 = time.perf_counter()
        r = oracle.fast_evaluate(s, cutoff=cfg["cutoff"], gradient=True, threads=threads, faithful_constants=faithful)
        dt = time.perf_counter() - t0
        rate = r["pairs"] / dt
    # repeat the evaluation of the sample until about 10 s of wall time (= 10 s x cores of CPU work) are spent
    reps = max(1, min(40, int(10.0 / max(dt, 1e-3))))
    if reps > 1:
        t0 = time.perf_counter()
        for _ in range(reps):
            r = oracle.fast_evaluate(s, cutoff=cfg["cutoff"], gradient=True, threads=threads, faithful_constants=faithful)
        dt = (time.perf_counter() - t0) / reps
        rate = r["pairs"] / dt
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{s.name}: {r['pairs']} pairs, energy+gradient, {reps} x {dt:.2f} s, "
                      f"{'per-pair R*/eps recomputation like the reference' if faithful else 'type-pair table'}; "
                      "the Java reference itself cannot run (no JVM in the image)"}, s, r


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, per-pair constants like the Java code)
    on all host threads; each step is one evaluation of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    cfg = WORKLOADS[args.workload]
    threads = host_threads()
    # the FULL workload of the GPU arm (same atoms, same pairs per step): about 1.4 s per step for ions1m on 16 cores
    s = build_system(args.workload, 1.0)
    times, pairs = [], 0
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = oracle.fast_evaluate(s, cutoff=cfg["cutoff"], gradient=True, threads=threads, faithful_constants=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            pairs = r["pairs"]
    total = sum(times)
    value = pairs * len(times) / total
    sample = f"{s.name}: the full workload, {pairs} pairs per step, OpenMP x{threads}, per-pair R*/eps like MMFF94VdwComponent.java:106-135"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "atoms": s.n_atoms, "pairs_per_step": pairs,
                       "cutoff": cfg["cutoff"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_opt5k(args):
    """BASELINE config 4: energy calls per second inside the reference's optimizer loop.  One "step" here is one
    call of the loop body = jme_set_coord1 + one full energy evaluation (+ the occasional undo move); the loop
    runs inside the library (jme_gd_minimize), results come back through mapped pinned memory, so the number
    is end to end by construction (host-driven, one device synchronisation per call)."""
    import torch
    from jmolecularenergy_b200 import MMFF94, NativeGradientDescentMinimizer
    torch.cuda.set_device(0)
    s = build_system("opt5k")
    ff = MMFF94(False)
    ff.assignParameters(s)
    pairs = ff.calculateComponents(s)["pairs"]
    sampler = ClockSampler(0)
    m = NativeGradientDescentMinimizer(ff)
    launches0 = ff.launchCount(s)
    sampler.start()
    t0 = time.perf_counter()
    m.minimize(s, 1, 0.01, 1e-4)          # initialising sweep + one step: 1 + 2 * 3N energy calls
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    calls = m.energyCalls
    # the same loop with every trial move answered by jme_energy_delta (the moved atom's terms only, O(N) per call):
    # reported beside the headline, which stays the full evaluation per call that the reference's loop makes
    s2 = build_system("opt5k")
    ff2 = MMFF94(False)
    ff2.assignParameters(s2)
    m2 = NativeGradientDescentMinimizer(ff2, incremental=True)
    t0 = time.perf_counter()
    m2.minimize(s2, 1, 0.01, 1e-4)
    dt2 = time.perf_counter() - t0
    incremental = {"energy_calls": m2.energyCalls, "energy_calls_per_s": m2.energyCalls / dt2, "us_per_call": 1e6 * dt2 / m2.energyCalls,
                   "optimizer_steps": m2.step, "energy_first_last": [m2.stepEnergies[0], m2.stepEnergies[-1]],
                   "final_energy_full_evaluation": ff2.calculateEnergy(s2),
                   "what": "jme_set_incremental(1): each trial = jme_energy_delta (N - 1 pairs + the atom's bonded terms, old and new position)"}
    ff2.release(s2)
    s3 = build_system("opt5k")
    ff3 = MMFF94(False)
    ff3.assignParameters(s3)
    m3 = NativeGradientDescentMinimizer(ff3, incremental="device")
    t0 = time.perf_counter()
    m3.minimize(s3, 1, 0.01, 1e-4)
    dt3 = time.perf_counter() - t0
    incremental["device_loop"] = {"energy_calls": m3.energyCalls, "energy_calls_per_s": m3.energyCalls / dt3,
                                  "us_per_call": 1e6 * dt3 / m3.energyCalls, "optimizer_steps": m3.step,
                                  "energy_first_last": [m3.stepEnergies[0], m3.stepEnergies[-1]],
                                  "final_energy_full_evaluation": ff3.calculateEnergy(s3),
                                  "what": "jme_set_incremental(2): the whole loop in one kernel (minimize_kernel), one cluster of 8 thread blocks; the time includes the initial full evaluation, the launch and the copy back"}
    ff3.release(s3)
    line = {"metric": METRIC.replace("energy+forces", "energy only, optimizer loop"), "value": calls * pairs / dt, "unit": UNIT,
            "n_gpus": 1, "steps": calls, "warmup": 0, "ms_per_step": 1e3 * dt / calls, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "opt5k", "description": WORKLOADS["opt5k"]["desc"], "atoms": s.n_atoms, "pairs_per_step": pairs,
                       "energy_calls": calls, "energy_calls_per_s": calls / dt, "optimizer_steps": m.step,
                       "energy_first_last": [m.stepEnergies[0], m.stepEnergies[-1]], "incremental": incremental},
            "e2e": {"value": calls * pairs / dt, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 80},
            "gpu_launches": ff.launchCount(s) - launches0, "clocks": clocks}
    if not args.no_cpu_baseline:
        from oracle import oracle
        t0 = time.perf_counter()
        r = oracle.fast_evaluate(s, cutoff=0.0, gradient=False, threads=host_threads())
        cdt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": r["pairs"] / cdt, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                "sample": f"one full energy of {s.name} ({r['pairs']} pairs) in {cdt:.3f} s = {1 / cdt:.1f} calls/s; "
                                          "the reference needs 3N such calls per optimizer step"}
    emit(line)


_JSON_OUT = None


def _claim_stdout():
    """Keep the real stdout for the ONE JSON line: file descriptor 1 is pointed at stderr for everything else
    (NCCL prints its version banner on stdout from C code, whatever NCCL_DEBUG says)."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="ions1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: coordinate / gradient exchange over NVLink peer memory (jme_peer_*) or NCCL collectives")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "opt5k":
        if int(os.environ.get("RANK", "0")) == 0:
            run_opt5k(args)
        return

    import torch
    import torch.distributed as dist
    from jmolecularenergy_b200 import _lib, measure_fp64_peak
    from jmolecularenergy_b200.distributed import ShardedMMFF94, unpack

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"

    # everything device-resident runs on one explicit (non-default) stream: CUDA events bracket it there, NCCL
    # uses it, and the library can replay its CUDA graph on it (the legacy default stream cannot be captured)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    cfg = WORKLOADS[args.workload]
    system = build_system(args.workload)
    n = system.n_atoms
    L = _lib.lib()
    ev = ShardedMMFF94(system, cutoff=cfg["cutoff"], path=cfg["path"], rank=rank, world=world)
    st = ev.state
    xyz_dev = torch.from_numpy(system.xyz).to(dev)
    out = ev.new_output(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    # N > 1: owner mode - a rank supplies the coordinates of the atoms it owns and receives their gradient rows
    # (all-gather of the coordinates, all-reduce of the 10 energies / counts, reduce-scatter of the gradient)
    a0, a1 = ev.owned_range()
    xyz_full, out_full, grad_own = ev.new_owner_buffers()
    xyz_own_dev = torch.zeros(3 * ev.chunk, dtype=torch.float64, device=dev)
    xyz_own_dev[:3 * (a1 - a0)] = xyz_dev[a0:a1].reshape(-1)

    def step_device():
        if world == 1:
            ev.set_coords(xyz_dev)            # AoS -> SoA on the device; invalidates the cell sort
            ev.evaluate_into(out, True)       # kernels on the current stream
        else:
            ev.step_owned(xyz_own_dev, out_full, grad_own, xyz_full)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the host entry points size the cutoff path's lists to the system (a device-side evaluation cannot: an overflow
    # only poisons its total); one host-API energy before the device-resident loop does that on every rank
    if world == 1:
        ev.set_coords(xyz_dev)
    else:
        ev.set_coords_owned(xyz_own_dev, xyz_full)
    ev.prime()
    exchange = "single GPU"
    if world > 1:
        # from here on step_owned goes over NVLink peer memory, not through NCCL (unless the GPUs cannot map each other)
        exchange = "peer" if args.exchange == "peer" and ev.open_peer_exchange() else "nccl"
    for _ in range(args.warmup):
        step_device()
    barrier()
    res = unpack((out if world == 1 else out_full).cpu().numpy(), n, False)
    pairs = res["pairs"]

    # ---- N > 1: the sharded result against an unsharded evaluation of the same system on this rank's GPU
    #      (energies and pair count on every rank, the gradient rows each rank owns)
    parity = None
    if world > 1:
        ev1 = ShardedMMFF94(system, cutoff=cfg["cutoff"], path=cfg["path"], rank=0, world=1)
        out1 = ev1.new_output(True)
        ev1.set_coords(xyz_dev)
        ev1.prime()
        ev1.evaluate_into(out1, True)
        torch.cuda.synchronize()
        ref1 = out1.cpu().numpy()

        def compare():
            got_head = out_full[:10].cpu().numpy()
            e_err = float(np.max(np.abs(got_head[:8] - ref1[:8]) / np.maximum(np.abs(ref1[:8]), 1e-3)))
            g_ref = ref1[10:10 + 3 * n].reshape(n, 3)[a0:a1]
            g_got = grad_own.cpu().numpy()[:3 * (a1 - a0)].reshape(-1, 3)
            floor = 1e-10 * max(1.0, float(np.abs(ref1[10:]).max()))
            g_err = float(np.max(np.abs(g_got - g_ref) / (np.abs(g_ref) + floor / 1e-8))) if a1 > a0 else 0.0
            ok = e_err <= 1e-10 and g_err <= 1e-8 and int(round(got_head[8])) == int(round(ref1[8]))    # (NaN compares false)
            t = torch.tensor([e_err if e_err == e_err else 1e300, g_err if g_err == g_err else 1e300, 0.0 if ok else 1.0],
                             dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return {"parity_checked": True, "ok": bool(t[2].item() == 0.0), "energy_rel_err_max": float(t[0].item()),
                    "gradient_rel_err_max": float(t[1].item()), "pairs_equal": int(round(got_head[8])) == int(round(ref1[8])),
                    "against": "an unsharded (world = 1) evaluation of the same coordinates on every rank's GPU; tolerances 1e-10 "
                               "(energies) and 1e-8 relative with a 1e-10 x max floor (owned gradient rows)"}

        parity = compare()
        if not parity["ok"] and exchange == "peer":
            # the exchange over peer memory did not reproduce the unsharded result on some rank (a wait that timed out
            # poisons the total): fall back to the collectives, on every rank, and check again
            ev.close_peer_exchange()
            exchange = "nccl"
            ev.peer_error = f"peer exchange failed its parity check ({parity})"
            for _ in range(args.warmup):
                step_device()
            barrier()
            parity = compare()
            pairs = unpack(out_full.cpu().numpy(), n, False)["pairs"]
        ev1.close()
        del out1
        torch.cuda.empty_cache()
        if not parity["ok"]:
            raise SystemExit(f"sharded evaluation differs from the unsharded one: {parity}")

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream, L2 flushed between
    sampler = ClockSampler(local_rank)
    launches0 = ev.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        starts[k].record()
        step_device()
        stops[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = ev.launch_count() - launches0
    ms = [starts[k].elapsed_time(stops[k]) for k in range(args.steps)]
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    # ---- the dominant kernel's own duration: the same step again with the library's CUDA events switched on
    #      (events on the launching stream around the nonbonded kernel; this disables the CUDA-graph replay, so it
    #      is kept out of the timed region above)
    n_prof = min(args.steps, 50)
    L.jme_set_profiling(st.h, 1)
    for k in range(n_prof):
        flush.fill_(k & 0xff)
        step_device()
    barrier()
    nb = (C.c_float * n_prof)()
    bl = (C.c_float * n_prof)()
    got = C.c_int()
    L.jme_profile_read(st.h, nb, bl, n_prof, C.byref(got))
    L.jme_set_profiling(st.h, 0)
    nb_ms = float(np.mean(nb[:got.value])) if got.value else float("nan")
    build_ms = float(np.mean(bl[:got.value])) if got.value else float("nan")

    # ---- N > 1: where a step goes - CUDA events around the phases of an owner-mode step, max over ranks per phase
    phases = None
    if world > 1:
        names = ["all_gather_coordinates", "local_evaluation", "all_reduce_energies", "reduce_scatter_gradient"]
        acc = [0.0] * 4
        n_ph = min(args.steps, 30)
        for k in range(n_ph):
            flush.fill_(k & 0xff)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            e[0].record()
            dist.all_gather_into_tensor(xyz_full, xyz_own_dev)
            e[1].record()
            ev.set_coords(xyz_full[:3 * n].view(n, 3))
            st.check(L.jme_eval_device(st.h, 1, C.c_void_p(out_full.data_ptr())))
            e[2].record()
            dist.all_reduce(out_full[:10])
            e[3].record()
            dist.reduce_scatter_tensor(grad_own, out_full[10:])
            e[4].record()
            torch.cuda.synchronize()
            for i in range(4):
                acc[i] += e[i].elapsed_time(e[i + 1])
        t = torch.tensor(acc, dtype=torch.float64, device=dev) / n_ph
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = {k: float(v) for k, v in zip(names, t.tolist())}
        phases["unit"] = "ms, max over ranks; the phases of the step with NCCL collectives, measured in a separate loop after the timed region (the timed step itself uses the exchange named in config.parallelism)"

    # ---- end to end through the host-buffer C ABI (rank-local; for N > 1 the partial results are reduced
    #      on the device and copied back, so the host round trip is included on every rank)
    host_xyz = torch.from_numpy(system.xyz.copy()).pin_memory()
    host_out = torch.empty(ev.n_out_grad, dtype=torch.float64).pin_memory()
    host_own_xyz = torch.zeros(3 * ev.chunk, dtype=torch.float64).pin_memory()
    host_own_xyz[:3 * (a1 - a0)] = torch.from_numpy(system.xyz[a0:a1].reshape(-1))
    host_own_out = torch.empty(10 + 3 * ev.chunk, dtype=torch.float64).pin_memory()
    total_e = C.c_double()
    comps = (C.c_double * 7)()

    def step_e2e():
        if world == 1:
            st.check(L.jme_set_coords(st.h, host_xyz.data_ptr(), n))
            st.check(L.jme_energy_gradient(st.h, C.byref(total_e), comps, host_out.data_ptr() + 80))
        else:
            # pinned host rows of the owned atoms -> device, all-gather over NVLink, evaluation, reduction,
            # energies + owned gradient rows -> pinned host
            xyz_own_dev.copy_(host_own_xyz, non_blocking=True)
            ev.step_owned(xyz_own_dev, out_full, grad_own, xyz_full)
            host_own_out[:10].copy_(out_full[:10], non_blocking=True)
            host_own_out[10:].copy_(grad_own, non_blocking=True)
            torch.cuda.synchronize()

    e2e_steps = max(5, min(args.steps, 30))
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        fp64_tflops, fp64_mhz = measure_fp64_peak()
        ms_per_step = total_ms / args.steps
        value = pairs / (ms_per_step * 1e-3)
        # dominant kernel: the nonbonded pair kernel of this rank's shard
        local_pairs = pairs / world
        achieved_tf = FLOP_PER_PAIR * local_pairs / (nb_ms * 1e-3) / 1e12
        if cfg["path"] == "cells":
            alg_bytes = 4.0 * local_pairs + 84.0 * n / world          # 4 B/pair + 84 B/atom
        else:
            alg_bytes = 36.0 / 128.0 * local_pairs
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get(args.workload)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "atoms": n, "pairs_per_step": pairs,
                       "cutoff": cfg["cutoff"], "path": cfg["path"], "parallelism": (f"atom-row shards x{world}; " + ("coordinate slices pushed to / gradient rows pulled from every rank over NVLink peer memory (jme_peer_step)" if exchange == "peer" else "all-gather of coordinates, all-reduce of energies, reduce-scatter of the gradient (NCCL)" + (f"; peer mapping failed: {ev.peer_error}" if ev.peer_error else "")) +
                                       " - every rank ends with the rows of its own atoms")
                       if world > 1 else "single GPU",
                       "scheme": {1: "all-pairs tiles", 3: "cell lists, per-atom survivor lists", 4: "cell lists, cluster lists + Newton's third law"}.get(int(L.jme_active_path(st.h)), "?"),
                       "l2": "256 MiB written between timed steps (L2 flush); step = marshal + cell build + pairs + bonded + reduce"},
            "e2e": {"value": pairs / (e2e_s / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": 24 * ev.chunk * world,
                    "d2h_bytes_per_step": (24 * ev.chunk + 80) * world, "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_tflops, "traffic": traffic,
                         "kernel": {1: "ap_kernel<GRAD>", 3: "cells_pair_kernel<GRAD> (per-atom lists)",
                                    4: "cl_pair_kernel<4,GRAD> (cluster lists, Newton's third law)"}.get(int(L.jme_active_path(st.h)), "?"),
                         "kernel_ms": nb_ms, "build_ms": build_ms,
                         "peak_source": f"DFMA chain microbenchmark measured in this run ({fp64_mhz:.0f} MHz equivalent at 64 FMA/clk/SM); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "flop_per_pair": FLOP_PER_PAIR},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (nb_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": alg_bytes / (nb_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs"),
                             "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"},
            "wall_s_timed_region": t_wall,
            "energy_kcal_mol": res["total"],
        }
        if parity is not None:
            line["parity"] = parity
            line["parity_checked"] = True
        if phases is not None:
            line["phases_ms"] = phases
        if not args.no_cpu_baseline:
            base, _, _ = cpu_baseline(args.workload, faithful=False)
            line["cpu_baseline"] = base
        emit(line)
    ev.close()                       # (collective while the peer exchange is open: before the process group goes)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
