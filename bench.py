#!/usr/bin/env python
"""Benchmark of the CP-PPD hot path (BASELINE.json metric: iterations/s + effective HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (N = 1 and N > 1 alike): the synthetic 4096x4096 Potts-segmentation LP of
BASELINE.json configs[4] (n = 50 323 456, m = 67 092 480, nnz = 201 277 440), fp64.
A *step* is ``--iters-per-step`` (default 50) solver iterations — one pass of the hot path
(A^T y + primal update, A xbar + dual update) over the whole LP per iteration.

* ``value``   : iterations/s with the LP resident in HBM, CUDA-event timed on the solver's
                stream, K steps bracketed by barrier + synchronize, max over ranks.
* ``e2e``     : the same metric through the public call ``chambolle_pock_ppd(host arrays...)``:
                every e2e step uploads the whole LP from pinned host memory, builds the operator,
                runs ``--e2e-iters`` iterations (stats block at iteration 0, as the reference
                does) and reads x back to the host.
* ``with_stats_block``: iterations/s of the same resident loop with the reference's stats block every
                ``--stats-interval`` (500) iterations.
* ``latency_bound_configs``: iterations/s of BASELINE.json's two small configs (Potts 50x50, netlib SC105), which fit
                the caches and are bound by launch latency, not bandwidth (extra information, N = 1).
* ``roofline``: algorithmic bytes of the dominant kernel / its CUDA-event time,
                against MEASURED_PEAKS.json's hbm_gbs (fallback 6650 GB/s).
* ``cpu_baseline`` / ``--impl reference``: the oracle port of the reference's CPU path
                (oracle/, numpy+scipy single-thread and plain-C OpenMP) on the host cores.

N > 1 is launched by torchrun (one rank per GPU); total work is fixed ("strong" scaling).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cp_ppd_iterations_per_second"
UNIT = "iterations/s"


def algorithmic_bytes(n, m, nnz):
    """per-iteration bytes, split per kernel (A pass = dual, A^T pass = primal)."""
    p = 8 if nnz >= 2**31 else 4
    dual = nnz * 12 + p * (m + 1) + 8 * n + 24 * m + 8 * m
    primal = nnz * 12 + p * (n + 1) + 8 * m + 40 * n + 16 * n
    return primal, dual


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while the timed region runs."""

    BAD = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")

    def __init__(self, index):
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

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksEventReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pinned_empty():
    """numpy allocator backed by pinned host memory (torch is only the allocator)."""
    import torch

    keep = []
    table = {np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}

    def empty(count, dtype=np.float64):
        t = torch.empty(int(count), dtype=table[np.dtype(dtype)], pin_memory=True)
        keep.append(t)
        return t.numpy()

    return empty, keep


def build_workload(size, pinned):
    from pysparselp_b200 import generators

    if pinned:
        empty, keep = pinned_empty()
        lp = generators.potts_lp(size, empty=empty)
        return lp, keep
    return generators.potts_lp(size), None


def lp_nbytes(lp):
    total = 0
    for v in lp:
        if v is None:
            continue
        if hasattr(v, "indptr"):
            total += v.indptr.nbytes + v.indices.nbytes + v.data.nbytes
        else:
            total += np.asarray(v).nbytes
    return total


# ------------------------------------------------------------------------------------------
def cpu_baselines(lp, sample_note, numpy_iters=2, c_iters=6):
    """Oracle port timed on the host cores: numpy/scipy (1 thread) and plain C + OpenMP (all cores)."""
    from oracle.c_port import COracle
    from oracle.cpppd_oracle import CpPpdOracle
    from pysparselp_b200 import generators

    args = generators.lp_args(lp)
    out = {}
    o = CpPpdOracle(*args)
    o.primal_step()
    o.dual_step()
    t0 = time.perf_counter()
    for _ in range(numpy_iters):
        o.primal_step()
        o.dual_step()
    out["numpy"] = numpy_iters / (time.perf_counter() - t0)
    del o
    from oracle import c_port

    cores = c_port.set_threads()  # every core of the box, whatever OMP_NUM_THREADS the launcher exported
    co = COracle(*args)
    co.iterate(1)
    t0 = time.perf_counter()
    co.iterate(c_iters)
    out["c_openmp"] = c_iters / (time.perf_counter() - t0)
    return {
        "value": out["c_openmp"], "unit": UNIT, "cores": cores, "kind": "port",
        "sample": "%s; oracle/cpppd_oracle.c (OpenMP, %d threads) %d iterations after 1 warm-up; "
                  "numpy+scipy restatement (the reference's own arithmetic, 1 thread) ran at %.3f it/s over %d iterations"
                  % (sample_note, cores, c_iters, out["numpy"], numpy_iters),
        "numpy_scipy_single_thread_value": out["numpy"],
    }


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_port
    from oracle.c_port import COracle
    from pysparselp_b200 import generators

    # torchrun exports OMP_NUM_THREADS=1 to its ranks: ask for every core explicitly and report what OpenMP uses
    cores = c_port.set_threads()
    # bounded sample: full size when the requested number of steps allows it, else a smaller image
    size = a.size
    budget_iters = (a.steps + a.warmup) * a.ref_iters_per_step
    while size > 512 and budget_iters * 0.45 * (size / 4096.0) ** 2 > 150:  # ~0.45 s/iteration at 4096^2 on 8 cores
        size //= 2
    lp, _ = build_workload(size, pinned=False)
    n, m, nnz = lp.c.size, lp.a_ineq.shape[0], lp.a_ineq.nnz
    scale = nnz / float(_potts_nnz(a.size))
    co = COracle(*generators.lp_args(lp))
    for _ in range(a.warmup):
        co.iterate(a.ref_iters_per_step)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        co.iterate(a.ref_iters_per_step)
    dt = time.perf_counter() - t0
    its = a.steps * a.ref_iters_per_step / dt * scale
    del co
    # beside it: the reference's own arithmetic (scipy csr_matvec / csc_matvec + numpy ufuncs, one thread by
    # construction) on the same sample — what a PySparseLP user runs today
    numpy_its = None
    if a.ref_numpy_iters > 0:
        from oracle.cpppd_oracle import CpPpdOracle

        o = CpPpdOracle(*generators.lp_args(lp))
        o.primal_step()
        o.dual_step()
        t0 = time.perf_counter()
        for _ in range(a.ref_numpy_iters):
            o.primal_step()
            o.dual_step()
        numpy_its = a.ref_numpy_iters / (time.perf_counter() - t0) * scale
        del o
    sample = ("each step = %d iteration(s) of the %dx%d Potts LP (nnz %d) on the CPU, value scaled by nnz ratio %.4f "
              "to the %dx%d workload; plain-C OpenMP oracle port of pysparselp/ChambollePockPPD.py:195-343"
              % (a.ref_iters_per_step, size, size, nnz, scale, a.size, a.size))
    line = {
        "impl": "reference", "metric": METRIC, "value": its, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": its, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "numpy_scipy_single_thread_value": numpy_its},
        "e2e": {"value": its, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _potts_nnz(size):
    return 3 * 2 * (size * (size - 1) * 2)


def workload_config(a):
    return {"workload": "potts_segmentation_lp_%dx%d" % (a.size, a.size),
            "generator": "pysparselp_b200.generators.potts_lp(seed=1, coef_potts=0.5, coef_mul=500)",
            "iters_per_step": a.iters_per_step, "theta": 1, "alpha": 1,
            "l2": "inputs larger than L2 (about 11 GB streamed per iteration vs 126 MB L2); no flush needed",
            "flags": a.flags,
            "parallelism": "1 GPU" if a.gpus == 1 else "%d GPUs, owner-computes row/column strips, %s halo exchange" % (
                a.gpus, "NCCL send/recv" if a.flags & 32 else (
                    "peer-memory (NVLink) stores fused into k_primal / k_dual" if a.flags & 128
                    else "peer-memory (NVLink) push-kernel"))}


def time_variant(make_solver, args, flags, iters_per_step, peak):
    """iterations/s of the same solve with opt-in storage flags (iterates stay bit-identical)."""
    vs = make_solver(*args, flags=flags)
    try:
        vs.iterate(2 * iters_per_step)
        vs.sync()
        ms = vs.time_iterations(4 * iters_per_step) / (4 * iters_per_step)
        info = vs.info()
    finally:
        vs.close()
    return {"flags": flags, "iterations_per_s": 1e3 / ms,
            "actual_bytes_per_iteration": info["bytes_per_iteration_actual"],
            "actual_GBs": info["bytes_per_iteration_actual"] / ms / 1e6,
            "actual_frac_of_peak": info["bytes_per_iteration_actual"] / ms / 1e6 / peak,
            "algorithmic_GBs": info["bytes_per_iteration_algorithmic"] / ms / 1e6,
            "value_bytes": info["value_bytes"], "const_vector_mask": info["const_vector_mask"]}


# ------------------------------------------------------------------------------------------
def run_b200(a):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (a.gpus, a.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pysparselp_b200 import generators
    from pysparselp_b200.ChambollePockPPD import chambolle_pock_ppd, make_solver

    def barrier():
        if dist is not None:
            dist.barrier()

    t_build = time.perf_counter()
    lp, keep = build_workload(a.size, pinned=True)
    t_build = time.perf_counter() - t_build
    args = generators.lp_args(lp)
    n, m, nnz = lp.c.size, lp.a_ineq.shape[0], lp.a_ineq.nnz

    # ---- device-resident timing -------------------------------------------------------------
    t_setup = time.perf_counter()
    solver = make_solver(*args, flags=a.flags)
    t_setup = time.perf_counter() - t_setup
    info = solver.info()
    for _ in range(a.warmup):
        solver.iterate(a.iters_per_step)
    solver.sync()
    barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        ms = solver.time_iterations(a.steps * a.iters_per_step)
        torch.cuda.synchronize()
    barrier()
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    its_per_s = a.steps * a.iters_per_step / (ms * 1e-3)

    # ---- per-kernel times (events between kernels) -> roofline of the dominant kernel --------
    peak, peak_src = measured_peak_gbs()
    kp, kd = solver.time_kernels(32)
    kp, kd = kp / 32, kd / 32
    bp, bd = algorithmic_bytes(n, m, nnz)
    if world > 1:  # per-rank share of the algorithmic bytes
        bp, bd = bp / world, bd / world
    dom = ("k_primal", kp, bp) if kp >= kd else ("k_dual", kd, bd)
    roofline = {
        "bound": "hbm", "kernel": dom[0], "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": load_ncu_traffic(dom[0]),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[2], "avg_launch_ms": dom[1],
        "traffic_source": "profiles/ncu_summary.json (ncu --set full of variant 1, round-1 v2 capture)",
        "kernels": {"k_primal": {"ms": kp, "algorithmic_GBs": bp / (kp * 1e-3) / 1e9},
                    "k_dual": {"ms": kd, "algorithmic_GBs": bd / (kd * 1e-3) / 1e9}},
        "iteration": {"algorithmic_bytes": info["bytes_per_iteration_algorithmic"],
                      "effective_GBs": info["bytes_per_iteration_algorithmic"] * its_per_s / 1e9,
                      "frac_of_peak": info["bytes_per_iteration_algorithmic"] * its_per_s / 1e9 / (peak * world),
                      "note": "whole job: algorithmic bytes of the full LP x iterations/s, against n_gpus x peak"},
    }
    # ---- the same loop with the reference's stats block every 500 iterations (its usual nb_iter_plot): SURVEY 8(d)
    with_stats = None
    try:
        interval, blocks = a.stats_interval, 2
        solver.sync()
        barrier()
        t0 = time.perf_counter()
        for _ in range(blocks):
            solver.primal_step(keep_d=True)
            solver.stats_step(False)
            solver.read_stats()  # the one host synchronisation of the interval
            solver.dual_step()
            solver.iterate(interval - 1)
        solver.sync()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        with_stats = {"iterations_per_s": blocks * interval / dt, "nb_iter_plot": interval, "iterations": blocks * interval,
                      "note": "device-resident, wall clock around %d stats intervals (stats block + 96-byte read-back each)" % blocks}
    except Exception as e:
        if world > 1:
            raise
        with_stats = {"error": repr(e)}
    solver.close()
    del solver

    # ---- opt-in variants of the same solve (bit-identical iterates, different storage), N = 1 only
    variants = None
    if world == 1 and a.variants:
        variants = {}
        for name, vflags in (("reorder", 8), ("compressed", 3), ("compressed+reorder", 11)):
            try:
                variants[name] = time_variant(make_solver, args, vflags, a.iters_per_step, peak)
            except Exception as e:  # a variant must never take the headline measurement down
                variants[name] = {"flags": vflags, "error": repr(e)}

    # ---- the two CPU-runnable configs of BASELINE.json (configs[0] Potts 50x50, configs[1] netlib SC105): launch-latency
    #      bound ("report it/s only"); CUDA graphs of 50 iterations vs the opt-in persistent CTA
    small = None
    if world == 1 and a.small_configs:
        small = {}
        try:
            small_lps = {"potts_50x50": generators.lp_args(generators.potts_lp(50))}
            try:
                from pysparselp_b200.netlib import get_problem
                from pysparselp_b200.SparseLP import SparseLP

                d = get_problem("SC105")
                lp105 = SparseLP()
                lp105.add_variables_array(len(d["cost_vector"]), lower_bounds=d["lower_bounds"],
                                          upper_bounds=np.minimum(d["upper_bounds"], np.max(d["solution"]) * 2),
                                          costs=d["cost_vector"])
                lp105.add_equality_constraints_sparse(d["a_eq"], d["b_eq"])
                lp105.add_inequality_constraints_sparse(d["a_ineq"], d["b_lower"], d["b_upper"])
                lp105.convert_to_one_sided_inequality_system()
                small_lps["netlib_sc105"] = (lp105.costsvector, lp105.a_equalities, lp105.b_equalities,
                                             lp105.a_inequalities, lp105.b_lower, lp105.b_upper, lp105.lower_bounds,
                                             lp105.upper_bounds)
            except Exception as e:
                small["netlib_sc105"] = {"error": repr(e)}
            for name, sargs in small_lps.items():
                small[name] = {}
                # (the persistent CTA has not run on hardware yet: only timed on request, --small-configs 2)
                for label, sflags in (("cuda_graphs", 0), ("persistent_cta", 512))[: 1 if a.small_configs < 2 else 2]:
                    try:
                        ss = make_solver(*sargs, flags=sflags)
                        try:
                            ss.iterate(a.small_iters)
                            ss.sync()
                            ms_small = ss.time_iterations(a.small_iters)
                            used = bool(ss.info()["tiny_persistent"])
                        finally:
                            ss.close()
                        if label == "persistent_cta" and not used:
                            small[name][label] = None  # the LP does not fit one CTA: the flag is ignored
                        else:
                            small[name][label] = {"iterations_per_s": a.small_iters / (ms_small * 1e-3)}
                    except Exception as e:
                        small[name][label] = {"error": repr(e)}
        except Exception as e:
            small = {"error": repr(e)}

    # ---- end to end through the public API with host buffers -----------------------------------
    h2d = lp_nbytes(lp)
    d2h = 8 * n + 96
    e2e_value, e2e_error = None, None
    if a.e2e_steps > 0:
        def one_call():
            x, best = chambolle_pock_ppd(*args, nb_max_iter=a.e2e_iters, nb_iter_plot=a.e2e_iters, flags=a.flags)
            return x

        try:  # (a failure here must not take the device-timed headline down with it; it is reported in the line)
            one_call()  # warm-up (allocator pools, graph instantiation, kernel-variant timing of this operand shape)
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                x = one_call()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if not np.all(np.isfinite(x)):
                raise FloatingPointError("end-to-end x holds non-finite entries")
            e2e_value = a.e2e_steps * a.e2e_iters / dt
        except Exception as e:
            if world > 1:
                raise  # the other ranks are inside collectives: fail together
            e2e_error = repr(e)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            cpu = cpu_baselines(lp, "the full %dx%d Potts workload (same arrays as the GPU arm)" % (a.size, a.size))
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": its_per_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a),
            "roofline": roofline, "cpu_baseline": cpu, "variants": variants, "with_stats_block": with_stats,
            "latency_bound_configs": small,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "iters_per_call": a.e2e_iters, "calls": a.e2e_steps, "error": e2e_error,
                    "note": "each call: upload LP from pinned host memory, build SELL operators + transpose + "
                            "preconditioners on device, iterate, read x back"},
            # k_primal + k_dual per iteration; with N > 1 also k_push + k_wait after each of them (peer memory)
            # or one k_pack before each NCCL send/recv group
            # ... or nothing more when the halo is fused into the two kernels (flag 128)
            "gpu_launches": (2 if world == 1 or a.flags & 128 else (4 if a.flags & 32 else 6)) * a.steps * a.iters_per_step,
            "kernel_variants": kernel_variants(info),
            "clocks": clocks.summary(),
            "partition": None if world == 1 else {k: info[k] for k in (
            "n_local", "m_local", "n_ghost", "m_ghost", "nnz_local_rows", "nnz_local_cols",
            "halo_send_bytes_per_iteration", "partition_granule")},
        "problem": {"n": n, "m": m, "nnz": nnz, "build_host_s": round(t_build, 2), "setup_device_s": round(t_setup, 2),
                        "device_bytes": info["device_bytes"], "padding_A": info["a_padded_entries"] / nnz,
                        "padding_AT": info["at_padded_entries"] / nnz},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


VARIANT_NAMES = ("loop-unroll4/8cta", "chunk2/8cta", "chunk4/6cta", "chunk4/4cta", "chunk8/4cta", "rows2-chunk4/3cta",
                 "rows2-chunk2/4cta")


def kernel_variants(info):
    """Which compiled variant of each hot kernel cpppd_create kept, and the per-launch times it measured."""
    out = {"autotuned": bool(info["autotuned"])}
    for kernel, key in (("k_primal", "primal_variant"), ("k_dual", "dual_variant")):
        v = info[key]
        out[kernel] = {"variant": v, "name": VARIANT_NAMES[v - 1] if 1 <= v <= len(VARIANT_NAMES) else None,
                       "create_time_ms_per_launch": dict(zip(VARIANT_NAMES, info["variant_ms"][kernel]))
                       if info["autotuned"] else None}
    return out


def load_ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu summary (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(path) as f:
            return json.load(f)["kernels"][kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="Potts image side (4096 = BASELINE configs[4])")
    ap.add_argument("--iters-per-step", type=int, default=50)
    ap.add_argument("--ref-iters-per-step", type=int, default=1)
    ap.add_argument("--ref-numpy-iters", type=int, default=2,
                    help="--impl reference: iterations of the numpy/scipy restatement timed beside the C port (0: skip)")
    ap.add_argument("--small-configs", type=int, default=1, help="also time Potts 50x50 and SC105 (N = 1); 2: also with the opt-in persistent CTA")
    ap.add_argument("--small-iters", type=int, default=5000)
    ap.add_argument("--stats-interval", type=int, default=500, help="nb_iter_plot of the with_stats_block measurement")
    ap.add_argument("--e2e-iters", type=int, default=500)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variants", type=int, default=1, help="also time the opt-in storage variants (N = 1)")
    ap.add_argument("--flags", type=int, default=0, help="CPPPD_FLAG_* bit mask (8 reorder, 32 NCCL halos instead of peer memory, 128 halo fused into the kernels)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
