#!/usr/bin/env python
"""Benchmark of the CP-PPD hot path (BASELINE.json metric: iterations/s + effective HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload potts|random|l1svm]

Headline workload (N = 1 and N > 1 alike): the synthetic 4096x4096 Potts-segmentation LP of
BASELINE.json configs[4] (n = 50 323 456, m = 67 092 480, nnz = 201 277 440), fp64.
``--workload random`` makes BASELINE configs[3] the headline instead (randomLP.py family: 20 M variables, 40 M
inequality rows, exactly 8 entries per row, 320 M entries), ``--workload l1svm`` the configs[2] family (``--size``
samples x 1 000 features).  The default run also measures, after the headline, the random LP (any N: with N > 1 its
distributed solve) and — one GPU — the L1-SVM LP with 100 000 samples, and reports them under ``secondary_workloads``
(device-resident iterations/s, per-half-iteration times, roofline, storage chosen, parity verdict).
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
                the caches and are bound by launch latency, not bandwidth (extra information, N = 1): CUDA graphs
                against the persistent kernel (one thread-block cluster).
* ``roofline``: algorithmic bytes of the dominant kernel / its CUDA-event time,
                against MEASURED_PEAKS.json's hbm_gbs (fallback 6650 GB/s).
* ``parity``  : x and y of a fresh solver after 6 iterations, sha256, against the digest the plain-C oracle port
                produced for this workload (tests/golden/bench_digests.json, minted by tools/mint_bench_digests.py): the
                timed configuration — any N, any transport — must reproduce the reference's bits.  A mismatch makes
                the process exit non-zero.
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


DEFAULT_SIZE = {"potts": 4096, "random": 20_000_000, "l1svm": 100_000}
SVM_FEATURES = 1000
DIGEST_ITERS = 6
DIGESTS = os.path.join(ROOT, "tests", "golden", "bench_digests.json")
# measured by tools/probe/gather_probe.cu on a B200 (profiles/r02_random_lp.md): nanoseconds per 1000 random 8-byte
# gathers while the gathered window stays L2-resident and the matrix streams past it
L2_GATHER_NS_PER_1000 = 4.75


def workload_name(kind, size):
    if kind == "potts":
        return "potts_segmentation_lp_%dx%d" % (size, size)
    if kind == "random":
        return "random_sparse_lp_%dx%d_8_per_row" % (size, 2 * size)
    return "l1svm_lp_%d_samples_x_%d_features" % (size, SVM_FEATURES)


def workload_generator(kind):
    return {"potts": "pysparselp_b200.generators.potts_lp(seed=1, coef_potts=0.5, coef_mul=500)",
            "random": "pysparselp_b200.generators.random_sparse_lp_chunked(n, 2 n, nnz_per_row=8, seed=0) "
                      "(restatement of pysparselp/randomLP.py:14-75, inequalities only)",
            "l1svm": "pysparselp_b200.generators.l1svm_lp(samples, 1000, nb_classes=3, seed=1) "
                     "(pysparselp/examples/example_l1_svm.py:13-68)"}[kind]


def build_workload(kind, size, pinned):
    from pysparselp_b200 import generators

    empty, keep = (pinned_empty() if pinned else (np.empty, None))
    if kind == "potts":
        lp = generators.potts_lp(size, empty=empty) if pinned else generators.potts_lp(size)
    elif kind == "random":
        lp, _ = generators.random_sparse_lp_chunked(size, 2 * size, nnz_per_row=8, seed=0, empty=empty)
    elif kind == "l1svm":
        lp, _ = generators.l1svm_lp(size, SVM_FEATURES)
    else:
        raise SystemExit("unknown workload %r" % (kind,))
    return lp, keep


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


def iterate_digest(x, y):
    import hashlib

    h = hashlib.sha256()
    h.update(np.ascontiguousarray(x, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(y, dtype=np.float64).tobytes())
    return h.hexdigest()


def iterate_fingerprint(x, y, count=64):
    """Workloads whose sums are not bit-reproducible (rows / columns above the long-row threshold are summed by a fixed
    tree instead of sequentially: the L1-SVM weight columns) are compared through sampled entries and 1-norms."""
    rng = np.random.default_rng(0)
    xi = np.sort(rng.choice(x.size, size=min(count, x.size), replace=False))
    yi = np.sort(rng.choice(y.size, size=min(count, y.size), replace=False))
    return {"x_idx": xi.tolist(), "x": x[xi].tolist(), "y_idx": yi.tolist(), "y": y[yi].tolist(),
            "max_abs_x": float(np.max(np.abs(x))), "max_abs_y": float(np.max(np.abs(y))),
            "sum_abs_x": float(np.sum(np.abs(x))), "sum_abs_y": float(np.sum(np.abs(y)))}


def fingerprint_error(x, y, want):
    """Largest deviation from a committed fingerprint, relative to the vector's max-norm (BASELINE: <= 1e-9)."""
    err = 0.0
    for v, key in ((x, "x"), (y, "y")):
        scale = max(want["max_abs_" + key], 1e-300)
        err = max(err, float(np.max(np.abs(v[np.asarray(want[key + "_idx"])] - np.asarray(want[key])))) / scale)
        err = max(err, abs(float(np.sum(np.abs(v))) - want["sum_abs_" + key]) / max(want["sum_abs_" + key], 1e-300))
    return err


def parity_check(make_solver, args, name, flags):
    """x, y of a fresh solver after DIGEST_ITERS iterations against what the C port produced for this workload: the
    sha256 of the bits, or — workloads with long rows — a sampled fingerprint within 1e-9 relative."""
    try:
        with open(DIGESTS) as f:
            want = json.load(f).get(name)
    except Exception:
        want = None
    s = make_solver(*args, flags=flags)
    try:
        s.iterate(DIGEST_ITERS)
        x, y = s.get_x(), s.get_y()
    finally:
        s.close()
    got = iterate_digest(x, y)
    if want is None:
        return {"status": "no digest committed for this workload", "sha256": got, "iterations": DIGEST_ITERS}
    if "fingerprint" in want:
        err = fingerprint_error(x, y, want["fingerprint"])
        return {"status": "ok" if err <= 1e-9 else "MISMATCH", "max_relative_error": err, "tolerance": 1e-9,
                "iterations": DIGEST_ITERS, "minted_by": want.get("minted_by"),
                "note": "sampled entries and 1-norms of x, y against the C port (long rows: fixed summation tree, not bit-exact)"}
    return {"status": "ok" if got == want["sha256"] else "MISMATCH", "sha256": got, "expected": want["sha256"],
            "iterations": DIGEST_ITERS, "minted_by": want.get("minted_by")}


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
    # bounded sample: full size when the requested number of steps allows it, else a smaller instance of the family
    size = a.size
    budget_iters = (a.steps + a.warmup) * a.ref_iters_per_step
    if a.workload == "potts":
        while size > 512 and budget_iters * 0.45 * (size / 4096.0) ** 2 > 150:  # ~0.45 s/iteration at 4096^2 on 8 cores
            size //= 2
    else:
        floor = {"random": 1_000_000, "l1svm": 5_000}[a.workload]
        full = {"random": 1.5, "l1svm": 2.0}[a.workload] * a.size / DEFAULT_SIZE[a.workload]  # s/iteration, 8 cores
        while size > floor and budget_iters * full * size / a.size > 150:
            size //= 2
    lp, _ = build_workload(a.workload, size, pinned=False)
    co = COracle(*generators.lp_args(lp))
    nnz = int(co.val.size)
    scale = nnz / float(workload_nnz(a.workload, a.size))
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
    sample = ("each step = %d iteration(s) of %s (nnz %d) on the CPU, value scaled by nnz ratio %.4f "
              "to %s; plain-C OpenMP oracle port of pysparselp/ChambollePockPPD.py:195-343, %d threads"
              % (a.ref_iters_per_step, workload_name(a.workload, size), nnz, scale, workload_name(a.workload, a.size), cores))
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


def workload_nnz(kind, size):
    if kind == "potts":
        return 3 * 2 * (size * (size - 1) * 2)
    if kind == "random":
        return 8 * 2 * size
    # L1-SVM, K = 3: two blocks of K (F + 1) two-entry rows, then 2 N rows (examples x the two other classes) of 2 (F + 1) + 1
    return 4 * 3 * (SVM_FEATURES + 1) + 2 * size * (2 * (SVM_FEATURES + 1) + 1)


def workload_config(a, kind=None, size=None):
    kind, size = kind or a.workload, size or a.size
    return {"workload": workload_name(kind, size), "generator": workload_generator(kind),
            "iters_per_step": a.iters_per_step, "theta": 1, "alpha": 1,
            "l2": "inputs larger than L2 (gigabytes streamed per iteration vs 126 MB L2); no flush needed",
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
            "value_bytes": info["value_bytes"], "const_vector_mask": info["const_vector_mask"],
            "banded": info["band_in_use"]}


def half_iteration_names(info):
    """What runs the primal / dual half-iteration of this handle, as profiles/ncu_summary.json keys it."""
    dict_tag = ",dict" if info["value_bytes"] == 0 else ""
    out = {}
    for kernel, key, band in (("k_primal", "primal_variant", 1), ("k_dual", "dual_variant", 0)):
        if info["band_in_use"][band]:
            out[kernel] = "%s_band x %d windows" % (kernel, info["band_windows"][band])
        else:
            v = info[key]
            out[kernel] = "%s[%s%s]" % (kernel, VARIANT_NAMES[v - 1] if 1 <= v <= len(VARIANT_NAMES) else "?", dict_tag)
    return out


def roofline_block(solver, info, its_per_s, world, peak, peak_src):
    """Per-half-iteration times (CUDA events between the halves) -> roofline of the dominant one."""
    n, m, nnz = info["n"], info["m_eq"] + info["m_ineq"], info["nnz"]
    kp, kd = solver.time_kernels(32)
    kp, kd = kp / 32, kd / 32
    bp, bd = algorithmic_bytes(n, m, nnz)
    if world > 1:  # per-rank share of the algorithmic bytes
        bp, bd = bp / world, bd / world
    names = half_iteration_names(info)
    dom = ("k_primal", kp, bp) if kp >= kd else ("k_dual", kd, bd)
    traffic, traffic_src = load_ncu_traffic(names[dom[0]])
    block = {
        "bound": "hbm", "kernel": names[dom[0]], "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom[2], "avg_launch_ms": dom[1],
        "traffic_source": traffic_src,
        "kernels": {"k_primal": {"runs": names["k_primal"], "ms": kp, "algorithmic_GBs": bp / (kp * 1e-3) / 1e9},
                    "k_dual": {"runs": names["k_dual"], "ms": kd, "algorithmic_GBs": bd / (kd * 1e-3) / 1e9}},
        "iteration": {"algorithmic_bytes": info["bytes_per_iteration_algorithmic"],
                      "effective_GBs": info["bytes_per_iteration_algorithmic"] * its_per_s / 1e9,
                      "frac_of_peak": info["bytes_per_iteration_algorithmic"] * its_per_s / 1e9 / (peak * world),
                      "actual_bytes": info["bytes_per_iteration_actual"],
                      "actual_GBs": info["bytes_per_iteration_actual"] * its_per_s / 1e9,
                      "note": "whole job: algorithmic bytes of the full LP x iterations/s, against n_gpus x peak"},
    }
    if any(info["band_in_use"]):
        # a banded half-iteration is `windows` launches; its gathers are served by the L2, whose sector throughput
        # (one 32-byte sector per 8-byte gather) is the second bound next to HBM
        gathers = 2 * nnz / world
        floor_ms = gathers * L2_GATHER_NS_PER_1000 / 1e3 / 1e6
        block["l2_gather_bound"] = {
            "gathers_per_iteration": gathers, "ns_per_1000_gathers_measured": L2_GATHER_NS_PER_1000,
            "floor_ms_per_iteration": floor_ms, "frac_of_floor": floor_ms / (kp + kd),
            "band_windows": info["band_windows"], "band_window_bytes": info["band_window_bytes"],
            "sectors_per_gather_sampled": info["band_sectors_per_gather"],
            "note": "tools/probe/gather_probe.cu: random 8-byte gathers into an L2-resident window cost a 32-byte sector "
                    "each; 640 M gathers per iteration of the 20Mx40M LP bound the iteration below the HBM roofline"}
    return block


def time_resident(solver, a, torch, dist, local_rank, steps, warmup):
    """W warm-up steps, then K steps between CUDA events on the solver's stream, max over ranks."""
    for _ in range(warmup):
        solver.iterate(a.iters_per_step)
    solver.sync()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        ms = solver.time_iterations(steps * a.iters_per_step)
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, clocks.summary()


def storage_summary(info, t_build, t_setup):
    nnz = max(info["nnz"], 1)
    return {"n": info["n"], "m": info["m_eq"] + info["m_ineq"], "nnz": info["nnz"], "build_host_s": round(t_build, 2),
            "setup_device_s": round(t_setup, 2), "device_bytes": info["device_bytes"],
            "padding_A": info["a_padded_entries"] / nnz, "padding_AT": info["at_padded_entries"] / nnz,
            "banded": {"in_use": info["band_in_use"], "windows": info["band_windows"],
                       "window_bytes": info["band_window_bytes"], "create_time_ms": info["band_ms"],
                       "sectors_per_gather": info["band_sectors_per_gather"]}}


def secondary_workload(kind, size, a, torch, local_rank, make_solver, chambolle_pock_ppd, peak, peak_src, dist=None,
                       world=1):
    """Device-resident measurement of another BASELINE config (after the headline, own roofline).  One GPU: also the
    SELL kernels beside the banded ones (random LP) and one end-to-end call.  N GPUs (every rank calls this): the
    distributed solve of the same LP, device-timed as the max over ranks, with its parity verdict."""
    from pysparselp_b200 import generators

    t0 = time.perf_counter()
    lp, keep = build_workload(kind, size, pinned=True)
    t_build = time.perf_counter() - t0
    args = generators.lp_args(lp)
    out = {"config": workload_config(a, kind, size)}
    t0 = time.perf_counter()
    solver = make_solver(*args, flags=a.flags)
    t_setup = time.perf_counter() - t0
    try:
        info = solver.info()
        steps = max(2, a.steps // 4)
        ms, clocks = time_resident(solver, a, torch, dist, local_rank, steps, 3)
        its = steps * a.iters_per_step / (ms * 1e-3)
        out.update({"value": its, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": ms / steps,
                    "clocks": clocks, "roofline": roofline_block(solver, info, its, world, peak, peak_src),
                    "kernel_variants": kernel_variants(info), "problem": storage_summary(info, t_build, t_setup)})
    finally:
        solver.close()
    out["parity"] = parity_check(make_solver, args, workload_name(kind, size), a.flags)
    if world > 1:
        out["partition"] = {k: info[k] for k in ("n_local", "m_local", "n_ghost", "m_ghost", "halo_send_bytes_per_iteration",
                                                 "balanced_split")}
        del lp, keep
        return out
    # the SELL kernels on the same LP (what ran before the banded operands existed)
    if kind == "random":
        try:
            out["sell_kernels"] = time_variant(make_solver, args, a.flags | 2048, a.iters_per_step // 2, peak)
        except Exception as e:
            out["sell_kernels"] = {"error": repr(e)}
    # end to end, one call: upload from pinned host memory, build both operand forms, iterate, read x back
    try:
        iters = max(100, a.e2e_iters // 2)
        chambolle_pock_ppd(*args, nb_max_iter=2, nb_iter_plot=2, flags=a.flags)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, _ = chambolle_pock_ppd(*args, nb_max_iter=iters, nb_iter_plot=iters, flags=a.flags)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": iters / dt, "unit": UNIT, "h2d_bytes_per_step": lp_nbytes(lp), "d2h_bytes_per_step": 8 * x.size + 96,
                      "iters_per_call": iters, "calls": 1, "finite": bool(np.all(np.isfinite(x)))}
    except Exception as e:
        out["e2e"] = {"value": None, "error": repr(e)}
    del lp, keep
    return out


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
    lp, keep = build_workload(a.workload, a.size, pinned=True)
    t_build = time.perf_counter() - t_build
    args = generators.lp_args(lp)

    # ---- device-resident timing -------------------------------------------------------------
    t_setup = time.perf_counter()
    solver = make_solver(*args, flags=a.flags)
    t_setup = time.perf_counter() - t_setup
    info = solver.info()
    n, m, nnz = info["n"], info["m_eq"] + info["m_ineq"], info["nnz"]
    ms, clock_summary = time_resident(solver, a, torch, dist, local_rank, a.steps, a.warmup)
    its_per_s = a.steps * a.iters_per_step / (ms * 1e-3)

    # ---- per-kernel times (events between kernels) -> roofline of the dominant kernel --------
    peak, peak_src = measured_peak_gbs()
    roofline = roofline_block(solver, info, its_per_s, world, peak, peak_src)
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
                      "note": "device-resident, wall clock around %d stats intervals (stats block + 128-byte read-back each)" % blocks}
    except Exception as e:
        if world > 1:
            raise
        with_stats = {"error": repr(e)}
    solver.close()
    del solver

    # ---- parity of the timed configuration (any N): digest of x, y after 6 iterations vs the C port's
    parity = parity_check(make_solver, args, workload_name(a.workload, a.size), a.flags)

    # ---- opt-in variants of the same solve (bit-identical iterates, different storage), N = 1 only
    variants = None
    if world == 1 and a.variants and a.workload == "potts":
        variants = {}
        for name, vflags in (("reorder", 8), ("compressed", 3), ("compressed+reorder", 11)):
            try:
                variants[name] = time_variant(make_solver, args, vflags, a.iters_per_step, peak)
            except Exception as e:  # a variant must never take the headline measurement down
                variants[name] = {"flags": vflags, "error": repr(e)}

    # ---- the two CPU-runnable configs of BASELINE.json (configs[0] Potts 50x50, configs[1] netlib SC105): launch-latency
    #      bound ("report it/s only"); CUDA graphs of 50 iterations vs one persistent CTA
    small = None
    if world == 1 and a.small_configs:
        small = small_configs(a, generators, make_solver)

    # ---- end to end through the public API with host buffers -----------------------------------
    h2d = lp_nbytes(lp)
    d2h = 8 * n + 96
    e2e_value, e2e_error, e2e_calls = None, None, []
    if a.e2e_steps > 0:
        e2e_phases = []

        def one_call():
            phases = {}
            x, best = chambolle_pock_ppd(*args, nb_max_iter=a.e2e_iters, nb_iter_plot=a.e2e_iters, flags=a.flags,
                                         timings=phases)
            e2e_phases.append({k: round(v, 4) for k, v in phases.items()})
            return x

        try:  # (a failure here must not take the device-timed headline down with it; it is reported in the line)
            # warm-up: allocator pools (device and pinned host), graph instantiation, kernel-variant timing of this
            # operand shape.  Twice, and the previous result is dropped before every call: a call that has to create a
            # fresh 400 MB pinned result buffer (cudaHostAlloc) because the last x is still referenced takes 0.1-0.5 s
            # longer (seen as 1.56 / 1.17 s calls among 1.03 s ones)
            one_call()
            one_call()
            barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(a.e2e_steps):
                x = None
                t_call = time.perf_counter()
                x = one_call()
                e2e_calls.append(round(time.perf_counter() - t_call, 4))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
                per_rank = [None] * world
                dist.all_gather_object(per_rank, e2e_calls)
                e2e_calls = per_rank  # [rank][call]: one slow rank holds every other one at the next collective
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
            cpu = cpu_baselines(lp, "the full %s workload (same arrays as the GPU arm)" % workload_name(a.workload, a.size))
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % (e,)}

    # ---- BASELINE configs[3] (random sparse LP) and the configs[2] family (L1-SVM) beside the headline, default run.
    #      N > 1: the random LP only — the distributed solve (balanced split, banded operands, dense halo over peer memory)
    secondary = None
    if a.secondary and a.workload == "potts":
        del lp, keep, args
        torch.cuda.empty_cache()
        secondary = {}
        for item in a.secondary.split(","):
            kind, _, own_size = item.strip().partition(":")
            if world > 1 and kind != "random":
                continue
            size2 = int(own_size) if own_size else (a.secondary_size or DEFAULT_SIZE[kind])
            try:
                secondary[workload_name(kind, size2)] = secondary_workload(kind, size2, a, torch, local_rank, make_solver,
                                                                           chambolle_pock_ppd, peak, peak_src, dist, world)
            except Exception as e:
                if world > 1:
                    raise  # the other ranks are inside collectives: fail together
                secondary[kind] = {"error": repr(e)}

    if rank == 0:
        split_wait = os.environ.get("CPPPD_SPLIT_HALO_WAIT", "0") not in ("", "0")
        launches_per_iteration = 2 if world == 1 or a.flags & 128 else (6 if split_wait and not a.flags & 32 else 4)
        if any(info["band_in_use"]):  # a banded half-iteration is one launch per window
            launches_per_iteration = sum(info["band_windows"][k] if info["band_in_use"][k] else 1 for k in (0, 1))
        line = {
            "metric": METRIC, "value": its_per_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(a),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "variants": variants,
            "with_stats_block": with_stats, "latency_bound_configs": small, "secondary_workloads": secondary,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "iters_per_call": a.e2e_iters, "calls": a.e2e_steps, "call_seconds": e2e_calls, "error": e2e_error,
                    "call_phases_rank0": e2e_phases[-a.e2e_steps:] if a.e2e_steps > 0 else None,
                    "note": "each call: upload LP from pinned host memory, build SELL operators + transpose + "
                            "preconditioners on device, iterate, read x back"},
            # k_primal + k_dual per iteration; with N > 1 also one k_push after each of them (peer memory; it also waits
            # for the incoming halo — a separate k_wait only with CPPPD_SPLIT_HALO_WAIT=1) or one k_pack before each
            # NCCL send/recv group
            # ... or nothing more when the halo is fused into the two kernels (flag 128)
            "gpu_launches": launches_per_iteration * a.steps * a.iters_per_step,
            "kernel_variants": kernel_variants(info),
            "clocks": clock_summary,
            "partition": None if world == 1 else {k: info[k] for k in (
                "n_local", "m_local", "n_ghost", "m_ghost", "nnz_local_rows", "nnz_local_cols",
                "halo_send_bytes_per_iteration", "partition_granule", "balanced_split")},
            "problem": storage_summary(info, t_build, t_setup),
        }
        print(json.dumps(line))
        sys.stdout.flush()
    if dist is not None:
        dist.destroy_process_group()
    if parity["status"] == "MISMATCH":
        raise SystemExit("parity: x, y after %d iterations differ from the C port's digest" % DIGEST_ITERS)


def small_configs(a, generators, make_solver):
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
            # (4096: CPPPD_FLAG_NO_TINY_PERSISTENT; the default picks one persistent CTA (k_tiny_iterate) when the LP fits one
            # SM, one persistent thread-block cluster of 16 CTAs (k_cluster_iterate) when it fits 16)
            for label, sflags in (("cuda_graphs", 4096), ("persistent", 0)):
                try:
                    ss = make_solver(*sargs, flags=sflags)
                    try:
                        ss.iterate(a.small_iters)
                        ss.sync()
                        ms_small = ss.time_iterations(a.small_iters)
                        used = int(ss.info()["tiny_persistent"])
                    finally:
                        ss.close()
                    if label == "persistent" and not used:
                        small[name][label] = None  # the LP fits neither: graph path
                    else:
                        small[name][label] = {"iterations_per_s": a.small_iters / (ms_small * 1e-3)}
                        if label == "persistent":
                            small[name][label]["kernel"] = {1: "k_tiny_iterate (one CTA)", 2: "k_cluster_iterate (one cluster)"}[used]
                except Exception as e:
                    small[name][label] = {"error": repr(e)}
    except Exception as e:
        small = {"error": repr(e)}
    return small


VARIANT_NAMES = ("loop-unroll4/8cta", "chunk2/8cta", "chunk4/6cta", "chunk4/4cta", "chunk8/4cta", "rows2-chunk4/3cta",
                 "rows2-chunk2/4cta", "stride-chunk2/8cta", "stride-chunk4/6cta")


def kernel_variants(info):
    """Which compiled variant of each hot kernel cpppd_create kept, and the per-launch times it measured."""
    out = {"autotuned": bool(info["autotuned"])}
    for kernel, key, band in (("k_primal", "primal_variant", 1), ("k_dual", "dual_variant", 0)):
        v = info[key]
        out[kernel] = {"variant": v, "name": VARIANT_NAMES[v - 1] if 1 <= v <= len(VARIANT_NAMES) else None,
                       "create_time_ms_per_launch": dict(zip(VARIANT_NAMES, info["variant_ms"][kernel]))
                       if info["autotuned"] else None,
                       "banded": bool(info["band_in_use"][band]),
                       "banded_create_time_ms_per_half_iteration": info["band_ms"][band] or None}
    return out


def load_ncu_traffic(key):
    """(dram bytes per launch, source) of the kernel variant `key` from the committed ncu summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    try:
        with open(path) as f:
            rec = json.load(f)["kernels"][key]
        return rec["dram_bytes_per_launch"], "profiles/ncu_summary.json[%s]: %s" % (key, rec.get("source"))
    except Exception:
        return None, "no ncu capture of %s under profiles/" % key


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="potts", choices=sorted(DEFAULT_SIZE),
                    help="headline LP: potts (BASELINE configs[4]), random (configs[3]) or l1svm (configs[2] family)")
    ap.add_argument("--size", type=int, default=0, help="Potts image side (4096) / random-LP variables (20 M; rows = 2x) / "
                                                         "L1-SVM samples (x 1000 features); 0: the workload's default")
    ap.add_argument("--secondary", default="random,l1svm",
                    help="--workload potts: also measure these workloads (comma separated, kind or kind:size, '' for none) "
                         "under secondary_workloads; with N > 1 only the random LP")
    ap.add_argument("--secondary-size", type=int, default=0)
    ap.add_argument("--iters-per-step", type=int, default=50)
    ap.add_argument("--ref-iters-per-step", type=int, default=1)
    ap.add_argument("--ref-numpy-iters", type=int, default=2,
                    help="--impl reference: iterations of the numpy/scipy restatement timed beside the C port (0: skip)")
    ap.add_argument("--small-configs", type=int, default=1, help="also time Potts 50x50 and SC105 (N = 1): CUDA graphs vs one persistent CTA")
    ap.add_argument("--small-iters", type=int, default=5000)
    ap.add_argument("--stats-interval", type=int, default=500, help="nb_iter_plot of the with_stats_block measurement")
    ap.add_argument("--e2e-iters", type=int, default=500)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variants", type=int, default=1, help="also time the opt-in storage variants (N = 1)")
    ap.add_argument("--flags", type=int, default=0, help="CPPPD_FLAG_* bit mask (8 reorder, 32 NCCL halos instead of peer memory, 128 halo fused into the kernels, 1024 banded operands, 2048 never banded)")
    a = ap.parse_args()
    a.size = a.size or DEFAULT_SIZE[a.workload]
    a.warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
