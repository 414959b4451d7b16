"""CPU: pieces of the bench / ABI contract that need no GPU."""
import ctypes as C
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_algorithmic_bytes_match_survey_numbers():
    sys.path.insert(0, ROOT)
    import bench

    # Potts 4096^2 -> 11.20 GB per iteration; random LP 20M x 40M x 320M -> 10.80 GB
    p, d = bench.algorithmic_bytes(50323456, 67092480, 201277440)
    assert round((p + d) / 1e9, 2) == 11.20
    p, d = bench.algorithmic_bytes(20_000_000, 40_000_000, 320_000_000)
    assert round((p + d) / 1e9, 2) == 10.80
    assert bench.workload_nnz("potts", 4096) == 201277440 and bench.workload_nnz("random", 20_000_000) == 320_000_000


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` runs the oracle port on the host cores (no GPU involved)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "128",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "iterations/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--size", "128", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


C_PROBE = r"""
#include <stdio.h>
#include <stddef.h>
#include "cpppd.h"
int main(void) {
  printf("{\"problem\": %zu, \"stats\": %zu, \"info\": %zu, "
         "\"p_indptr\": %zu, \"p_alpha\": %zu, \"p_alloc\": %zu, \"p_rank\": %zu, \"p_granule\": %zu, "
         "\"s_energy1\": %zu, \"s_feasible\": %zu, \"i_device_bytes\": %zu, \"i_n_local\": %zu, \"i_granule\": %zu, "
         "\"abi\": %d}\n",
         sizeof(cpppd_problem), sizeof(cpppd_stats), sizeof(cpppd_info),
         offsetof(cpppd_problem, indptr), offsetof(cpppd_problem, alpha), offsetof(cpppd_problem, alloc),
         offsetof(cpppd_problem, rank), offsetof(cpppd_problem, partition_granule),
         offsetof(cpppd_stats, energy1), offsetof(cpppd_stats, feasible),
         offsetof(cpppd_info, device_bytes), offsetof(cpppd_info, n_local), offsetof(cpppd_info, partition_granule),
         CPPPD_ABI_VERSION);
  return 0;
}
"""


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Compile a probe against include/cpppd.h with gcc and compare sizes / offsets with the binding."""
    from pysparselp_b200 import _cabi

    src = tmp_path / "probe.c"
    src.write_text(C_PROBE)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    c = json.loads(subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout)
    assert c["abi"] == _cabi.ABI_VERSION
    assert c["problem"] == C.sizeof(_cabi.Problem) and c["stats"] == C.sizeof(_cabi.Stats) and c["info"] == C.sizeof(_cabi.Info)
    P, S, I = _cabi.Problem, _cabi.Stats, _cabi.Info
    assert (c["p_indptr"], c["p_alpha"], c["p_alloc"], c["p_rank"], c["p_granule"]) == (
        P.indptr.offset, P.alpha.offset, P.alloc.offset, P.rank.offset, P.partition_granule.offset)
    assert (c["s_energy1"], c["s_feasible"]) == (S.energy1.offset, S.feasible.offset)
    assert (c["i_device_bytes"], c["i_n_local"], c["i_granule"]) == (
        I.device_bytes.offset, I.n_local.offset, I.partition_granule.offset)


def test_b200_arm_dry_run_on_the_emulated_library(monkeypatch, capsys):
    """The whole b200 arm of bench.py — device-timed steps, per-kernel roofline, storage variants, end-to-end
    calls, CPU baseline, the JSON line — executed here with the solver swapped for the CPU-emulated library and
    torch's CUDA entry points stubbed.  The numbers mean nothing; the point is that every statement of the script
    the driver runs at round end has been executed once, and that the line carries every key of the contract."""
    import argparse

    import numpy as np
    import torch

    sys.path.insert(0, ROOT)
    import bench
    import pysparselp_b200.ChambollePockPPD as front
    from emul.patch_plugin import _Adapter

    monkeypatch.setattr(front, "CpPpdSolver", _Adapter)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(bench, "pinned_empty", lambda: (lambda count, dtype=np.float64: np.empty(int(count), dtype=dtype), []))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    torch.cuda.empty_cache()  # (a no-op without a device; the script calls it between the workloads)
    a = argparse.Namespace(gpus=1, steps=2, warmup=3, impl="b200", workload="potts", size=24, iters_per_step=3,
                           ref_iters_per_step=1, ref_numpy_iters=1, e2e_iters=7, e2e_steps=2, no_cpu_baseline=False,
                           variants=1, flags=0, stats_interval=5, small_configs=2, small_iters=30, secondary="random,l1svm:60",
                           secondary_size=2000)
    bench.run_b200(a)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["value"] > 0
    assert d["config"]["workload"] == "potts_segmentation_lp_24x24" and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"].startswith(("k_primal[", "k_dual["))
    # the timed configuration reproduces the digest the C port minted for this workload (tests/golden/bench_digests.json)
    assert d["parity"]["status"] == "ok", d["parity"]
    sec = d["secondary_workloads"]["random_sparse_lp_2000x4000_8_per_row"]
    assert sec["value"] > 0 and sec["parity"]["status"] == "ok" and sec["roofline"]["frac"] > 0
    assert sec["e2e"]["value"] > 0 and sec["e2e"]["finite"] and "error" not in sec["sell_kernels"]
    svm = d["secondary_workloads"]["l1svm_lp_60_samples_x_1000_features"]  # (no digest committed for this size)
    assert svm["value"] > 0 and svm["roofline"]["frac"] > 0 and svm["e2e"]["finite"] and "sell_kernels" not in svm
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["achieved"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    assert e["error"] is None and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert len(e["call_seconds"]) == 2 and len(e["call_phases_rank0"]) == 2
    assert all(set(ph) == {"setup", "iterate", "close"} and min(ph.values()) >= 0 for ph in e["call_phases_rank0"])
    assert d["gpu_launches"] == 2 * 2 * 3
    assert d["with_stats_block"]["iterations"] == 10 and d["with_stats_block"]["iterations_per_s"] > 0
    small = d["latency_bound_configs"]
    for name in ("potts_50x50", "netlib_sc105"):  # (the emulation drives the cluster kernel phase by phase)
        assert small[name]["cuda_graphs"]["iterations_per_s"] > 0 and small[name]["persistent"]["iterations_per_s"] > 0
        assert "k_cluster_iterate" in small[name]["persistent"]["kernel"]
    assert set(d["variants"]) == {"reorder", "compressed", "compressed+reorder"}
    assert all("error" not in v for v in d["variants"].values()), d["variants"]
    assert d["kernel_variants"]["k_primal"]["variant"] == 1 and d["kernel_variants"]["autotuned"] is False


def test_smoke_dry_run_on_the_emulated_library(monkeypatch, capsys):
    """__graft_entry__.smoke() with the solver swapped for the CPU-emulated library: the statements the driver runs on
    the GPU box before the bench, executed once here."""
    import torch

    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    import pysparselp_b200.ChambollePockPPD as front
    from emul.patch_plugin import _Adapter

    monkeypatch.setattr(front, "CpPpdSolver", _Adapter)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


def test_committed_bench_digests_are_reproducible():
    """tests/golden/bench_digests.json (bench.py's `parity` key): the small entries are re-minted here from the C
    port and must equal the committed ones — the generators and the port are deterministic across machines and
    thread counts; the full-size entries were minted by the same script (tools/mint_bench_digests.py)."""
    sys.path.insert(0, ROOT)
    import bench
    from oracle.c_port import COracle
    from pysparselp_b200 import generators

    table = json.load(open(bench.DIGESTS))
    for name in (bench.workload_name("potts", 4096), bench.workload_name("random", 20_000_000)):
        assert name in table and len(table[name]["sha256"]) == 64 and table[name]["iterations"] == bench.DIGEST_ITERS
    for kind, size in (("potts", 256), ("random", 200000)):
        lp, _ = bench.build_workload(kind, size, pinned=False)
        co = COracle(*generators.lp_args(lp))
        co.iterate(bench.DIGEST_ITERS)
        assert bench.iterate_digest(co.x, co.y) == table[bench.workload_name(kind, size)]["sha256"]
