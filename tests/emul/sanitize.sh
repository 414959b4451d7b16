#!/bin/bash
# TEST INFRASTRUCTURE — the emulated library (every kernel and all host code of pysparselp_b200/csrc, compiled
# for the CPU by make_emul.py) under AddressSanitizer and UndefinedBehaviorSanitizer: an out-of-bounds load or
# store of a kernel, which a GPU may tolerate silently, aborts the run here.
#
#   bash tests/emul/sanitize.sh            (about an hour since the cluster kernel is emulated phase by phase; run from
#                                           the repo root.  The two deselected regression curves — 27 500 iterations
#                                           against a 150 s time-out — are too slow for the emulation.)
#
# Covers tests/test_library_on_cpu.py, tests/test_multi_rank_on_cpu.py and the dry run of the `-m gpu` tests.
set -e
cd "$(dirname "$0")/../.."
rebuild() { CPPPD_NVCC_DEFINES="$1" python -c "import sys; sys.path.insert(0, 'tests'); from emul import make_emul; make_emul.build(force=True)"; }
small="CPPPD_FULL_SIZE_POTTS=96 CPPPD_FULL_SIZE_RANDOM_N=20000 CPPPD_FULL_SIZE_SVM_SAMPLES=3000 CPPPD_FULL_SIZE_SVM_FEATURES=20"

rebuild "-fsanitize=address -fno-omit-frame-pointer"
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
LD_PRELOAD=$(gcc -print-file-name=libasan.so) python -m pytest tests/test_library_on_cpu.py tests/test_multi_rank_on_cpu.py tests/test_fuzz_on_cpu.py -x -q
env $small CPPPD_EMULATE_GPU_TESTS=1 PYTHONPATH=.:tests LD_PRELOAD=$(gcc -print-file-name=libasan.so) \
  python -m pytest tests -m gpu -p emul.patch_plugin -x -q \
  --deselect tests/test_gpu_variants.py::test_autotune_is_on_by_default_for_large_operands \
  --deselect tests/test_gpu_parity.py::test_reference_golden_curves_through_solve \
  --deselect tests/test_gpu_zz_opt_in_features.py::test_cluster_kernel_reproduces_the_potts50_regression_curve

rebuild "-fsanitize=undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer"
python -m pytest tests/test_library_on_cpu.py tests/test_multi_rank_on_cpu.py tests/test_fuzz_on_cpu.py -x -q

rebuild ""
echo "sanitizer runs clean"
