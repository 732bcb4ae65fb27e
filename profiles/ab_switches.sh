#!/usr/bin/env bash
# A/B of the opt-in kernel variants on one B200 (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/ab_switches.sh'
# Every step is bounded by its own `timeout`; logs land in gpurun_out/ab_*.log.  Parity first (the GPU test files that
# exercise each variant, with the switch set), then device-time microbenchmarks, then the full training-step bench.
set -u
mkdir -p gpurun_out
T="timeout 300"    # a hang of an unvalidated kernel must not hold the box
PY=python

echo "== default parity ==" | tee gpurun_out/ab_summary.log
$T $PY -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_default.log 2>&1; echo "default tests rc=$?" | tee -a gpurun_out/ab_summary.log

for sw in LVAE_CONV_F32_STAGE LVAE_DMOL_FAST LVAE_CONV_DYNAMIC; do
  case $sw in
    LVAE_CONV_F32_STAGE) files="tests/test_conv_tc_gpu.py tests/test_model_gpu.py tests/test_engine_gpu.py" ;;
    LVAE_DMOL_FAST)      files="tests/test_kernels_gpu.py tests/test_model_gpu.py" ;;
    LVAE_CONV_DYNAMIC)   files="tests/test_conv_tc_gpu.py tests/test_model_gpu.py tests/test_engine_gpu.py" ;;
  esac
  env $sw=1 $T $PY -m pytest $files -m gpu -q > gpurun_out/ab_tests_$sw.log 2>&1
  echo "$sw=1 tests rc=$? ($(tail -1 gpurun_out/ab_tests_$sw.log))" | tee -a gpurun_out/ab_summary.log
done

# the chained conv2 + gate kernel: kernel-level bit-exactness against the two-launch path, then the model tests with it on
LVAE_TEST_EXPERIMENTAL=1 $T $PY -m pytest tests/test_experimental_gpu.py -m gpu -q > gpurun_out/ab_tests_experimental.log 2>&1
echo "experimental tests rc=$? ($(tail -1 gpurun_out/ab_tests_experimental.log))" | tee -a gpurun_out/ab_summary.log
LVAE_CONV_GATE_CHAIN=1 LVAE_GATE_BWD_CHAIN=1 $T $PY -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -m gpu -q > gpurun_out/ab_tests_LVAE_GATE_CHAINS.log 2>&1
echo "LVAE_CONV_GATE_CHAIN=1 LVAE_GATE_BWD_CHAIN=1 tests rc=$? ($(tail -1 gpurun_out/ab_tests_LVAE_GATE_CHAINS.log))" | tee -a gpurun_out/ab_summary.log

echo "== microbenchmarks ==" | tee -a gpurun_out/ab_summary.log
$T $PY profiles/bench_hbm_kernels.py > gpurun_out/ab_hbm_default.log 2>&1
LVAE_DMOL_FAST=1 $T $PY profiles/bench_hbm_kernels.py > gpurun_out/ab_hbm_dmolfast.log 2>&1
$T $PY profiles/bench_conv_tc.py > gpurun_out/ab_conv_default.log 2>&1
LVAE_CONV_F32_STAGE=1 $T $PY profiles/bench_conv_tc.py > gpurun_out/ab_conv_f32stage.log 2>&1
LVAE_CONV_DYNAMIC=1 $T $PY profiles/bench_conv_tc.py > gpurun_out/ab_conv_dynamic.log 2>&1
grep -h "dmol" gpurun_out/ab_hbm_default.log gpurun_out/ab_hbm_dmolfast.log | grep -v '^{' | tee -a gpurun_out/ab_summary.log
grep -h "N=100" gpurun_out/ab_conv_default.log gpurun_out/ab_conv_f32stage.log | tee -a gpurun_out/ab_summary.log

echo "== training step (CIFAR-15, batch 256, bf16) ==" | tee -a gpurun_out/ab_summary.log
run_bench() {   # name, env assignments...
  local name=$1; shift
  env "$@" $T $PY bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err
  $PY - "$name" <<'PYEOF' | tee -a gpurun_out/ab_summary.log
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_bench_%s.json" % name).read().strip().splitlines()[-1])
    print("%-22s %8.3f ms/step  %9.1f images/s  loss %.2f" % (name, d["ms_per_step"], d["value"], d["loss"]))
except Exception as e:  # noqa: BLE001
    print("%-22s FAILED (%s)" % (name, e))
PYEOF
}
run_bench default LVAE_NOP=1
run_bench f32stage LVAE_CONV_F32_STAGE=1
run_bench dmolfast LVAE_DMOL_FAST=1
run_bench dynamic LVAE_CONV_DYNAMIC=1
run_bench gatechain LVAE_CONV_GATE_CHAIN=1
run_bench gatebwdchain LVAE_GATE_BWD_CHAIN=1
run_bench bothchains LVAE_CONV_GATE_CHAIN=1 LVAE_GATE_BWD_CHAIN=1
run_bench all LVAE_CONV_F32_STAGE=1 LVAE_DMOL_FAST=1 LVAE_CONV_DYNAMIC=1 LVAE_CONV_GATE_CHAIN=1 LVAE_GATE_BWD_CHAIN=1
echo "== IW-1000 (MNIST-12, batch 1000) ==" | tee -a gpurun_out/ab_summary.log
$T $PY bench.py --workload iw --steps 1 --no-cpu-baseline > gpurun_out/ab_iw_default.json 2>&1
LVAE_CONV_DYNAMIC=1 $T $PY bench.py --workload iw --steps 1 --no-cpu-baseline > gpurun_out/ab_iw_dynamic.json 2>&1
LVAE_CONV_GATE_CHAIN=1 $T $PY bench.py --workload iw --steps 1 --no-cpu-baseline > gpurun_out/ab_iw_gatechain.json 2>&1
for f in default dynamic gatechain; do $PY -c "
import json
d=json.loads(open('gpurun_out/ab_iw_$f.json').read().strip().splitlines()[-1]); print('iw $f', d['value'], d['unit'])" 2>&1 | tee -a gpurun_out/ab_summary.log; done
