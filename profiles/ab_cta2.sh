#!/usr/bin/env bash
# First run of the cta_group::2 (CTA pair) convolution kernels, LVAE_CONV_CTA2=1 -- written without a GPU, so every step
# has a SHORT timeout of its own (a hang must not hold the box until gpurun's limit):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash profiles/ab_cta2.sh'
set -u
mkdir -p gpurun_out
export LVAE_CONV_CTA2=1
# 1. one shape, eager, fp32-output kernel (16x16, B = 2: 4 tiles = 2 pairs)
timeout 120 python -m pytest "tests/test_conv_tc_gpu.py::test_tc_conv_forward_backward[case0]" -m gpu -x -q > gpurun_out/cta2_t1.log 2>&1
echo "cta2 single case rc=$? $(tail -1 gpurun_out/cta2_t1.log)" | tee gpurun_out/cta2_summary.log
# 2. all conv shapes (halo shapes with an even tile count take the pair kernels, the others the usual ones)
timeout 180 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q > gpurun_out/cta2_t2.log 2>&1
echo "cta2 conv tests rc=$? $(tail -1 gpurun_out/cta2_t2.log)" | tee -a gpurun_out/cta2_summary.log
# 3. microbenchmark per shape, pair vs single-CTA kernels
timeout 120 python profiles/bench_conv_tc.py > gpurun_out/cta2_conv_pair.log 2>&1; echo "bench pair rc=$?" | tee -a gpurun_out/cta2_summary.log
LVAE_CONV_CTA2=0 timeout 120 python profiles/bench_conv_tc.py > gpurun_out/cta2_conv_single.log 2>&1
paste -d'|' gpurun_out/cta2_conv_single.log gpurun_out/cta2_conv_pair.log | tee -a gpurun_out/cta2_summary.log
# 4. model parity and the training step
timeout 300 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -m gpu -q > gpurun_out/cta2_t3.log 2>&1
echo "cta2 model tests rc=$? $(tail -1 gpurun_out/cta2_t3.log)" | tee -a gpurun_out/cta2_summary.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/cta2_bench.json 2> gpurun_out/cta2_bench.err
echo "cta2 bench rc=$? $(cut -c1-260 gpurun_out/cta2_bench.json)" | tee -a gpurun_out/cta2_summary.log
