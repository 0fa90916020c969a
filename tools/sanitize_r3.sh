#!/usr/bin/env bash
# compute-sanitizer over the kernels added after profiles/r02_compute_sanitizer_*: the fp16 decoder dataflow (bn_relu_h,
# bnbwd_*_h, up_adjoint + statistics, wgrad_f16 with cout = 64, the decoder's fp16 data gradients), the tensor-core stem
# (forward and weight gradient), bias_grad_h and the fp16 u0 / du0 paths of the latent kernels -- through the small-shape
# training parity tests in both dispatches and the op-level stem tests.
#   gpurun --timeout 1500 -- 'bash tools/sanitize_r3.sh > gpurun_out/sanitize_r3.log 2>&1'
set -u
SEL='(golden and train_b2_g1_l128 and tf32_tc) or dropout_statistics'
STEM='2-1-512 or 3-3-1000 or 1-3-16'
for tool in memcheck racecheck synccheck initcheck; do
  for pm in 0 1; do
    echo "=== $tool: model path (NEF_TC_PERSIST_MIN=$pm)"
    NEF_TC_PERSIST_MIN=$pm compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -8
  done
  echo "=== $tool: tensor-core stem ops"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_stem_tc.py -m gpu -x -q -k "$STEM" 2>&1 | tail -8
done
