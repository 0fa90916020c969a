#!/usr/bin/env bash
# compute-sanitizer over the kernels changed after profiles/r03_compute_sanitizer_*: the coalesced drain of wgrad_f16_kernel
# (the drain warps reuse the pipeline's shared memory as a transpose slab), the L2-prefetching epilogue of the persistent conv
# kernel (NEF_TC_PERSIST_MIN=1 forces it at small shapes) and dec_out_bwd_h (small-shape training parity test).
#   gpurun --timeout 1500 -- 'bash tools/sanitize_r4.sh > gpurun_out/sanitize_r4.log 2>&1'
set -u
SEL='(golden and train_b2_g1_l128 and tf32_tc) or dropout_statistics'
WG='2-122-1-64-1 or 3-250-2-128-3 or 4-500-2-128-7 or 1-40-3-64-7'
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== $tool: fp16 weight gradient ops"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_f16_ops.py -m gpu -x -q -k "wgrad_f16 and ($WG)" 2>&1 | tail -8
  for pm in 0 1; do
    echo "=== $tool: model path (NEF_TC_PERSIST_MIN=$pm)"
    NEF_TC_PERSIST_MIN=$pm compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -8
  done
done
