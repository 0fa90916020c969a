#!/usr/bin/env bash
# compute-sanitizer passes over small-shape parity tests (one GPU, through gpurun; each tool replays every kernel, so keep the
# selection small).  Two dispatches: the default one (small row spaces take conv_tc_kernel<1, generic>) and, with
# NEF_TC_PERSIST_MIN=1, the persistent kernel with its specialised epilogues (16 epilogue warps, pipelined MMA issuer) that the
# benchmarked step runs -- including dropout (test_dropout_statistics) and the Model_nefnet2 variant.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned global + shared accesses (the zero-halo layout relies on in-bounds halo reads);
# racecheck: shared-memory hazards in the re-tile passes (wgrad_tc, latent_bwd); synccheck: mbarrier / bar.sync misuse;
# initcheck: reads of never-written workspace (the plan carves one caller-owned allocation).
set -u
SEL='(golden and train_b2_g1_l128 and tf32_tc) or dropout_statistics'
SEL2='oracle_fixed_upstream and 4-2-264 and tf32_tc'
CONV='2-40-2-128-128-7-1 or 5-333-2-128-64-3-1 or 4-50-7-128-64-1-1'   # small shapes, tcgen05 implementation (last id field 1)
for tool in memcheck racecheck synccheck initcheck; do
  for pm in 0 1; do
    echo "=== $tool: model path (NEF_TC_PERSIST_MIN=$pm)"
    NEF_TC_PERSIST_MIN=$pm compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -8
  done
  echo "=== $tool: Model_nefnet2 (persistent kernels)"
  NEF_TC_PERSIST_MIN=1 compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_nefnet2.py -m gpu -x -q -k "$SEL2" 2>&1 | tail -8
  echo "=== $tool: conv ops"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_conv_ops.py -m gpu -x -q -k "$CONV" 2>&1 | tail -8
  echo "=== $tool: fp16 operand ops (forward, data gradient, weight gradient)"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_f16_ops.py -m gpu -x -q 2>&1 | tail -8
done
