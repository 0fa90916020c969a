#!/usr/bin/env bash
# compute-sanitizer passes over the small-shape parity tests (one GPU, through gpurun; each tool replays every kernel, so
# keep the selection small).  NOT YET RUN: round 1 ended without GPU budget for it -- first thing to do in round 2.
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# memcheck: out-of-bounds / misaligned global + shared accesses (the zero-halo layout relies on in-bounds halo reads);
# racecheck: shared-memory hazards in the re-tile passes (wgrad_tc, latent_bwd); synccheck: mbarrier / bar.sync misuse;
# initcheck: reads of never-written workspace (the plan carves one caller-owned allocation).
set -u
SEL='golden and train_b2_g1_l128 and tf32_tc'
CONV='2-40-2-128-128-7-1 or 5-333-2-128-64-3-1 or 4-50-7-128-64-1-1'   # small shapes, tcgen05 implementation (last id field 1)
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== $tool: model path"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -25
  echo "=== $tool: conv ops"
  compute-sanitizer --tool "$tool" --error-exitcode 9 python -m pytest tests/test_gpu_conv_ops.py -m gpu -x -q -k "$CONV" 2>&1 | tail -25
done
