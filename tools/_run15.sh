cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r04q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r04q_tests.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:"stem_tc|latent_fwd" --clock-control none -c 6 --csv --log-file gpurun_out/r04q_stem.csv python tools/prof_step.py --steps 1 > /dev/null 2>&1
