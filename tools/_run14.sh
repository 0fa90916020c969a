cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r04p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r04p_tests.log
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:stem_tc --clock-control none -c 4 --csv --log-file gpurun_out/r04p_stem.csv python tools/prof_step.py --steps 1 > /dev/null 2>&1
python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r04p_bench.json 2>/dev/null
