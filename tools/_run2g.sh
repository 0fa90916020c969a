cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r04_bench_2gpu.json 2> gpurun_out/r04_bench_2gpu.err; echo "rc=$?" >> gpurun_out/r04_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r04_bench_2gpu_ref.json 2> gpurun_out/r04_bench_2gpu_ref.err; echo "rc=$?" >> gpurun_out/r04_bench_2gpu_ref.err
timeout 200 python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r04_bench_1gpu_same_box_as_2gpu.json 2>/dev/null
