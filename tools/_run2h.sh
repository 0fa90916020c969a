cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
NEF_DDP_OVERLAP=0 timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04j_2gpu_after.json 2> gpurun_out/r04j_2gpu_after.err
NEF_DDP_OVERLAP=1 timeout 300 $TR --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04j_2gpu_overlap.json 2> gpurun_out/r04j_2gpu_overlap.err
NEF_DDP_OVERLAP=0 timeout 300 $TR --master-port 29543 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r04j_2gpu_after2.json 2> gpurun_out/r04j_2gpu_after2.err
