# usage: bash scripts/gpu/run_multi.sh N   (inside gpurun --gpus N)
N=$1
set -x
nvidia-smi -L | head -8
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_multi_n${N}_tests.log 2>&1; tail -3 gpurun_out/r2_multi_n${N}_tests.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus $N --steps 10 --warmup 3 --no-ctb64 --no-plugin-leg > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; cat gpurun_out/r2_bench_n${N}.json | cut -c1-2500; tail -2 gpurun_out/r2_bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/config5_bench.py > gpurun_out/r2_config5_n${N}.json 2> gpurun_out/r2_config5_n${N}.err; cat gpurun_out/r2_config5_n${N}.json; tail -2 gpurun_out/r2_config5_n${N}.err
