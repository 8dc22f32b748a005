# inside gpurun --gpus 2: the 2-GPU picture equals the 1-GPU one with the tail overlap on (128 tiles per GPU: more than one K0 wave), short bench line
set -x
export B200_BENCH_TILE_CACHE=/tmp/b200_tiles_shared; mkdir -p $B200_BENCH_TILE_CACHE
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2s_multi_n2_tests.log 2>&1; tail -3 gpurun_out/r2s_multi_n2_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 5 --warmup 3 --no-ctb64 --no-plugin-leg > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n2.json')); print('n2', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernels_ms'], d['parity_checked'], d['parity'])"; tail -2 gpurun_out/r2s_bench_n2.err
