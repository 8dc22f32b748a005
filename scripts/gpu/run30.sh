set -x
export B200_BENCH_TILE_CACHE=/tmp/b200_tiles_shared; mkdir -p $B200_BENCH_TILE_CACHE
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2t_pytest_gpu.log
timeout 120 python bench.py --no-ctb64 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); print('final', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_parse_ms'], d['e2e_pipelined']['ms_per_step'], d.get('e2e_plugin',{}).get('value'), d.get('e2e_plugin_n2',{}), d['parity_checked'])"
