set -x
export B200_BENCH_TILE_CACHE=/tmp/b200_tiles_shared; mkdir -p $B200_BENCH_TILE_CACHE
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2s_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench.json')); print('final', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['roofline']['kernels_ms'], d.get('e2e_plugin',{}).get('value'), d.get('e2e_plugin_n2',{}).get('value'), d['parity_checked'], d['parity'], d['cpu_baseline']['value'], d['ctb64'], d['gpu_launches'], d['pipeline'])"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2s_bench_reference.json 2> gpurun_out/r2s_bench_reference.err; cut -c1-300 gpurun_out/r2s_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ctb64 --no-plugin-leg > gpurun_out/r2s_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2s_launches.csv | cut -c1-150
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sao_rows|hevc_recon" -s 3 -c 3 -o gpurun_out/r2s_sao_rec python scripts/decode_probe_n.py 16 16 > gpurun_out/r2s_ncu_sao_rec.log 2>&1; tail -2 gpurun_out/r2s_ncu_sao_rec.log
ls -la gpurun_out/*.ncu-rep
