set -x
timeout 900 python bench.py > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench19.json')); print('default', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'], d.get('e2e_plugin'), d.get('e2e_plugin_n2'), d.get('ctb64'), d['cpu_baseline'])"
B200_CHUNKS=0 timeout 600 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench19_nochunk.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench19_nochunk.json')); print('nochunk', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'])"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_t19_multi.log 2>&1; tail -3 gpurun_out/r2_t19_multi.log
timeout 300 python scripts/config5_bench.py > gpurun_out/r2_config5_n1.json 2> gpurun_out/r2_config5_n1.err; cat gpurun_out/r2_config5_n1.json
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_hevc_gpu.py -m gpu -x -q -k "chunked or heterogeneous or sequential_and_concurrent" > gpurun_out/r2_racecheck.log 2>&1; tail -15 gpurun_out/r2_racecheck.log
