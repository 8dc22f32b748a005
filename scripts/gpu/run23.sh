set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t23.log 2>&1; tail -3 gpurun_out/r2_t23.log
timeout 400 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench23.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench23.json')); print('dp2a', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['roofline']['kernels_ms'])"
