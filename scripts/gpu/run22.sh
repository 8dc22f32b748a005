set -x
timeout 1100 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_hevc_gpu.py -m gpu -x -q -k "x_444_basic or x_422_basic or x_422_pcm or x_444_pcm or x_422_tiles or tiles_2x2 or tiles_3x3 or pcm_nolf_wpp or bypass_mixed or grid_of_444 or scaling_sps_ctb64" > gpurun_out/r2_memcheck.log 2>&1; tail -12 gpurun_out/r2_memcheck.log; grep -c "Invalid\|out of bounds" gpurun_out/r2_memcheck.log
timeout 300 python bench.py --no-plugin-leg --no-ctb64 --steps 5 > gpurun_out/r2_bench22.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench22.json')); print('final2', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['roofline']['kernels_ms'], d['parity_checked'])"
