set -x
timeout 600 python -m pytest tests/test_hevc_gpu.py -m gpu -q -k "422 or 444" > gpurun_out/r2_t20a.log 2>&1; tail -40 gpurun_out/r2_t20a.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_hevc_gpu.py::test_422_and_444_coded_pictures_match_oracle --deselect tests/test_hevc_gpu.py::test_422_and_444_to_rgb_through_the_fused_entry_point --deselect tests/test_hevc_gpu.py::test_grid_of_444_tiles > gpurun_out/r2_t20.log 2>&1; tail -4 gpurun_out/r2_t20.log
timeout 600 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench20.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['roofline']['kernels_ms'])"
