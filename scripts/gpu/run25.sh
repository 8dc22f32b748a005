set -x
timeout 600 python -m pytest tests/test_color_gpu.py -x -q -m gpu -k "rgb_to_ycbcr" 2>&1 | tail -5
timeout 300 python scripts/rgb2ycc_probe.py > gpurun_out/r2h_rgb2ycc.json 2> gpurun_out/r2h_rgb2ycc.err; cat gpurun_out/r2h_rgb2ycc.json; tail -3 gpurun_out/r2h_rgb2ycc.err
