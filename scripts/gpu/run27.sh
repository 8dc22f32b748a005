set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
export B200_BENCH_TILE_CACHE=/tmp/b200_tiles_shared; mkdir -p $B200_BENCH_TILE_CACHE
timeout 120 python scripts/k0_variant_probe.py 16 4 2>&1 | tail -1
B200_SAO_ROW_PER_THREAD=1 timeout 120 python scripts/k0_variant_probe.py 16 4 2>&1 | tail -1
timeout 120 python scripts/k0_variant_probe.py 16 4 2>&1 | tail -1
