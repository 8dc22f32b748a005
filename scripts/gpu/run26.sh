set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
export B200_BENCH_TILE_CACHE=/tmp/b200_tiles_shared; mkdir -p $B200_BENCH_TILE_CACHE
for v in "" _w1 _w2 _w8; do
  B200_LIB=$PWD/libheif_b200/libb200heif$v.so timeout 120 python scripts/k0_variant_probe.py 16 4 2>&1 | tail -1
done
B200_TAIL_OVERLAP=0 timeout 120 python scripts/k0_variant_probe.py 16 4 2>&1 | tail -1
timeout 240 python scripts/pipe_probe.py 16 8 > gpurun_out/pipe_probe.json 2> gpurun_out/pipe_probe.err; cat gpurun_out/pipe_probe.err | cut -c1-300
