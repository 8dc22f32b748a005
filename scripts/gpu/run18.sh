set -x
timeout 300 python -m pytest tests/test_hevc_gpu.py -m gpu -x -q -k "chunked or single_picture" > gpurun_out/r2_t18a.log 2>&1; tail -3 gpurun_out/r2_t18a.log
(timeout 200 python scripts/decode_probe_n.py 16 16
 B200_LIB=$PWD/libheif_b200/libb200heif_inl.so timeout 200 python scripts/decode_probe_n.py 16 16
 B200_LIB=$PWD/libheif_b200/libb200heif_w6.so timeout 200 python scripts/decode_probe_n.py 16 16
 timeout 200 python scripts/decode_probe_n.py 4 4
 B200_LIB=$PWD/libheif_b200/libb200heif_inl.so timeout 200 python scripts/decode_probe_n.py 4 4
 B200_LIB=$PWD/libheif_b200/libb200heif_w6.so timeout 200 python scripts/decode_probe_n.py 4 4) > gpurun_out/r2_probe18.log 2>&1; cat gpurun_out/r2_probe18.log
timeout 900 python bench.py > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; cat gpurun_out/r2_bench18.json; tail -3 gpurun_out/r2_bench18.err
for t in 32 16; do B200_CHUNK_TILES=$t timeout 600 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench18_t$t.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench18_t$t.json')); print($t, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'])"; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t18.log 2>&1; tail -3 gpurun_out/r2_t18.log
