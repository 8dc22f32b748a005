set -x
timeout 300 python -m pytest tests/test_hevc_gpu.py -m gpu -x -q -k chunked > gpurun_out/r2_t17a.log 2>&1; tail -5 gpurun_out/r2_t17a.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t17.log 2>&1; tail -3 gpurun_out/r2_t17.log
timeout 900 python bench.py > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; cat gpurun_out/r2_bench17.json
(B200_CHUNKS=0 B200_ENTROPY_BLOCKS_PER_SM=3 timeout 200 python scripts/decode_probe_n.py 16 16
 B200_CHUNKS=0 timeout 200 python scripts/decode_probe_n.py 16 16
 timeout 200 python scripts/decode_probe_n.py 16 16) > gpurun_out/r2_probe17.log 2>&1; cat gpurun_out/r2_probe17.log
for t in 32 128; do B200_CHUNK_TILES=$t timeout 600 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench17_t$t.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench17_t$t.json')); print($t, d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'])"; done
B200_CHUNK_K0_BLOCKS=4 timeout 600 python bench.py --no-plugin-leg --no-ctb64 --no-cpu-baseline --steps 5 > gpurun_out/r2_bench17_k4.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench17_k4.json')); print('k0blocks4', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'])"
