set -x
timeout 900 python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench24.json')); print('final', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['roofline']['kernels_ms'], d.get('e2e_plugin',{}).get('value'), d.get('e2e_plugin_n2',{}).get('value'), d['parity_checked'], d['parity'], d['cpu_baseline']['value'], d['ctb64'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hevc_recon -s 1 -c 1 -o gpurun_out/r2g_rec python scripts/decode_probe_n.py 16 16 > gpurun_out/r2g_ncu_rec.log 2>&1; tail -2 gpurun_out/r2g_ncu_rec.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ctb64 --no-plugin-leg > gpurun_out/r2g_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2g_launches.csv | cut -c1-150
