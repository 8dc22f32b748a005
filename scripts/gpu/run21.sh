set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t21.log 2>&1; tail -3 gpurun_out/r2_t21.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench21.json')); print('final', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_pipelined']['ms_per_step'], d['pipeline'], d.get('e2e_plugin'), d.get('e2e_plugin_n2'), d['parity_checked'], d['cpu_baseline']['value'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hevc_entropy -s 1 -c 1 -o gpurun_out/r2f_ent python scripts/decode_probe_n.py 16 16 > gpurun_out/r2f_ncu_ent.log 2>&1; tail -2 gpurun_out/r2f_ncu_ent.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:hevc_recon -s 1 -c 1 -o gpurun_out/r2f_rec python scripts/decode_probe_n.py 16 16 > gpurun_out/r2f_ncu_rec.log 2>&1; tail -2 gpurun_out/r2f_ncu_rec.log
timeout 400 ncu --set full --clock-control none -k "regex:deblock|sao_kernel|k6_" -s 4 -c 4 -o gpurun_out/r2f_filt python scripts/decode_probe_n.py 16 16 > gpurun_out/r2f_ncu_filt.log 2>&1; tail -2 gpurun_out/r2f_ncu_filt.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ctb64 --no-plugin-leg > gpurun_out/r2f_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2f_launches.csv | cut -c1-200
ls -la gpurun_out/r2f_*
