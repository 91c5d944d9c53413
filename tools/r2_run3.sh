bash tools/r2_prof.sh r2a cfg2 128
bash tools/r2_prof.sh r2a_cfg4 cfg4 64
timeout 300 python bench.py --steps 10 --no-cpu --also none 2>gpurun_out/bench_e2e.err | tail -1 > gpurun_out/bench_e2e_t8.json
timeout 300 python bench.py --steps 10 --no-cpu --also none --e2e-threads 4 2>>gpurun_out/bench_e2e.err | tail -1 > gpurun_out/bench_e2e_t4.json
timeout 300 python bench.py --steps 10 --no-cpu --also none --e2e-threads 16 2>>gpurun_out/bench_e2e.err | tail -1 > gpurun_out/bench_e2e_t16.json
CHARLS_B200_HOST_CHUNK=8 timeout 300 python bench.py --steps 10 --no-cpu --also none --e2e-threads 2 2>>gpurun_out/bench_e2e.err | tail -1 > gpurun_out/bench_e2e_t2_chunk8.json
for f in gpurun_out/bench_e2e_*.json; do python -c "
import json,sys
d=json.load(open('$f')); e=d['e2e']; print('$f', round(d['value']), 'e2e', round(e['value']), 'threads', e['host_threads'], 'inflight', e['objects_in_flight_per_thread'], 'one_part', round(e['one_part_calls_value']), 'batch', round(e['host_batch_value']))
"; done
tail -5 gpurun_out/bench_e2e.err
