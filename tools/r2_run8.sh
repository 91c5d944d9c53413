AB_WORKLOADS="cfg2:128" timeout 600 tools/r2_ab.sh 8 cur unroll 2>&1 | tail -3
timeout 200 python bench.py --restart-interval 0 --frames 32 --steps 2 --warmup 3 --no-cpu --no-e2e --also none 2>gpurun_out/bench_ri0.err | tail -1 > gpurun_out/bench_ri0.json
python -c "
import json; d=json.load(open('gpurun_out/bench_ri0.json')); print('Ri=0: value', round(d['value']), 'enc', round(d['encode_mpix_s']), 'dec', round(d['decode_mpix_s']), 'ms', round(d['ms_per_step'],1))"
tail -3 gpurun_out/bench_ri0.err
