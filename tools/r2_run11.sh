timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --restart-interval 0 --frames 32 --steps 2 --warmup 3 --no-cpu --no-e2e --also none 2>gpurun_out/bench_ri0.err | tail -1 > gpurun_out/bench_ri0_b.json
python -c "
import json; d=json.load(open('gpurun_out/bench_ri0_b.json')); print('Ri=0 32 frames: value', round(d['value']), 'enc', round(d['encode_mpix_s']), 'dec', round(d['decode_mpix_s']), 'ms', round(d['ms_per_step'],1))"
timeout 300 python bench.py --restart-interval 0 --frames 256 --steps 1 --warmup 3 --no-cpu --no-e2e --also none 2>>gpurun_out/bench_ri0.err | tail -1 > gpurun_out/bench_ri0_c.json
python -c "
import json; d=json.load(open('gpurun_out/bench_ri0_c.json')); print('Ri=0 256 frames: value', round(d['value']), 'enc', round(d['encode_mpix_s']), 'dec', round(d['decode_mpix_s']), 'ms', round(d['ms_per_step'],1))"
tail -3 gpurun_out/bench_ri0.err
