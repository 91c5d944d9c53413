( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_default.json')); e=d['e2e']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'roof', round(d['roofline']['frac'],4), d['roofline']['kernel'], round(d['roofline']['ms_per_launch'],3), 'other', round(d['roofline']['other_kernel_ms_per_launch'],3))
print('e2e', round(e['value']), 'two-part', round(e['two_part_calls_value']), 'batch', round(e['host_batch_value']), 'two-way', round(e.get('host_batch_two_way_value',0)), 'ceiling', round(e['box_copy_ceiling_mpix_s']), 'frac', round(e['frac_of_box_copy_ceiling'],3))
print({k:round(v) for k,v in d['config'].items() if k.endswith('mpix_s')}, d['cpu_baseline'], d['clocks'], d['gpu_launches'])"
tail -3 gpurun_out/bench_r2_default.err
