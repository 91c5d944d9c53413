timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/split_frame_check.py 16384 16384 2>&1 | grep -E "split_frame_check|Error|error" | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 10 --no-cpu --also none 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_r2_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_n2.json')); e=d['e2e']
print('N2 value', round(d['value']), 'e2e', round(e['value']), 'two-part', round(e['two_part_calls_value']), 'batch', round(e['host_batch_value']), 'two-way', round(e.get('host_batch_two_way_value',0)), 'ceiling', round(e['box_copy_ceiling_mpix_s']), 'frac', round(e['frac_of_box_copy_ceiling'],3))"
tail -3 gpurun_out/bench_n2.err
