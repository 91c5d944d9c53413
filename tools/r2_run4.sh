nvidia-smi topo -m 2>&1 | head -20
lscpu | grep -E "NUMA|Model name|Socket|^CPU\(s\)" 
cat /sys/fs/cgroup/cpu.max
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 tools/pcie_probe_ranks.py 2>&1 | grep -v Warning | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 4 --steps 10 --no-cpu --also none 2>gpurun_out/bench_n4.err | tail -1 > gpurun_out/bench_n4.json
python -c "
import json
d=json.load(open('gpurun_out/bench_n4.json')); e=d['e2e']; print('N4', round(d['value']), 'e2e', round(e['value']), 'threads', e['host_threads'], 'inflight', e['objects_in_flight_per_thread'], 'one_part', round(e['one_part_calls_value']), e['one_part_calls_host_threads'], 'batch', round(e['host_batch_value']))
"
tail -3 gpurun_out/bench_n4.err
