timeout 100 python tools/prof_general.py 2>&1 | tail -1
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for extra in "" "--no-offset-table"; do
timeout 300 python bench.py --steps 10 --no-cpu --no-e2e --also none $extra 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg2 $extra', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3), round(d['ms_per_step'],3), d['gpu_launches'])"
done
timeout 300 python bench.py --steps 10 --no-cpu --no-e2e --also none --workload cfg4 --frames 128 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 128 frames', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3), round(d['ms_per_step'],3))"
timeout 300 python bench.py --steps 10 --no-cpu --no-e2e --also none --workload cfg3 --frames 128 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 128 frames', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3), round(d['ms_per_step'],3))"
