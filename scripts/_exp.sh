python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-atomic-peak 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e'], d['c2_isolate']['ms_per_graph'])"
