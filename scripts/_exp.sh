for e in "AMIRA_X=1" "AMIRA_UF_ORDERED=1"; do
env $e python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-atomic-peak --no-c2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$e', round(d['ms_per_step'],3), d['phases_ms'])"
done
