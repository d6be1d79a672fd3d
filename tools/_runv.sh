timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['fused_call']['ms_per_step'])"
