timeout 400 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|re-run|Error|error" | head -20
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_hmm']['ms_per_launch'])"
