timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d.get('roofline_hmm'))"
