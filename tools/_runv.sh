timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python tools/bench_extra.py 2>&1 | tail -1
