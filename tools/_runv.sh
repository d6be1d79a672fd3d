timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 100 python tools/_stats.py 2>&1 | tail -5
for nt in 1024 512; do echo "NT $nt"; ICNV_CELL_NT=$nt timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline'])"; done
