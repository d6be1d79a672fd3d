python -m pytest tests -m gpu -q 2>&1 | tail -3
for v in 2 3; do echo "variant $v"; ICNV_CELL_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_smooth']['ms_per_step'])"; done
