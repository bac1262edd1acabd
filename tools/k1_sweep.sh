run() { python bench.py --steps 20 --warmup 3 --no-fit --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])" 2>&1 | tail -1; }
echo "base $(run)"
echo "base $(run)"
