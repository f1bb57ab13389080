timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/matvec_probe.py h2o | tail -2
python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(d['value'], d['ms_per_step'], r['frac'], r['avg_launch_us'])"
