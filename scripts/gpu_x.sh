timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do for f in "" "--no-pdl"; do
  timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline $f 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench $f', d['value']/1e6, d['ms_per_step'])"
done; done
