timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python - <<'PY'
import sys; sys.path.insert(0, "scripts"); sys.argv=["x"]
import quick_bench as q
q.run("connect_four", 1024, 128, 256)
q.run("go_9x9", 1024, 800, 1600, moves=2, warm=1, pdl=False)
q.run("othello", 512, 200, 400, weighted=True, moves=4, pdl=True)
PY
for i in 1 2; do timeout 300 python bench.py --skip-cpu --skip-e2e --skip-roofline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value']/1e6, d['ms_per_step'])"; done
