python scripts/render_one.py --config 3 --spp 256 --repeat 3 2>&1 | tail -1
python bench.py --no-cpu-baseline --render-configs 3 --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for e in d['render']: print(e['config'], e['spp'], e['msamples_per_s'], e['seconds'], e['frame_ms_max_over_ranks'])
"
python bench.py --no-cpu-baseline --render-configs 1,3 --no-render-stats --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for e in d['render']: print(e['config'], e['spp'], e['msamples_per_s'], e['seconds'], e['frame_ms_max_over_ranks'])
"
