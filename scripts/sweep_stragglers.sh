for T in 0 1 2 3 4 5; do
  IZPI_NODE_STRAGGLERS=$T python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('T=$T', round(d['value'],1), round(d['roofline']['frac'],3))"
done
