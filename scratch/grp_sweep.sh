for q in 1 2 4 6 8 12 16; do
  r=$(SC_FE_GROUPS=$q timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-gl --no-sweep --no-e2e --no-probe --no-selfcheck 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1000,1), 'frac', round(d['roofline']['frac'],4))")
  echo "groups=$q -> $r"
done
SC_FE_GROUPS=4 python -m pytest tests/test_gpu_frontend.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -2
