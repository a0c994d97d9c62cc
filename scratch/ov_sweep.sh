for cfg in "1 24" "4 24" "4 32"; do
  set -- $cfg
  for mc in 8 32; do
  r=$(CUDA_DEVICE_MAX_CONNECTIONS=$mc SC_OV_Q=$1 SC_OV_R=$2 timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-gl --no-sweep --no-e2e --no-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step']*1000,1), [round(v['ms']*1000,1) for v in d['roofline']['kernels'].values()], round(d['frontend_fp32_mode']['ms_per_step']*1000,1))")
  echo "Q=$1 R=$2 conn=$mc -> $r"
  done
done
