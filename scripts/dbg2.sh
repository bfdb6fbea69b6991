for d in 0 8 15; do echo "DBG=$d"; MI_TC_DBG=$d python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --timesteps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print(r['us_gemm1'], r['us_gemm2'])"; done
