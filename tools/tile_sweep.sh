for v in "RHO_FUSED_TILE_BATCHES=4" "RHO_FUSED_TILE_BATCHES=6" "RHO_FUSED_TILE_BATCHES=8" "RHO_FUSED_TILE_BATCHES=11" "RHO_FUSED_TILE_BATCHES=16" "RHO_FUSED_SCHED=13,13,6" "RHO_FUSED_SCHED=11,11,6,4" "RHO_FUSED_SCHED=10,8,6,4,4" "RHO_FUSED_SCHED=16,8,4,2,2" "X=1"; do
  env $v python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu-baseline --configs "" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), round(d['kernels']['k_fused_features']['ms_per_launch'],4))"
done
