#!/usr/bin/env bash
# Multi-GPU gpurun session: tools/gpu_session_multi.sh <N> <tag> [steps...]
#   steps: dptest weak weak_nofused strong sample_strong sample_weak launcher
set -u
N=$1; TAG=$2; shift 2
STEPS=${*:-dptest weak strong sample_strong}
OUT=gpurun_out
mkdir -p $OUT
PORT=29531
run() { # name, extra env (as VAR=val words or ""), bench args...
  local name=$1; shift
  local envs=$1; shift
  PORT=$((PORT+1))
  env $envs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $N "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$?"
}
for s in $STEPS; do
  case $s in
    dptest) timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q --timeout 900 > $OUT/${TAG}_dptest.log 2>&1; echo "dptest rc=$?" ;;
    weak) run weak "NCCL_DEBUG=INFO" --steps 10 --warmup 3 --no-cpu-baseline ;;
    weak_plain) run weak_plain "A=1" --steps 10 --warmup 3 --no-cpu-baseline --no-prof ;;
    weak_fused) run weak_fused "RNVP_DP_FUSED=1" --steps 10 --warmup 3 --no-cpu-baseline --no-prof ;;
    weak_nofused) run weak_nofused "RNVP_DP_FUSED=0" --steps 10 --warmup 3 --no-cpu-baseline --no-prof ;;
    weakenv:*) name=$(echo $s | cut -d: -f2); envs=$(echo $s | cut -d: -f3 | tr ',' ' ')
      run weak_$name "$envs" --steps 10 --warmup 3 --no-cpu-baseline --no-prof ;;
    strong) run strong "A=1" --steps 10 --warmup 3 --no-cpu-baseline --no-prof --global-batch 2048 ;;
    sample_strong) run sample_strong "A=1" --mode sample --steps 5 --warmup 3 --no-cpu-baseline --no-prof --global-batch 4096 ;;
    launcher) PORT=$((PORT+1)); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        dl-normalizing-flows_b200/train_dp.py --synthetic 640 --batch-size 64 --image-size 32 --base-dim 8 --res-blocks 1 --epochs 2 \
        --output-dir $OUT/${TAG}_launcher_out > $OUT/${TAG}_launcher.log 2>&1; echo "launcher rc=$?" ;;
    *) echo "unknown step $s" ;;
  esac
done
tail -3 $OUT/${TAG}_*.log 2>/dev/null | tail -30
for f in $OUT/${TAG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("dp_check"), d["config"].get("bn_stat_exchange"))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
