#!/usr/bin/env bash
# A/B gpurun session: tools/gpu_session_ab.sh <tag> <steps...>
#   steps: tests cpl cpl_off bench bench_off sample sample_off convbench launches
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for s in "$@"; do
  t0=$(date +%s)
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/${TAG}_tests.log 2>&1 ;;
    quick) timeout 900 python -m pytest tests/test_gpu_coupling.py tests/test_gpu_flow.py tests/test_gpu_ops.py -m gpu -q -x --timeout 600 > $OUT/${TAG}_quick.log 2>&1 ;;
    cpl) timeout 600 python tools/bench_coupling.py > $OUT/${TAG}_cpl.log 2>&1 ;;
    cpl_off) RNVP_SKIP_FUSED=0 timeout 600 python tools/bench_coupling.py > $OUT/${TAG}_cpl_off.log 2>&1 ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ;;
    bench_full) timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err ;;
    bench_off) RNVP_SKIP_FUSED=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-prof > $OUT/${TAG}_bench_off.json 2> $OUT/${TAG}_bench_off.err ;;
    sample) timeout 600 python bench.py --mode sample --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_sample.json 2> $OUT/${TAG}_sample.err ;;
    sample_off) RNVP_SKIP_FUSED=0 timeout 600 python bench.py --mode sample --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-prof > $OUT/${TAG}_sample_off.json 2> $OUT/${TAG}_sample_off.err ;;
    convbench) timeout 600 python tools/bench_conv.py > $OUT/${TAG}_convbench.log 2>&1 ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6400 -c 2400 --csv --log-file $OUT/${TAG}_launches.csv \
                python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof --no-gpu-eager > $OUT/${TAG}_ncu.log 2>&1 ;;
    cplenv:*) # cplenv:<name>:<VAR=val,VAR=val>  -- bench_coupling under an environment
      name=$(echo $s | cut -d: -f2); envs=$(echo $s | cut -d: -f3 | tr ',' ' ')
      env $envs timeout 600 python tools/bench_coupling.py > $OUT/${TAG}_cpl_${name}.log 2>&1 ;;
    benchenv:*) name=$(echo $s | cut -d: -f2); envs=$(echo $s | cut -d: -f3 | tr ',' ' ')
      env $envs timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-prof > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err ;;
    targetsenv:*) name=$(echo $s | cut -d: -f2); envs=$(echo $s | cut -d: -f3 | tr ',' ' ')
      env $envs TIME=1 timeout 300 python tools/ncu_targets.py > $OUT/${TAG}_targets_${name}.log 2>&1 ;;
    hosttime) timeout 300 python tools/host_time.py > $OUT/${TAG}_hosttime.log 2>&1; B=64 timeout 300 python tools/host_time.py >> $OUT/${TAG}_hosttime.log 2>&1 ;;
    x3ops) timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x --timeout 300 -k "tf32x3" > $OUT/${TAG}_x3ops.log 2>&1 ;;
    x3rest) timeout 900 python -m pytest tests/test_gpu_coupling.py tests/test_gpu_flow.py tests/test_gpu_api.py -m gpu -q --timeout 600 -k "tf32x3" > $OUT/${TAG}_x3rest.log 2>&1 ;;
    x3bench) timeout 900 python bench.py --math tf32x3 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_x3bench.json 2> $OUT/${TAG}_x3bench.err ;;
    refarm) timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_refarm.json 2> $OUT/${TAG}_refarm.err ;;
    c3) timeout 600 python bench.py --config c3 --batch 512 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_c3.json 2> $OUT/${TAG}_c3.err ;;
    b2048) timeout 900 python bench.py --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-prof > $OUT/${TAG}_b2048.json 2> $OUT/${TAG}_b2048.err ;;
    traffic) timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 6400 -c 2400 --csv \
               --log-file $OUT/${TAG}_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof --no-gpu-eager > $OUT/${TAG}_traffic.log 2>&1 ;;
    targets) TIME=1 timeout 300 python tools/ncu_targets.py > $OUT/${TAG}_targets.log 2>&1 ;;
    ncu_targets) timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_(fwd|wgrad)_tf32" -c 27 \
                   -o $OUT/${TAG}_targets python tools/ncu_targets.py > $OUT/${TAG}_ncu_targets.log 2>&1 ;;
    sampleenv:*) name=$(echo $s | cut -d: -f2); envs=$(echo $s | cut -d: -f3 | tr ',' ' ')
      env $envs timeout 600 python bench.py --mode sample --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-prof > $OUT/${TAG}_sample_${name}.json 2> $OUT/${TAG}_sample_${name}.err ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1 ;;
    *) echo "unknown step $s" ;;
  esac
  echo "$s rc=$? $(( $(date +%s) - t0 ))s"
done
for f in $OUT/${TAG}_*.log; do echo "== $f"; tail -12 $f; done 2>/dev/null | tail -${TAILN:-80}
for f in $OUT/${TAG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %.0f ms/step %.2f launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("gpu_launches")))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
