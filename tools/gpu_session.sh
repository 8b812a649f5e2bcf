#!/usr/bin/env bash
# One gpurun session: per-op tests first (localises kernel bugs), then the rest of the GPU suite, the parity
# diagnostic and the bench.  Everything is wrapped in `timeout` so that a hung kernel cannot eat the box.
# usage: tools/gpu_session.sh <tag> [steps...]   steps: ops rest diag bench bench_noxf
set -u
TAG=${1:-s}; shift || true
STEPS=${*:-ops rest diag bench}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
for s in $STEPS; do
  case $s in
    ops)   timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 > $OUT/${TAG}_ops.log 2>&1; echo "ops rc=$?" ;;
    rest)  timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_ops.py > $OUT/${TAG}_rest.log 2>&1; echo "rest rc=$?" ;;
    rest_nocpl) RNVP_CPL_EPILOGUE=0 timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_ops.py > $OUT/${TAG}_rest_nocpl.log 2>&1; echo "rest_nocpl rc=$?" ;;
    cpl)   timeout 1200 python -m pytest tests/test_gpu_coupling.py "tests/test_gpu_flow.py::test_golden_model" tests/test_gpu_flow.py::test_cfg_a_against_oracle -m gpu -q --timeout 900 > $OUT/${TAG}_cpl.log 2>&1; echo "cpl rc=$?" ;;
    bench_nocpl) RNVP_CPL_EPILOGUE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_bench_nocpl.json 2> $OUT/${TAG}_bench_nocpl.err; echo "bench_nocpl rc=$?" ;;
    dbg)   timeout 600 python tools/debug_conv_bn.py > $OUT/${TAG}_dbg.log 2>&1; echo "dbg rc=$?" ;;
    diag)  timeout 1200 python tools/diag_precision.py > $OUT/${TAG}_diag.log 2>&1; echo "diag rc=$?" ;;
    diagbig) timeout 1500 python tools/diag_precision.py --big > $OUT/${TAG}_diagbig.log 2>&1; echo "diagbig rc=$?" ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?" ;;
    bench_noxf) RNVP_XFORM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_bench_noxf.json 2> $OUT/${TAG}_bench_noxf.err; echo "bench_noxf rc=$?" ;;
    sample) timeout 600 python bench.py --mode sample --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_sample.json 2> $OUT/${TAG}_sample.err; echo "sample rc=$?" ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    *) echo "unknown step $s" ;;
  esac
done
tail -5 $OUT/${TAG}_*.log 2>/dev/null | tail -60
