#!/bin/bash
# round-2 validation batch (run on the GPU box through scripts/gpu.sh)
O=gpurun_out/r2b; mkdir -p $O
run() { name=$1; shift; ( "$@" > $O/$name.log 2>&1; echo "rc=$?" >> $O/$name.log ); tail -4 $O/$name.log | cut -c1-300; }
echo "== gemm"; run gemm timeout 600 python -m pytest tests/test_gpu_backward.py -k gemm_nt -q --tb=short
grep -q "rc=0" $O/gemm.log || { export PNR_BWD_GEMM=simt; echo FALLBACK_SIMT_GEMM; }
echo "== fused render parity (tc tests)"; run fused_tc timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short
grep -q "rc=0" $O/fused_tc.log || { echo "== same with PNR_RENDER_FUSED=0"; PNR_RENDER_FUSED=0 run unfused_tc timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q --tb=short; grep -q "rc=0" $O/unfused_tc.log && { export PNR_RENDER_FUSED=0; echo FALLBACK_UNFUSED_RENDER; }; }
echo "== full suite"; run pytest_gpu timeout 1800 python -m pytest tests -m gpu -q --tb=short
grep -E "passed|failed" $O/pytest_gpu.log | tail -3
echo "== train"; for m in render torch; do python scripts/bench_train.py --mode $m 2> $O/train_$m.err | tail -1 | tee $O/train_$m.json; done
PNR_BWD_GEMM=simt python scripts/bench_train.py --mode render 2>/dev/null | tail -1 | tee $O/train_render_simt.json
echo "== bench c2 fused / unfused"
python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
PNR_RENDER_FUSED=0 python bench.py --no-cpu-baseline --no-parity > $O/bench_c2_unfused.json 2> $O/bench_c2_unfused.err
python bench.py --workload c3 --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
PNR_RENDER_FUSED=0 python bench.py --workload c3 --no-cpu-baseline --no-parity > $O/bench_c3_unfused.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"], 3), d.get("parity"))
    except Exception as e:
        print(f, "ERR", e)
PY
