#!/bin/bash
# last round-2 GPU call: validate the final library (split GEMM variants, masked epilogue, fp16 projection, polite
# ready-flag polling) and decide fused vs stage-by-stage at small ray counts
O=gpurun_out/r2f; mkdir -p $O
( timeout 100 python -m pytest tests/test_gpu_tc.py tests/test_gpu_backward.py -m gpu -q --tb=line -x > $O/tests.log 2>&1; echo rc=$? >> $O/tests.log ); tail -3 $O/tests.log | cut -c1-200
for f in 1 0; do PNR_RENDER_FUSED=$f timeout 60 python bench.py --workload c3 --rays 512 --steps 20 --no-parity --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_c3_512rays_fused$f.json; python -c "import json;d=json.loads(open('$O/bench_c3_512rays_fused$f.json').read());print('c3 512 rays fused=$f', round(d['value']), 'ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches'])"; done
timeout 50 python scripts/bench_train.py --mode render 2>/dev/null | tail -1 | tee $O/train_render.json | cut -c1-200
timeout 70 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > $O/bench_c2.json; python -c "import json;d=json.loads(open('$O/bench_c2.json').read());print('c2', round(d['value']), d['roofline']['frac'], d['parity'])"
