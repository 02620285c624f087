#!/bin/bash
# round-2 profiling batch (run on the GPU box through scripts/gpu.sh)
O=gpurun_out/r2c; mkdir -p $O
echo "== backward tests (tc recompute + bigger chunks)"
( timeout 900 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dropin_scripts.py -m gpu -q --tb=short > $O/bwd.log 2>&1; echo rc=$? >> $O/bwd.log ); tail -4 $O/bwd.log
echo "== train bench"
for m in render torch reference; do python scripts/bench_train.py --mode $m 2> $O/train_$m.err | tail -1 | tee $O/train_$m.json; done
echo "== train step launch list (ncu, serialised; shares only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train_launches.csv python scripts/bench_train.py --mode render --steps 1 --warmup 1 > $O/train_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2c/train_launches.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; kn = h.index("Kernel Name"); mv = h.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
# only the last step: take the second half of the launches
body = [r for r in rows[hdr + 1:] if len(r) > mv]
body = body[len(body) // 2:]
for r in body:
    name = r[kn].split("(")[0][-60:]
    try: t = float(r[mv].replace(",", ""))
    except ValueError: continue
    tot[name] += t; cnt[name] += 1
s = sum(tot.values())
print("last step: %d launches, %.2f ms serialised" % (len(body), s / 1e6))
for k, v in tot.most_common(18): print("  %6.2f%% %8.3f ms x%-4d %s" % (100 * v / s, v / 1e6, cnt[k], k))
PY
echo "== bench launch list c2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $O/launch_c2.log 2>&1
echo "== ncu full: fused render kernel, c2 (1 launch) and c4 (1 launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 3 -c 1 -o $O/r2_prof_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity > $O/ncu_c2.log 2>&1; tail -2 $O/ncu_c2.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_field_tc -s 1 -c 1 -o $O/r2_prof_c4 -f python bench.py --workload c4 --rays 30000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity > $O/ncu_c4.log 2>&1; tail -2 $O/ncu_c4.log
timeout 900 ncu --set full --clock-control none -k regex:k_gemm_bf16x3 -s 200 -c 6 -o $O/r2_prof_gemm -f python scripts/bench_train.py --mode render --steps 1 --warmup 1 > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log
ls -la $O
