#!/bin/bash
# round-2 eight-GPU batch: weak / strong scaling of every BASELINE config, single-process sharded API, reference DataParallel
O=gpurun_out/r2e; mkdir -p $O
nvidia-smi -L | wc -l
PORT=29520
tr() { n=$1; shift; PORT=$((PORT+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port $PORT --nproc-per-node $n bench.py --gpus $n --no-parity --no-cpu-baseline "$@" 2>> $O/torchrun.err | tail -1; }
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ", sys.argv[1].split("/")[-1], "rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d.get("e2e", {}).get("value", 0)))
except Exception as e:
    print("   ", sys.argv[1], "ERR", e)
PY
}
echo "== C2 weak"; for n in 1 4 8; do if [ $n -eq 1 ]; then python bench.py --no-parity --no-cpu-baseline 2>>$O/torchrun.err | tail -1 > $O/bench_c2_weak_${n}gpu.json; else tr $n > $O/bench_c2_weak_${n}gpu.json; fi; show $O/bench_c2_weak_${n}gpu.json; done
for w in c3 c4; do echo "== $w strong"; for n in 1 4 8; do if [ $n -eq 1 ]; then python bench.py --workload $w --scaling strong --no-parity --no-cpu-baseline 2>>$O/torchrun.err | tail -1 > $O/bench_${w}_strong_${n}gpu.json; else tr $n --workload $w --scaling strong > $O/bench_${w}_strong_${n}gpu.json; fi; show $O/bench_${w}_strong_${n}gpu.json; done; done
echo "== c3 strong N=8, stage-by-stage orchestration (PNR_RENDER_FUSED=0)"; PNR_RENDER_FUSED=0 tr 8 --workload c3 --scaling strong > $O/bench_c3_strong_8gpu_unfused.json; show $O/bench_c3_strong_8gpu_unfused.json
PNR_RENDER_FUSED=0 tr 8 --workload c4 --scaling strong > $O/bench_c4_strong_8gpu_unfused.json; show $O/bench_c4_strong_8gpu_unfused.json
echo "== single-process bind_parallel(net, 0..7)"
timeout 600 python scripts/bench_sharded.py --gpus 8 2>> $O/sharded.err | tail -1 | tee $O/sharded_c2_8gpu.json | cut -c1-420
timeout 600 python scripts/bench_sharded.py --gpus 8 --workload c3 --frames 1 2>> $O/sharded.err | tail -1 | tee $O/sharded_c3_1frame_8gpu.json | cut -c1-420
timeout 600 python scripts/bench_sharded.py --gpus 8 --workload c4 --frames 1 --steps 3 2>> $O/sharded.err | tail -1 | tee $O/sharded_c4_1frame_8gpu.json | cut -c1-420
echo "== reference DataParallel"
timeout 900 python bench.py --impl reference-gpu --gpus 8 --steps 3 --warmup 1 2>> $O/refdp.err | tail -1 | tee $O/ref_gpu_c2_weak_8gpu.json | cut -c1-250
timeout 900 python bench.py --impl reference-gpu --gpus 8 --workload c3 --scaling strong --steps 3 --warmup 1 2>> $O/refdp.err | tail -1 | tee $O/ref_gpu_c3_strong_8gpu.json | cut -c1-250
timeout 900 python bench.py --impl reference-gpu --gpus 1 --workload c3 --steps 3 --warmup 1 2>> $O/refdp.err | tail -1 | tee $O/ref_gpu_c3_1gpu.json | cut -c1-250
timeout 900 python bench.py --impl reference-gpu --gpus 1 --workload c4 --steps 2 --warmup 1 2>> $O/refdp.err | tail -1 | tee $O/ref_gpu_c4_1gpu.json | cut -c1-250
tail -3 $O/torchrun.err $O/sharded.err $O/refdp.err 2>/dev/null | cut -c1-300
