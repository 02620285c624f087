#!/bin/bash
# round-2 two-GPU batch: single-process sharded render (pnr_mgpu_*), torchrun bench at N=2, reference DataParallel
O=gpurun_out/r2d; mkdir -p $O
nvidia-smi -L | head -3
echo "== fullsize + dropin tests (incl. the 2-GPU sharded test)"
( timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dropin_scripts.py -m gpu -q --tb=short > $O/tests.log 2>&1; echo rc=$? >> $O/tests.log ); tail -6 $O/tests.log | cut -c1-300
echo "== bench_sharded 2 GPUs"
timeout 600 python scripts/bench_sharded.py --gpus 2 2> $O/sharded2.err | tail -1 | tee $O/sharded_c2_2gpu.json
timeout 600 python scripts/bench_sharded.py --gpus 1 2>> $O/sharded2.err | tail -1 | tee $O/sharded_c2_1gpu.json
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
echo "== torchrun bench N=2"
timeout 600 $TR --nproc-per-node 2 bench.py --gpus 2 --no-parity 2> $O/b2.err | tail -1 > $O/bench_c2_weak_2gpu.json; cut -c1-200 $O/bench_c2_weak_2gpu.json
timeout 600 $TR --nproc-per-node 2 bench.py --gpus 2 --workload c3 --scaling strong --no-parity 2>> $O/b2.err | tail -1 > $O/bench_c3_strong_2gpu.json; cut -c1-200 $O/bench_c3_strong_2gpu.json
echo "== reference DataParallel N=2"
timeout 900 python bench.py --impl reference-gpu --gpus 2 --steps 3 --warmup 1 2> $O/refdp.err | tail -1 | tee $O/ref_gpu_c2_2gpu.json | cut -c1-300
tail -3 $O/sharded2.err
