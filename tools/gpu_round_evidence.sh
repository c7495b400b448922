#!/bin/bash
# One GPU call that refreshes the judged evidence: parity tests, ncu full capture of the trimer kernel, ncu launch list of one
# bench step (about 250 s under ncu: 170 s cut the r01r list short), then the bench line itself (never under ncu).  Usage: gpurun --timeout 760 -- 'bash tools/gpu_round_evidence.sh r01r'
# Multi-GPU follow-ups (separate calls): gpurun --gpus 2 -- 'XR_TEST_NCCL=1 python -m pytest tests/test_zz_distributed_gpu.py -m gpu -q';
#   gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_hermitian_sharded.py 0 herm100'
tag=${1:-r01x}
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest_gpu.log
timeout 100 ncu --set full --clock-control none --import-source on -k regex:trimer_stream -c 1 -f -o gpurun_out/${tag}_trimer_full python tools/ncu_kernels.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 330 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${tag}_launches_cfg4.csv python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 330 python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_bench_cfg4_n1.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/${tag}_bench_cfg4_n1.json
# optional (XR_SANITIZE=1): memcheck of the trimer launcher's kernels, incl. the multi-block first-moment kernels added in r01s
if [ "${XR_SANITIZE:-0}" = "1" ]; then
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/trimer_sweep.py 18 130 330 300 > gpurun_out/${tag}_sanitizer.txt 2>&1; echo "sanitizer rc=$?"; tail -2 gpurun_out/${tag}_sanitizer.txt
fi
