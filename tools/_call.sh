mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02d_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 25 gpurun_out/r02d_pytest_gpu.log
timeout 300 python tools/profile_plan.py cfg1 30 > gpurun_out/r02d_profile_plan_cfg1.json 2> gpurun_out/r02d_profile_plan.err; echo profile rc=$?; cut -c1-2500 gpurun_out/r02d_profile_plan_cfg1.json; tail -n 3 gpurun_out/r02d_profile_plan.err
timeout 400 python bench.py --workload cfg1 --steps 10 --warmup 3 > gpurun_out/r02d_bench_cfg1.json 2> gpurun_out/r02d_bench_cfg1.err; echo cfg1 rc=$?; tail -c 1500 gpurun_out/r02d_bench_cfg1.json; tail -n 5 gpurun_out/r02d_bench_cfg1.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/r02d_bench_cfg2.json 2> gpurun_out/r02d_bench_cfg2.err; echo cfg2 rc=$?; tail -c 600 gpurun_out/r02d_bench_cfg2.json
timeout 400 python bench.py --workload herm100 --steps 5 --warmup 3 > gpurun_out/r02d_bench_herm100.json 2> gpurun_out/r02d_bench_herm100.err; echo herm100 rc=$?; tail -c 1500 gpurun_out/r02d_bench_herm100.json; tail -n 5 gpurun_out/r02d_bench_herm100.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02d_bench_cfg4_short.json 2> gpurun_out/r02d_bench_cfg4_short.err; echo cfg4 rc=$?; python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_cfg4_short.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['kernel_classes'])[:1500]); print(d.get('dimer_phase'))"
