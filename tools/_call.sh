mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider -k "gemm or blocks_match or cfg1 or plan or hermitian" > gpurun_out/r02i_pytest.log 2>&1; echo pytest rc=$?; tail -n 5 gpurun_out/r02i_pytest.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02i_bench_cfg4_short.json 2> gpurun_out/r02i_bench_cfg4_short.err; echo cfg4 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02i_bench_cfg4_short.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step']); print(json.dumps({k:(round(v['tflops'],2), round(v['frac'],3)) for k,v in d['kernel_classes'].items()})); print(d['dimer_phase'])"
for w in cfg1 cfg2 herm100; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_bench_$w.json 2> gpurun_out/r02i_bench_$w.err; echo $w rc=$?; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r02i_bench_$w.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['graph_streams'], d['launches_per_call'], d['host_seconds_per_call'])"; done
