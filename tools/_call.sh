mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; echo smoke rc=$?; tail -n 3 gpurun_out/r02g_smoke.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02g_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 12 gpurun_out/r02g_pytest_gpu.log
timeout 400 python bench.py --workload cfg1 --steps 10 --warmup 3 > gpurun_out/r02g_bench_cfg1.json 2> gpurun_out/r02g_bench_cfg1.err; echo cfg1 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02g_bench_cfg1.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['graph_streams'], d['host_seconds_per_call'], d['e2e']['seconds_per_step'])"; tail -n 3 gpurun_out/r02g_bench_cfg1.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_cfg2.json 2> gpurun_out/r02g_bench_cfg2.err; echo cfg2 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02g_bench_cfg2.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['graph_streams'], d['host_seconds_per_call'])"
timeout 300 python bench.py --workload herm100 --steps 5 --warmup 3 > gpurun_out/r02g_bench_herm100.json 2> gpurun_out/r02g_bench_herm100.err; echo herm100 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02g_bench_herm100.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['graph_streams'], d['host_seconds_per_call'])"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma_stream -c 1 -o gpurun_out/r02g_stream python tools/ncu_stream_kernel.py > gpurun_out/r02g_ncu_stream.log 2>&1; echo ncu rc=$?; tail -n 3 gpurun_out/r02g_ncu_stream.log
timeout 200 python tools/ncu_stream_kernel.py
timeout 400 python bench.py --steps 4 --warmup 3 > gpurun_out/r02g_bench_cfg4_n1.json 2> gpurun_out/r02g_bench_cfg4_n1.err; echo cfg4 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02g_bench_cfg4_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps({k:(round(v['tflops'],2), round(v['frac'],3)) for k,v in d['kernel_classes'].items()})); print(d['dimer_phase'])"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-600
