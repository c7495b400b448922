mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02r_bench_cfg4_n1.json 2> gpurun_out/r02r_bench_cfg4_n1.err; echo rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02r_bench_cfg4_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['trimer_consumers'])[:1200]); print(d['dimer_phase'])"; tail -n 5 gpurun_out/r02r_bench_cfg4_n1.err
