mkdir -p gpurun_out
( time timeout 800 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02k_reference_n1.json 2> gpurun_out/r02k_reference_n1.err ) 2>&1 | tail -n 3; cut -c1-400 gpurun_out/r02k_reference_n1.json
( time timeout 850 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02k_bench_n1.json 2> gpurun_out/r02k_bench_n1.err ) 2>&1 | tail -n 3; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02k_bench_n1.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['build_time_s'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])"; tail -n 3 gpurun_out/r02k_bench_n1.err
timeout 200 python3 bench.py --workload cfg3 --steps 5 --warmup 3 | cut -c1-700
