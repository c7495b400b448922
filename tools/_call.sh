mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; echo smoke rc=$?; tail -n 2 gpurun_out/r02n_smoke.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02n_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 5 gpurun_out/r02n_pytest_gpu.log
for w in cfg1 cfg2 herm100; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02n_bench_$w.json 2> gpurun_out/r02n_bench_$w.err; echo $w rc=$?; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r02n_bench_$w.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['graph_streams'], d['roofline']['frac'], d['host_seconds_per_call'], d['e2e']['seconds_per_step'], d['cpu_baseline']['seconds_per_call'])"; done
timeout 300 python tools/profile_plan.py cfg1 12 > gpurun_out/r02n_profile_plan_cfg1.json 2> gpurun_out/r02n_profile_plan.err; python -c "
import json; d=json.load(open('gpurun_out/r02n_profile_plan_cfg1.json')); print('total', d['total_ms_between_events'], 'critical path', d['critical_path_ms'], 'edges', d['dependency_edges'], d['by_kind'])"
