mkdir -p gpurun_out
timeout 600 python tools/gemm_variants.py run > gpurun_out/r02v_gemm_wholek.jsonl 2> gpurun_out/r02v_gemm_wholek.err; echo rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02v_gemm_wholek.jsonl'):
    try: r=json.loads(line)
    except Exception: print('BAD', line[:200]); continue
    c=r['classes']
    print(r['variant'], r['parity'][-300:].replace('\n',' '), {k:(round(v['ms'],3), round(v['tflops'],2)) for k,v in c.items()} if isinstance(c,dict) else c[-600:])
PY
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "gemm or hermitian or plan or blocks or recorded" > gpurun_out/r02v_pytest.log 2>&1; echo pytest rc=$?; tail -n 5 gpurun_out/r02v_pytest.log | cut -c1-300
for w in cfg1 cfg2; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02v_bench_$w.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02v_bench_$w.json') if l.startswith('{')][-1]); print('$w', d['ms_per_step'], d['graph_streams'])"; done
