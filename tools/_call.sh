mkdir -p gpurun_out
timeout 900 python tools/gemm_variants.py run > gpurun_out/r02l_gemm_variants.jsonl 2> gpurun_out/r02l_gemm_variants.err; echo variants rc=$?
python - <<'PY'
import json
for line in open('gpurun_out/r02l_gemm_variants.jsonl'):
    try: r=json.loads(line)
    except Exception: print('BAD', line[:200]); continue
    c=r['classes']
    print(r['variant'], r['parity'][:80], {k:(round(v['tflops'],2)) for k,v in c.items()} if isinstance(c,dict) else c[-400:])
PY
