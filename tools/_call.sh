mkdir -p gpurun_out
for st in 1 4 16 32; do timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --graph-streams $st > gpurun_out/r02m_cfg1_streams$st.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02m_cfg1_streams$st.json') if l.startswith('{')][-1]); print('cfg1 streams', $st, d['graph_streams'], d['ms_per_step'])"; done
for st in 16 32; do timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --graph-streams $st > gpurun_out/r02m_cfg2_streams$st.json 2>/dev/null; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02m_cfg2_streams$st.json') if l.startswith('{')][-1]); print('cfg2 streams', $st, d['graph_streams'], d['ms_per_step'])"; done
