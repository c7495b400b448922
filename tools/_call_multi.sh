mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 600 python -m pytest tests/test_zz_distributed_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02q_pytest_nccl.log 2>&1; echo pytest nccl rc=$?; tail -n 4 gpurun_out/r02q_pytest_nccl.log | cut -c1-300
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 8 --warmup 4 > gpurun_out/r02q_bench_cfg4_n8.json 2> gpurun_out/r02q_bench_cfg4_n8.err; echo n8 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02q_bench_cfg4_n8.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['build_time_s']); print({k:v for k,v in d['e2e'].items() if k!='note'}); print(d['step_ms_by_assemble_mode'], d['assemble'], d['dimer_phase'])"; tail -n 3 gpurun_out/r02q_bench_cfg4_n8.err | cut -c1-300
