# usage: bash tools/_call_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
timeout 900 python -m pytest tests/test_zz_distributed_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r02o_pytest_nccl.log 2>&1; echo pytest nccl rc=$?; tail -n 12 gpurun_out/r02o_pytest_nccl.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_cfg1_n2.json 2> gpurun_out/r02o_bench_cfg1_n2.err; echo cfg1 n2 rc=$?; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02o_bench_cfg1_n2.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['scaling'], d['e2e']['seconds_per_step'])"; tail -n 3 gpurun_out/r02o_bench_cfg1_n2.err | cut -c1-300
