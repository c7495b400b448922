# usage: bash tools/_call_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1
run() { # name, nproc, extra args
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $2 ${@:3} > gpurun_out/$1.json 2> gpurun_out/$1.err
  echo "$1 rc=$?"; tail -c 1800 gpurun_out/$1.json; tail -n 4 gpurun_out/$1.err | cut -c1-400
}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_zz_distributed_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r02d_pytest_nccl.log 2>&1; echo pytest nccl rc=$?; tail -n 8 gpurun_out/r02d_pytest_nccl.log
  run r02d_bench_cfg4_n2_nccl 2 --steps 4 --warmup 3
  run r02d_bench_cfg4_n2_ce 2 --steps 4 --warmup 3 --assemble ce --no-e2e
fi
if [ "$N" = "8" ]; then
  run r02f_bench_cfg4_n8 8 --steps 8 --warmup 4
  run r02f_bench_cfg4_n4 4 --steps 4 --warmup 3 --no-e2e
  run r02f_bench_cfg4_n8_nccl 8 --steps 4 --warmup 3 --assemble nccl --no-e2e
  run r02f_bench_cfg5_n8 8 --workload cfg5 --steps 2 --warmup 1
fi
