mkdir -p gpurun_out
env | grep -i nccl; echo "--- env above"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_cfg3_n2.json 2> gpurun_out/r02p_bench_cfg3_n2.err; echo rc=$?; wc -l gpurun_out/r02p_bench_cfg3_n2.json; head -c 150 gpurun_out/r02p_bench_cfg3_n2.json; echo; grep -c "NCCL" gpurun_out/r02p_bench_cfg3_n2.err
