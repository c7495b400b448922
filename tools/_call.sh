mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=15 -p no:cacheprovider > gpurun_out/r02s_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 8 gpurun_out/r02s_pytest_gpu.log | cut -c1-300
timeout 300 python tools/bench_recorded_general.py cfg3 > gpurun_out/r02s_recorded_general_cfg3.json 2> gpurun_out/r02s_recorded.err; cat gpurun_out/r02s_recorded_general_cfg3.json; tail -n 3 gpurun_out/r02s_recorded.err
timeout 300 python tools/bench_recorded_general.py cfg1 > gpurun_out/r02s_recorded_general_cfg1.json 2>> gpurun_out/r02s_recorded.err; cat gpurun_out/r02s_recorded_general_cfg1.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
