mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zz_smoke.log 2>&1; echo smoke rc=$?; tail -n 1 gpurun_out/r02zz_smoke.log
timeout 400 python -m pytest tests -m gpu -q --maxfail=5 -p no:cacheprovider > gpurun_out/r02zz_pytest_gpu.log 2>&1; echo pytest rc=$?; tail -n 3 gpurun_out/r02zz_pytest_gpu.log | cut -c1-200
