python -m pytest tests/test_factor_gpu.py -x -q 2>&1 | tail -3
python tools/gpu_probe2.py 2>&1 | grep -E "getrf"
python tools/ncu_small.py lu 1024 > gpurun_out/plain_lu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^lla_ --csv --log-file gpurun_out/small_lu.csv python tools/ncu_small.py lu 1024 > gpurun_out/ncu_small_lu.log 2>&1; echo "rc=$?"
