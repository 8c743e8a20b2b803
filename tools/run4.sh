for c in 0 1 2 3; do LLACUDA_GEMM_CFG=$c python tools/gemm_sweep.py 2>&1 | tail -6; done
python bench.py --profile --steps 1 --warmup 0 > gpurun_out/plain_profile.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^lla_ -c 40000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
wc -l gpurun_out/r01_launches.csv
