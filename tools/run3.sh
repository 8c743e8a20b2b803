set -x
free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r01a.json 2> gpurun_out/bench_r01a.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r01a.json; tail -5 gpurun_out/bench_r01a.err
python tools/ncu_gemm.py > gpurun_out/plain_gemm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lla_dgemm_kernel -s 2 -c 2 -o gpurun_out/r01_dgemm_full -f python tools/ncu_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu full rc=$?"
python bench.py --profile --steps 1 --warmup 0 > gpurun_out/plain_profile.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^lla_ -c 40000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --profile --steps 1 --warmup 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -2 gpurun_out/plain_profile.log | cut -c1-600
ls -la gpurun_out/
