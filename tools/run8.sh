python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/gpu_probe2.py 2>&1 | grep -E "getrf|gesv|potrf"
for w in chol lu; do
python tools/ncu_small.py $w 1024 > gpurun_out/plain_$w.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^lla_ --csv --log-file gpurun_out/small_$w.csv python tools/ncu_small.py $w 1024 > gpurun_out/ncu_small_$w.log 2>&1; echo "rc=$?"
done
