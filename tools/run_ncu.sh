set -x
CMD="python bench.py --steps 2 --warmup 3 --skip-extras --skip-cpu"
$CMD > gpurun_out/plain_r2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_l_r2.log 2>&1
tail -2 gpurun_out/ncu_l_r2.log
ncu --set full --clock-control none --import-source on -k regex:k_whole -s 6 -c 2 -o gpurun_out/prof_whole_r2 -f $CMD > gpurun_out/ncu_f_r2.log 2>&1
tail -3 gpurun_out/ncu_f_r2.log; ls -la gpurun_out/*.ncu-rep | tail -3
