set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_plain.json 2> gpurun_out/r2b_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dopri5_fwd_kernel|dopri5_backprop_bwd_kernel" -s 6 -c 4 -f -o gpurun_out/r2b_dopri5 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/ncu2.log 2>&1
ncu -i gpurun_out/r2b_dopri5.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_full_dopri5_kernels.csv 2>/dev/null
python scripts/wide_dp5_probe.py > gpurun_out/wide_probe_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"wide_dopri5|wide_dp5" -s 40 -c 6 -f -o gpurun_out/r2b_wide_dp5 python scripts/wide_dp5_probe.py > gpurun_out/ncu3.log 2>&1
ncu -i gpurun_out/r2b_wide_dp5.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_full_wide_dopri5.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"dopri5_fwd_kernel" -s 3 -c 1 -f -o gpurun_out/r2b_odernn python scripts/odernn_probe.py 8192 > gpurun_out/ncu4.log 2>&1
ncu -i gpurun_out/r2b_odernn.ncu-rep --page raw --csv > gpurun_out/r2b_ncu_full_odernn_persistent_fwd.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/r2b_*.csv | tail -12
