set -x
python bench.py > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err || echo BENCH_FAILED
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/k1_r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu --no-fit > gpurun_out/k1_r02_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/fit_r02_launches.csv python tools/fit_once.py 9 gls > gpurun_out/fit_r02_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:parity_kernel -s 4 -c 1 -f -o gpurun_out/k1_r02_full python bench.py --steps 5 --warmup 3 --no-cpu --no-fit > gpurun_out/k1_r02_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:potrf_dag -c 1 -f -o gpurun_out/potrf_r02_full python tools/fit_once.py 9 gls > gpurun_out/potrf_r02_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sytrd_panel -s 10 -c 1 -f -o gpurun_out/sytrd_r02_full python tools/eig_probe.py 6779 > gpurun_out/sytrd_r02_full.log 2>&1
ls -la gpurun_out | tail -20
