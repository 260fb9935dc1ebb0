set -e
cd /root/repo
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 240 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/b_r1f.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:tileKernel -s 60 -c 1 -o gpurun_out/prof_tile_r1f -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/b2.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:geomKernel -s 60 -c 1 -o gpurun_out/prof_geom_r1f -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/b3.log 2>&1 || true
tail -2 gpurun_out/b_r1f.log | cut -c1-300
