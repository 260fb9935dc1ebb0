# ncu evidence for profiles/: launch list with a few metrics, then one full capture each of the three kernels
set -e
cd /root/repo
TAG=${1:-r1g}
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 240 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/b_$TAG.log 2>&1 || true
for K in tileKernel geomKernel stepKernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 1 -o gpurun_out/prof_${K}_$TAG -f python bench.py --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/b2_$TAG.log 2>&1 || true
done
tail -1 gpurun_out/b_$TAG.log | cut -c1-200
