#!/bin/bash
# DRAM traffic of every kernel of every BASELINE workload, measured on the current kernels (run on the GPU box):
#   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, once with ncu's default cache control
#   (caches flushed before every kernel: cold) and once with --cache-control none (what a kernel finds in L2 from its
#   predecessor stays there).  Writes gpurun_out/traffic_<workload>_<streams>_<mode>.csv; tools/ncu_traffic.py turns them into
#   profiles/r02_dram_bytes.json, which bench.py reads for roofline.traffic.
set -u
FR=${FRAMES:-8}
for W in "heaac_v1_stereo 1024" "heaac_v1_stereo 8192" "heaac_v2_ps 4096" "lc_51_48k_tns 2048" "lc_stereo_44k 4096"; do
  set -- $W
  for MODE in all none; do
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control $MODE \
        --csv --log-file gpurun_out/traffic_$1_$2_$MODE.csv \
        python bench.py --workload $1 --streams $2 --frames $FR --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-configs \
        > gpurun_out/traffic_$1_$2_$MODE.log 2>&1
  done
done
ls -la gpurun_out/traffic_*.csv
