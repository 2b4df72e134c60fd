#!/bin/bash
# session 8: record of the committed state: whole GPU suite, default bench, launch list + DRAM traffic under ncu
O=gpurun_out/s8; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -rs > $O/gpu_tests.log 2>&1; echo "tests rc=$?"
tail -8 $O/gpu_tests.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -c 1200 $O/bench.json; tail -3 $O/bench.err
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/traffic_R256.csv python tools/prof_run.py --replicas 256 --days 180 > $O/traffic.log 2>&1
python tools/ncu_traffic.py $O/traffic_R256.csv 256 180 $O/r02_dram_traffic_R256.json "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none python tools/prof_run.py --replicas 256 --days 180" > $O/traffic_summary.log 2>&1
tail -30 $O/traffic_summary.log
gzip -f $O/traffic_R256.csv
