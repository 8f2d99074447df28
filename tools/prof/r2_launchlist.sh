# launch list of the bench command in the final state + DRAM bytes of the timed traversal launches; usage: TAG=.. bash tools/prof/r2_launchlist.sh
O=gpurun_out
export ZYG_BENCH_CACHE=/tmp/zyg_cache
python bench.py --steps 1 --warmup 1 --no-cpu --no-render > /dev/null 2>&1   # fills the ray cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2> $O/${TAG}_launches_bench.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'traceWide' -c 40 --csv --log-file $O/${TAG}_traffic_trace.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-render > /dev/null 2>&1
ls -la $O/${TAG}_*
