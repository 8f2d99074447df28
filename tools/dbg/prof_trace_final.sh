export ZYG_BENCH_CACHE=/tmp/zyg_cache
ncu --set full --import-source on --clock-control none -k regex:traceWidePersistent -s 12 -c 3 -o gpurun_out/s6_trace_final python bench.py --steps 1 --warmup 3 --no-cpu --no-render > /dev/null 2>&1
ls -la gpurun_out/s6_trace_final.ncu-rep
