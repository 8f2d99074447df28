python -m pytest tests -m gpu -x -q > gpurun_out/s6_tests_final.log 2>&1; tail -3 gpurun_out/s6_tests_final.log
python __graft_entry__.py smoke > gpurun_out/s6_smoke.log 2>&1; tail -2 gpurun_out/s6_smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/s6_bench_final.json 2> gpurun_out/s6_bench_final.log; tail -2 gpurun_out/s6_bench_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s6_bench_ref.json 2> gpurun_out/s6_bench_ref.log; tail -2 gpurun_out/s6_bench_ref.log; head -c 600 gpurun_out/s6_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s6_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --scenes config1_cornell_512x512x64 sphere1m_1024x1024x16 > /dev/null 2> gpurun_out/s6_launches_bench.log
