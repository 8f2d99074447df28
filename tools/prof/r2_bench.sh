# the bench line of the final state (+ reference arm); usage: TAG=.. bash tools/prof/r2_bench.sh
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; tail -n 2 $O/${TAG}_bench.log; head -c 300 $O/${TAG}_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.log; head -c 200 $O/${TAG}_bench_ref.json; echo
