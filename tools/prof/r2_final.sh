# Evidence of the round's final state (run under gpurun, one GPU): bench line, reference arm, launch list of the same bench command,
# --set full captures of the dominant kernels, DRAM bytes per timed traversal launch. usage: TAG=r2_xx bash tools/prof/r2_final.sh
O=gpurun_out
python __graft_entry__.py > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.log; tail -2 $O/${TAG}_bench.log; head -c 400 $O/${TAG}_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.log; head -c 300 $O/${TAG}_bench_ref.json; echo
export ZYG_BENCH_CACHE=/tmp/zyg_cache
# launch list of the bench command (traversal microbench + the five render scenes, 2 steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2> $O/${TAG}_launches_bench.log
# DRAM bytes of the three timed traversal launches of one step (roofline.traffic)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'traceWide|tracePool|traceRay' -c 40 --csv --log-file $O/${TAG}_traffic_trace.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-render > /dev/null 2>&1
# --set full: the traversal microbench kernels, the fused scene kernel on config 3
ncu --set full --import-source on --clock-control none -k regex:'traceWide|tracePool|traceRay' -s 12 -c 3 -o $O/${TAG}_trace python bench.py --steps 1 --warmup 3 --no-cpu --no-render > /dev/null 2>&1
SCENE=instanced_scene KW='{"grid":[100,100],"prototypes":20,"quads":[500,250],"sun":60.0}' ncu --set full --import-source on --clock-control none -k regex:sceneTracePersistent -s 2 -c 2 -o $O/${TAG}_config3_trace python tools/render_scene.py 1920 1080 2 1 > /dev/null 2>&1
ls -la $O/${TAG}_*
