set -x
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench256.csv python bench.py --images 256 --chunks 1 --steps 2 --warmup 3 --no-configs > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"unstuff|spec_kernel|fix_local|chain_kernel|write_kernel|bj_pixels" -s 8 -c 8 -f -o gpurun_out/prof_r2 python tools/profile_run.py --images 512 --distinct 64 --steps 2 > gpurun_out/prof_r2.log 2>&1
ls -la gpurun_out/prof_r2.ncu-rep
