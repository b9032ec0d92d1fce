# ncu evidence for round 2 (each target program has already exited 0 without ncu in the same call)
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:pair_ai_reg -s 2 -c 1 -o gpurun_out/r2_pair_c4 -f python tools/run_closure_once.py c4 3 > gpurun_out/r2_ncu_pair.log 2>&1
$NCU -k regex:project_stream -s 2 -c 1 -o gpurun_out/r2_proj_c2 -f python tools/run_closure_once.py c2 3 > gpurun_out/r2_ncu_proj.log 2>&1
$NCU -k regex:gram_tf32x3_small -s 1 -c 1 -o gpurun_out/r2_gram_c3 -f python tools/run_stats_once.py c3 2 > gpurun_out/r2_ncu_gram3.log 2>&1
$NCU -k regex:gram_tf32x3_kernel -s 1 -c 1 -o gpurun_out/r2_gram_c2 -f python tools/run_stats_once.py c2 2 > gpurun_out/r2_ncu_gram2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-table > gpurun_out/r2_ncu_bench.log 2>&1
for c in c2 c4 c3 c1 c5; do ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2_cl_$c.csv python tools/run_closure_once.py $c 4 > gpurun_out/r2_cl_$c.log 2>&1; done
for c in c3 c1 c4; do ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2_st_$c.csv python tools/run_stats_once.py $c 3 > gpurun_out/r2_st_$c.log 2>&1; done
