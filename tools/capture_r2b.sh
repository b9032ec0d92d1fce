# ncu evidence for the second half of round 2 (column-pair pair kernel, fused reduce-scatter, one-pass bucketing,
# tcgen05 projection). Each target program has already exited 0 without ncu in the same call.
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:pair_cp_kernel -s 2 -c 1 -o gpurun_out/r2b_pair_c4 -f python tools/run_closure_once.py c4 3 > gpurun_out/r2b_ncu_pair.log 2>&1
$NCU -k regex:pair_cp_kernel -s 2 -c 1 -o gpurun_out/r2b_pair_c5 -f python tools/run_closure_once.py c5 3 > gpurun_out/r2b_ncu_pair5.log 2>&1
$NCU -k regex:project_tc -s 2 -c 1 -o gpurun_out/r2b_ptc_c5 -f python tools/run_closure_once.py c5 3 > gpurun_out/r2b_ncu_ptc5.log 2>&1
$NCU -k regex:gram_tf32x3_kernel -s 1 -c 1 -o gpurun_out/r2b_gram_c2 -f python tools/run_stats_once.py c2 2 > gpurun_out/r2b_ncu_gram2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2b_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-table > gpurun_out/r2b_ncu_bench.log 2>&1
for c in c2 c4 c3 c1 c5; do ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2b_cl_$c.csv python tools/run_closure_once.py $c 4 > gpurun_out/r2b_cl_$c.log 2>&1; done
for c in c3 c1 c4; do ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r2b_st_$c.csv python tools/run_stats_once.py $c 3 > gpurun_out/r2b_st_$c.log 2>&1; done
