set -x
mkdir -p gpurun_out/final
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final/smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err
timeout 600 python bench.py --impl reference > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
for w in fwd fwd_bf16 train; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/l_$w.csv python bench.py --workload $w --only --no-cpu --steps 2 --warmup 1 > gpurun_out/final/ncu_$w.log 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:tc_layer_kernelILi1E -s 9 -c 1 -o gpurun_out/final/prof_tc_layer_m1 python bench.py --workload train --only --no-cpu --steps 1 --warmup 1 > gpurun_out/final/ncu_tl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -c 1 -o gpurun_out/final/prof_kmeans python bench.py --workload kmeans --only --no-cpu --steps 2 --warmup 1 > gpurun_out/final/ncu_km.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_chain_kernel -s 8 -c 4 -o gpurun_out/final/prof_tc_chain python bench.py --workload fwd_bf16 --only --no-cpu --steps 1 --warmup 1 > gpurun_out/final/ncu_chain.log 2>&1
tail -2 gpurun_out/final/pytest_gpu.txt; tail -1 gpurun_out/final/smoke.txt; wc -c gpurun_out/final/*.json
