# One gpurun call: tests, smoke, default bench, reference arm, ncu launch lists and full captures -> gpurun_out/final/
set -x
O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/smoke.txt 2>&1
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
cp gpurun_out/bench_detail_n1.json $O/bench_detail_n1.json 2>/dev/null
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
for w in fwd train fps; do
  AMP_BENCH_EAGER_TRAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/l_$w.csv python bench.py --workload $w --only --no-cpu --steps 2 --warmup 1 > $O/ncu_$w.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_chain32_kernel -s 10 -c 5 -o $O/prof_tc_chain32 python bench.py --workload fwd --only --no-cpu --steps 1 --warmup 1 > $O/ncu_chain32.log 2>&1
AMP_BENCH_EAGER_TRAIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_wgrad_kernel -s 15 -c 1 -o $O/prof_tc_wgrad python bench.py --workload train --only --no-cpu --steps 1 --warmup 1 > $O/ncu_wgrad.log 2>&1
AMP_BENCH_EAGER_TRAIN=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:tc_layer_kernelILi16E -s 10 -c 1 -o $O/prof_tc_layer_dgrad python bench.py --workload train --only --no-cpu --steps 1 --warmup 1 > $O/ncu_tl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_assign -c 1 -o $O/prof_kmeans python bench.py --workload kmeans --only --no-cpu --steps 2 --warmup 1 > $O/ncu_km.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_window_fast -c 1 -o $O/prof_kmeans_window python bench.py --workload tile --only --no-cpu --steps 1 --warmup 1 > $O/ncu_kw.log 2>&1
tail -2 $O/pytest_gpu.txt; tail -1 $O/smoke.txt; wc -c $O/*.json; ls -la $O
