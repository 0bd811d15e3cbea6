python -m pytest tests -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3j_tests.log; tail -4 gpurun_out/r3j_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3j_smoke.log 2>&1; tail -2 gpurun_out/r3j_smoke.log
python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; tail -2 gpurun_out/r3j_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3j_bench_ref.json 2> gpurun_out/r3j_bench_ref.err
timeout 600 python tools/fuzz_parity.py 1500 424242 > gpurun_out/r3j_fuzz.log 2>&1; tail -2 gpurun_out/r3j_fuzz.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3j_launches.csv python bench.py --steps 2 --warmup 1 --short --legs none --no-cpu-baseline > gpurun_out/r3j_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wiggle_kernel -s 1 -c 1 -o gpurun_out/r3j_wiggle -f python bench.py --steps 1 --warmup 1 --short --legs c5 --no-cpu-baseline > gpurun_out/r3j_ncu_wig.log 2>&1; tail -2 gpurun_out/r3j_ncu_wig.log
