# final one-GPU job: GPU tests, smoke, bench line, reference arm, launch list
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; tail -c 200 gpurun_out/r3k_bench.err; cut -c1-200 gpurun_out/r3k_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3k_bench_reference.json 2> gpurun_out/r3k_ref.err; cut -c1-200 gpurun_out/r3k_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3k_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3k_ncu_launches.log 2>&1
