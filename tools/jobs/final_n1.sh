python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/r2Z_bench.json 2> gpurun_out/r2Z_bench.err; tail -c 200 gpurun_out/r2Z_bench.err; cut -c1-200 gpurun_out/r2Z_bench.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2Z_bench_reference.json 2> gpurun_out/r2Z_ref.err; cut -c1-200 gpurun_out/r2Z_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2Z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2Z_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_row_totals32|k_columns|k_tri_points|k_final" -c 4 -o gpurun_out/r2Z_passes python tools/profile_transfers.py c5_chop16 > gpurun_out/r2Z_ncu_passes.log 2>&1; tail -n 1 gpurun_out/r2Z_ncu_passes.log | cut -c1-200
ls -la gpurun_out/r2Z*
