set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2N_bench_n8.json 2> gpurun_out/r2N_bench_n8.err
tail -c 1500 gpurun_out/r2N_bench_n8.err
QRAD_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/r2N_dbg_n8.json 2> gpurun_out/r2N_dbg_n8.err
grep -c "" gpurun_out/r2N_dbg_n8.err
