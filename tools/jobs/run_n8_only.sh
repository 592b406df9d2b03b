# bench line at 8 GPUs (run with gpurun --gpus 8)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3i_bench_n8.json 2> gpurun_out/r3i_bench_n8.err
tail -c 300 gpurun_out/r3i_bench_n8.err
