# bench line at 4 GPUs (run with gpurun --gpus 4)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3j_bench_n4.json 2> gpurun_out/r3j_bench_n4.err
tail -c 300 gpurun_out/r3j_bench_n4.err
