# bench lines at 8 and 4 GPUs (run with gpurun --gpus 8)
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_n8.json 2> gpurun_out/r3e_bench_n8.err
tail -c 300 gpurun_out/r3e_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r3e_bench_n4.json 2> gpurun_out/r3e_bench_n4.err
tail -c 300 gpurun_out/r3e_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/multi_gpu_check.py > gpurun_out/r3e_multi_gpu_check_n8.log 2>&1
tail -n 3 gpurun_out/r3e_multi_gpu_check_n8.log
nproc
