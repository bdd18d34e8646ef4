# 8-GPU check of the final code: the driver's launch line (all ranks, extras included) and the reference arm on the same box
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r2_final_ref_n$N.json 2> gpurun_out/r2_final_ref_n$N.err
