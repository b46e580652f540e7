# final multi-GPU bench lines: bash tools/final_multi.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_final_n$N.json 2> gpurun_out/r02_final_n$N.err
tail -c 400 gpurun_out/r02_final_n$N.json
