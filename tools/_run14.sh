python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-modes > gpurun_out/bench_n8_s3.json 2> gpurun_out/bench_n8_s3.err
tail -c 400 gpurun_out/bench_n8_s3.json; tail -3 gpurun_out/bench_n8_s3.err
