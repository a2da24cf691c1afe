set -x
O=gpurun_out/r2n2; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -c 2500 $O/bench_n2.json; grep -i "nranks\|error\|Traceback" $O/bench_n2.err | head
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
