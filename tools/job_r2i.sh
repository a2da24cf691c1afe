set -x
O=gpurun_out/r2i; mkdir -p $O
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 1500 $O/bench_n2.err; cut -c1-1200 $O/bench_n2.json
MPRB_GPUS=2 ./build/drivers/render_3d_table 2>&1 | tail -6
MPRB_GPUS=1 ./build/drivers/render_3d_table 2>&1 | tail -6
