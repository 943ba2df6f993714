mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02l_bench_n2.err
python -m pytest tests/test_gpu_sharded.py -q > gpurun_out/r02l_pytest_sharded.log 2>&1; tail -3 gpurun_out/r02l_pytest_sharded.log
