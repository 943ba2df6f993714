mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02m_bench_n8.json 2> gpurun_out/r02m_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r02m_bench_n8.err
nvidia-smi topo -m > gpurun_out/r02m_topo.txt 2>&1; lscpu | head -25 > gpurun_out/r02m_lscpu.txt; numactl -H >> gpurun_out/r02m_lscpu.txt 2>&1
