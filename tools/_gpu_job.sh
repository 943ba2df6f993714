set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02g_pytest.log
python tools/time_kernels.py --only k1w,k1r,k3 > gpurun_out/r02g_time.json 2> gpurun_out/r02g_time.err; cat gpurun_out/r02g_time.json
