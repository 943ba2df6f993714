set -x
mkdir -p gpurun_out
python tools/time_kernels.py --only k1r --reps 1 --warm 0 > gpurun_out/r02b_k1r_plain.json 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:fabrik_split -c 1 -f -o gpurun_out/r02b_k1_interior python tools/time_kernels.py --only k1r --reps 1 --warm 0 > gpurun_out/r02b_ncu_k1r.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r02b_pytest.log
python bench.py --steps 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02b_bench.err
