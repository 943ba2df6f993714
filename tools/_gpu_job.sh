set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/r02a_pytest.log
IK_REFERENCE_ROOT=.refscratch python tools/run_reference_suite.py --log gpurun_out/r02a_dropin_suite.log > gpurun_out/r02a_dropin_stdout.log 2>&1; echo "suite rc=$?"
tail -5 gpurun_out/r02a_dropin_suite.log
python tools/ann_accuracy.py models/roboarm_b200_r01 --rows 200000 --comp-sweep 0,0.5,1.5,2,3 > gpurun_out/r02a_acc_corrfirst.json 2> gpurun_out/r02a_acc_corrfirst.err; echo "acc rc=$?"
IKB200_LIB=gpurun_variants/ts_interleaved.so python tools/ann_accuracy.py models/roboarm_b200_r01 --rows 200000 --modes fp16x3_ts > gpurun_out/r02a_acc_interleaved.json 2> gpurun_out/r02a_acc_interleaved.err
python tools/time_kernels.py > gpurun_out/r02a_time_default.json 2> gpurun_out/r02a_time_default.err
IKB200_LIB=gpurun_variants/ts_interleaved.so python tools/time_kernels.py --only k2 > gpurun_out/r02a_time_interleaved.json 2>&1
cat gpurun_out/r02a_time_default.json gpurun_out/r02a_time_interleaved.json
