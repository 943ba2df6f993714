mkdir -p gpurun_out
for v in default k3_minb3 k3_minb5 k3_minb6; do
  if [ $v = default ]; then python tools/time_kernels.py --only k1w,k3 --reps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['k3'])";
  else IKB200_LIB=gpurun_variants/$v.so python tools/time_kernels.py --only k1w,k3 --reps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['k3'])"; fi
done 2>&1 | tee gpurun_out/r02k_k3_variants.log
