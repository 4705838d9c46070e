set -u
mkdir -p gpurun_out/r02f
python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/plain3_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tensor_proj<\(int\)2' -s 3 -c 1 -f -o gpurun_out/r02f/full_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/ncu_full_c3.log 2>&1
tail -1 gpurun_out/r02f/ncu_full_c3.log
python scripts/one_kernel.py generic &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_shared_tensor<\(int\)0, \(int\)2' -s 1 -c 1 -f -o gpurun_out/r02f/full_generic python scripts/one_kernel.py generic > gpurun_out/r02f/ncu_generic.log 2>&1
tail -1 gpurun_out/r02f/ncu_generic.log
