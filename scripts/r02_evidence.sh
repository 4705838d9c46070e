#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): tests, the four bench lines, the launch list of the
# default bench command and one `ncu --set full` capture of the top kernel of each workload.
set -u
mkdir -p gpurun_out/r02f
python -m pytest tests -m gpu -x -q > gpurun_out/r02f/pytest_gpu.log 2>&1; tail -2 gpurun_out/r02f/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f/smoke.log 2>&1; tail -1 gpurun_out/r02f/smoke.log
for w in c2 c3 c4 c5; do
  python bench.py --workload $w > gpurun_out/r02f/bench_$w.json 2> gpurun_out/r02f/bench_$w.err
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f/bench_reference.json 2> gpurun_out/r02f/bench_reference.err
python scripts/config_bench.py auto > gpurun_out/r02f/config_table.txt 2>&1
# launch list of the default bench command (kernel shares)
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f/launches_c2.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/ncu_launches_c2.log 2>&1
# one full capture per workload's top kernel
python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/plain2_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_shared_tensor -s 3 -c 1 -f -o gpurun_out/r02f/full_c2 \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/ncu_full_c2.log 2>&1
for w in c3 c5; do
  python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/plain2_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_tensor_proj<\(int\)[12]' -s 3 -c 1 -f -o gpurun_out/r02f/full_$w \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/ncu_full_$w.log 2>&1
done
python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/plain2_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_subset_select -s 3 -c 1 -f -o gpurun_out/r02f/full_c4 \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02f/ncu_full_c4.log 2>&1
ls -la gpurun_out/r02f
