for rep in 1 2; do
for lib in old new; do
if [ $lib = old ]; then export PSBAX_LIB=/root/repo/psbax_b200/libpsbax_old.so; else unset PSBAX_LIB; fi
python scripts/config_bench.py auto 2>&1 | grep "C3 GB1 top-k d=32\|C5 Ackley top-32\|C1 demo top-k (30" | sed "s/^/$lib: /"
done; done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line 2>&1 | tail -2
