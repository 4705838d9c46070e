for rep in 1 2; do
PSBAX_LIB=/root/repo/psbax_b200/libpsbax_old.so python scripts/config_bench.py auto 2>&1 | grep "C2 level-set 1.68e7\|C4 DiscoBAX values" | sed "s/^/old: /"
python scripts/config_bench.py auto 2>&1 | grep "C2 level-set 1.68e7\|C4 DiscoBAX values" | sed "s/^/new: /"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line 2>&1 | tail -2
