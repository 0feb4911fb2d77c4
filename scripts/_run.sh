timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
tail -c 400 gpurun_out/r3i_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
