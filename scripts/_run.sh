timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err
tail -c 200 gpurun_out/r3m_bench.err
