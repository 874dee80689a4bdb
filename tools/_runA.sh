( time python -m pytest tests -x -q -m gpu ) > gpurun_out/tA.log 2>&1; tail -6 gpurun_out/tA.log
( time python bench.py ) > gpurun_out/benchA.json 2> gpurun_out/benchA.err; tail -3 gpurun_out/benchA.err; cat gpurun_out/benchA.json | head -c 6000
