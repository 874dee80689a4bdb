set -x
python -m pytest tests/test_encode_gpu.py -x -q -m gpu -k "pipelined or many_tiny or all_classes or tiny_and" > gpurun_out/t2.log 2>&1; tail -5 gpurun_out/t2.log
python tools/probe_classes.py "4096,1048576,4,64,1048576,0" "4096,1048576,4,64,1048576,0;LZB_ENC_GROUP=148;LZB_ENC_LANES=8" "4096,1048576,4,64,1048576,0;LZB_ENC_GROUP=148;LZB_ENC_LANES=12" "4096,1048576,4,64,1048576,0;LZB_ENC_GROUP=512;LZB_ENC_LANES=4" "2048,1048576,4,64,1048576,0" "2048,1048576,4,64,1048576,0;LZB_ENC_PIPE=0" "512,4194304,4,32,4194304,0;LZB_ENC_PIPE=1" > gpurun_out/probe2.log 2>&1
grep -E "PROBE|pipeline|SPEC|rror" gpurun_out/probe2.log | tail -40
