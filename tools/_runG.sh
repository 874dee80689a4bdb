for w in 8 2; do
LZB_ENC_WARPS=$w ncu --set full --clock-control none --import-source on -k regex:lzb_parse_kernel -c 1 -f -o gpurun_out/parse_w$w python tools/quick_encode_bench.py 1184 131072 4 0 64 1048576 > gpurun_out/ncu_parse_w$w.log 2>&1
tail -2 gpurun_out/ncu_parse_w$w.log
done
ls -la gpurun_out/*.ncu-rep
