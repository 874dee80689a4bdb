python tools/probe_classes.py "4096,1048576,4,64,1048576,0" > gpurun_out/probe4.log 2>&1
grep -E "PROBE|pipeline|SPEC|rror|group" gpurun_out/probe4.log | tail -40
