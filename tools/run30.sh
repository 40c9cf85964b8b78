mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:nms_kernel -c 1 -f -o gpurun_out/nms30 python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ls -la gpurun_out/nms30.ncu-rep
