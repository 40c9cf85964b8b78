#!/bin/bash
# dram bytes of the 59 conv_tc_kernel launches of one eager YOLOv8n batch-32 forward (roofline.traffic of the default bench line)
mkdir -p gpurun_out/rtr; O=gpurun_out/rtr
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 59 -c 59 --csv --log-file $O/conv_traffic.csv python tools/ncu_target.py v8n 32 > /dev/null 2>&1
python tools/ncu_traffic.py $O/conv_traffic.csv v8n 32 > $O/r2_conv_traffic.json; cat $O/r2_conv_traffic.json
