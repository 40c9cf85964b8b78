mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench15.json 2> gpurun_out/bench15.err; tail -c 400 gpurun_out/bench15.json
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench15_ref.json 2> gpurun_out/bench15_ref.err; tail -c 300 gpurun_out/bench15_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
python tools/ncu_target.py v8n 32 list 2>&1 | tail -1 > gpurun_out/ncu_order.txt
for idx in 44 1; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((62+idx)) -c 1 -f -o gpurun_out/ncu15_v8n_tc$idx python tools/ncu_target.py v8n 32 > gpurun_out/ncu15_log_$idx.txt 2>&1
done
