cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/bench_line.json
python tools/summ.py gpurun_out/bench_line.json
python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_l.log 2>&1
echo launches done
ncu --set full --clock-control none --import-source on -k regex:"flow_strip|flow_layer_kernel" -s 9 -c 3 -o gpurun_out/flow_full -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"polyexp_fast|pyr3|classify_batch|thresholds_batch" -s 15 -c 6 -o gpurun_out/others_full -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_o.log 2>&1
ls gpurun_out/ | wc -l
