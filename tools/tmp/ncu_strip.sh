cd /root/repo
python bench.py --steps 2 --warmup 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"flow_strip" -s 6 -c 1 -o gpurun_out/strip_full -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_s.log 2>&1
ls -la gpurun_out/strip_full.ncu-rep
