cd /root/repo
python bench.py --steps 3 --warmup 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"polyexp_fast|pyr3" -s 12 -c 4 -o gpurun_out/poly_full -f python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_poly.log 2>&1
ls -la gpurun_out/*.ncu-rep
