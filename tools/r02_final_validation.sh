timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final2.log 2>&1; echo pytest_rc=$?; tail -3 gpurun_out/pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for spec in "res101 1" "res101 4" "vgg16 4" "mobile 4" "res152lg 4"; do set -- $spec; python bench.py --steps 40 --warmup 5 --net $1 --batch $2 > gpurun_out/final_$1_b$2.json 2> gpurun_out/final_$1_b$2.err; python tools/print_bench_line.py gpurun_out/final_$1_b$2.json || tail -5 gpurun_out/final_$1_b$2.err; done
python bench.py --layers --batch 4 > gpurun_out/final_layers_res101_b4.txt 2>&1; python bench.py --layers --batch 1 > gpurun_out/final_layers_res101_b1.txt 2>&1; head -15 gpurun_out/final_layers_res101_b1.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
for b in 1 4; do ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_res101_b$b.csv python bench.py --ncu --batch $b > /dev/null 2>&1; done
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -s 95 -c 4 -o /tmp/convb4 python bench.py --ncu --batch 4 > /dev/null 2>&1; ncu -i /tmp/convb4.ncu-rep --page raw --csv > gpurun_out/final_ncu_full_conv_block4_b4_raw.csv 2>/dev/null
du -sh gpurun_out
