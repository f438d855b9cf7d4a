timeout 300 python bench.py > gpurun_out/r1_bf16_bench.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/r1_bf16_bench.json
for w in cfg3 cfg4 cfg5; do timeout 400 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/r1_${w}_bench.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/r1_${w}_bench.json; done
timeout 300 python bench.py --net fp32 --steps 50 --warmup 3 > gpurun_out/r1_fp32_bench.json 2>> gpurun_out/bench.err; cut -c1-160 gpurun_out/r1_fp32_bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/r1_bf16_launches.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bf16_chain_pipe -s 60 -c 2 -o gpurun_out/pipe_full -f python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_f.log 2>&1
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_backup_select_sm -s 60 -c 2 -o gpurun_out/tree_full -f python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu_t.log 2>&1
ls -la gpurun_out | tail -12
