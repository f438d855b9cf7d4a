#!/bin/bash
# Round-2 evidence run (one B200): bench lines of every workload / mode, the ncu launch list of the bench command,
# and one `--set full` capture per hot kernel.  Everything lands in gpurun_out/; summaries are copied to profiles/.
set -u
O=gpurun_out
mkdir -p $O
t() { timeout "$@"; }
t 400 python bench.py --steps 20 --warmup 5 > $O/r2_bench_cfg2.json 2> $O/r2_bench_cfg2.err
for w in cfg3 cfg4 cfg5; do
  t 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_bench_$w.json 2> $O/r2_bench_$w.err
done
for m in tc32 f16; do
  t 300 python bench.py --net $m --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/r2_bench_$m.json 2> $O/r2_bench_$m.err
done
t 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_reference_arm.json 2>/dev/null
# launch list of the bench command (2 timed steps): per-launch device time, serialised, no PDL overlap
t 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_bf16_launches.csv \
  python bench.py --steps 2 --warmup 3 --profile-only > $O/r2_ncu_launch.log 2>&1
# full captures: bf16 chain, tc32 chain (3 products / 1 product), the shared-memory tree step
t 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bf16_chain_m32 -s 60 -c 2 -f -o $O/r2_m32_full \
  python bench.py --steps 1 --warmup 3 --profile-only > $O/r2_ncu_a.log 2>&1
t 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_tc32_chain_m64 -s 60 -c 2 -f -o $O/r2_tc32_full \
  python bench.py --net tc32 --steps 1 --warmup 3 --profile-only > $O/r2_ncu_b.log 2>&1
t 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_tc32_chain_m64 -s 60 -c 2 -f -o $O/r2_f16_full \
  python bench.py --net f16 --steps 1 --warmup 3 --profile-only > $O/r2_ncu_c.log 2>&1
t 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_backup_select_sm -s 60 -c 2 -f -o $O/r2_tree_full \
  python bench.py --steps 1 --warmup 3 --profile-only > $O/r2_ncu_d.log 2>&1
t 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_bf16_chain_pipeN -s 20 -c 2 -f -o $O/r2_pipe2_full \
  python bench.py --workload cfg4 --steps 1 --warmup 2 --profile-only > $O/r2_ncu_f.log 2>&1
t 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_cfg4_launches.csv \
  python bench.py --workload cfg4 --steps 1 --warmup 2 --profile-only > $O/r2_ncu_g.log 2>&1
t 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_cfg5_launches.csv \
  python bench.py --workload cfg5 --steps 1 --warmup 2 --profile-only > $O/r2_ncu_e.log 2>&1
SMZ_BF16_TIMELINE=1 SMZ_TREE_TIMELINE=1 t 120 python tools/diag_timeline.py > $O/r2_timeline.log 2>&1
ls -la $O | tail -30
