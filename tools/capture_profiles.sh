#!/bin/bash
# Round-2 profile capture (run under gpurun): one ncu --set full capture per dominant kernel at HEAD, each after the same
# command has exited 0 without ncu, plus the launch list of the bench command.  Reports land in gpurun_out/;
# tools/ncu_summary.py turns them into profiles/r2_*.txt (with the SASS opcode histogram of the captured kernel).
set -u
O=gpurun_out
P="ncu --set full --clock-control none --import-source on"
run() { # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  "$@" > $O/r2_${name}_plain.log 2>&1 && $P -k regex:$rx -s $skip -c $cnt -o $O/r2_prof_${name} "$@" > $O/r2_${name}_ncu.log 2>&1
  echo "$name rc=$?"
}
# two groups, one gpurun call each: a call may bring back at most 64 MiB and a report is ~10 MB
G=${1:-all}
if [ "$G" = "a" ] || [ "$G" = "all" ]; then
run lanes_4096 gx_step_lanes_kernel 20 2 python tools/prof_target.py 4096 30
run step_65536 gx_step_kernel 6 2 python tools/prof_target.py 65536 12
run step_1m gx_step_kernel 3 2 python tools/prof_target.py 1048576 6
run policy_rollout_4096 gx_policy_rollout_kernel 1 1 python tools/prof_policy.py 4096 20
fi
if [ "$G" = "b" ] || [ "$G" = "all" ]; then
run art_ant_65536 gx_art_kernel 31 1 python tools/prof_target.py 65536 3 0 xmls/ant.xml
run tree_doggo_4096 gx_tree_kernel 31 1 python tools/prof_target.py 4096 3 0 xmls/doggo.xml
run policy_4096 gx_policy_kernel 25 2 python tools/prof_policy.py 4096 20 graph
run task_button_car_16384 gxt_step_lanes_kernel 25 1 python tools/prof_task.py Button_Car 16384 30
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 20 --warmup 5 --no-sweep --no-legged --no-ppo > $O/r2_bench_plain.json 2> $O/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 20 --warmup 5 --no-sweep --no-legged --no-ppo > $O/r2_launches_ncu.log 2>&1
echo "launches rc=$?"
fi
