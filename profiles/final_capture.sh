#!/bin/bash
# Round-end captures on the GPU box: python runs first WITHOUT ncu (must exit 0), then one `ncu --set full` launch per kernel,
# exported as text right here (the library on the box is the build that was profiled, so attribute.py lines up).
#   gpurun -- 'bash profiles/final_capture.sh r2_final'
set -u
TAG=${1:-r2_final}
OUT=gpurun_out
python profiles/profile_step.py mass_td3 1048576 64 > $OUT/${TAG}_plain_run.log 2>&1 || { echo "plain run failed"; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:step_kernel -s 60 -c 1 -o $OUT/${TAG}_step python profiles/profile_step.py mass_td3 1048576 64 > /dev/null 2>&1
$NCU -k regex:outputs_kernel -s 121 -c 1 -o $OUT/${TAG}_outputs python profiles/profile_step.py mass_td3 1048576 64 > /dev/null 2>&1
$NCU -k regex:coop_step_kernel -s 60 -c 1 -o $OUT/${TAG}_coop python profiles/profile_step.py hss_td3 4096 64 > /dev/null 2>&1
$NCU -k regex:actor_mlp -s 4 -c 1 -o $OUT/${TAG}_actor python profiles/profile_actor.py > /dev/null 2>&1
LIB=marl-mass_b200/_build/libmarl_mass_b200.so
python profiles/ncu_summary.py $OUT/${TAG}_step.ncu-rep $OUT/${TAG}_outputs.ncu-rep $OUT/${TAG}_coop.ncu-rep $OUT/${TAG}_actor.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
python profiles/attribute.py $LIB $OUT/${TAG}_step.ncu-rep mms_mass11step_kernelILb0ELb0 40 > $OUT/${TAG}_step_by_function.txt 2>&1
python profiles/attribute.py $LIB $OUT/${TAG}_outputs.ncu-rep outputs_kernelILb1 30 > $OUT/${TAG}_outputs_by_function.txt 2>&1
python profiles/attribute.py $LIB $OUT/${TAG}_coop.ncu-rep coop_step_kernelILi1ELb0 30 > $OUT/${TAG}_coop_by_function.txt 2>&1
python profiles/traffic_from_ncu.py mass_td3 1048576 $OUT/${TAG}_traffic.json $OUT/${TAG}_step.ncu-rep $OUT/${TAG}_outputs.ncu-rep > /dev/null 2>&1
# launch list of the bench command itself (short run; the numbers printed under ncu are not bench values)
python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extra > $OUT/${TAG}_bench_plain.json 2> /dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extra > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/${TAG}_bench_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 2:]:
    if len(r) > vi:
        agg[r[ki][:70]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open("$OUT/${TAG}_bench_launch_shares.txt", "w") as f:
    f.write("launch list of: python bench.py --steps 3 --warmup 3 --skip-cpu --skip-extra (ncu --metrics gpu__time_duration.sum; whole process incl. the 100-step prologue, the QP microbenchmark and the host-path passes)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("%-70s n=%5d mean %9.1f us share %5.1f%%\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
print(open("$OUT/${TAG}_bench_launch_shares.txt").read())
PY
cat $OUT/${TAG}_ncu_summary.txt | head -60
